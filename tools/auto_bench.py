"""Timing of the section-8f entry points on one GPU: get_fock, AutoRCCSD.do_rccsd to convergence
(canonical and non-canonical reference), the (T) correction and mRCCD (DIIS), at a given shape.
Phase times are CUDA-event milliseconds on the library's stream (jues_b200_get_phases).

  python tools/auto_bench.py --nbf 120 --nocc 20 --out gpurun_out/auto_bench.json
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jues.jl_b200 as jb
from jues.jl_b200 import flops


def phases(ctx):
    agg = {}
    for k, ms in ctx.phases():
        agg.setdefault(k, []).append(ms)
    return agg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nbf", type=int, default=120)
    ap.add_argument("--nocc", type=int, default=20)
    ap.add_argument("--seed", type=int, default=2024)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    N, o = args.nbf, args.nocc
    v = N - o
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "profiles", "fp64_peak_r01.json")))["fp64_tflops"]
    t0 = time.time()
    g, h, Ca, eps = jb.synth.noncanonical_inputs(N, o, seed=args.seed)
    gc, Cao, Cav, _ = jb.synth.dense_inputs(N, o, seed=args.seed)
    hc = jb.synth.core_hamiltonian(gc, Cao, Cav, eps)
    t_inputs = time.time() - t0
    ctx = jb.Context(0)
    w_can = jb.Wfn(o, v, eps, Cao, Cav, gc, hao=hc)
    w_non = jb.Wfn(o, v, eps, Ca[:, :o].copy(), Ca[:, o:].copy(), g, hao=h, Ca=Ca)
    res = {"nbf": N, "nocc": o, "nvir": v, "fp64_peak_tflops": peak, "host_input_s": round(t_inputs, 2)}

    jb.get_fock(w_can, ctx=ctx)                      # warm-up (module load, pools)
    t0 = time.time()
    f = jb.get_fock(w_non, ctx=ctx)
    ph = phases(ctx)
    res["get_fock"] = {"wall_s": time.time() - t0, "fock.build_ms": ph.get("fock.build"), "h2d.gao_ms": ph.get("h2d.gao"),
                       "gemm_flops": ctx.counters()["gemm_flops"],
                       "max_offdiag_oo": float(np.abs(f[:o, :o] - np.diag(np.diag(f[:o, :o]))).max())}

    for name, w in (("auto_rccsd_canonical", w_can), ("auto_rccsd_noncanonical", w_non)):
        t0 = time.time()
        r = jb.AutoRCCSD.do_rccsd(w, ctx=ctx, do_pT=(name == "auto_rccsd_canonical"), _return_all=True)
        wall = time.time() - t0
        ph = phases(ctx)
        c = ctx.counters()
        it_ms = ph.get("cc.iteration", [])
        gf = ph.get("cc.iteration.gflop", [])
        d = {"wall_s": wall, "iterations": r["iterations"], "converged": r["converged"], "ecc": r["ecc"],
             "ept": r["ept"], "ms_per_sweep_median": float(np.median(it_ms)) if it_ms else None,
             "sweep_exec_tflops": float(np.median(gf) / np.median(it_ms)) if it_ms else None,
             "cc.transform_ms": ph.get("cc.transform"), "fock.build_ms": ph.get("fock.build"),
             "cc.triples_ms": ph.get("cc.triples"), "total_ms": ph.get("total"),
             "launches": [c["gemm_launches"], c["aux_launches"]]}
        tr = {k: round(float(sum(vv)), 3) for k, vv in ph.items() if k.startswith(("pt.", "cc.part", "cc.comm", "tei."))}
        if tr:
            d["trace_ms_sum"] = tr          # JUES_B200_TRACE=1: fine-grained CUDA-event timers, summed
        if d["sweep_exec_tflops"]:
            d["sweep_frac_of_fp64_peak"] = d["sweep_exec_tflops"] / peak
        if ph.get("cc.triples"):
            ft = flops.pt_flops(o, v)
            d["pt_flops"] = ft
            d["pt_tflops"] = ft / (ph["cc.triples"][0] * 1e-3) * 1e-12
            d["pt_frac_of_fp64_peak"] = d["pt_tflops"] / peak
        res[name] = d

    t0 = time.time()
    r = jb.mRCCD.do_rccd(w_can, ctx=ctx, _return_all=True)
    ph = phases(ctx)
    it_ms = ph.get("cc.iteration", [])
    res["mrccd_diis"] = {"wall_s": time.time() - t0, "iterations": r["iterations"], "ecc": r["ecc"],
                         "rms_last": float(r["rms_hist"][-1]), "ms_per_sweep_median": float(np.median(it_ms)),
                         "total_ms": ph.get("total")}
    e40 = jb.RCCD.do_rccd(w_can, ctx=ctx, _guess="mp2")
    res["mrccd_diis"]["minus_rccd_40_sweeps"] = r["ecc"] - e40
    s = json.dumps(res)
    print(s)
    if args.out:
        open(args.out, "w").write(s + "\n")
    ctx.close()


if __name__ == "__main__":
    main()
