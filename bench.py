#!/usr/bin/env python
"""bench.py -- the hot path's headline benchmark on B200.

BASELINE.json metric: "RCCSD s/iteration + 4-index transform TFLOP/s (FP64) at 1/2/4/8 B200 vs
CPU".  One JSON line is printed by rank 0:

  * a "step" is ONE RCCSD sweep (all intermediates, T1 and T2 updates, and the energy the reference
    evaluates every sweep -- RCCSD.jl:150-173,104) on device-resident integrals and amplitudes;
    `ms_per_step` is therefore the "RCCSD s/iteration" figure (x1000) and `value` the same thing
    as whole-job FP64 throughput: floating-point operations the GEMM kernels executed per sweep
    (never more than the algorithmic count F_alg of SURVEY.md section 8a) / seconds per sweep.
  * `tei_transform` holds the 4-index transform part of the metric (TFLOP/s of the full
    tei_transform(gao, C) of the same AO tensor, executed flops = 8 N^5).
  * `e2e` is the same metric through the reference-facing C-ABI call with HOST buffers (pinned):
    one complete do_rccsd (H2D of gao/C/eps + integral transformation + 40 sweeps + energy D2H).
  * `roofline` is the dominant kernel (the TMA+DMMA FP64 GEMM at the shape with the largest share of
    the sweep: the ring products (ov x ov)(ov x ov) on one GPU, the particle-particle-ladder slab on
    several), timed live with CUDA events on the library's stream, against the measured cuBLAS FP64
    peak of this pool (profiles/fp64_peak_r01.json; MEASURED_PEAKS.json records no FP64 figure).
  * `cpu_baseline` is the oracle (numpy restatement of the reference's literal algorithm, "port")
    timed on this box's host cores on a bounded sample.
  * `next_rows` (extra, N=1 only): AutoRCCSD.do_rccsd with the (T) correction on the same inputs through
    the C ABI from host buffers -- sweeps to convergence, Fock build and (T) times.

`--impl reference` times the reference's CPU algorithm (the oracle port; there is no Julia here)
on the same config and prints the same line shape.

Workload at N=1: BASELINE config 3 (RCCSD, synthetic ERIs, nbf=120, nocc=20), the largest
single-GPU configuration in BASELINE.json (the nbf=460 configuration the metric is quoted on
needs 205 GB for <vv|vv> alone and is the 8-GPU case).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

NBF, NOCC = 120, 20
WEAK_NVIR = {1: 100, 2: 124, 4: 152, 8: 192}   # F_alg(nocc=20, nvir) ~ N * F_alg(20, 100)
SEED = 2024
REF_MAXIT = 40          # RCCSD.jl:36


def flops_alg_rccsd(o, v):
    """F_alg of SURVEY.md section 8a (factorised RCCSD sweep): jues.jl_b200/flops.py."""
    from importlib import import_module
    return import_module("jues.jl_b200.flops").rccsd_iter_alg(o, v)


def flops_ref_rccsd(o, v):
    """F_ref: the reference's literal sweep (Wabef built and applied, 13 ring-type terms)."""
    from importlib import import_module
    return import_module("jues.jl_b200.flops").rccsd_iter_ref(o, v)


def make_inputs(pinned: bool):
    import jues.jl_b200 as jb
    scale = jb.synth.counter_scale(NBF)           # same element variance as the dense generator; converges
    Cao, Cav, eps = jb.synth.orbitals(NBF, NOCC, SEED)
    g = jb.synth.counter_eri(NBF, SEED, scale)
    if pinned:
        import torch
        t = torch.empty(g.size, dtype=torch.float64).pin_memory()
        gp = t.numpy().reshape(g.shape, order="F")
        gp[...] = g
        return gp, Cao, Cav, eps, scale, t
    return g, Cao, Cav, eps, scale, None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.

    NVML is queried in-process (two cheap calls every 100 ms).  An `nvidia-smi -lms 100` loop with
    the usual field list was measured to stall the GPUs for tens of milliseconds at a time (sweeps
    of 9.8 ms jittered up to 90 ms with it, on 2 GPUs), so it is only the fallback."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, device):
        self.device = device
        self.sm = []
        self.bits = 0
        self.max_mhz = None
        self.stop_flag = False
        self.period = float(os.environ.get("BENCH_SAMPLER_PERIOD", "0.1"))
        self.thread = None
        self.proc = None
        self.rows = []

    def start(self):
        if os.environ.get("BENCH_NO_SAMPLER"):
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.device]) if vis and vis.split(",")[0].isdigit() else self.device
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            self.t_clock = self.t_reason = 0.0

            def loop():
                while not self.stop_flag:
                    try:
                        t0 = time.perf_counter()
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        t1 = time.perf_counter()
                        self.bits |= int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        t2 = time.perf_counter()
                        self.t_clock = max(self.t_clock, t1 - t0)
                        self.t_reason = max(self.t_reason, t2 - t1)
                    except Exception:
                        pass
                    time.sleep(self.period)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "500"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=1)
            # under load = samples within 25 % of the maximum observed during the run
            busy = [x for x in self.sm if x >= 0.75 * max(self.sm)] if self.sm else []
            return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(n for b, n in self.REASONS.items() if self.bits & b),
                    "samples": len(self.sm), "how": f"NVML in-process, {self.period * 1e3:.0f} ms period",
                    "nvml_call_max_ms": [round(self.t_clock * 1e3, 3), round(self.t_reason * 1e3, 3)]}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for n, val in zip(names, r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm), "how": "nvidia-smi -lms 500"}


def fp64_peak():
    """Measured cuBLAS FP64 GEMM peak of this pool's B200 (tools/cublas_peak.py)."""
    p = os.path.join(ROOT, "profiles", "fp64_peak_r01.json")
    try:
        d = json.load(open(p))
        return float(d["fp64_tflops"]), "profiles/fp64_peak_r01.json (cuBLAS DGEMM burst, measured on this pool)"
    except Exception:
        return 37.0, "fallback: DMMA issue-rate ceiling 148 SM x 64 FMA/clk x 1.965 GHz"


def traffic_from_profile(name="ncu_dgemm_ladder_r01.json"):
    p = os.path.join(ROOT, "profiles", name)
    try:
        return float(json.load(open(p))["dram_bytes_per_launch"])
    except Exception:
        return None


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle on host cores
# ------------------------------------------------------------------------------------------
_ORACLE_STATE = {}


def run_oracle_sample(n_iter: int):
    """15 literal transforms (RCCSD.jl:117-142; done once per process and timed) + n_iter literal
    sweeps at the bench config.  Returns (t_transform, t_iter_avg, energy_after_first_sweep)."""
    from oracle import jues_oracle as orc
    o, v = NOCC, NBF - NOCC
    st = _ORACLE_STATE
    if not st:
        g, Cao, Cav, eps, scale, _ = make_inputs(False)
        t0 = time.perf_counter()
        st["I"] = orc.make_rccsd_integrals(g, Cao, Cav)
        st["t_tr"] = time.perf_counter() - t0
        st["Dia"], st["D"] = orc.form_Dia(o, v, eps), orc.form_Dijab(o, v, eps)
        del g
    I, Dia, D = st["I"], st["Dia"], st["D"]
    T1, T2 = np.zeros((o, v)), I["oovv"] / D
    e1 = None
    t0 = time.perf_counter()
    for k in range(n_iter):
        T1, T2 = orc.rccsd_iteration(I, T1, T2, Dia, D)
        e = orc.rccsd_energy(I["oovv"], T1, T2)
        if k == 0:
            e1 = e
    t_it = (time.perf_counter() - t0) / max(n_iter, 1)
    return st["t_tr"], t_it, e1


def reference_arm(args, rank):
    if rank != 0:
        return
    o, v = NOCC, NBF - NOCC
    cores = os.cpu_count() or 1
    t_tr_s, t_it_s = [], []
    n_iter = 1
    for step in range(args.warmup + args.steps):
        t_tr, t_it, e = run_oracle_sample(n_iter)
        if step >= args.warmup:
            t_tr_s.append(t_tr); t_it_s.append(t_it)
    t_tr, t_it = float(np.mean(t_tr_s)), float(np.mean(t_it_s))
    F = flops_alg_rccsd(o, v)
    t_call = t_tr + REF_MAXIT * t_it            # a complete do_rccsd: transforms + 40 sweeps
    tf = F / t_it * 1e-12
    e2e_tf = REF_MAXIT * F / t_call * 1e-12     # same numerator as the GPU arm: 40 sweeps of F_alg
    line = {
        "impl": "reference", "metric": "rccsd_iteration_fp64_tflops", "value": tf, "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_it * 1e3,
        "s_per_iteration": t_it, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"RCCSD nbf={NBF} nocc={NOCC} nvir={v} (BASELINE config 3), synthetic counter-based ERIs",
                   "algorithm": "reference literal: 15 tei_transforms + sweeps with materialised Wabef (numpy/OpenBLAS port)",
                   "note": "the reference's CPU path does not shard: for every --gpus N it is timed on the 1-GPU workload "
                           "(BASELINE config 3); the metric is F_alg-normalised TFLOP/s, comparable across the weak-scaling shapes"},
        "cpu_baseline": {"value": tf, "unit": "TFLOP/s", "cores": cores, "kind": "port",
                         "sample": f"15 literal transforms ({t_tr:.2f} s) + {n_iter} literal sweep(s) ({t_it:.2f} s each) "
                                   f"per step; do_rccsd extrapolated to {REF_MAXIT} sweeps"},
        "e2e": {"value": e2e_tf, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "s_per_do_rccsd": t_call, "s_per_iteration": t_call / REF_MAXIT},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def gpu_arm(args, rank, world):
    import torch
    import jues.jl_b200 as jb
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world)
    global NBF
    # weak scaling: nocc fixed, nvir grows so that F_alg per GPU stays (approximately) that of the
    # 1-GPU workload (BASELINE config 3); nvir is a multiple of 2*N so the slabs need no padding
    NBF = NOCC + WEAK_NVIR.get(world, 100)
    o, v = NOCC, NBF - NOCC
    ctx = jb.Context(local)
    if world > 1:
        ctx.init_dist(rank, world)          # NCCL communicator inside the library (id via torch.distributed)
    scale = jb.synth.counter_scale(NBF)
    Cao, Cav, eps = jb.synth.orbitals(NBF, NOCC, SEED)
    gdev = jb.DeviceFourTensor.synth_eri(NBF, seed=SEED, scale=scale, ctx=ctx)   # inputs resident in HBM
    # the same tensor in pinned host memory for the end-to-end leg (copied back from the device:
    # bit-identical to synth.counter_eri, without minutes of numpy at the larger shapes)
    keep = torch.empty(NBF ** 4, dtype=torch.float64).pin_memory()
    g = keep.numpy().reshape((NBF,) * 4, order="F")
    for s0 in range(0, NBF, 16):
        s1 = min(NBF, s0 + 16)
        g[:, :, :, s0:s1] = gdev[:, :, :, s0:s1]
    wdev = jb.Wfn(NOCC, v, eps, Cao, Cav, gdev)
    whost = jb.Wfn(NOCC, v, eps, Cao, Cav, g)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    # ---- value: K timed sweeps after W warm-up sweeps, device-resident inputs --------------
    # untimed warm-up call: lazy loading of every kernel variant, memory pools, NCCL channels
    jb.RCCSD.do_rccsd(wdev, ctx=ctx, _maxit=2)
    sampler = ClockSampler(local)
    if rank != 0:
        os.environ["BENCH_NO_SAMPLER"] = "1"      # one sampler per job: NVML queries perturb the GPUs
    barrier()
    sampler.start()
    hist = []
    e = jb.RCCSD.do_rccsd(wdev, ctx=ctx, _maxit=args.warmup + args.steps, _e_hist=hist)
    barrier()
    ph = ctx.phases()
    cnt = ctx.counters()
    it_ms = [ms for k, ms in ph if k == "cc.iteration"]
    timed = it_ms[args.warmup:]
    assert len(timed) == args.steps
    ms_step = float(np.sum(timed)) / args.steps
    tr_ms = [ms for k, ms in ph if k == "cc.transform"][0]
    # executed GEMM flops per sweep: subtract the transform's share by running the counters on
    # a zero-sweep call
    jb.RCCSD.do_rccsd(wdev, ctx=ctx, _maxit=0)
    c0 = ctx.counters()
    flops_step = (cnt["gemm_flops"] - c0["gemm_flops"]) / (args.warmup + args.steps)
    launches_step = ((cnt["gemm_launches"] - c0["gemm_launches"]) + (cnt["aux_launches"] - c0["aux_launches"])) \
        / (args.warmup + args.steps)

    # ---- the 4-index transform part of the metric -------------------------------------------------
    # (a) the integral classes of the CC run (sharded: every rank transforms its virtual slab);
    #     flops and time of the zero-sweep call above
    tr0_ms = [ms for k, ms in ctx.phases() if k == "cc.transform"][0]
    tr_flops = c0["gemm_flops"]
    if dist is not None:
        tt_ = torch.tensor([tr_flops, -tr0_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt_[0:1], op=dist.ReduceOp.SUM)
        dist.all_reduce(tt_[1:2], op=dist.ReduceOp.MIN)
        tr_flops, tr0_ms = float(tt_[0]), -float(tt_[1])
    tei = {"cc_classes": {"workload": f"7 MO classes of RCCSD (<oo|vv>,<ov|ov>,<oo|oo>,<oo|ov>,<vv|vv>,<vv|ov>,<vo|vv>) "
                                      f"nbf={NBF}, virtual slab per rank", "ms": tr0_ms,
                          "tflops": tr_flops / tr0_ms * 1e-9}}
    # (b) full tei_transform(gao, C) (all four indices, 8 N^5 flop) -- single GPU
    if world == 1:
        Cfull = np.asfortranarray(np.hstack([Cao, Cav]))
        for _ in range(2):
            out = jb.tei_transform(gdev, Cfull, "bench", ctx=ctx)
            out.free()
        tt_ms = [ms for k, ms in ctx.phases() if k == "tei.transform"][0]
        tei["full"] = {"workload": f"tei_transform(gao, C) nbf={NBF} (all four indices, 8 N^5 flop)", "ms": tt_ms,
                       "tflops": ctx.counters()["gemm_flops"] / tt_ms * 1e-9}

    # ---- e2e: one complete do_rccsd through the C ABI from pinned HOST buffers ----------------
    e2e_t = []
    for rep in range(2):
        barrier()
        t0 = time.perf_counter()
        e_host = jb.RCCSD.do_rccsd(whost, ctx=ctx)          # 40 sweeps, the reference's count
        torch.cuda.synchronize()
        e2e_t.append(time.perf_counter() - t0)
    t_call = min(e2e_t)
    c_full = ctx.counters()
    clocks = sampler.stop()

    # ---- roofline of the dominant kernel, timed live (CUDA events on the library's stream) ----
    if world == 1:
        # one rank: the pp-ladder runs in the packed symmetric/antisymmetric pair space (half its flops),
        # which leaves the seven ring products (ov x ov)(ov x ov) as the largest share of the sweep
        # (41 % of its kernel time, profiles/ncu_launches_rccsd_sweep_r01b.csv)
        M = Nn = K = o * v
        tA, shape, prof = "T", f"ring GEMM (ov x ov)(ov x ov) M=N=K={M}", "ncu_dgemm_ring_r01.json"
    else:
        # several ranks: this rank's column block of the packed (symmetric/antisymmetric) pp-ladder,
        # tau+-(ij,(ef)) x W+-((ef),(ab)): two such products per sweep in one batched launch
        M, Nn, K = o * o, (v // world) * (v // 2 + 1), v * (v + 1) // 2
        tA, shape, prof = "N", f"packed pp-ladder GEMM M={M} N={Nn} K={K} (x2 per sweep)", "ncu_dgemm_ladder_r01.json"
    ms_gemm = ctx.gemm_bench(tA, "N", M, Nn, K, reps=5)
    peak, peak_src = fp64_peak()
    ach = 2.0 * M * Nn * K / ms_gemm * 1e-9
    roofline = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                # measured DRAM bytes per launch (ncu --set full) exist for the 1-GPU shape only
                "traffic": traffic_from_profile(prof) if (world == 1 and NBF == 120) else None,
                "kernel": "jues::gemm::dgemm_tma_dmma (FP64 DMMA.8x8x4 fed by TMA)",
                "shape": shape, "ms_per_launch": ms_gemm, "peak_source": peak_src,
                "sweep_frac_of_peak": flops_step / (ms_step * 1e-3) * 1e-12 / peak}

    # ---- max over ranks -----------------------------------------------------------------------
    comm = ctx.comm_counters()
    if dist is not None:
        t = torch.tensor([ms_step, t_call], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, t_call = float(t[0]), float(t[1])
        fs = torch.tensor([flops_step, c_full["gemm_flops"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(fs, op=dist.ReduceOp.SUM)      # whole-job executed flops
        flops_step, c_full["gemm_flops"] = float(fs[0]), float(fs[1])
        roofline["sweep_frac_of_peak"] = flops_step / (ms_step * 1e-3) * 1e-12 / (peak * world)

    if rank != 0:
        return
    # ---- CPU baseline on this box's host cores (rank 0, N=1 only), bounded sample ------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        t_tr, t_it, e_cpu = run_oracle_sample(1)
        cpu = {"value": flops_alg_rccsd(o, v) / t_it * 1e-12, "unit": "TFLOP/s", "cores": os.cpu_count() or 1,
               "kind": "port", "s_per_iteration": t_it, "s_transforms": t_tr,
               "sample": f"oracle (numpy/OpenBLAS port of the reference's literal algorithm): 15 transforms "
                         f"({t_tr:.1f} s) + 1 sweep ({t_it:.1f} s) at nbf={NBF} nocc={NOCC}",
               "energy_check": abs(e_cpu - hist[1]) if len(hist) > 1 else None}
    # ---- the callers either side of the path, same inputs, one GPU: AutoRCCSD to
    #      |dE|, rms <= 1e-10 with the (T) correction, through the C ABI from host buffers.  Extra keys only;
    #      never allowed to break the line.
    next_rows = None
    if world == 1 and not args.no_next_rows:
        try:
            from importlib import import_module
            fl = import_module("jues.jl_b200.flops")
            t0 = time.perf_counter()
            hao = jb.synth.core_hamiltonian(g, Cao, Cav, eps)      # host plumbing: makes (C, eps) the RHF solution
            t_h = time.perf_counter() - t0
            wa = jb.Wfn(NOCC, v, eps, Cao, Cav, g, hao=hao)
            t0 = time.perf_counter()
            r = jb.AutoRCCSD.do_rccsd(wa, ctx=ctx, do_pT=True, _return_all=True)
            t_auto = time.perf_counter() - t0
            pa = ctx.phases()
            sw = [ms for k, ms in pa if k == "cc.iteration"]
            tr = [ms for k, ms in pa if k == "cc.triples"]
            next_rows = {"what": "AutoRCCSD.do_rccsd(do_pT=true) from host buffers: get_fock + transform + sweeps to "
                                 "1e-10 + (T)", "s_per_call": t_auto, "host_core_hamiltonian_s": t_h,
                         "iterations": r["iterations"], "converged": r["converged"], "e_ccsd": r["ecc"], "e_pt": r["ept"],
                         "ms_per_sweep_median": float(np.median(sw)) if sw else None,
                         "fock_build_ms": [ms for k, ms in pa if k == "fock.build"],
                         "triples_ms": tr[0] if tr else None,
                         "triples_tflops": fl.pt_flops(NOCC, v) / (tr[0] * 1e-3) * 1e-12 if tr else None}
        except Exception as ex:     # noqa: BLE001
            next_rows = {"error": str(ex)[:300]}
    F_alg = flops_alg_rccsd(o, v)
    line = {
        # algorithmic FP64 throughput: F_alg per sweep / seconds per sweep (both arms use F_alg);
        # the flops the kernels actually executed (<= F_alg) are in config and roofline
        "metric": "rccsd_iteration_fp64_tflops", "value": F_alg / (ms_step * 1e-3) * 1e-12,
        "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "s_per_iteration": ms_step * 1e-3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"RCCSD nbf={NBF} nocc={NOCC} nvir={v}"
                               + (" (BASELINE config 3)" if world == 1 else
                                  f" (config 3 grown for weak scaling: F_alg = {F_alg / flops_alg_rccsd(NOCC, 100):.2f} x the 1-GPU workload)")
                               + f", synthetic counter-based ERIs seed={SEED}",
                   "parallelism": f"virtual-index slabs over {world} GPU(s), NCCL all-gather of H and T2 + one all-reduce per sweep"
                                  if world > 1 else "single GPU",
                   "collectives_per_call": comm,
                   "step": "one RCCSD sweep (intermediates + T1 + T2 + energy) on device-resident data",
                   "l2": "inputs larger than L2 (<vv|vv> >= 0.8 GB per GPU, amplitudes/intermediates >= 32 MB each, re-streamed every sweep)",
                   "flops_executed_per_sweep": flops_step, "F_alg_per_sweep": F_alg, "F_ref_per_sweep": flops_ref_rccsd(o, v)},
        "tei_transform": tei,
        "e2e": {"value": REF_MAXIT * F_alg / t_call * 1e-12, "unit": "TFLOP/s",
                "flops_executed": c_full["gemm_flops"],
                "h2d_bytes_per_step": int(g.nbytes + Cao.nbytes + Cav.nbytes + eps.nbytes),
                "d2h_bytes_per_step": int(8 * (REF_MAXIT + 2)),
                "s_per_do_rccsd": t_call, "s_per_iteration": t_call / REF_MAXIT,
                "what": "jues_b200_rccsd(host gao, Cao, Cav, eps, maxit=40): H2D + transform + 40 sweeps + energies D2H"},
        "gpu_launches": int(round(launches_step)),
        "ms_each_step_rank0": [round(x, 3) for x in it_ms],
        # host clock (ms since the first sweep was issued) at which each sweep had been handed to the
        # driver, and how many sweeps were replayed from a CUDA graph: the launching thread runs ahead
        "host_issue_ms_rank0": [round(ms, 3) for k, ms in ph if k == "cc.iteration.host_ms"],
        "graph_replayed_sweeps": int(sum(ms for k, ms in ph if k == "cc.graph_launches")),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "clocks": clocks,
        "energy": {"E_ccsd_after_timed_sweeps": e, "E_ccsd_40_sweeps_e2e": e_host},
        "integral_transform_ms": tr_ms,
        "next_rows": next_rows,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-next-rows", action="store_true", help="skip the AutoRCCSD(T) extra keys")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout must carry exactly ONE JSON line: libraries (NCCL prints its version banner on fd 1) are
    # diverted to stderr for the duration of the run; the line is written to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    gpu_arm(args, rank, world)


if __name__ == "__main__":
    main()
