#!/usr/bin/env python
"""bench.py -- the hot path's headline benchmark on B200.

BASELINE.json metric: "RCCSD s/iteration + 4-index transform TFLOP/s (FP64) at 1/2/4/8 B200 vs
CPU".  One JSON line is printed by rank 0:

  * a "step" is ONE RCCSD sweep (all intermediates, T1 and T2 updates, and the energy the reference
    evaluates every sweep -- RCCSD.jl:150-173,104) on device-resident integrals and amplitudes;
    `ms_per_step` is the "RCCSD s/iteration" figure (x1000) and `value` the same thing as whole-job
    FP64 throughput: floating-point operations the GEMM kernels of all ranks EXECUTED per sweep
    (counted by the library per launch) / seconds per sweep.  The algorithm executes fewer flops than
    SURVEY.md section 8a's F_alg and far fewer than the reference's literal F_ref; throughput normalised
    by those counts is reported as labelled extras (`config.alg_normalised_tflops`, `.ref_normalised_tflops`),
    never as `value`.
  * `tei_transform` holds the 4-index transform part of the metric: the sharded one-pass transform that
    produces every integral class of the run (executed flops, ms, fraction of the FP64 peak) and, on one
    GPU, the full tei_transform(gao, C).
  * `e2e` is the same metric through the reference-facing C-ABI call with HOST buffers (pinned): one
    complete do_rccsd (H2D of this rank's share of gao + integral transformation + 40 sweeps + D2H).
  * `roofline` is the dominant kernel (the TMA+DMMA FP64 GEMM at the shape with the largest share of
    the sweep), its launch duration measured IN the sweep (one CUDA-event pair per launch on the library's
    stream, `jues_b200_set_trace(ctx, 2)`), the isolated back-to-back figure beside it, against the measured
    cuBLAS FP64 peak of this pool (profiles/fp64_peak_r01.json; MEASURED_PEAKS.json has no FP64 figure).
  * `parity`: the per-sweep energies of the timed run against the committed oracle trace for this shape
    (tests/golden/bench_ehist_nbf*_nocc20.npz, made by tests/golden/make_bench_golden.py from the literal
    reference algorithm); the process exits non-zero when |dE| exceeds 1e-10 Eh.
  * `cpu_baseline` is the oracle (numpy restatement of the reference's literal algorithm, "port") timed on
    this box's host cores on a bounded sample.
  * `large`: the configurations the metric's targets are quoted on -- RCCSD nbf=300/nocc=60 (fits one GPU:
    strong scaling over N), the 4-index transform + RMP2 at BASELINE config 4 (nbf=500/nocc=60, strong scaling
    over N) and, at N=8, BASELINE config 5 (RCCSD nbf=460/nocc=60): s/iteration, executed TFLOP/s per GPU,
    transform time and TFLOP/s, communication ms per sweep, energies against the same inputs at other rank counts.
  * `next_rows` (extra, N=1 only): AutoRCCSD.do_rccsd with the (T) correction on the same inputs.

`--impl reference` times the reference's CPU algorithm (the oracle port; there is no Julia here) on the
same shapes, with every host core, on a bounded sample, and prints the same line shape.

Workload at N=1: BASELINE config 3 (RCCSD, synthetic ERIs, nbf=120, nocc=20), the largest of BASELINE's
named single-GPU configurations; at N>1 it grows with N at fixed nocc (weak scaling).
"""
from __future__ import annotations

import os
import sys

if "--impl" in sys.argv and "reference" in sys.argv:
    # The reference arm uses every host core: torch.distributed.run exports OMP_NUM_THREADS=1 for N>1,
    # and OpenBLAS reads its thread count when numpy is imported -- so this comes first.
    _cores = str(os.cpu_count() or 1)
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = _cores

import argparse  # noqa: E402
import json  # noqa: E402
import subprocess  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

NBF, NOCC = 120, 20
WEAK_NVIR = {1: 100, 2: 124, 4: 152, 8: 192}   # F_alg(nocc=20, nvir) ~ N * F_alg(20, 100)
SEED = 2024
REF_MAXIT = 40          # RCCSD.jl:36
E_TOL = 1e-10           # north_star: correlation energies within 1e-10 Eh
LARGE = {"strong": (300, 60), "c4": (500, 60), "c5": (460, 60)}


def _flops():
    from importlib import import_module
    return import_module("jues.jl_b200.flops")


def golden_trace(nbf, nocc):
    p = os.path.join(ROOT, "tests", "golden", f"bench_ehist_nbf{nbf}_nocc{nocc}.npz")
    if not os.path.exists(p):
        return None, os.path.relpath(p, ROOT)
    return np.load(p), os.path.relpath(p, ROOT)


def make_inputs(nbf, nocc):
    import jues.jl_b200 as jb
    scale = jb.synth.counter_scale(nbf)           # same element variance as the dense generator; converges
    Cao, Cav, eps = jb.synth.orbitals(nbf, nocc, SEED)
    g = jb.synth.counter_eri(nbf, SEED, scale)
    return g, Cao, Cav, eps


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


def traffic_from_profile(name):
    p = os.path.join(ROOT, "profiles", name)
    try:
        return float(json.load(open(p))["dram_bytes_per_launch"])
    except Exception:
        return None


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle on host cores
# ------------------------------------------------------------------------------------------
def cpu_info():
    model = None
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    threads = None
    try:
        from threadpoolctl import threadpool_info
        threads = max((d.get("num_threads", 0) for d in threadpool_info()), default=None)
    except Exception:
        pass
    n = 4096
    a = np.random.default_rng(0).standard_normal((n, n))
    b = a.T.copy()
    a @ b
    t0 = time.perf_counter()
    a @ b
    dt = time.perf_counter() - t0
    return {"cpu_model": model, "host_cores": os.cpu_count() or 1, "blas_threads": threads,
            "dgemm_4096_tflops": 2.0 * n ** 3 / dt * 1e-12,
            "env": {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS")}}


def oracle_classes(nbf, nocc, real_inputs):
    """The 15 arrays of make_rccsd_integrals (RCCSD.jl:117-142).  real_inputs: from the bench inputs through
    the oracle's literal 15 transforms (timed).  Otherwise arrays of the right shapes filled with random
    numbers of the right magnitude -- the sweep's cost does not depend on the values -- and ONE literal
    transform (the <vv|vv> one, the most expensive of the 15) timed on a random AO tensor, the other 14
    extrapolated by their flop counts."""
    from oracle import jues_oracle as orc
    fl = _flops()
    o, v = nocc, nbf - nocc
    if real_inputs:
        g, Cao, Cav, eps = make_inputs(nbf, nocc)
        t0 = time.perf_counter()
        I = orc.make_rccsd_integrals(g, Cao, Cav)
        return I, eps, time.perf_counter() - t0, "15 literal transforms timed"
    rng = np.random.default_rng(1)
    g = rng.standard_normal((nbf,) * 4)
    C = rng.standard_normal((nbf, nbf)) / np.sqrt(nbf)
    t0 = time.perf_counter()
    orc.tei_transform(g, C[:, o:], C[:, o:], C[:, o:], C[:, o:])
    t_v = time.perf_counter() - t0
    del g
    t_tr = t_v * fl.rccsd_transforms_ref(nbf, o, v) / fl.tei_flops_ref(nbf, v, v, v, v)
    shape = {"o": o, "v": v}
    I = {k: 1e-3 * rng.standard_normal(tuple(shape[c] for c in k)) for k in
         ("vvvv", "ovvv", "vovv", "vvov", "vvvo", "oovv", "ovvo", "vovo", "ovov", "voov", "ooov", "oovo", "ovoo",
          "vooo", "oooo")}
    eps = np.concatenate([-1.0 - rng.random(o), 1.0 + rng.random(v)])
    return I, eps, t_tr, f"<vv|vv> literal transform timed ({t_v:.1f} s) on a random AO tensor, the 15 extrapolated by flops"


_ORACLE_STATE = {}


def run_oracle_sample(nbf, nocc, n_iter, real_inputs=True):
    """Integral classes (once per process) + n_iter literal sweeps.  Returns (t_transform, t_iter_avg,
    energy_after_first_sweep, how)."""
    from oracle import jues_oracle as orc
    o, v = nocc, nbf - nocc
    st = _ORACLE_STATE
    key = (nbf, nocc, real_inputs)
    if st.get("key") != key:
        st.clear()
        st["key"] = key
        st["I"], eps, st["t_tr"], st["how"] = oracle_classes(nbf, nocc, real_inputs)
        st["Dia"], st["D"] = orc.form_Dia(o, v, eps), orc.form_Dijab(o, v, eps)
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
    return st["t_tr"], t_it, e1, st["how"]


def reference_arm(args, rank, world):
    if rank != 0:
        return
    fl = _flops()
    world = max(world, args.gpus)      # the shape of the N-GPU arm, whichever way this arm was launched
    nbf = NOCC + WEAK_NVIR.get(world, 100)
    o, v = NOCC, nbf - NOCC
    info = cpu_info()
    real = world == 1           # N=1: the bench inputs themselves (energy cross-check with the GPU arm)
    budget_s = float(os.environ.get("BENCH_REF_BUDGET_S", "240"))
    t_start = time.perf_counter()
    t_tr, t_first, e1, how = run_oracle_sample(nbf, NOCC, 1, real)      # first sweep = warm-up
    t_it_s = []
    want = args.warmup + args.steps
    for step in range(want):
        if time.perf_counter() - t_start + t_first > budget_s and len(t_it_s) >= 1:
            break
        _, t_it, _, _ = run_oracle_sample(nbf, NOCC, 1, real)
        t_it_s.append(t_it)
    timed = t_it_s[min(args.warmup, len(t_it_s) - 1):] or t_it_s
    t_it = float(np.mean(timed))
    F_ref = fl.rccsd_iter_ref(o, v)                 # the flops this algorithm executes per sweep
    F_tr = fl.rccsd_transforms_ref(nbf, o, v)
    t_call = t_tr + REF_MAXIT * t_it                # a complete do_rccsd: transforms + 40 sweeps
    # The unit of work is the SAME for both arms -- the FP64 operations of one sweep (one complete do_rccsd) as
    # the B200 arm executes them (flops.rccsd_iter_exec / cc_transform_exec: its counted flops in closed form) --
    # so that value(b200) / value(reference) is the ratio of the two times.  What this algorithm itself
    # executes per second (F_ref / s, 2.3x more operations for the same sweep) is a labelled extra.
    F_unit, F_unit_tr = fl.rccsd_iter_exec(o, v), fl.cc_transform_exec(nbf)
    tf = F_unit / t_it * 1e-12
    e2e_tf = (F_unit_tr + REF_MAXIT * F_unit) / t_call * 1e-12
    own_tf = F_ref / t_it * 1e-12
    own_e2e_tf = (F_tr + REF_MAXIT * F_ref) / t_call * 1e-12
    sample = (f"{how}; {len(timed)} literal sweep(s) timed ({t_it:.2f} s each, {len(t_it_s) - len(timed)} warm-up) "
              f"within a {budget_s:.0f} s budget; do_rccsd extrapolated to {REF_MAXIT} sweeps")
    line = {
        "impl": "reference", "metric": "rccsd_iteration_fp64_tflops", "value": tf, "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_it * 1e3,
        "s_per_iteration": t_it, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"RCCSD nbf={nbf} nocc={NOCC} nvir={v}"
                               + (" (BASELINE config 3)" if world == 1 else " (config 3 grown for weak scaling, the GPU arm's shape)")
                               + ", synthetic ERIs",
                   "algorithm": "reference literal: 15 tei_transforms + sweeps with materialised Wabef (numpy/OpenBLAS port)",
                   "flops_unit_per_sweep": F_unit, "flops_executed_per_sweep": F_ref,
                   "own_executed_tflops": own_tf, "own_executed_e2e_tflops": own_e2e_tf,
                   "F_alg_per_sweep": fl.rccsd_iter_alg(o, v),
                   "alg_normalised_tflops": fl.rccsd_iter_alg(o, v) / t_it * 1e-12,
                   "value_is": "flops_unit_per_sweep / s_per_iteration: the sweep's FP64 operations as the B200 arm "
                               "executes them (the common unit of work of both arms) / this arm's seconds, so "
                               "value(b200) / value(reference) = ratio of times; own_executed_tflops = what this "
                               "algorithm itself executes per second (F_ref / s)",
                   "note": "the CPU path does not shard, so rank 0 alone runs the whole workload of the N-GPU arm"},
        "cpu_baseline": {"value": tf, "unit": "TFLOP/s", "cores": info["blas_threads"] or info["host_cores"],
                         "kind": "port", "sample": sample, "s_per_iteration": t_it, "s_transforms": t_tr,
                         "energy_after_first_sweep": e1 if real else None, **info},
        "e2e": {"value": e2e_tf, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "s_per_do_rccsd": t_call, "s_per_iteration": t_call / REF_MAXIT},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def split_sweeps(ph):
    """Phases of a traced call -> list of per-sweep lists [(name, ms)], closed by each cc.iteration entry."""
    out, cur = [], []
    for k, ms in ph:
        if k == "cc.iteration":
            out.append((ms, cur))
            cur = []
        elif k.startswith("gemm ") or k.startswith("cc.comm.") or k.startswith("cc.part."):
            cur.append((k, ms))
    return out


def insitu_roofline(ctx, jb, wdev, peak, peak_src, world, nbf):
    """Dominant GEMM of the sweep, its launch duration measured inside the sweep."""
    ctx.set_trace(2)
    try:
        jb.RCCSD.do_rccsd(wdev, ctx=ctx, _maxit=5)
        sweeps = split_sweeps(ctx.phases())
    finally:
        ctx.set_trace(0)
    sweeps = sweeps[2:] or sweeps           # the first sweeps build lazily created operand copies
    tot = {}
    for ms_sweep, items in sweeps:
        for k, ms in items:
            if k.startswith("gemm "):
                tot.setdefault(" ".join(k.split()[:2]), []).append(ms)     # by shape: drop the layout / beta suffix
    name = max(tot, key=lambda k: sum(tot[k]))
    M, N, K, B = (int(x) for x in name.split()[1].split("x"))
    per_sweep = len(tot[name]) / len(sweeps)
    ms_launch = float(np.mean(tot[name]))
    ms_sweep = float(np.mean([s for s, _ in sweeps]))
    ach = 2.0 * M * N * K * B / ms_launch * 1e-9
    iso = B * ctx.gemm_bench("T" if (M == N == K) else "N", "N", M, N, K, reps=5) if B <= 2 else None
    comm = {}
    for _, items in sweeps:
        for k, ms in items:
            if k.startswith("cc.comm."):
                comm[k] = comm.get(k, 0.0) + ms / len(sweeps)
    all_gemm = sum(sum(v) for v in tot.values()) / len(sweeps)
    return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "traffic": traffic_from_profile("ncu_dgemm_ring_r02.json") if (world == 1 and nbf == 120 and M == 2000) else None,
            "kernel": "jues::gemm::dgemm_tma_dmma (FP64 DMMA.8x8x4 fed by TMA)",
            "shape": f"M={M} N={N} K={K} batch={B} ({per_sweep:.0f} launches per sweep)",
            "ms_per_launch": ms_launch, "how": "CUDA-event pair around every launch inside traced eager sweeps",
            "share_of_sweep": ms_launch * per_sweep / ms_sweep, "all_gemms_share_of_sweep": all_gemm / ms_sweep,
            "traced_sweep_ms": ms_sweep, "isolated_ms_per_launch": iso,
            "isolated_tflops": (2.0 * M * N * K * B / iso * 1e-9) if iso else None,
            "peak_source": peak_src}, comm


def large_invariance():
    try:
        return json.load(open(os.path.join(ROOT, "tests", "golden", "large_invariance.json")))
    except Exception:
        return {}


def large_block(make_ctx, jb, dist, torch, world, rank, peak):
    """The shapes the metric's targets are quoted on, storage-less synthetic AO tensor: RCCSD nbf=300/nocc=60
    (strong scaling), RMP2 + transform nbf=500/nocc=60 (BASELINE config 4) and, on 8 GPUs, RCCSD nbf=460/nocc=60
    (BASELINE config 5).  Energies are compared with the committed values of the same inputs at other rank counts
    (tests/golden/large_invariance.json: sharding invariance, not an oracle).  Every shape runs on a FRESH
    context: the cached device blocks of the previous shape would otherwise be flushed inside the timed
    transform (config 5: 1.15 s against 0.82 s)."""
    res = {}
    inv = large_invariance()
    todo = [("strong", "rccsd", *LARGE["strong"]), ("c4", "rmp2", *LARGE["c4"])] \
        + ([("c5", "rccsd", *LARGE["c5"])] if world == 8 else [])

    def allmax(vals):
        if dist is None:
            return vals
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def allsum(vals):
        if dist is None:
            return vals
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(x) for x in t]

    for tag, what, nbf, nocc in todo:
        ctx = None
        try:
            ctx = make_ctx()
            v = nbf - nocc
            Cao, Cav, eps = jb.synth.orbitals(nbf, nocc, SEED)
            g = jb.DeviceFourTensor.synth_eri(nbf, seed=SEED, ctx=ctx, virtual=True)
            w = jb.Wfn(nocc, v, eps, Cao, Cav, g)
            label = f"nbf={nbf} nocc={nocc} nvir={v}" + {"c4": " (BASELINE config 4)", "c5": " (BASELINE config 5)"}.get(tag, "") \
                + f", generated AO integrals, {world} GPU(s)"
            if what == "rmp2":
                jb.do_rmp2(w, ctx=ctx)                      # warm-up: block cache, kernel variants
                torch.cuda.synchronize()
                if dist is not None:
                    dist.barrier()
                ctx.set_trace(1)
                try:
                    t0 = time.perf_counter()
                    e = jb.do_rmp2(w, ctx=ctx)
                    wall = time.perf_counter() - t0
                finally:
                    ctx.set_trace(0)
                ph, c = ctx.phases(), ctx.counters()
                tr = sum(ms for k, ms in ph if k == "mp2.transform")
                en = sum(ms for k, ms in ph if k == "mp2.energy")
                exch = sum(ms for k, ms in ph if k == "tei.exchange")
                gen = sum(ms for k, ms in ph if k == "tei.block")
                wall, tr, en, gen = allmax([wall, tr, en, gen])
                (fl_tr,) = allsum([c["gemm_flops"]])
                ref = inv.get(f"rmp2_nbf{nbf}_nocc{nocc}", {}).get("energy")
                res[tag] = {"workload": "4-index transform + RMP2 " + label, "scaling": "strong",
                            "s_per_call": wall, "transform_s": tr * 1e-3, "mp2_energy_ms": en,
                            "flops_executed": fl_tr, "transform_tflops_per_gpu": fl_tr / (tr * 1e-3) * 1e-12 / world,
                            "transform_frac_of_fp64_peak_per_gpu": fl_tr / (tr * 1e-3) * 1e-12 / world / peak,
                            "transform_exchange_ms_overlapped": exch,
                            # the AO blocks are GENERATED inside the timed transform (counter-based hash, integer-ALU
                            # bound at ~2.2 TB/s of output): not part of the reference's algorithm, so also without it
                            "ao_generation_ms": gen,
                            "transform_tflops_per_gpu_excl_generation": fl_tr / max((tr - gen) * 1e-3, 1e-9) * 1e-12 / world,
                            "energy": e,
                            "abs_dE_vs_other_rank_counts": abs(e - ref) if ref is not None else None,
                            "peak_device_GB": c["bytes_peak"] / 1e9}
            else:
                # one call: the transform is timed on a context that has never seen this shape, device block
                # allocation included (a second call was measured SLOWER at these memory-heavy shapes -- 1.03 s
                # against 0.74 s at nbf=300: the block cache of the first call is flushed to make room)
                hist = []
                ctx.set_trace(1)
                try:
                    jb.RCCSD.do_rccsd(w, ctx=ctx, _maxit=3, _e_hist=hist)
                finally:
                    ctx.set_trace(0)
                ph, c = ctx.phases(), ctx.counters()
                it = [ms for k, ms in ph if k == "cc.iteration"]
                gf = [ms for k, ms in ph if k == "cc.iteration.gflop"]
                tr = [ms for k, ms in ph if k == "cc.transform"][0]
                ms_it = float(np.median(it[1:]))
                comm_ms = sum(ms for k, ms in ph if k.startswith("cc.comm.")) / len(it)
                exch_ms = sum(ms for k, ms in ph if k == "tei.exchange")
                gen = sum(ms for k, ms in ph if k == "tei.block")
                tr_flops = c["gemm_flops"] - sum(gf) * 1e9
                ms_it, tr, comm_ms, gen = allmax([ms_it, tr, comm_ms, gen])
                fl_it, fl_tr = allsum([float(np.median(gf)) * 1e9, tr_flops])
                ref = inv.get(f"rccsd_nbf{nbf}_nocc{nocc}", {}).get("e_hist")
                d_e = float(np.abs(np.asarray(hist[:len(ref)]) - np.asarray(ref[:len(hist)])).max()) if ref else None
                res[tag] = {"workload": "RCCSD " + label, "scaling": "strong",
                            "s_per_iteration": ms_it * 1e-3, "executed_tflops_per_gpu": fl_it / (ms_it * 1e-3) * 1e-12 / world,
                            "frac_of_fp64_peak_per_gpu": fl_it / (ms_it * 1e-3) * 1e-12 / world / peak,
                            "transform_s": tr * 1e-3, "transform_tflops_per_gpu": fl_tr / (tr * 1e-3) * 1e-12 / world,
                            "transform_frac_of_fp64_peak_per_gpu": fl_tr / (tr * 1e-3) * 1e-12 / world / peak,
                            "ao_generation_ms": gen,
                            "transform_tflops_per_gpu_excl_generation": fl_tr / max((tr - gen) * 1e-3, 1e-9) * 1e-12 / world,
                            "comm_ms_per_sweep": comm_ms, "transform_exchange_ms_overlapped": exch_ms,
                            "e_hist": hist, "max_abs_dE_vs_other_rank_counts": d_e,
                            "peak_device_GB": c["bytes_peak"] / 1e9}
            g.free()
        except Exception as ex:     # noqa: BLE001
            res[tag] = {"error": str(ex)[:300]}
        finally:
            if ctx is not None:
                try:
                    ctx.close()
                except Exception:       # noqa: BLE001
                    pass
    res["invariance_reference"] = "tests/golden/large_invariance.json (same inputs at other rank counts / round-1 algorithm; not an oracle)"
    return res


def gpu_arm(args, rank, world):
    import torch
    import jues.jl_b200 as jb
    fl = _flops()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world)
    # weak scaling: nocc fixed, nvir grows so that F_alg per GPU stays (approximately) that of the
    # 1-GPU workload (BASELINE config 3); nvir is a multiple of 2*N so the slabs need no padding
    nbf = NOCC + WEAK_NVIR.get(world, 100)
    o, v = NOCC, nbf - NOCC
    ctx = jb.Context(local)
    if world > 1:
        ctx.init_dist(rank, world)          # NCCL communicator inside the library (id via torch.distributed)
    scale = jb.synth.counter_scale(nbf)
    Cao, Cav, eps = jb.synth.orbitals(nbf, NOCC, SEED)
    gdev = jb.DeviceFourTensor.synth_eri(nbf, seed=SEED, scale=scale, ctx=ctx)   # inputs resident in HBM
    # the same tensor in pinned host memory for the end-to-end leg (copied back from the device:
    # bit-identical to synth.counter_eri, without minutes of numpy at the larger shapes)
    keep = torch.empty(nbf ** 4, dtype=torch.float64).pin_memory()
    g = keep.numpy().reshape((nbf,) * 4, order="F")
    for s0 in range(0, nbf, 16):
        s1 = min(nbf, s0 + 16)
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

    peak, peak_src = fp64_peak()
    # ---- the 4-index transform part of the metric -------------------------------------------------
    # (a) every integral class of the CC run from ONE sharded pass (flops and time of the zero-sweep call)
    tr0_ms = [ms for k, ms in ctx.phases() if k == "cc.transform"][0]
    tr_flops = c0["gemm_flops"]
    if dist is not None:
        tt_ = torch.tensor([tr_flops, tr0_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt_[0:1], op=dist.ReduceOp.SUM)
        dist.all_reduce(tt_[1:2], op=dist.ReduceOp.MAX)
        tr_flops, tr0_ms = float(tt_[0]), float(tt_[1])
    tei = {"cc_classes": {"workload": f"all MO integral classes of RCCSD from one sharded pass over gao (device-resident), "
                                      f"nbf={nbf}, last index split over {world} rank(s)", "ms": tr0_ms,
                          "flops_executed": tr_flops, "tflops": tr_flops / tr0_ms * 1e-9,
                          "frac_of_fp64_peak": tr_flops / tr0_ms * 1e-9 / (peak * world)}}
    # (b) full tei_transform(gao, C) (all four indices, 8 N^5 flop) -- single GPU
    if world == 1:
        Cfull = np.asfortranarray(np.hstack([Cao, Cav]))
        for _ in range(2):
            out = jb.tei_transform(gdev, Cfull, "bench", ctx=ctx)
            out.free()
        tt_ms = [ms for k, ms in ctx.phases() if k == "tei.transform"][0]
        tei["full"] = {"workload": f"tei_transform(gao, C) nbf={nbf} (all four indices, 8 N^5 flop)", "ms": tt_ms,
                       "tflops": ctx.counters()["gemm_flops"] / tt_ms * 1e-9,
                       "frac_of_fp64_peak": ctx.counters()["gemm_flops"] / tt_ms * 1e-9 / peak}

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
    e2e_phases = {}
    for k, ms in ctx.phases():
        if k in ("cc.transform", "cc.static", "total"):
            e2e_phases[k] = ms
    clocks = sampler.stop()

    # ---- roofline of the dominant kernel, measured inside the sweep ---------------------------------
    roofline, comm_ms = insitu_roofline(ctx, jb, wdev, peak, peak_src, world, nbf)

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
    roofline["comm_ms_per_traced_sweep"] = comm_ms

    def run_large():
        """The configurations the targets are quoted on, on a FRESH context: the block cache of the config-3
        runs would otherwise be flushed inside the timed transform (seconds of cudaFree)."""
        gdev.free()
        if args.no_large:
            return None
        ctx.close()

        def make_ctx():
            c2 = jb.Context(local)
            if world > 1:
                c2.init_dist(rank, world)
            return c2

        return large_block(make_ctx, jb, dist, torch, world, rank, peak)

    if rank != 0:
        run_large()
        return 0
    # ---- parity: the timed run's energies against the committed oracle trace ----------------------
    gold, gold_path = golden_trace(nbf, NOCC)
    if gold is None:
        parity = {"status": "unpinned", "golden": gold_path, "tol": E_TOL,
                  "note": "no committed oracle trace for this shape"}
    else:
        ref = np.asarray(gold["e_hist"])
        n = min(len(ref), len(hist))
        d_run = float(np.abs(np.asarray(hist[:n]) - ref[:n]).max())
        d_e2e = float(abs(e_host - ref[REF_MAXIT])) if len(ref) > REF_MAXIT else None
        worst = max(d_run, d_e2e or 0.0)
        parity = {"status": "ok" if worst <= E_TOL else "FAILED", "max_abs_dE": worst, "tol": E_TOL,
                  "max_abs_dE_timed_run": d_run, "sweeps_compared": n, "abs_dE_e2e_40_sweeps": d_e2e,
                  "golden": gold_path,
                  "oracle": (str(gold["model"]) if "model" in gold.files else
                             "oracle/jues_oracle.py (literal RCCSD.jl:150-289)") + ", tests/golden/make_bench_golden.py"}
    # ---- the callers either side of the path, same inputs, one GPU: AutoRCCSD to
    #      |dE|, rms <= 1e-10 with the (T) correction, through the C ABI from host buffers.  Extra keys only;
    #      never allowed to break the line.
    next_rows = None
    if world == 1 and not args.no_next_rows:
        try:
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
                         "triples_tflops": fl.pt_flops(NOCC, v) / (tr[0] * 1e-3) * 1e-12 if tr else None,
                         "triples_frac_of_fp64_peak": fl.pt_flops(NOCC, v) / (tr[0] * 1e-3) * 1e-12 / peak if tr else None}
        except Exception as ex:     # noqa: BLE001
            next_rows = {"error": str(ex)[:300]}
    large = run_large()
    # ---- CPU baseline on this box's host cores (rank 0, N=1 only), bounded sample.  LAST: the BLAS worker
    #      threads of the oracle keep spinning on every host core for a while after a call, which starves the
    #      thread that launches kernels (AutoRCCSD sweeps measured 21 ms instead of 6.4 ms right behind it)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        t_tr, t_it, e_cpu, how = run_oracle_sample(nbf, NOCC, 1, True)
        info = cpu_info()
        cpu = {"value": flops_step / t_it * 1e-12, "unit": "TFLOP/s",
               "own_executed_tflops": fl.rccsd_iter_ref(o, v) / t_it * 1e-12,
               "cores": info["blas_threads"] or info["host_cores"],
               "kind": "port", "s_per_iteration": t_it, "s_transforms": t_tr,
               "sample": f"oracle (numpy/OpenBLAS port of the reference's literal algorithm): 15 transforms "
                         f"({t_tr:.1f} s) + 1 sweep ({t_it:.1f} s) at nbf={nbf} nocc={NOCC}; value = the sweep's FP64 "
                         f"operations as THIS line counts them (flops_executed_per_sweep, the common unit of work) / CPU "
                         f"seconds; own_executed_tflops = F_ref / s",
               "energy_check": abs(e_cpu - hist[1]) if len(hist) > 1 else None, **info}
    if large:
        ds = [x for x in ((large.get(t) or {}).get(k) for t in ("strong", "c4", "c5")
                          for k in ("max_abs_dE_vs_other_rank_counts", "abs_dE_vs_other_rank_counts")) if x is not None]
        if ds:
            parity["large_invariance"] = {"max_abs_dE": max(ds), "tol": E_TOL,
                                          "status": "ok" if max(ds) <= E_TOL else "FAILED",
                                          "what": "energies of the `large` runs against the same inputs at other rank "
                                                  "counts / the round-1 algorithm (sharding invariance; not an oracle, "
                                                  "does not change the exit code)"}
    F_alg, F_ref = fl.rccsd_iter_alg(o, v), fl.rccsd_iter_ref(o, v)
    s_it = ms_step * 1e-3
    h2d = int(g.nbytes // world + Cao.nbytes + Cav.nbytes + eps.nbytes)
    line = {
        # executed FP64 throughput of the whole job: flops the GEMM kernels of all ranks ran per sweep / s
        "metric": "rccsd_iteration_fp64_tflops", "value": flops_step / s_it * 1e-12,
        "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "s_per_iteration": s_it, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"RCCSD nbf={nbf} nocc={NOCC} nvir={v}"
                               + (" (BASELINE config 3)" if world == 1 else
                                  f" (config 3 grown for weak scaling: F_alg = {F_alg / fl.rccsd_iter_alg(NOCC, 100):.2f} x the 1-GPU workload)")
                               + f", synthetic counter-based ERIs seed={SEED}",
                   "parallelism": f"virtual-index slabs over {world} GPU(s): NCCL all-gathers of the ladder blocks, H and T2 + one "
                                  f"all-reduce per sweep; point-to-point exchange in the transform" if world > 1 else "single GPU",
                   "collectives_per_call": comm,
                   "step": "one RCCSD sweep (intermediates + T1 + T2 + energy) on device-resident data",
                   "l2": "inputs larger than L2 (<vv|vv> >= 0.4 GB per GPU, amplitudes/intermediates >= 32 MB each, re-streamed every sweep)",
                   "flops_executed_per_sweep": flops_step, "F_alg_per_sweep": F_alg, "F_ref_per_sweep": F_ref,
                   "flops_unit_model_per_sweep": fl.rccsd_iter_exec(o, v),
                   "flops_unit_model_rel_err": fl.rccsd_iter_exec(o, v) / flops_step - 1.0,
                   "alg_normalised_tflops": F_alg / s_it * 1e-12, "ref_normalised_tflops": F_ref / s_it * 1e-12,
                   "value_is": "flops_executed_per_sweep (all ranks) / s_per_iteration"},
        "tei_transform": tei,
        "e2e": {"value": c_full["gemm_flops"] / t_call * 1e-12, "unit": "TFLOP/s",
                "flops_executed": c_full["gemm_flops"],
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(8 * (REF_MAXIT + 2)),
                "s_per_do_rccsd": t_call, "s_per_iteration": t_call / REF_MAXIT,
                "alg_normalised_tflops": REF_MAXIT * F_alg / t_call * 1e-12,
                "phases_ms_rank0": e2e_phases,
                "what": "jues_b200_rccsd(host gao, Cao, Cav, eps, maxit=40): H2D of this rank's 1/N of gao + one-pass "
                        "transform + 40 sweeps + energies D2H; value = executed flops (transform + sweeps, all ranks) / s"},
        "gpu_launches": int(round(launches_step)),
        "ms_each_step_rank0": [round(x, 3) for x in it_ms],
        # host clock (ms since the first sweep was issued) at which each sweep had been handed to the
        # driver, and how many sweeps were replayed from a CUDA graph: the launching thread runs ahead
        "host_issue_ms_rank0": [round(ms, 3) for k, ms in ph if k == "cc.iteration.host_ms"],
        "graph_replayed_sweeps": int(sum(ms for k, ms in ph if k == "cc.graph_launches")),
        "roofline": roofline,
        "parity": parity,
        "cpu_baseline": cpu,
        "clocks": clocks,
        "energy": {"E_ccsd_after_timed_sweeps": e, "E_ccsd_40_sweeps_e2e": e_host},
        "integral_transform_ms": tr_ms,
        "large": large,
        "next_rows": next_rows,
    }
    print(json.dumps(line), flush=True)
    return 0 if parity["status"] != "FAILED" else 3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-next-rows", action="store_true", help="skip the AutoRCCSD(T) extra keys")
    ap.add_argument("--no-large", action="store_true", help="skip the nbf=300 / config-4 / config-5 block")
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
        reference_arm(args, rank, world)
        return 0
    if args.warmup < 3:
        args.warmup = 3
    return gpu_arm(args, rank, world)


if __name__ == "__main__":
    sys.exit(main() or 0)
