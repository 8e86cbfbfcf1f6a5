"""Why is AutoRCCSD(T) slow inside bench.py (next_rows)?  Replays the bench sequence and dumps the phases."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jues.jl_b200 as jb
N, o = 120, 20
ctx = jb.Context(0)
Cao, Cav, eps = jb.synth.orbitals(N, o, 2024)
gdev = jb.DeviceFourTensor.synth_eri(N, seed=2024, scale=jb.synth.counter_scale(N), ctx=ctx)
keep = torch.empty(N ** 4, dtype=torch.float64).pin_memory()
g = keep.numpy().reshape((N,) * 4, order="F")
for s0 in range(0, N, 16):
    g[:, :, :, s0:min(N, s0 + 16)] = gdev[:, :, :, s0:min(N, s0 + 16)]
wdev = jb.Wfn(o, N - o, eps, Cao, Cav, gdev)
whost = jb.Wfn(o, N - o, eps, Cao, Cav, g)
if "rccsd_first" in sys.argv:
    jb.RCCSD.do_rccsd(wdev, ctx=ctx, _maxit=13)
    jb.RCCSD.do_rccsd(whost, ctx=ctx)
hao = jb.synth.core_hamiltonian(g, Cao, Cav, eps)
wa = jb.Wfn(o, N - o, eps, Cao, Cav, g, hao=hao)
for rep in range(2):
    t0 = time.perf_counter()
    r = jb.AutoRCCSD.do_rccsd(wa, ctx=ctx, do_pT=True, _return_all=True)
    dt = time.perf_counter() - t0
    ph = ctx.phases()
    agg = {}
    for k, ms in ph:
        agg.setdefault(k, []).append(ms)
    print("rep", rep, "wall", round(dt, 3), "iters", r["iterations"], {k: (round(sum(v), 2), len(v)) for k, v in agg.items() if not k.startswith("gemm")})
