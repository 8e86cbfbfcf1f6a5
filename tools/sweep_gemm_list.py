"""Per-launch list of the DGEMMs of one RCCSD sweep with their in-sweep durations (trace level 2):
shape, ms, TFLOP/s, share of the sweep.  python tools/sweep_gemm_list.py [nbf nocc]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jues.jl_b200 as jb
N, o = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (120, 20)
ctx = jb.Context(0)
Cao, Cav, eps = jb.synth.orbitals(N, o, 2024)
g = (jb.DeviceFourTensor.synth_eri(N, seed=2024, ctx=ctx, virtual=True) if N > 160 else
     jb.DeviceFourTensor.synth_eri(N, seed=2024, scale=jb.synth.counter_scale(N), ctx=ctx))
w = jb.Wfn(o, N - o, eps, Cao, Cav, g)
if N <= 160:
    jb.RCCSD.do_rccsd(w, ctx=ctx, _maxit=2)
ctx.set_trace(2)
jb.RCCSD.do_rccsd(w, ctx=ctx, _maxit=4 if N <= 160 else 3)
ph = ctx.phases()
ctx.set_trace(0)
sweeps, cur = [], []
for k, ms in ph:
    if k == "cc.iteration":
        sweeps.append((ms, cur)); cur = []
    elif k.startswith("gemm ") or k.startswith("perm ") or k.startswith("aux ") or k.startswith("cc.part") or k.startswith("cc.comm"):
        cur.append((k, ms))
ms_sweep, items = sweeps[-1]
tot = 0.0
ptot = 0.0
atot = 0.0
for k, ms in items:
    if k.startswith("gemm "):
        M, Nn, K, B = (int(x) for x in k.split()[1].split("x"))
        fl = 2.0 * M * Nn * K * B
        tot += ms
        print(f"{ms*1e3:9.1f} us  {fl/ms*1e-9:6.2f} TF/s  {100*ms/ms_sweep:5.1f}%  {k}")
    elif k.startswith("perm "):
        ptot += ms
        dims = [int(x) for x in k.split()[2].rstrip("+").split("x")]
        mb = 8e-6 * np.prod(dims) * (3 if k.endswith("+") else 2)
        print(f"{ms*1e3:9.1f} us  permute {mb/ms*1e-3:5.2f} TB/s {100*ms/ms_sweep:5.1f}%  {k}")
    elif k.startswith("aux "):
        atot += ms
        mb = float(k.split()[2][:-2])
        print(f"{ms*1e3:9.1f} us  aux     {mb/ms*1e-3:5.2f} TB/s {100*ms/ms_sweep:5.1f}%  {k}")
    else:
        print(f"{ms*1e3:9.1f} us  ------ {k}")
print(f"sweep {ms_sweep:.3f} ms (traced, eager), gemm sum {tot:.3f} ms, permute sum {ptot:.3f} ms, other element-wise {atot:.3f} ms")
