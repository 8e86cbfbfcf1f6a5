"""1-GPU determinism stress of the slab transform, quarter by quarter, on the device."""
import os, sys, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jues.jl_b200 as jb
from jues.jl_b200 import _p, _f
nbf, nocc, P = [int(x) for x in os.environ.get("DIAG_SHAPE", "144,20,2").split(",")]
reps = int(os.environ.get("DIAG_REPS", "12"))
v = nbf - nocc
ctx = jb.Context(0)
Cao, Cav, eps = jb.synth.orbitals(nbf, nocc, 2024)
g = jb.DeviceFourTensor.synth_eri(nbf, seed=2024, scale=jb.synth.counter_scale(nbf), ctx=ctx)
vs = v // P
Cv = _f(Cav)
for r in range(int(os.environ.get("DIAG_SLABS", P))):
    Cs = np.asfortranarray(Cav[:, r * vs:(r + 1) * vs])
    st = (C.c_double * (16 * reps))()
    ctx._check(ctx._lib.jues_b200_transform_stress(ctx._h, g._h, _p(Cv), v, _p(Cv), v, _p(Cv), v, _p(Cs), vs, reps, st))
    a = np.array(list(st)).reshape(reps, 4, 4)
    for k in range(reps):
        print(json.dumps({"slab": r, "rep": k, "q": [[int(x[0]), int(x[1]), int(x[2]), float(x[3])] for x in a[k]]}), flush=True)
