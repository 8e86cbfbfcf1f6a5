"""Short driver for ncu: RCCSD sweeps on device-resident inputs at the bench shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jues.jl_b200 as jb
N, o = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (120, 20)
its = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ctx = jb.Context(0)
scale = jb.synth.counter_scale(N)
Cao, Cav, eps = jb.synth.orbitals(N, o, 2024)
g = jb.DeviceFourTensor.synth_eri(N, seed=2024, scale=scale, ctx=ctx)
w = jb.Wfn(o, N - o, eps, Cao, Cav, g)
print(jb.RCCSD.do_rccsd(w, ctx=ctx, _maxit=its))
