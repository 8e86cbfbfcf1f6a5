"""Stand-alone (T) run for ncu: jues_b200_compute_pt on random amplitudes/integrals of a given
shape (the launch sequence does not depend on the numbers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jues.jl_b200 as jb

o, v = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(1)
r = lambda *s: np.asfortranarray(rng.standard_normal(s) * 0.05)
kw = dict(T1=r(o, v), T2=r(o, o, v, v), Vvvvo=r(v, v, v, o), Vvooo=r(v, o, o, o), Vvovo=r(v, o, v, o),
          fo=np.linspace(-2.0, -0.6, o), fv=np.linspace(0.4, 2.5, v))
ctx = jb.Context(0)
e = jb.compute_pT(ctx=ctx, **kw)
print("E(T) =", e, dict(ctx.phases()), ctx.counters())
ctx.close()
