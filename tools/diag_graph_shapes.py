"""Does the CUDA-graph replay of the sweep engage at the bench shapes of every rank count (run on ONE GPU:
the allocation pattern -- temporaries of 64 MB and more at nbf >= 172 -- is what matters)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jues.jl_b200 as jb
for N in (120, 144, 172, 212):
    ctx = jb.Context(0)
    Cao, Cav, eps = jb.synth.orbitals(N, 20, 2024)
    g = jb.DeviceFourTensor.synth_eri(N, seed=2024, scale=jb.synth.counter_scale(N), ctx=ctx)
    w = jb.Wfn(20, N - 20, eps, Cao, Cav, g)
    h = []
    jb.RCCSD.do_rccsd(w, ctx=ctx, _maxit=14, _e_hist=h)
    ph = dict((k, ms) for k, ms in ctx.phases() if k in ("cc.graph_launches", "cc.arena_mb", "cc.arena_misses"))
    it = [ms for k, ms in ctx.phases() if k == "cc.iteration"]
    gold = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", f"bench_ehist_nbf{N}_nocc20.npz"))["e_hist"]
    print(N, ph, "median sweep ms", round(float(np.median(it[3:])), 3), "max|dE|", float(np.abs(np.asarray(h) - gold[:len(h)]).max()))
    g.free(); ctx.close()
