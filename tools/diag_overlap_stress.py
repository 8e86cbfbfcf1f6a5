"""Determinism stress of the RCCSD sweep (side branch on the second stream): repeated 40-sweep runs must give
bit-identical energy traces and amplitudes, and match the committed oracle trace.
python tools/diag_overlap_stress.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jues.jl_b200 as jb
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
bad = 0
for (N, o) in [(120, 20), (60, 10), (33, 7)]:
    ctx = jb.Context(0)
    Cao, Cav, eps = jb.synth.orbitals(N, o, 2024)
    g = jb.DeviceFourTensor.synth_eri(N, seed=2024, scale=jb.synth.counter_scale(N), ctx=ctx)
    w = jb.Wfn(o, N - o, eps, Cao, Cav, g)
    ref_h, ref_T = None, None
    for r in range(reps):
        h = []
        e, T1, T2 = jb.RCCSD.do_rccsd(w, ctx=ctx, _return_T=True, _e_hist=h)
        h = np.asarray(h)
        if ref_h is None:
            ref_h, ref_T = h, (T1.copy(), T2.copy())
        else:
            same = np.array_equal(h, ref_h) and np.array_equal(T1, ref_T[0]) and np.array_equal(T2, ref_T[1])
            if not same:
                bad += 1
                print(f"N={N} rep {r}: DIFFERS  max|dE|={np.abs(h - ref_h).max():.3e} max|dT2|={np.abs(T2 - ref_T[1]).max():.3e}")
    gp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", f"bench_ehist_nbf{N}_nocc{o}.npz")
    if os.path.exists(gp):
        gold = np.load(gp)["e_hist"]
        print(f"N={N}: max|dE| vs oracle trace {np.abs(ref_h - gold[:len(ref_h)]).max():.3e}")
    print(f"N={N} o={o}: {reps} runs, phases graph launches {[ms for k, ms in ctx.phases() if k == 'cc.graph_launches']}")
    g.free(); ctx.close()
print("STRESS", "OK" if bad == 0 else f"FAILED ({bad})")
