"""Time the phases of do_rccsd / do_rmp2 at a given shape on the device (development helper)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jues.jl_b200 as jb
N, o = int(sys.argv[1]), int(sys.argv[2]); maxit = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g, Cao, Cav, eps = jb.synth.dense_inputs(N, o, seed=2024)
w = jb.Wfn(o, N - o, eps, Cao, Cav, g)
ctx = jb.Context(0)
for name, fn in [("rmp2", lambda: jb.do_rmp2(w, ctx=ctx)), ("rccsd", lambda: jb.RCCSD.do_rccsd(w, ctx=ctx, _maxit=maxit)),
                 ("rccd", lambda: jb.RCCD.do_rccd(w, ctx=ctx, _maxit=maxit))]:
    for rep in range(2):
        t = time.time(); e = fn(); dt = time.time() - t
    ph = ctx.phases(); c = ctx.counters()
    agg = {}
    for k, ms in ph: agg.setdefault(k, []).append(ms)
    print(name, "E=%.12f wall=%.3fs" % (e, dt), {k: (len(v), round(sum(v) / len(v), 3)) for k, v in agg.items()},
          "gemm_tflop=%.3f launches=%d/%d peakGB=%.2f" % (c["gemm_flops"] / 1e12, c["gemm_launches"], c["aux_launches"], c["bytes_peak"] / 1e9), flush=True)
    if "cc.iteration" in agg:
        it_ms = min(agg["cc.iteration"])
        print("   best iteration %.3f ms" % it_ms)
