"""Full tei_transform(gao, C) (8 N^5 flop) on a device-resident AO tensor at several sizes."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jues.jl_b200 as jb
ctx = jb.Context(0)
res = {}
for N in [int(x) for x in sys.argv[1:]] or [120, 200]:
    g = jb.DeviceFourTensor.synth_eri(N, seed=1, ctx=ctx)
    Cao, Cav, eps = jb.synth.orbitals(N, N // 6, 1)
    C = np.asfortranarray(np.hstack([Cao, Cav]))
    for rep in range(2):
        out = jb.tei_transform(g, C, "bench", ctx=ctx); out.free()
    ms = [m for k, m in ctx.phases() if k == "tei.transform"][0]
    fl = ctx.counters()["gemm_flops"]
    res[N] = {"ms": ms, "tflops": fl / ms * 1e-9, "flops": fl}
    print(N, res[N], flush=True)
    g.free()
json.dump(res, open(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "tei_bench.json"), "w"), indent=1)
