"""TFLOP/s of the DGEMM at the shapes that dominate the path (isolated launches, CUDA events)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jues.jl_b200 as jb
ctx = jb.Context(0)
shapes = [("ring C3", "T", "N", 2000, 2000, 2000), ("packed ladder C3", "N", "N", 400, 5100, 5050),
          ("packed ladder N=2", "N", "N", 400, 3906, 7750), ("Q1 slab", "N", "N", 144 ** 3, 62, 144),
          ("Q2 slab", "T", "N", 124, 144 * 144 * 62, 144), ("8192^3", "N", "N", 8192, 8192, 8192),
          ("dense ladder C5/GPU", "N", "N", 3600, 20000, 20000), ("hh ladder C3", "T", "N", 400, 10000, 400),
          ("Fae-like", "T", "N", 100, 100, 40000)]
for name, tA, tB, M, N, K in shapes:
    ms = ctx.gemm_bench(tA, tB, M, N, K, reps=5)
    print(json.dumps({"gemm": name, "M": M, "N": N, "K": K, "ms": round(ms, 4), "tflops": round(2.0 * M * N * K / ms * 1e-9, 2)}), flush=True)
