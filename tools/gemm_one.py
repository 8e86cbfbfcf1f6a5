"""Short driver for ncu: a few launches of the DGEMM at one shape (device-resident operands)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jues.jl_b200 as jb
tA, tB, M, N, K = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
ctx = jb.Context(0)
ms = ctx.gemm_bench(tA, tB, M, N, K, reps=2)
print("ms", ms, "TFLOP/s", 2.0 * M * N * K / ms * 1e-9)
