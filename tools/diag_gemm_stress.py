"""1-GPU determinism stress of the DGEMM at the quarter-transform shapes of a (np, v, vs) slab transform."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jues.jl_b200 as jb
np_, v, vs = [int(x) for x in os.environ.get("DIAG_SHAPE", "144,124,62").split(",")]
reps = int(os.environ.get("DIAG_REPS", "30"))
ctx = jb.Context(0)
shapes = [("Q1 axis3", "N", "N", np_ ** 3, vs, np_, 1), ("Q2 axis0", "T", "N", v, np_ * np_ * vs, np_, 1),
          ("Q3 axis1", "N", "N", v, v, np_, np_ * vs), ("Q4 axis2", "N", "N", v * v, v, np_, vs),
          ("ring", "T", "N", 2000, 2000, 2000, 1), ("ladder", "N", "N", 400, 3906, 7750, 1)]
for name, tA, tB, M, N, K, b in shapes:
    nb, w = ctx.gemm_stress(tA, tB, M, N, K, b, reps)
    print(json.dumps({"gemm": name, "tA": tA, "tB": tB, "M": M, "N": N, "K": K, "batch": b, "reps": reps,
                      "n_bad": nb, "worst_sqdiff": w}), flush=True)
