"""First-light check of the sm_100a DGEMM through the C ABI (host operands) + kernel timing."""
import ctypes as C, os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, "jues.jl_b200", "libjues_b200.so"))
dp = C.POINTER(C.c_double)
lib.jues_b200_init.argtypes = [C.POINTER(C.c_void_p), C.c_int]
lib.jues_b200_last_error.restype = C.c_char_p; lib.jues_b200_last_error.argtypes = [C.c_void_p]
lib.jues_b200_dgemm.argtypes = [C.c_void_p, C.c_char, C.c_char, C.c_int64, C.c_int64, C.c_int64, C.c_double, dp, C.c_int64, dp, C.c_int64, C.c_double, dp, C.c_int64]
lib.jues_b200_dgemm_bench.argtypes = [C.c_void_p, C.c_char, C.c_char, C.c_int64, C.c_int64, C.c_int64, C.c_int, dp]
ctx = C.c_void_p()
rc = lib.jues_b200_init(C.byref(ctx), 0)
if rc: print("init failed", rc, lib.jues_b200_last_error(None)); sys.exit(1)
rng = np.random.default_rng(0)
def P(a): return a.ctypes.data_as(dp)
def gemm(tA, tB, M, N, K, alpha=1.0, beta=0.0):
    A = np.asfortranarray(rng.standard_normal((K, M) if tA else (M, K)))
    B = np.asfortranarray(rng.standard_normal((N, K) if tB else (K, N)))
    Cm = np.asfortranarray(rng.standard_normal((M, N)))
    ref = alpha * ((A.T if tA else A) @ (B.T if tB else B)) + beta * Cm
    rc = lib.jues_b200_dgemm(ctx, b'T' if tA else b'N', b'T' if tB else b'N', M, N, K, alpha, P(A), A.shape[0], P(B), B.shape[0], beta, P(Cm), M)
    if rc: return None, lib.jues_b200_last_error(ctx).decode()
    return np.abs(Cm - ref).max() / max(1.0, np.abs(ref).max()), None
bad = 0
for cfg in ["-1", "0", "1", "2", "3"]:
    os.environ["JUES_B200_GEMM_CFG"] = cfg
    for (M, N, K) in [(128, 128, 64), (7, 5, 3), (25, 361, 361), (130, 250, 17), (400, 1000, 333), (64, 64, 16), (1, 1, 1), (257, 129, 100)]:
        for tA in (False, True):
            for tB in (False, True):
                for (al, be) in [(1.0, 0.0), (-0.5, 2.0)]:
                    err, msg = gemm(tA, tB, M, N, K, al, be)
                    ok = err is not None and err < 1e-12
                    if not ok:
                        bad += 1
                        print("FAIL cfg", cfg, M, N, K, tA, tB, al, be, err, msg, flush=True)
    print("cfg", cfg, "done, failures so far", bad, flush=True)
os.environ["JUES_B200_GEMM_CFG"] = "-1"
print("GEMM_PARITY", "OK" if bad == 0 else f"{bad} FAILURES", flush=True)
res = {}
ms = C.c_double()
for name, (tA, tB, M, N, K) in {
    "nn_8192": (b'N', b'N', 8192, 8192, 8192), "tn_8192": (b'T', b'N', 8192, 8192, 8192),
    "nt_8192": (b'N', b'T', 8192, 8192, 8192), "tt_8192": (b'T', b'T', 8192, 8192, 8192),
    "ladder_c3": (b'N', b'N', 400, 10000, 10000), "ladder_like": (b'N', b'N', 3600, 16384, 16384),
    "q1_like": (b'N', b'N', 460 * 460 * 16, 460, 460), "ring_c3": (b'N', b'N', 2000, 2000, 2000),
}.items():
    for cfg in ["0", "1", "2", "3"]:
        os.environ["JUES_B200_GEMM_CFG"] = cfg
        rc = lib.jues_b200_dgemm_bench(ctx, tA, tB, M, N, K, 3, C.byref(ms))
        if rc: print(name, cfg, "ERR", lib.jues_b200_last_error(ctx)); continue
        tf = 2.0 * M * N * K / ms.value * 1e-9
        res[f"{name}_cfg{cfg}"] = {"ms": ms.value, "tflops": tf}
        print(f"{name:14s} cfg{cfg} {ms.value:9.3f} ms  {tf:7.2f} TFLOP/s", flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "gemm_first_light.json"), "w"), indent=1)
