"""Exhaustive (tile configuration x split-K) timing of the DGEMM at the skinny / small shapes of a BASELINE
config-3 sweep: what the kernel can do at best for each shape vs what the launcher's model picks.
python tools/gemm_autotune.py > gpurun_out/gemm_autotune.json"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jues.jl_b200 as jb
ctx = jb.Context(0)
NCFG = 10
shapes = [("N", "T", 100, 100, 40000), ("N", "T", 20, 20, 200000), ("T", "N", 400, 400, 10000), ("T", "N", 10000, 1, 2000),
          ("N", "T", 20, 100, 40000), ("N", "T", 20, 100, 200000), ("T", "N", 2000, 1, 2000), ("N", "N", 8000, 20, 100),
          ("N", "N", 200000, 20, 100), ("N", "N", 40000, 100, 20), ("T", "N", 400, 10000, 400), ("N", "N", 40000, 100, 100),
          ("N", "N", 200000, 20, 20), ("T", "N", 400, 2000, 10000), ("N", "N", 20, 40000, 100), ("N", "N", 20, 200000, 100),
          ("N", "N", 100, 40000, 20), ("T", "N", 20, 200000, 20), ("T", "N", 40000, 100, 20), ("T", "T", 200000, 20, 100)]
if len(sys.argv) > 1 and sys.argv[1] == "skinny":
    shapes = [s for s in shapes if min(s[2], s[3]) <= 20 or s[4] <= 20]
out = []
for tA, tB, M, N, K in shapes:
    os.environ.pop("JUES_B200_GEMM_CFG", None)
    base = ctx.gemm_bench(tA, tB, M, N, K, reps=5)
    best = (base, "model")
    table = {}
    KT = (K + 15) // 16
    for cfg in range(NCFG):
        for sp in (1, 2, 4, 8, 16, 32, 64, 128, 256):
            if sp > max(1, KT // 2):
                continue
            os.environ["JUES_B200_GEMM_CFG"] = str(cfg + NCFG * (sp - 1))
            try:
                ms = ctx.gemm_bench(tA, tB, M, N, K, reps=3)
            except Exception as ex:       # noqa: BLE001
                continue
            table[f"{cfg}/{sp}"] = round(ms * 1e3, 1)
            if ms < best[0]:
                best = (ms, f"cfg{cfg} split{sp}")
    os.environ.pop("JUES_B200_GEMM_CFG", None)
    rec = {"tA": tA, "tB": tB, "M": M, "N": N, "K": K, "model_us": round(base * 1e3, 1), "best_us": round(best[0] * 1e3, 1),
           "best": best[1], "bytes_MB": round(8e-6 * (M * K + K * N + M * N), 1),
           "hbm_floor_us": round(8.0 * (M * K + K * N + M * N) / 6.0e6, 1), "table_us": table}
    out.append(rec)
    print(json.dumps({k: v for k, v in rec.items() if k != "table_us"}), flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "gemm_autotune_full.json"), "w"))
