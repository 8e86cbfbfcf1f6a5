"""Measure the FP64 GEMM peak of the box with cuBLAS (torch.matmul float64): the roofline
denominator for every FP64 tensor-pipe fraction this repo reports."""
import json, sys, time
import torch

def bench(M, N, K, reps=5):
    a = torch.randn(M, K, dtype=torch.float64, device="cuda")
    b = torch.randn(K, N, dtype=torch.float64, device="cuda")
    torch.matmul(a, b); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    # sustained: back to back for ~3 s
    n = max(3, int(3000 / best)); 
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): c = torch.matmul(a, b)
    e1.record(); torch.cuda.synchronize()
    sus = e0.elapsed_time(e1) / n
    fl = 2.0 * M * N * K
    return fl / best * 1e-9, fl / sus * 1e-9

if __name__ == "__main__":
    out = {"gpu": torch.cuda.get_device_name(0)}
    for (M, N, K) in [(8192, 8192, 8192), (16384, 16384, 4096), (4096, 4096, 4096), (3600, 16384, 16384)]:
        burst, sus = bench(M, N, K)
        out[f"dgemm_{M}x{N}x{K}"] = {"burst_tflops": round(burst, 2), "sustained_tflops": round(sus, 2)}
        print(M, N, K, burst, sus, flush=True)
    out["fp64_tflops"] = max(v["burst_tflops"] for k, v in out.items() if k.startswith("dgemm"))
    out["fp64_tflops_sustained"] = max(v["sustained_tflops"] for k, v in out.items() if k.startswith("dgemm"))
    json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/fp64_peak.json", "w"), indent=1)
    print(json.dumps(out))
