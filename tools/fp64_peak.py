"""Calibrates the FP64 denominators MEASURED_PEAKS.json lacks: cuBLAS DGEMM / ZGEMM throughput on this B200
(torch.matmul float64 / complex128, 8192^3 and 4096^3), best of 10 with CUDA events.  Writes gpurun_out/fp64_peaks.json."""
import json
import os
import sys

import torch


def bench(n, dtype, reps=10):
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    for _ in range(3):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    fl = 2.0 * n ** 3 * (4 if dtype.is_complex else 1)
    return fl / (best * 1e-3) / 1e12, best


out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
for n in (4096, 8192):
    t, ms = bench(n, torch.float64)
    out[f"dgemm_{n}_tflops"] = t
    out[f"dgemm_{n}_ms"] = ms
    t, ms = bench(n, torch.complex128, reps=5)
    out[f"zgemm_{n}_tflops"] = t
    out[f"zgemm_{n}_ms"] = ms
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fp64_peaks.json", "w"), indent=1)
print(json.dumps(out))
