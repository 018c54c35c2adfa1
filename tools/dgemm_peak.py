"""cuBLAS DGEMM throughput on this box (roofline denominator for the FP64 kernels).

Burst = best of 10 single 8192^3 float64 matmuls; sustained = back-to-back for ~3 s.
Prints one JSON line. Library call used only as a yard-stick, never on the product path.
"""
import json
import os
import torch


def main():
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    flops = 2.0 * n ** 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cnt = 0
    while True:
        for _ in range(5):
            torch.matmul(a, b, out=c)
        cnt += 5
        e1.record()
        e1.synchronize()
        if e0.elapsed_time(e1) > 3000:
            break
    sustained = flops * cnt / (e0.elapsed_time(e1) * 1e-3) / 1e12
    # tall-skinny shape like the synthesis stage-2 contraction (M=86400, K=193, N=720)
    a2 = torch.randn(86400, 193, dtype=torch.float64, device="cuda")
    b2 = torch.randn(193, 720, dtype=torch.float64, device="cuda")
    c2 = torch.empty(86400, 720, dtype=torch.float64, device="cuda")
    for _ in range(3):
        torch.matmul(a2, b2, out=c2)
    torch.cuda.synchronize()
    best2 = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a2, b2, out=c2)
        e1.record()
        e1.synchronize()
        best2 = min(best2, e0.elapsed_time(e1))
    print(json.dumps({
        "dgemm_8192_burst_tflops": flops / (best * 1e-3) / 1e12,
        "dgemm_8192_sustained_tflops": sustained,
        "dgemm_86400x193x720_ms": best2,
        "dgemm_86400x193x720_tflops": 2.0 * 86400 * 193 * 720 / (best2 * 1e-3) / 1e12,
        "host_cpus": os.cpu_count(),
    }))


if __name__ == "__main__":
    main()
