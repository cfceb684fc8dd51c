"""Measures the FP64 GEMM denominators MEASURED_PEAKS.json lacks (cuBLAS DGEMM / ZGEMM via torch).
Library comparator only - not part of the product path."""
import json, sys, time
import torch
out = {}
dev = torch.device("cuda:0")
for name, dt, n, flops_per in (("dgemm", torch.float64, 8192, 2), ("zgemm", torch.complex128, 4096, 8)):
    a = torch.randn(n, n, dtype=dt, device=dev); b = torch.randn(n, n, dtype=dt, device=dev)
    for _ in range(3): c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[name + "_tflops_burst"] = flops_per * n ** 3 / best / 1e9
    t0 = time.time(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); it = 0
    while time.time() - t0 < 3.0:
        c = a @ b; it += 1
        if it % 8 == 0: torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    out[name + "_tflops_sustained"] = flops_per * n ** 3 * it / e0.elapsed_time(e1) / 1e9
    out[name + "_n"] = n
print(json.dumps(out))
