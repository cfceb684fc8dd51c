"""Library comparators (context only, never on the product path): cuSOLVER Zheevd and cuBLAS ZGEMM through torch,
timed with CUDA events beside libtevo's own eigensolver / ZGEMM at the hot-path shapes.
    python tools/comparators.py > gpurun_out/comparators.json"""
import json, os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tevo_b200 as t
from timeevomps_jl_b200._lib import load, check, default_ctx

dev = torch.device("cuda", 0)
out = {"device": torch.cuda.get_device_name(0)}


def ev_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(min(ts))


# --- cuSOLVER Hermitian eigensolver (torch.linalg.eigh -> Zheevd / syevj batched as torch chooses) ---
for n in (512, 1024, 2048):
    g = torch.Generator(device="cpu").manual_seed(n)
    X = torch.randn(n, n // 2, dtype=torch.complex128, generator=g).to(dev)
    A = X @ X.conj().T
    med, best = ev_time(lambda: torch.linalg.eigh(A), reps=3, warm=1)
    out["cusolver_zheevd_ms_n%d" % n] = {"median": med, "best": best}

# --- cuBLAS ZGEMM at the hot-path shapes ---
shapes = {"theta_2048x2048x1024": (2048, 2048, 1024), "rho_2048x2048x2048": (2048, 2048, 2048),
          "her2k_1920x1920x128": (1920, 1920, 128), "back_2048x1024x128": (2048, 1024, 128),
          "square_4096": (4096, 4096, 4096)}
for name, (M, N, K) in shapes.items():
    A = torch.randn(M, K, dtype=torch.complex128, device=dev)
    B = torch.randn(K, N, dtype=torch.complex128, device=dev)
    med, best = ev_time(lambda: torch.matmul(A, B), reps=7, warm=3)
    out["cublas_zgemm_" + name] = {"ms": med, "tflops": 8.0 * M * N * K / (med * 1e-3) / 1e12,
                                   "tflops_best": 8.0 * M * N * K / (best * 1e-3) / 1e12}

# --- libtevo ZGEMM (device-resident timing through the test hook is host-buffer based; use the profile of heig instead) ---
ctx = default_ctx()
lib = load()
for n in (512, 1024, 2048):
    rng = np.random.default_rng(n)
    X = rng.standard_normal((n, n // 2)) + 1j * rng.standard_normal((n, n // 2))
    A = np.asfortranarray(X @ X.conj().T)
    w = np.zeros(n); U = np.zeros((n, n), dtype=np.complex128, order="F")
    for it in range(3):
        t.profile_enable(True)
        check(lib.tevo_test_heig(ctx.h, n, A.ctypes.data, w.ctypes.data_as(C.POINTER(C.c_double)), U.ctypes.data), ctx.h)
        prof = t.profile_read()
    out["libtevo_heig_ms_n%d" % n] = {"total": sum(v["ms"] for v in prof.values()),
                                      "phases": {k: round(v["ms"], 3) for k, v in prof.items()}}
print(json.dumps(out, indent=1))
