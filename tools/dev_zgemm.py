"""Device-resident throughput of the DMMA ZGEMM at the hot-path shapes (CUDA events, no host copies), next to the
measured cuBLAS ZGEMM peak (profiles/r01_fp64_peaks.json / profiles/r02_comparators.json).
    python tools/dev_zgemm.py            [TEVO_ZGEMM_CTSIGN=1 python tools/dev_zgemm.py  for the compile-time-sign variant]"""
import os, sys, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tevo_b200 as t
from timeevomps_jl_b200._lib import load, check, default_ctx
ctx = default_ctx()
lib = load()
shapes = [  # name, kind, ta, tb, M, N, K, accumulate
    ("theta NN 2048x2048x1024", 0, 0, 0, 2048, 2048, 1024, 0),
    ("R=U^H theta CN 1024x2048x2048", 0, 2, 0, 1024, 2048, 2048, 0),
    ("rho herm NC 2048x2048x2048", 1, 0, 2, 2048, 2048, 2048, 0),
    ("rank-2k herm NC 1920x1920x128 (+C)", 1, 0, 2, 1920, 1920, 128, 1),
    ("back-transform NN 2048x1024x128 (+C)", 0, 0, 0, 2048, 1024, 128, 1),
    ("square 4096", 0, 0, 0, 4096, 4096, 4096, 0),
]
out = {"ctsign": bool(os.environ.get("TEVO_ZGEMM_CTSIGN"))}
for (name, kind, ta, tb, M, N, K, acc) in shapes:
    ms = C.c_double()
    check(lib.tevo_test_zgemm_time(ctx.h, kind, ta, tb, M, N, K, acc, 20, C.byref(ms)), ctx.h)
    flops = 8.0 * M * N * K * (0.5 if kind == 1 else 1.0)
    out[name] = {"ms": round(ms.value, 4), "tflops": round(flops / (ms.value * 1e-3) / 1e12, 2)}
print(json.dumps(out))
