"""Device-resident time of the DMMA ZGEMM at SMALL shapes for each CTA tile (csrc/zgemm.cu, pick_tile): the panel /
trailing GEMMs of the blocked Cholesky chain of the centre moves, and the chi = 256 shapes of BASELINE config C3.
    python tools/dev_zgemm_tiles.py"""
import os, sys, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tevo_b200 as t  # noqa: F401
from timeevomps_jl_b200._lib import load, check, default_ctx
ctx = default_ctx()
lib = load()
shapes = [  # name, kind, ta, tb, M, N, K, accumulate
    ("chol panel NC 960x64x64", 0, 0, 2, 960, 64, 64, 0),
    ("chol trailing herm NC 960x960x64 (+C)", 1, 0, 2, 960, 960, 64, 1),
    ("chol panel NC 448x64x64", 0, 0, 2, 448, 64, 64, 0),
    ("chol trailing herm NC 448x448x64 (+C)", 1, 0, 2, 448, 448, 64, 1),
    ("inverse merge NN 256x256x256", 0, 0, 0, 256, 256, 256, 0),
    ("C3 theta NN 512x512x256", 0, 0, 0, 512, 512, 256, 0),
    ("C3 rho herm NC 512x512x512", 1, 0, 2, 512, 512, 512, 0),
    ("C3 R=U^H theta CN 256x512x512", 0, 2, 0, 256, 512, 512, 0),
    ("hop NN 1024x2048x1024", 0, 0, 0, 1024, 2048, 1024, 0),
    ("gram herm CN 1024x1024x2048", 1, 2, 0, 1024, 1024, 2048, 0),
    ("theta NN 2048x2048x1024", 0, 0, 0, 2048, 2048, 1024, 0),
]
names = {-1: "auto", 0: "128x64", 1: "64x32", 2: "32x32"}
out = {}
for (name, kind, ta, tb, M, N, K, acc) in shapes:
    row = {}
    for tile in (-1, 0, 1, 2):
        if tile == 2 and M * N >= 1024 * 1024:
            continue
        assert lib.tevo_debug_set_zgemm_tile(tile) == 0
        ms = C.c_double()
        check(lib.tevo_test_zgemm_time(ctx.h, kind, ta, tb, M, N, K, acc, 50, C.byref(ms)), ctx.h)
        flops = 8.0 * M * N * K * (0.5 if kind == 1 else 1.0)
        row[names[tile]] = {"us": round(ms.value * 1e3, 1), "tflops": round(flops / (ms.value * 1e-3) / 1e12, 2)}
    lib.tevo_debug_set_zgemm_tile(-1)
    out[name] = row
    print(name, row, flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "zgemm_tiles.json"), "w"), indent=1)
