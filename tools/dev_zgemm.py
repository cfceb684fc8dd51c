import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tevo_b200 as t
from timeevomps_jl_b200._lib import load, check, default_ctx, tevo_complex
ctx = default_ctx()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rng = np.random.default_rng(0)
A = np.asfortranarray(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
B = np.asfortranarray(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
Cm = np.zeros((n, n), dtype=np.complex128, order="F")
for ta, tb in ((0, 0), (2, 0), (0, 2)):
    t0 = time.time()
    check(load().tevo_test_zgemm(ctx.h, ta, tb, n, n, n, tevo_complex(1, 0), A.ctypes.data, n, B.ctypes.data, n, tevo_complex(0, 0), Cm.ctypes.data, n), ctx.h)
print("done")
