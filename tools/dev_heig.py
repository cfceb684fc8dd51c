import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sys, time, ctypes as C
import numpy as np
import tevo_b200 as t
from timeevomps_jl_b200._lib import load, check, default_ctx
ctx = default_ctx()
for n in [int(x) for x in sys.argv[1:]]:
    rng = np.random.default_rng(n)
    X = (rng.standard_normal((n, n // 2)) + 1j * rng.standard_normal((n, n // 2)))
    X = X @ X.conj().T
    A = np.asfortranarray(X)
    w = np.zeros(n); U = np.zeros((n, n), dtype=np.complex128, order="F")
    for it in range(3):
        t.profile_enable(True)
        check(load().tevo_test_heig(ctx.h, n, A.ctypes.data, w.ctypes.data_as(C.POINTER(C.c_double)), U.ctypes.data), ctx.h)
        prof = t.profile_read()
    print(n, {k: (round(v["ms"], 2), v["count"]) for k, v in prof.items()}, "sum", round(sum(v["ms"] for v in prof.values()), 2))
    wr = np.linalg.eigvalsh(X)[::-1]
    print("   eig err", np.abs(w - wr).max() / wr[0], "resid", np.abs(X @ U - U * w).max() / wr[0])
