"""Per-phase cycle counts of the tridiagonalisation panel kernel (needs libtevo_timing.so built with
-DTEVO_PANEL_TIMING: `make -C timeevomps.jl_b200/csrc timing`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import tevo_b200  # registers the package alias
import numpy as np
import timeevomps_jl_b200._lib as L
L.LIB_PATH = os.path.join(os.path.dirname(L.LIB_PATH), "libtevo_timing%s.so" % os.environ.get("TEVO_TIMING_VARIANT", ""))
from timeevomps_jl_b200._lib import load, check, default_ctx
ctx = default_ctx()
lib = load()
names = ["A:fetch", "A:scalars", "A:rows", "A:norm", "barrier1", "B:loads", "B:scalars+scale", "B:hemv+dots", "B:dots-atomics", "B:yv", "barrier2"]
out = (C.c_ulonglong * 16)()
for n in [int(x) for x in sys.argv[1:]]:
    rng = np.random.default_rng(n)
    X = rng.standard_normal((n, n // 2)) + 1j * rng.standard_normal((n, n // 2))
    A = np.asfortranarray(X @ X.conj().T)
    w = np.zeros(n); U = np.zeros((n, n), dtype=np.complex128, order="F")
    for it in range(2):
        lib.tevo_debug_panel_cycles(out, 1)
        check(lib.tevo_test_heig(ctx.h, n, A.ctypes.data, w.ctypes.data_as(C.POINTER(C.c_double)), U.ctypes.data), ctx.h)
        lib.tevo_debug_panel_cycles(out, 0)
    ncol = n - 1
    d = {nm: round(out[i] / ncol) for i, nm in enumerate(names)}
    print(n, d, "total", sum(d.values()))
