"""Per-phase cycle counts of the stage-1 kernel of the two-stage tridiagonalisation (CTA 0's view; needs
libtevo_timing.so: `make -C timeevomps.jl_b200/csrc timing`).   python tools/dev_sbr_cycles.py 1024 2048"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import tevo_b200  # registers the package alias
import numpy as np
import timeevomps_jl_b200._lib as L
L.LIB_PATH = os.path.join(os.path.dirname(L.LIB_PATH), "libtevo_timing.so")
from timeevomps_jl_b200._lib import load, check, default_ctx
import tevo_b200 as t
ctx = default_ctx()
lib = load()
lib.tevo_debug_set_heig_mode(2)   # the two-stage reduction is opt-in
names = ["mini-phase", "barrier a", "panel QR (CTA 0)", "barrier b", "pass: tiles", "pass: flush", "barrier c", "-",
         "QR: load panel", "QR: 8 columns", "QR: T factor / taus", "QR: write V, band"]
out = (C.c_ulonglong * 16)()
PD = C.POINTER(C.c_double)
for n in [int(x) for x in sys.argv[1:]]:
    rng = np.random.default_rng(n)
    X = rng.standard_normal((n, n // 2)) + 1j * rng.standard_normal((n, n // 2))
    A = np.asfortranarray(X @ X.conj().T)
    w = np.zeros(n); U = np.zeros((n, n), dtype=np.complex128, order="F")
    for it in range(2):
        lib.tevo_debug_sbr_cycles(out, 1)
        t.profile_enable(True)
        check(lib.tevo_test_heig(ctx.h, n, A.ctypes.data, w.ctypes.data_as(PD), U.ctypes.data), ctx.h)
        prof = t.profile_read()
        lib.tevo_debug_sbr_cycles(out, 0)
    nb = (n - 10) // 8 + 1
    d = {nm: round(out[i] / nb) for i, nm in enumerate(names)}
    print(n, "cycles per block of 8 columns:", d, "total", sum(d.values()))
    print("    ", {kk: round(v["ms"], 3) for kk, v in prof.items()})
