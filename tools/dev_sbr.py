"""Stage-by-stage check of the two-stage tridiagonalisation (csrc/sbr.cu) against proto/two_stage.py and LAPACK.
    python tools/dev_sbr.py 40 96 257 1024 2048"""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "proto"))
import numpy as np
import tevo_b200 as t
from timeevomps_jl_b200._lib import load, check, default_ctx
import two_stage as ts

ctx = default_ctx()
lib = load()
lib.tevo_debug_set_heig_mode(2)   # the two-stage reduction is opt-in
PD = C.POINTER(C.c_double)
for n in [int(x) for x in sys.argv[1:]]:
    rng = np.random.default_rng(n)
    X = rng.standard_normal((n, n // 2 + 1)) + 1j * rng.standard_normal((n, n // 2 + 1))
    A = np.asfortranarray(X @ X.conj().T)
    w_ref = np.linalg.eigvalsh(A)
    band = np.zeros((16, n), dtype=np.complex128, order="F")
    V = np.zeros((n, n), dtype=np.complex128, order="F")
    taus = np.zeros(n, dtype=np.complex128)
    d = np.zeros(n); e = np.zeros(n)
    k = min(n, 64)
    Tz = None
    # first call: stages only
    check(lib.tevo_test_sbr(ctx.h, n, A.ctypes.data, band.ctypes.data, V.ctypes.data, taus.ctypes.data,
                            d.ctypes.data_as(PD), e.ctypes.data_as(PD), 0, None, None), ctx.h)
    B = np.zeros((n, n), dtype=complex)
    for c in range(n):
        for dd in range(9):
            if c + dd < n:
                B[c + dd, c] = band[dd, c]
                B[c, c + dd] = np.conj(band[dd, c])
    wb = np.linalg.eigvalsh(B)
    msg = "n=%d  stage1 eig err %.2e" % (n, np.abs(wb - w_ref).max() / w_ref[-1])
    if n <= 300:
        Bp, Vp, tp = ts.stage1(A, 8)
        msg += "  band vs proto %.2e  V vs proto %.2e  taus %.2e" % (np.abs(B - Bp).max() / w_ref[-1], np.abs(V - Vp).max(),
                                                                     np.abs(taus - tp).max())
    T = np.diag(d) + np.diag(e[:n - 1], -1) + np.diag(e[:n - 1], 1)
    wt, Z = np.linalg.eigh(T)
    msg += "  stage2 eig err %.2e" % (np.abs(wt - w_ref).max() / w_ref[-1])
    Zk = np.asfortranarray(Z[:, -k:])
    Xo = np.zeros((n, k), dtype=np.complex128, order="F")
    check(lib.tevo_test_sbr(ctx.h, n, A.ctypes.data, None, None, None, d.ctypes.data_as(PD), e.ctypes.data_as(PD), k,
                            Zk.ctypes.data_as(PD), Xo.ctypes.data), ctx.h)
    msg += "  Q2: band resid %.2e orth %.2e" % (np.abs(B @ Xo - Xo * wt[-k:]).max() / w_ref[-1],
                                               np.abs(Xo.conj().T @ Xo - np.eye(k)).max())
    # whole eigensolver
    w = np.zeros(n); U = np.zeros((n, n), dtype=np.complex128, order="F")
    t0 = time.time()
    for it in range(3):
        t.profile_enable(True)
        check(lib.tevo_test_heig(ctx.h, n, A.ctypes.data, w.ctypes.data_as(PD), U.ctypes.data), ctx.h)
        prof = t.profile_read()
    msg += "  heig: eig err %.2e resid %.2e orth %.2e" % (np.abs(w - w_ref[::-1]).max() / w_ref[-1],
                                                         np.abs(A @ U - U * w).max() / w_ref[-1],
                                                         np.abs(U.conj().T @ U - np.eye(n)).max())
    print(msg)
    print("    ", {kk: round(v["ms"], 3) for kk, v in prof.items()}, "sum %.2f ms" % sum(v["ms"] for v in prof.values()))
