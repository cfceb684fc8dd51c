"""The two-stage Hermitian tridiagonalisation (csrc/sbr.cu: dense -> band of width 8 -> tridiagonal, opt-in through
tevo_debug_set_heig_mode(2)) against LAPACK: every stage on its own (band matrix, tridiagonal matrix, Q2
back-transformation) and the whole eigensolver, for sizes around the block boundaries, rank-deficient, clustered and
graded spectra.  Replaces LAPACK inside replacebond! (src/bondgate.jl:51) exactly like the one-stage default."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
PD = C.POINTER(C.c_double)


def _herm(n, kind, seed):
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    if kind == "random":
        w = rng.standard_normal(n)
    elif kind == "psd_rank_half":
        w = np.concatenate([rng.random(n // 2), np.zeros(n - n // 2)])
    elif kind == "graded":
        w = np.exp(-np.arange(n) * 30.0 / n)
    elif kind == "clustered":
        w = np.repeat(rng.random(max(n // 8, 1)), 8)[:n]
        w = np.concatenate([w, np.ones(n - len(w))])
    else:
        raise ValueError(kind)
    A = (Q * w) @ Q.conj().T
    return np.asfortranarray((A + A.conj().T) / 2), np.sort(w)


def _band_to_dense(band, n):
    B = np.zeros((n, n), dtype=complex)
    for d in range(min(9, n)):
        idx = np.arange(n - d)
        B[idx + d, idx] = band[d, :n - d]
        B[idx, idx + d] = np.conj(band[d, :n - d])
    return B


@pytest.mark.parametrize("n", [2, 3, 9, 10, 11, 17, 18, 26, 64, 65, 130, 300])
def test_stages_against_lapack(tevo, n):
    from timeevomps_jl_b200._lib import load, check, default_ctx
    ctx, lib = default_ctx(), load()
    A, _ = _herm(n, "random", n)
    w_ref = np.linalg.eigvalsh(A)
    scale = np.abs(w_ref).max()
    band = np.zeros((16, n), dtype=np.complex128, order="F")
    V = np.zeros((n, n), dtype=np.complex128, order="F")
    taus = np.zeros(n, dtype=np.complex128)
    d, e = np.zeros(n), np.zeros(n)
    check(lib.tevo_test_sbr(ctx.h, n, A.ctypes.data, band.ctypes.data, V.ctypes.data, taus.ctypes.data,
                            d.ctypes.data_as(PD), e.ctypes.data_as(PD), 0, None, None), ctx.h)
    B = _band_to_dense(band, n)
    assert np.abs(np.linalg.eigvalsh(B) - w_ref).max() <= 1e-13 * scale           # stage 1 is a similarity transform
    # the Householder vectors reproduce it: Q1^H A Q1 = B with Q1 = prod (I - tau v v^H)
    Q1 = np.eye(n, dtype=complex)
    for c in range(n):
        if taus[c] != 0:
            v = V[:, c]
            assert np.all(V[:c + 8, c] == 0) and V[c + 8, c] == 1                  # explicit zeros above the unit entry
            Q1 = Q1 - taus[c] * np.outer(Q1 @ v, v.conj())
    assert np.abs(Q1.conj().T @ A @ Q1 - B).max() <= 1e-13 * scale
    T = np.diag(d) + np.diag(e[:n - 1], -1) + np.diag(e[:n - 1], 1)
    assert np.abs(np.linalg.eigvalsh(T) - w_ref).max() <= 1e-13 * scale           # stage 2 as well


@pytest.mark.parametrize("kind", ["random", "psd_rank_half", "graded", "clustered"])
@pytest.mark.parametrize("n", [7, 40, 129, 512])
def test_eigensolver_two_stage(tevo, n, kind):
    from timeevomps_jl_b200._lib import load, check, default_ctx
    ctx, lib = default_ctx(), load()
    A, _ = _herm(n, kind, 7 * n)
    w_ref = np.linalg.eigvalsh(A)[::-1]
    scale = max(np.abs(w_ref).max(), 1e-300)
    w = np.zeros(n)
    U = np.zeros((n, n), dtype=np.complex128, order="F")
    lib.tevo_debug_set_heig_mode(2)
    try:
        check(lib.tevo_test_heig(ctx.h, n, A.ctypes.data, w.ctypes.data_as(PD), U.ctypes.data), ctx.h)
    finally:
        lib.tevo_debug_set_heig_mode(0)
    assert np.abs(w - w_ref).max() <= 1e-13 * scale
    assert np.abs(A @ U - U * w).max() <= 1e-13 * scale * np.sqrt(n)
    assert np.abs(U.conj().T @ U - np.eye(n)).max() <= 1e-12


def test_split_with_two_stage_matches_one_stage(tevo, oracle):
    """The truncated split of a chi = 128 two-site wavefunction gives the same spectrum and state with either
    reduction (same tolerances as the split tests of test_gpu_kernels.py)."""
    from timeevomps_jl_b200._lib import load
    lib = load()
    N = 12
    res = []
    for mode in (1, 2):
        lib.tevo_debug_set_heig_mode(mode)
        try:
            o = oracle.MPS.random(N, 64, seed=3)
            p = tevo.MPS.from_tensors(o.A, llim=o.llim, rlim=o.rlim)
            H = oracle.xxz_bondop(N, 1.0, 0.1).gates()
            specs = []
            for b in (5, 6, 4):
                g = oracle.gate_exp(H[b], -0.05j)
                specs.append(tevo.apply_gate(p, tevo.BondGate(g.G, b), maxdim=48, cutoff=1e-9))
            res.append((p, specs))
        finally:
            lib.tevo_debug_set_heig_mode(0)
    (p1, s1), (p2, s2) = res
    assert p1.bond_dims() == p2.bond_dims()
    for a, b in zip(s1, s2):
        np.testing.assert_allclose(a.eigs, b.eigs, rtol=1e-10, atol=1e-15)
        assert abs(a.truncerr - b.truncerr) <= 1e-10 * a.truncerr + 1e-18
    assert abs(abs(tevo.inner(p1, p2)) - 1.0) <= 1e-10
