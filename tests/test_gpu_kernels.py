"""GPU parity of the primitive kernels, called through the C ABI test hooks (include/tevo.h)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rand(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def _zgemm(tevo, ta, tb, M, N, K, alpha, A, B, beta, Cm):
    from timeevomps_jl_b200._lib import load, check, tevo_complex, default_ctx
    ctx = default_ctx()
    A = np.asfortranarray(A.astype(np.complex128))
    B = np.asfortranarray(B.astype(np.complex128))
    Cm = np.asfortranarray(Cm.astype(np.complex128)).copy(order="F")
    check(load().tevo_test_zgemm(ctx.h, ta, tb, M, N, K, tevo_complex(alpha.real, alpha.imag), A.ctypes.data,
                                 A.shape[0], B.ctypes.data, B.shape[0], tevo_complex(beta.real, beta.imag),
                                 Cm.ctypes.data, Cm.shape[0]), ctx.h)
    return Cm


def _opnp(X, t):
    return X if t == 0 else (X.T if t == 1 else X.conj().T)


@pytest.fixture
def zgemm_tile():
    """Forces one CTA tile of the DMMA ZGEMM for the duration of a test (the automatic choice depends on the problem
    size: small products would never reach the 128 x 64 kernel, large ones never the 32 x 32 one)."""
    from timeevomps_jl_b200._lib import load
    yield lambda tile, m3=1: load().tevo_debug_set_zgemm_tile(tile) + load().tevo_debug_set_zgemm_3m(m3)
    load().tevo_debug_set_zgemm_tile(-1)
    load().tevo_debug_set_zgemm_3m(1)


@pytest.mark.parametrize("tile", [-1, 0, 1, 2, 11, 12])
@pytest.mark.parametrize("ta", [0, 1, 2])
@pytest.mark.parametrize("tb", [0, 1, 2])
@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 4, 2), (7, 5, 3), (33, 65, 17), (128, 64, 16), (130, 70, 50),
                                   (256, 192, 129), (300, 20, 515), (520, 500, 40), (1100, 1090, 20)])
def test_zgemm_matches_numpy(tevo, zgemm_tile, ta, tb, shape, tile):
    """DMMA ZGEMM vs numpy complex128; tolerance 1e-13 relative to |A||B| (fp64 accumulation order differs).  Every
    shape with the automatic tile choice and with each of the three CTA tiles forced; the last two shapes reach the
    64 x 32 and the 128 x 64 tile on their own.  Tiles 11 / 12: the 64 x 32 / 32 x 32 kernels with four real
    multiplications per complex product instead of their default three."""
    M, N, K = shape
    assert zgemm_tile(tile - 10 if tile >= 10 else tile, 0 if tile >= 10 else 1) == 0
    rng = np.random.default_rng(M * 1000 + N * 10 + K + ta * 3 + tb)
    A = _rand(rng, M, K) if ta == 0 else _rand(rng, K, M)
    B = _rand(rng, K, N) if tb == 0 else _rand(rng, N, K)
    C0 = _rand(rng, M, N)
    alpha, beta = 0.7 - 0.3j, -0.2 + 1.1j
    got = _zgemm(tevo, ta, tb, M, N, K, alpha, A, B, beta, C0)
    ref = alpha * (_opnp(A, ta) @ _opnp(B, tb)) + beta * C0
    scale = np.linalg.norm(_opnp(A, ta), axis=1).max() * np.linalg.norm(_opnp(B, tb), axis=0).max() + 1
    assert np.abs(got - ref).max() <= 1e-13 * scale
    got0 = _zgemm(tevo, ta, tb, M, N, K, 1.0 + 0j, A, B, 0j, np.full((M, N), np.nan + 0j))
    assert np.abs(got0 - _opnp(A, ta) @ _opnp(B, tb)).max() <= 1e-13 * scale


def _split(tevo, theta, left, normalize=False, **tr):
    from timeevomps_jl_b200._lib import load, check, tevo_spec, default_ctx, make_trunc
    ctx = default_ctx()
    m, n = theta.shape
    th = np.asfortranarray(theta.astype(np.complex128))
    k = min(m, n)
    L = np.zeros((m, k), dtype=np.complex128, order="F")
    R = np.zeros((k, n), dtype=np.complex128, order="F")
    eigs = np.zeros(k)
    spec = tevo_spec()
    t = make_trunc(**tr)
    check(load().tevo_test_split(ctx.h, m, n, th.ctypes.data, C.byref(t), 1 if left else 0, 1 if normalize else 0,
                                 C.byref(spec), eigs.ctypes.data_as(C.POINTER(C.c_double)), k, L.ctypes.data,
                                 R.ctypes.data, k), ctx.h)
    kk = spec.nkeep
    Lf = np.ndarray((m, kk), dtype=np.complex128, buffer=L.data, order="F").copy()
    Rf = np.ndarray((kk, n), dtype=np.complex128, buffer=R.data, order="F").copy()
    return Lf, Rf, eigs[:kk].copy(), spec


@pytest.mark.parametrize("left", [True, False])
@pytest.mark.parametrize("shape", [(1, 1), (2, 2), (4, 2), (2, 4), (8, 8), (20, 20), (64, 32), (32, 64), (96, 96)])
def test_split_untruncated(tevo, oracle, left, shape):
    m, n = shape
    rng = np.random.default_rng(m * 100 + n)
    th = _rand(rng, m, n)
    L, R, eigs, spec = _split(tevo, th, left)
    k = min(m, n)
    assert spec.nkeep == k and spec.truncerr == 0.0
    S = np.linalg.svd(th, compute_uv=False)
    np.testing.assert_allclose(eigs, S ** 2, rtol=1e-11, atol=1e-13 * S[0] ** 2)
    assert np.abs(L @ R - th).max() <= 1e-12 * S[0]
    iso = L.conj().T @ L if left else R @ R.conj().T
    assert np.abs(iso - np.eye(k)).max() <= 1e-12


@pytest.mark.parametrize("left", [True, False])
@pytest.mark.parametrize("tr", [dict(maxdim=7), dict(cutoff=1e-3), dict(maxdim=20, cutoff=1e-6),
                                dict(cutoff=1e-2, mindim=12), dict(cutoff=0.05, use_absolute_cutoff=True),
                                dict(maxdim=5, mindim=5)])
def test_split_truncation_matches_oracle(tevo, oracle, left, tr):
    """Same bond dimension, truncerr, spectrum, entropy and truncated state as the oracle's replacebond!."""
    rng = np.random.default_rng(7)
    m, n = 24, 32
    # decaying spectrum so that cutoffs bite
    U, _ = np.linalg.qr(_rand(rng, m, m))
    V, _ = np.linalg.qr(_rand(rng, n, n))
    s = np.exp(-0.6 * np.arange(m))
    th = (U * s) @ V[:m]
    L, R, eigs, spec = _split(tevo, th, left, normalize=True, **tr)
    psi = oracle.MPS([np.zeros((m // 2, 2, 1)), np.zeros((1, 2, n // 2))])
    ospec = psi.replacebond(0, th.reshape(m // 2, 2, 2, n // 2), ortho="left" if left else "right", normalize=True,
                            **tr)
    assert spec.nkeep == len(ospec.eigs)
    np.testing.assert_allclose(eigs, ospec.eigs, rtol=1e-10)
    assert abs(spec.truncerr - ospec.truncerr) <= 1e-10 * max(ospec.truncerr, 1e-30) + 1e-18
    assert abs(spec.entropy - ospec.entropy()) <= 1e-10
    oth = np.tensordot(psi.A[0], psi.A[1], axes=(2, 0)).reshape(m, n)
    assert np.abs(L @ R - oth).max() <= 1e-12
    assert abs(np.linalg.norm(L @ R) - 1.0) <= 1e-13


def test_split_rank_deficient(tevo):
    """Exact zero singular values (product-state growth, test/test_tebd.jl:31-33): isometry stays unitary."""
    rng = np.random.default_rng(3)
    a, b = _rand(rng, 16, 3), _rand(rng, 3, 16)
    th = a @ b
    for left in (True, False):
        L, R, eigs, spec = _split(tevo, th, left)
        assert spec.nkeep == 16
        assert np.abs(L @ R - th).max() <= 1e-12 * np.linalg.norm(th)
        iso = L.conj().T @ L if left else R @ R.conj().T
        assert np.abs(iso - np.eye(16)).max() <= 1e-12
        assert np.all(np.isfinite(L)) and np.all(np.isfinite(R))


def _heig(tevo, A):
    from timeevomps_jl_b200._lib import load, check, default_ctx
    ctx = default_ctx()
    n = A.shape[0]
    Af = np.asfortranarray(A.astype(np.complex128))
    w = np.zeros(n)
    U = np.zeros((n, n), dtype=np.complex128, order="F")
    check(load().tevo_test_heig(ctx.h, n, Af.ctypes.data, w.ctypes.data_as(C.POINTER(C.c_double)), U.ctypes.data),
          ctx.h)
    return w, U


@pytest.mark.parametrize("n", [1, 2, 3, 5, 16, 31, 32, 33, 64, 65, 100, 129, 200, 257, 512])
def test_heig_random_hermitian(tevo, n):
    """tridiagonalisation + divide&conquer + back-transformation vs numpy.linalg.eigh"""
    rng = np.random.default_rng(n)
    X = _rand(rng, n, n)
    A = X + X.conj().T
    w, U = _heig(tevo, A)
    wr = np.linalg.eigvalsh(A)[::-1]
    nrm = np.abs(wr).max()
    assert np.abs(w - wr).max() <= 1e-13 * nrm * max(1, n ** 0.5)
    assert np.abs(U.conj().T @ U - np.eye(n)).max() <= 1e-13 * max(1, n ** 0.5)
    assert np.abs(A @ U - U * w).max() <= 1e-13 * nrm * max(1, n)


@pytest.mark.parametrize("n,ctas", [(300, 7), (300, 3), (520, 5), (333, 1), (200, 40)])
def test_heig_capped_grid(tevo, n, ctas):
    """The panel kernel with fewer CTAs than 16-row strips (several strips per CTA, some beyond the shared-memory
    resident ones): the configuration of the lanes in the layer-parallel mode."""
    from timeevomps_jl_b200._lib import load, check, default_ctx
    ctx = default_ctx()
    rng = np.random.default_rng(1000 * n + ctas)
    X = _rand(rng, n, n)
    A = np.asfortranarray(X + X.conj().T)
    w = np.zeros(n)
    U = np.zeros((n, n), dtype=np.complex128, order="F")
    check(load().tevo_test_heig_ctas(ctx.h, n, A.ctypes.data, w.ctypes.data_as(C.POINTER(C.c_double)), U.ctypes.data,
                                     ctas), ctx.h)
    wr = np.linalg.eigvalsh(A)[::-1]
    nrm = np.abs(wr).max()
    assert np.abs(w - wr).max() <= 1e-13 * nrm * n ** 0.5
    assert np.abs(U.conj().T @ U - np.eye(n)).max() <= 1e-13 * n ** 0.5
    assert np.abs(A @ U - U * w).max() <= 1e-13 * nrm * n
    # and the default grid still gives the same spectrum afterwards (the cap is restored)
    w2, _ = _heig(tevo, A)
    assert np.abs(w2 - wr).max() <= 1e-13 * nrm * n ** 0.5


@pytest.mark.parametrize("kind", ["lowrank", "graded", "clustered", "zero", "identity"])
def test_heig_hard_spectra(tevo, kind):
    rng = np.random.default_rng(11)
    n = 160
    Q, _ = np.linalg.qr(_rand(rng, n, n))
    if kind == "lowrank":
        s = np.concatenate([rng.random(9) + 0.5, np.zeros(n - 9)])
    elif kind == "graded":
        s = 10.0 ** (-np.arange(n) / 8.0)
    elif kind == "clustered":
        s = np.concatenate([np.ones(50), 0.5 * np.ones(50) * (1 + 1e-13 * rng.random(50)), 1e-3 * np.ones(60)])
    elif kind == "zero":
        s = np.zeros(n)
    else:
        s = np.ones(n)
    A = (Q * s) @ Q.conj().T
    A = 0.5 * (A + A.conj().T)
    w, U = _heig(tevo, A)
    nrm = max(np.abs(s).max(), 1e-300)
    assert np.abs(np.sort(w) - np.sort(s)).max() <= 1e-13 * nrm * n ** 0.5 + 1e-300
    assert np.abs(U.conj().T @ U - np.eye(n)).max() <= 2e-13 * n ** 0.5
    assert np.abs(A @ U - U * w).max() <= 1e-13 * nrm * n + 1e-300


def _potrf_block(G, variant, reps=0, cycles=False):
    from timeevomps_jl_b200._lib import load, check, default_ctx
    ctx = default_ctx()
    nb = G.shape[0]
    Gf = np.asfortranarray(G.astype(np.complex128))
    L = np.zeros((nb, nb), dtype=np.complex128, order="F")
    X = np.zeros((nb, nb), dtype=np.complex128, order="F")
    stats = np.zeros(2)
    us = C.c_double(0.0)
    cyc = np.zeros(8)
    dp = C.POINTER(C.c_double)
    check(load().tevo_test_potrf_block(ctx.h, variant, nb, Gf.ctypes.data, L.ctypes.data, X.ctypes.data,
                                       stats.ctypes.data_as(dp), reps, C.byref(us),
                                       cyc.ctypes.data_as(dp) if cycles else None), ctx.h)
    return L, X, stats, us.value, cyc


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("nb", [1, 2, 3, 7, 8, 9, 15, 16, 17, 31, 32, 33, 40, 47, 48, 49, 56, 57, 63, 64])
def test_potrf_block_matches_lapack(tevo, nb, variant):
    """The diagonal-block step of the CholeskyQR factorisation (csrc/cholqr.cu; the centre moves of orthogonalize!,
    src/bondgate.jl:48): L = chol(G) and inv(L) of one nb x nb block, all three kernel variants, block sizes around
    every 8-column panel boundary of the default kernel."""
    rng = np.random.default_rng(100 * nb + variant)
    B = _rand(rng, 2 * nb + 3, nb) * np.linspace(1.0, 6.0, nb)
    G = B.conj().T @ B
    G = (G + G.conj().T) / 2
    L, X, stats, _, _ = _potrf_block(G, variant)
    Lr = np.linalg.cholesky(G)
    scale = np.abs(Lr).max()
    assert np.abs(np.triu(L, 1)).max() == 0.0 and np.abs(np.triu(X, 1)).max() == 0.0
    assert np.abs(np.diag(L).imag).max() == 0.0
    assert np.abs(L - Lr).max() <= 1e-13 * scale * np.linalg.cond(Lr)
    assert np.abs(L @ L.conj().T - G).max() <= 1e-14 * np.abs(G).max() * nb
    assert np.abs(X @ Lr - np.eye(nb)).max() <= 1e-13 * np.linalg.cond(Lr)
    d = np.diag(Lr).real
    np.testing.assert_allclose(stats, [d.min(), d.max()], rtol=1e-12)


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("where", [0, 5, 8, 37, 63])
def test_potrf_block_flags_indefinite_block(tevo, where, variant):
    """A non-positive pivot anywhere in the block sets stats[0] < 0 (the caller then takes the eigen-decomposition
    route) instead of producing NaNs."""
    rng = np.random.default_rng(7 + where)
    B = _rand(rng, 140, 64)
    G = B.conj().T @ B
    G = (G + G.conj().T) / 2
    G[where, where] = -1.0 if where == 0 else 0.5 * np.real(G[where, :where] @ np.linalg.solve(G[:where, :where], G[:where, where]))
    L, X, stats, _, _ = _potrf_block(G, variant)
    assert stats[0] < 0.0
