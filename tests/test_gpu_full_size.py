"""BASELINE-size checks (chi = 1024: the split of a 2048 x 2048 two-site wavefunction) through properties that do not
need the oracle to run a full step: isometry, reconstruction, consistency of the reported discarded weight, spectrum
vs LAPACK, and norm preservation of a sequential and a layer-parallel gate on a chi=1024 chain."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _split(theta, left, **tr):
    from timeevomps_jl_b200._lib import load, check, tevo_spec, default_ctx, make_trunc
    ctx = default_ctx()
    m, n = theta.shape
    th = np.asfortranarray(theta)
    k = min(m, n)
    L = np.zeros((m, k), dtype=np.complex128, order="F")
    R = np.zeros((k, n), dtype=np.complex128, order="F")
    eigs = np.zeros(k)
    spec = tevo_spec()
    t = make_trunc(**tr)
    check(load().tevo_test_split(ctx.h, m, n, th.ctypes.data, C.byref(t), 1 if left else 0, 0, C.byref(spec),
                                 eigs.ctypes.data_as(C.POINTER(C.c_double)), k, L.ctypes.data, R.ctypes.data, k), ctx.h)
    kk = spec.nkeep
    Lf = np.ndarray((m, kk), dtype=np.complex128, buffer=L.data, order="F").copy()
    Rf = np.ndarray((kk, n), dtype=np.complex128, buffer=R.data, order="F").copy()
    return Lf, Rf, eigs[:kk].copy(), spec


@pytest.mark.parametrize("left", [True, False])
def test_split_2048_maxdim_1024(tevo, left):
    rng = np.random.default_rng(2048)
    m = n = 2048
    # a two-site wavefunction with a decaying spectrum, like an evolved state: random isometries x graded weights
    A = rng.standard_normal((m, 1400)) + 1j * rng.standard_normal((m, 1400))
    B = rng.standard_normal((1400, n)) + 1j * rng.standard_normal((1400, n))
    th = (A * np.exp(-np.arange(1400) / 150.0)) @ B
    th /= np.linalg.norm(th)
    L, R, eigs, spec = _split(th, left, maxdim=1024)
    assert spec.nkeep == 1024
    iso = L.conj().T @ L if left else R @ R.conj().T
    assert np.abs(iso - np.eye(1024)).max() < 1e-12                       # exact isometry at full size
    res = np.linalg.norm(th - L @ R) ** 2
    assert abs(res - spec.truncerr * spec.sum_all) <= 1e-9 * res + 1e-18     # reported discarded weight = actual error
    assert abs(spec.sum_all - 1.0) < 1e-12
    assert abs(spec.sum_kept + res - spec.sum_all) < 1e-12
    S = np.linalg.svd(th, compute_uv=False)
    dev = np.abs(eigs - S[:1024] ** 2) / S[:1024] ** 2
    print("2048^2 split, left=%s: kept sigma^2 %.2e .. %.2e, max rel dev vs LAPACK %.2e" % (left, eigs[0], eigs[-1], dev.max()))
    np.testing.assert_allclose(eigs, S[:1024] ** 2, rtol=1e-10, atol=1e-15)                   # north-star tolerance
    assert abs(spec.truncerr - np.sum(S[1024:] ** 2) / np.sum(S ** 2)) <= 1e-10 * spec.truncerr


def test_gates_on_chi1024_chain_preserve_norm(tevo):
    """One sequential and one layer-parallel TEBD layer on a saturated chi=1024 chain: norm 1, bond dims kept."""
    N, chi = 22, 1024
    psi = tevo.MPS.random(N, chi, seed=3)
    H = tevo.gates(tevo.xxz_bondop(N, 1.0))
    layer = [tevo.exp(-0.05j * g) for g in H[0::2]]
    dims = psi.bond_dims()
    tevo.apply_gates(psi, layer, maxdim=chi, mindim=chi)
    assert psi.bond_dims() == dims
    assert abs(tevo.inner(psi, psi) - 1.0) < 1e-11
    par = psi.copy().to_bform()
    tevo.apply_gates_parallel(par, [tevo.exp(-0.05j * g) for g in H[1::2]], maxdim=chi, mindim=chi)
    assert par.bond_dims() == dims
    lam = par.schmidt_values(N // 2)
    assert abs(np.sum(lam ** 2) - 1.0) < 1e-12
    assert abs(abs(tevo.inner(par, par)) - 1.0) < 1e-6      # B-form after truncation is canonical only up to O(eps_trunc)
