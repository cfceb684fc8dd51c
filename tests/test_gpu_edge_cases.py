"""Edge cases of the two-site update through the C ABI: minimal chains, bond dimension 1, extreme truncation
keywords, boundary bonds, rank-deficient and non-finite inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _to_dev(tevo, o):
    return tevo.MPS.from_tensors(o.A, llim=o.llim, rlim=o.rlim)


def _compare(tevo, oracle, o, p):
    assert p.bond_dims() == o.bond_dims()
    ov = oracle.inner(o, oracle.MPS(p.tensors()))
    assert abs(abs(ov) - abs(oracle.inner(o, o))) <= TOL
    np.testing.assert_allclose(p.copy().expect_all("Sz"), oracle.expect_all(o, "Sz"), atol=TOL)


def test_two_site_chain_tebd_and_tdvp(tevo, oracle):
    N = 2
    o = oracle.MPS.product([0, 1])
    p = tevo.MPS.product([0, 1])
    oracle.tebd(o, oracle.tfi_bondop(N, 1.0, 0.7), 0.05, 0.2, oracle.TEBD4())
    tevo.tebd(p, tevo.tfi_bondop(N, 1.0, 0.7), 0.05, 0.2, tevo.TEBD4())
    _compare(tevo, oracle, o, p)
    oracle.tdvp(o, oracle.tfi_mpo(N, 1.0, 0.7), 0.05, 0.1, exp_tol=1e-12)
    tevo.tdvp(p, tevo.tfi_mpo(1.0, 0.7, N), 0.05, 0.1, exp_tol=1e-12)
    _compare(tevo, oracle, o, p)


@pytest.mark.parametrize("kw", [dict(maxdim=1), dict(cutoff=0.5), dict(cutoff=1e-30), dict(maxdim=3, mindim=3),
                                dict(cutoff=0.3, mindim=2), dict(cutoff=1e-3, use_absolute_cutoff=True),
                                dict(maxdim=1000), dict(cutoff=0.0)])
def test_extreme_truncation_keywords(tevo, oracle, kw):
    N = 6
    o = oracle.MPS.random(N, 8, seed=3)
    p = _to_dev(tevo, o)
    H = oracle.tfi_bondop(N, 1.0, 0.4).gates()
    for b in (2, 0, 4, 3):
        g = oracle.gate_exp(H[b], -0.2j)
        ospec = oracle.apply_gate(o, g, **kw)
        spec = tevo.apply_gate(p, tevo.BondGate(g.G, b), **kw)
        assert len(spec) == len(ospec.eigs), (b, kw)
        # north-star tolerance 1e-10 on the weights; the absolute floor covers weights below ~1e-5 of the largest,
        # which the density-matrix route resolves absolutely
        np.testing.assert_allclose(spec.eigs, ospec.eigs, rtol=1e-10, atol=1e-15)
        assert abs(spec.truncerr - ospec.truncerr) <= 1e-10 * ospec.truncerr + 1e-15
    _compare(tevo, oracle, o, p)


def test_bond_dimension_one_product_state_identity_gate(tevo, oracle):
    N = 5
    p = tevo.MPS.product([0, 1, 0, 1, 1])
    for b in range(N - 1):
        spec = tevo.apply_gate(p, tevo.BondGate(np.eye(4), b), cutoff=1e-14)
        assert len(spec) == 1 and abs(spec.eigs[0] - 1.0) < 1e-14 and spec.truncerr <= 1e-28
    assert p.bond_dims() == [1, 1, 1, 1]
    np.testing.assert_allclose(p.expect_all("Sz"), [0.5, -0.5, 0.5, -0.5, -0.5], atol=1e-14)


def test_layer_matches_gate_by_gate(tevo, oracle):
    """tevo_apply_layer (one call, ortho chosen by direction) == literal per-gate apply_gate! with ortho="left"."""
    N = 9
    o = oracle.MPS.random(N, 6, seed=8)
    a, b = _to_dev(tevo, o), _to_dev(tevo, o)
    H = tevo.gates(tevo.xxz_bondop(N, 0.9, 0.2))
    layer = [tevo.exp(-0.1j * g) for g in H[1::2]]
    for rev in (False, True):
        tevo.apply_gates(a, layer, reversed=rev, maxdim=5)
        for g in (layer[::-1] if rev else layer):
            tevo.apply_gate(b, g, maxdim=5)
        assert a.bond_dims() == b.bond_dims()
        assert abs(abs(tevo.inner(a, b)) - 1.0) <= TOL
        np.testing.assert_allclose(a.copy().expect_all("Sx"), b.copy().expect_all("Sx"), atol=TOL)


def test_reorthogonalize_and_copy(tevo, oracle):
    o = oracle.MPS.random(7, 5, seed=12, canonical=False)
    p = _to_dev(tevo, o)
    q = p.copy()
    p.reorthogonalize()
    assert p.limits() == (-1, 1)
    assert abs(tevo.inner(p, p) - 1.0) <= 1e-12
    nq = tevo.inner(q, q).real
    assert abs(abs(tevo.inner(q, p)) / np.sqrt(nq) - 1.0) <= 1e-12     # same state up to norm; copy untouched
    assert abs(nq - oracle.inner(o, o).real) <= 1e-10 * nq


def test_non_finite_input_is_an_error_not_a_hang(tevo):
    t = [np.ones((1, 2, 2), complex), np.ones((2, 2, 2), complex), np.ones((2, 2, 1), complex)]
    t[1][0, 0, 0] = np.nan
    p = tevo.MPS.from_tensors(t)
    with pytest.raises(tevo.TevoError):
        tevo.apply_gate(p, tevo.BondGate(np.eye(4), 0), maxdim=2)


def test_measure_bond_requires_centre(tevo):
    p = tevo.MPS.product([0] * 6)
    p.orthogonalize(0)
    with pytest.raises(tevo.TevoError):
        p.measure_bond(4, ["Sz"])
    p.orthogonalize(4)
    m = p.measure_bond(4, ["Sz", "Id"])
    assert abs(m[0, 0] - 0.5) < 1e-14 and abs(m[1, 1] - 1.0) < 1e-14
