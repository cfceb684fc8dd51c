"""GPU parity of the two-site TDVP path against the CPU oracle, through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _to_dev(tevo, opsi):
    return tevo.MPS.from_tensors(opsi.A, llim=opsi.llim, rlim=opsi.rlim)


def test_heff_apply_matches_oracle(tevo, oracle):
    N = 8
    o = oracle.MPS.random(N, 7, seed=11)
    p = _to_dev(tevo, o)
    for (Wo, Hd) in ((oracle.tfi_mpo(N, 1.0, 0.5), tevo.tfi_mpo(1.0, 0.5, N)),
                     (oracle.xxz_mpo(N, 0.8, 0.3), tevo.xxz_mpo(N, 0.8, 0.3))):
        PH = oracle.ProjMPO(Wo)
        env = tevo.ProjMPO(Hd)
        rng = np.random.default_rng(0)
        for pos, nsite in ((3, 2), (0, 2), (6, 2), (4, 1), (0, 1), (7, 1), (2, 2)):
            PH.nsite = nsite
            PH.position(o, pos)
            shape = o.theta(pos).shape if nsite == 2 else o.A[pos].shape
            v = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
            ref = PH.product(pos, v)
            got = env.product(p, pos, nsite, v)
            assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
        e_o = oracle.inner_mpo(o, Wo, o)
        e_p = tevo.inner_mpo(p, Hd, p)
        assert abs(e_o - e_p) <= 1e-12 * max(1, abs(e_o))


def test_tdvp_config2_matches_oracle(tevo, oracle):
    """BASELINE config 2: TFIM N=10 two-site TDVP dt=0.1 maxdim=30 with LocalMeasurementCallback Sx/Sy/Sz
    (tf shortened to 1.0; exp_tol 1e-12)."""
    N = 10
    ocb = oracle.LocalMeasurementCallback(["Sz", "Sx", "Sy"], N, 0.2)
    o = oracle.MPS.product([0] * N)
    ost = {}
    oracle.tdvp(o, oracle.tfi_mpo(N, 1.0, 0.5), 0.1, 1.0, maxdim=30, callback=ocb, exp_tol=1e-12, stats=ost)
    cb = tevo.LocalMeasurementCallback(["Sz", "Sx", "Sy"], N, 0.2)
    p = tevo.MPS.product([0] * N)
    st = {}
    tevo.tdvp(p, tevo.tfi_mpo(1.0, 0.5, N), 0.1, 1.0, maxdim=30, callback=cb, exp_tol=1e-12, stats=st)
    assert p.bond_dims() == o.bond_dims()
    np.testing.assert_allclose(cb.ts, ocb.ts)
    for name in ("Sz", "Sx", "Sy"):
        np.testing.assert_allclose(np.array(cb.measurements[name]), np.array(ocb.measurements[name]), atol=TOL)
    ov = oracle.inner(o, oracle.MPS(p.tensors()))
    assert abs(abs(ov) - 1.0) <= TOL
    assert abs(tevo.inner(p, p) - 1.0) <= TOL
    np.testing.assert_allclose(oracle.bond_entropies(oracle.MPS(p.tensors())), oracle.bond_entropies(o), atol=TOL)


def test_tdvp_truncated_xxz_matches_oracle(tevo, oracle):
    N = 8
    o = oracle.MPS.random(N, 4, seed=21)
    p = _to_dev(tevo, o)
    osc = oracle.SpecCallback(0.1, N)
    sc = tevo.SpecCallback(0.1, N)
    oracle.tdvp(o, oracle.xxz_mpo(N, 1.0, 0.2), 0.05, 0.2, maxdim=6, cutoff=1e-9, exp_tol=1e-12, callback=osc)
    tevo.tdvp(p, tevo.xxz_mpo(N, 1.0, 0.2), 0.05, 0.2, maxdim=6, cutoff=1e-9, exp_tol=1e-12, callback=sc)
    assert p.bond_dims() == o.bond_dims()
    assert [list(x) for x in sc.bonddims] == [list(x) for x in osc.bonddims]
    np.testing.assert_allclose(np.array(sc.entropies), np.array(osc.entropies), atol=TOL)
    ov = oracle.inner(o, oracle.MPS(p.tensors()))
    assert abs(abs(ov) - 1.0) <= TOL
    for name in ("Sz", "Sx"):
        np.testing.assert_allclose(p.copy().expect_all(name), oracle.expect_all(o, name), atol=TOL)
