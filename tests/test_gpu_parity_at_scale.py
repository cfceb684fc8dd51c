"""Parity with the oracle AT THE BASELINE SCALES (chi = 1024 two-site updates, chi = 64 TEBD4 with an active cutoff,
chi = 256 TDVP with the w = 5 XXZ MPO), where the multi-CTA tridiagonalisation, the multi-level divide & conquer, the
blocked back-transformation and the blocked CholeskyQR centre moves all engage together.  Everything goes through the
C ABI; tolerance 1e-10 on gauge-invariant quantities (north star), stated per assertion.

Reference call sites: apply_gate! src/bondgate.jl:46-53, tebd! src/tebd.jl:91-179, tdvp! src/tdvp.jl:54-95,
SpecCallback src/callback.jl:139-160."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-10
SIGMA2_ATOL = 1e-15   # absolute floor on sigma^2 (sum of all sigma^2 = 1); measured on the B200: max rel dev 5.2e-11 over
#                       kept weights from 2e-2 down to 1e-9, max abs dev 1.4e-16 (profiles/r02_scale_parity_tests.log)


def _to_dev(tevo, opsi):
    return tevo.MPS.from_tensors(opsi.A, llim=opsi.llim, rlim=opsi.rlim)


def test_apply_gate_chi1024_matches_oracle(tevo, oracle):
    """Three interior two-site updates at chi = 1024 (theta 2048 x 2048, mindim = maxdim = 1024: the bench's own
    configuration) against oracle.apply_gate (QR centre move + LAPACK SVD)."""
    N, chi = 24, 1024
    o = oracle.MPS.random(N, chi, seed=7)
    p = _to_dev(tevo, o)
    H = oracle.xxz_bondop(N, 1.0, 0.3).gates()
    for b in (10, 12, 11):
        g = oracle.gate_exp(H[b], -0.05j)
        ospec = oracle.apply_gate(o, g, maxdim=chi, mindim=chi)
        spec = tevo.apply_gate(p, tevo.BondGate(g.G, b), maxdim=chi, mindim=chi)
        assert len(spec) == len(ospec.eigs) == chi
        rel = np.abs(spec.eigs - ospec.eigs) / ospec.eigs
        print("bond %d: kept sigma^2 from %.2e to %.2e, max rel dev %.2e (max abs dev %.2e), truncerr %.3e vs %.3e"
              % (b, ospec.eigs[0], ospec.eigs[-1], rel.max(), np.abs(spec.eigs - ospec.eigs).max(), spec.truncerr,
                 ospec.truncerr))
        # kept sigma^2 span 1e-3 .. 1e-9 here: 1e-10 relative on the weights that matter, SIGMA2_ATOL absolute on the
        # smallest ones (density-matrix route: DESIGN.md section 5)
        np.testing.assert_allclose(spec.eigs, ospec.eigs, rtol=TOL, atol=SIGMA2_ATOL)
        assert abs(spec.truncerr - ospec.truncerr) <= TOL * ospec.truncerr            # discarded weight
        assert abs(spec.entropy() - ospec.entropy()) <= TOL
        assert p.limits() == (o.llim, o.rlim)
    assert p.bond_dims() == o.bond_dims()
    po = _to_dev(tevo, o)
    ov = tevo.inner(p, po)
    assert abs(abs(ov) - 1.0) <= TOL                                                  # same state
    assert abs(tevo.inner(p, p).real - 1.0) <= TOL
    sz_o = po.expect_all("Sz")          # the oracle's state measured by the device (the CPU sweep would take minutes)
    sz_p = p.copy().expect_all("Sz")
    np.testing.assert_allclose(sz_p, sz_o, atol=TOL)
    # and three sites measured by the oracle itself, so the device observable path is not its own judge
    for j in (10, 11, 12):
        assert abs(sz_p[j] - oracle.expect_site(o, "Sz", j).real) <= TOL


def test_tebd4_chi64_cutoff_active_matches_oracle(tevo, oracle):
    """SURVEY 8(d) parity run: C3-shaped TEBD4 (XXZ, dt = 0.05) at N = 16, chi = 64 with maxdim AND an active cutoff,
    full tebd! incl. SpecCallback, from a random state whose bonds are saturated so truncation bites on every gate."""
    N, chi = 16, 64
    o = oracle.MPS.random(N, chi, seed=16)
    p = _to_dev(tevo, o)
    osc = oracle.SpecCallback(0.05, N)
    sc = tevo.SpecCallback(0.05, N)
    kw = dict(maxdim=chi, cutoff=1e-7)
    oracle.tebd(o, oracle.xxz_bondop(N, 1.0, 0.2), 0.05, 0.1, oracle.TEBD4(), callback=osc, **kw)
    tevo.tebd(p, tevo.xxz_bondop(N, 1.0, 0.2), 0.05, 0.1, tevo.TEBD4(), callback=sc, **kw)
    assert p.bond_dims() == o.bond_dims()
    assert min(o.bond_dims()[4:-4]) < chi or max(osc.truncerrs) > 0          # the cutoff / maxdim really acted
    assert [list(x) for x in sc.bonddims] == [list(x) for x in osc.bonddims]
    np.testing.assert_allclose(np.array(sc.entropies), np.array(osc.entropies), atol=TOL)
    np.testing.assert_allclose(np.array(sc.truncerrs), np.array(osc.truncerrs), rtol=1e-8, atol=1e-16)
    ov = oracle.inner(o, oracle.MPS(p.tensors()))
    assert abs(abs(ov) - 1.0) <= TOL
    for name in ("Sz", "Sx", "Sy"):
        np.testing.assert_allclose(p.copy().expect_all(name), oracle.expect_all(o, name), atol=TOL)
    np.testing.assert_allclose(oracle.bond_entropies(oracle.MPS(p.tensors())), oracle.bond_entropies(o), atol=TOL)


def test_tdvp_chi256_w5_matches_oracle(tevo, oracle):
    """One full two-site TDVP step (forward two-site + backward one-site updates on every bond, src/tdvp.jl:56-88) of
    the w = 5 XXZ MPO at chi = 256 (N = 18: the five central bonds are saturated), mindim = maxdim = chi."""
    N, chi = 18, 256
    o = oracle.MPS.random(N, chi, seed=18)
    p = _to_dev(tevo, o)
    ost, st = {}, {}
    kw = dict(maxdim=chi, mindim=chi, exp_tol=1e-12, krylovdim=30)
    oracle.tdvp(o, oracle.xxz_mpo(N, 1.0, 0.3), 0.05, 0.05, stats=ost, **kw)
    tevo.tdvp(p, tevo.xxz_mpo(N, 1.0, 0.3), 0.05, 0.05, stats=st, **kw)
    assert p.bond_dims() == o.bond_dims()
    assert max(p.bond_dims()) == chi
    ov = oracle.inner(o, oracle.MPS(p.tensors()))
    assert abs(abs(ov) - 1.0) <= TOL
    assert abs(tevo.inner(p, p).real - 1.0) <= TOL
    for name in ("Sz", "Sx"):
        np.testing.assert_allclose(p.copy().expect_all(name), oracle.expect_all(o, name), atol=TOL)
    np.testing.assert_allclose(oracle.bond_entropies(oracle.MPS(p.tensors())), oracle.bond_entropies(o), atol=TOL)
    e_o = oracle.inner_mpo(o, oracle.xxz_mpo(N, 1.0, 0.3), o)
    e_p = tevo.inner_mpo(p, tevo.xxz_mpo(N, 1.0, 0.3), p)
    assert abs(e_o - e_p) <= TOL * max(1.0, abs(e_o))
