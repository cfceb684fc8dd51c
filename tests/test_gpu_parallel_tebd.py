"""Layer-parallel (B-form, Hastings) TEBD on the GPU vs its oracle restatement, and vs the sequential reference
algorithm where the two must coincide (no truncation)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-10


def test_to_bform_matches_oracle(tevo, oracle):
    N = 8
    o = oracle.MPS.random(N, 6, seed=3)
    ob = oracle.to_bform(o)
    p = tevo.MPS.from_tensors(o.A, llim=o.llim, rlim=o.rlim).to_bform()
    assert p.bond_dims() == ob.bond_dims()
    for j in range(N + 1):
        np.testing.assert_allclose(p.schmidt_values(j), ob.lam[j], atol=1e-12)
    ts = p.tensors()
    for j in range(N):
        M = ts[j].reshape(ts[j].shape[0], -1)
        assert np.abs(M @ M.conj().T - np.eye(M.shape[0])).max() < 1e-12      # right-orthonormal B_j
    assert abs(abs(oracle.inner(o, oracle.MPS(ts))) - 1.0) < 1e-12
    for name in ("Sz", "Sx"):
        np.testing.assert_allclose(p.bform_expect_all(name), ob.expect_all(name), atol=1e-12)


@pytest.mark.parametrize("kw", [dict(), dict(maxdim=4), dict(cutoff=1e-6), dict(maxdim=5, cutoff=1e-9)])
def test_parallel_tebd_matches_oracle_parallel(tevo, oracle, kw):
    N = 9
    o = oracle.MPS.random(N, 5, seed=5)
    ob = oracle.to_bform(o)
    p = tevo.MPS.from_tensors(o.A, llim=o.llim, rlim=o.rlim).to_bform()
    oracle.tebd_parallel(ob, oracle.xxz_bondop(N, 0.8, 0.3), 0.05, 0.2, oracle.TEBD4(), **kw)
    tevo.tebd_parallel(p, tevo.xxz_bondop(N, 0.8, 0.3), 0.05, 0.2, tevo.TEBD4(), **kw)
    assert p.bond_dims() == ob.bond_dims()
    for j in range(1, N):
        # Schmidt weights: 1e-10 relative down to sigma^2 ~ 1e-8 sigma_max^2; below that the density-matrix route
        # resolves them to ~n*eps*sigma_max^2 absolute
        np.testing.assert_allclose(p.schmidt_values(j) ** 2, ob.lam[j] ** 2, rtol=TOL, atol=2e-13)
    for name in ("Sz", "Sx", "Sy"):
        np.testing.assert_allclose(p.bform_expect_all(name), ob.expect_all(name), atol=TOL)
    om = ob.to_mps()
    ov = oracle.inner(om, oracle.MPS(p.tensors()))
    assert abs(abs(ov) - abs(oracle.inner(om, om))) <= TOL


def test_parallel_equals_sequential_without_truncation(tevo, oracle):
    """Without truncation the layer-parallel mode and the reference's sequential sweep give the same state."""
    N = 8
    o = oracle.MPS.random(N, 4, seed=7)
    oracle.tebd(o, oracle.tfi_bondop(N, 1.0, 0.6), 0.05, 0.2, oracle.TEBD4())
    p0 = oracle.MPS.random(N, 4, seed=7)
    p = tevo.MPS.from_tensors(p0.A, llim=p0.llim, rlim=p0.rlim).to_bform()
    tevo.tebd_parallel(p, tevo.tfi_bondop(N, 1.0, 0.6), 0.05, 0.2, tevo.TEBD4())
    ov = oracle.inner(o, oracle.MPS(p.tensors()))
    assert abs(abs(ov) - 1.0) <= TOL
    np.testing.assert_allclose(p.bform_expect_all("Sz"), oracle.expect_all(o, "Sz"), atol=TOL)
    ent = [float(-np.sum((l ** 2)[l ** 2 > 1e-13] * np.log((l ** 2)[l ** 2 > 1e-13]))) for l in
           (p.schmidt_values(j) for j in range(1, N))]
    np.testing.assert_allclose(ent, oracle.bond_entropies(o), atol=TOL)


def test_device_segment_single_rank_matches_direct(tevo, oracle):
    """DeviceSegment + ShardedTEBD with world_size 1 (no exchange) == direct layer-parallel TEBD."""
    from timeevomps_jl_b200.sharded import DeviceSegment, ShardedTEBD
    N = 8
    o = oracle.MPS.random(N, 5, seed=9)
    a = tevo.MPS.from_tensors(o.A, llim=o.llim, rlim=o.rlim).to_bform()
    b = tevo.MPS.from_tensors(o.A, llim=o.llim, rlim=o.rlim).to_bform()
    seg = DeviceSegment(a, 0, N, ghost=False)
    sh = ShardedTEBD(seg, N, None, 0, 1, device="cuda")
    H = tevo.gates(tevo.tfi_bondop(N, 1.0, 0.5))
    sh.run(H, 0.05, 2, tevo.time_evo_gates, tevo.TEBD2(), maxdim=6)
    tevo.tebd_parallel(b, tevo.tfi_bondop(N, 1.0, 0.5), 0.05, 0.1, tevo.TEBD2(), maxdim=6)
    np.testing.assert_allclose(sh.gather_expect("Sz"), b.bform_expect_all("Sz"), atol=1e-13)
