"""GPU parity of the TEBD path (apply_gate!/apply_gates!/tebd!) against the CPU oracle, through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-10  # north-star tolerance on gauge-invariant observables at fixed truncation


def _to_dev(tevo, opsi):
    return tevo.MPS.from_tensors(opsi.A, llim=opsi.llim, rlim=opsi.rlim)


def _dense(tensors):
    v = tensors[0]
    for t in tensors[1:]:
        v = np.tensordot(v, t, axes=(v.ndim - 1, 0))
    return v.reshape(-1)


def test_roundtrip_and_inner(tevo, oracle):
    o = oracle.MPS.random(8, 6, seed=1)
    p = _to_dev(tevo, o)
    for a, b in zip(o.A, p.tensors()):
        assert np.array_equal(a, b)
    # download into caller-owned buffers (the bench's end-to-end path hands pinned host memory here)
    bufs = [np.empty(a.shape, dtype=np.complex128, order="F") for a in o.A]
    bufs[3] = np.empty((1, 2, 1), dtype=np.complex128, order="F")      # wrong shape: a fresh array is returned instead
    outs = p.tensors(out=bufs)
    for j, (a, b) in enumerate(zip(o.A, outs)):
        assert np.array_equal(a, b)
        assert (b is bufs[j]) == (j != 3)
    assert abs(tevo.inner(p, p) - oracle.inner(o, o)) < 1e-13
    o2 = oracle.MPS.random(8, 5, seed=2)
    p2 = _to_dev(tevo, o2)
    assert abs(tevo.inner(p, p2) - oracle.inner(o, o2)) < 1e-13


def test_orthogonalize_preserves_state(tevo, oracle):
    o = oracle.MPS.random(7, 8, seed=4, canonical=False)
    p = _to_dev(tevo, o)
    ref = _dense(o.A)
    for j in (3, 0, 6, 2):
        p.orthogonalize(j)
        assert p.limits() == (j - 1, j + 1)
        got = _dense(p.tensors())
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
        ts = p.tensors()
        for i in range(j):
            A = ts[i].reshape(-1, ts[i].shape[2])
            assert np.abs(A.conj().T @ A - np.eye(A.shape[1])).max() < 1e-10  # Gram-route centre move: ~eps*cond^2
        for i in range(j + 1, 7):
            A = ts[i].reshape(ts[i].shape[0], -1)
            assert np.abs(A @ A.conj().T - np.eye(A.shape[0])).max() < 1e-10


@pytest.mark.parametrize("spread", [1.0, 30.0])
@pytest.mark.parametrize("chi", [5, 37, 63, 64, 65, 100, 128, 130, 193, 200, 320, 449])
def test_centre_moves_across_cholesky_block_sizes(tevo, chi, spread):
    """The centre moves factor a D x D Gram matrix in 64-wide diagonal blocks (csrc/cholqr.cu): bond dimensions below,
    at and just above block boundaries.  Chain with d = 4 and bond dimensions 1, 2, 4, ..., chi, ..., 4, 2, 1 of generic
    Gaussian tensors: every site matrix is at least 2:1 tall in the direction it is factored, so the moves stay well
    conditioned (cond <= 10; the Gram route loses eps * cond^2) while the triangular factors are far from the
    identity.  `spread` grades the chi bond by 1..spread, which pushes the pivot ratio of that move above the
    single-pass threshold and exercises the second pass.  No oracle needed: exact isometry of every moved site and
    the overlap with the untouched copy."""
    rng = np.random.default_rng(1000 + chi)
    k = int(np.ceil(np.log2(chi))) - 1
    up = [2 ** i for i in range(k + 1)]
    dims = up + [chi] + up[::-1]
    N, d = len(dims) - 1, 4
    A = []
    for i in range(N):
        a = rng.standard_normal((dims[i], d, dims[i + 1])) + 1j * rng.standard_normal((dims[i], d, dims[i + 1]))
        A.append(a / np.linalg.norm(a) * np.sqrt(max(dims[i], dims[i + 1])))
    mid = dims.index(chi)
    w = np.linspace(1.0, spread, chi)
    A[mid - 1] = A[mid - 1] * w
    A[mid] = A[mid] / w[:, None, None]
    p0 = tevo.MPS.from_tensors([np.asfortranarray(a) for a in A], llim=-1, rlim=N)
    p = p0.copy()
    nrm = tevo.inner(p0, p0).real
    p.orthogonalize(N - 1)
    ts = p.tensors()
    assert [t.shape[2] for t in ts] == dims[1:]
    for i in range(N - 1):
        M = ts[i].reshape(-1, ts[i].shape[2])
        assert np.abs(M.conj().T @ M - np.eye(M.shape[1])).max() < 1e-12, (i, M.shape)
    assert abs(tevo.inner(p0, p) / nrm - 1.0) < 1e-12
    p.orthogonalize(0)
    ts = p.tensors()
    for i in range(1, N):
        M = ts[i].reshape(ts[i].shape[0], -1)
        assert np.abs(M @ M.conj().T - np.eye(M.shape[0])).max() < 1e-12, (i, M.shape)
    assert abs(tevo.inner(p0, p) / nrm - 1.0) < 1e-12
    assert abs(tevo.inner(p, p).real / nrm - 1.0) < 1e-12


@pytest.mark.parametrize("ortho", ["left", "right"])
def test_apply_gate_matches_oracle(tevo, oracle, ortho):
    N = 8
    o = oracle.MPS.random(N, 8, seed=5)
    p = _to_dev(tevo, o)
    H = oracle.tfi_bondop(N, 1.0, 0.7).gates()
    for b in (3, 4, 0, 6, 2):
        g = oracle.gate_exp(H[b], -0.1j)
        ospec = oracle.apply_gate(o, g, maxdim=6, cutoff=1e-9, ortho=ortho)
        spec = tevo.apply_gate(p, tevo.BondGate(g.G, b), maxdim=6, cutoff=1e-9, ortho=ortho)
        assert len(spec) == len(ospec.eigs)
        np.testing.assert_allclose(spec.eigs, ospec.eigs, rtol=TOL)
        assert abs(spec.truncerr - ospec.truncerr) <= TOL * ospec.truncerr + 1e-18
        assert abs(spec.entropy() - ospec.entropy()) <= TOL
        assert p.limits() == (o.llim, o.rlim)
    assert p.bond_dims() == o.bond_dims()
    ov = oracle.inner(o, oracle.MPS(p.tensors()))
    assert abs(abs(ov) - 1.0) <= TOL
    for name in ("Sz", "Sx", "Sy"):
        np.testing.assert_allclose(p.copy().expect_all(name), oracle.expect_all(o, name), atol=TOL)


@pytest.mark.parametrize("algname", ["TEBD2", "TEBD4", "TEBD2Sweep"])
def test_tebd_config1_matches_oracle(tevo, oracle, algname):
    """BASELINE config 1: TFIM N=10 J=1 h=0.5 dt=0.01 maxdim=10 from the polarised product state (tf shortened
    to 0.2 for TEBD4/TEBD2Sweep to bound the run time)."""
    N, J, h = 10, 1.0, 0.5
    tf = 1.0 if algname == "TEBD2" else 0.2
    o = oracle.MPS.product([0] * N)
    oracle.tebd(o, oracle.tfi_bondop(N, J, h), 0.01, tf, getattr(oracle, algname)(), maxdim=10)
    p = tevo.MPS.product([0] * N)
    tevo.tebd(p, tevo.tfi_bondop(N, J, h), 0.01, tf, getattr(tevo, algname)(), maxdim=10)
    assert p.bond_dims() == o.bond_dims()
    assert abs(tevo.inner(p, p) - 1.0) <= TOL
    for name in ("Sz", "Sx", "Sy"):
        np.testing.assert_allclose(p.copy().expect_all(name), oracle.expect_all(o, name), atol=TOL)
    np.testing.assert_allclose(oracle.bond_entropies(oracle.MPS(p.tensors())), oracle.bond_entropies(o), atol=TOL)
    ov = oracle.inner(o, oracle.MPS(p.tensors()))
    assert abs(abs(ov) - 1.0) <= TOL


def test_tebd_callbacks_match_oracle(tevo, oracle):
    """LocalMeasurementCallback + SpecCallback at every callback time (src/callback.jl) vs the oracle."""
    N = 10
    ocb = oracle.LocalMeasurementCallback(["Sz", "Sx", "Sy"], N, 0.1)
    o = oracle.MPS.product([0] * N)
    oracle.tebd(o, oracle.tfi_bondop(N, 1.0, 0.5), 0.05, 0.5, oracle.TEBD2(), maxdim=8, cutoff=1e-10, callback=ocb)
    cb = tevo.LocalMeasurementCallback(["Sz", "Sx", "Sy"], N, 0.1)
    p = tevo.MPS.product([0] * N)
    tevo.tebd(p, tevo.tfi_bondop(N, 1.0, 0.5), 0.05, 0.5, tevo.TEBD2(), maxdim=8, cutoff=1e-10, callback=cb)
    assert len(cb.ts) == len(ocb.ts) == 5
    np.testing.assert_allclose(cb.ts, ocb.ts)
    for name in ("Sz", "Sx", "Sy"):
        np.testing.assert_allclose(np.array(cb.measurements[name]), np.array(ocb.measurements[name]), atol=TOL)
    osc = oracle.SpecCallback(0.1, N)
    o = oracle.MPS.product([0] * N)
    oracle.tebd(o, oracle.tfi_bondop(N, 1.0, 0.5), 0.05, 0.5, oracle.TEBD4(), maxdim=8, cutoff=1e-10, callback=osc)
    sc = tevo.SpecCallback(0.1, N)
    p = tevo.MPS.product([0] * N)
    tevo.tebd(p, tevo.tfi_bondop(N, 1.0, 0.5), 0.05, 0.5, tevo.TEBD4(), maxdim=8, cutoff=1e-10, callback=sc)
    assert [list(x) for x in sc.bonddims] == [list(x) for x in osc.bonddims]
    np.testing.assert_allclose(np.array(sc.entropies), np.array(osc.entropies), atol=TOL)
    np.testing.assert_allclose(sc.truncerrs, osc.truncerrs, rtol=1e-8, atol=1e-16)


def test_measure_energy(tevo, oracle):
    N = 8
    o = oracle.MPS.random(N, 6, seed=9)
    p = _to_dev(tevo, o)
    Hg = oracle.tfi_bondop(N, 1.0, 0.5).gates()
    e_o = oracle.measure_gates(Hg, o)
    e_p = tevo.measure(tevo.gates(tevo.tfi_bondop(N, 1.0, 0.5)), p)
    assert abs(e_o - e_p) <= 1e-12


def test_apply_gate_spin_one_sites(tevo, oracle):
    """d = 3 (spin-1) sites: random two-site unitaries, truncated split vs the oracle (generality row of SURVEY 8(f))."""
    N, d = 6, 3
    rng = np.random.default_rng(17)
    o = oracle.MPS.random(N, 9, d=d, seed=6)
    p = tevo.MPS.from_tensors(o.A, llim=o.llim, rlim=o.rlim)
    for b in (2, 3, 0, 4, 1):
        X = rng.standard_normal((d * d, d * d)) + 1j * rng.standard_normal((d * d, d * d))
        G, _ = np.linalg.qr(X)
        ospec = oracle.apply_gate(o, oracle.BondGate(G, b), maxdim=7, cutoff=1e-9)
        spec = tevo.apply_gate(p, tevo.BondGate(G, b), maxdim=7, cutoff=1e-9)
        assert len(spec) == len(ospec.eigs)
        np.testing.assert_allclose(spec.eigs, ospec.eigs, rtol=TOL)
        assert abs(spec.truncerr - ospec.truncerr) <= TOL * ospec.truncerr + 1e-18
    assert p.bond_dims() == o.bond_dims()
    ov = oracle.inner(o, oracle.MPS(p.tensors()))
    assert abs(abs(ov) - 1.0) <= TOL
