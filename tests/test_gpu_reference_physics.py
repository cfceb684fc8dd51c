"""The reference's own physics-level tests (test/test_tebd.jl, test/test_tdvp.jl, test/test_callbacks.jl) re-run on
the CUDA path through the public host API - same assertions, same tolerances (shortened evolutions where noted)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _algs(t):
    return [t.TEBD2, t.TEBD2Sweep, t.TEBD4]


@pytest.mark.parametrize("ialg", [0, 1, 2])
def test_tebd_identity_and_reversibility(tevo, oracle, ialg):
    alg = _algs(tevo)[ialg]
    N = 10
    o0 = oracle.MPS.random(N, 4, seed=1)
    psi0 = tevo.MPS.from_tensors(o0.A, llim=-1, rlim=1)
    psi = psi0.copy()
    tevo.tebd(psi, tevo.BondOperator(N), 0.01, 0.2, alg())
    assert abs(tevo.inner(psi0, psi) - 1) < 1.5e-8                       # test_tebd.jl:14-16 (tf shortened)
    H = tevo.tfi_bondop(N, 0.3, -0.7)
    psi = psi0.copy()
    tevo.tebd(psi, H, 0.01, 0.2, alg())
    tevo.tebd(psi, H, -0.01, -0.2, alg())
    assert abs(tevo.inner(psi0, psi) - 1) < 1.5e-8                       # test_tebd.jl:25-28


def test_tebd_bond_dimension_grows_without_truncation(tevo):
    N = 10
    psi = tevo.MPS.product([0] * N)
    tevo.tebd(psi, tevo.tfi_bondop(N, 0.3, -0.7), 0.01, 0.5, tevo.TEBD2())
    assert psi.maxlinkdim() == 2 ** 5                                     # test_tebd.jl:31-33


def test_tebd_norm_with_truncation(tevo):
    N = 20
    psi = tevo.MPS.product([0] * N)
    tevo.tebd(psi, tevo.tfi_bondop(N, 0.3, -0.7), 0.1, 5, tevo.TEBD2(), maxdim=5)
    assert abs(tevo.inner(psi, psi) - 1.0) < 1e-10                        # test_tebd.jl:39-46 (tf shortened)


def test_tebd_imaginary_time_energy(tevo, oracle):
    N = 10
    psi = tevo.MPS.product([0] * N)
    H = tevo.tfi_bondop(N, 1.0, 0.5)
    for dt, nst in ((0.1, 500), (0.01, 500)):
        tevo.tebd(psi, H, -1j * dt, -1j * nst * dt, tevo.TEBD2(), maxdim=50, cutoff=1e-8, orthogonalize=10)
    E = tevo.measure(tevo.gates(H), psi)
    assert abs(E.real - oracle.tfi_exact_critical_energy(N)) < 1e-4      # test_tebd.jl:50-72


def test_tebd2_vs_tebd4(tevo):
    N, J, h, dt, tf = 10, 1.0, 5.87, 0.01, 0.5
    psi2 = tevo.MPS.product([0] * N)
    psi4 = tevo.MPS.product([0] * N)
    H = tevo.tfi_bondop(N, J, h)
    tevo.tebd(psi2, H, dt, tf, cutoff=1e-12)
    tevo.tebd(psi4, H, dt, tf, tevo.TEBD4(), cutoff=1e-12)
    assert abs(tevo.inner(psi2, psi4) - 1) < dt ** 2                      # test_tebd.jl:90
    assert np.max(np.abs(psi2.expect_all("Sz") - psi4.expect_all("Sz"))) < dt ** 2   # :95


def test_tebd_argument_errors(tevo):
    psi = tevo.MPS.product([0] * 4)
    with pytest.raises(ValueError):
        tevo.tebd(psi, tevo.tfi_bondop(4, 1, 1), 0.1, 0.35)              # Int(tf/dt), tebd.jl:95
    with pytest.raises(ValueError):
        tevo.tebd(psi, tevo.tfi_bondop(4, 1, 1), 0.1, 1.0, callback=tevo.LocalMeasurementCallback(["Sz"], 4, 0.25))
    with pytest.raises(tevo.TevoError):
        tevo.apply_gate(psi, tevo.BondGate(np.eye(4), 7))                 # bond out of range
    with pytest.raises(tevo.TevoError):
        psi.orthogonalize(9)


def test_tdvp_reversibility_and_energy(tevo, oracle):
    N = 10
    H = tevo.tfi_mpo(1.0, 0.5, N)
    rng = np.random.default_rng(5)
    states = [int(rng.random() > 0.5) for _ in range(N)]
    psi0 = tevo.MPS.product(states)
    psi = psi0.copy()
    dt, nsteps = 0.1, 5
    E0 = tevo.inner_mpo(psi0, H, psi0)
    tevo.tdvp(psi, H, dt, nsteps * dt, exp_tol=1e-12 / dt, cutoff=1e-14)
    E1 = tevo.inner_mpo(psi, H, psi)
    assert abs(E1 - E0) < 1.5e-8 * max(1, abs(E0))                        # test_tdvp.jl:29-33
    tevo.tdvp(psi, H, -dt, -nsteps * dt, exp_tol=1e-12 / dt, cutoff=1e-14)
    assert abs(tevo.inner(psi, psi) - 1) < 1.5e-8                         # test_tdvp.jl:20
    assert abs(tevo.inner(psi0, psi) - 1) < 1.5e-8                        # test_tdvp.jl:21


def test_tdvp_norm_with_truncation(tevo):
    N = 20
    psi = tevo.MPS.product([0] * N)
    tevo.tdvp(psi, tevo.tfi_mpo(1.0, 0.5, N), 1, 3, maxdim=5, exp_tol=1e-12)
    assert abs(tevo.inner(psi, psi) - 1.0) < 1e-10                        # test_tdvp.jl:36-42 (tf shortened)


def test_tdvp_vs_tebd(tevo):
    N, J, h, dt, tf = 10, 1.0, 5.87, 0.01, 0.2
    a = tevo.MPS.product([0] * N)
    b = tevo.MPS.product([0] * N)
    tevo.tdvp(a, tevo.tfi_mpo(J, h, N), dt, tf, cutoff=1e-12, exp_tol=1e-12)
    tevo.tebd(b, tevo.tfi_bondop(N, J, h), dt, tf, cutoff=1e-12)
    assert abs(tevo.inner(a, b) - 1) < dt ** 2                             # test_tdvp.jl:64
    assert np.max(np.abs(a.expect_all("Sz") - b.expect_all("Sz"))) < dt ** 2   # :69


def test_tdvp_imaginary_time_energy(tevo, oracle):
    N = 10
    H = tevo.tfi_mpo(1.0, 0.5, N)
    psi = tevo.MPS.product([0] * N)
    dt = 0.1
    tevo.tdvp(psi, H, -1j * dt, -1j * 60 * dt, maxdim=50, exp_tol=1e-13)
    E = tevo.inner_mpo(psi, H, psi)
    assert abs(E.real - oracle.tfi_exact_critical_energy(N)) < 1e-4      # test_tdvp.jl:75-94 (60 steps)


def test_tdvp_free_fermions(tevo, oracle):
    N, dt, tf = 12, 0.1, 0.5
    W = tevo.MPO(oracle.hopping_mpo(N))
    psi = tevo.MPS.product([1 if i % 2 == 0 else 0 for i in range(N)])
    cb = tevo.LocalMeasurementCallback(["N"], N, 0.1)
    tevo.tdvp(psi, W, dt, tf, cutoff=1e-10, callback=cb, exp_tol=1e-12)
    ns_exact = oracle.free_fermions_densities(N, 0.1, tf)
    ns = cb.measurements["N"]
    assert len(ns) == 5
    for i in range(len(ns)):
        assert np.all(np.abs(ns[i] - ns_exact[i + 1]) < 5e-5)             # test_tdvp.jl:96-117 (N=12, tf=0.5)


def test_tdvp_arnoldi_branch_matches_lanczos_and_oracle(tevo, oracle):
    """hermitian=false (KrylovKit Arnoldi, src/tdvp.jl:25-27,40): same evolution as the Lanczos branch for a
    Hermitian H, and parity with the oracle's Arnoldi restatement."""
    N = 8
    o0 = oracle.MPS.random(N, 3, seed=4)
    a = tevo.MPS.from_tensors(o0.A, llim=-1, rlim=1)
    b = tevo.MPS.from_tensors(o0.A, llim=-1, rlim=1)
    H = tevo.xxz_mpo(N, 0.7, 0.2)
    tevo.tdvp(a, H, 0.05, 0.1, maxdim=8, exp_tol=1e-12, hermitian=True)
    tevo.tdvp(b, H, 0.05, 0.1, maxdim=8, exp_tol=1e-12, hermitian=False)
    assert a.bond_dims() == b.bond_dims()
    assert abs(abs(tevo.inner(a, b)) - 1.0) < 1e-10
    oracle.tdvp(o0, oracle.xxz_mpo(N, 0.7, 0.2), 0.05, 0.1, maxdim=8, exp_tol=1e-12, hermitian=False)
    np.testing.assert_allclose(b.copy().expect_all("Sz"), oracle.expect_all(o0, "Sz"), atol=1e-10)


def test_callbacks_during_evolution(tevo, oracle):
    N = 10
    psi = tevo.MPS.product([0] * N)
    cb = tevo.LocalMeasurementCallback(["Sz", "Sx"], N, 0.5)
    tevo.tebd(psi, tevo.tfi_bondop(N, 1.0, 1.0), 0.1, 2, tevo.TEBD2(), maxdim=30, callback=cb)
    assert len(cb.ts) == 4                                                # test_callbacks.jl:55 (tf=2)
    np.testing.assert_allclose(cb.measurements["Sz"][-1], psi.copy().expect_all("Sz"), rtol=1.5e-8, atol=1e-12)
    psi = tevo.MPS.product([0] * N)
    sc = tevo.SpecCallback(0.1, N)
    tevo.tebd(psi, tevo.tfi_bondop(N, 1.0, 1.0), 0.1, 1, tevo.TEBD2(), maxdim=30, callback=sc)
    assert list(sc.bonddims[-1]) == psi.bond_dims()                        # test_callbacks.jl:87-89
    np.testing.assert_allclose(sc.entropies[-1], oracle.bond_entropies(oracle.MPS(psi.tensors())), rtol=1.5e-8,
                               atol=1e-12)                                 # :90-98
    psi = tevo.MPS.product([0] * N)
    sc = tevo.SpecCallback(0.1, N)
    tevo.tdvp(psi, tevo.tfi_mpo(1.0, 1.0, N), 0.1, 1, maxdim=30, callback=sc, exp_tol=1e-12)
    assert list(sc.bonddims[-1]) == psi.bond_dims()                        # :105-107
    np.testing.assert_allclose(sc.entropies[-1], oracle.bond_entropies(oracle.MPS(psi.tensors())), rtol=1.5e-8,
                               atol=1e-12)


def test_custom_callback_gets_coherent_state(tevo, oracle):
    """A user callback (needs_state=True) is fired after every gate with a state it can measure."""
    N = 6

    class Norms(tevo.TEvoCallback):
        def __init__(self):
            self.seen = []

        def apply(self, psi, *, t, bond, spec, **kw):
            self.seen.append((bond, abs(tevo.inner(psi, psi)), len(spec)))

    cb = Norms()
    psi = tevo.MPS.product([0] * N)
    tevo.tebd(psi, tevo.tfi_bondop(N, 1.0, 0.5), 0.05, 0.1, tevo.TEBD2(), maxdim=8, callback=cb)
    assert len(cb.seen) > 0
    assert all(abs(n - 1.0) < 1e-10 for (_, n, _) in cb.seen)
