"""Pins the CPU oracle (oracle/tevo_oracle.py) against every known-answer check the reference's own test-suite
holds for the hot path, plus dense exact-diagonalisation anchors.
CPU only - runs under `-m "not gpu"`."""
import math

import numpy as np
import pytest
import scipy.linalg as sla

from oracle import tevo_oracle as o


# ---- test/test_bondop.jl ------------------------------------------------------------------------
def test_siteterm_products():
    assert np.array_equal(o.op_matrix("Sx*Sx") * 4, o.op_matrix("Id"))          # test_bondop.jl:7-8
    assert np.array_equal(4 * o.op_matrix("Sx*Sz*Sx"), -1 * o.op_matrix("Sz"))  # test_bondop.jl:9-10


def test_bond_operator_gates_match_hand_built():
    N, J, h = 10, -1, 0.7
    H = o.BondOperator(N)
    for b in range(N - 1):
        H.add(J, "Sz", "Sz", b)
        H.add(h, "Sx", b)
    H.add(h, "Sx", N - 1)
    I, Sz, Sx = o.op_matrix("Id"), o.op_matrix("Sz"), o.op_matrix("Sx")
    fac = lambda i: 1.0 if (i == 0 or i == N - 1) else 0.5
    for b in range(N - 1):  # test_bondop.jl:16-21,45-47 (exact ==)
        G = J * np.kron(Sz, Sz) + h * (fac(b) * np.kron(Sx, I) + fac(b + 1) * np.kron(I, Sx))
        assert np.array_equal(H.bondgate(b).G, G)
    H2 = o.BondOperator(N)
    H2.add(1j, "Sx", 0)
    assert H2.siteterms[0] == [(1j, "Sx")]
    H2.add(1.0, "Sz", "Sx", 0)
    assert H2.bondterms[0] == [(1.0, "Sz", "Sx")]


# ---- truncate! -------------------------------------------------------------------------------------
def test_truncate_rules():
    P = np.array([0.5, 0.3, 0.1, 0.06, 0.03, 0.01])
    assert o.truncate_probs(P, maxdim=3)[0] == 3
    n, err = o.truncate_probs(P, maxdim=3)
    assert abs(err - 0.1) < 1e-15
    assert o.truncate_probs(P, cutoff=0.05)[0] == 4           # 0.01+0.03 <= 0.05 < 0.01+0.03+0.06
    assert o.truncate_probs(P, cutoff=0.05, mindim=5)[0] == 5
    assert o.truncate_probs(P, cutoff=0.05, use_absolute_cutoff=True)[0] == 4
    assert o.truncate_probs(P, cutoff=0.0) == (6, 0.0)
    assert o.truncate_probs(np.array([0.7, 0.3, 0.0, 0.0]), maxdim=10)[0] == 2  # exact zeros dropped at cutoff 0
    assert o.truncate_probs(np.array([0.0, 0.0]))[0] == 1


# ---- test/test_tebd.jl -----------------------------------------------------------------------------
ALGS = [o.TEBD2, o.TEBD2Sweep, o.TEBD4]


@pytest.mark.parametrize("alg", ALGS)
def test_tebd_identity_and_reversibility(alg):
    N = 10
    psi0 = o.MPS.random(N, 4, seed=1)
    psi = psi0.copy()
    o.tebd(psi, o.BondOperator(N), 0.01, 1.0, alg())
    assert abs(o.inner(psi0, psi) - 1) < 1.5e-8                   # test_tebd.jl:14-16
    H = o.tfi_bondop(N, 0.3, -0.7)
    psi0 = o.MPS.random(N, 2, seed=2)
    psi = psi0.copy()
    o.tebd(psi, H, 0.01, 1.0, alg())
    o.tebd(psi, H, -0.01, -1.0, alg())
    assert abs(o.inner(psi0, psi) - 1) < 1.5e-8                   # test_tebd.jl:25-28


@pytest.mark.parametrize("alg", ALGS)
def test_tebd_bond_dimension_grows_to_maximum(alg):
    N = 10
    psi = o.MPS.product([0] * N)
    o.tebd(psi, o.tfi_bondop(N, 0.3, -0.7), 0.01, 1.0 if alg is o.TEBD4 else 2.0, alg())
    assert psi.maxlinkdim() == 2 ** 5                             # test_tebd.jl:31-33 (no truncation w/o kwargs)


def test_tebd_norm_with_truncation():
    N = 20
    psi = o.MPS.product([0] * N)
    o.tebd(psi, o.tfi_bondop(N, 0.3, -0.7), 0.1, 20, o.TEBD2(), maxdim=5)
    assert abs(o.inner(psi, psi) - 1.0) < 1e-10                   # test_tebd.jl:39-46


@pytest.mark.parametrize("alg", [o.TEBD2, o.TEBD4])
def test_tebd_imaginary_time_ground_state_energy(alg):
    N, J, h = 10, 1.0, 0.5
    psi = o.MPS.product([0] * N)
    H = o.tfi_bondop(N, J, h)
    for dt in (0.1, 0.01):
        o.tebd(psi, H, -1j * dt, -1j * 500 * dt, alg(), maxdim=50, cutoff=1e-8, orthogonalize=10)
    E = o.measure_gates(H.gates(), psi)
    assert abs(E.real - o.tfi_exact_critical_energy(N)) < 1e-4   # test_tebd.jl:50-72
    assert abs(o.tfi_exact_critical_energy(N) - (-3.0953724999137)) < 1e-10


def test_tebd2_vs_tebd4():
    N, J, h, dt, tf = 10, 1.0, 5.87, 0.01, 2
    psi2 = o.MPS.product([0] * N)
    psi4 = psi2.copy()
    H = o.tfi_bondop(N, J, h)
    o.tebd(psi2, H, dt, tf, cutoff=1e-12)
    o.tebd(psi4, H, dt, tf, o.TEBD4(), cutoff=1e-12)
    assert abs(o.inner(psi2, psi4) - 1) < dt ** 2                 # test_tebd.jl:90
    sz2, sz4 = o.expect_all(psi2, "Sz"), o.expect_all(psi4, "Sz")
    assert np.max(np.abs(sz2 - sz4)) < dt ** 2                    # test_tebd.jl:95
    psi4 = o.MPS.product([0] * N)
    o.tebd(psi4, H, 0.1, tf, o.TEBD4())
    assert abs(o.inner(psi2, psi4) - 1) < dt ** 2                 # test_tebd.jl:103
    assert np.max(np.abs(sz2 - o.expect_all(psi4, "Sz"))) < dt ** 2


def test_incommensurate_steps_throw():
    psi = o.MPS.product([0] * 4)
    with pytest.raises(ValueError):
        o.tebd(psi, o.tfi_bondop(4, 1, 1), 0.1, 0.35)            # Int(tf/dt) InexactError, tebd.jl:95
    with pytest.raises(ValueError):
        o.tebd(psi, o.tfi_bondop(4, 1, 1), 0.1, 1.0, callback=o.LocalMeasurementCallback(["Sz"], 4, 0.25))


# ---- dense ED anchors ----------------------------------------------------------
def test_exact_diagonalisation_anchor():
    N = 10
    Hd = o.mpo_to_dense(o.tfi_mpo(N, 1.0, 0.5))
    v0 = np.zeros(2 ** N, complex)
    v0[0] = 1
    psi = o.MPS.product([0] * N)
    o.tebd(psi, o.tfi_bondop(N, 1.0, 0.5), 0.01, 1.0, o.TEBD4(), cutoff=1e-14)
    ve = sla.expm(-1j * Hd) @ v0
    assert abs(abs(np.vdot(ve, psi.to_dense())) - 1) < 1e-9
    sz = o.expect_all(psi, "Sz")
    assert abs(sz[0] - 0.440050585745) < 1e-8 and abs(sz[4] - 0.443547817926) < 1e-8
    assert abs(o.bond_entropies(psi)[4] - 0.003183460651) < 1e-6  # Trotter error of dt=0.01 TEBD4


# ---- test/test_tdvp.jl -----------------------------------------------------------------------------
def test_tdvp_reversibility_and_energy():
    N = 10
    W = o.tfi_mpo(N, 1.0, 0.5)
    rng = np.random.default_rng(5)
    psi0 = o.MPS.product([int(rng.random() > 0.5) for _ in range(N)])
    psi = psi0.copy()
    dt, nsteps = 0.1, 20
    o.tdvp(psi, W, dt, nsteps * dt, exp_tol=1e-12 / dt, cutoff=1e-14)
    E1 = o.inner_mpo(psi, W, psi)
    assert abs(E1 - o.inner_mpo(psi0, W, psi0)) < 1.5e-8 * max(1, abs(E1))   # test_tdvp.jl:29-33
    o.tdvp(psi, W, -dt, -nsteps * dt, exp_tol=1e-12 / dt, cutoff=1e-14)
    assert abs(o.inner(psi, psi) - 1) < 1.5e-8                                 # test_tdvp.jl:20
    assert abs(o.inner(psi0, psi) - 1) < 1.5e-8                                # test_tdvp.jl:21


def test_tdvp_norm_with_truncation():
    N = 20
    psi = o.MPS.product([0] * N)
    o.tdvp(psi, o.tfi_mpo(N, 1.0, 0.5), 1, 10, maxdim=5, exp_tol=1e-12)
    assert abs(o.inner(psi, psi) - 1.0) < 1e-10                                # test_tdvp.jl:36-42


def test_tdvp_vs_tebd():
    N, J, h, dt, tf = 10, 1.0, 5.87, 0.01, 0.5
    a = o.MPS.product([0] * N)
    b = a.copy()
    o.tdvp(a, o.tfi_mpo(N, J, h), dt, tf, cutoff=1e-12, exp_tol=1e-12)
    o.tebd(b, o.tfi_bondop(N, J, h), dt, tf, cutoff=1e-12)
    assert abs(o.inner(a, b) - 1) < dt ** 2                                     # test_tdvp.jl:64
    assert np.max(np.abs(o.expect_all(a, "Sz") - o.expect_all(b, "Sz"))) < dt ** 2  # :69


def test_tdvp_imaginary_time_energy():
    N = 10
    W = o.tfi_mpo(N, 1.0, 0.5)
    psi = o.MPS.product([0] * N)
    dt = 0.1
    o.tdvp(psi, W, -1j * dt, -1j * 100 * dt, maxdim=50, exp_tol=1e-14 / dt)
    E = o.inner_mpo(psi, W, psi)
    assert abs(E.real - o.tfi_exact_critical_energy(N)) < 1e-4                 # test_tdvp.jl:75-94


def test_tdvp_free_fermions():
    N, dt, tf = 20, 0.1, 2
    W = o.hopping_mpo(N)
    psi = o.MPS.product([1 if i % 2 == 0 else 0 for i in range(N)])  # odd (1-based) sites "Occ"
    cb = o.LocalMeasurementCallback(["N"], N, 0.1)
    o.tdvp(psi, W, dt, tf, cutoff=1e-10, callback=cb, exp_tol=1e-12)
    ns_exact = o.free_fermions_densities(N, 0.1, tf)
    ns = cb.measurements["N"]
    assert len(ns) == 20
    for i in range(len(ns)):
        assert np.all(np.abs(ns[i] - ns_exact[i + 1]) < 5e-5)                  # test_tdvp.jl:96-117


# ---- test/test_callbacks.jl ------------------------------------------------------------------------
def test_no_callback():
    cb = o.NoTEvoCallback()
    assert cb.apply(None, t=0, bond=1, sweepend=True) is None
    assert cb.checkdone(None) is False and cb.callback_dt() == 0


def test_local_measurement_gating():
    N = 10
    psi = o.MPS.random(N, 4, seed=3)
    cb = o.LocalMeasurementCallback(["Sz", "Sx"], N, 0.1)
    assert cb.callback_dt() == 0.1 and len(cb.ts) == 0
    for (t, b, se) in [(0.03, 0, True), (0.1, 0, False)]:           # test_callbacks.jl:21-24
        cb.apply(psi, t=t, bond=b, sweepend=se, sweepdir="right", alg=o.TEBD2())
        assert len(cb.ts) == 0
    psi.orthogonalize(2)
    cb.apply(psi, t=0.1, bond=2, sweepend=True, sweepdir="left", alg=o.TEBD2())
    assert len(cb.ts) == 1
    for name in ("Sz", "Sx"):
        assert np.count_nonzero(cb.measurements[name][-1]) == 2    # :29-32
    cb = o.LocalMeasurementCallback(["Sz", "Sx"], N, 0.1)
    psi.orthogonalize(0)
    cb.apply(psi, t=0.1, bond=0, sweepend=True, sweepdir="left", alg=o.TDVP2())
    for name in ("Sz", "Sx"):
        assert np.count_nonzero(cb.measurements[name][-1]) == 2    # :37-40
    cb = o.LocalMeasurementCallback(["Sz", "Sx"], N, 0.1)
    psi.orthogonalize(2)
    cb.apply(psi, t=0.1, bond=2, sweepend=True, sweepdir="left", alg=o.TDVP2())
    for name in ("Sz", "Sx"):
        assert np.count_nonzero(cb.measurements[name][-1]) == 1    # :45-48


def test_local_measurement_during_evolution():
    N = 10
    psi = o.MPS.product([0] * N)
    cb = o.LocalMeasurementCallback(["Sz", "Sx"], N, 0.5)
    o.tebd(psi, o.tfi_bondop(N, 1.0, 1.0), 0.1, 5, o.TEBD2(), maxdim=30, callback=cb)
    assert len(cb.ts) == 10                                          # test_callbacks.jl:55
    np.testing.assert_allclose(cb.measurements["Sz"][-1], o.expect_all(psi, "Sz"), rtol=1.5e-8, atol=1e-12)
    psi = o.MPS.product([0] * N)
    cb2 = o.LocalMeasurementCallback(["Sz", "Sx"], N, 0.5)
    o.tdvp(psi, o.tfi_mpo(N, 1.0, 1.0), 0.1, 5, maxdim=30, callback=cb2, exp_tol=1e-12)
    assert len(cb2.ts) == 10                                         # :70
    np.testing.assert_allclose(cb2.measurements["Sz"][-1], o.expect_all(psi, "Sz"), rtol=1.5e-8, atol=1e-12)


def test_spec_callback():
    N = 10
    psi = o.MPS.product([0] * N)
    cb = o.SpecCallback(0.1, N)
    o.tebd(psi, o.tfi_bondop(N, 1.0, 1.0), 0.1, 5, o.TEBD2(), maxdim=30, callback=cb)
    assert list(cb.bonddims[-1]) == psi.bond_dims()                  # test_callbacks.jl:87-89
    np.testing.assert_allclose(cb.entropies[-1], o.bond_entropies(psi), rtol=1.5e-8, atol=1e-12)  # :90-98
    psi = o.MPS.product([0] * N)
    cb = o.SpecCallback(0.1, N)
    o.tdvp(psi, o.tfi_mpo(N, 1.0, 1.0), 0.1, 5, maxdim=30, callback=cb, exp_tol=1e-12)
    assert list(cb.bonddims[-1]) == psi.bond_dims()                  # :105-107
    np.testing.assert_allclose(cb.entropies[-1], o.bond_entropies(psi), rtol=1.5e-8, atol=1e-12)
    with pytest.raises(ValueError):
        o.SpecCallback(0.1, N, bonds=[0, N - 1])                     # callback.jl:114-116


def test_exponentiate_krylovkit_literal_mode_gives_the_same_vector():
    """The oracle (and the CUDA path) test KrylovKit's error estimate at every Krylov dimension; KrylovKit 0.4.0
    (benchmarks/Manifest.toml:58-62) expands to the full dimension first.  Both control flows integrate the same ODE
    to the same tolerance: the vectors agree to exp_tol, also across restarts (long times) and for the Arnoldi branch."""
    import scipy.linalg as sla
    rng = np.random.default_rng(0)
    n = 300
    M = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    H = 3.0 * (M + M.conj().T) / 2 / np.sqrt(n)
    v = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    v /= np.linalg.norm(v)
    for t, herm, A in ((-0.05j, True, H), (-1.0j, True, H), (-6.0j, True, H), (0.3, True, H), (-0.5j, False, H + 0.1 * M / np.sqrt(n))):
        ref = sla.expm(t * A) @ v
        out = {}
        for mode in ("early", "krylovkit"):
            w, info = o.exponentiate(lambda x: A @ x, t, v, tol=1e-12, krylovdim=30, ishermitian=herm, mode=mode)
            assert info.converged == 1
            assert np.linalg.norm(w - ref) < 5e-11 * max(1.0, abs(t))
            out[mode] = (w, info)
        assert np.linalg.norm(out["early"][0] - out["krylovkit"][0]) < 5e-11 * max(1.0, abs(t))
        assert out["early"][1].numops <= out["krylovkit"][1].numops


def test_tdvp_krylovkit_literal_mode_same_observables():
    """One TDVP step of the TFIM chain with both Krylov control flows: gauge-invariant results agree to 1e-10."""
    N = 8
    res = {}
    for mode in ("early", "krylovkit"):
        p = o.MPS.random(N, 6, seed=5)
        st = {}
        o.tdvp(p, o.tfi_mpo(N, 1.0, 0.7), 0.1, 0.2, maxdim=8, exp_tol=1e-12, krylov_mode=mode, stats=st)
        res[mode] = (p, st)
    a, b = res["early"][0], res["krylovkit"][0]
    assert a.bond_dims() == b.bond_dims()
    assert abs(abs(o.inner(a, b)) - 1.0) < 1e-10
    np.testing.assert_allclose(o.expect_all(a, "Sz"), o.expect_all(b, "Sz"), atol=1e-10)
    assert res["early"][1]["matvec2"] <= res["krylovkit"][1]["matvec2"]
