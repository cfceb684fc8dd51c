"""Generates the golden vectors of tests/golden/*.npz from the CPU oracle (oracle/tevo_oracle.py) with fixed seeds.
The reference itself holds no golden vectors and cannot be executed here (no Julia), so these pin the oracle - and
through it the CUDA path - against accidental change.  Run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import tevo_oracle as o  # noqa: E402


def pack_state(psi):
    return {"site%d" % j: a for j, a in enumerate(psi.A)}


def case_tebd(name, alg, N, chi0, seed, dt, tf, model, **kw):
    psi = o.MPS.random(N, chi0, seed=seed)
    init = pack_state(psi)
    H = o.tfi_bondop(N, 1.0, 0.5) if model == "tfim" else o.xxz_bondop(N, 1.0, 0.3)
    sc = o.SpecCallback(dt, N)
    o.tebd(psi, H, dt, tf, getattr(o, alg)(), callback=sc, **kw)
    out = dict(init)
    out.update(N=N, dt=dt, tf=tf, alg=alg, model=model, kw=str(kw),
               bond_dims=np.array(psi.bond_dims()), sz=o.expect_all(psi, "Sz"), sx=o.expect_all(psi, "Sx"),
               sy=o.expect_all(psi, "Sy"), entropies=np.array(o.bond_entropies(psi)),
               norm=abs(o.inner(psi, psi)), spec_entropies=np.array(sc.entropies), spec_bonddims=np.array(sc.bonddims),
               truncerrs=np.array(sc.truncerrs))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def case_tdvp(name, N, chi0, seed, dt, tf, **kw):
    psi = o.MPS.random(N, chi0, seed=seed)
    init = pack_state(psi)
    W = o.xxz_mpo(N, 1.0, 0.3)
    cb = o.LocalMeasurementCallback(["Sz", "Sx", "Sy"], N, dt)
    o.tdvp(psi, W, dt, tf, callback=cb, exp_tol=1e-12, **kw)
    out = dict(init)
    out.update(N=N, dt=dt, tf=tf, kw=str(kw), bond_dims=np.array(psi.bond_dims()),
               sz=np.array(cb.measurements["Sz"]), sx=np.array(cb.measurements["Sx"]), sy=np.array(cb.measurements["Sy"]),
               ts=np.array(cb.ts), entropies=np.array(o.bond_entropies(psi)), energy=o.inner_mpo(psi, W, psi))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def case_split():
    rng = np.random.default_rng(42)
    m, n = 24, 20
    U, _ = np.linalg.qr(rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m)))
    V, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    s = np.exp(-0.45 * np.arange(n))
    th = (U[:, :n] * s) @ V.conj().T
    np.savez_compressed(os.path.join(HERE, "split_24x20.npz"), theta=th, sigma2=s ** 2)


if __name__ == "__main__":
    case_tebd("tebd4_tfim_n8", "TEBD4", 8, 4, 1, 0.05, 0.2, "tfim", maxdim=8, cutoff=1e-9)
    case_tebd("tebd2_xxz_n9", "TEBD2", 9, 3, 2, 0.05, 0.2, "xxz", maxdim=6)
    case_tdvp("tdvp_xxz_n8", 8, 3, 3, 0.05, 0.2, maxdim=8, cutoff=1e-10)
    case_split()
    print("golden vectors written to", HERE)
