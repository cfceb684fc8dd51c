"""CPU oracle for the TimeEvoMPS.jl two-site-update hot path.  TEST INFRASTRUCTURE ONLY.

This file is a NumPy complex128 restatement of the reference algorithm
(orialb/TimeEvoMPS.jl, Julia) in the reference's exact operation order.  It is the
checker for the CUDA path; it is NOT part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The product package never imports anything from ``oracle/``.

PARITY STATUS: *parity unpinned vs ITensors.jl*.  The reference cannot be executed here
(no Julia, ITensors.jl/KrylovKit.jl not vendored) and its own tests hold no golden vectors,
only physics-level known answers (norms, reversibility, ground-state energy, free-fermion
densities, exact 4x4 gate matrices).  This oracle is pinned against *every one of those*
(tests/test_oracle_reference_tests.py) and against dense exact diagonalisation, and the
vectors it generates are committed under tests/golden/.

Conventions
-----------
* Site tensors are ``A[j]`` with shape ``(chi_left, d, chi_right)``; storage order in the C ABI
  is column-major (l fastest), i.e. ``A.transpose(2,1,0)`` C-contiguous.
* Sites are 0-based here; ``bond b`` (0-based) joins sites ``b`` and ``b+1``.  The reference is
  1-based; every docstring cites reference lines with the 1-based meaning.
* Two-site gate ``G`` is a ``(d*d, d*d)`` matrix, row = (s1', s2') out, col = (s1, s2) in,
  with s1 the slow index:  row = s1'*d + s2'.
* MPO tensors ``W[j]`` have shape ``(wl, d_out, d_in, wr)``.

Third-party behaviour restated from documentation (not in /root/reference):
ITensors.jl v0.1.x (May 2020 master; unpinned, Project.toml:6-11) ``orthogonalize!``,
``replacebond!``/``factorize``/``svd``/``truncate!``/``Spectrum``/``entropy``, ``ProjMPO``,
``sweepnext``; KrylovKit.jl 0.4.0 ``exponentiate`` (benchmarks/Manifest.toml:58-62).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import scipy.linalg as sla

# --------------------------------------------------------------------------------------
# S=1/2 site operators  [UPSTREAM ITensors "S=1/2" site type; state 1 = "Up"]
# --------------------------------------------------------------------------------------
_SPIN_HALF_OPS = {
    "Id": np.eye(2, dtype=complex),
    "Sz": np.array([[0.5, 0], [0, -0.5]], dtype=complex),
    "Sx": np.array([[0, 0.5], [0.5, 0]], dtype=complex),
    "Sy": np.array([[0, -0.5j], [0.5j, 0]], dtype=complex),
    "S+": np.array([[0, 1], [0, 0]], dtype=complex),
    "S-": np.array([[0, 0], [1, 0]], dtype=complex),
    # spinless fermion on a site written in the same 2-dim space: state 0 = "Occ" here is NOT
    # used; for the free-fermion test we map Emp=|0>, Occ=|1>:
    "N": np.array([[0, 0], [0, 1]], dtype=complex),
    "Cdag": np.array([[0, 0], [1, 0]], dtype=complex),  # |1><0|  (nearest-neighbour: no JW string)
    "C": np.array([[0, 1], [0, 0]], dtype=complex),
}


def op_matrix(name: str) -> np.ndarray:
    """Single-site operator, incl. "A*B" product strings (src/bondop.jl:17-25)."""
    parts = name.split("*")
    o = _SPIN_HALF_OPS[parts[0]].copy()
    for p in parts[1:]:
        o = o @ _SPIN_HALF_OPS[p]
    return o


# --------------------------------------------------------------------------------------
# Spectrum / truncate!   [UPSTREAM ITensors truncate!, Spectrum, entropy]
# --------------------------------------------------------------------------------------
@dataclass
class Spectrum:
    eigs: np.ndarray  # kept sigma^2 (un-normalised), descending
    truncerr: float

    def entropy(self) -> float:
        p = self.eigs[self.eigs > 1e-13]
        return float(-np.sum(p * np.log(p)))


def truncate_probs(P: np.ndarray, *, maxdim=None, mindim=1, cutoff=0.0,
                   use_absolute_cutoff=False, do_rel_cutoff=True) -> Tuple[int, float]:
    """Upstream ``truncate!`` on a descending vector of sigma^2.  Returns (nkeep, truncerr).

    Restated from ITensors.jl v0.1 ``truncate!``.
    """
    P = np.asarray(P, dtype=float).copy()
    origm = len(P)
    maxdim = origm if maxdim is None else min(int(maxdim), origm)
    mindim = max(int(mindim), 1)
    cutoff = max(float(cutoff), 0.0)
    if P[0] <= 0.0:
        return 1, 0.0
    if origm == 1:
        return 1, 0.0
    P[P < 0.0] = 0.0
    n = origm
    truncerr = 0.0
    while n > maxdim:
        truncerr += P[n - 1]
        n -= 1
    if use_absolute_cutoff:
        while P[n - 1] <= cutoff and n > mindim:
            truncerr += P[n - 1]
            n -= 1
    else:
        scale = 1.0
        if do_rel_cutoff:
            scale = float(np.sum(P))
            if scale == 0.0:
                scale = 1.0
        while (truncerr + P[n - 1] <= cutoff * scale) and n > mindim:
            truncerr += P[n - 1]
            n -= 1
        truncerr /= scale
    n = max(n, 1)
    return n, truncerr


# --------------------------------------------------------------------------------------
# MPS
# --------------------------------------------------------------------------------------
class MPS:
    """Open-boundary MPS with ITensors-style orthogonality limits (llim, rlim).

    Sites 0..N-1; tensors j<=llim are left-orthonormal, j>=rlim right-orthonormal
    (0-based version of ITensors leftlim/rightlim).
    """

    def __init__(self, tensors: Sequence[np.ndarray], llim: int = -1, rlim: Optional[int] = None):
        self.A: List[np.ndarray] = [np.asarray(t, dtype=complex).copy() for t in tensors]
        self.llim = llim
        self.rlim = len(self.A) if rlim is None else rlim

    def __len__(self):
        return len(self.A)

    def copy(self) -> "MPS":
        return MPS(self.A, self.llim, self.rlim)

    @property
    def N(self):
        return len(self.A)

    def bond_dims(self) -> List[int]:
        return [self.A[j].shape[2] for j in range(self.N - 1)]

    def maxlinkdim(self) -> int:
        return max(self.bond_dims()) if self.N > 1 else 1

    # ---- constructors ----
    @staticmethod
    def product(states: Sequence[int], d: int = 2) -> "MPS":
        ts = []
        for s in states:
            t = np.zeros((1, d, 1), dtype=complex)
            t[0, s, 0] = 1.0
            ts.append(t)
        return MPS(ts, llim=-1, rlim=1) if len(ts) else MPS(ts)

    @staticmethod
    def random(N: int, chi: int, d: int = 2, seed: int = 0, canonical: bool = True) -> "MPS":
        """Random saturated MPS: iid complex Gaussian site tensors, chi_j=min(d^j,d^(N-j),chi)
        (SURVEY.md section 8(d) synthetic inputs), canonicalised to site 0 and normalised."""
        rng = np.random.default_rng(seed)
        dims = [min(d ** min(j, N - j, 40), chi) for j in range(N + 1)]
        ts = []
        for j in range(N):
            shape = (dims[j], d, dims[j + 1])
            ts.append(rng.standard_normal(shape) + 1j * rng.standard_normal(shape))
        psi = MPS(ts)
        if canonical:
            psi.llim, psi.rlim = -1, N
            # bring to right-canonical form, centre at 0
            for j in range(N - 1, 0, -1):
                psi._hop_left(j)
                psi.A[j - 1] /= np.linalg.norm(psi.A[j - 1])  # keeps long chains from overflowing
            psi.llim, psi.rlim = -1, 1
        return psi

    # ---- orthogonalize! [UPSTREAM]: QR hops moving the centre ----
    def _hop_right(self, j: int):
        A = self.A[j]
        l, d, r = A.shape
        Q, R = np.linalg.qr(A.reshape(l * d, r))
        k = Q.shape[1]
        self.A[j] = Q.reshape(l, d, k)
        self.A[j + 1] = np.tensordot(R, self.A[j + 1], axes=(1, 0))

    def _hop_left(self, j: int):
        A = self.A[j]
        l, d, r = A.shape
        Q, R = np.linalg.qr(A.reshape(l, d * r).T)  # A^T = Q R  ->  A = R^T Q^T
        k = Q.shape[1]
        self.A[j] = Q.T.reshape(k, d, r)
        self.A[j - 1] = np.tensordot(self.A[j - 1], R.T, axes=(2, 0))

    def orthogonalize(self, j: int):
        """``orthogonalize!(psi, j)`` (called at src/bondgate.jl:48, src/tdvp.jl:51)."""
        while self.llim < j - 1:
            self._hop_right(self.llim + 1)
            self.llim += 1
            if self.rlim < self.llim + 2:
                self.rlim = self.llim + 2
        while self.rlim > j + 1:
            self._hop_left(self.rlim - 1)
            self.rlim -= 1
            if self.llim > self.rlim - 2:
                self.llim = self.rlim - 2

    def reorthogonalize(self):
        """src/itensor.jl:28-33."""
        self.llim, self.rlim = -1, self.N
        self.orthogonalize(0)
        self.A[0] /= math.sqrt(inner(self, self).real)

    # ---- replacebond! [UPSTREAM] ----
    def replacebond(self, b: int, theta: np.ndarray, *, ortho: str = "left", normalize: bool = False,
                    maxdim=None, mindim=1, cutoff=None, use_absolute_cutoff=False,
                    do_rel_cutoff=True) -> Spectrum:
        """theta has shape (l, d, d, r).  SVD split, truncate only if maxdim or cutoff given
        (upstream ``svd``: ``truncate = haskey(maxdim) || haskey(cutoff)``)."""
        l, d1, d2, r = theta.shape
        M = theta.reshape(l * d1, d2 * r)
        try:
            U, S, Vh = np.linalg.svd(M, full_matrices=False)
        except np.linalg.LinAlgError:
            U, S, Vh = sla.svd(M, full_matrices=False, lapack_driver="gesvd")
        P = S ** 2
        if maxdim is not None or cutoff is not None:
            nkeep, truncerr = truncate_probs(P, maxdim=maxdim, mindim=mindim,
                                             cutoff=0.0 if cutoff is None else cutoff,
                                             use_absolute_cutoff=use_absolute_cutoff,
                                             do_rel_cutoff=do_rel_cutoff)
        else:
            nkeep, truncerr = len(P), 0.0
        U, S, Vh, P = U[:, :nkeep], S[:nkeep], Vh[:nkeep], P[:nkeep]
        if ortho == "left":
            self.A[b] = U.reshape(l, d1, nkeep)
            C = (S[:, None] * Vh).reshape(nkeep, d2, r)
            if normalize:
                C = C / np.linalg.norm(C)
            self.A[b + 1] = C
            if self.llim == b - 1:
                self.llim += 1
            if self.rlim == b + 1:
                self.rlim += 1
        elif ortho == "right":
            C = (U * S[None, :]).reshape(l, d1, nkeep)
            if normalize:
                C = C / np.linalg.norm(C)
            self.A[b] = C
            self.A[b + 1] = Vh.reshape(nkeep, d2, r)
            if self.llim == b:
                self.llim -= 1
            if self.rlim == b + 2:
                self.rlim -= 1
        else:
            raise ValueError(ortho)
        return Spectrum(P.copy(), float(truncerr))

    def theta(self, b: int) -> np.ndarray:
        """psi[b]*psi[b+1] -> (l, d, d, r)."""
        return np.tensordot(self.A[b], self.A[b + 1], axes=(2, 0))

    def to_dense(self) -> np.ndarray:
        v = self.A[0]
        for j in range(1, self.N):
            v = np.tensordot(v, self.A[j], axes=(v.ndim - 1, 0))
        return v.reshape(-1)


def inner(a: MPS, b: MPS) -> complex:
    """<a|b>."""
    E = np.ones((1, 1), dtype=complex)
    for j in range(a.N):
        T = np.tensordot(E, b.A[j], axes=(1, 0))  # (a', s, r)
        E = np.tensordot(a.A[j].conj(), T, axes=([0, 1], [0, 1]))
    return complex(E[0, 0])


def inner_mpo(a: MPS, W: Sequence[np.ndarray], b: MPS) -> complex:
    """<a|H|b> for an MPO list W[j] (wl, s_out, s_in, wr)."""
    E = np.ones((1, 1, 1), dtype=complex)  # (bra, w, ket)
    for j in range(a.N):
        T = np.tensordot(E, b.A[j], axes=(2, 0))  # (bra, w, s, r)
        T = np.tensordot(T, W[j], axes=([1, 2], [0, 2]))  # (bra, r, s_out, wr)
        E = np.tensordot(a.A[j].conj(), T, axes=([0, 1], [0, 2]))  # (r_bra, r, wr)
        E = E.transpose(0, 2, 1)
    return complex(E[0, 0, 0])


def expect_site(psi: MPS, opname_or_mat, j: int) -> complex:
    """``measure!(psi, op, i)`` (src/testutils.jl:55-58): moves the centre to j."""
    O = op_matrix(opname_or_mat) if isinstance(opname_or_mat, str) else opname_or_mat
    psi.orthogonalize(j)
    A = psi.A[j]
    return complex(np.einsum("atb,ts,asb->", A.conj(), O, A, optimize=True))


def expect_all(psi: MPS, op) -> np.ndarray:
    p = psi.copy()
    return np.array([expect_site(p, op, j).real for j in range(p.N)])


def bond_entropies(psi: MPS) -> List[float]:
    """von Neumann entropy at every bond from a fresh SVD (as test/test_callbacks.jl:92-97)."""
    p = psi.copy()
    out = []
    for b in range(p.N - 1):
        p.orthogonalize(b)
        th = p.theta(b)
        l, d1, d2, r = th.shape
        S = np.linalg.svd(th.reshape(l * d1, d2 * r), compute_uv=False)
        out.append(Spectrum(S ** 2, 0.0).entropy())
    return out


# --------------------------------------------------------------------------------------
# BondOperator / gates  (src/bondop.jl)
# --------------------------------------------------------------------------------------
@dataclass
class BondGate:
    """(G, b): src/bondgate.jl:9-22.  G is (d*d, d*d), b is the 0-based bond."""
    G: np.ndarray
    b: int


class BondOperator:
    """src/bondop.jl:43-88.  Sites are 0-based here; ``add`` mirrors the four ``add!`` methods."""

    def __init__(self, N: int, d: int = 2):
        self.N, self.d = N, d
        self.bondterms: Dict[int, list] = {b: [] for b in range(N - 1)}
        self.siteterms: Dict[int, list] = {i: [] for i in range(N)}

    def add(self, *args):
        args = list(args)
        c = 1.0
        if not isinstance(args[0], str):
            c = args.pop(0)
        if len(args) == 2:  # (op, i)
            self.siteterms[args[1]].append((c, args[0]))
        elif len(args) == 3:  # (op1, op2, b)
            self.bondterms[args[2]].append((c, args[0], args[1]))
        else:
            raise ValueError(args)
        return self

    def bondgate(self, b: int) -> BondGate:
        """src/bondop.jl:72-86: site terms enter with 1/2 on interior sites, 1 at chain ends."""
        d = self.d
        I = np.eye(d, dtype=complex)
        g = np.zeros((d * d, d * d), dtype=complex)
        for (c, o1, o2) in self.bondterms[b]:
            g += c * np.kron(op_matrix(o1), op_matrix(o2))
        for i in (b, b + 1):
            fac = 1.0 if (i == self.N - 1 or i == 0) else 0.5
            for (c, o) in self.siteterms[i]:
                O = c * op_matrix(o)
                g += fac * (np.kron(O, I) if i == b else np.kron(I, O))
        return BondGate(g, b)

    def gates(self) -> List[BondGate]:
        return [self.bondgate(b) for b in range(self.N - 1)]


def tfi_bondop(N: int, J: float, h: float) -> BondOperator:
    """src/testutils.jl:43-52."""
    H = BondOperator(N)
    for b in range(N - 1):
        H.add(-J, "Sz", "Sz", b)
        H.add(-h, "Sx", b)
    H.add(-h, "Sx", N - 1)
    return H


def xxz_bondop(N: int, delta: float = 1.0, hz: float = 0.0) -> BondOperator:
    """XXZ chain H = sum SxSx+SySy+delta SzSz - hz sum Sz (benchmarks/tebd/tebd_benchmark.jl:17-26
    for delta=1, hz=0; the field is this build's sweep parameter)."""
    H = BondOperator(N)
    for b in range(N - 1):
        H.add(0.5, "S+", "S-", b)
        H.add(0.5, "S-", "S+", b)
        H.add(delta, "Sz", "Sz", b)
    if hz != 0.0:
        for i in range(N):
            H.add(-hz, "Sz", i)
    return H


def gate_exp(g: BondGate, c: complex) -> BondGate:
    """exp(c * G) (src/bondgate.jl:18,24-25) - dense matrix exponential of the d^2 x d^2 gate."""
    return BondGate(sla.expm(c * g.G), g.b)


# --------------------------------------------------------------------------------------
# TEBD  (src/bondgate.jl:46-79, src/tebd.jl)
# --------------------------------------------------------------------------------------
def apply_gate(psi: MPS, g: BondGate, **kw) -> Spectrum:
    """src/bondgate.jl:46-53."""
    b = g.b
    psi.orthogonalize(b)
    th = psi.theta(b)
    l, d1, d2, r = th.shape
    th = np.tensordot(g.G.reshape(d1, d2, d1, d2), th, axes=([2, 3], [1, 2]))  # (s1',s2',l,r)
    th = th.transpose(2, 0, 1, 3)
    tk = {k: kw[k] for k in ("maxdim", "mindim", "cutoff", "use_absolute_cutoff", "ortho") if k in kw}
    return psi.replacebond(b, th, normalize=kw.get("normalize", True), **tk)


def apply_gates(psi: MPS, Gs: Sequence[BondGate], *, reversed: bool = False, cb=None, **kw):
    """src/bondgate.jl:55-63."""
    order = list(Gs)[::-1] if reversed else list(Gs)
    for g in order:
        spec = apply_gate(psi, g, **kw)
        if cb is not None:
            cb(psi, bond=g.b, spec=spec)


def measure_gates(H: Sequence[BondGate], psi: MPS) -> complex:
    """``measure(H::GateList, psi)`` (src/bondgate.jl:71-79)."""
    e = 0.0
    for g in H:
        psi.orthogonalize(g.b)
        th = psi.theta(g.b)
        l, d1, d2, r = th.shape
        hth = np.tensordot(g.G.reshape(d1, d2, d1, d2), th, axes=([2, 3], [1, 2])).transpose(2, 0, 1, 3)
        e += np.vdot(th, hth)
    return complex(e)


class TEBD2:
    name = "TEBD2"


class TEBD2Sweep:
    name = "TEBD2Sweep"


class TEBD4:
    name = "TEBD4"


class TDVP2:
    name = "TDVP2"


def _is_tebd(alg) -> bool:
    return isinstance(alg, (TEBD2, TEBD2Sweep, TEBD4))


def time_evo_gates(dt, H: Sequence[BondGate], alg):
    """src/tebd.jl:56-87.  Julia H[1:2:end] = 1-based odd bonds = 0-based even indices."""
    o = list(H[0::2])
    e = list(H[1::2])
    ex = lambda c, gs: [gate_exp(g, c) for g in gs]
    if isinstance(alg, TEBD2):
        Uhalf = ex(-1j * dt / 2, o)
        Us = [ex(-1j * dt, e), ex(-1j * dt, o)]
        return [Uhalf], Us, [Us[0], Uhalf]
    if isinstance(alg, TEBD2Sweep):
        Us = [ex(-1j * dt / 2, list(H)), ex(-1j * dt / 2, list(H))]
        return [], Us, []
    if isinstance(alg, TEBD4):
        t1 = 1 / (4 - 4 ** (1 / 3)) * dt
        t2 = t1
        t3 = dt - 2 * t1 - 2 * t2
        Ustart = [ex(-1j * t1 / 2, o)]
        seq = [(t1, e), (t1, o), (t2, e), ((t2 + t3) / 2, o), (t3, e),
               ((t2 + t3) / 2, o), (t2, e), (t2, o), (t1, e), (t1, o)]
        Us = [ex(-1j * x[0], x[1]) for x in seq]
        eseq = seq[:-1] + [(t1 / 2, o)]
        Uend = [ex(-1j * x[0], x[1]) for x in eseq]
        return Ustart, Us, Uend
    raise ValueError(alg)


def _nsteps(tf, dt) -> int:
    """``Int(tf/dt)`` (src/tebd.jl:95, src/tdvp.jl:38): InexactError if not integral."""
    q = tf / dt
    if isinstance(q, complex):
        if q.imag != 0:
            raise ValueError("InexactError: Int(%r)" % (q,))
        q = q.real
    if q != int(q):
        raise ValueError("InexactError: Int(%r)" % (q,))
    return int(q)


def tebd(psi: MPS, H, dt, tf, alg=None, *, callback=None, orthogonalize: int = 0, **kw) -> MPS:
    """``tebd!`` (src/tebd.jl:89-179)."""
    alg = TEBD2() if alg is None else alg
    if isinstance(H, BondOperator):
        H = H.gates()
    nsteps = _nsteps(tf, dt)
    cb = NoTEvoCallback() if callback is None else callback

    def cb_func(dt_, step, rev, se):
        return lambda p, bond, spec: cb.apply(p, t=dt_ * step, bond=bond, spec=spec,
                                              sweepdir="left" if rev else "right",
                                              sweepend=se, alg=alg)

    dtm = cb.callback_dt()
    if dtm > 0:
        if math.floor(_real(dtm / dt)) != dtm / dt:
            raise ValueError("Measurement time step %s incommensurate with time-evolution time step %s" % (dtm, dt))
        mstep = math.floor(_real(dtm / dt))
        nbunch = math.gcd(mstep, nsteps)
    else:
        nbunch = nsteps
    if orthogonalize > 0:
        nbunch = math.gcd(nbunch, orthogonalize)
    Ustart, Us, Uend = time_evo_gates(dt, H, alg)
    if len(Ustart) == 0 and len(Uend) == 0:
        nbunch += 1
    step = 0
    rev = False
    while step < nsteps:
        for U in Ustart:
            apply_gates(psi, U, reversed=rev, **kw)
            rev = not rev
        brk = False
        for _ in range(nbunch - 1):
            for U in Us:
                apply_gates(psi, U, reversed=rev, cb=cb_func(dt, step, rev, False), **kw)
                rev = not rev
            step += 1
            if cb.checkdone(psi):
                brk = True
                break
        for i, U in enumerate(Uend):  # Julia i is 1-based: i >= length(Uend)-1
            apply_gates(psi, U, reversed=rev, cb=cb_func(dt, step + 1, rev, (i + 1) >= len(Uend) - 1), **kw)
            rev = not rev
        if len(Uend) > 0:
            step += 1
        if orthogonalize > 0 and step % orthogonalize == 0:
            psi.reorthogonalize()
        if cb.checkdone(psi):
            break
    return psi


def _real(x):
    return x.real if isinstance(x, complex) else x


# --------------------------------------------------------------------------------------
# Callbacks  (src/callback.jl)
# --------------------------------------------------------------------------------------
def _approx(a, b):
    """Julia isapprox default: |a-b| <= sqrt(eps)*max(|a|,|b|)."""
    return abs(a - b) <= math.sqrt(np.finfo(float).eps) * max(abs(a), abs(b))


class NoTEvoCallback:
    def apply(self, *a, **k):
        return None

    def checkdone(self, *a, **k):
        return False

    def callback_dt(self):
        return 0


class LocalMeasurementCallback:
    """src/callback.jl:33-100."""

    def __init__(self, ops: Sequence[str], N: int, dt_measure: float):
        self.ops = list(ops)
        self.N = N
        self.measurements: Dict[str, List[np.ndarray]] = {o: [] for o in ops}
        self.ts: List[float] = []
        self.dt_measure = dt_measure

    def callback_dt(self):
        return self.dt_measure

    def checkdone(self, *a, **k):
        return False

    def _measure_localops(self, wf: np.ndarray, which: int, i: int):
        for o in self.ops:
            O = op_matrix(o)
            if which == 0:
                m = np.einsum("atub,ts,asub->", wf.conj(), O, wf, optimize=True)
            else:
                m = np.einsum("autb,ts,ausb->", wf.conj(), O, wf, optimize=True)
            self.measurements[o][-1][i] = m.real

    def apply(self, psi: MPS, *, t, sweepend, sweepdir, bond, alg, **kw):
        t = _real(t)
        prev_t = self.ts[-1] if self.ts else 0
        if (_approx(t - prev_t, self.dt_measure) or t == prev_t) and sweepend and \
                ((bond + 1) % 2 == 1 or not _is_tebd(alg)):
            if t != prev_t:
                self.ts.append(t)
                for o in self.ops:
                    self.measurements[o].append(np.zeros(self.N))
            wf = psi.theta(bond)
            self._measure_localops(wf, 1, bond + 1)
            if _is_tebd(alg):
                self._measure_localops(wf, 0, bond)
            elif bond == 0:
                self._measure_localops(wf, 0, bond)


class SpecCallback:
    """src/callback.jl:102-162."""

    def __init__(self, dt: float, N: int, bonds: Optional[Sequence[int]] = None):
        bonds = sorted(set(range(N - 1) if bonds is None else bonds))
        if max(bonds) > N - 2 or min(bonds) < 0:
            raise ValueError("bonds must be between 1 and %d" % (N - 1))
        self.N = N
        self.truncerrs: List[float] = []
        self.current_truncerr = 0.0
        self.entropies: List[np.ndarray] = []
        self.bonddims: List[np.ndarray] = []
        self.bonds = bonds
        self.ts: List[float] = []
        self.dt_measure = dt

    def callback_dt(self):
        return self.dt_measure

    def checkdone(self, *a, **k):
        return False

    def apply(self, psi: MPS, *, t, sweepend, bond, spec, sweepdir, **kw):
        t = _real(t)
        self.current_truncerr += spec.truncerr
        prev_t = self.ts[-1] if self.ts else 0
        if (_approx(t - prev_t, self.dt_measure) or t == prev_t) and sweepend:
            if t != prev_t:
                self.ts.append(t)
                self.bonddims.append(np.zeros(len(self.bonds), dtype=np.int64))
                self.entropies.append(np.zeros(len(self.bonds)))
            if bond in self.bonds:
                i = self.bonds.index(bond)
                self.bonddims[-1][i] = len(spec.eigs)
                self.entropies[-1][i] = spec.entropy()
            if sweepdir == "right" and bond == self.N - 2:
                self.truncerrs.append(self.current_truncerr)
            elif sweepdir == "left" and bond == 0:
                self.truncerrs.append(self.current_truncerr)


# --------------------------------------------------------------------------------------
# MPOs (AutoMPO results written analytically; src/testutils.jl:4-12)
# --------------------------------------------------------------------------------------
def _mpo_from_bulk(Wb: np.ndarray, N: int) -> List[np.ndarray]:
    """Lower-triangular bulk W (w, d, d, w): left boundary = last row, right = first column."""
    out = []
    for j in range(N):
        W = Wb
        if j == 0:
            W = W[-1:, :, :, :]
        if j == N - 1:
            W = W[:, :, :, :1]
        out.append(W.copy())
    return out


def tfi_mpo(N: int, J: float, h: float) -> List[np.ndarray]:
    """H = -J sum SzSz - h sum Sx, MPO bond dimension 3."""
    I, Sz, Sx = op_matrix("Id"), op_matrix("Sz"), op_matrix("Sx")
    W = np.zeros((3, 2, 2, 3), dtype=complex)
    W[0, :, :, 0] = I
    W[1, :, :, 0] = Sz
    W[2, :, :, 0] = -h * Sx
    W[2, :, :, 1] = -J * Sz
    W[2, :, :, 2] = I
    return _mpo_from_bulk(W, N)


def xxz_mpo(N: int, delta: float = 1.0, hz: float = 0.0) -> List[np.ndarray]:
    """H = sum [1/2 (S+S- + S-S+) + delta SzSz] - hz sum Sz, MPO bond dimension 5."""
    I, Sz, Sp, Sm = op_matrix("Id"), op_matrix("Sz"), op_matrix("S+"), op_matrix("S-")
    W = np.zeros((5, 2, 2, 5), dtype=complex)
    W[0, :, :, 0] = I
    W[1, :, :, 0] = Sm
    W[2, :, :, 0] = Sp
    W[3, :, :, 0] = Sz
    W[4, :, :, 0] = -hz * Sz
    W[4, :, :, 1] = 0.5 * Sp
    W[4, :, :, 2] = 0.5 * Sm
    W[4, :, :, 3] = delta * Sz
    W[4, :, :, 4] = I
    return _mpo_from_bulk(W, N)


def hopping_mpo(N: int) -> List[np.ndarray]:
    """H = sum c+_b c_{b+1} + h.c. (test/test_tdvp.jl:101-105; '-C Cdag' = 'Cdag C' reordered,
    nearest-neighbour so no Jordan-Wigner string survives)."""
    I, Cd, C = op_matrix("Id"), op_matrix("Cdag"), op_matrix("C")
    W = np.zeros((4, 2, 2, 4), dtype=complex)
    W[0, :, :, 0] = I
    W[1, :, :, 0] = C
    W[2, :, :, 0] = Cd
    W[3, :, :, 1] = Cd
    W[3, :, :, 2] = C
    W[3, :, :, 3] = I
    return _mpo_from_bulk(W, N)


def mpo_to_dense(W: Sequence[np.ndarray]) -> np.ndarray:
    N = len(W)
    cur = W[0][0]  # (so, si, wr)
    for j in range(1, N):
        cur = np.tensordot(cur, W[j], axes=(cur.ndim - 1, 0))  # (..., so, si, wr)
    cur = cur[..., 0]
    outs = list(range(0, 2 * N, 2))
    ins = list(range(1, 2 * N, 2))
    cur = cur.transpose(outs + ins)
    D = int(np.prod(cur.shape[:N]))
    return cur.reshape(D, D)


# --------------------------------------------------------------------------------------
# ProjMPO  [UPSTREAM ITensors ProjMPO: position!, product]
# --------------------------------------------------------------------------------------
class ProjMPO:
    """Environment cache.  LR[j] for j<lpos is the left env including sites 0..j;
    for j>rpos the right env including sites j..N-1.  Shapes (ket, w, bra)."""

    def __init__(self, W: Sequence[np.ndarray]):
        self.W = list(W)
        self.N = len(W)
        self.nsite = 2
        self.lpos = -1
        self.rpos = self.N
        self.LR: List[Optional[np.ndarray]] = [None] * self.N

    def _L(self, j):
        return self.LR[j] if j >= 0 else np.ones((1, 1, 1), dtype=complex)

    def _R(self, j):
        return self.LR[j] if j < self.N else np.ones((1, 1, 1), dtype=complex)

    def position(self, psi: MPS, pos: int):
        """``position!(PH, psi, pos)`` (src/tdvp.jl:53,59,80)."""
        while self.lpos < pos - 1:
            j = self.lpos + 1
            L = self._L(j - 1)
            A = psi.A[j]
            T = np.tensordot(L, A, axes=(0, 0))  # (w, bra, s, r)
            T = np.tensordot(T, self.W[j], axes=([0, 2], [0, 2]))  # (bra, r, s_out, wr)
            E = np.tensordot(T, A.conj(), axes=([0, 2], [0, 1]))  # (r, wr, r_bra)
            self.LR[j] = E
            self.lpos = j
        self.lpos = pos - 1
        end = pos + self.nsite - 1
        while self.rpos > end + 1:
            j = self.rpos - 1
            R = self._R(j + 1)
            A = psi.A[j]
            T = np.tensordot(A, R, axes=(2, 0))  # (l, s, w, bra)
            T = np.tensordot(T, self.W[j], axes=([1, 2], [2, 3]))  # (l, bra, wl, s_out)
            E = np.tensordot(T, A.conj(), axes=([3, 1], [1, 2]))  # (l, wl, l_bra)
            self.LR[j] = E
            self.rpos = j
        self.rpos = end + 1

    def product(self, pos: int, v: np.ndarray) -> np.ndarray:
        """H_eff v for nsite=2 (v: (l,s1,s2,r)) or nsite=1 (v: (l,s,r))."""
        L = self._L(pos - 1)
        R = self._R(pos + self.nsite)
        if self.nsite == 2:
            T = np.tensordot(L, v, axes=(0, 0))  # (w, l', s1, s2, r)
            T = np.tensordot(T, self.W[pos], axes=([0, 2], [0, 2]))  # (l', s2, r, s1', x)
            T = np.tensordot(T, self.W[pos + 1], axes=([4, 1], [0, 2]))  # (l', r, s1', s2', y)
            T = np.tensordot(T, R, axes=([1, 4], [0, 1]))  # (l', s1', s2', r')
            return T
        T = np.tensordot(L, v, axes=(0, 0))  # (w, l', s, r)
        T = np.tensordot(T, self.W[pos], axes=([0, 2], [0, 2]))  # (l', r, s', x)
        T = np.tensordot(T, R, axes=([1, 3], [0, 1]))  # (l', s', r')
        return T


# --------------------------------------------------------------------------------------
# Krylov exponentiate  [UPSTREAM KrylovKit.exponentiate 0.4.0 semantics]
# --------------------------------------------------------------------------------------
@dataclass
class KrylovInfo:
    converged: int
    numops: int
    numiter: int
    normres: float


def _small_exp_col(Hk: np.ndarray, c, ishermitian: bool):
    """First column of exp(c*Hk) and the last-row/first-column elements needed by the estimate."""
    if ishermitian:
        D, U = np.linalg.eigh(Hk)
        f = lambda x: (U * np.exp(x * D)[None, :]) @ U[0].conj()
    else:
        f = lambda x: sla.expm(x * Hk)[:, 0]
    return f


def _signif(x: float, digits: int) -> float:
    """Julia's round(x, sigdigits=digits) as used by KrylovKit's step control."""
    if x == 0 or not math.isfinite(x):
        return x
    return float("%.*e" % (digits - 1, x))


def exponentiate(Aop: Callable[[np.ndarray], np.ndarray], t, v0: np.ndarray, *, ishermitian=True,
                 tol=1e-12, krylovdim=30, maxiter=100, mode: str = "early") -> Tuple[np.ndarray, KrylovInfo]:
    """exp(t*A) v0 by restarted Lanczos (Hermitian) / Arnoldi with full re-orthogonalisation.

    KrylovKit 0.4 semantics kept: ``tol`` is an error per unit time; the a-posteriori estimate is
    eps = normres * (2|e1|/3 + |e2|/6) with e_x = [exp(sgn*x*dtau*H_K)]_{K,1} (x = 1/2, 1); if it is
    not met at ``krylovdim`` the sub-step is shrunk by (tol/eps)^(1/krylovdim) and the iteration
    restarts from the advanced vector (<= maxiter restarts); ``converged`` = 1 iff the whole interval
    was integrated.  mode="early" (default; what the CUDA path does): the estimate is also tested before
    the space is full (same answer to within tol, fewer operator applications).  mode="krylovkit": the
    literal KrylovKit 0.4.0 control flow (src/matrixfun/exponentiate.jl of the pinned dependency,
    benchmarks/Manifest.toml:58-62): the space is expanded while normres > tol and dim < krylovdim, the
    estimate is evaluated only then, a failing step is shrunk to signif(0.9 (tol/eps)^(1/krylovdim) dtau, 2).
    tests/test_oracle_reference_tests.py shows both modes give the same vector to within tol.
    """
    if mode not in ("early", "krylovkit"):
        raise ValueError(mode)
    literal = mode == "krylovkit"
    shape = v0.shape
    v = v0.reshape(-1).astype(complex)
    n = v.size
    if np.linalg.norm(v) == 0:
        return v0.copy(), KrylovInfo(1, 0, 0, 0.0)
    tau_total = abs(t)
    if tau_total == 0:
        return v0.copy(), KrylovInfo(1, 0, 0, 0.0)
    sgn = t / tau_total
    tau_done = 0.0
    numops = 0
    numiter = 0
    w = v.copy()
    eta = float(tol)
    conv = 0
    while True:
        numiter += 1
        beta = np.linalg.norm(w)
        kd = min(krylovdim, n)
        V = np.zeros((kd + 1, n), dtype=complex)
        Hm = np.zeros((kd + 1, kd), dtype=complex)
        V[0] = w / beta
        dtau = tau_total - tau_done
        k = 0
        done_sub = False
        while k < kd:
            r = Aop(V[k].reshape(shape)).reshape(-1)
            numops += 1
            h = V[:k + 1].conj() @ r          # full two-pass orthogonalisation
            r = r - V[:k + 1].T @ h
            h2 = V[:k + 1].conj() @ r
            r = r - V[:k + 1].T @ h2
            h = h + h2
            Hm[:k + 1, k] = h
            nb = np.linalg.norm(r)
            Hm[k + 1, k] = nb
            k += 1
            Hk = Hm[:k, :k]
            if ishermitian:
                Hk = 0.5 * (Hk + Hk.conj().T)
            f = _small_exp_col(Hk, 1.0, ishermitian)
            e1 = f(sgn * dtau / 2)[-1]
            e2 = f(sgn * dtau)[-1]
            eps = nb * (2 * abs(e1) / 3 + abs(e2) / 6)
            if literal:
                # KrylovKit: `while normres(fact) > eta && length(fact) < krylovdim; expand!` - no estimate yet
                if not (nb > eta and k < kd):
                    break
            elif eps < eta or k == n:
                done_sub = True
                break
            if k == n:
                break
            V[k] = r / nb
        if not done_sub:
            f = _small_exp_col(Hk, 1.0, ishermitian)
            while True:
                e1 = f(sgn * dtau / 2)[-1]
                e2 = f(sgn * dtau)[-1]
                eps = nb * (2 * abs(e1) / 3 + abs(e2) / 6)
                if eps < eta or numiter == maxiter:
                    break
                if literal:
                    dtau = _signif(0.9 * (eta / eps) ** (1.0 / krylovdim) * dtau, 2)
                else:
                    dtau = dtau * (eta / eps) ** (1.0 / kd)
        y = f(sgn * dtau)
        w = beta * (V[:k].T @ y)
        tau_done += dtau
        if tau_done >= tau_total * (1 - 1e-15):
            conv = 1
            break
        if numiter >= maxiter:
            break
    return w.reshape(shape), KrylovInfo(conv, numops, numiter, 0.0)


# --------------------------------------------------------------------------------------
# two-site TDVP  (src/tdvp.jl:37-96)
# --------------------------------------------------------------------------------------
def sweepnext(N: int):
    """[UPSTREAM] ``sweepnext(N)``: (b,ha) for b=1..N-1 (ha=1) then N-1..1 (ha=2); 0-based b here."""
    for b in range(0, N - 1):
        yield b, 1
    for b in range(N - 2, -1, -1):
        yield b, 2


def tdvp(psi: MPS, W: Sequence[np.ndarray], dt, tf, *, callback=None, hermitian=True, exp_tol=1e-14,
         krylovdim=30, maxiter=100, normalize=True, stats: Optional[dict] = None, krylov_mode: str = "early", **kw):
    """``tdvp!`` (src/tdvp.jl:37-96).  krylov_mode: see exponentiate ("krylovkit" = the literal KrylovKit 0.4 flow)."""
    nsteps = _nsteps(tf, dt)
    cb = NoTEvoCallback() if callback is None else callback
    tau = 1j * dt
    if isinstance(tau, complex) and tau.imag == 0:
        tau = tau.real
    imag_time = isinstance(dt, complex)
    N = psi.N
    psi.orthogonalize(0)
    PH = ProjMPO(W)
    PH.nsite = 2
    PH.position(psi, 0)
    tk = {k: kw[k] for k in ("maxdim", "mindim", "cutoff", "use_absolute_cutoff") if k in kw}
    for s in range(1, nsteps + 1):
        for (b, ha) in sweepnext(N):
            PH.nsite = 2
            PH.position(psi, b)
            wf = psi.theta(b)
            wf, info = exponentiate(lambda x: PH.product(b, x), -tau / 2, wf, ishermitian=hermitian,
                                    tol=exp_tol, krylovdim=krylovdim, mode=krylov_mode)
            if stats is not None:
                stats["matvec2"] = stats.get("matvec2", 0) + info.numops
            if info.converged == 0:
                raise RuntimeError("exponentiate did not converge")
            spec = psi.replacebond(b, wf, normalize=normalize, ortho="left" if ha == 1 else "right", **tk)
            cb.apply(psi, t=s * dt, bond=b, sweepend=(ha == 2),
                     sweepdir="right" if ha == 1 else "left", spec=spec, alg=TDVP2())
            i = b + 1 if ha == 1 else b
            if 0 < i < N - 1 and not imag_time:
                PH.nsite = 1
                PH.position(psi, i)
                new, info = exponentiate(lambda x: PH.product(i, x), tau / 2, psi.A[i],
                                         ishermitian=hermitian, tol=exp_tol, krylovdim=krylovdim,
                                         maxiter=maxiter, mode=krylov_mode)
                if stats is not None:
                    stats["matvec1"] = stats.get("matvec1", 0) + info.numops
                if info.converged == 0:
                    raise RuntimeError("exponentiate did not converge")
                psi.A[i] = new
            elif i == 0 and imag_time:
                psi.A[i] = psi.A[i] / np.linalg.norm(psi.A[i])
        if cb.checkdone():
            break
    return None


# --------------------------------------------------------------------------------------
# closed forms used by the reference tests
# --------------------------------------------------------------------------------------
def free_fermions_densities(N: int, dt: float, tf: float) -> List[np.ndarray]:
    """test/testutils.jl:3-40."""
    Nf = N // 2
    U = np.zeros((N, Nf), dtype=complex)
    for i in range(Nf):
        U[2 * i, i] = 1.0
    h = np.diag(np.ones(N - 1), 1) + np.diag(np.ones(N - 1), -1)
    expH = sla.expm(-1j * dt * h)
    ns = [np.real(np.diag(U @ U.conj().T))]
    for _ in range(int(math.floor(tf / dt))):
        Q, _ = np.linalg.qr(expH @ U)
        U = Q
        ns.append(np.real(np.diag(U @ U.conj().T)))
    return ns


def tfi_exact_critical_energy(N: int) -> float:
    """test/test_tebd.jl:68."""
    return 0.25 - 0.25 / math.sin(math.pi / (4 * N + 2))


# --------------------------------------------------------------------------------------
# Layer-parallel TEBD (Vidal / Hastings form)  -  NOT the reference's algorithm.
#
# The reference applies the gates of a layer sequentially with exact re-canonicalisation between gates
# (src/bondgate.jl:48,55-63).  Sharding a layer over GPUs needs gates that are independent of each other; the
# standard way is the canonical form B_j (right-orthonormal) + Schmidt values Lambda_j on every bond, updated with
# Hastings' inverse-free rule.  Without truncation it is the same state as the sequential sweep; with truncation the
# two differ at the order of the discarded weight.  This restatement is the checker of the
# CUDA layer-parallel / multi-GPU mode.
# --------------------------------------------------------------------------------------
class BForm:
    """B[j]: (chi_j, d, chi_{j+1}) right-orthonormal; lam[j]: Schmidt values on the bond left of site j
    (lam[0] = lam[N] = [1])."""

    def __init__(self, B, lam):
        self.B = [np.asarray(b, dtype=complex).copy() for b in B]
        self.lam = [np.asarray(l, dtype=float).copy() for l in lam]

    @property
    def N(self):
        return len(self.B)

    def copy(self):
        return BForm(self.B, self.lam)

    def to_mps(self) -> "MPS":
        """A B-form state is a right-canonical MPS with its centre on site 0."""
        return MPS(self.B, llim=-1, rlim=1)

    def bond_dims(self):
        return [self.B[j].shape[2] for j in range(self.N - 1)]

    def expect_site(self, op, j) -> complex:
        O = op_matrix(op) if isinstance(op, str) else op
        T = self.lam[j][:, None, None] * self.B[j]
        return complex(np.einsum("atb,ts,asb->", T.conj(), O, T, optimize=True))

    def expect_all(self, op) -> np.ndarray:
        return np.array([self.expect_site(op, j).real for j in range(self.N)])

    def entropies(self):
        out = []
        for j in range(1, self.N):
            p = self.lam[j] ** 2
            p = p[p > 1e-13]
            out.append(float(-np.sum(p * np.log(p))))
        return out


def to_bform(psi: MPS) -> BForm:
    """Left-canonicalise, then sweep leftwards with SVDs: B_j = V^dagger (right factor), lam_j = S."""
    p = psi.copy()
    N = p.N
    p.orthogonalize(N - 1)
    p.A[N - 1] = p.A[N - 1] / np.linalg.norm(p.A[N - 1])
    B = [None] * N
    lam = [None] * (N + 1)
    lam[N] = np.ones(1)
    C = p.A[N - 1]
    for j in range(N - 1, 0, -1):
        l, d, r = C.shape
        U, S, Vh = np.linalg.svd(C.reshape(l, d * r), full_matrices=False)
        k = len(S)
        B[j] = Vh.reshape(k, d, r)
        lam[j] = S.copy()
        C = np.tensordot(p.A[j - 1], U * S[None, :], axes=(2, 0))
    B[0] = C
    lam[0] = np.ones(1)
    return BForm(B, lam)


def apply_gate_bform(st: BForm, g: BondGate, *, normalize=True, maxdim=None, mindim=1, cutoff=None,
                     use_absolute_cutoff=False, do_rel_cutoff=True) -> Spectrum:
    """Hastings' inverse-free two-site update of bond b in B-form."""
    b = g.b
    C = np.tensordot(st.B[b], st.B[b + 1], axes=(2, 0))  # (l, s1, s2, r)
    l, d1, d2, r = C.shape
    C = np.tensordot(g.G.reshape(d1, d2, d1, d2), C, axes=([2, 3], [1, 2])).transpose(2, 0, 1, 3)
    theta = st.lam[b][:, None, None, None] * C
    U, S, Vh = np.linalg.svd(theta.reshape(l * d1, d2 * r), full_matrices=False)
    P = S ** 2
    if maxdim is not None or cutoff is not None:
        nkeep, truncerr = truncate_probs(P, maxdim=maxdim, mindim=mindim, cutoff=0.0 if cutoff is None else cutoff,
                                         use_absolute_cutoff=use_absolute_cutoff, do_rel_cutoff=do_rel_cutoff)
    else:
        nkeep, truncerr = len(P), 0.0
    S, Vh, P = S[:nkeep], Vh[:nkeep], P[:nkeep]
    nrm = np.sqrt(np.sum(P)) if normalize else 1.0
    st.B[b + 1] = Vh.reshape(nkeep, d2, r)
    st.B[b] = (C.reshape(l * d1, d2 * r) @ Vh.conj().T).reshape(l, d1, nkeep) / nrm
    st.lam[b + 1] = S / nrm
    return Spectrum(P.copy(), float(truncerr))


def apply_layer_bform(st: BForm, Gs: Sequence[BondGate], **kw) -> List[Spectrum]:
    """All gates of one even/odd layer act on disjoint bonds and only read the Schmidt values of bonds the layer does
    not touch, so the order is irrelevant (that is what makes the layer shardable)."""
    return [apply_gate_bform(st, g, **kw) for g in Gs]


def tebd_parallel(st: BForm, H, dt, tf, alg=None, **kw) -> BForm:
    """The schedule of tebd! (src/tebd.jl:113-177, no callbacks) driving layer-parallel updates."""
    alg = TEBD2() if alg is None else alg
    if isinstance(H, BondOperator):
        H = H.gates()
    nsteps = _nsteps(tf, dt)
    Ustart, Us, Uend = time_evo_gates(dt, H, alg)
    nbunch = nsteps
    if len(Ustart) == 0 and len(Uend) == 0:
        nbunch += 1
    step = 0
    while step < nsteps:
        for U in Ustart:
            apply_layer_bform(st, U, **kw)
        for _ in range(nbunch - 1):
            for U in Us:
                apply_layer_bform(st, U, **kw)
            step += 1
        for U in Uend:
            apply_layer_bform(st, U, **kw)
        if len(Uend) > 0:
            step += 1
    return st
