"""BondOperator / BondGate / gates  (src/bondop.jl, src/bondgate.jl:9-25,71-79) - host-side, setup-time only."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence

import numpy as np

from ._lib import check, load, tevo_complex
from .mps import MPS, op


class SiteTerm:
    """src/bondop.jl:9-15."""

    def __init__(self, i: int, coeff, opname: str):
        self.i, self.coeff, self.opname = i, coeff, opname

    def matrix(self) -> np.ndarray:
        return self.coeff * op(self.opname)


class BondTerm:
    """src/bondop.jl:27-35."""

    def __init__(self, b: int, c, op1: str, op2: str):
        self.b, self.coeff = b, c
        self.leftop, self.rightop = SiteTerm(b, 1, op1), SiteTerm(b + 1, 1, op2)


def coeff(t):
    return t.coeff


class BondGate:
    """(G, b) pair (src/bondgate.jl:9-22).  G: (d^2, d^2) matrix, row=(s1',s2'), col=(s1,s2), s1 slow."""

    def __init__(self, G: np.ndarray, b: int):
        self.G = np.asarray(G, dtype=complex)
        self.b = b

    def __mul__(self, c):
        return BondGate(c * self.G, self.b)

    __rmul__ = __mul__

    def __eq__(self, other):
        return isinstance(other, BondGate) and self.b == other.b and np.array_equal(self.G, other.G)


def bond(g: BondGate) -> int:
    return g.b


GateList = List[BondGate]


def expm(A: np.ndarray) -> np.ndarray:
    """Dense matrix exponential (scaling and squaring, Taylor) for the d^2 x d^2 gate generators."""
    A = np.asarray(A, dtype=complex)
    nrm = np.linalg.norm(A, 1)
    s = 0
    if nrm > 0.25:
        s = int(np.ceil(np.log2(nrm / 0.25)))
    B = A / (2.0 ** s)
    E = np.eye(A.shape[0], dtype=complex)
    term = np.eye(A.shape[0], dtype=complex)
    for k in range(1, 25):
        term = term @ B / k
        E = E + term
    for _ in range(s):
        E = E @ E
    return E


def exp(g: BondGate, ishermitian: bool = False) -> BondGate:
    """exp(G::BondGate) (src/bondgate.jl:24-25)."""
    return BondGate(expm(g.G), g.b)


class BondOperator:
    """src/bondop.jl:43-88."""

    def __init__(self, N: int, d: int = 2):
        self.N, self.d = N, d
        self.bondterms: Dict[int, List[BondTerm]] = {b: [] for b in range(N - 1)}
        self.siteterms: Dict[int, List[SiteTerm]] = {i: [] for i in range(N)}

    def __len__(self):
        return self.N

    def add(self, *args):
        """The four add! methods (src/bondop.jl:67-70): add(op,i) add(c,op,i) add(op1,op2,b) add(c,op1,op2,b)."""
        args = list(args)
        c = 1.0
        if not isinstance(args[0], str):
            c = args.pop(0)
        if len(args) == 2:
            self.siteterms[args[1]].append(SiteTerm(args[1], c, args[0]))
        elif len(args) == 3:
            self.bondterms[args[2]].append(BondTerm(args[2], c, args[0], args[1]))
        else:
            raise ValueError("add!: bad arguments %r" % (args,))
        return self


def add(bo: BondOperator, *args):
    return bo.add(*args)


def siteterms(bo: BondOperator, i: int):
    return bo.siteterms[i]


def bondterms(bo: BondOperator, b: int):
    return bo.bondterms[b]


def bondgate(bo: BondOperator, b: int) -> BondGate:
    """src/bondop.jl:72-86: single-site terms enter with 1/2 on interior sites, 1 at the chain ends."""
    d = bo.d
    I = np.eye(d, dtype=complex)
    g = np.zeros((d * d, d * d), dtype=complex)
    for t in bo.bondterms[b]:
        g += t.coeff * np.kron(t.leftop.matrix(), t.rightop.matrix())
    for i in (b, b + 1):
        fac = 1.0 if (i == bo.N - 1 or i == 0) else 0.5
        for st in bo.siteterms[i]:
            O = st.matrix()
            g += fac * (np.kron(O, I) if i == b else np.kron(I, O))
    return BondGate(g, b)


def gates(bo: BondOperator) -> GateList:
    return [bondgate(bo, b) for b in range(bo.N - 1)]


def tfi_bondop(N: int, J: float, h: float) -> BondOperator:
    """src/testutils.jl:43-52."""
    H = BondOperator(N)
    for b in range(N - 1):
        H.add(-J, "Sz", "Sz", b)
        H.add(-h, "Sx", b)
    H.add(-h, "Sx", N - 1)
    return H


def xxz_bondop(N: int, delta: float = 1.0, hz: float = 0.0) -> BondOperator:
    H = BondOperator(N)
    for b in range(N - 1):
        H.add(0.5, "S+", "S-", b)
        H.add(0.5, "S-", "S+", b)
        H.add(delta, "Sz", "Sz", b)
    if hz != 0.0:
        for i in range(N):
            H.add(-hz, "Sz", i)
    return H


def measure(H: Sequence[BondGate], psi: MPS) -> complex:
    """measure(H::GateList, psi) (src/bondgate.jl:71-79)."""
    e = 0.0
    for g in H:
        m = np.ascontiguousarray(g.G.astype(np.complex128))
        out = tevo_complex()
        check(load().tevo_gate_expect(psi.h, g.b, m.ctypes.data, C.byref(out)), psi.ctx.h)
        e += complex(out.re, out.im)
    return e
