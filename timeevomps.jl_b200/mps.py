"""Device-resident MPS / MPO handles (the ITensors `MPS` / `MPO` arguments of tebd! / tdvp!).

Sites and bonds are 0-based in this Python host layer; bond b joins sites b and b+1."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import check, load, tevo_complex, tevo_spec

# S=1/2 operators as ITensors' "S=1/2" site type defines them (state 0 = "Up")
_OPS = {
    "Id": np.eye(2, dtype=complex),
    "Sz": np.array([[0.5, 0], [0, -0.5]], dtype=complex),
    "Sx": np.array([[0, 0.5], [0.5, 0]], dtype=complex),
    "Sy": np.array([[0, -0.5j], [0.5j, 0]], dtype=complex),
    "S+": np.array([[0, 1], [0, 0]], dtype=complex),
    "S-": np.array([[0, 0], [1, 0]], dtype=complex),
    "N": np.array([[0, 0], [0, 1]], dtype=complex),
    "Cdag": np.array([[0, 0], [1, 0]], dtype=complex),
    "C": np.array([[0, 1], [0, 0]], dtype=complex),
}


def op(name) -> np.ndarray:
    """op(sites, name, i) for S=1/2 sites, with "A*B" products as src/bondop.jl:17-25."""
    if not isinstance(name, str):
        return np.asarray(name, dtype=complex)
    parts = name.split("*")
    o = _OPS[parts[0]].copy()
    for p in parts[1:]:
        o = o @ _OPS[p]
    return o


def _addr(a: np.ndarray) -> int:
    return a.ctypes.data


class Spectrum:
    """ITensors.Spectrum of one update: kept sigma^2 (`eigs`), `truncerr`, `entropy()`."""

    def __init__(self, s: tevo_spec, eigs: Optional[np.ndarray] = None):
        self.nkeep = s.nkeep
        self.nfull = s.nfull
        self.truncerr = s.truncerr
        self._entropy = s.entropy
        self.sum_all = s.sum_all
        self.sum_kept = s.sum_kept
        self.eigs = eigs

    def __len__(self):
        return self.nkeep

    def entropy(self) -> float:
        return self._entropy


class MPS:
    def __init__(self, N: int, d: int = 2, ctx: Optional[_lib.Context] = None, _handle=None):
        self.ctx = ctx or _lib.default_ctx()
        self.N, self.d = N, d
        if _handle is None:
            h = C.c_void_p()
            check(load().tevo_mps_create(self.ctx.h, N, d, C.byref(h)), self.ctx.h)
            _handle = h
        self.h = _handle

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.ctx.h:
                load().tevo_mps_free(self.h)
                self.h = None
        except Exception:
            pass

    def __len__(self):
        return self.N

    # ---- constructors ----
    @classmethod
    def from_tensors(cls, tensors: Sequence[np.ndarray], ctx=None, llim: Optional[int] = None,
                     rlim: Optional[int] = None) -> "MPS":
        N = len(tensors)
        d = tensors[0].shape[1]
        psi = cls(N, d, ctx)
        for j, t in enumerate(tensors):
            psi.set_site(j, t)
        if llim is not None or rlim is not None:
            psi.set_limits(-1 if llim is None else llim, N if rlim is None else rlim)
        return psi

    @classmethod
    def product(cls, states: Sequence[int], d: int = 2, ctx=None) -> "MPS":
        """productMPS(sites, states): 0 = "Up"."""
        ts = []
        for s in states:
            t = np.zeros((1, d, 1), dtype=complex)
            t[0, int(s), 0] = 1.0
            ts.append(t)
        return cls.from_tensors(ts, ctx, llim=-1, rlim=1)

    @classmethod
    def random(cls, N: int, chi: int, d: int = 2, seed: int = 0, ctx=None) -> "MPS":
        """Random saturated, canonical (centre 0), normalised MPS generated on the device."""
        psi = cls(N, d, ctx)
        check(load().tevo_mps_random(psi.h, chi, seed), psi.ctx.h)
        return psi

    def copy(self) -> "MPS":
        h = C.c_void_p()
        check(load().tevo_mps_copy(self.h, C.byref(h)), self.ctx.h)
        return MPS(self.N, self.d, self.ctx, _handle=h)

    # ---- marshalling ----
    def set_site(self, j: int, t: np.ndarray):
        t = np.asfortranarray(np.asarray(t, dtype=np.complex128))
        assert t.ndim == 3 and t.shape[1] == self.d
        check(load().tevo_mps_upload_site(self.h, j, t.shape[0], t.shape[2], _addr(t)), self.ctx.h)

    def site(self, j: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Site tensor j as a (chi_l, d, chi_r) column-major array.  `out`: a caller-owned complex128 F-ordered array of
        that shape (e.g. a view of pinned host memory, which the device copies into directly); a fresh array
        otherwise (also when the bond dimensions no longer match `out`)."""
        l, r = C.c_int(), C.c_int()
        check(load().tevo_mps_site_dims(self.h, j, C.byref(l), C.byref(r)), self.ctx.h)
        shape = (l.value, self.d, r.value)
        if (out is None or out.shape != shape or out.dtype != np.complex128 or not out.flags["F_CONTIGUOUS"]
                or not out.flags["WRITEABLE"]):
            out = np.empty(shape, dtype=np.complex128, order="F")
        check(load().tevo_mps_download_site(self.h, j, _addr(out), out.size), self.ctx.h)
        return out

    def tensors(self, out: Optional[Sequence[np.ndarray]] = None) -> List[np.ndarray]:
        """All site tensors; `out` (optional) is a list of per-site buffers reused where the shapes still match."""
        return [self.site(j, None if out is None else out[j]) for j in range(self.N)]

    def bond_dims(self) -> List[int]:
        arr = (C.c_int * (self.N - 1))()
        check(load().tevo_mps_bonddims(self.h, arr), self.ctx.h)
        return list(arr)

    def maxlinkdim(self) -> int:
        return max(self.bond_dims())

    def limits(self):
        a, b = C.c_int(), C.c_int()
        check(load().tevo_mps_get_limits(self.h, C.byref(a), C.byref(b)), self.ctx.h)
        return a.value, b.value

    def set_limits(self, llim: int, rlim: int):
        check(load().tevo_mps_set_limits(self.h, llim, rlim), self.ctx.h)

    # ---- gauge ----
    def orthogonalize(self, j: int):
        """orthogonalize!(psi, j)."""
        check(load().tevo_mps_orthogonalize(self.h, j), self.ctx.h)

    def reorthogonalize(self):
        """reorthogonalize!(psi) (src/itensor.jl:28-33)."""
        check(load().tevo_mps_reorthogonalize(self.h), self.ctx.h)

    def normalize_site(self, j: int):
        check(load().tevo_mps_normalize_site(self.h, j), self.ctx.h)

    # ---- layer-parallel (B-form) mode ----
    def to_bform(self):
        """Canonical form B_j + Schmidt values on every bond (input of the layer-parallel / sharded TEBD mode)."""
        check(load().tevo_mps_to_bform(self.h), self.ctx.h)
        return self

    def schmidt_values(self, j: int) -> np.ndarray:
        """Schmidt values on the bond left of site j (B-form only)."""
        n = C.c_int()
        check(load().tevo_mps_get_lambda(self.h, j, None, 0, C.byref(n)), self.ctx.h)
        out = np.zeros(n.value)
        check(load().tevo_mps_get_lambda(self.h, j, out.ctypes.data_as(C.POINTER(C.c_double)), n.value, None),
              self.ctx.h)
        return out

    def bform_expect(self, ops: Sequence, j: int) -> np.ndarray:
        mats = np.ascontiguousarray(np.stack([op(o) for o in ops]).astype(np.complex128))
        out = np.zeros(len(ops), dtype=np.complex128)
        check(load().tevo_bform_expect_site(self.h, j, len(ops), _addr(mats), _addr(out)), self.ctx.h)
        return out

    def bform_expect_all(self, o) -> np.ndarray:
        return np.array([self.bform_expect([o], j)[0].real for j in range(self.N)])

    # ---- observables ----
    def measure_bond(self, b: int, ops: Sequence) -> np.ndarray:
        """<O> on sites b, b+1 from theta = psi[b]*psi[b+1] (src/callback.jl:65-73).  Returns (nops, 2) complex."""
        mats = np.ascontiguousarray(np.stack([op(o) for o in ops]).astype(np.complex128))
        out = np.zeros((len(ops), 2), dtype=np.complex128)
        check(load().tevo_measure_bond(self.h, b, len(ops), _addr(mats), _addr(out)), self.ctx.h)
        return out

    def expect(self, o, j: int) -> complex:
        """measure!(psi, op, j) (src/testutils.jl:55-58)."""
        m = np.ascontiguousarray(op(o).astype(np.complex128))
        out = tevo_complex()
        check(load().tevo_mps_expect_site(self.h, j, _addr(m), C.byref(out)), self.ctx.h)
        return complex(out.re, out.im)

    def expect_all(self, o) -> np.ndarray:
        return np.array([self.expect(o, j).real for j in range(self.N)])


def inner(a: MPS, b: MPS) -> complex:
    out = tevo_complex()
    check(load().tevo_mps_inner(a.h, b.h, C.byref(out)), a.ctx.h)
    return complex(out.re, out.im)


class MPO:
    """ITensors MPO argument of tdvp!: list of (wl, d_out, d_in, wr) tensors."""

    def __init__(self, tensors: Sequence[np.ndarray], ctx=None):
        self.ctx = ctx or _lib.default_ctx()
        self.N = len(tensors)
        self.d = tensors[0].shape[1]
        h = C.c_void_p()
        check(load().tevo_mpo_create(self.ctx.h, self.N, self.d, C.byref(h)), self.ctx.h)
        self.h = h
        for j, W in enumerate(tensors):
            W = np.asfortranarray(np.asarray(W, dtype=np.complex128))
            check(load().tevo_mpo_upload_site(self.h, j, W.shape[0], W.shape[3], _addr(W)), self.ctx.h)

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.ctx.h:
                load().tevo_mpo_free(self.h)
                self.h = None
        except Exception:
            pass


def inner_mpo(a: MPS, H: MPO, b: MPS) -> complex:
    """inner(a, H, b)."""
    out = tevo_complex()
    check(load().tevo_mps_inner_mpo(a.h, H.h, b.h, C.byref(out)), a.ctx.h)
    return complex(out.re, out.im)


def _bulk_to_mpo(Wb: np.ndarray, N: int) -> List[np.ndarray]:
    out = []
    for j in range(N):
        W = Wb
        if j == 0:
            W = W[-1:, :, :, :]
        if j == N - 1:
            W = W[:, :, :, :1]
        out.append(W.copy())
    return out


def tfi_mpo(J: float, h: float, N: int, ctx=None) -> MPO:
    """tfi_mpo(J, h, sites) (src/testutils.jl:4-12): AutoMPO of -J sum SzSz - h sum Sx, written analytically."""
    I, Sz, Sx = op("Id"), op("Sz"), op("Sx")
    W = np.zeros((3, 2, 2, 3), dtype=complex)
    W[0, :, :, 0] = I
    W[1, :, :, 0] = Sz
    W[2, :, :, 0] = -h * Sx
    W[2, :, :, 1] = -J * Sz
    W[2, :, :, 2] = I
    return MPO(_bulk_to_mpo(W, N), ctx)


def xxz_mpo(N: int, delta: float = 1.0, hz: float = 0.0, ctx=None) -> MPO:
    """XXZ chain sum [SxSx + SySy + delta SzSz] - hz sum Sz, MPO bond dimension 5."""
    I, Sz, Sp, Sm = op("Id"), op("Sz"), op("S+"), op("S-")
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
    return MPO(_bulk_to_mpo(W, N), ctx)
