"""Two-site TDVP sweep scheduler tdvp! (src/tdvp.jl:37-96); arithmetic in libtevo."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from ._lib import (TevoError, TEVO_ENOTCONVERGED, check, load, tevo_complex, tevo_krylov, tevo_krylov_info,
                   tevo_spec)
from .callback import NoTEvoCallback, TEvoCallback
from .mps import MPO, MPS, Spectrum
from .tebd import _nsteps, _trunc_from


class TDVP2:
    """src/tdvp.jl:7."""


class ProjMPO:
    """ITensors.ProjMPO handle: MPO + cached environments (src/tdvp.jl:52-53)."""

    def __init__(self, H: MPO):
        self.H = H
        self.ctx = H.ctx
        h = C.c_void_p()
        check(load().tevo_env_create(H.h, C.byref(h)), self.ctx.h)
        self.h = h

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.ctx.h:
                load().tevo_env_free(self.h)
                self.h = None
        except Exception:
            pass

    def position(self, psi: MPS, pos: int, nsite: int = 2):
        check(load().tevo_env_position(self.h, psi.h, pos, nsite), self.ctx.h)

    def product(self, psi: MPS, pos: int, nsite: int, v: np.ndarray) -> np.ndarray:
        v = np.asfortranarray(np.asarray(v, dtype=np.complex128))
        out = np.empty(v.shape, dtype=np.complex128, order="F")
        check(load().tevo_env_apply_host(self.h, psi.h, pos, nsite, v.ctypes.data, out.ctypes.data), self.ctx.h)
        return out


def sweepnext(N: int):
    """ITensors.sweepnext(N), 0-based bonds."""
    for b in range(0, N - 1):
        yield b, 1
    for b in range(N - 2, -1, -1):
        yield b, 2


def _cplx(z) -> tevo_complex:
    z = complex(z)
    return tevo_complex(z.real, z.imag)


def tdvp(psi: MPS, H: MPO, dt, tf, *, callback: Optional[TEvoCallback] = None, hermitian: bool = True,
         exp_tol: float = 1e-14, krylovdim: int = 30, maxiter: int = 100, normalize: bool = True,
         progress: bool = False, stats: Optional[dict] = None, **kw):
    """tdvp!(psi, H::MPO, dt, tf; kwargs...) (src/tdvp.jl:37-96)."""
    nsteps = _nsteps(tf, dt)
    cb = NoTEvoCallback() if callback is None else callback
    tau = 1j * dt
    if isinstance(tau, complex) and tau.imag == 0:
        tau = tau.real
    imag_time = isinstance(dt, complex)
    N = psi.N
    lib = load()
    psi.orthogonalize(0)
    PH = ProjMPO(H)
    PH.position(psi, 0, 2)
    tr = _trunc_from(kw)
    kp2 = tevo_krylov(1 if hermitian else 0, float(exp_tol), int(krylovdim), 100)  # :61 does not forward maxiter
    kp1 = tevo_krylov(1 if hermitian else 0, float(exp_tol), int(krylovdim), int(maxiter))
    is_noop = isinstance(cb, NoTEvoCallback)
    for s in range(1, nsteps + 1):
        for (b, ha) in sweepnext(N):
            spec = tevo_spec()
            info = tevo_krylov_info()
            st = lib.tevo_tdvp_twosite(PH.h, psi.h, b, _cplx(-tau / 2), C.byref(kp2), C.byref(tr),
                                       1 if ha == 1 else 0, 1 if normalize else 0, C.byref(spec), None, 0,
                                       C.byref(info))
            if st == TEVO_ENOTCONVERGED:
                raise RuntimeError("exponentiate did not converge")
            check(st, psi.ctx.h)
            if stats is not None:
                stats["matvec2"] = stats.get("matvec2", 0) + info.numops
            if not is_noop:
                cb.apply(psi, t=s * dt, bond=b, sweepend=(ha == 2), sweepdir="right" if ha == 1 else "left",
                         spec=Spectrum(spec), alg=TDVP2())
            i = b + 1 if ha == 1 else b
            if 0 < i < N - 1 and not imag_time:
                info = tevo_krylov_info()
                st = lib.tevo_tdvp_onesite(PH.h, psi.h, i, _cplx(tau / 2), C.byref(kp1), C.byref(info))
                if st == TEVO_ENOTCONVERGED:
                    raise RuntimeError("exponentiate did not converge")
                check(st, psi.ctx.h)
                if stats is not None:
                    stats["matvec1"] = stats.get("matvec1", 0) + info.numops
            elif i == 0 and imag_time:
                psi.normalize_site(i)
        if cb.checkdone():
            break
    return None
