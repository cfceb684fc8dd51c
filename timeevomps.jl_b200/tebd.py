"""TEBD: apply_gate! / apply_gates! (src/bondgate.jl:46-63) and the even/odd Trotter layer scheduler tebd!
(src/tebd.jl).  Host-side mirror of the reference interface; all arithmetic is in libtevo."""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import numpy as np

from ._lib import check, load, make_trunc, tevo_spec
from .bondop import BondGate, BondOperator, gates
from .callback import NoTEvoCallback, TEvoCallback
from .mps import MPS, Spectrum, op


class TEBDalg:
    pass


class TEBD2(TEBDalg):
    """src/tebd.jl:7-24."""


class TEBD2Sweep(TEBDalg):
    """src/tebd.jl:26-40."""


class TEBD4(TEBDalg):
    """src/tebd.jl:43-54."""


_TRUNC_KEYS = ("maxdim", "mindim", "cutoff", "use_absolute_cutoff", "do_rel_cutoff")


def _trunc_from(kw):
    k = {x: kw[x] for x in _TRUNC_KEYS if x in kw}
    if "absoluteCutoff" in kw:  # the spelling of src/tdvp.jl:22
        k["use_absolute_cutoff"] = kw["absoluteCutoff"]
    return make_trunc(**k)


def apply_gate(psi: MPS, G: BondGate, *, want_eigs: bool = True, **kw) -> Spectrum:
    """apply_gate!(psi, G; kwargs...) -> Spectrum  (src/bondgate.jl:46-53)."""
    d = psi.d
    g = np.ascontiguousarray(G.G.astype(np.complex128))
    assert g.shape == (d * d, d * d)
    tr = _trunc_from(kw)
    spec = tevo_spec()
    ortho = kw.get("ortho", "left")
    if ortho not in ("left", "right"):
        raise ValueError("ortho must be \"left\" or \"right\"")
    cap = 0
    eigs = None
    if want_eigs:
        l, r = C.c_int(), C.c_int()
        check(load().tevo_mps_site_dims(psi.h, G.b, C.byref(l), None), psi.ctx.h)
        check(load().tevo_mps_site_dims(psi.h, G.b + 1, None, C.byref(r)), psi.ctx.h)
        cap = d * min(l.value, r.value)
        eigs = np.zeros(cap)
    check(load().tevo_apply_gate(psi.h, G.b, g.ctypes.data, C.byref(tr), 1 if ortho == "left" else 0,
                                 1 if kw.get("normalize", True) else 0, C.byref(spec),
                                 eigs.ctypes.data_as(C.POINTER(C.c_double)) if want_eigs else None, cap), psi.ctx.h)
    return Spectrum(spec, eigs[:spec.nkeep].copy() if want_eigs else None)


def apply_gates(psi: MPS, Gs: Sequence[BondGate], *, reversed: bool = False, cb=None, callback_obj=None,
                measure_ops: Sequence[str] = (), literal_ortho: bool = False, **kw):
    """apply_gates!(psi, Gs; reversed, cb, kwargs...)  (src/bondgate.jl:55-63).

    One tevo_apply_layer call applies the whole layer sequentially on the device; `cb(psi; bond, spec[, meas])`
    is then fired per gate in application order.  A callback that needs the coherent state after each gate
    (callback_obj.needs_state) forces the literal gate-by-gate path."""
    Gs = list(Gs)
    if not Gs:
        return
    if cb is not None and callback_obj is not None and getattr(callback_obj, "needs_state", True):
        order = Gs[::-1] if reversed else Gs
        for g in order:
            spec = apply_gate(psi, g, **kw)
            cb(psi, bond=g.b, spec=spec)
        return
    d = psi.d
    n = len(Gs)
    bonds = (C.c_int * n)(*[g.b for g in Gs])
    G = np.ascontiguousarray(np.stack([g.G for g in Gs]).astype(np.complex128))
    tr = _trunc_from(kw)
    specs = (tevo_spec * n)()
    nops = len(measure_ops) if cb is not None else 0
    ops_arr = np.ascontiguousarray(np.stack([op(o) for o in measure_ops]).astype(np.complex128)) if nops else None
    meas = np.zeros((n, max(nops, 1), 2), dtype=np.complex128)
    check(load().tevo_apply_layer(psi.h, n, bonds, G.ctypes.data, 1 if reversed else 0, C.byref(tr),
                                  1 if kw.get("normalize", True) else 0, 1 if literal_ortho else 0, specs, None, 0,
                                  nops, ops_arr.ctypes.data if nops else None, meas.ctypes.data if nops else None),
          psi.ctx.h)
    if cb is not None:
        order = range(n - 1, -1, -1) if reversed else range(n)
        for g in order:
            if nops:
                cb(psi, bond=Gs[g].b, spec=Spectrum(specs[g]), meas=meas[g])
            else:
                cb(psi, bond=Gs[g].b, spec=Spectrum(specs[g]))


def time_evo_gates(dt, H: Sequence[BondGate], alg: TEBDalg, ishermitian: bool = False):
    """src/tebd.jl:56-87.  Julia H[1:2:end] (odd bonds, 1-based) = H[0::2] here."""
    from .bondop import exp as gexp
    H = list(H)
    o, e = H[0::2], H[1::2]
    ex = lambda c, gs: [gexp(c * g, ishermitian=ishermitian) for g in gs]
    if isinstance(alg, TEBD2):
        Uhalf = ex(-1j * dt / 2, o)
        Us = [ex(-1j * dt, e), ex(-1j * dt, o)]
        return [Uhalf], Us, [Us[0], Uhalf]
    if isinstance(alg, TEBD2Sweep):
        Us = [ex(-1j * dt / 2, H), ex(-1j * dt / 2, H)]
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
    raise TypeError("unknown TEBD algorithm %r" % (alg,))


def _nsteps(tf, dt) -> int:
    """Int(tf/dt) (src/tebd.jl:95): raises like Julia's InexactError when not integral."""
    q = tf / dt
    if isinstance(q, complex):
        if q.imag != 0:
            raise ValueError("InexactError: Int(%r)" % (q,))
        q = q.real
    if q != int(q):
        raise ValueError("InexactError: Int(%r)" % (q,))
    return int(q)


def _real(x):
    return x.real if isinstance(x, complex) else x


def tebd(psi: MPS, H, dt, tf, alg: Optional[TEBDalg] = None, *, callback: Optional[TEvoCallback] = None,
         orthogonalize: int = 0, progress: bool = False, ishermitian: bool = False, **kw) -> MPS:
    """tebd!(psi, H, dt, tf, alg=TEBD2(); kwargs...) (src/tebd.jl:89-179).  Mutates and returns psi."""
    alg = TEBD2() if alg is None else alg
    if isinstance(H, BondOperator):
        H = gates(H)
    nsteps = _nsteps(tf, dt)
    cb = NoTEvoCallback() if callback is None else callback

    def cb_func(dt_, step, rev, se):
        def f(p, bond, spec, **extra):
            cb.apply(p, t=dt_ * step, bond=bond, spec=spec, sweepdir="left" if rev else "right", sweepend=se,
                     alg=alg, **extra)
        return f

    dtm = cb.callback_dt()
    if dtm > 0:
        ratio = dtm / dt
        if isinstance(ratio, complex) or math.floor(ratio) != ratio:
            raise ValueError("Measurement time step %s incommensurate with time-evolution time step %s" % (dtm, dt))
        mstep = math.floor(ratio)
        nbunch = math.gcd(mstep, nsteps)
    else:
        nbunch = nsteps
    if orthogonalize > 0:
        nbunch = math.gcd(nbunch, orthogonalize)
    Ustart, Us, Uend = time_evo_gates(dt, H, alg, ishermitian=ishermitian)
    if len(Ustart) == 0 and len(Uend) == 0:
        nbunch += 1
    is_noop = isinstance(cb, NoTEvoCallback)
    dev_ops = tuple(getattr(cb, "device_ops", ()))
    step = 0
    rev = False
    while step < nsteps:
        for U in Ustart:
            apply_gates(psi, U, reversed=rev, **kw)
            rev = not rev
        for _ in range(nbunch - 1):
            for U in Us:
                apply_gates(psi, U, reversed=rev, cb=None if is_noop else cb_func(dt, step, rev, False),
                            callback_obj=cb, **kw)
                rev = not rev
            step += 1
            if cb.checkdone(psi):
                break
        for i, U in enumerate(Uend):
            se = (i + 1) >= len(Uend) - 1
            apply_gates(psi, U, reversed=rev, cb=None if is_noop else cb_func(dt, step + 1, rev, se),
                        callback_obj=cb, measure_ops=dev_ops if se else (), **kw)
            rev = not rev
        if len(Uend) > 0:
            step += 1
        if orthogonalize > 0 and step % orthogonalize == 0:
            psi.reorthogonalize()
        if cb.checkdone(psi):
            break
    return psi


# ------------------------------------------------------------------------------------------------------------
# layer-parallel mode (NOT the reference's sequential algorithm; the multi-GPU building block)
# ------------------------------------------------------------------------------------------------------------
def apply_gates_parallel(psi: MPS, Gs: Sequence[BondGate], **kw):
    """All gates of one even/odd layer at once on a B-form state (psi.to_bform()); returns the Spectrum of each."""
    Gs = list(Gs)
    if not Gs:
        return []
    n = len(Gs)
    bonds = (C.c_int * n)(*[g.b for g in Gs])
    G = np.ascontiguousarray(np.stack([g.G for g in Gs]).astype(np.complex128))
    tr = _trunc_from(kw)
    specs = (tevo_spec * n)()
    check(load().tevo_apply_layer_bform(psi.h, n, bonds, G.ctypes.data, C.byref(tr), 1 if kw.get("normalize", True) else 0,
                                        specs, None, 0), psi.ctx.h)
    return [Spectrum(specs[g]) for g in range(n)]


def tebd_parallel(psi: MPS, H, dt, tf, alg: Optional[TEBDalg] = None, **kw) -> MPS:
    """The schedule of tebd! (bunched half steps, no callbacks) driving layer-parallel updates of a B-form state."""
    alg = TEBD2() if alg is None else alg
    if isinstance(alg, TEBD2Sweep):
        raise ValueError("tebd_parallel: TEBD2Sweep layers contain adjacent bonds and are not layer-parallel; "
                         "use tebd() (sequential) for TEBD2Sweep")
    if isinstance(H, BondOperator):
        H = gates(H)
    nsteps = _nsteps(tf, dt)
    Ustart, Us, Uend = time_evo_gates(dt, H, alg)
    nbunch = nsteps + (1 if (len(Ustart) == 0 and len(Uend) == 0) else 0)
    step = 0
    while step < nsteps:
        for U in Ustart:
            apply_gates_parallel(psi, U, **kw)
        for _ in range(nbunch - 1):
            for U in Us:
                apply_gates_parallel(psi, U, **kw)
            step += 1
        for U in Uend:
            apply_gates_parallel(psi, U, **kw)
        if len(Uend) > 0:
            step += 1
    return psi
