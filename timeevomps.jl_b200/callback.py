"""Observer protocol of the time evolution (src/callback.jl).

apply(psi; t, bond, spec, sweepdir, sweepend, alg[, meas]) / checkdone(psi) / callback_dt().
The built-in callbacks never pull the state to the host: LocalMeasurementCallback receives the local
expectation values computed on the device right after the update of `bond` (tevo_apply_layer's `meas` or
tevo_measure_bond), SpecCallback reads the device-computed Spectrum summary.  A user callback that wants the
state itself sets `needs_state = True` (the default of the base class) and is then called gate by gate."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np


def _approx(a, b):
    """Julia isapprox default."""
    return abs(a - b) <= math.sqrt(np.finfo(float).eps) * max(abs(a), abs(b))


def _real(x):
    return x.real if isinstance(x, complex) else x


class TEvoCallback:
    needs_state = True     # called after every single gate with a coherent psi
    device_ops: Sequence[str] = ()

    def apply(self, psi, **kw):
        return None

    def checkdone(self, *a, **k):
        return False

    def callback_dt(self):
        return 0


class NoTEvoCallback(TEvoCallback):
    needs_state = False


def measurement_ts(cb):
    return cb.ts


def measurements(cb):
    return cb.measurements


class LocalMeasurementCallback(TEvoCallback):
    """src/callback.jl:33-100."""
    needs_state = False

    def __init__(self, ops: Sequence[str], N: int, dt_measure: float):
        self.ops = list(ops)
        self.device_ops = self.ops
        self.N = N
        self.measurements: Dict[str, List[np.ndarray]] = {o: [] for o in self.ops}
        self.ts: List[float] = []
        self.dt_measure = dt_measure

    def callback_dt(self):
        return self.dt_measure

    def wants(self, t, sweepend, bond, alg) -> bool:
        """The gating condition of src/callback.jl:84 (0-based bond: reference bond = bond+1)."""
        from .tebd import TEBDalg
        t = _real(t)
        prev_t = self.ts[-1] if self.ts else 0
        return bool((_approx(t - prev_t, self.dt_measure) or t == prev_t) and sweepend and
                    ((bond + 1) % 2 == 1 or not isinstance(alg, TEBDalg)))

    def apply(self, psi, *, t, sweepend, sweepdir, bond, alg, meas: Optional[np.ndarray] = None, **kw):
        from .tebd import TEBDalg
        if not self.wants(t, sweepend, bond, alg):
            return
        t = _real(t)
        prev_t = self.ts[-1] if self.ts else 0
        if t != prev_t:
            self.ts.append(t)
            for o in self.ops:
                self.measurements[o].append(np.zeros(self.N))
        if meas is None:
            meas = psi.measure_bond(bond, self.ops)
        for k, o in enumerate(self.ops):
            m = meas[k]
            self.measurements[o][-1][bond + 1] = m[1].real
            if isinstance(alg, TEBDalg) or bond == 0:
                self.measurements[o][-1][bond] = m[0].real


class SpecCallback(TEvoCallback):
    """src/callback.jl:102-162."""
    needs_state = False

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

    @property
    def measurements(self):
        return {"entropy": self.entropies, "bonddim": self.bonddims, "truncerrs": self.truncerrs}

    def callback_dt(self):
        return self.dt_measure

    def apply(self, psi, *, t, sweepend, bond, spec, sweepdir, **kw):
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
                self.bonddims[-1][i] = len(spec)
                self.entropies[-1][i] = spec.entropy()
            if sweepdir == "right" and bond == self.N - 2:
                self.truncerrs.append(self.current_truncerr)
            elif sweepdir == "left" and bond == 0:
                self.truncerrs.append(self.current_truncerr)
