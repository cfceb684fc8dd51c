"""The Trotter layer schedule of tebd! pinned to traces written down BY HAND from src/tebd.jl:56-87 (gate sequences)
and src/tebd.jl:113-177 (bunching, direction alternation, callback arguments) for a 4-site chain - independent of
both implementations.  The product's host scheduler (timeevomps.jl_b200/tebd.py) and the oracle's
(oracle/tevo_oracle.py) are each run with `apply_gates` replaced by a recorder, so no GPU and no arithmetic is
involved: a schedule bug shared by the two restatements would show up here.

Events, in order of occurrence:
  ("layer", bonds as applied, tau of the layer, had_cb)      one apply_gates! call (src/bondgate.jl:55-63)
  ("cb", bond, t, sweepdir, sweepend)                        one callback invocation (src/tebd.jl:103-109)
  ("reorth",)                                                reorthogonalize!(psi) (src/tebd.jl:174)
Bonds are 0-based: the reference's odd bonds 1, 3 are 0 and 2 here, its even bond 2 is 1."""
import math

import numpy as np
import pytest

DT = 0.1
T1 = DT / (4 - 4 ** (1 / 3))      # tau_1 = tau_2 of TEBD4 (src/tebd.jl:70-71)
T3 = DT - 4 * T1                  # tau_3 (src/tebd.jl:72)
ODD, EVEN = (0, 2), (1,)


class Recorder:
    """A TEvoCallback that records its arguments; dt_measure = 0 means "no measurement step"."""
    needs_state = True
    device_ops = ()

    def __init__(self, events, dt_measure=0.0):
        self.events = events
        self.dt_measure = dt_measure

    def apply(self, psi, *, t, bond, spec, sweepdir, sweepend, alg, **kw):
        self.events.append(("cb", bond, round(float(np.real(t)), 12), sweepdir, bool(sweepend)))

    def checkdone(self, *a, **k):
        return False

    def callback_dt(self):
        return self.dt_measure


class FakePsi:
    def __init__(self, events):
        self.events = events

    def reorthogonalize(self):
        self.events.append(("reorth",))


def _tau_of(G):
    """The layers are exp(-i tau Sz.Sz): the (up,up) entry is exp(-i tau / 4)."""
    return round(float(-4.0 * np.angle(G[0, 0])), 12)


def _run(mod, is_oracle, alg_name, nsteps, dt_measure=0.0, orthogonalize=0):
    events = []

    def fake_apply_gates(psi, Gs, *, reversed=False, cb=None, **kw):
        Gs = list(Gs)
        order = Gs[::-1] if reversed else Gs
        events.append(("layer", tuple(g.b for g in order), _tau_of(np.asarray(Gs[0].G)), cb is not None))
        for g in order:
            if cb is not None:
                cb(psi, bond=g.b, spec=None)

    if is_oracle:
        H = mod.BondOperator(4)
        for b in range(3):
            H.add(1.0, "Sz", "Sz", b)
        target = mod
    else:
        H = mod.BondOperator(4)
        for b in range(3):
            mod.add(H, 1.0, "Sz", "Sz", b)
        import sys
        target = sys.modules["timeevomps_jl_b200.tebd"]   # (the package attribute `tebd` is the function)
    saved = target.apply_gates
    target.apply_gates = fake_apply_gates
    try:
        mod.tebd(FakePsi(events), H, DT, DT * nsteps, getattr(mod, alg_name)(), callback=Recorder(events, dt_measure),
                 orthogonalize=orthogonalize)
    finally:
        target.apply_gates = saved
    return events


def _layer(bonds, tau, cb):
    return ("layer", tuple(bonds), round(tau, 12), cb)


def _cbs(bonds, t, d, se):
    return [("cb", b, round(t, 12), d, se) for b in bonds]


def expected_tebd2_two_steps_bunched():
    """TEBD2, nsteps = 2, no measurement step: nbunch = 2 (src/tebd.jl:119).  Ustart = [o(dt/2)], one bunched step
    Us = [e(dt), o(dt)] with t = dt*0 and sweepend = false, then Uend = [e(dt), o(dt/2)] with t = dt*2, both layers
    sweepend (i >= length(Uend)-1 holds for i = 1, 2)."""
    ev = [_layer(ODD, DT / 2, False)]                                   # rev: false -> true
    ev += [_layer(EVEN, DT, True)] + _cbs(EVEN, 0.0, "left", False)     # reversed; rev -> false
    ev += [_layer(ODD, DT, True)] + _cbs(ODD, 0.0, "right", False)      # rev -> true; step = 1
    ev += [_layer(EVEN, DT, True)] + _cbs(EVEN, 2 * DT, "left", True)   # Uend[1]; rev -> false
    ev += [_layer(ODD, DT / 2, True)] + _cbs(ODD, 2 * DT, "right", True)
    return ev


def expected_tebd2_measure_every_step():
    """TEBD2, nsteps = 2, dt_measure = dt: mstep = 1, nbunch = gcd(1, 2) = 1: every step is Ustart + Uend, and the
    direction keeps alternating ACROSS steps (rev is never reset, src/tebd.jl:130-135)."""
    ev = [_layer(ODD, DT / 2, False)]
    ev += [_layer(EVEN, DT, True)] + _cbs(EVEN, DT, "left", True)
    ev += [_layer(ODD, DT / 2, True)] + _cbs(ODD, DT, "right", True)              # rev -> true; step = 1
    ev += [_layer(ODD[::-1], DT / 2, False)]                                       # Ustart reversed; rev -> false
    ev += [_layer(EVEN, DT, True)] + _cbs(EVEN, 2 * DT, "right", True)             # rev -> true
    ev += [_layer(ODD[::-1], DT / 2, True)] + _cbs(ODD[::-1], 2 * DT, "left", True)
    return ev


TEBD4_SEQ = [(T1, EVEN), (T1, ODD), (T1, EVEN), ((T1 + T3) / 2, ODD), (T3, EVEN),
             ((T1 + T3) / 2, ODD), (T1, EVEN), (T1, ODD), (T1, EVEN), (T1, ODD)]      # src/tebd.jl:78-79 with tau2 = tau1


def expected_tebd4(nsteps):
    """TEBD4 without measurement step: nbunch = nsteps.  Ustart = [o(tau1/2)] (forward), nsteps-1 bunched steps of the
    10 layers (t = dt*step before the increment, sweepend false), then the 10 layers of Uend (last one o(tau1/2),
    t = dt*nsteps, sweepend for the last two).  Starting from rev = true after Ustart, even layers run reversed
    ("left"), odd layers forward ("right"); ten flips per step leave rev unchanged."""
    ev = [_layer(ODD, T1 / 2, False)]
    for step in range(nsteps - 1):
        for (tau, bonds) in TEBD4_SEQ:
            rev = bonds == EVEN
            bb = bonds[::-1] if rev else bonds
            ev += [_layer(bb, tau, True)] + _cbs(bb, DT * step, "left" if rev else "right", False)
    end = TEBD4_SEQ[:-1] + [(T1 / 2, ODD)]
    for i, (tau, bonds) in enumerate(end):
        rev = bonds == EVEN
        bb = bonds[::-1] if rev else bonds
        ev += [_layer(bb, tau, True)] + _cbs(bb, DT * nsteps, "left" if rev else "right", i >= 8)
    return ev


def expected_tebd2sweep_two_steps():
    """TEBD2Sweep: Ustart = Uend = [] so nbunch = nsteps + 1 = 3 (src/tebd.jl:125); each step is a forward sweep and a
    backward sweep over ALL bonds at dt/2; callbacks only ever see sweepend = false (src/tebd.jl:143)."""
    ev = []
    for step in range(2):
        ev += [_layer((0, 1, 2), DT / 2, True)] + _cbs((0, 1, 2), DT * step, "right", False)
        ev += [_layer((2, 1, 0), DT / 2, True)] + _cbs((2, 1, 0), DT * step, "left", False)
    return ev


def expected_tebd2_orthogonalize():
    """TEBD2, nsteps = 4, orthogonalize = 2, no measurement step: nbunch = gcd(4, 2) = 2, so the loop body (Ustart, one
    bunched step, Uend) runs twice and reorthogonalize! fires after steps 2 and 4 (src/tebd.jl:121,174).  The second
    body starts with rev = true: all its directions are mirrored."""
    ev = []
    rev = False
    step = 0
    for _ in range(2):
        def lay(bonds, tau, t, se, cb=True):
            nonlocal rev
            bb = bonds[::-1] if rev else bonds
            out = [_layer(bb, tau, cb)] + (_cbs(bb, t, "left" if rev else "right", se) if cb else [])
            rev = not rev
            return out
        ev += lay(ODD, DT / 2, None, None, cb=False)
        ev += lay(EVEN, DT, DT * step, False)
        ev += lay(ODD, DT, DT * step, False)
        step += 1
        ev += lay(EVEN, DT, DT * (step + 1), True)
        ev += lay(ODD, DT / 2, DT * (step + 1), True)
        step += 1
        ev.append(("reorth",))
    return ev


CASES = [
    ("TEBD2", 2, 0.0, 0, expected_tebd2_two_steps_bunched),
    ("TEBD2", 2, DT, 0, expected_tebd2_measure_every_step),
    ("TEBD4", 1, 0.0, 0, lambda: expected_tebd4(1)),
    ("TEBD4", 2, 0.0, 0, lambda: expected_tebd4(2)),
    ("TEBD2Sweep", 2, 0.0, 0, expected_tebd2sweep_two_steps),
    ("TEBD2Sweep", 2, DT, 0, expected_tebd2sweep_two_steps),     # nbunch = gcd(1,2)+1 = 2: same trace
    ("TEBD2", 4, 0.0, 2, expected_tebd2_orthogonalize),
]


def _compare(got, want):
    assert len(got) == len(want), (len(got), len(want), got[:6], want[:6])
    for i, (g, w) in enumerate(zip(got, want)):
        if g[0] == "layer":
            assert w[0] == "layer" and g[1] == w[1] and g[3] == w[3] and abs(g[2] - w[2]) < 1e-9, (i, g, w)
        elif g[0] == "cb":
            assert w[0] == "cb" and g[1] == w[1] and abs(g[2] - w[2]) < 1e-9 and g[3:] == w[3:], (i, g, w)
        else:
            assert g == w, (i, g, w)


@pytest.mark.parametrize("alg,nsteps,dtm,orth,expected", CASES)
def test_product_scheduler_matches_hand_trace(tevo, alg, nsteps, dtm, orth, expected):
    _compare(_run(tevo, False, alg, nsteps, dtm, orth), expected())


@pytest.mark.parametrize("alg,nsteps,dtm,orth,expected", CASES)
def test_oracle_scheduler_matches_hand_trace(oracle, alg, nsteps, dtm, orth, expected):
    _compare(_run(oracle, True, alg, nsteps, dtm, orth), expected())


def test_incommensurate_measurement_step_throws(tevo, oracle):
    """src/tebd.jl:115."""
    for mod, is_o in ((tevo, False), (oracle, True)):
        with pytest.raises(ValueError, match="incommensurate"):
            _run(mod, is_o, "TEBD2", 2, dt_measure=0.15)


def test_tebd4_time_steps_sum_to_dt():
    """src/tebd.jl:70-84: the ten layers of a bulk step advance even and odd bonds by dt each."""
    assert math.isclose(sum(t for t, b in TEBD4_SEQ if b == EVEN), DT)
    assert math.isclose(sum(t for t, b in TEBD4_SEQ if b == ODD), DT)
