"""timeevomps.jl_b200 - B200-native two-site MPS update (TEBD / two-site TDVP) behind the TimeEvoMPS.jl API.

Host-side mirror of the reference's public interface (names follow src/*.jl; `!` dropped, indices 0-based):
    tebd, TEBD2, TEBD2Sweep, TEBD4, apply_gate, apply_gates, tdvp,
    BondOperator, BondGate, SiteTerm, BondTerm, add, bondgate, gates, siteterms, bondterms, coeff, exp, measure,
    TEvoCallback, NoTEvoCallback, LocalMeasurementCallback, SpecCallback, measurement_ts, measurements,
    tfi_mpo, tfi_bondop (test helpers of src/testutils.jl)
All arithmetic runs in libtevo.so (hand-written sm_100a CUDA, include/tevo.h); there is no CPU fallback."""
from ._lib import Context, TevoError, default_ctx, load, LIB_PATH, profile_enable, profile_read
from .mps import MPS, MPO, Spectrum, inner, inner_mpo, op, tfi_mpo, xxz_mpo
from .bondop import (BondOperator, BondGate, SiteTerm, BondTerm, GateList, add, bond, bondgate, gates, siteterms,
                     bondterms, coeff, exp, expm, measure, tfi_bondop, xxz_bondop)
from .callback import (TEvoCallback, NoTEvoCallback, LocalMeasurementCallback, SpecCallback, measurement_ts,
                       measurements)
from .tebd import (TEBDalg, TEBD2, TEBD2Sweep, TEBD4, apply_gate, apply_gates, time_evo_gates, tebd,
                   apply_gates_parallel, tebd_parallel)
from .tdvp import TDVP2, ProjMPO, sweepnext, tdvp

__all__ = [n for n in dir() if not n.startswith("_")]
