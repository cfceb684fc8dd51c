import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sys, time
import numpy as np
import tevo_b200 as t
chi = int(sys.argv[1]); N = int(sys.argv[2]) if len(sys.argv) > 2 else 12
ctx = t.default_ctx()
t0 = time.time()
psi = t.MPS.random(N, chi, seed=1)
ctx.synchronize()
print("random+canonicalise", time.time() - t0, psi.bond_dims())
H = t.xxz_bondop(N, 1.0)
Ustart, Us, Uend = t.time_evo_gates(0.05, t.gates(H), t.TEBD4())
kw = dict(maxdim=chi, mindim=chi)
rev = False
for U in Ustart:
    t.apply_gates(psi, U, reversed=rev, **kw); rev = not rev
ctx.synchronize()
t.profile_enable(True)
for it in range(2):
    n0 = ctx.launch_count()
    t0 = time.time()
    ng = 0
    for U in Us:
        t.apply_gates(psi, U, reversed=rev, **kw); rev = not rev; ng += len(U)
    ctx.synchronize()
    dt = time.time() - t0
    prof = t.profile_read()
    tot = sum(v["ms"] for v in prof.values())
    print({k: (round(v["ms"], 1), v["count"]) for k, v in prof.items()}, "sum", round(tot, 1))
    print("chi", chi, "step", dt, "s;", ng, "gates;", dt / ng * 1e3, "ms/gate; launches", ctx.launch_count() - n0)
print(abs(t.inner(psi, psi)))

import ctypes as C
from timeevomps_jl_b200._lib import load
out=(C.c_double*11)()
load().tevo_debug_counters.argtypes=[C.c_void_p, C.POINTER(C.c_double), C.c_int]
load().tevo_debug_counters(ctx.h, out, 11)
print("cholqr single/double/fail", list(out[:3]), "ratio hist [1,2),[2,4),[4,8),[8,16),[16,64),[64,1e3),[1e3,1e6),fail:", list(out[3:]))
