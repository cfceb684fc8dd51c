import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sys, time
import numpy as np
import tevo_b200 as t
N, chi = int(sys.argv[1]), int(sys.argv[2])
ctx = t.default_ctx()
psi = t.MPS.random(N, chi, seed=1)
H = t.xxz_mpo(N, 1.0, 0.3)
t.profile_enable(True)
for it in range(2):
    st = {}
    ctx.synchronize(); t0 = time.time()
    t.tdvp(psi, H, 0.05, 0.05, maxdim=chi, mindim=chi, exp_tol=1e-12, krylovdim=30, stats=st)
    ctx.synchronize(); dt = time.time() - t0
    prof = t.profile_read()
    print("N", N, "chi", chi, "tdvp step", round(dt, 3), "s", st, {k: (round(v["ms"], 1), v["count"]) for k, v in prof.items()})
print(abs(t.inner(psi, psi)))
