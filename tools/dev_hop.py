import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sys, time
import tevo_b200 as t
N, chi = 24, 1024
ctx = t.default_ctx()
psi = t.MPS.random(N, chi, seed=0)
psi.orthogonalize(10); ctx.synchronize()
t.profile_enable(True)
for it in range(3):
    t.profile_read()
    t0 = time.time()
    psi.orthogonalize(14); psi.orthogonalize(10)
    ctx.synchronize()
    dt = time.time() - t0
    prof = t.profile_read()
    print("8 bulk hops:", round(dt * 1e3, 1), "ms", {k: (round(v["ms"], 2), v["count"]) for k, v in prof.items()})
