import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sys, time
import tevo_b200 as t
N, chi = int(sys.argv[1]), int(sys.argv[2])
ctx = t.default_ctx()
t0 = time.time()
try:
    psi = t.MPS.random(N, chi, seed=0)
    ctx.synchronize()
    print(N, chi, "ok", round(time.time() - t0, 2), "s", abs(t.inner(psi, psi)))
except Exception as e:
    print(N, chi, "FAIL", e)
