"""Timing of the CholeskyQR diagonal-block kernels (csrc/cholqr.cu) through tevo_test_potrf_block: time per launch
over back-to-back launches and, for the default panel kernel, thread 0's clock64 phase sums."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import tevo_b200  # noqa: F401  (loads the library)
from test_gpu_kernels import _potrf_block

rng = np.random.default_rng(0)
names = ["load", "publish+wait", "panel: load rows", "panel: column loop", "write back + barrier + own tiles", "L out + diag inv", "inverse merge", "total"]
for nb in (64, 40):
    B = rng.standard_normal((150, nb)) + 1j * rng.standard_normal((150, nb))
    G = B.conj().T @ B
    for variant, nm in ((0, "panel"), (1, "register-tiled"), (2, "shared-memory")):
        L, X, stats, us, cyc = _potrf_block(G, variant, reps=200, cycles=(variant == 0))
        err = np.abs(L - np.linalg.cholesky(G)).max()
        print("nb=%d %-15s %.1f us per launch (back to back), max |L - lapack| %.1e" % (nb, nm, us, err))
        if variant == 0:
            print("   cycles of thread 0: " + ", ".join("%s %d" % (n, c) for n, c in zip(names, cyc)))
