"""Error of the three-multiplication complex product (csrc/zgemm.cu, small-tile kernels) against the four-multiplication
one, both in float64 with an 80-bit reference (NumPy on the CPU; no GPU needed):
  (1) dense random 256 x 2048 x 256: max error of Re / Im relative to max row-norm x max column-norm;
  (2) R = U^+ theta with singular values from 1 to 1e-5: relative accuracy of the row weights |R_i|^2 (the refined
      Schmidt weights of the split) from 1 down to 1e-10.
Output of the run recorded in DESIGN.md section 7.3:
  4M max err re 4.99e-17 im 4.88e-17 ; 3M max err re 4.99e-17 im 9.01e-17
  4M row weights: max rel dev 9.2e-13 (at weight 1.7e-10) ; 3M: 1.26e-12 (at weight 1.2e-10)"""
import numpy as np

rng = np.random.default_rng(0)
L = np.longdouble


def prod4(Ar, Ai, Br, Bi):
    return Ar @ Br - Ai @ Bi, Ar @ Bi + Ai @ Br


def prod3(Ar, Ai, Br, Bi):
    t1, t2, t3 = Ar @ Br, Ai @ Bi, (Ar + Ai) @ (Br + Bi)
    return t1 - t2, t3 - t1 - t2


M, N, K = 256, 256, 2048
A = rng.standard_normal((M, K)) + 1j * rng.standard_normal((M, K))
B = rng.standard_normal((K, N)) + 1j * rng.standard_normal((K, N))
Ar, Ai, Br, Bi = A.real, A.imag, B.real, B.imag
ref_re = Ar.astype(L) @ Br.astype(L) - Ai.astype(L) @ Bi.astype(L)
ref_im = Ar.astype(L) @ Bi.astype(L) + Ai.astype(L) @ Br.astype(L)
scale = np.linalg.norm(A, axis=1).max() * np.linalg.norm(B, axis=0).max()
for name, f in (("4M", prod4), ("3M", prod3)):
    re, im = f(Ar, Ai, Br, Bi)
    print(name, "max err re %.2e im %.2e (relative to max row-norm x col-norm)"
          % (float(np.abs(re - ref_re).max()) / scale, float(np.abs(im - ref_im).max()) / scale))

U, _ = np.linalg.qr(rng.standard_normal((K, K)) + 1j * rng.standard_normal((K, K)))
V, _ = np.linalg.qr(rng.standard_normal((K, K)) + 1j * rng.standard_normal((K, K)))
theta = (U * np.logspace(0, -5, K)) @ V
Uh = U.conj().T
w_ref = np.array((np.abs(Uh.astype(np.clongdouble) @ theta.astype(np.clongdouble)) ** 2).sum(axis=1), dtype=np.float64)
for name, f in (("4M", prod4), ("3M", prod3)):
    re, im = f(Uh.real, Uh.imag, theta.real, theta.imag)
    w = (re ** 2 + im ** 2).sum(axis=1)
    rel = np.abs(w - w_ref) / w_ref
    print(name, "row weights |R_i|^2 from 1 to 1e-10: max rel dev %.2e (at weight %.1e)" % (rel.max(), w_ref[rel.argmax()]))
