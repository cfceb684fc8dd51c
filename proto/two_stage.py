"""NumPy prototype of the two-stage Hermitian tridiagonalisation (the data flow of csrc/sbr.cu, checked against
numpy.linalg.eigh before any CUDA was written):

  stage 1  dense -> band (bandwidth b): blocked Householder QR of the sub-band panel, two-sided rank-2b update
  stage 2  band -> real tridiagonal: bulge chasing, one sweep per column, reflectors of length <= b
  back     eigenvectors of A = Q1 Q2 Z, Z = eigenvectors of the tridiagonal matrix

    python proto/two_stage.py 96 8
"""
import sys

import numpy as np


def zlarfg(x):
    """LAPACK zlarfg: H = I - tau v v^H with v[0] = 1, H^H x = beta e1, beta real."""
    x = np.asarray(x, dtype=complex)
    alpha = x[0]
    xn2 = np.sum(np.abs(x[1:]) ** 2)
    v = np.zeros_like(x)
    v[0] = 1.0
    if xn2 == 0.0 and alpha.imag == 0.0:
        return v, 0.0, alpha.real
    beta = -np.copysign(np.sqrt(abs(alpha) ** 2 + xn2), alpha.real)
    tau = (beta - alpha) / beta
    v[1:] = x[1:] / (alpha - beta)
    return v, tau, beta


def stage1(A, b):
    """Returns (B band matrix in full storage, V with the reflectors in its columns (unit at row c+b), taus)."""
    n = A.shape[0]
    A = A.astype(complex).copy()
    V = np.zeros((n, n), dtype=complex)
    taus = np.zeros(n, dtype=complex)
    j = 0
    while j + b <= n - 2:        # rows R = j+b .. n-1 hold at least 2 entries
        R = slice(j + b, n)
        m = n - j - b
        bw = min(b, m - 1)       # reflectors of this block; the panel itself always has b columns (the last block may
        #                          have more columns than reflectors: those only receive Q^H from the left)
        P = A[R, j:j + b].copy()
        Vp = np.zeros((m, bw), dtype=complex)
        tp = np.zeros(bw, dtype=complex)
        for i in range(bw):
            v, tau, beta = zlarfg(P[i:, i])
            Vp[i:, i] = v
            tp[i] = tau
            P[i, i] = beta
            P[i + 1:, i] = 0
            if i + 1 < b:  # H^H applied to the remaining panel columns
                w = v.conj() @ P[i:, i + 1:]
                P[i:, i + 1:] -= np.conj(tau) * np.outer(v, w)
        # compact WY: Q = H_0 H_1 ... = I - V T V^H
        T = np.zeros((bw, bw), dtype=complex)
        for i in range(bw):
            T[i, i] = tp[i]
            if i > 0:
                T[:i, i] = -tp[i] * (T[:i, :i] @ (Vp[:, :i].conj().T @ Vp[:, i]))
        A[R, j:j + b] = P
        A[j:j + b, R] = P.conj().T
        # two-sided update of the trailing matrix  A22 <- Q^H A22 Q = A22 - V W^H - W V^H
        A22 = A[R, R]
        Y = A22 @ Vp
        YT = Y @ T
        W = YT - 0.5 * Vp @ (T.conj().T @ (Vp.conj().T @ YT))
        A[R, R] = A22 - Vp @ W.conj().T - W @ Vp.conj().T
        V[R, j:j + bw] = Vp
        taus[j:j + bw] = tp
        j += b
    return A, V, taus


def stage2(B, b):
    """Bulge chasing.  Returns d, e (real tridiagonal) and the reflector list [(row0, v, tau)] in application order."""
    n = B.shape[0]
    B = B.copy()
    refl = []
    for s in range(n - 1):
        r0, r1 = s + 1, min(s + b, n - 1)
        v, tau, beta = zlarfg(B[r0:r1 + 1, s])
        refl.append((r0, v, tau))
        B[r0:r1 + 1, s] = 0
        B[r0, s] = beta
        B[s, r0:r1 + 1] = B[r0:r1 + 1, s].conj()
        H = np.eye(r1 - r0 + 1, dtype=complex) - tau * np.outer(v, v.conj())
        while True:
            D = B[r0:r1 + 1, r0:r1 + 1]
            B[r0:r1 + 1, r0:r1 + 1] = H.conj().T @ D @ H
            q0, q1 = r1 + 1, min(r1 + b, n - 1)
            if q0 > n - 1:
                break
            E = B[q0:q1 + 1, r0:r1 + 1] @ H
            v2, tau2, beta2 = zlarfg(E[:, 0])
            H2 = np.eye(q1 - q0 + 1, dtype=complex) - tau2 * np.outer(v2, v2.conj())
            E = H2.conj().T @ E
            E[1:, 0] = 0
            E[0, 0] = beta2
            B[q0:q1 + 1, r0:r1 + 1] = E
            B[r0:r1 + 1, q0:q1 + 1] = E.conj().T
            refl.append((q0, v2, tau2))
            r0, r1, H = q0, q1, H2
    d = B.diagonal().real.copy()
    e = B.diagonal(-1).real.copy()
    off = np.abs(B - np.diag(B.diagonal()) - np.diag(B.diagonal(-1), -1) - np.diag(B.diagonal(1), 1)).max()
    return d, e, refl, off, np.abs(B.diagonal(-1).imag).max()


def back(Z, refl, V, taus, b):
    """X = Q1 Q2 Z."""
    X = Z.astype(complex).copy()
    for (r0, v, tau) in reversed(refl):      # Q2 = H_first ... H_last  ->  apply the last one first
        L = len(v)
        X[r0:r0 + L] -= tau * np.outer(v, v.conj() @ X[r0:r0 + L])
    n = V.shape[0]
    for c in range(n - 1, -1, -1):
        if taus[c] != 0:
            v = V[:, c]
            X -= taus[c] * np.outer(v, v.conj() @ X)
    return X


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    rng = np.random.default_rng(n)
    M = rng.standard_normal((n, n // 2)) + 1j * rng.standard_normal((n, n // 2))
    A = M @ M.conj().T
    Bm, V, taus = stage1(A, b)
    band_off = max(np.abs(np.tril(Bm, -b - 1)).max(), np.abs(np.triu(Bm, b + 1)).max())
    w_ref = np.linalg.eigvalsh(A)
    print("stage 1: outside band %.2e, eig err %.2e" % (band_off, np.abs(np.linalg.eigvalsh(Bm) - w_ref).max() / w_ref[-1]))
    d, e, refl, off, im = stage2(Bm, b)
    T = np.diag(d) + np.diag(e, -1) + np.diag(e, 1)
    w, Z = np.linalg.eigh(T)
    print("stage 2: off-tridiagonal %.2e, imag(e) %.2e, eig err %.2e, reflectors %d"
          % (off, im, np.abs(w - w_ref).max() / w_ref[-1], len(refl)))
    X = back(Z, refl, V, taus, b)
    print("back: resid %.2e, orth %.2e" % (np.abs(A @ X - X * w).max() / w_ref[-1],
                                            np.abs(X.conj().T @ X - np.eye(n)).max()))


if __name__ == "__main__":
    main()
