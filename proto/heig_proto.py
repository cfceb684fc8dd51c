"""NumPy prototype of the device Hermitian eigensolver (blocked Householder tridiagonalisation + Cuppen divide &
conquer with Gu-Eisenstat vectors + compact-WY back-transformation), structured phase by phase like the CUDA
kernels in csrc/heig.cu.  Development aid only (not imported by the product or the tests' checker)."""
import numpy as np

EPS = np.finfo(float).eps


def larfg(alpha, x):
    xnorm = np.linalg.norm(x)
    if xnorm == 0 and alpha.imag == 0:
        return alpha.real, 0.0, np.zeros_like(x)  # beta, tau, v-tail  (H = I)
    beta = -np.copysign(np.sqrt(abs(alpha) ** 2 + xnorm ** 2), alpha.real)
    tau = complex((beta - alpha.real) / beta, -alpha.imag / beta)
    return beta, tau, x / (alpha - beta)


def tridiagonalize(A, nb=8):
    """In place: returns d, e, taus; A's strictly-lower part holds the reflectors (explicit 1, zeros above)."""
    A = A.copy()
    n = A.shape[0]
    d = np.zeros(n)
    e = np.zeros(max(n - 1, 0))
    taus = np.zeros(max(n - 1, 0), complex)
    p0 = 0
    while p0 < n - 1:
        pw = min(nb, n - 1 - p0)
        W = np.zeros((n, pw), complex)
        for i in range(pw):
            c = p0 + i
            Vp = A[:, p0:p0 + i]
            a = A[c:, c].copy()
            if i > 0:
                a -= Vp[c:, :] @ W[c, :i].conj() + W[c:, :i] @ Vp[c, :].conj()
            d[c] = a[0].real
            beta, tau, vt = larfg(a[1], a[2:])
            e[c] = beta
            taus[c] = tau
            v = np.concatenate([[1.0], vt])
            # store reflector
            A[:c + 1, c] = 0
            A[c + 1:, c] = v
            w = A[c + 1:, c + 1:] @ v
            if i > 0:
                w -= Vp[c + 1:, :] @ (W[c + 1:, :i].conj().T @ v)
                w -= W[c + 1:, :i] @ (Vp[c + 1:, :].conj().T @ v)
            w *= tau
            alpha2 = -0.5 * tau * np.vdot(w, v)
            w += alpha2 * v
            W[c + 1:, i] = w
        q0 = p0 + pw
        V = A[q0:, p0:q0]
        Wt = W[q0:, :]
        A[q0:, q0:] -= V @ Wt.conj().T + Wt @ V.conj().T
        p0 = q0
    d[n - 1] = A[n - 1, n - 1].real
    A[:, n - 1] = 0
    return A, d, e, taus


def apply_q(V, taus, Z, nb=8):
    """U = H_0 H_1 ... H_{n-2} Z using compact WY blocks (applied last block first)."""
    n = V.shape[0]
    U = Z.astype(complex).copy()
    starts = list(range(0, n - 1, nb))
    for p0 in reversed(starts):
        pw = min(nb, n - 1 - p0)
        Vp = V[:, p0:p0 + pw]
        S = Vp.conj().T @ Vp
        T = np.zeros((pw, pw), complex)
        for j in range(pw):
            T[j, j] = taus[p0 + j]
            if j > 0:
                T[:j, j] = -taus[p0 + j] * (T[:j, :j] @ S[:j, j])
        X = Vp.conj().T @ U
        U -= Vp @ (T @ X)
    return U


# ------------------------------------------------------------------------------------------------
def secular_root(j, K, dd, z2, rho):
    """Root j of 1 + rho*sum z2_i/(dd_i - lam) in (dd_j, dd_{j+1}); returns (origin index, mu)."""
    def f(org, mu):
        return 1.0 + rho * np.sum(z2 / ((dd - dd[org]) - mu))
    if j < K - 1:
        delta = dd[j + 1] - dd[j]
        fm = f(j, 0.5 * delta)
        if fm > 0:
            org, lo, hi, sgn = j, 0.0, 0.5 * delta, 1.0     # mu in (0, delta/2], f increasing in mu
        elif fm < 0:
            org, lo, hi, sgn = j + 1, 0.0, 0.5 * delta, -1.0  # mu = -x, x in (0, delta/2]
        else:
            return j, 0.5 * delta
    else:
        org, lo, hi, sgn = j, 0.0, rho * np.sum(z2), 1.0
    # find x in (0, hi] with g(x) = f(org, sgn*x) = 0; g is increasing in x for sgn=+1, decreasing for sgn=-1
    def g(x):
        return sgn * f(org, sgn * x)     # increasing in x in both cases, g(0+) = -inf
    x = hi
    if g(x) < 0 and j == K - 1:
        return org, hi
    # geometric descent to bracket
    xlo = 0.0
    for _ in range(80):
        xt = x / 16.0
        if xt == 0.0 or g(xt) <= 0:
            xlo = xt
            break
        x = xt
    xhi = x
    for _ in range(120):
        xm = 0.5 * (xlo + xhi) if (xlo > 0 and xhi / xlo < 4) or xlo == 0 else np.sqrt(xlo * xhi)
        if xm <= xlo or xm >= xhi:
            break
        if g(xm) <= 0:
            xlo = xm
        else:
            xhi = xm
    return org, sgn * 0.5 * (xlo + xhi)


def merge(D1, Q1, D2, Q2, e):
    """One Cuppen merge.  D1,D2 ascending; returns ascending D and Q."""
    n1, n2 = len(D1), len(D2)
    N = n1 + n2
    rho = abs(e)
    s = 1.0 if e >= 0 else -1.0
    Q = np.zeros((N, N))
    Q[:n1, :n1] = Q1
    Q[n1:, n1:] = Q2
    D = np.concatenate([D1, D2])
    z = np.concatenate([Q1[-1, :], s * Q2[0, :]]) / np.sqrt(2.0)
    rho = 2.0 * rho
    perm = np.argsort(D, kind="stable")
    D, z, Q = D[perm], z[perm], Q[:, perm]
    if rho == 0.0:
        return D, Q
    tol = 8.0 * EPS * max(np.max(np.abs(D)), np.max(np.abs(z)))
    defl = np.zeros(N, bool)
    if rho * np.max(np.abs(z)) <= tol:
        return D, Q
    prev = -1
    for j in range(N):
        if rho * abs(z[j]) <= tol:
            defl[j] = True
            continue
        if prev >= 0:
            sn, cs = z[prev], z[j]
            tau = np.hypot(cs, sn)
            t = D[j] - D[prev]
            cs, sn = cs / tau, -sn / tau
            if abs(t * cs * sn) <= tol:
                z[j] = tau
                z[prev] = 0.0
                qp, qj = Q[:, prev].copy(), Q[:, j].copy()
                Q[:, prev] = cs * qp + sn * qj
                Q[:, j] = -sn * qp + cs * qj
                tt = D[prev] * cs * cs + D[j] * sn * sn
                D[j] = D[prev] * sn * sn + D[j] * cs * cs
                D[prev] = tt
                defl[prev] = True
                # keep deflated values sorted is not required here (final sort below)
        prev = j
    idx = np.where(~defl)[0]
    K = len(idx)
    dd, zz = D[idx], z[idx]
    z2 = zz * zz
    org = np.zeros(K, int)
    mu = np.zeros(K)
    for j in range(K):
        org[j], mu[j] = secular_root(j, K, dd, z2, rho)
    lam = dd[org] + mu
    # Gu-Eisenstat: zhat_i^2 = prod_j (lam_j - d_i) / prod_{j != i} (d_j - d_i) / rho
    # delta[i, j] = d_i - lam_j computed from the stored offsets
    delta = (dd[:, None] - dd[org][None, :]) - mu[None, :]
    zhat = np.zeros(K)
    for i in range(K):
        p = -delta[i, i]  # lam_i - d_i
        for j in range(K):
            if j != i:
                p *= (-delta[i, j]) / (dd[j] - dd[i])
        zhat[i] = np.copysign(np.sqrt(abs(p) / rho), zz[i])
    Umat = zhat[:, None] / delta
    Umat /= np.linalg.norm(Umat, axis=0)[None, :]
    Qn = Q[:, idx] @ Umat
    Dall = np.concatenate([lam, D[defl]])
    Qall = np.concatenate([Qn, Q[:, defl]], axis=1)
    p2 = np.argsort(Dall, kind="stable")
    return Dall[p2], Qall[:, p2]


def stedc(d, e, leaf=4):
    n = len(d)
    d = d.copy()
    bounds = list(range(0, n, leaf)) + [n]
    for b in bounds[1:-1]:
        d[b - 1] -= abs(e[b - 1])
        d[b] -= abs(e[b - 1])
    blocks = []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        T = np.diag(d[lo:hi]) + np.diag(e[lo:hi - 1], 1) + np.diag(e[lo:hi - 1], -1)
        w, q = np.linalg.eigh(T)
        blocks.append((lo, hi, w, q))
    while len(blocks) > 1:
        nxt = []
        for k in range(0, len(blocks) - 1, 2):
            lo, mid, D1, Q1 = blocks[k]
            _, hi, D2, Q2 = blocks[k + 1]
            D, Q = merge(D1, Q1, D2, Q2, e[mid - 1])
            nxt.append((lo, hi, D, Q))
        if len(blocks) % 2:
            nxt.append(blocks[-1])
        blocks = nxt
    return blocks[0][2], blocks[0][3]


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 5, 16, 33, 70):
        X = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        A = X + X.conj().T
        V, d, e, taus = tridiagonalize(A, nb=8)
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        Q = apply_q(V, taus, np.eye(n), nb=8)
        print(n, "Q^H A Q - T", np.abs(Q.conj().T @ A @ Q - T).max(), "unitary", np.abs(Q.conj().T @ Q - np.eye(n)).max())
        if n >= 2:
            w, Z = stedc(d, e, leaf=4)
            print("   stedc eig err", np.abs(w - np.linalg.eigvalsh(T)).max(), "orth", np.abs(Z.T @ Z - np.eye(n)).max(),
                  "resid", np.abs(T @ Z - Z * w).max())
    # hard cases: clustered / rank deficient / graded
    for name, d, e in (("zeros", np.zeros(40), np.zeros(39)), ("wilk", np.abs(np.arange(-20, 21)).astype(float), np.ones(40)),
                       ("graded", 10.0 ** -np.arange(30), 10.0 ** -(np.arange(29) + 0.5)),
                       ("glued", np.tile(np.abs(np.arange(-5, 6)).astype(float), 4), np.concatenate([np.ones(10), [1e-10]] * 4)[:-1])):
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        w, Z = stedc(d, e, leaf=4)
        n = len(d)
        print(name, "eig err", np.abs(w - np.linalg.eigvalsh(T)).max(), "orth", np.abs(Z.T @ Z - np.eye(n)).max(),
              "resid", np.abs(T @ Z - Z * w).max())
    # PSD low-rank Gram
    X = rng.standard_normal((60, 7)) + 1j * rng.standard_normal((60, 7))
    A = X @ X.conj().T
    V, d, e, taus = tridiagonalize(A, nb=8)
    w, Z = stedc(d, e, leaf=8)
    U = apply_q(V, taus, Z, nb=8)
    print("lowrank: resid", np.abs(A @ U - U * w).max() / np.abs(w).max(), "orth", np.abs(U.conj().T @ U - np.eye(60)).max())
