"""CPU emulation of csrc/tail.cu, CTA by CTA between the cluster barriers: checks the data flow (cyclic row ownership,
gathered column / y vectors, lagging rank-2 update) against Q^H A Q = tridiag(d, e).  Development aid."""
import numpy as np
def run(Aglob, t0, C):
    """Emulates tail.cu CTA by CTA (serially between the cluster barriers)."""
    A = Aglob.copy(); n = A.shape[0]; m = n - t0
    ldr = (m + 1) & ~1; slots = (m + C - 1) // C
    d = np.zeros(n); e = np.zeros(n); taus = np.zeros(n, complex)
    cta = [dict(Ar=np.zeros((slots, ldr), complex), xs=np.zeros(ldr, complex), ys=np.zeros(ldr, complex),
                vb=np.zeros((2, ldr), complex), wb=np.zeros((2, ldr), complex)) for _ in range(C)]
    for rank in range(C):
        T = cta[rank]
        for s in range(slots):
            i = s * C + rank
            if i >= m: break
            col = A[t0:, t0 + i]           # column i of the trailing matrix (global col-major)
            T['Ar'][s, :m] = np.conj(col[:m])
        T['xs'][:m] = A[t0:, t0]
    for c in range(m - 1):
        scal_cache = {}
        for rank in range(C):              # phase (A) + (B)
            T = cta[rank]; xs = T['xs']; vs = T['vb'][c & 1]; vp = T['vb'][(c + 1) & 1]; wp = T['wb'][(c + 1) & 1]
            xn2 = np.sum(np.abs(xs[c + 2:m]) ** 2); alpha = xs[c + 1]; s2 = abs(alpha) ** 2 + xn2
            if (xn2 == 0 and alpha.imag == 0) or not (s2 > 1e-290):
                beta = alpha.real; tau = 0j; scal = 0j
            else:
                beta = -np.copysign(np.sqrt(s2), alpha.real); tau = complex((beta - alpha.real) / beta, -alpha.imag / beta)
                scal = 1.0 / (alpha - beta)
            if rank == 0:
                d[t0 + c] = xs[c].real; e[t0 + c] = beta; taus[t0 + c] = tau
            for i in range(m):
                vs[i] = 0 if i <= c else (1 if i == c + 1 else xs[i] * scal)
            for i in range(rank, m, C):    # reflector store (any distribution covers all i)
                A[t0 + i, t0 + c] = vs[i]
            T['tau'] = tau
            for s in range(slots):
                i = s * C + rank
                if i >= m or i <= c: continue
                row = T['Ar'][s]
                acc = 0j
                for j in range(c + 1, m):
                    aij = row[j] - vp[i] * np.conj(wp[j]) - wp[i] * np.conj(vp[j])
                    row[j] = aij; acc += aij * vs[j]
                for p in range(C): cta[p]['ys'][i] = acc
        for rank in range(C):              # phase (C)
            T = cta[rank]; vs = T['vb'][c & 1]; ws = T['wb'][c & 1]; ys = T['ys']; tau = T['tau']
            yv = np.sum(np.conj(ys[c + 1:m]) * vs[c + 1:m])
            aw = -0.5 * tau * (np.conj(tau) * yv)
            for i in range(m):
                ws[i] = 0 if i <= c else tau * ys[i] + aw * vs[i]
        newx = [dict() for _ in range(C)]
        for rank in range(C):
            T = cta[rank]; vs = T['vb'][c & 1]; ws = T['wb'][c & 1]; wc1 = ws[c + 1]
            for s in range(slots):
                i = s * C + rank
                if i >= m or i < c + 1: continue
                x = T['Ar'][s, c + 1] - vs[i] * np.conj(wc1) - ws[i]
                for p in range(C): newx[p][i] = x
        for p in range(C):
            for i, x in newx[p].items(): cta[p]['xs'][i] = x
    A[n - 1, n - 1] = cta[0]['xs'][m - 1].real
    for k in range(m - 1):
        A[:t0, t0 + k] = 0
    return A, d, e, taus

rng = np.random.default_rng(3)
for n, t0, C in [(9, 0, 1), (13, 4, 2), (21, 5, 4), (40, 8, 16), (37, 0, 16), (6, 4, 8)]:
    X = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)); A0 = X + X.conj().T
    # pretend the leading t0 columns were already reduced: only the trailing block matters here
    A, d, e, taus = run(A0, t0, C)
    m = n - t0; B = A0[t0:, t0:]
    d[n - 1] = A[n - 1, n - 1].real
    Q = np.eye(m, dtype=complex)
    for c in range(m - 1):
        v = A[t0:, t0 + c].copy()
        Q = Q @ (np.eye(m) - taus[t0 + c] * np.outer(v, v.conj()))
    T = Q.conj().T @ B @ Q
    Tref = np.diag(d[t0:]) + np.diag(e[t0:n - 1], -1) + np.diag(e[t0:n - 1], 1)
    print(n, t0, C, np.abs(T - Tref).max(), np.abs(A[:t0, t0:n - 1]).max() if t0 else 0.0)
