// In-shared-memory Cholesky factor and triangular inverse of one <= 64 x 64 block, right-looking (outer-product)
// form: every step is a pivot + an embarrassingly parallel rank-1 update, so there are no dot-product chains.
// All functions are called by every thread of the CTA (any blockDim.x, 1024 intended).
#pragma once
#include "common.cuh"

namespace tevo {

constexpr int TRB = 64;

// a[p][i] = A(i,p) (column p contiguous over rows i), Hermitian positive definite, lower part referenced.
// On return a holds L (lower).  Returns false (uniformly) on a non-positive pivot.  pmin/pmax: pivot extremes.
__device__ inline bool chol_lower_smem(cplx (*a)[TRB], int nb, double& pmin, double& pmax) {
  pmin = 1e300;
  pmax = 0.0;
  for (int k = 0; k < nb; ++k) {
    const double dkk = a[k][k].x;  // finalised by the trailing update of step k-1 (barrier below)
    if (!(dkk > 0.0) || !isfinite(dkk)) return false;
    const double l = sqrt(dkk), inv = 1.0 / l;
    pmin = fmin(pmin, l);
    pmax = fmax(pmax, l);
    __syncthreads();
    for (int i = k + threadIdx.x; i < nb; i += blockDim.x) a[k][i] = (i == k) ? mk(l, 0.0) : cscale(inv, a[k][i]);
    __syncthreads();
    const int rem = nb - k - 1;
    for (int e = threadIdx.x; e < rem * rem; e += blockDim.x) {
      const int i = k + 1 + e % rem, j = k + 1 + e / rem;
      if (j <= i) a[j][i] = csub(a[j][i], cmul(a[k][i], cconj(a[k][j])));
    }
    __syncthreads();
  }
  return true;
}

// a[p][i] = L(i,p) lower triangular (entries with i < p ignored); x[p][c] <- inv(L)(p,c), zero above the diagonal.
__device__ inline void trinv_lower_smem(cplx (*a)[TRB], cplx (*x)[TRB], int nb) {
  for (int e = threadIdx.x; e < TRB * TRB; e += blockDim.x) {
    const int c = e % TRB, p = e / TRB;
    x[p][c] = mk((p == c && p < nb) ? 1.0 : 0.0, 0.0);
  }
  __syncthreads();
  for (int k = 0; k < nb; ++k) {
    const cplx d = a[k][k];
    const double dn = 1.0 / cabs2(d);
    const cplx dinv = mk(d.x * dn, -d.y * dn);
    for (int c = threadIdx.x; c <= k; c += blockDim.x) x[k][c] = cmul(dinv, x[k][c]);
    __syncthreads();
    const int rem = nb - k - 1, w = k + 1;
    for (int e = threadIdx.x; e < rem * w; e += blockDim.x) {
      const int c = e % w, i = k + 1 + e / w;
      x[i][c] = csub(x[i][c], cmul(a[k][i], x[k][c]));
    }
    __syncthreads();
  }
}

// Cholesky factor and its inverse in one right-looking sweep with a single barrier per column.  The pivot scaling is
// deferred: column k of `a` and row k of inv(L) stay unscaled (they are final once step k has read them) and the
// updates divide by the pivot instead; sinv[k] = 1/L(k,k) is applied by the caller:
//   L(i,k) = a[k][i] * sinv[k] (i >= k),   inv(L)(i,c) = xt[c][i] * sinv[i] (c <= i).
// a[p][i] = A(i,p), Hermitian positive definite, lower part referenced; xt is the inverse stored by columns (thread
// i works on row i of both matrices, so both are laid out with i contiguous: no bank conflicts).  Called by all
// threads (blockDim.x a multiple of 64).  Returns false (uniformly) on a non-positive pivot; pmin/pmax: extremes of
// L(k,k).
__device__ inline bool chol_inv_lower_smem(cplx (*a)[TRB], cplx (*xt)[TRB], double* sinv, int nb, double& pmin,
                                           double& pmax) {
  pmin = 1e300;
  pmax = 0.0;
  for (int e = threadIdx.x; e < TRB * TRB; e += blockDim.x) {
    const int i = e % TRB, c = e / TRB;
    xt[c][i] = mk((i == c && i < nb) ? 1.0 : 0.0, 0.0);
  }
  __syncthreads();
  const int i = threadIdx.x % TRB, tj = threadIdx.x / TRB, nj = blockDim.x / TRB;
  for (int k = 0; k < nb; ++k) {
    const double dkk = a[k][k].x;  // final: the update of step k-1 finished at the barrier below
    if (!(dkk > 0.0) || !isfinite(dkk)) return false;
    const double inv = rsqrt(dkk), l = dkk * inv, rd = inv * inv;
    pmin = fmin(pmin, l);
    pmax = fmax(pmax, l);
    if (threadIdx.x == 0) sinv[k] = inv;
    if (i > k && i < nb) {
      const cplx aki = cscale(rd, a[k][i]);
#pragma unroll 4
      for (int j = k + 1 + tj; j <= i; j += nj) a[j][i] = csub(a[j][i], cmul(aki, cconj(a[k][j])));
#pragma unroll 4
      for (int c = tj; c <= k; c += nj) xt[c][i] = csub(xt[c][i], cmul(aki, xt[c][k]));
    }
    __syncthreads();
  }
  return true;
}

}  // namespace tevo
