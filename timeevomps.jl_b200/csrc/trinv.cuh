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

}  // namespace tevo
