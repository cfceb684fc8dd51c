// CholeskyQR2 for the centre moves of orthogonalize! (src/bondgate.jl:48): A (m x n, m >= n) = Q C with Q^H Q = 1.
// Two passes of  G = A^H A (DMMA GEMM),  G = L L^H (blocked Cholesky, 64-wide diagonal blocks factored and inverted in
// shared memory by one CTA, panels and trailing updates by GEMM),  Q = A L^-H (block forward substitution by GEMM).
// Valid while cond(A) < ~1e6 (then the second pass restores orthogonality to machine precision); a non-positive
// pivot or a large diagonal ratio makes factor() return false and the caller uses the eigen-decomposition route.
#include "cholqr.h"

#include <algorithm>
#include <cmath>
#include <vector>

#include "kernels.h"
#include "prof.h"
#include "zgemm.h"

namespace tevo {
namespace {

constexpr int NB = 64;

// Cholesky of one nb x nb (nb <= 64) diagonal block of G at (j,j), left-looking, one thread per row: writes L_jj
// into Lm (lower part) and its inverse into Xm (lower part) at the same position (both column-major, ld),
// and the extreme pivots into stats[0..1] (stats[0] < 0: breakdown).
__global__ void __launch_bounds__(64) potrf_block_kernel(int nb, const cplx* __restrict__ G, cplx* __restrict__ Lm,
                                                         cplx* __restrict__ Xm, long ld, double* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char smraw[];
  cplx(*a)[NB] = reinterpret_cast<cplx(*)[NB]>(smraw);  // a[p][i] = L(i,p)   (column p contiguous over rows i)
  cplx(*x)[NB] = a + NB;                                // x[p][c] = inv(L)(p,c)
  __shared__ double s_piv;
  __shared__ int s_fail;
  const int i = threadIdx.x;
  for (int p = 0; p < nb; ++p) a[p][i] = (i < nb) ? G[(long)p * ld + i] : mk(0, 0);
  if (i == 0) s_fail = 0;
  __syncthreads();
  double pmin = 1e300, pmax = 0.0;
  for (int j = 0; j < nb; ++j) {
    cplx sres = mk(0, 0);
    if (i >= j && i < nb) {
      sres = a[j][i];
      for (int p = 0; p < j; ++p) sres = csub(sres, cmul(a[p][i], cconj(a[p][j])));
    }
    if (i == j) {
      const double dkk = sres.x;
      if (!(dkk > 0.0) || !isfinite(dkk)) {
        s_fail = 1;
        s_piv = 1.0;
      } else {
        s_piv = sqrt(dkk);
      }
    }
    __syncthreads();
    if (s_fail) break;
    const double l = s_piv;
    pmin = fmin(pmin, l);
    pmax = fmax(pmax, l);
    if (i >= j && i < nb) a[j][i] = (i == j) ? mk(l, 0.0) : cscale(1.0 / l, sres);
    __syncthreads();
  }
  if (s_fail) {
    if (i == 0) {
      stats[0] = -1.0;
      stats[1] = 0.0;
    }
    return;
  }
  // inverse by forward substitution, thread c owns column c
  const int c = i;
  if (c < nb) {
    for (int r = 0; r < nb; ++r) {
      if (r < c) {
        x[r][c] = mk(0, 0);
        continue;
      }
      cplx sres = mk(r == c ? 1.0 : 0.0, 0.0);
      for (int p = c; p < r; ++p) sres = csub(sres, cmul(a[p][r], x[p][c]));
      x[r][c] = cscale(1.0 / a[r][r].x, sres);
    }
  }
  __syncthreads();
  for (int p = 0; p < nb; ++p) {
    if (i < nb) {
      Lm[(long)p * ld + i] = (i >= p) ? a[p][i] : mk(0, 0);
      Xm[(long)p * ld + i] = (i >= p) ? x[i][p] : mk(0, 0);
    }
  }
  if (i == 0) {
    stats[0] = pmin;
    stats[1] = pmax;
  }
}

}  // namespace

void CholQR::init() {
  TEVO_CUDA_CHECK(cudaMalloc(&d_stats_, sizeof(double) * 2 * 128));
  TEVO_CUDA_CHECK(cudaMallocHost(&h_stats_, sizeof(double) * 2 * 128));
  TEVO_CUDA_CHECK(cudaFuncSetAttribute(potrf_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(2 * sizeof(cplx) * NB * NB)));
}

void CholQR::destroy() {
  auto fr = [](void* p) {
    if (p) cudaFree(p);
  };
  fr(G_); fr(L1_); fr(L2_); fr(Linv_); fr(T_); fr(Q1_); fr(d_stats_);
  if (h_stats_) cudaFreeHost(h_stats_);
  G_ = L1_ = L2_ = Linv_ = T_ = Q1_ = nullptr;
  d_stats_ = h_stats_ = nullptr;
  ncap_ = mcap_ = 0;
}

void CholQR::ensure(int m, int n) {
  if (n > ncap_) {
    auto fr = [](void* p) {
      if (p) cudaFree(p);
    };
    fr(G_); fr(L1_); fr(L2_); fr(Linv_);
    const size_t nn = (size_t)n * n;
    TEVO_CUDA_CHECK(cudaMalloc(&G_, sizeof(cplx) * nn));
    TEVO_CUDA_CHECK(cudaMalloc(&L1_, sizeof(cplx) * nn));
    TEVO_CUDA_CHECK(cudaMalloc(&L2_, sizeof(cplx) * nn));
    TEVO_CUDA_CHECK(cudaMalloc(&Linv_, sizeof(cplx) * nn));
    ncap_ = n;
    mcap_ = 0;
  }
  if ((size_t)m * n > mcap_) {
    if (T_) cudaFree(T_);
    if (Q1_) cudaFree(Q1_);
    TEVO_CUDA_CHECK(cudaMalloc(&T_, sizeof(cplx) * std::max((size_t)m * NB, (size_t)n * n)));
    TEVO_CUDA_CHECK(cudaMalloc(&Q1_, sizeof(cplx) * (size_t)m * n));
    mcap_ = (size_t)m * n;
  }
}

// one pass: Q = A L^-H with A^H A = L L^H; pivot extremes per diagonal block go to d_stats_
void CholQR::pass(cudaStream_t st, const cplx* A, int m, int n, cplx* L, cplx* Q) {
  const cplx ONE = mk(1, 0), ZERO = mk(0, 0), MONE = mk(-1, 0);
  const int nblk = (n + NB - 1) / NB;
  cplx* X = Linv_;  // n x n: inverse of L, assembled below
  TEVO_CUDA_CHECK(cudaMemsetAsync(L, 0, sizeof(cplx) * (size_t)n * n, st));
  TEVO_CUDA_CHECK(cudaMemsetAsync(X, 0, sizeof(cplx) * (size_t)n * n, st));
  zgemm(st, OP_C, OP_N, n, n, m, ONE, A, m, A, m, ZERO, G_, n);
  for (int jb = 0; jb < nblk; ++jb) {
    const int j = jb * NB, nb = std::min(NB, n - j), rest = n - j - nb;
    potrf_block_kernel<<<1, 64, 2 * sizeof(cplx) * NB * NB, st>>>(nb, G_ + (long)j * n + j, L + (long)j * n + j, X + (long)j * n + j, n,
                                         d_stats_ + 2 * jb);
    TEVO_LAUNCHED();
    if (rest > 0) {
      // panel: L(j+nb:, j:j+nb) = G(j+nb:, j:j+nb) * inv(L_jj)^H ; trailing: G(j+nb:, j+nb:) -= panel * panel^H
      zgemm(st, OP_N, OP_C, rest, nb, nb, ONE, G_ + (long)j * n + j + nb, n, X + (long)j * n + j, n, ZERO,
            L + (long)j * n + j + nb, n);
      zgemm(st, OP_N, OP_C, rest, rest, nb, MONE, L + (long)j * n + j + nb, n, L + (long)j * n + j + nb, n, ONE,
            G_ + (long)(j + nb) * n + j + nb, n);
    }
  }
  // inverse of L by pairwise merging of diagonal ranges: inv([A 0; B C]) = [A^-1 0; -C^-1 B A^-1, C^-1]
  std::vector<int> bounds;
  for (int b = 0; b < n; b += NB) bounds.push_back(b);
  bounds.push_back(n);
  while (bounds.size() > 2) {
    std::vector<int> nbnd;
    size_t k = 0;
    for (; k + 2 < bounds.size(); k += 2) {
      const int lo = bounds[k], mid = bounds[k + 1], hi = bounds[k + 2];
      const int s1 = mid - lo, s2 = hi - mid;
      // T = B * A^-1 (s2 x s1) ; X21 = -C^-1 * T
      zgemm(st, OP_N, OP_N, s2, s1, s1, ONE, L + (long)lo * n + mid, n, X + (long)lo * n + lo, n, ZERO, T_, s2);
      zgemm(st, OP_N, OP_N, s2, s1, s2, MONE, X + (long)mid * n + mid, n, T_, s2, ZERO, X + (long)lo * n + mid, n);
      nbnd.push_back(lo);
    }
    for (; k < bounds.size(); ++k) nbnd.push_back(bounds[k]);
    bounds = nbnd;
  }
  // Q = A * inv(L)^H
  zgemm(st, OP_N, OP_C, m, n, n, ONE, A, m, X, n, ZERO, Q, m);
}

bool CholQR::factor(cudaStream_t st, const cplx* A, int m, int n, cplx* Q, cplx* C) {
  if (m < n || n < 1) return false;
  ensure(m, n);
  ProfScope ps(st, P_HOP_GEMM);
  const int nblk = (n + NB - 1) / NB;
  if (nblk > 64) return false;
  auto check = [&](double maxratio) {
    TEVO_CUDA_CHECK(cudaMemcpyAsync(h_stats_, d_stats_, sizeof(double) * 2 * nblk, cudaMemcpyDeviceToHost, st));
    TEVO_CUDA_CHECK(cudaStreamSynchronize(st));
    double mn = 1e300, mx = 0.0;
    for (int b = 0; b < nblk; ++b) {
      if (!(h_stats_[2 * b] > 0.0)) return false;
      mn = std::min(mn, h_stats_[2 * b]);
      mx = std::max(mx, h_stats_[2 * b + 1]);
    }
    return std::isfinite(mx) && mx <= maxratio * mn;
  };
  pass(st, A, m, n, L1_, Q1_);
  if (!check(1e6)) return false;
  pass(st, Q1_, m, n, L2_, Q);
  if (!check(1.0 + 1e-3)) return false;  // second-pass factor must be ~identity
  // C = L2^H L1^H = (L1 L2)^H
  const cplx ONE = mk(1, 0), ZERO = mk(0, 0);
  zgemm(st, OP_N, OP_N, n, n, n, ONE, L1_, n, L2_, n, ZERO, G_, n);
  conj_transpose(st, n, n, G_, n, C, n, true);
  return true;
}

}  // namespace tevo
