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
#include "trinv.cuh"
#include "prof.h"
#include "zgemm.h"

namespace tevo {
namespace {

constexpr int NB = 64;

// Cholesky of one nb x nb (nb <= 64) diagonal block of G at (j,j), left-looking, 4 threads per row: writes L_jj into
// Lm (lower part) and its inverse into Xm (lower part) at the same position (both column-major, ld), and the
// extreme pivots into stats[0..1] (stats[0] < 0: breakdown).
__global__ void __launch_bounds__(256) potrf_block_kernel(int nb, const cplx* __restrict__ G, cplx* __restrict__ Lm,
                                                          cplx* __restrict__ Xm, long ld, double* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char smraw[];
  cplx(*a)[NB] = reinterpret_cast<cplx(*)[NB]>(smraw);  // a[p][i] = L(i,p)   (column p contiguous over rows i)
  cplx(*x)[NB] = a + NB;                                // x[p][c] = inv(L)(p,c)
  __shared__ double s_piv;
  __shared__ int s_fail;
  const int tid = threadIdx.x, i = tid >> 2, part = tid & 3;
  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e % NB, p = e / NB;
    a[p][r] = (r < nb && p < nb) ? G[(long)p * ld + r] : mk(r == p ? 1.0 : 0.0, 0.0);
  }
  if (tid == 0) s_fail = 0;
  __syncthreads();
  double pmin = 1e300, pmax = 0.0;
  for (int j = 0; j < nb; ++j) {
    cplx sres = mk(0, 0);
    if (i >= j && i < nb) {
      for (int p = part; p < j; p += 4) cfma(sres, a[p][i], cconj(a[p][j]));
    }
    sres.x += __shfl_xor_sync(0xffffffffu, sres.x, 1);
    sres.y += __shfl_xor_sync(0xffffffffu, sres.y, 1);
    sres.x += __shfl_xor_sync(0xffffffffu, sres.x, 2);
    sres.y += __shfl_xor_sync(0xffffffffu, sres.y, 2);
    if (i >= j && i < nb) sres = csub(a[j][i], sres);
    if (i == j && part == 0) {
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
    if (i >= j && i < nb && part == 0) a[j][i] = (i == j) ? mk(l, 0.0) : cscale(1.0 / l, sres);
    __syncthreads();
  }
  if (s_fail) {
    if (tid == 0) {
      stats[0] = -1.0;
      stats[1] = 0.0;
    }
    return;
  }
  trinv_lower_256(a, x, nb);
  for (int e = tid; e < nb * nb; e += 256) {
    const int r = e % nb, p = e / nb;
    Lm[(long)p * ld + r] = (r >= p) ? a[p][r] : mk(0, 0);
    Xm[(long)p * ld + r] = (r >= p) ? x[r][p] : mk(0, 0);
  }
  if (tid == 0) {
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
  {
    ProfScope ps(st, P_CHOL_GRAM);
    zgemm(st, OP_C, OP_N, n, n, m, ONE, A, m, A, m, ZERO, G_, n);
  }
  ProfScope* pf = new ProfScope(st, P_CHOL_FACTOR);
  for (int jb = 0; jb < nblk; ++jb) {
    const int j = jb * NB, nb = std::min(NB, n - j), rest = n - j - nb;
    potrf_block_kernel<<<1, 256, 2 * sizeof(cplx) * NB * NB, st>>>(nb, G_ + (long)j * n + j, L + (long)j * n + j, X + (long)j * n + j, n,
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
  delete pf;
  ProfScope* pi = new ProfScope(st, P_CHOL_INV);
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
  delete pi;
  // Q = A * inv(L)^H
  ProfScope pq(st, P_CHOL_Q);
  zgemm(st, OP_N, OP_C, m, n, n, ONE, A, m, X, n, ZERO, Q, m);
}

bool CholQR::factor(cudaStream_t st, const cplx* A, int m, int n, cplx* Q, cplx* C) {
  if (m < n || n < 1) return false;
  ensure(m, n);
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
    last_ratio_ = mx / mn;
    return std::isfinite(mx) && mx <= maxratio * mn;
  };
  pass(st, A, m, n, L1_, Q);
  {
    const bool ok1 = check(1e6);
    const double r = last_ratio_;
    int bin = !ok1 ? 7 : r < 2 ? 0 : r < 4 ? 1 : r < 8 ? 2 : r < 16 ? 3 : r < 64 ? 4 : r < 1e3 ? 5 : 6;
    ratio_hist[bin] += 1;
    if (!ok1) {
      ++n_fail;
      return false;
    }
  }
  if (last_ratio_ <= kSinglePassRatio) {
    // well-conditioned centre tensor (pivot ratio <= 8 => cond^2 * eps far below 1e-12): one pass is enough,
    // C = L1^H
    conj_transpose(st, n, n, L1_, n, C, n, true);
    ++n_single;
    return true;
  }
  TEVO_CUDA_CHECK(cudaMemcpyAsync(Q1_, Q, sizeof(cplx) * (size_t)m * n, cudaMemcpyDeviceToDevice, st));
  pass(st, Q1_, m, n, L2_, Q);
  if (!check(1.0 + 1e-3)) {  // second-pass factor must be ~identity
    ++n_fail;
    return false;
  }
  ++n_double;
  // C = L2^H L1^H = (L1 L2)^H
  const cplx ONE = mk(1, 0), ZERO = mk(0, 0);
  zgemm(st, OP_N, OP_N, n, n, n, ONE, L1_, n, L2_, n, ZERO, G_, n);
  conj_transpose(st, n, n, G_, n, C, n, true);
  return true;
}

}  // namespace tevo
