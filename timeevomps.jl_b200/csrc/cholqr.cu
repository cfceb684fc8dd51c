// CholeskyQR2 for the centre moves of orthogonalize! (src/bondgate.jl:48): A (m x n, m >= n) = Q C with Q^H Q = 1.
// Two passes of  G = A^H A (DMMA GEMM),  G = L L^H (blocked Cholesky, 64-wide diagonal blocks factored and inverted in
// shared memory by one CTA, panels and trailing updates by GEMM),  Q = A L^-H (block forward substitution by GEMM).
// Valid while cond(A) < ~1e6 (then the second pass restores orthogonality to machine precision); a non-positive
// pivot or a large diagonal ratio makes factor() return false and the caller uses the eigen-decomposition route.
#include "cholqr.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <optional>
#include <vector>

#include "kernels.h"
#include "trinv.cuh"
#include "prof.h"
#include "zgemm.h"

namespace tevo {
namespace {

constexpr int NB = 64;

// Cholesky of one nb x nb (nb <= 64) diagonal block of G at (j,j) and the inverse of its factor, both in shared
// memory (right-looking, 1024 threads): writes L_jj into Lm (lower part) and inv(L_jj) into Xm (lower part) at the
// same position (column-major, ld), and the extreme pivots into stats[0..1] (stats[0] < 0: breakdown).
__global__ void __launch_bounds__(1024) potrf_block_kernel(int nb, const cplx* __restrict__ G, cplx* __restrict__ Lm,
                                                           cplx* __restrict__ Xm, long ld, double* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char smraw[];
  cplx(*a)[NB] = reinterpret_cast<cplx(*)[NB]>(smraw);  // a[p][i] = L(i,p)   (column p contiguous over rows i)
  cplx(*x)[NB] = a + NB;                                // x[c][i] = inv(L)(i,c) up to the deferred row scaling
  const int tid = threadIdx.x;
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    const int r = e % NB, p = e / NB;
    a[p][r] = (r < nb && p < nb) ? G[(long)p * ld + r] : mk(r == p ? 1.0 : 0.0, 0.0);
  }
  __syncthreads();
  __shared__ double sinv[NB];
  double pmin, pmax;
  if (!chol_inv_lower_smem(a, x, sinv, nb, pmin, pmax)) {
    if (tid == 0) {
      stats[0] = -1.0;
      stats[1] = 0.0;
    }
    return;
  }
  for (int e = tid; e < nb * nb; e += blockDim.x) {
    const int r = e % nb, p = e / nb;
    Lm[(long)p * ld + r] = (r > p) ? cscale(sinv[p], a[p][r]) : (r == p ? mk(a[p][p].x * sinv[p], 0.0) : mk(0, 0));
    Xm[(long)p * ld + r] = (r >= p) ? cscale(sinv[r], x[p][r]) : mk(0, 0);
  }
  if (tid == 0) {
    stats[0] = pmin;
    stats[1] = pmax;
  }
}

// Register-tiled version of the same factorisation (default): 256 threads, thread (ti, tj) keeps the 4 x 4 tiles
// A(4ti.., 4tj..) and inv(L)(4ti.., 4tj..) in registers; per column k only the column of A and the row of inv(L) go
// through (double-buffered) shared memory, one barrier per column.  Same deferred pivot scaling as
// chol_inv_lower_smem: the updates divide by the pivot, L(i,k) = a(i,k)/l_k and inv(L)(i,c) = x(i,c)/l_i are applied
// when the results are written.  The 1024-thread shared-memory version spends ~90 us per block on index arithmetic
// (issue-bound); this one does 32 complex FMAs per 12 shared-memory loads.
__global__ void __launch_bounds__(256) potrf_block_reg_kernel(int nb, const cplx* __restrict__ G, cplx* __restrict__ Lm,
                                                              cplx* __restrict__ Xm, long ld,
                                                              double* __restrict__ stats) {
  __shared__ cplx scol[2][NB];  // column k of A (unscaled), rows 0..63
  __shared__ cplx srow[2][NB];  // row k of inv(L) (unscaled), columns 0..63
  __shared__ double sinv[NB];
  const int t = threadIdx.x, ti = t / 16, tj = t % 16;
  cplx a[4][4], x[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = 4 * ti + r, j = 4 * tj + c;
      a[r][c] = (i < nb && j < nb) ? (i >= j ? G[(long)j * ld + i] : mk(0, 0)) : mk(i == j ? 1.0 : 0.0, 0.0);
      x[r][c] = mk((i == j && i < nb) ? 1.0 : 0.0, 0.0);
    }
  double pmin = 1e300, pmax = 0.0;
  for (int kb = 0; kb < NB / 4; ++kb) {
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
      const int k = 4 * kb + kc;
      if (k >= nb) continue;  // uniform
      const int buf = k & 1;
      if (tj == kb) {
#pragma unroll
        for (int r = 0; r < 4; ++r) scol[buf][4 * ti + r] = a[r][kc];
      }
      if (ti == kb) {
#pragma unroll
        for (int c = 0; c < 4; ++c) srow[buf][4 * tj + c] = x[kc][c];
      }
      __syncthreads();
      const double dkk = scol[buf][k].x;
      if (!(dkk > 0.0) || !isfinite(dkk)) {  // uniform: every thread reads the same pivot
        if (t == 0) {
          stats[0] = -1.0;
          stats[1] = 0.0;
        }
        return;
      }
      const double inv = rsqrt(dkk), l = dkk * inv, rd = inv * inv;
      pmin = fmin(pmin, l);
      pmax = fmax(pmax, l);
      if (t == 0) sinv[k] = inv;
      if (ti >= kb) {
        cplx li[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) li[r] = (4 * ti + r > k) ? cscale(rd, scol[buf][4 * ti + r]) : mk(0, 0);
        if (tj >= kb) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const cplx lj = (4 * tj + c > k) ? cconj(scol[buf][4 * tj + c]) : mk(0, 0);
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r][c] = csub(a[r][c], cmul(li[r], lj));
          }
        }
        if (tj <= kb) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const cplx xr = srow[buf][4 * tj + c];  // zero for columns > k
#pragma unroll
            for (int r = 0; r < 4; ++r) x[r][c] = csub(x[r][c], cmul(li[r], xr));
          }
        }
      }
      // no second barrier: column k+1 uses the other buffer, and its barrier orders the writes of column k+2 after
      // these reads
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = 4 * ti + r, j = 4 * tj + c;
      if (i < nb && j < nb) {
        Lm[(long)j * ld + i] = (i > j) ? cscale(sinv[j], a[r][c]) : (i == j ? mk(a[r][c].x * sinv[j], 0.0) : mk(0, 0));
        Xm[(long)j * ld + i] = (i >= j) ? cscale(sinv[i], x[r][c]) : mk(0, 0);
      }
    }
  if (t == 0) {
    stats[0] = pmin;
    stats[1] = pmax;
  }
}

// E = G - I in place; stats[0] = max |E_ij| (non-negative doubles order like their bit patterns)
__global__ void gram_minus_identity_kernel(int n, cplx* __restrict__ G, unsigned long long* __restrict__ maxbits) {
  double m = 0.0;
  const long total = (long)n * n;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int i = (int)(e % n);
    const long j = e / n;
    cplx v = G[e];
    if (i == j) v.x -= 1.0;
    G[e] = v;
    m = fmax(m, fmax(fabs(v.x), fabs(v.y)));
  }
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(maxbits, (unsigned long long)__double_as_longlong(m));
}

// S = I - E/2 + (3/8) E2 ,  Sinv = I + E/2 - E2/8      ((I+E)^(-1/2) and (I+E)^(1/2) to second order)
__global__ void invsqrt_series_kernel(int n, const cplx* __restrict__ E, const cplx* __restrict__ E2, int use_e2,
                                      cplx* __restrict__ S, cplx* __restrict__ Sinv) {
  const long total = (long)n * n;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int i = (int)(e % n);
    const long j = e / n;
    const cplx x = E[e];
    const cplx y = use_e2 ? E2[e] : mk(0, 0);
    const double d = (i == j) ? 1.0 : 0.0;
    S[e] = mk(d - 0.5 * x.x + 0.375 * y.x, -0.5 * x.y + 0.375 * y.y);
    Sinv[e] = mk(d + 0.5 * x.x - 0.125 * y.x, 0.5 * x.y - 0.125 * y.y);
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
    zgemm_herm(st, OP_C, OP_N, n, m, ONE, A, m, A, m, ZERO, G_, n, false);  // Cholesky reads the lower part only
  }
  std::optional<ProfScope> pf;
  pf.emplace(st, P_CHOL_FACTOR);
  for (int jb = 0; jb < nblk; ++jb) {
    const int j = jb * NB, nb = std::min(NB, n - j), rest = n - j - nb;
    static const bool smem_version = getenv("TEVO_POTRF_SMEM") != nullptr;  // the 1024-thread shared-memory kernel
    if (smem_version)
      potrf_block_kernel<<<1, 1024, 2 * sizeof(cplx) * NB * NB, st>>>(nb, G_ + (long)j * n + j, L + (long)j * n + j,
                                                                       X + (long)j * n + j, n, d_stats_ + 2 * jb);
    else
      potrf_block_reg_kernel<<<1, 256, 0, st>>>(nb, G_ + (long)j * n + j, L + (long)j * n + j, X + (long)j * n + j, n,
                                                d_stats_ + 2 * jb);
    TEVO_LAUNCHED();
    if (rest > 0) {
      // panel: L(j+nb:, j:j+nb) = G(j+nb:, j:j+nb) * inv(L_jj)^H ; trailing: G(j+nb:, j+nb:) -= panel * panel^H
      zgemm(st, OP_N, OP_C, rest, nb, nb, ONE, G_ + (long)j * n + j + nb, n, X + (long)j * n + j, n, ZERO,
            L + (long)j * n + j + nb, n);
      zgemm_herm(st, OP_N, OP_C, rest, nb, MONE, L + (long)j * n + j + nb, n, L + (long)j * n + j + nb, n, ONE,
                 G_ + (long)(j + nb) * n + j + nb, n, false);
    }
  }
  pf.reset();
  std::optional<ProfScope> pi;
  pi.emplace(st, P_CHOL_INV);
  // inverse of L by pairwise merging of diagonal ranges: inv([A 0; B C]) = [A^-1 0; -C^-1 B A^-1, C^-1]
  // regular part: n = q * NB (+ tail); level s merges blocks [lo, lo+s) and [lo+s, lo+2s), all pairs of a level in
  // two strided-batched GEMMs.  Irregular leftovers (n not a multiple of 2s) are merged one by one afterwards.
  std::vector<int> bounds;
  for (int b = 0; b < n; b += NB) bounds.push_back(b);
  bounds.push_back(n);
  while (bounds.size() > 2) {
    std::vector<int> nbnd;
    size_t k = 0;
    // count leading pairs with identical sizes
    const int s = bounds[1] - bounds[0];
    int nreg = 0;
    while (2 * (size_t)nreg + 2 < bounds.size() && bounds[2 * nreg + 1] - bounds[2 * nreg] == s &&
           bounds[2 * nreg + 2] - bounds[2 * nreg + 1] == s)
      ++nreg;
    if (nreg > 0) {
      const long bs = (long)2 * s * (n + 1);
      // T_b = B_b * inv(A_b) ; X21_b = -inv(C_b) * T_b
      zgemm_batched(st, OP_N, OP_N, s, s, s, ONE, L + s, n, bs, X, n, bs, ZERO, T_, s, (long)s * s, nreg);
      zgemm_batched(st, OP_N, OP_N, s, s, s, MONE, X + (long)s * n + s, n, bs, T_, s, (long)s * s, ZERO, X + s, n, bs, nreg);
      for (int q = 0; q < nreg; ++q) nbnd.push_back(bounds[2 * q]);
      k = 2 * (size_t)nreg;
    }
    for (; k + 2 < bounds.size(); k += 2) {
      const int lo = bounds[k], mid = bounds[k + 1], hi = bounds[k + 2];
      const int s1 = mid - lo, s2 = hi - mid;
      zgemm(st, OP_N, OP_N, s2, s1, s1, ONE, L + (long)lo * n + mid, n, X + (long)lo * n + lo, n, ZERO, T_, s2);
      zgemm(st, OP_N, OP_N, s2, s1, s2, MONE, X + (long)mid * n + mid, n, T_, s2, ZERO, X + (long)lo * n + mid, n);
      nbnd.push_back(lo);
    }
    for (; k < bounds.size(); ++k) nbnd.push_back(bounds[k]);
    bounds = nbnd;
  }
  pi.reset();
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
  const cplx ONE = mk(1, 0), ZERO = mk(0, 0);
  TEVO_CUDA_CHECK(cudaMemcpyAsync(Q1_, Q, sizeof(cplx) * (size_t)m * n, cudaMemcpyDeviceToDevice, st));
  // second pass.  G2 = Q1^H Q1 = I + E: when E is tiny (the usual case) (I+E)^(-1/2) is applied as a series with two
  // GEMMs instead of a second Cholesky factorisation; otherwise the full CholeskyQR pass is repeated.
  {
    ProfScope ps(st, P_CHOL_GRAM);
    zgemm_herm(st, OP_C, OP_N, n, m, ONE, Q1_, m, Q1_, m, ZERO, G_, n, true);
  }
  unsigned long long* maxbits = reinterpret_cast<unsigned long long*>(d_stats_ + 2 * 127);
  TEVO_CUDA_CHECK(cudaMemsetAsync(maxbits, 0, sizeof(unsigned long long), st));
  {
    const long total = (long)n * n;
    int blocks = (int)std::min<long>((total + 255) / 256, 148 * 8);
    gram_minus_identity_kernel<<<blocks, 256, 0, st>>>(n, G_, maxbits);
    TEVO_LAUNCHED();
  }
  TEVO_CUDA_CHECK(cudaMemcpyAsync(h_stats_ + 2 * 127, maxbits, sizeof(double), cudaMemcpyDeviceToHost, st));
  TEVO_CUDA_CHECK(cudaStreamSynchronize(st));
  const double emax = h_stats_[2 * 127];
  if (std::isfinite(emax) && emax * n <= 1e-5) {
    ProfScope ps(st, P_CHOL_Q);
    const int use_e2 = (emax * n > 1e-8) ? 1 : 0;
    if (use_e2) zgemm(st, OP_N, OP_N, n, n, n, ONE, G_, n, G_, n, ZERO, L2_, n);  // E^2
    {
      const long total = (long)n * n;
      int blocks = (int)std::min<long>((total + 255) / 256, 148 * 8);
      invsqrt_series_kernel<<<blocks, 256, 0, st>>>(n, G_, L2_, use_e2, Linv_, T_);
      TEVO_LAUNCHED();
    }
    zgemm(st, OP_N, OP_N, m, n, n, ONE, Q1_, m, Linv_, n, ZERO, Q, m);  // Q = Q1 (I+E)^(-1/2)
    zgemm(st, OP_N, OP_C, n, n, n, ONE, T_, n, L1_, n, ZERO, C, n);     // C = (I+E)^(1/2) L1^H
    ++n_series;
    return true;
  }
  pass(st, Q1_, m, n, L2_, Q);
  // C = L2^H L1^H = (L1 L2)^H
  zgemm(st, OP_N, OP_N, n, n, n, ONE, L1_, n, L2_, n, ZERO, G_, n);
  conj_transpose(st, n, n, G_, n, C, n, true);
  return true;
}

}  // namespace tevo
