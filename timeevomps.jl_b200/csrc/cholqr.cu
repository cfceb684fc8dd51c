// CholeskyQR2 for the centre moves of orthogonalize! (src/bondgate.jl:48): A (m x n, m >= n) = Q C with Q^H Q = 1.
// Two passes of  G = A^H A (DMMA GEMM),  G = L L^H (blocked Cholesky, 64-wide diagonal blocks factored and inverted in
// shared memory by one CTA, panels and trailing updates by GEMM),  Q = A L^-H (block forward substitution by GEMM).
// Valid while cond(A) < ~1e6 (then the second pass restores orthogonality to machine precision); a non-positive
// pivot or a large diagonal ratio makes factor() return false and the caller uses the eigen-decomposition route.
#include "cholqr.h"

#include <algorithm>
#include <cmath>

#include "kernels.h"
#include "prof.h"
#include "zgemm.h"

namespace tevo {
namespace {

constexpr int NB = 64;

// Cholesky of one NB x NB (or smaller, nb) diagonal block of G at (j,j); writes L_jj into Lm (upper part zeroed),
// its inverse into Linv (NB x NB, ld NB), and the extreme pivots into stats[2*blk], stats[2*blk+1] (min < 0: failed)
__global__ void __launch_bounds__(256) potrf_block_kernel(int nb, const cplx* __restrict__ G, long ldg, cplx* __restrict__ Lm,
                                                          long ldl, cplx* __restrict__ Linv, double* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char smraw[];
  cplx(*a)[NB + 1] = reinterpret_cast<cplx(*)[NB + 1]>(smraw);        // a[i][j]
  cplx(*x)[NB + 1] = a + NB;                                           // inverse
  __shared__ double s_min, s_max;
  __shared__ int s_fail;
  const int tid = threadIdx.x;
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    const int i = e % NB, j = e / NB;
    a[i][j] = (i < nb && j < nb) ? G[(long)j * ldg + i] : mk(i == j ? 1.0 : 0.0, 0.0);
    x[i][j] = mk(0, 0);
  }
  if (tid == 0) {
    s_min = 1e300;
    s_max = 0.0;
    s_fail = 0;
  }
  __syncthreads();
  for (int k = 0; k < nb; ++k) {
    const double dkk = a[k][k].x;
    __syncthreads();
    if (!(dkk > 0.0) || !isfinite(dkk)) {
      if (tid == 0) s_fail = 1;
      break;
    }
    const double l = sqrt(dkk);
    if (tid == 0) {
      a[k][k] = mk(l, 0.0);
      s_min = fmin(s_min, l);
      s_max = fmax(s_max, l);
    }
    const double inv = 1.0 / l;
    for (int i = k + 1 + tid; i < nb; i += blockDim.x) a[i][k] = cscale(inv, a[i][k]);
    __syncthreads();
    // trailing update of the lower triangle: a[i][j] -= a[i][k] conj(a[j][k]),  k < j <= i
    const int rem = nb - k - 1;
    for (int e = tid; e < rem * rem; e += blockDim.x) {
      const int i = k + 1 + e % rem, j = k + 1 + e / rem;
      if (j <= i) {
        const cplx p = cmul(a[i][k], cconj(a[j][k]));
        a[i][j] = csub(a[i][j], p);
      }
    }
    __syncthreads();
  }
  __syncthreads();
  if (s_fail) {
    if (tid == 0) {
      stats[0] = -1.0;
      stats[1] = 0.0;
    }
    return;
  }
  // inverse of the lower-triangular factor, one thread per column (forward substitution)
  if (tid < nb) {
    const int j = tid;
    for (int i = j; i < nb; ++i) {
      cplx s = mk(i == j ? 1.0 : 0.0, 0.0);
      for (int p = j; p < i; ++p) s = csub(s, cmul(a[i][p], x[p][j]));
      x[i][j] = cscale(1.0 / a[i][i].x, s);
    }
  }
  __syncthreads();
  for (int e = tid; e < nb * nb; e += blockDim.x) {
    const int i = e % nb, j = e / nb;
    Lm[(long)j * ldl + i] = (j <= i) ? a[i][j] : mk(0, 0);
  }
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    const int i = e % NB, j = e / NB;
    Linv[(long)j * NB + i] = (i < nb && j < nb) ? x[i][j] : mk(0, 0);
  }
  if (tid == 0) {
    stats[0] = s_min;
    stats[1] = s_max;
  }
}

__global__ void zero_upper_block_kernel(int rows, int cols, cplx* __restrict__ M, long ld) {
  const long total = (long)rows * cols;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x)
    M[(e / rows) * ld + (e % rows)] = mk(0, 0);
}

inline int gridfor(long total) {
  long b = (total + 255) / 256;
  if (b < 1) b = 1;
  if (b > 148 * 8) b = 148 * 8;
  return (int)b;
}

}  // namespace

void CholQR::init() {
  TEVO_CUDA_CHECK(cudaMalloc(&d_stats_, sizeof(double) * 2 * 128));
  TEVO_CUDA_CHECK(cudaMallocHost(&h_stats_, sizeof(double) * 2 * 128));
  TEVO_CUDA_CHECK(cudaFuncSetAttribute(potrf_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(2 * sizeof(cplx) * NB * (NB + 1))));
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
    TEVO_CUDA_CHECK(cudaMalloc(&Linv_, sizeof(cplx) * (size_t)NB * NB * ((n + NB - 1) / NB)));
    ncap_ = n;
    mcap_ = 0;
  }
  if ((size_t)m * n > mcap_) {
    if (T_) cudaFree(T_);
    if (Q1_) cudaFree(Q1_);
    TEVO_CUDA_CHECK(cudaMalloc(&T_, sizeof(cplx) * (size_t)m * NB));
    TEVO_CUDA_CHECK(cudaMalloc(&Q1_, sizeof(cplx) * (size_t)m * n));
    mcap_ = (size_t)m * n;
  }
}

// one pass: Q = A L^-H with A^H A = L L^H; returns the pivot extremes through stats (device)
void CholQR::pass(cudaStream_t st, const cplx* A, int m, int n, cplx* L, cplx* Q) {
  const cplx ONE = mk(1, 0), ZERO = mk(0, 0), MONE = mk(-1, 0);
  const int nblk = (n + NB - 1) / NB;
  zgemm(st, OP_C, OP_N, n, n, m, ONE, A, m, A, m, ZERO, G_, n);
  for (int jb = 0; jb < nblk; ++jb) {
    const int j = jb * NB, nb = std::min(NB, n - j), rest = n - j - nb;
    potrf_block_kernel<<<1, 256, 2 * sizeof(cplx) * NB * (NB + 1), st>>>(nb, G_ + (long)j * n + j, n, L + (long)j * n + j, n,
                                                                       Linv_ + (size_t)jb * NB * NB, d_stats_ + 2 * jb);
    TEVO_LAUNCHED();
    if (j > 0) {  // zero the block row above the diagonal block (strictly upper part of L)
      zero_upper_block_kernel<<<gridfor((long)j * nb), 256, 0, st>>>(j, nb, L + (long)j * n, n);
      TEVO_LAUNCHED();
    }
    if (rest > 0) {
      // panel: L(j+nb:, j:j+nb) = G(j+nb:, j:j+nb) * inv(L_jj)^H
      zgemm(st, OP_N, OP_C, rest, nb, nb, ONE, G_ + (long)j * n + j + nb, n, Linv_ + (size_t)jb * NB * NB, NB, ZERO,
            L + (long)j * n + j + nb, n);
      // trailing: G(j+nb:, j+nb:) -= panel * panel^H
      zgemm(st, OP_N, OP_C, rest, rest, nb, MONE, L + (long)j * n + j + nb, n, L + (long)j * n + j + nb, n, ONE,
            G_ + (long)(j + nb) * n + j + nb, n);
    }
  }
  // Q = A L^-H by block forward substitution over the block columns
  for (int jb = 0; jb < nblk; ++jb) {
    const int j = jb * NB, nb = std::min(NB, n - j);
    TEVO_CUDA_CHECK(cudaMemcpy2DAsync(T_, sizeof(cplx) * m, A + (long)j * m, sizeof(cplx) * m, sizeof(cplx) * m, nb,
                                      cudaMemcpyDeviceToDevice, st));
    if (j > 0)  // T -= Q(:, 0:j) * L(j:j+nb, 0:j)^H
      zgemm(st, OP_N, OP_C, m, nb, j, MONE, Q, m, L + j, n, ONE, T_, m);
    zgemm(st, OP_N, OP_C, m, nb, nb, ONE, T_, m, Linv_ + (size_t)jb * NB * NB, NB, ZERO, Q + (long)j * m, m);
  }
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
