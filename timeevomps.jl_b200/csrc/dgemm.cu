// Real FP64 GEMM on the DMMA pipe for the divide-and-conquer eigenvector updates (stedc.cu):
//   for every merge node g of one tree level:  C_g(M_g x K_g) = A_g(M_g x K_g) * B_g(K_g x K_g)
// where K_g (the number of non-deflated eigenvalues) lives in device memory - the grid is sized for the worst
// case and surplus CTAs exit.  Column-major, non-transposed, common leading dimension.
// CTA tile 128x128, 8 warps (2 x 4), warp tile 64x32 = 8x4 DMMA blocks (64 fp64 accumulators/thread).
#include <mutex>

#include "common.cuh"
#include "stedc.h"

namespace tevo {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 3, NT = 256;
constexpr int APAD = 4;  // A tile stored [k][m] (m contiguous): stride (BM+4) doubles -> 8*4=32 B shift per k
constexpr int BPAD = 4;  // B tile stored [n][k] (k contiguous): stride (BK+4) doubles

__device__ inline void cp_async8(void* smem_dst, const void* gsrc, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  int sz = pred ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}
__device__ inline void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ inline void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ inline void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(NT, 1)
dgemm_merge_kernel(const MergeDesc* __restrict__ descs, const int* __restrict__ Kdev, const double* __restrict__ A,
                   const double* __restrict__ B, double* __restrict__ C, long ld) {
  const MergeDesc md = descs[blockIdx.z];
  const int M = md.hi - md.lo;
  const int Kk = Kdev[blockIdx.z];
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  if (m0 >= M || n0 >= Kk) return;
  const long off = (long)md.lo * ld + md.lo;
  const double* Ag = A + off;
  const double* Bg = B + off;
  double* Cg = C + off;
  const int N = Kk, K = Kk;

  extern __shared__ __align__(16) unsigned char smraw[];
  double* sA = reinterpret_cast<double*>(smraw);           // STAGES * BK * (BM+APAD)
  double* sB = sA + STAGES * BK * (BM + APAD);              // STAGES * BN * (BK+BPAD)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = (warp & 1) * 64, wn = (warp >> 1) * 32;
  const int g = lane >> 2, tq = lane & 3;
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto load = [&](int stage, int k0) {
    double* a = sA + stage * BK * (BM + APAD);
    double* b = sB + stage * BN * (BK + BPAD);
    for (int e = threadIdx.x; e < BM * BK; e += NT) {
      int k = e / BM, r = e % BM;
      bool p = (m0 + r < M) && (k0 + k < K);
      cp_async8(&a[k * (BM + APAD) + r], p ? Ag + (long)(k0 + k) * ld + (m0 + r) : Ag, p);
    }
    for (int e = threadIdx.x; e < BN * BK; e += NT) {
      int c = e / BK, k = e % BK;
      bool p = (n0 + c < N) && (k0 + k < K);
      cp_async8(&b[c * (BK + BPAD) + k], p ? Bg + (long)(n0 + c) * ld + (k0 + k) : Bg, p);
    }
  };
  const int nk = (K + BK - 1) / BK;
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) load(s, s * BK);
    cp_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_wait<STAGES - 2>();
    __syncthreads();
    {
      int nx = kt + STAGES - 1;
      if (nx < nk) load(nx % STAGES, nx * BK);
      cp_commit();
    }
    const double* a = sA + (kt % STAGES) * BK * (BM + APAD);
    const double* b = sB + (kt % STAGES) * BN * (BK + BPAD);
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double af[8], bf[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) af[i] = a[(kk + tq) * (BM + APAD) + wm + i * 8 + g];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = b[(wn + j * 8 + g) * (BK + BPAD) + kk + tq];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_wait<0>();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + wm + i * 8 + g;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = n0 + wn + j * 8 + tq * 2 + e;
        if (col < N) Cg[(long)col * ld + row] = acc[i][j][e];
      }
  }
}

}  // namespace

void dgemm_merge(cudaStream_t st, const MergeDesc* descs_dev, const int* K_dev, int nmerge, int maxN, const double* A,
                 const double* B, double* C, long ld) {
  if (nmerge <= 0) return;
  const size_t smem = (size_t)STAGES * (BK * (BM + APAD) + BN * (BK + BPAD)) * sizeof(double);
  static PerDeviceOnce once;
  once([&]() {
    TEVO_CUDA_CHECK(cudaFuncSetAttribute(dgemm_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  });
  dim3 grid((maxN + BM - 1) / BM, (maxN + BN - 1) / BN, nmerge);
  dgemm_merge_kernel<<<grid, NT, smem, st>>>(descs_dev, K_dev, A, B, C, ld);
  TEVO_LAUNCHED();
}

}  // namespace tevo
