// In-shared-memory Cholesky factor and triangular inverse of one <= 64 x 64 block with 256 threads
// (4 threads share every row / column dot product and combine by warp shuffles).
#pragma once
#include "common.cuh"

namespace tevo {

constexpr int TRB = 64;

// a[p][i] = L(i,p), lower triangular (entries with i < p ignored).  x[p][c] <- inv(L)(p,c) (zero above the diagonal).
// blockDim.x == 256.  Ends with __syncthreads().
__device__ inline void trinv_lower_256(cplx (*a)[TRB], cplx (*x)[TRB], int nb) {
  const int c = threadIdx.x >> 2, part = threadIdx.x & 3;
  for (int r = 0; r < nb; ++r) {  // uniform control flow: every lane takes part in the shuffles
    const bool active = (c < nb) && (r >= c);
    cplx s = mk(0, 0);
    if (active)
      for (int p = c + part; p < r; p += 4) cfma(s, a[p][r], x[p][c]);
    s.x += __shfl_xor_sync(0xffffffffu, s.x, 1);
    s.y += __shfl_xor_sync(0xffffffffu, s.y, 1);
    s.x += __shfl_xor_sync(0xffffffffu, s.x, 2);
    s.y += __shfl_xor_sync(0xffffffffu, s.y, 2);
    if (c < nb && part == 0) {
      cplx q = mk(0, 0);
      if (active) {
        const cplx d = a[r][r];
        const double dn = 1.0 / cabs2(d);
        const cplx num = mk((r == c ? 1.0 : 0.0) - s.x, -s.y);
        q = mk((num.x * d.x + num.y * d.y) * dn, (num.y * d.x - num.x * d.y) * dn);  // num / d
      }
      x[r][c] = q;
    }
    __syncwarp();
  }
  __syncthreads();
}

}  // namespace tevo
