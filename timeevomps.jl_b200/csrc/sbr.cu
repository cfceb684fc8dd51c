// Two-stage Hermitian tridiagonalisation (successive band reduction) for the truncated split
// (replacebond!, src/bondgate.jl:51, src/tdvp.jl:64 -> LAPACK zheev*/zgesdd in the reference).
//
//   stage 1  dense -> band of width SB = 8   sbr_stage1_kernel: persistent cooperative kernel; per block of 8 columns a
//            Householder QR of the sub-band panel by one CTA (panel in registers) and ONE pass over the trailing
//            matrix that applies the rank-16 two-sided update of the previous block and multiplies the updated
//            matrix with the new reflectors in the same sweep (A is read and written once per 8 columns instead of
//            being read once per column by the one-stage Householder reduction of heig.cu)
//   stage 2  band -> real tridiagonal        band_chase_kernel: bulge chasing, one warp per sweep, sweeps pipelined
//            through progress counters (a sweep may run step k once its predecessor has finished step k+1)
//   back     X = Q2 Z                        q2_apply_kernel: column tiles of the eigenvector matrix stay in shared memory
//            while every length-8 reflector of stage 2 is applied (sweeps in reverse order, the reflectors of one
//            sweep act on disjoint rows); Q1 is applied afterwards by the compact-WY GEMMs of heig.cu
//
// proto/two_stage.py is the NumPy statement of the same data flow (checked against numpy.linalg.eigh).
#include "sbr.h"

#include <algorithm>
#include <cooperative_groups.h>

#include "prof.h"

namespace cg = cooperative_groups;

namespace tevo {
namespace {

constexpr int SB = SBR_BAND;  // 8
constexpr int LDB = SBR_LDB;  // 16 stored diagonals (bulge included)

__device__ inline cplx ldcg_c(const cplx* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }
__device__ inline cplx shfl_c(cplx v, int src) {
  return mk(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
__device__ inline cplx shfl_xor_c(cplx v, int m) {
  return mk(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}
// a -= x * y
__device__ inline void cfms(cplx& a, cplx x, cplx y) {
  a.x -= x.x * y.x - x.y * y.y;
  a.y -= x.x * y.y + x.y * y.x;
}

// zlarfg: H = I - tau v v^H, v_0 = 1, v_j = x_j * scal (j >= 1), H^H (alpha; x) = (beta; 0), beta real
__device__ inline void larfg(cplx alpha, double xn2, double& beta, cplx& tau, cplx& scal) {
  const double s2 = cabs2(alpha) + xn2;
  if ((xn2 == 0.0 && alpha.y == 0.0) || !(s2 > 1e-290)) {
    beta = alpha.x;
    tau = mk(0, 0);
    scal = mk(0, 0);
    return;
  }
  const double nx = sqrt(s2);
  beta = -copysign(nx, alpha.x);
  const double ib = 1.0 / beta;
  tau = mk((beta - alpha.x) * ib, -alpha.y * ib);
  const cplx den = mk(alpha.x - beta, alpha.y);
  const double rdn = 1.0 / cabs2(den);
  scal = mk(den.x * rdn, -den.y * rdn);
}

// ------------------------------------------------------------------------------------------------------------------
// stage 2: band (width 8, lower storage Bd[d + LDB*c] = B(c+d, c), d < 16) -> real symmetric tridiagonal (d, e)
// ------------------------------------------------------------------------------------------------------------------
// One warp per sweep s (eliminates column s below the sub-diagonal and chases the bulge to the end of the matrix).
// Step k of sweep s works on the window rows r0 .. r0+15, columns r0 .. r0+7, r0 = s + 1 + 8k:
//   D (rows 0..7, Hermitian) <- H^H D H,  E (rows 8..15) <- E H,  new reflector H2 from E(:,0),  E <- H2^H E.
// Lane l holds the 4 window entries (4g+i, c), c = l & 7, g = l >> 3.
// The reflector of step (s, k) goes to RV[(s*KS + k)*9 + 0..7] (v, zero padded) and [8] (tau).
__global__ void __launch_bounds__(256) band_chase_kernel(int n, cplx* __restrict__ Bd, cplx* __restrict__ RV, int KS,
                                                         double* __restrict__ dd, double* __restrict__ ee,
                                                         int* prog) {
  const int lane = threadIdx.x & 31;
  const int c = lane & 7, g = lane >> 3;
  const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int NW = gridDim.x * (blockDim.x >> 5);
  const int DONE = 1 << 30;
  for (int s = gw; s <= n - 2; s += NW) {
    const int Ks = (n - 1 - s + SB - 1) / SB;
    cplx vc = mk(0, 0), tau = mk(0, 0);
    cplx vr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) vr[i] = mk(0, 0);
    for (int k = 0; k < Ks; ++k) {
      const int r0 = s + 1 + SB * k;
      if (s > 0) {  // predecessor must have finished step k+1 (or the whole sweep)
        if (lane == 0) {
          int v;
          do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(prog + s - 1) : "memory");
          } while (v < k + 2);
        }
        __syncwarp();
      }
      if (k == 0) {
        // first reflector of the sweep from column s, rows s+1 .. s+8
        const cplx* col = Bd + (long)LDB * s;
        const cplx xc = (r0 + c < n) ? ldcg_c(col + 1 + c) : mk(0, 0);
        cplx xr[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = 4 * (g & 1) + i;
          xr[i] = (r0 + j < n) ? ldcg_c(col + 1 + j) : mk(0, 0);
        }
        double xn2 = (c >= 1) ? cabs2(xc) : 0.0;
        xn2 += __shfl_xor_sync(0xffffffffu, xn2, 1);
        xn2 += __shfl_xor_sync(0xffffffffu, xn2, 2);
        xn2 += __shfl_xor_sync(0xffffffffu, xn2, 4);
        xn2 = __shfl_sync(0xffffffffu, xn2, 0);
        const cplx alpha = shfl_c(xc, 0);
        double beta;
        cplx scal;
        larfg(alpha, xn2, beta, tau, scal);
        vc = (c == 0) ? mk(1, 0) : cmul(xc, scal);
#pragma unroll
        for (int i = 0; i < 4; ++i) vr[i] = (4 * (g & 1) + i == 0) ? mk(1, 0) : cmul(xr[i], scal);
        if (lane < SB && r0 + lane < n) Bd[(long)LDB * s + 1 + lane] = (lane == 0) ? mk(beta, 0) : mk(0, 0);
        if (lane == 0) {
          ee[s] = beta;
          dd[s] = ldcg_c(col).x;
        }
      }
      // ---- load the window ----
      cplx w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = 4 * g + i;
        const int rg = r0 + r, cg_ = r0 + c;
        cplx x = mk(0, 0);
        if (rg < n && cg_ < n) {
          if (r >= c)
            x = ldcg_c(Bd + (long)LDB * cg_ + (r - c));
          else  // upper triangle of D by symmetry
            x = cconj(ldcg_c(Bd + (long)LDB * rg + (c - r)));
        }
        w[i] = x;
      }
      // record the reflector of this step
      {
        cplx* rv = RV + ((long)s * KS + k) * 9;
        if (lane < SB) rv[lane] = vc;
        if (lane == SB) rv[SB] = tau;
      }
      // ---- W <- W H :  z_r = sum_c W(r,c) v_c ;  W(r,c) -= tau z_r conj(v_c) ----
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        cplx z = cmul(w[i], vc);
        z = cadd(z, shfl_xor_c(z, 1));
        z = cadd(z, shfl_xor_c(z, 2));
        z = cadd(z, shfl_xor_c(z, 4));
        cfms(w[i], cmul(tau, z), cconj(vc));
      }
      // ---- D <- H^H D (rows 0..7: g < 2):  y_c = sum_r conj(v_r) D(r,c) ;  D(r,c) -= conj(tau) v_r y_c ----
      {
        cplx y = mk(0, 0);
        if (g < 2) {
#pragma unroll
          for (int i = 0; i < 4; ++i) cfma_conj(y, vr[i], w[i]);
        }
        y = cadd(y, shfl_xor_c(y, 8));
        if (g < 2) {
          const cplx ty = cmul(cconj(tau), y);
#pragma unroll
          for (int i = 0; i < 4; ++i) cfms(w[i], vr[i], ty);
        }
      }
      // ---- next reflector from E(:,0) (lanes 16 and 24), E <- H2^H E ----
      const bool hasE = (r0 + SB < n);
      cplx vc2 = mk(0, 0), tau2 = mk(0, 0);
      cplx vr2[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) vr2[i] = mk(0, 0);
      if (hasE) {
        // x_j (j < 8) lives in lane 16 + 8*(j >> 2), register j & 3
        cplx xc = mk(0, 0);
        cplx xr[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const cplx t = shfl_c(w[i], 16 + 8 * (c >> 2));
          if ((c & 3) == i) xc = t;
          xr[i] = shfl_c(w[i], 16 + 8 * (g & 1));
        }
        double xn2 = (c >= 1) ? cabs2(xc) : 0.0;
        xn2 += __shfl_xor_sync(0xffffffffu, xn2, 1);
        xn2 += __shfl_xor_sync(0xffffffffu, xn2, 2);
        xn2 += __shfl_xor_sync(0xffffffffu, xn2, 4);
        xn2 = __shfl_sync(0xffffffffu, xn2, 0);
        const cplx alpha = shfl_c(xc, 0);
        double beta2;
        cplx scal;
        larfg(alpha, xn2, beta2, tau2, scal);
        vc2 = (c == 0) ? mk(1, 0) : cmul(xc, scal);
#pragma unroll
        for (int i = 0; i < 4; ++i) vr2[i] = (4 * (g & 1) + i == 0) ? mk(1, 0) : cmul(xr[i], scal);
        cplx y = mk(0, 0);
        if (g >= 2) {
#pragma unroll
          for (int i = 0; i < 4; ++i) cfma_conj(y, vr2[i], w[i]);
        }
        y = cadd(y, shfl_xor_c(y, 8));
        if (g >= 2) {
          const cplx ty = cmul(cconj(tau2), y);
#pragma unroll
          for (int i = 0; i < 4; ++i) cfms(w[i], vr2[i], ty);
          if (c == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) w[i] = (g == 2 && i == 0) ? mk(beta2, 0) : mk(0, 0);
          }
        }
      }
      // ---- store the lower part of the window ----
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = 4 * g + i;
        const int rg = r0 + r, cg_ = r0 + c;
        if (rg < n && cg_ < n && r >= c) {
          cplx x = w[i];
          if (r == c) x.y = 0.0;  // Hermitian diagonal
          Bd[(long)LDB * cg_ + (r - c)] = x;
        }
      }
      vc = vc2;
      tau = tau2;
#pragma unroll
      for (int i = 0; i < 4; ++i) vr[i] = vr2[i];
      __syncwarp();
      if (lane == 0) {
        __threadfence();
        const int val = (k + 1 == Ks) ? DONE : k + 1;
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(prog + s), "r"(val) : "memory");
      }
    }
    if (s == n - 2 && lane == 0) {
      // the last diagonal entry is final once the last sweep is done (it was written by this warp)
      __threadfence();
      dd[n - 1] = ldcg_c(Bd + (long)LDB * (n - 1)).x;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// X <- Q2 X for NC columns per CTA held in shared memory.  X(:, col) starts as the real eigenvector Z(:, perm[col]) of the
// tridiagonal matrix.  Shared layout: row r of local column q at  q*npad + (r & 7)*(npad/8) + (r >> 3)  so that the
// threads of a warp (consecutive reflectors = row offsets 8 apart) touch consecutive addresses.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) q2_apply_kernel(int n, int k, int NC, const double* __restrict__ Z, long ldz,
                                                       const int* __restrict__ perm, const cplx* __restrict__ RV, int KS,
                                                       cplx* __restrict__ U, long ldu) {
  extern __shared__ __align__(16) unsigned char smraw[];
  cplx* X = reinterpret_cast<cplx*>(smraw);
  const int npad = (n + 7) & ~7;
  const int n8 = npad >> 3;
  const int col0 = blockIdx.x * NC;
  const int nc = min(NC, k - col0);
  if (nc <= 0) return;
  const int tid = threadIdx.x;
  for (int q = 0; q < nc; ++q) {
    const double* zc = Z + (long)(perm ? perm[col0 + q] : col0 + q) * ldz;
    for (int r = tid; r < npad; r += blockDim.x)
      X[q * npad + (r & 7) * n8 + (r >> 3)] = (r < n) ? mk(zc[r], 0.0) : mk(0, 0);
  }
  __syncthreads();
  for (int s = n - 2; s >= 0; --s) {
    const int Ks = (n - 1 - s + SB - 1) / SB;
    const int ntask = Ks * nc;
    for (int id = tid; id < ntask; id += blockDim.x) {
      const int kk = id % Ks, q = id / Ks;
      const cplx* rv = RV + ((long)s * KS + kk) * 9;
      const cplx tau = rv[SB];
      if (tau.x == 0.0 && tau.y == 0.0) continue;
      const int r0 = s + 1 + SB * kk;
      cplx* xq = X + q * npad;
      cplx x[SB], v[SB];
      cplx dot = mk(0, 0);
#pragma unroll
      for (int i = 0; i < SB; ++i) {
        const int r = r0 + i;  // r < npad always holds for rows with v != 0; padded rows hold zeros
        v[i] = rv[i];
        x[i] = (r < npad) ? xq[(r & 7) * n8 + (r >> 3)] : mk(0, 0);
        cfma_conj(dot, v[i], x[i]);
      }
      const cplx td = cmul(tau, dot);
#pragma unroll
      for (int i = 0; i < SB; ++i) {
        const int r = r0 + i;
        cfms(x[i], v[i], td);
        if (r < npad) xq[(r & 7) * n8 + (r >> 3)] = x[i];
      }
    }
    __syncthreads();
  }
  for (int q = 0; q < nc; ++q)
    for (int r = tid; r < n; r += blockDim.x) U[(long)(col0 + q) * ldu + r] = X[q * npad + (r & 7) * n8 + (r >> 3)];
}

}  // namespace

// ==================================================================================================================
// host side
// ==================================================================================================================
void Sbr::destroy() {
  auto fr = [](void* p) {
    if (p) cudaFree(p);
  };
  fr(Bd_); fr(RV_); fr(prog_); fr(Vb_); fr(Wb_); fr(Yb_); fr(small_);
  Bd_ = RV_ = Vb_ = Wb_ = Yb_ = nullptr;
  small_ = nullptr;
  prog_ = nullptr;
  cap_ = 0;
}

void Sbr::ensure(int n) {
  if (n <= cap_) return;
  destroy();
  const int KS = (n + SB - 1) / SB;
  TEVO_CUDA_CHECK(cudaMalloc(&Bd_, sizeof(cplx) * (size_t)LDB * (n + 16)));
  TEVO_CUDA_CHECK(cudaMalloc(&RV_, sizeof(cplx) * (size_t)n * KS * 9));
  TEVO_CUDA_CHECK(cudaMalloc(&prog_, sizeof(int) * (size_t)(n + 8)));
  TEVO_CUDA_CHECK(cudaMalloc(&Vb_, sizeof(cplx) * (size_t)2 * SB * (n + 64)));
  TEVO_CUDA_CHECK(cudaMalloc(&Wb_, sizeof(cplx) * (size_t)2 * SB * (n + 64)));
  TEVO_CUDA_CHECK(cudaMalloc(&Yb_, sizeof(cplx) * (size_t)2 * SB * (n + 64)));
  TEVO_CUDA_CHECK(cudaMalloc(&small_, sizeof(double) * SBR_SMALL_DOUBLES));
  cap_ = n;
}

void Sbr::chase(cudaStream_t st, int n, double* d, double* e) {
  // band -> tridiagonal; Bd_ holds the band (diagonals 0..8 filled, 9..15 zero)
  const int KS = (n + SB - 1) / SB;
  ks_ = KS;
  TEVO_CUDA_CHECK(cudaMemsetAsync(prog_, 0, sizeof(int) * (size_t)(n + 8), st));
  if (n == 1) {
    // d[0] = B(0,0): handled by the caller (no sweep exists)
    return;
  }
  int ctas = std::min(16, std::max(1, (n - 1 + 7) / 8));
  void* args[] = {(void*)&n, (void*)&Bd_, (void*)&RV_, (void*)&KS, (void*)&d, (void*)&e, (void*)&prog_};
  TEVO_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)band_chase_kernel, dim3(ctas), dim3(256), args, 0, st));
  TEVO_LAUNCHED();
}

void Sbr::apply_q2(cudaStream_t st, int n, int k, const double* Z, long ldz, const int* perm_dev, cplx* U, long ldu) {
  if (k < 1) return;
  static PerDeviceOnce once;
  once([&]() {
    TEVO_CUDA_CHECK(cudaFuncSetAttribute(q2_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  });
  const int npad = (n + 7) & ~7;
  const size_t per_col = (size_t)npad * sizeof(cplx);
  int ncmax = (int)((size_t)(227 * 1024) / per_col);
  if (ncmax < 1) throw TevoError(1, "sbr: matrix too large for the shared-memory back-transformation");
  int NC = std::min(ncmax, std::max(1, (k + num_sms_ - 1) / num_sms_));
  const int grid = (k + NC - 1) / NC;
  q2_apply_kernel<<<grid, 256, (size_t)NC * per_col, st>>>(n, k, NC, Z, ldz, perm_dev, RV_, ks_, U, ldu);
  TEVO_LAUNCHED();
}

}  // namespace tevo
