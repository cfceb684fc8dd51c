// HBM-bound helper kernels of the two-site update: gate application, gauge permutes/transposes, column
// gathers with scaling, norms, local expectation values and the on-device truncate! logic.
#include <mutex>

#include "common.cuh"
#include "kernels.h"

namespace tevo {

static inline int grid_for(long total, int block = 256, int maxblocks = 148 * 16) {
  long b = (total + block - 1) / block;
  if (b < 1) b = 1;
  if (b > maxblocks) b = maxblocks;
  return (int)b;
}

// ------------------------------------------------------------------------------------------------
// theta'(l,s1',s2',r) = sum_{s1,s2} G[(s1',s2'),(s1,s2)] theta(l,s1,s2,r)      (src/bondgate.jl:50  G*wf)
// theta is the (chil*d) x (d*chir) column-major matrix of psi[b]*psi[b+1]: element (l,s1,s2,r) at
// l + chil*s1 + (chil*d)*(s2 + d*r).  G is row-major d^2 x d^2, s1 slow.
// One thread per (l,r); l fastest -> every one of the d^2 loads/stores of a warp is a 512 B contiguous run.
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void gate_apply_kernel(cplx* __restrict__ theta, int chil, int chir, const cplx* __restrict__ G) {
  __shared__ cplx g[D * D * D * D];
  for (int i = threadIdx.x; i < D * D * D * D; i += blockDim.x) g[i] = G[i];
  __syncthreads();
  const long total = (long)chil * chir;
  const long m = (long)chil * D;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int l = (int)(e % chil);
    const long r = e / chil;
    cplx v[D * D];
#pragma unroll
    for (int s1 = 0; s1 < D; ++s1)
#pragma unroll
      for (int s2 = 0; s2 < D; ++s2) v[s1 * D + s2] = theta[l + (long)chil * s1 + m * (s2 + D * r)];
#pragma unroll
    for (int o1 = 0; o1 < D; ++o1)
#pragma unroll
      for (int o2 = 0; o2 < D; ++o2) {
        cplx acc = mk(0, 0);
#pragma unroll
        for (int c = 0; c < D * D; ++c) cfma(acc, g[(o1 * D + o2) * D * D + c], v[c]);
        theta[l + (long)chil * o1 + m * (o2 + D * r)] = acc;
      }
  }
}

void gate_apply(cudaStream_t st, cplx* theta, int chil, int d, int chir, const cplx* G_dev) {
  const long total = (long)chil * chir;
  const int grid = grid_for(total);
  switch (d) {
    case 2: gate_apply_kernel<2><<<grid, 256, 0, st>>>(theta, chil, chir, G_dev); break;
    case 3: gate_apply_kernel<3><<<grid, 256, 0, st>>>(theta, chil, chir, G_dev); break;
    case 4: gate_apply_kernel<4><<<grid, 256, 0, st>>>(theta, chil, chir, G_dev); break;
    default: throw TevoError(1, "gate_apply: physical dimension d must be 2, 3 or 4");
  }
  TEVO_LAUNCHED();
}

// ------------------------------------------------------------------------------------------------
// out(n x m, ldo) = conj(in(m x n, ldi))^T : 32x32 shared-memory tile transpose (coalesced both ways)
// ------------------------------------------------------------------------------------------------
__global__ void conj_transpose_kernel(int m, int n, const cplx* __restrict__ in, long ldi, cplx* __restrict__ out,
                                      long ldo, int do_conj) {
  __shared__ cplx tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int r = bx + threadIdx.x, c = by + j;
    if (r < m && c < n) tile[j][threadIdx.x] = in[(long)c * ldi + r];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c = by + threadIdx.x, r = bx + j;  // out(c, r) = conj(in(r, c))
    if (r < m && c < n) {
      cplx v = tile[threadIdx.x][j];
      if (do_conj) v.y = -v.y;
      out[(long)r * ldo + c] = v;
    }
  }
}

void conj_transpose(cudaStream_t st, int m, int n, const cplx* in, long ldi, cplx* out, long ldo, bool do_conj) {
  if (m <= 0 || n <= 0) return;
  dim3 grid((m + 31) / 32, (n + 31) / 32), block(32, 8);
  conj_transpose_kernel<<<grid, block, 0, st>>>(m, n, in, ldi, out, ldo, do_conj ? 1 : 0);
  TEVO_LAUNCHED();
}

__global__ void set_identity_kernel(int n, cplx* V, long ld) {
  const long total = (long)n * n;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    int i = (int)(e % n);
    long j = e / n;
    V[j * ld + i] = mk(i == j ? 1.0 : 0.0, 0.0);
  }
}
void set_identity(cudaStream_t st, int n, cplx* V, long ld) {
  set_identity_kernel<<<grid_for((long)n * n), 256, 0, st>>>(n, V, ld);
  TEVO_LAUNCHED();
}

// ------------------------------------------------------------------------------------------------
// One-sided (Hestenes) Jacobi round: CTA k handles the column pair (p,q) of the round-robin ordering,
// rotates the two columns of X (nrx rows) so that they become orthogonal and applies the same rotation
// to the two columns of V (nrv rows).  Used by the first-generation split; see svd.cu.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
jacobi_round_kernel(cplx* __restrict__ X, long ldx, int nrx, cplx* __restrict__ V, long ldv, int nrv, int nc,
                    int nceven, int round, double tol, int* __restrict__ rotcount) {
  __shared__ double red[3 * 32];
  const int k = blockIdx.x;
  const int nm1 = nceven - 1;
  int p, q;
  if (k == 0) {
    p = nm1;
    q = round;
  } else {
    p = (round + k) % nm1;
    q = (round - k + nm1) % nm1;
  }
  if (p >= nc || q >= nc) return;
  if (p > q) {
    int t = p;
    p = q;
    q = t;
  }
  cplx* xp = X + (long)p * ldx;
  cplx* xq = X + (long)q * ldx;
  double v[4] = {0, 0, 0, 0};  // a=|xp|^2, b=|xq|^2, c=<xp,xq>
  for (int i = threadIdx.x; i < nrx; i += blockDim.x) {
    cplx a = xp[i], b = xq[i];
    v[0] += cabs2(a);
    v[1] += cabs2(b);
    v[2] += a.x * b.x + a.y * b.y;
    v[3] += a.x * b.y - a.y * b.x;
  }
  {
    double w[3] = {v[0], v[1], v[2]};
    block_sum<3>(w, red);
    v[0] = w[0];
    v[1] = w[1];
    v[2] = w[2];
    double u[1] = {v[3]};
    block_sum<1>(u, red);
    v[3] = u[0];
  }
  const double a = v[0], b = v[1];
  const double cabs = sqrt(v[2] * v[2] + v[3] * v[3]);
  if (!(cabs > tol * sqrt(a * b)) || cabs == 0.0) return;
  if (threadIdx.x == 0) atomicAdd(rotcount, 1);
  const double zeta = (b - a) / (2.0 * cabs);
  const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
  const cplx ph = mk(v[2] / cabs, v[3] / cabs);  // e^{i phi}
  const cplx s1 = mk(sn * ph.x, -sn * ph.y);     // sn e^{-i phi}
  const cplx s2 = mk(sn * ph.x, sn * ph.y);      // sn e^{+i phi}
  // xp' = cs xp - s1 xq ; xq' = s2 xp + cs xq
  for (int i = threadIdx.x; i < nrx; i += blockDim.x) {
    cplx xa = xp[i], xb = xq[i];
    cplx na = csub(cscale(cs, xa), cmul(s1, xb));
    cplx nb = cadd(cmul(s2, xa), cscale(cs, xb));
    xp[i] = na;
    xq[i] = nb;
  }
  cplx* vp = V + (long)p * ldv;
  cplx* vq = V + (long)q * ldv;
  for (int i = threadIdx.x; i < nrv; i += blockDim.x) {
    cplx xa = vp[i], xb = vq[i];
    cplx na = csub(cscale(cs, xa), cmul(s1, xb));
    cplx nb = cadd(cmul(s2, xa), cscale(cs, xb));
    vp[i] = na;
    vq[i] = nb;
  }
}

void jacobi_round(cudaStream_t st, cplx* X, long ldx, int nrx, cplx* V, long ldv, int nrv, int nc, int round,
                  double tol, int* rotcount) {
  const int nceven = (nc + 1) & ~1;
  jacobi_round_kernel<<<nceven / 2, 256, 0, st>>>(X, ldx, nrx, V, ldv, nrv, nc, nceven, round, tol, rotcount);
  TEVO_LAUNCHED();
}

// column squared norms: out[j] = sum_i |X(i,j)|^2 ; one CTA per column
__global__ void colnorm2_kernel(const cplx* __restrict__ X, long ld, int nr, double* __restrict__ out) {
  __shared__ double red[32];
  const cplx* x = X + (long)blockIdx.x * ld;
  double v[1] = {0};
  for (int i = threadIdx.x; i < nr; i += blockDim.x) v[0] += cabs2(x[i]);
  block_sum<1>(v, red);
  if (threadIdx.x == 0) out[blockIdx.x] = v[0];
}
void colnorm2(cudaStream_t st, const cplx* X, long ld, int nr, int nc, double* out) {
  if (nc <= 0) return;
  colnorm2_kernel<<<nc, 128, 0, st>>>(X, ld, nr, out);
  TEVO_LAUNCHED();
}

// ------------------------------------------------------------------------------------------------
// On-device truncate! : sort sigma^2 descending (bitonic, one CTA), then the upstream rule
// (ITensors truncate!); writes permutation, sorted values and TruncInfo.  The sums over the
// spectrum (total weight, weight beyond maxdim, kept weight, entropy) are CTA-wide reductions; only the cutoff loop,
// whose length is the number of values the cutoff removes, runs on thread 0 (as sequential loops over all 2048
// values these sums took ~0.2 ms per split: ncu launch list of round 2).
// ------------------------------------------------------------------------------------------------
// sum of one double per thread over the CTA (blockDim.x a multiple of 32), returned to every thread; sm: 32 doubles
__device__ inline double block_sum_all(double v, double* sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) sm[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = (lane < (int)(blockDim.x >> 5)) ? sm[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) sm[0] = t;
  }
  __syncthreads();
  const double r = sm[0];
  __syncthreads();
  return r;
}

__device__ inline void bitonic_desc(double* key, int* idx, int npow2) {
  for (int k = 2; k <= npow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          bool desc = ((i & k) == 0);
          double a = key[i], b = key[ixj];
          int ia = idx[i], ib = idx[ixj];
          bool a_first = (a > b) || (a == b && ia < ib);  // "a should precede b" in descending order
          bool swap = desc ? !a_first : a_first;
          if (swap) {
            key[i] = b;
            key[ixj] = a;
            idx[i] = ib;
            idx[ixj] = ia;
          }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(1024)
sort_truncate_kernel(const double* __restrict__ lam, int nc, int nfull, TruncParams tp, double* __restrict__ sorted,
                     int* __restrict__ perm, TruncInfo* __restrict__ info, int npow2) {
  extern __shared__ unsigned char sm_raw[];
  double* key = reinterpret_cast<double*>(sm_raw);
  int* idx = reinterpret_cast<int*>(key + npow2);
  __shared__ int s_bad;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
    double v = (i < nc) ? lam[i] : -INFINITY;  // padding sorts below every value
    if (!(v == v)) {                           // NaN guard: flagged, sorted last
      v = -INFINITY;
      s_bad = 1;
    }
    key[i] = v;
    idx[i] = i;
  }
  __syncthreads();
  bitonic_desc(key, idx, npow2);
  for (int i = threadIdx.x; i < nc; i += blockDim.x) {
    sorted[i] = key[i];
    perm[i] = idx[i];
  }
  __syncthreads();
  __shared__ double sred[32];
  __shared__ int s_n;
  __shared__ double s_truncerr;
  const int origm = nfull;
  double loc = 0.0;
  for (int i = threadIdx.x; i < origm; i += blockDim.x) {
    double p = key[i];
    if (p < 0.0) {
      if (p == -INFINITY) s_bad = 1;
      p = 0.0;
    }
    key[i] = p;
    loc += p;
  }
  const double sum_all = block_sum_all(loc, sred);
  if (tp.inflate_rel > 0.0) {
    const double infl = tp.inflate_rel * key[0];
    __syncthreads();
    for (int i = threadIdx.x; i < origm; i += blockDim.x) key[i] += infl;
    __syncthreads();
  }
  const bool trunc_path = !(key[0] <= 0.0 || origm == 1) && tp.do_truncate;
  int maxdim = origm;
  if (trunc_path) {
    maxdim = tp.has_maxdim ? (tp.maxdim < origm ? tp.maxdim : origm) : origm;
    if (maxdim < 1) maxdim = 1;
  }
  double loc2 = 0.0;  // weight beyond maxdim
  for (int i = maxdim + threadIdx.x; i < origm; i += blockDim.x) loc2 += key[i];
  const double beyond = block_sum_all(loc2, sred);
  if (threadIdx.x == 0) {
    int n = origm;
    double truncerr = 0.0;
    if (key[0] <= 0.0 || origm == 1) {
      n = 1;
    } else if (tp.do_truncate) {
      int mindim = tp.mindim > 1 ? tp.mindim : 1;
      double cutoff = tp.cutoff > 0.0 ? tp.cutoff : 0.0;
      n = maxdim;
      truncerr = beyond;
      if (tp.use_absolute_cutoff) {
        while (key[n - 1] <= cutoff && n > mindim) {
          truncerr += key[n - 1];
          --n;
        }
      } else {
        double scale = 1.0;
        if (tp.do_rel_cutoff) {
          scale = sum_all;
          if (scale == 0.0) scale = 1.0;
        }
        while ((truncerr + key[n - 1] <= cutoff * scale) && n > mindim) {
          truncerr += key[n - 1];
          --n;
        }
        truncerr /= scale;
      }
      if (n < 1) n = 1;
    }
    s_n = n;
    s_truncerr = truncerr;
  }
  __syncthreads();
  const int n = s_n;
  double lk = 0.0, le = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double p = key[i];
    lk += p;
    if (p > 1e-13) le -= p * log(p);
  }
  const double sum_kept = block_sum_all(lk, sred);
  const double ent = block_sum_all(le, sred);
  if (threadIdx.x == 0) {
    info->nkeep = n;
    info->nfull = origm;
    info->truncerr = s_truncerr;
    info->entropy = ent;
    info->sum_all = sum_all;
    info->sum_kept = sum_kept;
    info->bad = s_bad;
    info->identity = 1;
  }
}

__global__ void __launch_bounds__(1024)
refine_truncate_kernel(const double* __restrict__ lam2, int k1, int nfull, const double* __restrict__ sum_all_dev,
                       const double* __restrict__ resid2_dev, int resid_by_difference, TruncParams tp, double* __restrict__ sorted,
                       int* __restrict__ perm, TruncInfo* __restrict__ info, int npow2) {
  extern __shared__ unsigned char sm_raw[];
  double* key = reinterpret_cast<double*>(sm_raw);
  int* idx = reinterpret_cast<int*>(key + npow2);
  __shared__ int s_bad, s_ident;
  if (threadIdx.x == 0) {
    s_bad = 0;
    s_ident = 1;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
    double v = (i < k1) ? lam2[i] : -INFINITY;
    if (!(v == v)) {
      v = -INFINITY;
      s_bad = 1;
    }
    key[i] = v;
    idx[i] = i;
  }
  __syncthreads();
  bitonic_desc(key, idx, npow2);
  for (int i = threadIdx.x; i < k1; i += blockDim.x) {
    sorted[i] = key[i];
    perm[i] = idx[i];
    if (idx[i] != i) s_ident = 0;
  }
  __syncthreads();
  __shared__ double sred[32];
  __shared__ int s_n;
  __shared__ double s_truncerr;
  const double sum_all = *sum_all_dev;
  double lk1 = 0.0;
  if (k1 < nfull && resid_by_difference)
    for (int i = threadIdx.x; i < k1; i += blockDim.x) lk1 += key[i];
  const double kept1 = block_sum_all(lk1, sred);
  if (threadIdx.x == 0) {
    double truncerr = 0.0;  // absolute weight discarded in stage 1
    if (k1 < nfull) truncerr = resid_by_difference ? fmax(sum_all - kept1, 0.0) : fmax(*resid2_dev, 0.0);
    int n = k1;
    if (tp.do_truncate && !(key[0] <= 0.0) && k1 > 1) {
      const int mindim = tp.mindim > 1 ? tp.mindim : 1;
      const double cutoff = tp.cutoff > 0.0 ? tp.cutoff : 0.0;
      if (tp.use_absolute_cutoff) {
        while (key[n - 1] <= cutoff && n > mindim) {
          truncerr += key[n - 1];
          --n;
        }
      } else {
        double scale = 1.0;
        if (tp.do_rel_cutoff) scale = (sum_all == 0.0) ? 1.0 : sum_all;
        while ((truncerr + key[n - 1] <= cutoff * scale) && n > mindim) {
          truncerr += key[n - 1];
          --n;
        }
      }
    } else if (key[0] <= 0.0) {
      n = 1;
    }
    if (!tp.use_absolute_cutoff && tp.do_truncate) {
      double scale = 1.0;
      if (tp.do_rel_cutoff) scale = (sum_all == 0.0) ? 1.0 : sum_all;
      truncerr /= scale;
    }
    if (n == nfull) truncerr = 0.0;
    s_n = n;
    s_truncerr = truncerr;
  }
  __syncthreads();
  const int n = s_n;
  double lk = 0.0, le = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double p = key[i];
    lk += p;
    if (p > 1e-13) le -= p * log(p);
  }
  const double sum_kept = block_sum_all(lk, sred);
  const double ent = block_sum_all(le, sred);
  if (threadIdx.x == 0) {
    info->nkeep = n;
    info->nfull = nfull;
    info->truncerr = s_truncerr;
    info->entropy = ent;
    info->sum_all = sum_all;
    info->sum_kept = sum_kept;
    info->bad = s_bad;
    info->identity = s_ident;
  }
}

void refine_truncate(cudaStream_t st, const double* lam2, int k1, int nfull, const double* sum_all_dev,
                     const double* resid2_dev, bool resid_by_difference, const TruncParams& tp, double* sorted, int* perm, TruncInfo* info) {
  int npow2 = 1;
  while (npow2 < k1) npow2 <<= 1;
  if (npow2 > 8192) throw TevoError(1, "refine_truncate: more than 8192 singular values not supported");
  size_t smem = (size_t)npow2 * (sizeof(double) + sizeof(int));
  static PerDeviceOnce once;
  once([&]() {
    TEVO_CUDA_CHECK(cudaFuncSetAttribute(refine_truncate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 12));
  });
  int threads = npow2 < 1024 ? (npow2 < 32 ? 32 : npow2) : 1024;
  refine_truncate_kernel<<<1, threads, smem, st>>>(lam2, k1, nfull, sum_all_dev, resid2_dev, resid_by_difference ? 1 : 0, tp,
                                                   sorted, perm, info, npow2);
  TEVO_LAUNCHED();
}

// dst(i, :) = scale * src(perm[i], :)   (k x n, row gather)
__global__ void gather_rows_kernel(int k, int n, const cplx* __restrict__ src, long lds, const int* __restrict__ perm,
                                   cplx* __restrict__ dst, long ldd, const double* __restrict__ norm2) {
  double sc = 1.0;
  if (norm2 && (*norm2 > 0.0)) sc = rsqrt(*norm2);
  const long total = (long)k * n;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int i = (int)(e % k);
    const long j = e / k;
    dst[j * ldd + i] = cscale(sc, src[j * lds + perm[i]]);
  }
}
void gather_rows(cudaStream_t st, int k, int n, const cplx* src, long lds, const int* perm, cplx* dst, long ldd,
                 const double* norm2) {
  if (k <= 0 || n <= 0) return;
  gather_rows_kernel<<<grid_for((long)k * n), 256, 0, st>>>(k, n, src, lds, perm, dst, ldd, norm2);
  TEVO_LAUNCHED();
}

void sort_truncate(cudaStream_t st, const double* lam, int nc, int nfull, const TruncParams& tp, double* sorted,
                   int* perm, TruncInfo* info) {
  int npow2 = 1;
  while (npow2 < nc) npow2 <<= 1;
  if (npow2 > 8192) throw TevoError(1, "sort_truncate: more than 8192 singular values not supported");
  size_t smem = (size_t)npow2 * (sizeof(double) + sizeof(int));
  static PerDeviceOnce once;
  once([&]() {
    TEVO_CUDA_CHECK(cudaFuncSetAttribute(sort_truncate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 12));
  });
  int threads = npow2 < 1024 ? (npow2 < 32 ? 32 : npow2) : 1024;
  sort_truncate_kernel<<<1, threads, smem, st>>>(lam, nc, nfull, tp, sorted, perm, info, npow2);
  TEVO_LAUNCHED();
}

// ------------------------------------------------------------------------------------------------
// gathers: dst(:, i) = scale * src(:, perm[i])          and   dst(i, :) = scale * conj(src(:, perm[i]))^T
// scale is read from device memory as 1/sqrt(*norm2) when norm2 != nullptr
// ------------------------------------------------------------------------------------------------
__global__ void gather_cols_kernel(int nr, int k, const cplx* __restrict__ src, long lds, const int* __restrict__ perm,
                                   cplx* __restrict__ dst, long ldd, const double* __restrict__ norm2) {
  const double s = norm2 ? rsqrt(*norm2) : 1.0;
  const double sc = (norm2 && !(*norm2 > 0.0)) ? 1.0 : s;
  const long total = (long)nr * k;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    int r = (int)(e % nr);
    int i = (int)(e / nr);
    cplx v = src[(long)perm[i] * lds + r];
    dst[(long)i * ldd + r] = cscale(sc, v);
  }
}
void gather_cols(cudaStream_t st, int nr, int k, const cplx* src, long lds, const int* perm, cplx* dst, long ldd,
                 const double* norm2) {
  if (nr <= 0 || k <= 0) return;
  gather_cols_kernel<<<grid_for((long)nr * k), 256, 0, st>>>(nr, k, src, lds, perm, dst, ldd, norm2);
  TEVO_LAUNCHED();
}

__global__ void gather_cols_conjT_kernel(int nr, int k, const cplx* __restrict__ src, long lds,
                                         const int* __restrict__ perm, cplx* __restrict__ dst, long ldd,
                                         const double* __restrict__ norm2) {
  // dst is k x nr (ldd >= k): dst(i, r) = conj(src(r, perm[i])).  32x32 tiles through shared memory.
  __shared__ cplx tile[32][33];
  double sc = 1.0;
  if (norm2 && (*norm2 > 0.0)) sc = rsqrt(*norm2);
  const int r0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int r = r0 + threadIdx.x, i = i0 + j;
    if (r < nr && i < k) tile[j][threadIdx.x] = src[(long)perm[i] * lds + r];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int i = i0 + threadIdx.x, r = r0 + j;
    if (r < nr && i < k) {
      cplx v = tile[threadIdx.x][j];
      dst[(long)r * ldd + i] = mk(sc * v.x, -sc * v.y);
    }
  }
}
void gather_cols_conjT(cudaStream_t st, int nr, int k, const cplx* src, long lds, const int* perm, cplx* dst,
                       long ldd, const double* norm2) {
  if (nr <= 0 || k <= 0) return;
  dim3 grid((nr + 31) / 32, (k + 31) / 32), block(32, 8);
  gather_cols_conjT_kernel<<<grid, block, 0, st>>>(nr, k, src, lds, perm, dst, ldd, norm2);
  TEVO_LAUNCHED();
}

// ------------------------------------------------------------------------------------------------
// vector helpers
// ------------------------------------------------------------------------------------------------
__global__ void zscal_kernel(long n, cplx a, cplx* x) {
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x)
    x[e] = cmul(a, x[e]);
}
void zscal(cudaStream_t st, long n, cplx a, cplx* x) {
  if (n <= 0) return;
  zscal_kernel<<<grid_for(n), 256, 0, st>>>(n, a, x);
  TEVO_LAUNCHED();
}

// x *= 1/sqrt(*norm2) (device scalar)
__global__ void zscal_rsqrt_kernel(long n, const double* __restrict__ norm2, cplx* x) {
  double v = *norm2;
  double s = v > 0.0 ? rsqrt(v) : 1.0;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x)
    x[e] = cscale(s, x[e]);
}
void zscal_rsqrt(cudaStream_t st, long n, const double* norm2_dev, cplx* x) {
  if (n <= 0) return;
  zscal_rsqrt_kernel<<<grid_for(n), 256, 0, st>>>(n, norm2_dev, x);
  TEVO_LAUNCHED();
}

// out[0..1] += sum conj(x) y  (out must be zeroed); grid-stride + block reduce + atomics
__global__ void zdotc_kernel(long n, const cplx* __restrict__ x, const cplx* __restrict__ y, double* out) {
  __shared__ double red[2 * 32];
  double v[2] = {0, 0};
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
    cplx a = x[e], b = y[e];
    v[0] += a.x * b.x + a.y * b.y;
    v[1] += a.x * b.y - a.y * b.x;
  }
  block_sum<2>(v, red);
  if (threadIdx.x == 0) {
    atomicAdd(out, v[0]);
    atomicAdd(out + 1, v[1]);
  }
}
void zdotc(cudaStream_t st, long n, const cplx* x, const cplx* y, double* out2_dev) {
  TEVO_CUDA_CHECK(cudaMemsetAsync(out2_dev, 0, 2 * sizeof(double), st));
  if (n <= 0) return;
  zdotc_kernel<<<grid_for(n, 256, 148 * 4), 256, 0, st>>>(n, x, y, out2_dev);
  TEVO_LAUNCHED();
}

// ------------------------------------------------------------------------------------------------
// Local expectation values from a two-site wavefunction (src/callback.jl:65-73):
//   out[2*o+0] = <theta| O_o on site b |theta>,  out[2*o+1] = <theta| O_o on site b+1 |theta>   (complex each)
// ops: nops row-major d x d matrices O[s', s].  out: 2*nops complex, must be zeroed.
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void expect2_kernel(const cplx* __restrict__ theta, int chil, int chir, int nops,
                               const cplx* __restrict__ ops, double* __restrict__ out) {
  extern __shared__ double red_dyn[];  // 32 * 2 doubles
  const long total = (long)chil * chir;
  const long m = (long)chil * D;
  for (int o = 0; o < nops; ++o) {
    const cplx* O = ops + (long)o * D * D;
    for (int which = 0; which < 2; ++which) {
      double v[2] = {0, 0};
      for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int l = (int)(e % chil);
        const long r = e / chil;
        cplx t[D * D];
#pragma unroll
        for (int s1 = 0; s1 < D; ++s1)
#pragma unroll
          for (int s2 = 0; s2 < D; ++s2) t[s1 * D + s2] = theta[l + (long)chil * s1 + m * (s2 + D * r)];
        cplx acc = mk(0, 0);
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
          for (int b = 0; b < D; ++b)
#pragma unroll
            for (int c = 0; c < D; ++c) {
              // which==0: <t[a,c]| O[a,b] |t[b,c]> ; which==1: <t[c,a]| O[a,b] |t[c,b]>
              cplx bra = which == 0 ? t[a * D + c] : t[c * D + a];
              cplx ket = which == 0 ? t[b * D + c] : t[c * D + b];
              cplx ok = cmul(O[a * D + b], ket);
              cfma_conj(acc, bra, ok);
            }
        v[0] += acc.x;
        v[1] += acc.y;
      }
      block_sum<2>(v, red_dyn);
      if (threadIdx.x == 0) {
        atomicAdd(out + 2 * (2 * o + which), v[0]);
        atomicAdd(out + 2 * (2 * o + which) + 1, v[1]);
      }
      __syncthreads();
    }
  }
}

void expect2(cudaStream_t st, const cplx* theta, int chil, int d, int chir, int nops, const cplx* ops_dev,
             double* out_dev) {
  TEVO_CUDA_CHECK(cudaMemsetAsync(out_dev, 0, sizeof(double) * 4 * nops, st));
  const long total = (long)chil * chir;
  const int grid = grid_for(total, 256, 148 * 2);
  const size_t sm = 64 * sizeof(double);
  switch (d) {
    case 2: expect2_kernel<2><<<grid, 256, sm, st>>>(theta, chil, chir, nops, ops_dev, out_dev); break;
    case 3: expect2_kernel<3><<<grid, 256, sm, st>>>(theta, chil, chir, nops, ops_dev, out_dev); break;
    case 4: expect2_kernel<4><<<grid, 256, sm, st>>>(theta, chil, chir, nops, ops_dev, out_dev); break;
    default: throw TevoError(1, "expect2: physical dimension d must be 2, 3 or 4");
  }
  TEVO_LAUNCHED();
}

// ------------------------------------------------------------------------------------------------
// Middle-index contraction used by H_eff and environment updates:
//   out[a + oq[q] + oc*c] = sum_p Mx[q*P + p] * in[a + na*(p + P*c)]      a<na, p<P, q<Q, c<nc
// (a is the fast, coalesced index; P,Q <= 64.)  Mx row-major Q x P in global memory; oq: Q offsets.
// ------------------------------------------------------------------------------------------------
__global__ void mid_contract_kernel(int na, int P, int Q, int nc, const cplx* __restrict__ Mx,
                                    const long* __restrict__ oq, long oc, const cplx* __restrict__ in,
                                    cplx* __restrict__ out) {
  extern __shared__ unsigned char smraw[];
  cplx* sM = reinterpret_cast<cplx*>(smraw);
  long* sO = reinterpret_cast<long*>(sM + (long)P * Q);
  for (int i = threadIdx.x; i < P * Q; i += blockDim.x) sM[i] = Mx[i];
  for (int i = threadIdx.x; i < Q; i += blockDim.x) sO[i] = oq[i];
  __syncthreads();
  const long total = (long)na * nc;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int a = (int)(e % na);
    const long c = e / na;
    const cplx* src = in + a + (long)na * P * c;
    cplx* dst = out + a + oc * c;
    for (int q = 0; q < Q; ++q) {
      cplx acc = mk(0, 0);
      for (int p = 0; p < P; ++p) {
        cplx m = sM[q * P + p];
        if (m.x != 0.0 || m.y != 0.0) cfma(acc, m, src[(long)na * p]);
      }
      dst[sO[q]] = acc;
    }
  }
}

void mid_contract(cudaStream_t st, int na, int P, int Q, int nc, const cplx* Mx_dev, const long* oq_dev, long oc,
                  const cplx* in, cplx* out) {
  const long total = (long)na * nc;
  if (total <= 0) return;
  size_t sm = (size_t)P * Q * sizeof(cplx) + (size_t)Q * sizeof(long);
  if (sm > 48 * 1024) throw TevoError(1, "mid_contract: operator block too large");
  mid_contract_kernel<<<grid_for(total), 256, sm, st>>>(na, P, Q, nc, Mx_dev, oq_dev, oc, in, out);
  TEVO_LAUNCHED();
}

// ------------------------------------------------------------------------------------------------
// Krylov BLAS-1: multi-dot h[i] = <V_i, r> (i<k) and r -= sum_i h[i] V_i ; V is n x k column-major (ld = n)
// ------------------------------------------------------------------------------------------------
__global__ void multi_dotc_kernel(long n, int k, const cplx* __restrict__ V, long ldv, const cplx* __restrict__ r,
                                  double* __restrict__ out /* 2k, zeroed */) {
  __shared__ double red[2 * 32];
  const int i = blockIdx.y;
  const cplx* v = V + (long)i * ldv;
  double acc[2] = {0, 0};
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
    cplx a = v[e], b = r[e];
    acc[0] += a.x * b.x + a.y * b.y;
    acc[1] += a.x * b.y - a.y * b.x;
  }
  block_sum<2>(acc, red);
  if (threadIdx.x == 0) {
    atomicAdd(out + 2 * i, acc[0]);
    atomicAdd(out + 2 * i + 1, acc[1]);
  }
}
void multi_dotc(cudaStream_t st, long n, int k, const cplx* V, long ldv, const cplx* r, double* out_dev) {
  TEVO_CUDA_CHECK(cudaMemsetAsync(out_dev, 0, sizeof(double) * 2 * k, st));
  if (n <= 0 || k <= 0) return;
  int gx = grid_for(n, 256, 148);
  dim3 grid(gx, k);
  multi_dotc_kernel<<<grid, 256, 0, st>>>(n, k, V, ldv, r, out_dev);
  TEVO_LAUNCHED();
}

// r -= sum_i h[i] V_i   (h device, complex as 2 doubles);  hacc[i] += h[i] if hacc != nullptr
__global__ void multi_axpy_sub_kernel(long n, int k, const cplx* __restrict__ V, long ldv,
                                      const double* __restrict__ h, cplx* __restrict__ r) {
  extern __shared__ double sh[];
  for (int i = threadIdx.x; i < 2 * k; i += blockDim.x) sh[i] = h[i];
  __syncthreads();
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
    cplx acc = r[e];
    for (int i = 0; i < k; ++i) {
      cplx v = V[(long)i * ldv + e];
      acc.x -= sh[2 * i] * v.x - sh[2 * i + 1] * v.y;
      acc.y -= sh[2 * i] * v.y + sh[2 * i + 1] * v.x;
    }
    r[e] = acc;
  }
}
void multi_axpy_sub(cudaStream_t st, long n, int k, const cplx* V, long ldv, const double* h_dev, cplx* r) {
  if (n <= 0 || k <= 0) return;
  multi_axpy_sub_kernel<<<grid_for(n), 256, sizeof(double) * 2 * k, st>>>(n, k, V, ldv, h_dev, r);
  TEVO_LAUNCHED();
}

// out = sum_i y[i] V_i  (y device complex), scaled by real factor
__global__ void lincomb_kernel(long n, int k, const cplx* __restrict__ V, long ldv, const double* __restrict__ y,
                               double scale, cplx* __restrict__ out) {
  extern __shared__ double sh[];
  for (int i = threadIdx.x; i < 2 * k; i += blockDim.x) sh[i] = y[i];
  __syncthreads();
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
    cplx acc = mk(0, 0);
    for (int i = 0; i < k; ++i) {
      cplx v = V[(long)i * ldv + e];
      acc.x += sh[2 * i] * v.x - sh[2 * i + 1] * v.y;
      acc.y += sh[2 * i] * v.y + sh[2 * i + 1] * v.x;
    }
    out[e] = cscale(scale, acc);
  }
}
void lincomb(cudaStream_t st, long n, int k, const cplx* V, long ldv, const double* y_dev, double scale, cplx* out) {
  if (n <= 0) return;
  lincomb_kernel<<<grid_for(n), 256, sizeof(double) * 2 * k, st>>>(n, k, V, ldv, y_dev, scale, out);
  TEVO_LAUNCHED();
}

// out = s * in  (real scale), separate buffers
__global__ void copy_scale_kernel(long n, double s, const cplx* __restrict__ in, cplx* __restrict__ out) {
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x)
    out[e] = cscale(s, in[e]);
}
void copy_scale(cudaStream_t st, long n, double s, const cplx* in, cplx* out) {
  if (n <= 0) return;
  copy_scale_kernel<<<grid_for(n), 256, 0, st>>>(n, s, in, out);
  TEVO_LAUNCHED();
}


// ------------------------------------------------------------------------------------------------
// B-form (layer-parallel TEBD) helpers
// ------------------------------------------------------------------------------------------------
// out(r, c) = lam[r % chil] * in(r, c)     (rows are (l, s) with l fastest)
__global__ void scale_rows_kernel(int m, long n, int chil, const double* __restrict__ lam, const cplx* __restrict__ in,
                                  cplx* __restrict__ out) {
  const long total = (long)m * n;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int r = (int)(e % m);
    out[e] = cscale(lam[r % chil], in[e]);
  }
}
void scale_rows(cudaStream_t st, int m, long n, int chil, const double* lam, const cplx* in, cplx* out) {
  if (m <= 0 || n <= 0) return;
  scale_rows_kernel<<<grid_for((long)m * n), 256, 0, st>>>(m, n, chil, lam, in, out);
  TEVO_LAUNCHED();
}

// out[i] = sqrt(max(in[i], 0) / (*denom))   (denom == nullptr: 1)
__global__ void sqrt_scaled_kernel(int n, const double* __restrict__ in, const double* __restrict__ denom,
                                   double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double dn = denom ? *denom : 1.0;
  if (!(dn > 0.0)) dn = 1.0;
  out[i] = sqrt(fmax(in[i], 0.0) / dn);
}
void sqrt_scaled(cudaStream_t st, int n, const double* in, const double* denom, double* out) {
  if (n <= 0) return;
  sqrt_scaled_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, in, denom, out);
  TEVO_LAUNCHED();
}

// <T|O_o|T> with T(l,s,r) = lam[l] * B(l,s,r); out: nops complex (zeroed here)
template <int D>
__global__ void expect1_kernel(const cplx* __restrict__ B, int chil, int chir, const double* __restrict__ lam, int nops,
                               const cplx* __restrict__ ops, double* __restrict__ out) {
  __shared__ double red[2 * 32];
  const long total = (long)chil * chir;
  for (int o = 0; o < nops; ++o) {
    const cplx* O = ops + (long)o * D * D;
    double v[2] = {0, 0};
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
      const int l = (int)(e % chil);
      const long r = e / chil;
      const double w = lam[l] * lam[l];
      cplx t[D];
#pragma unroll
      for (int s = 0; s < D; ++s) t[s] = B[l + (long)chil * (s + D * r)];
      cplx acc = mk(0, 0);
#pragma unroll
      for (int a = 0; a < D; ++a)
#pragma unroll
        for (int b = 0; b < D; ++b) cfma_conj(acc, t[a], cmul(O[a * D + b], t[b]));
      v[0] += w * acc.x;
      v[1] += w * acc.y;
    }
    block_sum<2>(v, red);
    if (threadIdx.x == 0) {
      atomicAdd(out + 2 * o, v[0]);
      atomicAdd(out + 2 * o + 1, v[1]);
    }
    __syncthreads();
  }
}
void expect1(cudaStream_t st, const cplx* B, int chil, int d, int chir, const double* lam, int nops, const cplx* ops_dev,
             double* out_dev) {
  TEVO_CUDA_CHECK(cudaMemsetAsync(out_dev, 0, sizeof(double) * 2 * nops, st));
  const int grid = grid_for((long)chil * chir, 256, 148 * 2);
  switch (d) {
    case 2: expect1_kernel<2><<<grid, 256, 0, st>>>(B, chil, chir, lam, nops, ops_dev, out_dev); break;
    case 3: expect1_kernel<3><<<grid, 256, 0, st>>>(B, chil, chir, lam, nops, ops_dev, out_dev); break;
    case 4: expect1_kernel<4><<<grid, 256, 0, st>>>(B, chil, chir, lam, nops, ops_dev, out_dev); break;
    default: throw TevoError(1, "expect1: physical dimension d must be 2, 3 or 4");
  }
  TEVO_LAUNCHED();
}

}  // namespace tevo
