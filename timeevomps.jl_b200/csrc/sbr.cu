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
  // |(alpha; x)| and its reciprocal from one reciprocal square root (+ a Newton correction each), 1/|alpha - beta|^2
  // from one reciprocal: dependent FP64 operations cost ~45 cycles each on this part and this chain sits on the
  // critical path of every reflector (same sequence as the one-stage panel kernel, heig.cu)
  double rs = rsqrt(s2);
  double nx = s2 * rs;
  nx = fma(0.5 * rs, fma(-nx, nx, s2), nx);
  rs = fma(rs, fma(-nx, rs, 1.0), rs);
  beta = -copysign(nx, alpha.x);
  const double ib = -copysign(rs, alpha.x);
  tau = mk((beta - alpha.x) * ib, -alpha.y * ib);
  const cplx den = mk(alpha.x - beta, alpha.y);
  const double rdn = __drcp_rn(cabs2(den));
  scal = mk(den.x * rdn, -den.y * rdn);
}

// ------------------------------------------------------------------------------------------------------------------
// stage 1: dense -> band (width 8).  Persistent cooperative kernel, 512 threads per CTA, one CTA per SM.
// ------------------------------------------------------------------------------------------------------------------
#ifdef TEVO_SBR_TIMING
__device__ unsigned long long g_sbr_cycles[16];
#define ST_MARK(slot)                                                    \
  do {                                                                   \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                           \
      long long _now = clock64();                                        \
      atomicAdd(&g_sbr_cycles[slot], (unsigned long long)(_now - _t0));  \
      _t0 = _now;                                                        \
    }                                                                    \
  } while (0)
#else
#define ST_MARK(slot)
#endif

struct S1Args {
  int n;
  cplx* A;
  long lda;
  cplx* Bd;       // band out (lower storage, LDB diagonals per column)
  cplx* taus;     // n
  cplx* Vb;       // [2][bstride]: reflectors of the block, row-major (row r: 8 entries), rows >= j+8 valid
  cplx* Wb;       // [2][bstride]: W of the block
  double* Yb;     // [2][2*bstride]: Y = A V accumulators (re, im)
  double* small;  // [2][256]: S = V^H Y accumulator (8x8 complex) followed by the T factor (8x8 complex)
  unsigned* bar;  // grid barrier counter (zeroed before the launch)
  long bstride;   // (n + 64) * 8
};

__device__ inline void grid_bar(unsigned* ctr, unsigned epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    const unsigned target = epoch * gridDim.x;
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

constexpr int S1_NT = 512;          // threads per CTA of the stage-1 kernel
constexpr int S1_NW = S1_NT / 32;   // warps
constexpr int S1_RPT = 4;           // panel rows per thread in the QR (SBR_MAX_N - 8 <= S1_NT * S1_RPT)

// sum of NV doubles over the CTA; result in all threads.  red: >= NV*S1_NW doubles of shared memory
template <int NV>
__device__ inline void cta_sum(double (&v)[NV], double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[i * S1_NW + wid] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    // pairwise tree (4 dependent additions instead of 16): dependent FP64 operations are what this kernel waits for
    double x[S1_NW];
#pragma unroll
    for (int w = 0; w < S1_NW; ++w) x[w] = red[i * S1_NW + w];
#pragma unroll
    for (int st = S1_NW / 2; st > 0; st >>= 1)
#pragma unroll
      for (int w = 0; w < st; ++w) x[w] += x[w + st];
    v[i] = x[0];
  }
}

// operand rows of one 64-column tile in shared memory: entry (c, q) at c*8 + q + (c >> 4): the four column groups of a
// warp (16 columns apart) fall into different banks
constexpr int OPS = 64 * 8 + 4;
__device__ inline int opi(int c, int q) { return c * 8 + q + (c >> 4); }

// Householder QR of the sub-band panel of block bi (columns j .. j+7, rows j+8 .. n-1) by one CTA.  Thread t owns the
// panel rows t, t+512, ...: columns 0..3 of its rows live in registers, columns 4..7 in shared memory (sP, private to
// the thread, so no synchronisation is needed for it).  Writes the band entries of the 8 columns, the reflectors
// (explicit form) to Vb[p] and into A, their scalars, and the compact-WY T factor.
constexpr int QR_MP = SBR_MAX_N;  // row stride of the shared-memory half of the panel
#define PGET(u, c) (((c) < 4) ? PR[u][(c)&3] : sP[((c)-4) * QR_MP + t + S1_NT * (u)])
#define PSET(u, c, val)                              \
  do {                                               \
    if ((c) < 4)                                     \
      PR[u][(c)&3] = (val);                          \
    else                                             \
      sP[((c)-4) * QR_MP + t + S1_NT * (u)] = (val); \
  } while (0)

// one column of the panel QR (IC is a compile-time constant so that every access to PR has a static index and the
// register half of the panel really stays in registers)
template <int IC>
__device__ __forceinline__ void qr_column(cplx (&PR)[S1_RPT][4], cplx* sP, int t, int bw, cplx (&tauv)[SB], double (&betav)[SB],
                                          double* red, cplx* sbc, cplx* sTq) {
  tauv[IC] = mk(0, 0);
  betav[IC] = 0.0;
  if (IC >= bw) return;
  double part[1] = {0.0};
#pragma unroll
  for (int u = 0; u < S1_RPT; ++u) {
    const int ir = t + S1_NT * u;
    if (ir > IC) part[0] += cabs2(PGET(u, IC));
  }
  if (t == IC) sbc[0] = PGET(0, IC);
  cta_sum<1>(part, red);
  const cplx alpha = sbc[0];
  double beta;
  cplx tau, scal;
  larfg(alpha, part[0], beta, tau, scal);
  tauv[IC] = tau;
  betav[IC] = beta;
  cplx vcol[S1_RPT];  // the reflector entries of the own rows (0 above the unit, 1 at the unit)
#pragma unroll
  for (int u = 0; u < S1_RPT; ++u) {
    const int ir = t + S1_NT * u;
    cplx x = PGET(u, IC);
    if (ir > IC) {
      x = cmul(x, scal);
      PSET(u, IC, x);
    } else if (ir == IC) {
      x = mk(1, 0);
      PSET(u, IC, x);
    }
    vcol[u] = (ir >= IC) ? x : mk(0, 0);
  }
  // one reduction of 7 complex numbers: slots c-1 (c > IC): w_c = v^H P(:, c) for the panel update;
  // slots pc (pc < IC): G(pc, IC) = v_pc^H v_IC for the T factor
  double ws[2 * (SB - 1)];
#pragma unroll
  for (int x = 0; x < 2 * (SB - 1); ++x) ws[x] = 0.0;
#pragma unroll
  for (int u = 0; u < S1_RPT; ++u) {
    const cplx v = vcol[u];
#pragma unroll
    for (int c = IC + 1; c < SB; ++c) {
      const cplx x = PGET(u, c);
      ws[2 * (c - 1)] += v.x * x.x + v.y * x.y;
      ws[2 * (c - 1) + 1] += v.x * x.y - v.y * x.x;
    }
#pragma unroll
    for (int pc = 0; pc < IC; ++pc) {
      const cplx x = PGET(u, pc);  // rows >= IC > pc lie below the unit of column pc; v is 0 on the others
      ws[2 * pc] += x.x * v.x + x.y * v.y;
      ws[2 * pc + 1] += x.x * v.y - x.y * v.x;
    }
  }
  cta_sum<2 * (SB - 1)>(ws, red);
  const cplx ct = cconj(tau);
#pragma unroll
  for (int u = 0; u < S1_RPT; ++u) {
    const cplx tv = cmul(ct, vcol[u]);
#pragma unroll
    for (int c = IC + 1; c < SB; ++c) {
      cplx x = PGET(u, c);
      cfms(x, tv, mk(ws[2 * (c - 1)], ws[2 * (c - 1) + 1]));
      PSET(u, c, x);
    }
  }
  if (t < IC) {
    cplx acc = mk(0, 0);
#pragma unroll
    for (int l = 0; l < IC; ++l)
      if (l >= t) cfma(acc, sTq[t + SB * l], mk(ws[2 * l], ws[2 * l + 1]));
    sTq[t + SB * IC] = cmul(mk(-tau.x, -tau.y), acc);
  } else if (t == IC) {
    sTq[t + SB * IC] = tau;
  }
}

__device__ __forceinline__ void panel_qr(const S1Args& a, int j, int p, double* red, cplx* sbc, cplx* sTq, cplx* sP) {
  const int n = a.n, t = threadIdx.x;
  const int m = n - j - SB;
  const int bw = min(SB, m - 1);
#ifdef TEVO_SBR_TIMING
  long long _t0 = clock64();
#endif
  cplx PR[S1_RPT][4];
#pragma unroll
  for (int u = 0; u < S1_RPT; ++u) {
    const int ir = t + S1_NT * u;
#pragma unroll
    for (int c = 0; c < SB; ++c) {
      const cplx x = (ir < m) ? ldcg_c(a.A + (long)(j + c) * a.lda + (j + SB + ir)) : mk(0, 0);
      PSET(u, c, x);
    }
  }
  __syncthreads();
  ST_MARK(8);
  if (t < 64) {  // diagonal block -> band
    const int rr = t & 7, cc = t >> 3;
    if (rr >= cc) {
      cplx x = ldcg_c(a.A + (long)(j + cc) * a.lda + (j + rr));
      if (rr == cc) x.y = 0.0;
      a.Bd[(long)LDB * (j + cc) + (rr - cc)] = x;
    }
  }
  cplx tauv[SB];
  double betav[SB];
  // sTq: T factor, row pr written by thread pr only (T(pr, q) = -tau_q sum_l T(pr, l) G(l, q), G = V^H V)
  if (t < 64) sTq[t] = mk(0, 0);
  qr_column<0>(PR, sP, t, bw, tauv, betav, red, sbc, sTq);
  qr_column<1>(PR, sP, t, bw, tauv, betav, red, sbc, sTq);
  qr_column<2>(PR, sP, t, bw, tauv, betav, red, sbc, sTq);
  qr_column<3>(PR, sP, t, bw, tauv, betav, red, sbc, sTq);
  qr_column<4>(PR, sP, t, bw, tauv, betav, red, sbc, sTq);
  qr_column<5>(PR, sP, t, bw, tauv, betav, red, sbc, sTq);
  qr_column<6>(PR, sP, t, bw, tauv, betav, red, sbc, sTq);
  qr_column<7>(PR, sP, t, bw, tauv, betav, red, sbc, sTq);
  __syncthreads();
  ST_MARK(9);
  if (t < 64) reinterpret_cast<cplx*>(a.small + 256 * p + 128)[t] = sTq[t];
  if (t < SB) {
    cplx tv = mk(0, 0);
#pragma unroll
    for (int c = 0; c < SB; ++c)
      if (t == c) tv = tauv[c];
    a.taus[j + t] = tv;
  }
  ST_MARK(10);
  // band entries of the R factor (rows j+8+ir, ir <= c) and explicit reflectors
  cplx* Vp = a.Vb + (long)p * a.bstride;
#pragma unroll
  for (int u = 0; u < S1_RPT; ++u) {
    const int ir = t + S1_NT * u;
    if (ir < m) {
      cplx vrow[SB];
#pragma unroll
      for (int c = 0; c < SB; ++c) {
        const cplx pv = PGET(u, c);
        if (ir <= c) {
          cplx x = pv;
          if (ir == c && c < bw) x = mk(betav[c], 0);
          a.Bd[(long)LDB * (j + c) + (SB + ir - c)] = x;
        }
        cplx v = mk(0, 0);
        if (c < bw) v = (ir < c) ? mk(0, 0) : ((ir == c) ? mk(1, 0) : pv);
        vrow[c] = v;
        a.A[(long)(j + c) * a.lda + (j + SB + ir)] = v;
      }
#pragma unroll
      for (int c = 0; c < SB; ++c) Vp[(long)(j + SB + ir) * SB + c] = vrow[c];
    }
  }
  __syncthreads();
  ST_MARK(11);
}
#undef PGET
#undef PSET

__global__ void __launch_bounds__(S1_NT, 1) sbr_stage1_kernel(S1Args a) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];  // panel QR: columns 4..7 of the panel (4 x QR_MP complex)
  __shared__ double red[14 * S1_NW];
  __shared__ cplx sbc[4];
  __shared__ cplx sT[64], sS[64], sX[64], sM[64];
  __shared__ cplx sWc[9 * 8], sVc[9 * 8];
  __shared__ cplx sOp[3][OPS];
  cplx(*sSp)[64] = reinterpret_cast<cplx(*)[64]>(dyn_smem);  // pass only (the panel QR is done by then)
  const int n = a.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x;
  const int NB = (n >= SB + 2) ? (n - 2 - SB) / SB + 1 : 0;  // blocks j = 8 bi with j + 8 <= n - 2
  unsigned epoch = 0;
#ifdef TEVO_SBR_TIMING
  long long _t0 = clock64();
#endif
  for (int bi = 0; bi <= NB; ++bi) {
    const int j = SB * bi;
    const int p = bi & 1, pp = p ^ 1;
    const int ncol = (bi < NB) ? SB : n - j;  // <= 9
    // ---------------- (1) W of block bi-1, its update applied to columns j .. j+ncol-1 (rows j .. n-1) ----------------
    if (bi > 0) {
      const cplx* Vq = a.Vb + (long)pp * a.bstride;
      cplx* Wq = a.Wb + (long)pp * a.bstride;
      const cplx* Yq = reinterpret_cast<const cplx*>(a.Yb + 2 * (long)pp * a.bstride);
      const cplx* Sg = reinterpret_cast<const cplx*>(a.small + 256 * pp);
      const cplx* Tg = Sg + 64;
      if (tid < 64) {
        sS[tid] = ldcg_c(Sg + tid);
        sT[tid] = ldcg_c(Tg + tid);
      }
      __syncthreads();
      if (tid < 64) {  // X = S T
        const int x = tid & 7, y = tid >> 3;
        cplx acc = mk(0, 0);
#pragma unroll
        for (int l = 0; l < SB; ++l) cfma(acc, sS[x + SB * l], sT[l + SB * y]);
        sX[tid] = acc;
      }
      __syncthreads();
      if (tid < 64) {  // M = 1/2 T^H X
        const int x = tid & 7, y = tid >> 3;
        cplx acc = mk(0, 0);
#pragma unroll
        for (int l = 0; l < SB; ++l) cfma_conj(acc, sT[l + SB * x], sX[l + SB * y]);
        sM[tid] = cscale(0.5, acc);
      }
      __syncthreads();
      if (tid < ncol * SB) {  // W and V of the rows j .. j+ncol-1 (the columns being updated)
        const int cc = tid >> 3, q = tid & 7;
        const long r = j + cc;
        cplx acc = mk(0, 0);
#pragma unroll
        for (int l = 0; l < SB; ++l) {
          cfma(acc, ldcg_c(Yq + r * SB + l), sT[l + SB * q]);
          cfms(acc, ldcg_c(Vq + r * SB + l), sM[l + SB * q]);
        }
        sWc[tid] = acc;
        sVc[tid] = ldcg_c(Vq + r * SB + q);
      }
      __syncthreads();
      const int q = lane & 7, sub = lane >> 3;
      for (long rb = j + (long)blockIdx.x * (4 * S1_NW) + warp * 4; rb < n; rb += (long)(4 * S1_NW) * G) {
        // the loop bound is warp-uniform (the shuffles below need all lanes); rows past the end are inactive
        const long r = rb + sub;
        const bool act = r < n;
        const cplx yq = act ? ldcg_c(Yq + r * SB + q) : mk(0, 0);
        const cplx vq = act ? ldcg_c(Vq + r * SB + q) : mk(0, 0);
        cplx yv[SB], vv[SB];
#pragma unroll
        for (int l = 0; l < SB; ++l) {
          yv[l] = shfl_c(yq, (lane & 24) + l);
          vv[l] = shfl_c(vq, (lane & 24) + l);
        }
        cplx wq = mk(0, 0);
#pragma unroll
        for (int l = 0; l < SB; ++l) {
          cfma(wq, yv[l], sT[l + SB * q]);
          cfms(wq, vv[l], sM[l + SB * q]);
        }
        if (act) Wq[r * SB + q] = wq;
        cplx wv[SB];
#pragma unroll
        for (int l = 0; l < SB; ++l) wv[l] = shfl_c(wq, (lane & 24) + l);
        for (int c = q; c < ncol; c += SB) {
          if (act) {
            cplx* ap = a.A + (long)(j + c) * a.lda + r;
            cplx x = ldcg_c(ap);
#pragma unroll
            for (int l = 0; l < SB; ++l) {
              cfms(x, vv[l], cconj(sWc[c * SB + l]));
              cfms(x, wv[l], cconj(sVc[c * SB + l]));
            }
            *ap = x;
          }
        }
      }
    }
    // accumulators of block bi
    {
      double* Yz = a.Yb + 2 * (long)p * a.bstride;
      for (long e = (long)blockIdx.x * S1_NT + tid; e < 2L * (n + 8) * SB; e += (long)S1_NT * G) Yz[e] = 0.0;
      if (blockIdx.x == 0 && tid < 128) a.small[256 * p + tid] = 0.0;
    }
    ST_MARK(0);
    grid_bar(a.bar, ++epoch);
    ST_MARK(1);
    if (bi == NB) {
      // remaining columns j .. n-1: all inside the band
      for (int e = blockIdx.x * S1_NT + tid; e < ncol * ncol; e += S1_NT * G) {
        const int rr = e % ncol, cc = e / ncol;
        if (rr >= cc) {
          cplx x = ldcg_c(a.A + (long)(j + cc) * a.lda + (j + rr));
          if (rr == cc) x.y = 0.0;
          a.Bd[(long)LDB * (j + cc) + (rr - cc)] = x;
        }
      }
      if (blockIdx.x == 0 && tid < ncol) a.taus[j + tid] = mk(0, 0);
      break;
    }
    // ---------------- (2) panel QR of block bi ----------------
    if (blockIdx.x == 0) panel_qr(a, j, p, red, sbc, sX, reinterpret_cast<cplx*>(dyn_smem));
    ST_MARK(2);
    grid_bar(a.bar, ++epoch);
    ST_MARK(3);
    // ---------------- (3) one pass over A[R, R], R = j+8 .. n-1: update of block bi-1, Y = A V, S = V^H Y ----------------
    {
      const int r_lo = j + SB, m = n - r_lo;
      constexpr int TR = 8 * S1_NW;  // tile: TR rows (8 per warp) x 64 columns
      const int nt = (m + 63) / 64;
      const int ntr = (m + TR - 1) / TR;
      const int ntiles = ntr * nt;
      // contiguous, balanced ranges of the row-major tile list (a CTA mostly stays inside one row block)
      const int tbase = ntiles / G, trem = ntiles % G;
      const int t_lo = blockIdx.x * tbase + min((int)blockIdx.x, trem);
      const int t_hi = t_lo + tbase + ((int)blockIdx.x < trem ? 1 : 0);
      const bool has_prev = bi > 0;
      const cplx* Vprev = a.Vb + (long)pp * a.bstride;
      const cplx* Wprev = a.Wb + (long)pp * a.bstride;
      const cplx* Vnew = a.Vb + (long)p * a.bstride;
      double* Yacc = a.Yb + 2 * (long)p * a.bstride;
      double* Sacc = a.small + 256 * p;
      const int rl = 8 * warp + (lane & 7);  // row inside the tile
      const int cgp = lane >> 3;             // column group: columns 16*cgp .. 16*cgp+15 of the tile
      cplx vp[SB], wp[SB], yacc[SB];
      int curI = -1;
      long r = 0;
      auto flush = [&]() {
        // reduce over the four column groups, add to Y, accumulate S = V^H Y over the rows of the warp
#pragma unroll
        for (int q = 0; q < SB; ++q) {
          yacc[q] = cadd(yacc[q], shfl_xor_c(yacc[q], 8));
          yacc[q] = cadd(yacc[q], shfl_xor_c(yacc[q], 16));
        }
        const bool own = (cgp == 0) && (r < n);
        cplx vn[SB];
#pragma unroll
        for (int q = 0; q < SB; ++q) vn[q] = own ? ldcg_c(Vnew + r * SB + q) : mk(0, 0);
        if (own) {
#pragma unroll
          for (int q = 0; q < SB; ++q) {
            atomicAdd(Yacc + 2 * (r * SB + q), yacc[q].x);
            atomicAdd(Yacc + 2 * (r * SB + q) + 1, yacc[q].y);
          }
        }
#pragma unroll
        for (int x = 0; x < SB; ++x) {
#pragma unroll
          for (int y = 0; y < SB; ++y) {
            cplx sp = mk(0, 0);
            cfma_conj(sp, vn[x], yacc[y]);
            sp = cadd(sp, shfl_xor_c(sp, 1));
            sp = cadd(sp, shfl_xor_c(sp, 2));
            sp = cadd(sp, shfl_xor_c(sp, 4));
            if (lane == 0) sSp[warp][x + SB * y] = sp;
          }
        }
        __syncthreads();
        if (tid < 64) {
          cplx acc = mk(0, 0);
#pragma unroll
          for (int w = 0; w < S1_NW; ++w) acc = cadd(acc, sSp[w][tid]);
          atomicAdd(Sacc + 2 * tid, acc.x);
          atomicAdd(Sacc + 2 * tid + 1, acc.y);
        }
        __syncthreads();
      };
      for (int tI = t_lo; tI < t_hi; ++tI) {
        const int I = tI / nt, J = tI % nt;
        if (I != curI) {
          if (curI >= 0) flush();
          curI = I;
          r = r_lo + (long)TR * I + rl;
#pragma unroll
          for (int q = 0; q < SB; ++q) {
            yacc[q] = mk(0, 0);
            vp[q] = (has_prev && r < n) ? ldcg_c(Vprev + r * SB + q) : mk(0, 0);
            wp[q] = (has_prev && r < n) ? ldcg_c(Wprev + r * SB + q) : mk(0, 0);
          }
        }
        const long c0 = r_lo + 64L * J;
        __syncthreads();  // the previous tile's operand rows are no longer read
        for (int e = tid; e < 64 * SB; e += S1_NT) {
          const int c = e >> 3, q = e & 7;
          const long cg_ = c0 + c;
          const bool in = cg_ < n;
          sOp[0][opi(c, q)] = (in && has_prev) ? ldcg_c(Wprev + cg_ * SB + q) : mk(0, 0);
          sOp[1][opi(c, q)] = (in && has_prev) ? ldcg_c(Vprev + cg_ * SB + q) : mk(0, 0);
          sOp[2][opi(c, q)] = in ? ldcg_c(Vnew + cg_ * SB + q) : mk(0, 0);
        }
        __syncthreads();
        cplx* ap = a.A + (c0 + 16 * cgp) * a.lda + r;
#pragma unroll
        for (int h = 0; h < 2; ++h) {  // the 16 columns of the thread in two batches of 8 loads in flight
          cplx av[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            av[i] = (r < n && c0 + 16 * cgp + 8 * h + i < n) ? ldcg_c(ap + (long)(8 * h + i) * a.lda) : mk(0, 0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = 16 * cgp + 8 * h + i;
            cplx x = av[i];
            if (has_prev) {
#pragma unroll
              for (int q = 0; q < SB; ++q) {
                cfms(x, vp[q], cconj(sOp[0][opi(c, q)]));
                cfms(x, wp[q], cconj(sOp[1][opi(c, q)]));
              }
              if (r < n && c0 + c < n) ap[(long)(8 * h + i) * a.lda] = x;
            }
#pragma unroll
            for (int q = 0; q < SB; ++q) cfma(yacc[q], x, sOp[2][opi(c, q)]);
          }
        }
      }
      ST_MARK(4);
      if (curI >= 0) flush();
      ST_MARK(5);
    }
    grid_bar(a.bar, ++epoch);
    ST_MARK(6);
  }
}

// explicit zeros above the reflectors: column c of A keeps rows >= c + 8 (the back-transformation reads A as V)
__global__ void sbr_zero_upper_kernel(int n, cplx* A, long lda) {
  const int c = blockIdx.x;
  const int top = (c + SB <= n - 2) ? c + SB : n;  // columns without reflector are cleared entirely
  for (int r = threadIdx.x; r < top; r += blockDim.x) A[(long)c * lda + r] = mk(0, 0);
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
            asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(prog + s - 1) : "memory");
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
      if (lane == 0) {  // the release store orders the window stores of all lanes (made visible to lane 0 by __syncwarp)
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
// stage 2, second version: the same chase with the hand-offs inside a CTA.  A CTA advances a GROUP of CH_W consecutive
// sweeps in lock-step (one warp per sweep, warp w runs step T - 2w in slot T, one named barrier per slot), with the
// part of the band the group is working on in a shared-memory ring of CH_RING columns.  A loader warp streams new
// columns in ahead of the first sweep (up to CH_L slots ahead, so the L2 latency is off the critical path) and writes
// finished columns back behind the last sweep; groups follow each other through global progress counters
// (gprog[g] = F: band columns < last_sweep(g) + 1 + 8 F are final w.r.t. group g and visible in global memory).
// ------------------------------------------------------------------------------------------------------------------
constexpr int CH_W = 16;
constexpr int CH_NT = (CH_W + 1) * 32;
constexpr int CH_RING = 512;
constexpr int CH_L = 8;   // slots the loader may run ahead of the compute warps
constexpr int CH_B = 1;   // slots per load / write-back batch (4 measured slower: 12.1 vs 10.7 ms at n = 2048)
constexpr int CH_DONE = 1 << 30;

__device__ inline cplx* ring_col(cplx* ring, int c) { return ring + (size_t)(c & (CH_RING - 1)) * LDB; }

__global__ void __launch_bounds__(CH_NT, 1) band_chase2_kernel(int n, cplx* __restrict__ Bd, cplx* __restrict__ RV, int KS,
                                                              double* __restrict__ dd, double* __restrict__ ee,
                                                              int* gprog) {
  extern __shared__ __align__(16) unsigned char smraw[];
  cplx* ring = reinterpret_cast<cplx*>(smraw);
  __shared__ volatile int s_loaded, s_done;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = lane & 7, g4 = lane >> 3;  // window column / row group of the lane
  const int NG = (n - 1 + CH_W - 1) / CH_W;
  for (int grp = blockIdx.x; grp < NG; grp += gridDim.x) {
    const int s0 = grp * CH_W;
    const int nact = min(CH_W, n - 1 - s0);  // sweeps s0 .. s0 + nact - 1 (<= n - 2)
    const int wl = nact - 1, sl = s0 + wl;
    const int KsL = (n - 1 - sl + SB - 1) / SB;
    const int Tend = 2 * wl + KsL;
    if (threadIdx.x == 0) {
      s_loaded = 0;
      s_done = 0;
    }
    __syncthreads();
    if (warp == CH_W) {
      // ------------------------------ loader / write-back warp ------------------------------
      int Tl = 0, Tw = 0;
      int wb = sl + 1;  // columns < wb need no write-back (nobody reads the group's own columns again)
      auto need_excl = [&](int T) { return min(n, s0 + 9 + 8 * T); };
      auto copy_in = [&](int lo, int hi) {
        for (int e = lo * LDB + lane; e < hi * LDB; e += 32) {
          const int col = e / LDB, d = e % LDB;
          ring_col(ring, col)[d] = ldcg_c(Bd + (long)col * LDB + d);
        }
      };
      auto copy_out = [&](int lo, int hi) {
        for (int e = lo * LDB + lane; e < hi * LDB; e += 32) {
          const int col = e / LDB, d = e % LDB;
          Bd[(long)col * LDB + d] = ring_col(ring, col)[d];
        }
      };
      // loads and write-backs are issued for CH_B slots at a time: every batch costs a few dependent L2 round trips
      // (progress counter, columns, release), which a one-slot batch would put on the compute warps' critical path
      while (Tw < Tend) {
        const int done = s_done;
        bool progressed = false;
        if (Tl < Tend && Tl <= done + CH_L) {
          const int Tn = min(Tend, Tl + CH_B);  // load the columns of slots Tl .. Tn-1
          const int lo = (Tl == 0) ? s0 : need_excl(Tl - 1), hi = need_excl(Tn - 1);
          bool ready = true;
          if (hi > lo && grp > 0) {
            // the previous group's last sweep is s0 - 1: columns < s0 + 8 F are final after F of its steps.  Relaxed
            // poll (an acquire load costs an L1 invalidation per attempt); the columns themselves are read with
            // ld.cg from L2, behind the control dependency on the polled value
            const int req = max(1, (hi - s0 + 7) / 8);
            int v = 0;
            if (lane == 0) asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(gprog + grp - 1) : "memory");
            v = __shfl_sync(0xffffffffu, v, 0);
            ready = v >= req;
          }
          if (ready) {
            if (hi > lo) copy_in(lo, hi);
            __syncwarp();
            __threadfence_block();
            if (lane == 0) s_loaded = Tn;
            Tl = Tn;
            progressed = true;
          }
        }
        if (Tw < done && (done - Tw >= CH_B || done == Tend)) {
          const int kL = (done - 1) - 2 * wl;  // steps the last sweep has completed: kL + 1
          if (kL >= 0) {
            const int cfin = min(n, sl + 1 + 8 * (kL + 1));
            if (cfin > wb) {
              copy_out(wb, cfin);
              wb = cfin;
            }
            __syncwarp();
            if (lane == 0) {
              const int val = kL + 1;
              asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(gprog + grp), "r"(val) : "memory");
            }
          }
          Tw = done;
          progressed = true;
        }
        if (!progressed) __nanosleep(40);
      }
      {
        const int top = need_excl(Tend - 1);
        if (top > wb) copy_out(wb, top);
        __syncwarp();
        if (lane == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(gprog + grp), "r"(CH_DONE) : "memory");
      }
    } else {
      // ------------------------------ compute warps: one sweep each ------------------------------
      const int w = warp;
      const int s = s0 + w;
      const bool active = w < nact;
      const int Ks = active ? (n - 1 - s + SB - 1) / SB : 0;
      cplx vc = mk(0, 0), tau = mk(0, 0);
      cplx vr[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) vr[i] = mk(0, 0);
      for (int T = 0; T < Tend; ++T) {
        if (w == 0) {  // the newest columns are only touched by the first sweep
          if (lane == 0)
            while (s_loaded <= T) {
            }
          __syncwarp();
          __threadfence_block();
        }
        const int k = T - 2 * w;
        if (active && k >= 0 && k < Ks) {
          const int r0 = s + 1 + SB * k;
          if (k == 0) {
            cplx* col = ring_col(ring, s);
            const cplx xc = (r0 + c < n) ? col[1 + c] : mk(0, 0);
            cplx xr[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int j = 4 * (g4 & 1) + i;
              xr[i] = (r0 + j < n) ? col[1 + j] : mk(0, 0);
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
            for (int i = 0; i < 4; ++i) vr[i] = (4 * (g4 & 1) + i == 0) ? mk(1, 0) : cmul(xr[i], scal);
            if (lane == 0) {
              ee[s] = beta;
              dd[s] = col[0].x;
            }
            __syncwarp();
            if (lane < SB && r0 + lane < n) col[1 + lane] = (lane == 0) ? mk(beta, 0) : mk(0, 0);
          }
          // ---- load the window (rows r0 .. r0+15, columns r0 .. r0+7) from the ring ----
          cplx wv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = 4 * g4 + i;
            const int rg = r0 + r, cgl = r0 + c;
            cplx x = mk(0, 0);
            if (rg < n && cgl < n) {
              if (r >= c)
                x = ring_col(ring, cgl)[r - c];
              else
                x = cconj(ring_col(ring, rg)[c - r]);
            }
            wv[i] = x;
          }
          {
            cplx* rv = RV + ((long)s * KS + k) * 9;
            if (lane < SB) rv[lane] = vc;
            if (lane == SB) rv[SB] = tau;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {  // W <- W H
            cplx z = cmul(wv[i], vc);
            z = cadd(z, shfl_xor_c(z, 1));
            z = cadd(z, shfl_xor_c(z, 2));
            z = cadd(z, shfl_xor_c(z, 4));
            cfms(wv[i], cmul(tau, z), cconj(vc));
          }
          {  // D <- H^H D
            cplx y = mk(0, 0);
            if (g4 < 2) {
#pragma unroll
              for (int i = 0; i < 4; ++i) cfma_conj(y, vr[i], wv[i]);
            }
            y = cadd(y, shfl_xor_c(y, 8));
            if (g4 < 2) {
              const cplx ty = cmul(cconj(tau), y);
#pragma unroll
              for (int i = 0; i < 4; ++i) cfms(wv[i], vr[i], ty);
            }
          }
          const bool hasE = (r0 + SB < n);
          cplx vc2 = mk(0, 0), tau2 = mk(0, 0);
          cplx vr2[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) vr2[i] = mk(0, 0);
          if (hasE) {
            cplx xc = mk(0, 0);
            cplx xr[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const cplx t = shfl_c(wv[i], 16 + 8 * (c >> 2));
              if ((c & 3) == i) xc = t;
              xr[i] = shfl_c(wv[i], 16 + 8 * (g4 & 1));
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
            for (int i = 0; i < 4; ++i) vr2[i] = (4 * (g4 & 1) + i == 0) ? mk(1, 0) : cmul(xr[i], scal);
            cplx y = mk(0, 0);
            if (g4 >= 2) {
#pragma unroll
              for (int i = 0; i < 4; ++i) cfma_conj(y, vr2[i], wv[i]);
            }
            y = cadd(y, shfl_xor_c(y, 8));
            if (g4 >= 2) {
              const cplx ty = cmul(cconj(tau2), y);
#pragma unroll
              for (int i = 0; i < 4; ++i) cfms(wv[i], vr2[i], ty);
              if (c == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) wv[i] = (g4 == 2 && i == 0) ? mk(beta2, 0) : mk(0, 0);
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {  // store the lower part of the window
            const int r = 4 * g4 + i;
            const int rg = r0 + r, cgl = r0 + c;
            if (rg < n && cgl < n && r >= c) {
              cplx x = wv[i];
              if (r == c) x.y = 0.0;
              ring_col(ring, cgl)[r - c] = x;
            }
          }
          vc = vc2;
          tau = tau2;
#pragma unroll
          for (int i = 0; i < 4; ++i) vr[i] = vr2[i];
          if (s == n - 2 && k + 1 == Ks) {
            __syncwarp();
            if (lane == 0) dd[n - 1] = ring_col(ring, n - 1)[0].x;
          }
        }
        // end of the slot: every sweep of the group has finished its step
        asm volatile("bar.sync 1, %0;" ::"n"(CH_W * 32) : "memory");
        if (threadIdx.x == 0) {
          __threadfence_block();
          s_done = T + 1;
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------------
// X <- Q2 X for NC columns per CTA held in shared memory.  X(:, col) starts as the real eigenvector Z(:, perm[col]) of the
// tridiagonal matrix.  Shared layout: row r of local column q at  q*npad + (r & 7)*(npad/8) + (r >> 3)  so that the
// threads of a warp (consecutive reflectors = row offsets 8 apart) touch consecutive addresses.
// ------------------------------------------------------------------------------------------------------------------
constexpr int Q2_NT = 512;
__global__ void __launch_bounds__(Q2_NT) q2_apply_kernel(int n, int k, int NC, const double* __restrict__ Z, long ldz,
                                                       const int* __restrict__ perm, const cplx* __restrict__ RV, int KS,
                                                       cplx* __restrict__ U, long ldu) {
  extern __shared__ __align__(16) unsigned char smraw[];
  cplx* X = reinterpret_cast<cplx*>(smraw);
  const int npad = (n + 7) & ~7;
  const int n8 = npad >> 3;
  cplx* Rb = X + (size_t)NC * npad;  // two buffers of KS*9 entries: the reflectors of the current and the next sweep
  const int rvsz = KS * 9;
  const int col0 = blockIdx.x * NC;
  const int nc = min(NC, k - col0);
  if (nc <= 0) return;
  const int tid = threadIdx.x;
  auto stage = [&](int s, int buf) {
    const int cnt = ((n - 1 - s + SB - 1) / SB) * 9;
    const cplx* src = RV + (long)s * rvsz;
    cplx* dst = Rb + (size_t)buf * rvsz;
    for (int e = tid; e < cnt; e += Q2_NT) {
      const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + e);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + e) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage(n - 2, 0);
  for (int q = 0; q < nc; ++q) {
    const double* zc = Z + (long)(perm ? perm[col0 + q] : col0 + q) * ldz;
    for (int r = tid; r < npad; r += blockDim.x)
      X[q * npad + (r & 7) * n8 + (r >> 3)] = (r < n) ? mk(zc[r], 0.0) : mk(0, 0);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  for (int s = n - 2; s >= 0; --s) {
    const int buf = (n - 2 - s) & 1;
    if (s > 0) stage(s - 1, buf ^ 1);  // that buffer was last read two sweeps ago (a barrier lies in between)
    const cplx* R = Rb + (size_t)buf * rvsz;
    const int Ks = (n - 1 - s + SB - 1) / SB;
    const int ntask = Ks * nc;
    // two independent (reflector, column) tasks per iteration: the dot-product / update chains of one task are
    // latency bound, a second one in flight hides them
    for (int id = tid; id < ntask; id += 2 * Q2_NT) {
      const int idb = id + Q2_NT;
      const bool hb = idb < ntask;
      const int ka = id % Ks, qa = id / Ks;
      const int kb = hb ? idb % Ks : ka, qb = hb ? idb / Ks : qa;
      const cplx* rva = R + ka * 9;
      const cplx* rvb = R + kb * 9;
      const cplx taua = rva[SB];
      const cplx taub = hb ? rvb[SB] : mk(0, 0);
      const int ra = s + 1 + SB * ka, rb = s + 1 + SB * kb;
      cplx* xqa = X + qa * npad;
      cplx* xqb = X + qb * npad;
      cplx xa[SB], va[SB], xb[SB], vb[SB];
      cplx da0 = mk(0, 0), da1 = mk(0, 0), db0 = mk(0, 0), db1 = mk(0, 0);
#pragma unroll
      for (int i = 0; i < SB; ++i) {  // rows past the end carry v = 0 and are neither read nor written
        va[i] = rva[i];
        vb[i] = rvb[i];
        xa[i] = (ra + i < npad) ? xqa[((ra + i) & 7) * n8 + ((ra + i) >> 3)] : mk(0, 0);
        xb[i] = (hb && rb + i < npad) ? xqb[((rb + i) & 7) * n8 + ((rb + i) >> 3)] : mk(0, 0);
      }
#pragma unroll
      for (int i = 0; i < SB; i += 2) {
        cfma_conj(da0, va[i], xa[i]);
        cfma_conj(da1, va[i + 1], xa[i + 1]);
        cfma_conj(db0, vb[i], xb[i]);
        cfma_conj(db1, vb[i + 1], xb[i + 1]);
      }
      const cplx tda = cmul(taua, cadd(da0, da1));
      const cplx tdb = cmul(taub, cadd(db0, db1));
#pragma unroll
      for (int i = 0; i < SB; ++i) {
        cfms(xa[i], va[i], tda);
        if (ra + i < npad) xqa[((ra + i) & 7) * n8 + ((ra + i) >> 3)] = xa[i];
      }
      if (hb) {
#pragma unroll
        for (int i = 0; i < SB; ++i) {
          cfms(xb[i], vb[i], tdb);
          if (rb + i < npad) xqb[((rb + i) & 7) * n8 + ((rb + i) >> 3)] = xb[i];
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
  for (int q = 0; q < nc; ++q)
    for (int r = tid; r < n; r += blockDim.x) U[(long)(col0 + q) * ldu + r] = X[q * npad + (r & 7) * n8 + (r >> 3)];
}

// ------------------------------------------------------------------------------------------------------------------
// X <- Q2 X, second version: no CTA-wide barrier.  Thread k owns the reflectors (s, k) of ALL sweeps (rows s+1+8k ..);
// reflector (s-1, k) touches rows written by (s, k) (same thread) and by (s, k-1) (left neighbour) only, so the lanes
// of a warp advance sweep by sweep with __syncwarp and a warp follows its left neighbour warp through one progress
// word in shared memory.  The NC columns of the tile are processed two at a time for instruction-level parallelism;
// the reflector of the next sweep is prefetched from L2 into registers.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(288) q2_apply2_kernel(int n, int k, int NC, const double* __restrict__ Z, long ldz,
                                                        const int* __restrict__ perm, const cplx* __restrict__ RV, int KS,
                                                        cplx* __restrict__ U, long ldu) {
  extern __shared__ __align__(16) unsigned char smraw[];
  cplx* X = reinterpret_cast<cplx*>(smraw);
  const int npad = (n + 7) & ~7;
  const int n8 = npad >> 3;
  volatile int* wprog = reinterpret_cast<volatile int*>(X + (size_t)NC * npad);
  const int col0 = blockIdx.x * NC;
  const int nc = min(NC, k - col0);
  if (nc <= 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int q = 0; q < nc; ++q) {
    const double* zc = Z + (long)(perm ? perm[col0 + q] : col0 + q) * ldz;
    for (int r = tid; r < npad; r += blockDim.x)
      X[q * npad + (r & 7) * n8 + (r >> 3)] = (r < n) ? mk(zc[r], 0.0) : mk(0, 0);
  }
  if (tid < (int)(blockDim.x >> 5)) wprog[tid] = n - 1;  // "no sweep done yet" (sweeps run from n-2 down to 0)
  __syncthreads();
  const int kk = tid;  // reflector index of this thread
  // prefetch the reflector of the first sweep in which this thread is active
  cplx v[SB], vn[SB];
  cplx tau = mk(0, 0), taun = mk(0, 0);
  auto fetch = [&](int s, cplx (&vv)[SB], cplx& tt) {
    const int Ks = (n - 1 - s + SB - 1) / SB;
    if (s >= 0 && kk < Ks) {
      const cplx* rv = RV + ((long)s * KS + kk) * 9;
#pragma unroll
      for (int i = 0; i < SB; ++i) vv[i] = ldcg_c(rv + i);
      tt = ldcg_c(rv + SB);
    } else {
      tt = mk(0, 0);
    }
  };
  fetch(n - 2, vn, taun);
  for (int s = n - 2; s >= 0; --s) {
#pragma unroll
    for (int i = 0; i < SB; ++i) v[i] = vn[i];
    tau = taun;
    fetch(s - 1, vn, taun);  // in flight while this sweep is applied
    if (warp > 0) {           // the left neighbour warp must have finished sweep s+1
      if (lane == 0)
        while (wprog[warp - 1] > s + 1) {
        }
      __syncwarp();
      __threadfence_block();
    }
    const int Ks = (n - 1 - s + SB - 1) / SB;
    if (kk < Ks && !(tau.x == 0.0 && tau.y == 0.0)) {
      const int r0 = s + 1 + SB * kk;
      int off[SB];
#pragma unroll
      for (int i = 0; i < SB; ++i) off[i] = ((r0 + i) & 7) * n8 + ((r0 + i) >> 3);  // rows past n carry v = 0
      for (int q = 0; q < nc; q += 2) {
        const bool hb = q + 1 < nc;
        cplx* xa = X + q * npad;
        cplx* xb = X + (hb ? q + 1 : q) * npad;
        cplx a[SB], b[SB];
        cplx da0 = mk(0, 0), da1 = mk(0, 0), db0 = mk(0, 0), db1 = mk(0, 0);
#pragma unroll
        for (int i = 0; i < SB; ++i) {
          const bool in = r0 + i < npad;
          a[i] = in ? xa[off[i]] : mk(0, 0);
          b[i] = (in && hb) ? xb[off[i]] : mk(0, 0);
        }
#pragma unroll
        for (int i = 0; i < SB; i += 2) {
          cfma_conj(da0, v[i], a[i]);
          cfma_conj(da1, v[i + 1], a[i + 1]);
          cfma_conj(db0, v[i], b[i]);
          cfma_conj(db1, v[i + 1], b[i + 1]);
        }
        const cplx tda = cmul(tau, cadd(da0, da1));
        const cplx tdb = cmul(tau, cadd(db0, db1));
#pragma unroll
        for (int i = 0; i < SB; ++i) {
          const bool in = r0 + i < npad;
          cfms(a[i], v[i], tda);
          if (in) xa[off[i]] = a[i];
          if (hb) {
            cfms(b[i], v[i], tdb);
            if (in) xb[off[i]] = b[i];
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      wprog[warp] = s;
    }
  }
  __syncthreads();
  for (int q = 0; q < nc; ++q)
    for (int r = tid; r < n; r += blockDim.x) U[(long)(col0 + q) * ldu + r] = X[q * npad + (r & 7) * n8 + (r >> 3)];
}

}  // namespace

#ifdef TEVO_SBR_TIMING
extern "C" int tevo_debug_sbr_cycles(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_sbr_cycles, sizeof(unsigned long long) * 16);
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_sbr_cycles, z, sizeof(z));
  }
  return 0;
}
#endif

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

void Sbr::reduce_to_band(cudaStream_t st, int n, cplx* A, long lda, cplx* taus, int max_ctas) {
  if (n > SBR_MAX_N) throw TevoError(1, "sbr: n too large for the two-stage reduction");
  ensure(n);
  // band array: diagonals 0..8 written by stage 1, the bulge diagonals 9..15 start as zeros
  TEVO_CUDA_CHECK(cudaMemsetAsync(Bd_, 0, sizeof(cplx) * (size_t)LDB * (n + 16), st));
  TEVO_CUDA_CHECK(cudaMemsetAsync(small_, 0, sizeof(double) * SBR_SMALL_DOUBLES, st));
  S1Args a;
  a.n = n;
  a.A = A;
  a.lda = lda;
  a.Bd = Bd_;
  a.taus = taus;
  a.Vb = Vb_;
  a.Wb = Wb_;
  a.Yb = reinterpret_cast<double*>(Yb_);
  a.small = small_;
  a.bar = reinterpret_cast<unsigned*>(small_ + 512);
  a.bstride = (long)(n + 64) * SB;
  int grid = num_sms_;
  if (max_ctas > 0) grid = std::min(grid, max_ctas);
  grid = std::max(1, grid);
  void* args[] = {&a};
  const size_t qr_smem = sizeof(cplx) * 4 * (size_t)QR_MP;
  static PerDeviceOnce once;
  once([&]() {
    TEVO_CUDA_CHECK(cudaFuncSetAttribute(sbr_stage1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qr_smem));
  });
  TEVO_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)sbr_stage1_kernel, dim3(grid), dim3(S1_NT), args, qr_smem, st));
  TEVO_LAUNCHED();
  sbr_zero_upper_kernel<<<n, 128, 0, st>>>(n, A, lda);
  TEVO_LAUNCHED();
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
  static const int chase_version = [] {
    const char* e = getenv("TEVO_SBR_CHASE");  // 1: one warp per sweep with hand-offs through L2; 2: lock-step groups
    return e ? atoi(e) : 2;
  }();
  void* args[] = {(void*)&n, (void*)&Bd_, (void*)&RV_, (void*)&KS, (void*)&d, (void*)&e, (void*)&prog_};
  if (chase_version == 1) {
    int ctas = std::min(16, std::max(1, (n - 1 + 7) / 8));
    TEVO_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)band_chase_kernel, dim3(ctas), dim3(256), args, 0, st));
  } else {
    const size_t ring_bytes = sizeof(cplx) * (size_t)CH_RING * LDB;
    static PerDeviceOnce once;
    once([&]() {
      TEVO_CUDA_CHECK(cudaFuncSetAttribute(band_chase2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes));
    });
    const int ngroups = (n - 1 + CH_W - 1) / CH_W;
    const int ctas = std::max(1, std::min(8, ngroups));
    TEVO_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)band_chase2_kernel, dim3(ctas), dim3(CH_NT), args, ring_bytes, st));
  }
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
  const size_t budget = (size_t)227 * 1024;
  static const int q2_version = [] {
    const char* e = getenv("TEVO_SBR_Q2");  // 1: one CTA barrier per sweep (default); 2: warp-pipelined (measured slower)
    return e ? atoi(e) : 1;
  }();
  if (q2_version != 1 && ks_ <= 288) {
    const int threads = 32 * ((ks_ + 31) / 32);
    const size_t extra = 64 * sizeof(int);
    if (per_col + extra > budget) throw TevoError(1, "sbr: matrix too large for the shared-memory back-transformation");
    const int ncmax = (int)((budget - extra) / per_col);
    const int waves = std::max(1, (k + ncmax * num_sms_ - 1) / (ncmax * num_sms_));
    const int NC = std::min(ncmax, std::max(1, (k + waves * num_sms_ - 1) / (waves * num_sms_)));
    const int grid = (k + NC - 1) / NC;
    static PerDeviceOnce once2;
    once2([&]() {
      TEVO_CUDA_CHECK(cudaFuncSetAttribute(q2_apply2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    });
    q2_apply2_kernel<<<grid, threads, (size_t)NC * per_col + extra, st>>>(n, k, NC, Z, ldz, perm_dev, RV_, ks_, U, ldu);
    TEVO_LAUNCHED();
    return;
  }
  const size_t rv_bytes = (size_t)2 * ks_ * 9 * sizeof(cplx);  // reflectors of two sweeps (double buffer)
  if (rv_bytes + per_col > budget) throw TevoError(1, "sbr: matrix too large for the shared-memory back-transformation");
  const int ncmax = (int)((budget - rv_bytes) / per_col);
  // as many columns per CTA as fit, but no more than an even split of the k columns over whole waves of CTAs needs
  const int waves = std::max(1, (k + ncmax * num_sms_ - 1) / (ncmax * num_sms_));
  int NC = std::min(ncmax, std::max(1, (k + waves * num_sms_ - 1) / (waves * num_sms_)));
  const int grid = (k + NC - 1) / NC;
  q2_apply_kernel<<<grid, Q2_NT, (size_t)NC * per_col + rv_bytes, st>>>(n, k, NC, Z, ldz, perm_dev, RV_, ks_, U, ldu);
  TEVO_LAUNCHED();
}

}  // namespace tevo
