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

// Panel version of the same factorisation (default since round 2).  ncu on the register-tiled kernel above: 87 us per
// block, 30-40 % of the warp samples at its 64 barriers, FP64 pipe 19 % busy.  Here the block is factored in 8 panels
// of 8 columns, 2 barriers per panel:
//   (b) warp 0 factors the panel in registers, lane l holding rows c0+l and c0+32+l; pivots and the multipliers of
//       the 8 x 8 diagonal block travel by warp shuffles (a first version let every lane factor the diagonal block
//       redundantly: one warp issues one FP64 instruction per ~4.8 cycles, measured, which made that phase 12 k
//       cycles per panel);
//   (c) the 136 4 x 4 tiles of the lower triangle are dealt to threads 0..135 column by column (tiles left of the
//       current panel retire whole warps) and take the rank-8 update from a copy of the panel stored row-interleaved
//       (sP[j][(i%4)*16 + i/4]) so that a warp's loads of 4 rows per thread are contiguous.
// inv(L) follows by block rows (8 x 8 diagonal inverses in parallel, then X(bi, 0:bi) = -X(bi,bi) [L(bi, 0:bi)
// X(0:bi, 0:bi)], 2 barriers per block row).  sL[c][i] = L(i,c), sX[c][i] = inv(L)(i,c): columns contiguous over rows.
constexpr int PW = 8;        // panel width
constexpr int LDX = NB + 1;  // sX row pitch: threads of a warp read 4 columns c at the same row k
constexpr int NTILE = 136;   // 4 x 4 tiles of the lower triangle of a 64 x 64 block, diagonal included
constexpr size_t POTRF_PANEL_SMEM = sizeof(cplx) * (NB * NB + NB * LDX + 2 * PW * NB) + sizeof(double) * NB;

__device__ __forceinline__ void csubmulc(cplx& a, cplx x, cplx y) {  // a -= x * conj(y), four FMAs
  a.x = fma(-x.x, y.x, a.x);
  a.y = fma(-x.y, y.x, a.y);
  a.x = fma(-x.y, y.y, a.x);
  a.y = fma(x.x, y.y, a.y);
}
__device__ __forceinline__ void caddmul(cplx& a, cplx x, cplx y) {  // a += x * y, four FMAs
  a.x = fma(x.x, y.x, a.x);
  a.y = fma(x.x, y.y, a.y);
  a.x = fma(-x.y, y.y, a.x);
  a.y = fma(x.y, y.x, a.y);
}
__device__ __forceinline__ cplx shfl_c(cplx v, int src) {
  return mk(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}

__global__ void __launch_bounds__(256) potrf_block_panel_kernel(int nb, const cplx* __restrict__ G,
                                                                cplx* __restrict__ Lm, cplx* __restrict__ Xm, long ld,
                                                                double* __restrict__ stats,
                                                                long long* __restrict__ dbg) {
  // dbg (tests only): thread 0's clock64 sums - [0] load, [1] publish + wait for the other warps' updates, [2] panel
  // factor, [3] barrier after it, [4] own tile update, [5] L out + diagonal inverses, [6] inverse block rows, [7] total
  long long tk0 = 0, tk = 0, acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define POTRF_MARK(slot)                \
  if (dbg) { /* uniform: no divergence inside warp 0, whose shuffles would otherwise take their slow path */ \
    const long long now_ = clock64();   \
    acc[slot] += now_ - tk;             \
    tk = now_;                          \
  }
  if (dbg) tk0 = tk = clock64();
  extern __shared__ __align__(16) unsigned char smraw[];
  cplx(*sL)[NB] = reinterpret_cast<cplx(*)[NB]>(smraw);
  cplx(*sX)[LDX] = reinterpret_cast<cplx(*)[LDX]>(smraw + sizeof(cplx) * NB * NB);
  cplx(*sS)[NB] = reinterpret_cast<cplx(*)[NB]>(smraw + sizeof(cplx) * (NB * NB + NB * LDX));  // sS[m][c], m < 8
  cplx(*sP)[NB] = sS + PW;                                                                     // interleaved panel
  double* sinv = reinterpret_cast<double*>(smraw + sizeof(cplx) * (NB * NB + NB * LDX + 2 * PW * NB));
  __shared__ int sfail;
  __shared__ double spiv[2];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  // tile of this thread: column tj holds tiles ti = tj..15, columns laid one after the other
  int ti = 0, tj = 0;
  const bool tile = t < NTILE;
  if (tile) {
    int off = 0;
    while (t >= off + 16 - tj) {
      off += 16 - tj;
      ++tj;
    }
    ti = tj + (t - off);
  }
  cplx a[4][4];
  if (tile) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = 4 * ti + r, j = 4 * tj + c;
        a[r][c] = (i < nb && j < nb) ? (i >= j ? G[(long)j * ld + i] : mk(0, 0)) : mk(i == j ? 1.0 : 0.0, 0.0);
      }
  }
  if (t == 0) sfail = 0;
  for (int e = t; e < NB * LDX; e += 256) (&sX[0][0])[e] = mk(0, 0);
  for (int e = PW * ((nb + PW - 1) / PW) * NB + t; e < NB * NB; e += 256) (&sL[0][0])[e] = mk(0, 0);
  POTRF_MARK(0)
  double pmin = 1e300, pmax = 0.0;  // warp 0 only
  const int npan = (nb + PW - 1) / PW;  // rows and columns >= nb are identity padding: whole panels of it are skipped
  for (int p = 0; p < npan; ++p) {
    const int c0 = PW * p;
    // (a) the owners of the panel's two tile columns publish the current values (rows >= 4 tj)
    if (tile && (tj >> 1) == p) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int r = 0; r < 4; ++r) sL[4 * tj + c][4 * ti + r] = a[r][c];
    }
    __syncthreads();
    POTRF_MARK(1)
    // (b) factor the panel: warps 0..2, each with the 8 rows of the diagonal block in lanes 0..7 (redundantly, so that
    // pivots and multipliers travel by warp shuffles) and 24 of the rows below it in lanes 8..31.  Entries above the
    // diagonal (lane < c <= 7) are carried along as zeros.
    {
      const int rowbase = c0 + PW + 24 * warp;
      if (warp < 3 && (warp == 0 || rowbase < NB)) {
        POTRF_MARK(4)
        const int irow = lane < PW ? c0 + lane : rowbase + (lane - PW);
        const bool v = irow < NB;
        cplx r[PW];
#pragma unroll
        for (int c = 0; c < PW; ++c) r[c] = (v && (lane >= PW || lane >= c)) ? sL[c0 + c][irow] : mk(0, 0);
        bool bad = false;
        POTRF_MARK(2)
        // fully unrolled on purpose: a rolled version (current column always in r[0], the others rotating down) was
        // measured slower (810 vs 550 cycles per column): its shuffles and FMAs serialise on the in-order issue
#pragma unroll
        for (int j = 0; j < PW; ++j) {
          const double d = __shfl_sync(0xffffffffu, r[j].x, j);
          if (!(d > 0.0) || !isfinite(d)) bad = true;  // uniform
          const double inv = rsqrt(d), l = d * inv;
          if (c0 + j < nb) {
            pmin = fmin(pmin, l);
            pmax = fmax(pmax, l);
          }
          r[j] = (lane > j) ? cscale(inv, r[j]) : (lane == j ? mk(l, 0.0) : mk(0, 0));
          if (warp == 0 && lane == j) sinv[c0 + j] = inv;
          cplx lc[PW];
#pragma unroll
          for (int c = j + 1; c < PW; ++c) lc[c] = shfl_c(r[j], c);  // L(c0+c, c0+j): all shuffles first
#pragma unroll
          for (int c = j + 1; c < PW; ++c) csubmulc(r[c], r[j], lc[c]);
        }
        if (!bad && v && (warp == 0 || lane >= PW)) {
#pragma unroll
          for (int c = 0; c < PW; ++c) {
            const cplx val = (lane >= PW || lane >= c) ? r[c] : mk(0, 0);
            sL[c0 + c][irow] = val;
            sP[c][(irow & 3) * 16 + (irow >> 2)] = val;
          }
        }
        POTRF_MARK(3)
        if (bad && t == 0) sfail = 1;
      }
    }
    POTRF_MARK(4)
    __syncthreads();
    if (sfail) {
      if (t == 0) {
        stats[0] = -1.0;
        stats[1] = 0.0;
      }
      return;
    }
    // (c) rank-8 update of the lower tiles right of the panel
    if (tile && (tj >> 1) > p) {
#pragma unroll 2
      for (int j = 0; j < PW; ++j) {
        cplx li[4], lj[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) li[r] = sP[j][r * 16 + ti];
#pragma unroll
        for (int c = 0; c < 4; ++c) lj[c] = sP[j][c * 16 + tj];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int r = 0; r < 4; ++r) csubmulc(a[r][c], li[r], lj[c]);
      }
    }
    // sP is rewritten only after the next barrier; sL columns of the next panel are disjoint from everything read here
    POTRF_MARK(4)
  }
  if (t == 0) {
    spiv[0] = pmin;
    spiv[1] = pmax;
  }
  // L out (zeros above the diagonal)
  for (int e = t; e < NB * NB; e += 256) {
    const int i = e % NB, j = e / NB;
    if (i < nb && j < nb) Lm[(long)j * ld + i] = (i >= j) ? sL[j][i] : mk(0, 0);
  }
  // inverse, diagonal 8 x 8 blocks: thread (w, cc) solves column cc of block w by forward substitution
  if (t < PW * npan) {
    const int w = t / PW, cc = t % PW, b0 = PW * w;
    cplx x[PW];
#pragma unroll
    for (int i = 0; i < PW; ++i) {
      cplx s = mk(0, 0);
#pragma unroll
      for (int k = 0; k < i; ++k) caddmul(s, sL[b0 + k][b0 + i], x[k]);
      const double di = sinv[b0 + i];
      x[i] = (i == cc) ? mk(di, 0.0) : (i > cc ? cscale(-di, s) : mk(0, 0));
    }
#pragma unroll
    for (int i = 0; i < PW; ++i) sX[b0 + cc][b0 + i] = x[i];
  }
  __syncthreads();
  POTRF_MARK(5)
  // pairwise merging of diagonal ranges, inv([A 0; B C]) = [A^-1 0; -C^-1 (B A^-1), C^-1], block sizes h = 8, 16, 32:
  // all pairs of a level together, 2 x 2 outputs per thread, the triangular structure of A^-1 and C^-1 in the loop
  // bounds (everything above the diagonal of sX is zero).  A pair whose lower half is all padding is skipped; rows of
  // padding inside a pair only ever produce rows of padding.
  cplx* sT = &sS[0][0];  // 1024 entries: sS and sP, both free now
  for (int h = PW; h < NB; h *= 2) {
    const int hh = h / 2, tpp = hh * hh, ntile = (NB / (2 * h)) * tpp;
    for (int e = t; e < ntile; e += 256) {  // T = B A^-1
      const int pr = e / tpp, q = e % tpp, tr = q / hh, tc = q % hh;
      const int lo = pr * 2 * h;
      if (lo + h >= PW * npan) continue;
      const int i0 = lo + h + 2 * tr, cc = lo + 2 * tc;
      cplx s00 = mk(0, 0), s01 = mk(0, 0), s10 = mk(0, 0), s11 = mk(0, 0);
      for (int m8 = (2 * tc) & ~7; m8 < h; m8 += 8) {  // whole chunks of 8 terms: A^-1 is zero above its diagonal
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int m = m8 + u;
          const cplx b0 = sL[lo + m][i0], b1 = sL[lo + m][i0 + 1];
          const cplx a0 = sX[cc][lo + m], a1 = sX[cc + 1][lo + m];
          caddmul(s00, b0, a0);
          caddmul(s01, b0, a1);
          caddmul(s10, b1, a0);
          caddmul(s11, b1, a1);
        }
      }
      cplx* T = sT + pr * h * h + (2 * tr) * h + 2 * tc;
      T[0] = s00;
      T[1] = s01;
      T[h] = s10;
      T[h + 1] = s11;
    }
    __syncthreads();
    for (int e = t; e < ntile; e += 256) {  // X(lower left) = -C^-1 T
      const int pr = e / tpp, q = e % tpp, tr = q / hh, tc = q % hh;
      const int lo = pr * 2 * h;
      if (lo + h >= PW * npan) continue;
      const int i0 = lo + h + 2 * tr, cc = lo + 2 * tc;
      cplx s00 = mk(0, 0), s01 = mk(0, 0), s10 = mk(0, 0), s11 = mk(0, 0);
      const cplx* T = sT + pr * h * h + 2 * tc;
      for (int k8 = 0; k8 <= 2 * tr + 1; k8 += 8) {  // whole chunks: C^-1 is zero above its diagonal
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int k = k8 + u;
          const cplx c0v = sX[lo + h + k][i0], c1v = sX[lo + h + k][i0 + 1];
          const cplx t0 = T[k * h], t1 = T[k * h + 1];
          caddmul(s00, c0v, t0);
          caddmul(s01, c0v, t1);
          caddmul(s10, c1v, t0);
          caddmul(s11, c1v, t1);
        }
      }
      sX[cc][i0] = mk(-s00.x, -s00.y);
      sX[cc + 1][i0] = mk(-s01.x, -s01.y);
      sX[cc][i0 + 1] = mk(-s10.x, -s10.y);
      sX[cc + 1][i0 + 1] = mk(-s11.x, -s11.y);
    }
    __syncthreads();
  }
  POTRF_MARK(6)
  for (int e = t; e < NB * NB; e += 256) {
    const int i = e % NB, j = e / NB;
    if (i < nb && j < nb) Xm[(long)j * ld + i] = (i >= j) ? sX[j][i] : mk(0, 0);
  }
  if (t == 0) {
    stats[0] = spiv[0];
    stats[1] = spiv[1];
    if (dbg) {
      acc[7] = clock64() - tk0;
      for (int i = 0; i < 8; ++i) dbg[i] = acc[i];
    }
  }
#undef POTRF_MARK
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
  TEVO_CUDA_CHECK(cudaFuncSetAttribute(potrf_block_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)POTRF_PANEL_SMEM));
  TEVO_CUDA_CHECK(cudaStreamCreateWithFlags(&side_, cudaStreamNonBlocking));
  for (cudaEvent_t* e : {&ev_diag_, &ev_inv_})
    TEVO_CUDA_CHECK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
}

void CholQR::block_step(cudaStream_t st, int variant, int nb, const cplx* G, cplx* L, cplx* X, long ld, double* stats,
                        long long* cycles) {
  if (variant == 0)
    potrf_block_panel_kernel<<<1, 256, POTRF_PANEL_SMEM, st>>>(nb, G, L, X, ld, stats, cycles);
  else if (variant == 2)
    potrf_block_kernel<<<1, 1024, 2 * sizeof(cplx) * NB * NB, st>>>(nb, G, L, X, ld, stats);
  else
    potrf_block_reg_kernel<<<1, 256, 0, st>>>(nb, G, L, X, ld, stats);
  TEVO_LAUNCHED();
}

void CholQR::destroy() {
  if (side_) {
    cudaStreamSynchronize(side_);
    cudaStreamDestroy(side_);
    side_ = nullptr;
  }
  for (cudaEvent_t* e : {&ev_diag_, &ev_inv_})
    if (*e) {
      cudaEventDestroy(*e);
      *e = nullptr;
    }
  auto fr = [](void* p) {
    if (p) cudaFree(p);
  };
  fr(part_);
  part_ = nullptr;
  part_cap_ = 0;
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

// one pass: Q = A L^-H with A^H A = L L^H; pivot extremes per diagonal block go to d_stats_.
//
// The blocked Cholesky is a chain of nblk dependent steps (diagonal block by one CTA, panel GEMM, trailing update), each
// far too small to fill the GPU.  The rows of inv(L) are therefore assembled beside the chain on a second stream
// (side_): row b as soon as potrf(b) is done, X(b, 0:b) = -X_bb [L(b, 0:b) X(0:b, 0:b)] (split-K GEMM), so that only the
// last block row remains after the chain: 0.62 -> 0.07 ms per pass at n = 1024 (gpurun_out -> profiles/r02_late_checks.log).
// The pairwise merging used before ran entirely after the chain; TEVO_CHOL_PIPELINE=0 selects it again.  (Also tried:
// splitting the trailing update into the next block column on the main stream and the rest on a third stream, so that
// the next diagonal block starts earlier.  No gain: a step costs ~104 us either way, because each of its small GEMMs is
// bound by its fixed launch + prologue cost, not by its flops.)
void CholQR::pass(cudaStream_t st, const cplx* A, int m, int n, cplx* L, cplx* Q) {
  const cplx ONE = mk(1, 0), ZERO = mk(0, 0), MONE = mk(-1, 0);
  const int nblk = (n + NB - 1) / NB;
  cplx* X = Linv_;  // n x n: inverse of L, assembled below
  static const bool pipeline = [] {
    const char* e = getenv("TEVO_CHOL_PIPELINE");
    return e ? atoi(e) != 0 : true;
  }();
  const bool piped = pipeline && side_ && nblk >= 3;
  TEVO_CUDA_CHECK(cudaMemsetAsync(L, 0, sizeof(cplx) * (size_t)n * n, st));
  TEVO_CUDA_CHECK(cudaMemsetAsync(X, 0, sizeof(cplx) * (size_t)n * n, st));
  {
    ProfScope ps(st, P_CHOL_GRAM);
    zgemm_herm(st, OP_C, OP_N, n, m, ONE, A, m, A, m, ZERO, G_, n, false);  // Cholesky reads the lower part only
  }
  static const bool smem_version = getenv("TEVO_POTRF_SMEM") != nullptr;  // the 1024-thread shared-memory kernel
  static const bool reg_version = getenv("TEVO_POTRF_REG") != nullptr;    // round 1's default (one barrier per column)
  const int variant = smem_version ? 2 : (reg_version ? 1 : 0);
  std::optional<ProfScope> pf;
  pf.emplace(st, P_CHOL_FACTOR);
  for (int jb = 0; jb < nblk; ++jb) {
    const int j = jb * NB, nb = std::min(NB, n - j), rest = n - j - nb;
    block_step(st, variant, nb, G_ + (long)j * n + j, L + (long)j * n + j, X + (long)j * n + j, n, d_stats_ + 2 * jb,
               nullptr);
    if (piped && jb > 0) {  // block row jb of inv(L) (its rows of L were written by the panels of the steps before)
      TEVO_CUDA_CHECK(cudaEventRecord(ev_diag_, st));
      TEVO_CUDA_CHECK(cudaStreamWaitEvent(side_, ev_diag_, 0));
      zgemm_splitk(side_, OP_N, OP_N, nb, j, j, ONE, L + j, n, X, n, ZERO, T_, nb, &part_, &part_cap_);
      zgemm(side_, OP_N, OP_N, nb, j, nb, MONE, X + (long)j * n + j, n, T_, nb, ZERO, X + j, n);
    }
    if (rest > 0) {
      // panel: L(j+nb:, j:j+nb) = G(j+nb:, j:j+nb) * inv(L_jj)^H ; trailing: G(j+nb:, j+nb:) -= panel * panel^H
      zgemm(st, OP_N, OP_C, rest, nb, nb, ONE, G_ + (long)j * n + j + nb, n, X + (long)j * n + j, n, ZERO,
            L + (long)j * n + j + nb, n);
      zgemm_herm(st, OP_N, OP_C, rest, nb, MONE, L + (long)j * n + j + nb, n, L + (long)j * n + j + nb, n, ONE,
                 G_ + (long)(j + nb) * n + j + nb, n, false);
    }
  }
  pf.reset();
  if (piped) {
    ProfScope pi(st, P_CHOL_INV);  // what is left of the inverse after the chain: the join with the side stream
    TEVO_CUDA_CHECK(cudaEventRecord(ev_inv_, side_));
    TEVO_CUDA_CHECK(cudaStreamWaitEvent(st, ev_inv_, 0));
  } else {
    ProfScope pi(st, P_CHOL_INV);
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
  }
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
