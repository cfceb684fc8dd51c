#include <mutex>
// Hermitian eigensolver on the device: blocked Householder tridiagonalisation (persistent cooperative panel kernel +
// DMMA rank-2k trailing updates), tridiagonal divide & conquer (stedc.cu) and compact-WY back-transformation.
// It is the core of the truncated split (replacebond!, src/bondgate.jl:51, src/tdvp.jl:64): the density matrix
// rho = theta theta^+ (or theta^+ theta) is diagonalised and the isometry is read off its leading eigenvectors.
#include "heig.h"

#include <algorithm>
#include <atomic>
#include <cstdlib>

#include "prof.h"
#include "trinv.cuh"
#include "zgemm.h"

namespace tevo {
namespace {

constexpr int PW = 64;  // panel width
constexpr int RS = 16;  // rows per strip
constexpr int NT = 256;
constexpr int NREP = 8;        // replicas of the panel-dot accumulators (same-address L2 atomics serialise)
constexpr int PANEL_SMEM = 200 * 1024;  // dynamic shared memory budget of the panel kernel
constexpr int MAXCTA = 256;    // sanity cap on the cooperative grid

struct PanelArgs {
  int n, p0, pw;
  int smax;    // strips per CTA whose rows of V, W (and y, next column) are kept in shared memory
  int chunks;  // column chunks per strip: CTA b works on strip b % nstrips, chunk b / nstrips (1: CTA b owns strips b, b+G, ..)
  int ncache;  // 16-column groups of the CTA's trailing-matrix slice kept in shared memory for the whole panel
  cplx* A;
  long lda;
  cplx* P;  // n x 2*PW : [V | W]
  cplx* Q;  // n x 2*PW : [W | V]
  cplx* Y;  // n x PW   : raw y = A22 v
  long ldp;
  double* acc_yv;    // NREP x PW complex: y^H v, accumulated like the panel dots
  double* acc_cw;    // NREP x 2*PW*PW : W^H v, accumulated with atomics into replica blockIdx % NREP
  double* acc_cv;    // NREP x 2*PW*PW : V^H v
  double* cv_out;    // 2*PW*PW  (kept: V^H v, the strictly upper part of V^H V for the compact-WY T factor)
  unsigned* bar;     // grid barrier counter (zeroed before the launch)
  double* d;
  double* e;
  cplx* taus;
};

#ifdef TEVO_PANEL_TIMING
__device__ unsigned long long g_panel_cycles[16];
#define PT_MARK(slot)                                                     \
  do {                                                                    \
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {                \
      long long _now = clock64();                                         \
      atomicAdd(&g_panel_cycles[slot], (unsigned long long)(_now - _t0)); \
      _t0 = _now;                                                         \
    }                                                                     \
  } while (0)
#else
#define PT_MARK(slot)
#endif

__device__ inline cplx ldcg_c(const cplx* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }

// grid-wide barrier for the co-resident (cooperatively launched) panel kernel: one release-reduction per CTA on a
// monotonically increasing counter, polled with acquire loads.  Data produced by other CTAs is read with ld.cg
// afterwards.  (Measured alternatives, all slower on B200: fence + atomicAdd + relaxed poll + fence (+3 %), one flag
// word per CTA polled by everybody (+20 %), gather/release through CTA 0 with one line per CTA (+35 %).)
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

// sum over the 16 "column lanes" (cl = tid/16) for each of the 16 rows (rr = tid%16); valid on cl == 0 threads
__device__ inline void reduce_cl2(cplx& v0, cplx& v1, cplx* sred) {
  __syncthreads();
  sred[threadIdx.x] = v0;
  sred[NT + threadIdx.x] = v1;
  __syncthreads();
  if (threadIdx.x < RS) {
    cplx s0 = mk(0, 0), s1 = mk(0, 0);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      s0 = cadd(s0, sred[c * RS + threadIdx.x]);
      s1 = cadd(s1, sred[NT + c * RS + threadIdx.x]);
    }
    v0 = s0;
    v1 = s1;
  }
}
__device__ inline cplx reduce_cl(cplx v, cplx* sred) {
  __syncthreads();
  sred[threadIdx.x] = v;
  __syncthreads();
  cplx s = mk(0, 0);
  if (threadIdx.x < RS) {
#pragma unroll
    for (int c = 0; c < 16; ++c) s = cadd(s, sred[c * RS + threadIdx.x]);
  }
  return s;
}

// One panel (pw <= 64 columns) of the blocked Householder tridiagonalisation (zlatrd, lower), persistent over its
// columns with two grid barriers per column.  CTA b owns the 16-row strips s = b, b+G, ... of the trailing rows.
//   phase A(i): finalise w_{i-1} on the own rows (scalars from the reductions of column i-1), bring column c = p0+i
//               up to date, accumulate |x|^2                                                        -> barrier
//   phase B(i): reflector (beta, tau, v), y = A22 v for the own rows (HEMV over full rows, L2-resident), panel
//               dots W^H v, V^H v and y^H v (w^H v follows algebraically, so w needs no third reduction) -> barrier
__global__ void __launch_bounds__(NT, 1) tridiag_panel_kernel(PanelArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  // v (at most n - p0 entries) with 16 zero entries in front and behind: the HEMV walks whole 16-column groups and
  // lets the columns <= c / >= n of the first / last group multiply zeros instead of testing every element
  cplx* sv = reinterpret_cast<cplx*>(smraw) + 16;
  // resident copies of the own rows: per strip [V: PW x RS | W: PW x RS | y: RS | next column: RS]
  constexpr int SB = 2 * PW * RS + 2 * RS;
  cplx* sres = sv + ((a.n - a.p0 + 1) & ~1) + 16;
  const int smax = a.smax;
  cplx* sAc = sres + smax * SB;  // resident part of the trailing-matrix slice: [group][thread]
  __shared__ cplx sred[2 * NT];
  __shared__ cplx pivV[PW], pivW[PW], scw[PW], scv[PW];
  __shared__ cplx s_tau_prev, s_alpha_prev, s_yv, s_ypiv;
  __shared__ double sblk[64];
  const int n = a.n, p0 = a.p0, pw = a.pw;
  const int tid = threadIdx.x, rr = tid % RS, cl = tid / RS;
  const int nstrips = (n - p0 + RS - 1) / RS;
  // work split: rows in 16-row strips; if there are fewer strips than CTAs the columns of a strip are split into
  // `chunks` interleaved sets of 16-column groups and the partial y are combined with atomics.  Chunk 0 owns the strip
  // (phase A, panel dots, V/W rows), chunks > 0 only help with the HEMV.
  const int chunks = a.chunks;
  const int my_q = blockIdx.x / nstrips;
  const int s_first = blockIdx.x % nstrips;
  const int s_step = (chunks > 1) ? nstrips : (int)gridDim.x;
  const bool helper = my_q > 0;
  const int own_first = helper ? nstrips : s_first;
  const int NGL = (nstrips - my_q + chunks - 1) / chunks;          // 16-column groups of this CTA: g = my_q + chunks*gl
  const int glc = (a.ncache > 0) ? max(NGL - a.ncache, 0) : NGL;  // groups gl >= glc are resident in shared memory
  const long jstep = 16L * chunks;
  // streamed groups stop before the last group of the matrix if that one is partial (n not a multiple of 16)
  const int NGf = (n - p0) / 16;
  const int NGLf = (NGf > my_q) ? (NGf - my_q + chunks - 1) / chunks : 0;
  const int glcs = min(glc, NGLf);
  const bool tail_streamed = (NGLf < NGL) && (NGLf < glc);
  cplx* Vp = a.P;
  cplx* Wp = a.P + (long)PW * a.ldp;
  cplx* Yp = a.Y;
  unsigned epoch = 0;
  cplx tau_cur = mk(0, 0);
  // own rows of the panel vectors: shared memory for the first smax strips of this CTA, global memory beyond
  auto ldV = [&](int sidx, int k, int r) -> cplx {
    return sidx < smax ? sres[sidx * SB + k * RS + (threadIdx.x % RS)] : Vp[(long)k * a.ldp + r];
  };
  auto ldW = [&](int sidx, int k, int r) -> cplx {
    return sidx < smax ? sres[sidx * SB + (PW + k) * RS + (threadIdx.x % RS)] : Wp[(long)k * a.ldp + r];
  };
#ifdef TEVO_PANEL_TIMING
  long long _t0 = clock64();
#endif

  if (tid < 16) sv[tid - 16] = mk(0, 0);
  if (a.ncache > 0) {
    // the trailing matrix is not modified inside a panel (columns > c are read, column c is written): keep the last
    // groups of this CTA's slice on chip for all pw columns
    const int r = p0 + s_first * RS + rr;
    for (int gl = glc; gl < NGL; gl += 8) {
      cplx x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = p0 + 16 * (my_q + chunks * (gl + u)) + cl;
        x[u] = (gl + u < NGL && r < n && j < n) ? a.A[(long)j * a.lda + r] : mk(0, 0);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (gl + u < NGL) sAc[(gl + u - glc) * NT + tid] = x[u];
    }
  }
  for (int i = 0; i <= pw; ++i) {
    const int c = p0 + i;
    // ---------------- phase A ----------------
    // everything produced by other CTAs that this phase needs is fetched in one round of ld.cg:
    // the reductions of column i-1 (cw, cv, y^H v), the pivot row c of V / W and the raw y of the pivot row
    if (i > 0) {
      const int ip = i - 1;
      for (int k = tid; k < i; k += NT) {
        if (k < ip) {
          cplx w[NREP], v[NREP];
#pragma unroll
          for (int q = 0; q < NREP; ++q) {
            w[q] = ldcg_c(reinterpret_cast<const cplx*>(a.acc_cw) + (long)q * PW * PW + ip * PW + k);
            v[q] = ldcg_c(reinterpret_cast<const cplx*>(a.acc_cv) + (long)q * PW * PW + ip * PW + k);
          }
          if (i < pw) pivW[k] = ldcg_c(&Wp[(long)k * a.ldp + c]);
#pragma unroll
          for (int q = 1; q < NREP; ++q) {
            w[0] = cadd(w[0], w[q]);
            v[0] = cadd(v[0], v[q]);
          }
          scw[k] = w[0];
          scv[k] = v[0];
          if (blockIdx.x == 0) reinterpret_cast<cplx*>(a.cv_out)[ip * PW + k] = v[0];
        }
        if (i < pw) pivV[k] = ldcg_c(&Vp[(long)k * a.ldp + c]);
      }
      if (tid == NT - 1) {  // a thread of another warp than the one that reduces below
        cplx y[NREP];
#pragma unroll
        for (int q = 0; q < NREP; ++q) y[q] = ldcg_c(reinterpret_cast<const cplx*>(a.acc_yv) + q * PW + ip);
        const cplx ypiv = (i < pw) ? ldcg_c(&Yp[(long)ip * a.ldp + c]) : mk(0, 0);
#pragma unroll
        for (int q = 1; q < NREP; ++q) y[0] = cadd(y[0], y[q]);
        s_yv = y[0];
        s_ypiv = ypiv;
      }
      __syncthreads();
      PT_MARK(0);
      if (tid < 32) {  // one warp: Re(cw^H cv) and the pivot-row correction sum
        double re = 0.0;
        cplx acc = mk(0, 0);
        for (int k = tid; k < ip; k += 32) {
          re += scw[k].x * scv[k].x + scw[k].y * scv[k].y;
          if (i < pw) {
            cfma(acc, pivV[k], scw[k]);
            cfma(acc, pivW[k], scv[k]);
          }
        }
        re = warp_sum(re);
        acc.x = warp_sum(acc.x);
        acc.y = warp_sum(acc.y);
        if (tid == 0) {
          const cplx yv = mk(s_yv.x - 2.0 * re, s_yv.y);
          const cplx wv = cmul(cconj(tau_cur), yv);  // w^H v = conj(tau) (y^H v - 2 Re(cw^H cv))
          const cplx alphap = cscale(-0.5, cmul(tau_cur, wv));
          s_tau_prev = tau_cur;
          s_alpha_prev = alphap;
          if (i < pw) pivW[ip] = cadd(cmul(tau_cur, csub(s_ypiv, acc)), cmul(alphap, pivV[ip]));
        }
      }
      // no barrier here: the scalars are consumed only behind the barriers of the row reductions below, so the other
      // warps start their panel products while warp 0 finishes the scalar chain
    }
    if (i == pw) {
      // last pass: only finalise w_{pw-1} on the own rows, then leave the column loop
      const int ip = pw - 1;
      const int cp = p0 + ip;
      int sidx = 0;
      for (int s = own_first; s < nstrips; s += s_step, ++sidx) {
        const int r = p0 + s * RS + rr;
        const bool act = (r < n) && (r >= cp + 1);
        cplx acc = mk(0, 0);
        if (act) {
          for (int k = cl; k < ip; k += 16) {
            cfma(acc, ldV(sidx, k, r), scw[k]);
            cfma(acc, ldW(sidx, k, r), scv[k]);
          }
        }
        const cplx tot = reduce_cl(acc, sred);
        if (cl == 0 && act) {
          const cplx y = (chunks > 1) ? ldcg_c(&Yp[(long)ip * a.ldp + r])
                                      : (sidx < smax ? sres[sidx * SB + 2 * PW * RS + rr] : Yp[(long)ip * a.ldp + r]);
          const cplx v = ldV(sidx, ip, r);
          Wp[(long)ip * a.ldp + r] = cadd(cmul(s_tau_prev, csub(y, tot)), cmul(s_alpha_prev, v));
        }
      }
      break;
    }
    PT_MARK(1);
    {
      const int ip = i - 1;
      int sidx = 0;
      for (int s = own_first; s < nstrips; s += s_step, ++sidx) {
        const int r = p0 + s * RS + rr;
        const bool act = (r < n) && (r >= c);
        cplx acc1 = mk(0, 0), acc2 = mk(0, 0);
        cplx ypre = mk(0, 0);  // partial y combined by the chunks: one L2 round trip, overlapped with the products
        if (chunks > 1 && cl == 0 && act && i > 0) ypre = ldcg_c(&Yp[(long)ip * a.ldp + r]);
        if (act && i > 0) {
#pragma unroll 4
          for (int k = cl; k < ip; k += 16) {
            const cplx v = ldV(sidx, k, r);
            const cplx w = ldW(sidx, k, r);
            cfma(acc1, v, scw[k]);
            cfma(acc1, w, scv[k]);
            cfma(acc2, v, cconj(pivW[k]));
            cfma(acc2, w, cconj(pivV[k]));
          }
        }
        if (i > 0) reduce_cl2(acc1, acc2, sred);
        if (cl == 0 && act) {
          const bool res = sidx < smax;
          // the column was prefetched into shared memory during the previous column's phase B (it is not modified
          // inside this launch before this point)
          cplx av = (res && i > 0) ? sres[sidx * SB + 2 * PW * RS + RS + rr] : a.A[(long)c * a.lda + r];
          if (i > 0) {
            const cplx y = (chunks > 1) ? ypre : (res ? sres[sidx * SB + 2 * PW * RS + rr] : Yp[(long)ip * a.ldp + r]);
            const cplx v = ldV(sidx, ip, r);
            const cplx wf = cadd(cmul(s_tau_prev, csub(y, acc1)), cmul(s_alpha_prev, v));
            Wp[(long)ip * a.ldp + r] = wf;
            if (res) sres[sidx * SB + (PW + ip) * RS + rr] = wf;
            cfma(acc2, v, cconj(pivW[ip]));
            cfma(acc2, wf, cconj(pivV[ip]));
            av = csub(av, acc2);
            a.A[(long)c * a.lda + r] = av;
          }
          if (chunks > 1) Yp[(long)i * a.ldp + r] = mk(0, 0);  // the chunks add their partial y in phase B
        }
      }
    }
    PT_MARK(2);
    PT_MARK(3);
    grid_bar(a.bar, ++epoch);
    PT_MARK(4);
    // ---------------- phase B: reflector, v, y = A22 v, panel dots ----------------
    const int nv = n - (c + 1);
    // software pipelining across the scalar phase: the first 8 trailing-matrix loads of this CTA's first strip do not
    // depend on v, so they are issued now and consumed by the HEMV below
    const int g_min = (c + 1 - p0) >> 4;  // first group with a column > c
    const int gl0 = (g_min <= my_q) ? 0 : (g_min - my_q + chunks - 1) / chunks;
    cplx pre[8];
    {
      const int r0 = p0 + s_first * RS + rr;
      const bool act0 = (r0 < n) && (r0 >= c + 1);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = p0 + 16 * (my_q + chunks * (gl0 + u)) + cl;
        pre[u] = (act0 && gl0 + u < glcs) ? a.A[(long)j * a.lda + r0] : mk(0, 0);
      }
    }
    const cplx alpha = ldcg_c(&a.A[(long)c * a.lda + c + 1]);
    double nloc;
    {
      // scalars and the raw column are fetched in one round of independent loads; scaling happens in shared memory
      const cplx* colp = a.A + (long)c * a.lda + c + 1;
      int j = tid;
      nloc = 0.0;
      for (; j + 7 * NT < nv; j += 8 * NT) {
        cplx x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = ldcg_c(colp + j + u * NT);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          sv[j + u * NT] = x[u];
          if (j + u * NT > 0) nloc += cabs2(x[u]);
        }
      }
      for (; j + 3 * NT < nv; j += 4 * NT) {
        const cplx x0 = ldcg_c(colp + j), x1 = ldcg_c(colp + j + NT), x2 = ldcg_c(colp + j + 2 * NT),
                   x3 = ldcg_c(colp + j + 3 * NT);
        sv[j] = x0;
        sv[j + NT] = x1;
        sv[j + 2 * NT] = x2;
        sv[j + 3 * NT] = x3;
        nloc += ((j > 0) ? cabs2(x0) : 0.0) + cabs2(x1) + cabs2(x2) + cabs2(x3);
      }
      for (; j < nv; j += NT) {
        const cplx x0 = ldcg_c(colp + j);
        sv[j] = x0;
        if (j > 0) nloc += cabs2(x0);
      }
    }
    // |x|^2 (below the pivot element): every CTA holds the whole column, so the norm needs no grid-wide reduction;
    // all CTAs add the same numbers in the same order and get bit-identical reflectors
    nloc = warp_sum(nloc);
    if ((tid & 31) == 0) sblk[tid >> 5] = nloc;
    __syncthreads();
    PT_MARK(5);
    double xn2 = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) xn2 += sblk[w];
    double beta;
    cplx tau, scal;
    const double s2 = cabs2(alpha) + xn2;
    if ((xn2 == 0.0 && alpha.y == 0.0) || !(s2 > 1e-290)) {  // nothing to annihilate (or a column of underflows)
      beta = alpha.x;
      tau = mk(0, 0);
      scal = mk(0, 0);
    } else {
      // |x| and 1/|x| from one reciprocal square root (+ a Newton correction), 1/(alpha - beta) from one
      // reciprocal: the sqrt -> divide chain is on the critical path of every column
      double rs = rsqrt(s2);
      double nx = s2 * rs;
      nx = fma(0.5 * rs, fma(-nx, nx, s2), nx);
      rs = fma(rs, fma(-nx, rs, 1.0), rs);
      beta = -copysign(nx, alpha.x);
      const double ib = -copysign(rs, alpha.x);
      tau = mk((beta - alpha.x) * ib, -alpha.y * ib);
      const cplx den = mk(alpha.x - beta, alpha.y);
      const double dn = cabs2(den);
      double rdn = __drcp_rn(dn);
      scal = mk(den.x * rdn, -den.y * rdn);
    }
    tau_cur = tau;
    if (blockIdx.x == 0 && tid == 0) {
      a.d[c] = ldcg_c(&a.A[(long)c * a.lda + c]).x;
      a.e[c] = beta;
      a.taus[c] = tau;
    }
    for (int j = tid; j < nv; j += NT) sv[j] = (j == 0) ? mk(1, 0) : cmul(sv[j], scal);
    if (tid < 16) sv[nv + tid] = mk(0, 0);
    __syncthreads();
    PT_MARK(6);
    cplx cwa[PW / 16], cva[PW / 16];
#pragma unroll
    for (int q = 0; q < PW / 16; ++q) cwa[q] = cva[q] = mk(0, 0);
    double yv0 = 0.0, yv1 = 0.0;
    bool has_rows = false;
    int sidx = 0;
    for (int s = s_first; s < nstrips; s += s_step, ++sidx) {
      const int r = p0 + s * RS + rr;
      const bool act = (r < n) && (r >= c + 1);
      has_rows = has_rows || (!helper && p0 + s * RS + RS - 1 >= c + 1);  // uniform over the CTA
      cplx acc = mk(0, 0);
      // next column of the own rows, consumed by phase A of column i+1 (untouched by this launch so far)
      cplx anext = mk(0, 0);
      if (cl == 0 && act && !helper && sidx < smax && i + 1 < pw) anext = a.A[(long)(c + 1) * a.lda + r];
      if (act) {
        const cplx* arow = a.A + r;
        cplx ac[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) ac[u] = mk(0, 0);
        const int jb0 = p0 + 16 * my_q + cl;  // column of local group gl: jb0 + jstep * gl
        const cplx* svb = sv - c - 1;          // v_j = svb[j]; zero for c - 15 <= j <= c and n <= j < n + 16
        const long astep = jstep * a.lda;
        int gl = gl0;
        if (sidx == 0) {  // prefetched groups
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (gl0 + u < glcs) cfma(ac[u & 3], pre[u], svb[jb0 + (int)jstep * (gl0 + u)]);
          gl += 8;
        }
        // streamed groups (L2): 16 independent 16-byte loads in flight per thread, no per-element tests
        for (; gl + 15 < glcs; gl += 16) {
          cplx av[16];
          const cplx* ap = arow + (long)(jb0 + (int)jstep * gl) * a.lda;
          const cplx* vp = svb + jb0 + (int)jstep * gl;
#pragma unroll
          for (int u = 0; u < 16; ++u) av[u] = ap[u * astep];
#pragma unroll
          for (int u = 0; u < 16; ++u) cfma(ac[u & 3], av[u], vp[u * (int)jstep]);
        }
        for (; gl + 3 < glcs; gl += 4) {
          cplx av[4];
          const cplx* ap = arow + (long)(jb0 + (int)jstep * gl) * a.lda;
          const cplx* vp = svb + jb0 + (int)jstep * gl;
#pragma unroll
          for (int u = 0; u < 4; ++u) av[u] = ap[u * astep];
#pragma unroll
          for (int u = 0; u < 4; ++u) cfma(ac[u], av[u], vp[u * (int)jstep]);
        }
        for (; gl < glcs; ++gl) {
          const int j = jb0 + (int)jstep * gl;
          cfma(ac[0], arow[(long)j * a.lda], svb[j]);
        }
        if (tail_streamed && NGLf >= gl0) {  // the last, partial group of the matrix (n not a multiple of 16)
          const int j = jb0 + (int)jstep * NGLf;
          if (j < n) cfma(ac[1], arow[(long)j * a.lda], svb[j]);
        }
        // resident groups (shared memory; only CTAs with a single strip have any; columns >= n are stored as zeros)
        int gc = max(gl0, glc);
        for (; gc + 3 < NGL; gc += 4) {
          const cplx* vp = svb + jb0 + (int)jstep * gc;
          const cplx* cp = sAc + (gc - glc) * NT + tid;
#pragma unroll
          for (int u = 0; u < 4; ++u) cfma(ac[u], cp[u * NT], vp[u * (int)jstep]);
        }
        for (; gc < NGL; ++gc) cfma(ac[0], sAc[(gc - glc) * NT + tid], svb[jb0 + (int)jstep * gc]);
#pragma unroll
        for (int u = 1; u < 4; ++u) ac[0] = cadd(ac[0], ac[u]);
        acc = ac[0];
        if (!helper) {
          const cplx vr = sv[r - c - 1];
#pragma unroll
          for (int q = 0; q < PW / 16; ++q) {
            const int k = cl + 16 * q;
            if (k < i) {
              cfma_conj(cwa[q], ldW(sidx, k, r), vr);
              cfma_conj(cva[q], ldV(sidx, k, r), vr);
            }
          }
        }
      }
      const cplx y = reduce_cl(acc, sred);
      if (cl == 0 && act) {
        const cplx vr = sv[r - c - 1];
        if (!helper) {
          Vp[(long)i * a.ldp + r] = vr;
          if (sidx < smax) {
            sres[sidx * SB + i * RS + rr] = vr;
            sres[sidx * SB + 2 * PW * RS + RS + rr] = anext;
          }
        }
        if (chunks > 1) {
          double* yp = reinterpret_cast<double*>(&Yp[(long)i * a.ldp + r]);
          atomicAdd(yp, y.x);
          atomicAdd(yp + 1, y.y);
        } else {
          Yp[(long)i * a.ldp + r] = y;
          if (sidx < smax) sres[sidx * SB + 2 * PW * RS + rr] = y;
        }
        yv0 += y.x * vr.x + y.y * vr.y;
        yv1 += y.x * vr.y - y.y * vr.x;
      }
    }
    PT_MARK(7);
    // reduce the panel dots over the 16 rows of the half-warp, then one atomic per (k, CTA)
#pragma unroll
    for (int q = 0; q < PW / 16; ++q) {
      if (16 * q >= i) continue;  // no panel column k = cl + 16 q < i in this group (uniform across the CTA)
      double x0 = cwa[q].x, x1 = cwa[q].y, x2 = cva[q].x, x3 = cva[q].y;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        x0 += __shfl_xor_sync(0xffffffffu, x0, o);
        x1 += __shfl_xor_sync(0xffffffffu, x1, o);
        x2 += __shfl_xor_sync(0xffffffffu, x2, o);
        x3 += __shfl_xor_sync(0xffffffffu, x3, o);
      }
      const int k = cl + 16 * q;
      if (rr == 0 && k < i && has_rows) {
        double* cw = a.acc_cw + (long)(blockIdx.x % NREP) * 2 * PW * PW + 2 * (i * PW + k);
        double* cv = a.acc_cv + (long)(blockIdx.x % NREP) * 2 * PW * PW + 2 * (i * PW + k);
        atomicAdd(cw, x0);
        atomicAdd(cw + 1, x1);
        atomicAdd(cv, x2);
        atomicAdd(cv + 1, x3);
      }
    }
    PT_MARK(8);
    if (tid < 32) {  // y^H v: the partial sums live in the cl == 0 threads (lanes 0..15 of warp 0)
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        yv0 += __shfl_xor_sync(0xffffffffu, yv0, o);
        yv1 += __shfl_xor_sync(0xffffffffu, yv1, o);
      }
      if (tid == 0 && (yv0 != 0.0 || yv1 != 0.0)) {
        double* yp = a.acc_yv + 2 * ((blockIdx.x % NREP) * PW + i);
        atomicAdd(yp, yv0);
        atomicAdd(yp + 1, yv1);
      }
    }
    PT_MARK(9);
    grid_bar(a.bar, ++epoch);
    PT_MARK(10);
  }
  // ---------------- panel epilogue: build [V|W], [W|V], store reflectors into A ----------------
  __syncthreads();
  for (int s = own_first; s < nstrips; s += s_step) {
    for (int k = cl; k < pw; k += 16) {
      const int r = p0 + s * RS + rr;
      if (r >= n) continue;
      const int c = p0 + k;
      const cplx v = (r <= c) ? mk(0, 0) : Vp[(long)k * a.ldp + r];
      const cplx w = (r <= c) ? mk(0, 0) : Wp[(long)k * a.ldp + r];
      Vp[(long)k * a.ldp + r] = v;
      Wp[(long)k * a.ldp + r] = w;
      a.Q[(long)k * a.ldp + r] = w;
      a.Q[(long)(PW + k) * a.ldp + r] = v;
      a.A[(long)c * a.lda + r] = v;
    }
  }
  // rows above the panel: explicit zeros in the reflector columns
  for (long e = (long)blockIdx.x * NT + tid; e < (long)p0 * pw; e += (long)gridDim.x * NT) {
    const int r = (int)(e % p0);
    const int k = (int)(e / p0);
    a.A[(long)(p0 + k) * a.lda + r] = mk(0, 0);
  }
}

// compact-WY T factor of one reflector panel from the identity  T^-1 = triu(V^H V, 1) + diag(1/tau)  (zlarft):
// the upper-triangular M is inverted in shared memory (as the transpose of a lower-triangular inverse).
// S holds (V^H v_l)_j for j < l at S[2*(l*PW + j)] (the panel kernel's acc_cv).  tau_j = 0 (identity reflector):
// row and column j of T are zero.  T is written column-major PW x PW.
// One CTA per panel: panel b covers columns b*PW .. of the n x n problem.
__global__ void __launch_bounds__(1024) wy_T_kernel(int n, const cplx* __restrict__ taus_all,
                                                   const double* __restrict__ S_all, cplx* __restrict__ T_all) {
  const int pw = min(PW, n - 1 - (int)blockIdx.x * PW);
  const cplx* taus = taus_all + (long)blockIdx.x * PW;
  const double* S = S_all + (long)blockIdx.x * 2 * PW * PW;
  cplx* T = T_all + (long)blockIdx.x * PW * PW;
  extern __shared__ __align__(16) unsigned char smraw[];
  cplx(*a)[TRB] = reinterpret_cast<cplx(*)[TRB]>(smraw);  // a[p][i] = M^T(i,p) = M(p,i)
  cplx(*x)[TRB] = a + TRB;
  __shared__ cplx stau[PW];
  const int tid = threadIdx.x;
  for (int e = tid; e < PW; e += blockDim.x) stau[e] = e < pw ? taus[e] : mk(0, 0);
  __syncthreads();
  for (int e = tid; e < PW * PW; e += blockDim.x) {
    const int i = e % PW, p = e / PW;  // M(p,i), upper triangular: p <= i
    cplx v = mk(0, 0);
    const bool zp = (p >= pw) || (stau[p].x == 0.0 && stau[p].y == 0.0);
    const bool zi = (i >= pw) || (stau[i].x == 0.0 && stau[i].y == 0.0);
    if (p == i) {
      if (zp) {
        v = mk(1, 0);
      } else {
        const double dn = 1.0 / cabs2(stau[p]);
        v = mk(stau[p].x * dn, -stau[p].y * dn);  // 1 / tau
      }
    } else if (p < i && !zp && !zi) {
      v = mk(S[2 * (i * PW + p)], S[2 * (i * PW + p) + 1]);
    }
    a[p][i] = v;
  }
  __syncthreads();
  trinv_lower_smem(a, x, PW);
  // T(c,p) = inv(M)(c,p) = inv(M^T)(p,c) = x[p][c]
  for (int e = tid; e < PW * PW; e += blockDim.x) {
    const int c = e % PW, p = e / PW;
    cplx v = (c <= p) ? x[p][c] : mk(0, 0);
    const bool zc = (c >= pw) || (stau[c].x == 0.0 && stau[c].y == 0.0);
    const bool zp = (p >= pw) || (stau[p].x == 0.0 && stau[p].y == 0.0);
    if (zc || zp) v = mk(0, 0);
    T[(long)p * PW + c] = v;
  }
}

__global__ void last_diag_kernel(int n, cplx* A, long lda, double* d) {
  d[n - 1] = A[(long)(n - 1) * lda + n - 1].x;
  for (int r = threadIdx.x; r < n; r += blockDim.x) A[(long)(n - 1) * lda + r] = mk(0, 0);
}

// U0(:, i) = Z(:, perm[i])  (real -> complex)
__global__ void gather_real_cols_kernel(int n, int k, const double* __restrict__ Z, long ldz, const int* __restrict__ perm,
                                        cplx* __restrict__ U, long ldu) {
  const long total = (long)n * k;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int r = (int)(e % n);
    const int i = (int)(e / n);
    U[(long)i * ldu + r] = mk(Z[(long)perm[i] * ldz + r], 0.0);
  }
}

}  // namespace

#ifdef TEVO_PANEL_TIMING
extern "C" int tevo_debug_panel_cycles(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_panel_cycles, sizeof(unsigned long long) * 16);
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_panel_cycles, z, sizeof(z));
  }
  return 0;
}
#endif

// algorithmic work of the tridiagonalisation panels launched so far (bench.py's roofline numerator): factorisations,
// bytes of the matrix-vector products (one read of the trailing matrix per column) and panel launches
std::atomic<long long> g_heig_calls{0}, g_panel_launches{0};
std::atomic<double> g_panel_alg_bytes{0.0};
void heig_counters(double* calls, double* bytes, double* launches) {
  *calls = (double)g_heig_calls.load();
  *bytes = g_panel_alg_bytes.load();
  *launches = (double)g_panel_launches.load();
}

std::atomic<int> g_heig_mode{0};  // 0: default (one-stage unless TEVO_HEIG_TWOSTAGE is set), 1: one-stage, 2: two-stage
void heig_set_mode(int mode) { g_heig_mode.store(mode); }
std::atomic<long long> g_sbr_calls{0};
std::atomic<double> g_sbr_alg_bytes{0.0};
void sbr_counters(double* calls, double* bytes) {
  *calls = (double)g_sbr_calls.load();
  *bytes = g_sbr_alg_bytes.load();
}

void Heig::init() {
  int dev = 0;
  TEVO_CUDA_CHECK(cudaGetDevice(&dev));
  TEVO_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms_, cudaDevAttrMultiProcessorCount, dev));
  // [barrier counter (one 128-byte line) | acc_yv NREP*2*PW | acc_cw NREP*2*PW*PW | acc_cv NREP*2*PW*PW]
  TEVO_CUDA_CHECK(cudaMalloc(&acc_, sizeof(double) * (16 + NREP * 2 * PW + 2 * NREP * 2 * PW * PW)));
  sbr_.init(num_sms_);
  // second stream: the compact-WY set-up of the back-transformation runs beside the divide & conquer (see factor())
  TEVO_CUDA_CHECK(cudaStreamCreateWithFlags(&aux_, cudaStreamNonBlocking));
  TEVO_CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
  TEVO_CUDA_CHECK(cudaEventCreateWithFlags(&ev_join_, cudaEventDisableTiming));
}

void Heig::destroy() {
  if (aux_) {
    cudaStreamSynchronize(aux_);
    cudaStreamDestroy(aux_);
    cudaEventDestroy(ev_fork_);
    cudaEventDestroy(ev_join_);
    aux_ = nullptr;
  }
  stedc_.destroy();
  sbr_.destroy();
  auto fr = [](void* p) {
    if (p) cudaFree(p);
  };
  fr(P_); fr(Q_); fr(Y_); fr(Sall_); fr(VT_); fr(d_); fr(e_); fr(taus_); fr(evals_); fr(Z_); fr(acc_); fr(Tblk_); fr(X_); fr(part_);
  fr(pair_);
  pair_ = nullptr;
  Y_ = nullptr;
  P_ = Q_ = VT_ = nullptr;
  Sall_ = nullptr;
  d_ = e_ = evals_ = Z_ = acc_ = nullptr;
  taus_ = Tblk_ = X_ = part_ = nullptr;
  cap_ = 0;
}

void Heig::ensure(int n) {
  if (n <= cap_) return;
  auto fr = [](void* p) {
    if (p) cudaFree(p);
  };
  fr(P_); fr(Q_); fr(Y_); fr(Sall_); fr(VT_); fr(d_); fr(e_); fr(taus_); fr(evals_); fr(Z_); fr(X_); fr(Tblk_); fr(pair_);
  const size_t nn = (size_t)n * n;
  TEVO_CUDA_CHECK(cudaMalloc(&Tblk_, sizeof(cplx) * PW * PW * (size_t)((n + PW - 1) / PW)));
  TEVO_CUDA_CHECK(cudaMalloc(&P_, sizeof(cplx) * (size_t)n * 2 * PW));
  TEVO_CUDA_CHECK(cudaMalloc(&Q_, sizeof(cplx) * (size_t)n * 2 * PW));
  TEVO_CUDA_CHECK(cudaMalloc(&Y_, sizeof(cplx) * (size_t)n * PW));
  TEVO_CUDA_CHECK(cudaMalloc(&Sall_, sizeof(double) * 2 * PW * PW * (size_t)((n + PW - 1) / PW)));
  TEVO_CUDA_CHECK(cudaMalloc(&VT_, sizeof(cplx) * nn));
  TEVO_CUDA_CHECK(cudaMalloc(&d_, sizeof(double) * n));
  TEVO_CUDA_CHECK(cudaMalloc(&e_, sizeof(double) * n));
  TEVO_CUDA_CHECK(cudaMalloc(&taus_, sizeof(cplx) * ((size_t)n + PW)));
  TEVO_CUDA_CHECK(cudaMalloc(&evals_, sizeof(double) * n));
  TEVO_CUDA_CHECK(cudaMalloc(&Z_, sizeof(double) * nn));
  TEVO_CUDA_CHECK(cudaMalloc(&X_, sizeof(cplx) * (size_t)2 * PW * n));
  TEVO_CUDA_CHECK(cudaMalloc(&pair_, sizeof(cplx) * 3 * PW * PW * (size_t)(n / (2 * PW) + 1)));
  xcap_ = n;
  cap_ = n;
}

void Heig::factor(cudaStream_t st, int n, cplx* A, long lda) {
  if (n < 1) throw TevoError(1, "heig: empty matrix");
  ensure(n);
  n_ = n;
  A_ = A;
  lda_ = lda;
  static PerDeviceOnce once;
  once([&]() {
    TEVO_CUDA_CHECK(cudaFuncSetAttribute(tridiag_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM));
    TEVO_CUDA_CHECK(cudaFuncSetAttribute(wy_T_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(2 * sizeof(cplx) * TRB * TRB)));
  });
  if (n > 4096) throw TevoError(1, "heig: n > 4096 not supported");
  g_heig_calls.fetch_add(1, std::memory_order_relaxed);
  if (wy_pending_) {  // a previous factorisation whose vectors() was never called: its WY set-up still owns the buffers
    TEVO_CUDA_CHECK(cudaStreamWaitEvent(st, ev_join_, 0));
    wy_pending_ = false;
  }
  double* zero_from = acc_;  // the barrier counter and the replicated accumulators are cleared before every panel
  const size_t zero_bytes = sizeof(double) * (16 + NREP * 2 * PW + 2 * NREP * 2 * PW * PW);
  const cplx ONE = mk(1, 0), MONE = mk(-1, 0), ZERO = mk(0, 0);
  TEVO_CUDA_CHECK(cudaMemsetAsync(Sall_, 0, sizeof(double) * 2 * PW * PW * (size_t)((n + PW - 1) / PW), st));
  // the last columns go to the cluster-resident kernel (tail.cu) once the trailing matrix fits the shared memory of
  // one 16-CTA cluster: validated on the B200 with the eigensolver / split tests at cluster sizes 1, 8 and 16 and
  // measured at -5 % (n = 2048) / -14 % (n = 1024) of the panel time (profiles/r02_deferred_checks.log).
  // TEVO_CLUSTER_TAIL=0 switches it off, another value selects the cluster size.
  static const int tail_cluster = [] {
    const char* e = getenv("TEVO_CLUSTER_TAIL");
    return e ? atoi(e) : 16;
  }();
  static const int tail_max = tail_cluster > 0 ? tridiag_tail_max_m(tail_cluster) : 0;
  int tail_t0 = -1;
  int p0 = 0;
  // default: the one-stage panel kernel below.  The two-stage reduction (sbr.cu: dense -> band of width 8 ->
  // tridiagonal) is parity-green but not yet faster: TEVO_HEIG_TWOSTAGE=1 or
  // tevo_debug_set_heig_mode(2) selects it.
  static const bool env_two_stage = getenv("TEVO_HEIG_TWOSTAGE") != nullptr;
  two_stage_ = (env_two_stage || g_heig_mode.load() == 2) && g_heig_mode.load() != 1 && n >= 2 && n <= SBR_MAX_N;
  if (two_stage_) {
    {
      ProfScope ps(st, P_SBR_STAGE1);
      sbr_.reduce_to_band(st, n, A, lda, taus_, max_ctas_);
    }
    {
      ProfScope ps(st, P_SBR_CHASE);
      sbr_.chase(st, n, d_, e_);
    }
    g_sbr_calls.fetch_add(1, std::memory_order_relaxed);
    double by = 0.0;
    for (int j = 0; j + SBR_BAND <= n - 2; j += SBR_BAND) by += 32.0 * (double)(n - j - SBR_BAND) * (double)(n - j - SBR_BAND);
    double cur = g_sbr_alg_bytes.load(std::memory_order_relaxed);
    while (!g_sbr_alg_bytes.compare_exchange_weak(cur, cur + by, std::memory_order_relaxed)) {
    }
    tail_t0 = 0;  // V^H V of every 64-column panel by GEMM below
    p0 = n;       // skip the one-stage loop
  }
  while (p0 < n - 1) {
    // single-lane configuration and matrices of at least 128 rows only: the lanes of the layer-parallel mode keep the
    // panel kernel (several 16-CTA clusters next to capped cooperative grids were not validated), and tiny problems
    // gain nothing from a cluster
    if (tail_max > 0 && max_ctas_ == 0 && n >= 128 && n - p0 <= tail_max && n - p0 >= 2) {
      ProfScope ps(st, P_TRI_PANEL);
      tridiag_tail(st, tail_cluster, n, p0, A, lda, d_, e_, taus_);
      tail_t0 = p0;
      break;
    }
    const int pw = std::min(PW, n - 1 - p0);
    TEVO_CUDA_CHECK(cudaMemsetAsync(zero_from, 0, zero_bytes, st));
    PanelArgs pa;
    pa.n = n; pa.p0 = p0; pa.pw = pw; pa.A = A; pa.lda = lda; pa.P = P_; pa.Q = Q_; pa.Y = Y_; pa.ldp = n;
    pa.bar = reinterpret_cast<unsigned*>(zero_from);
    pa.acc_yv = zero_from + 16;
    pa.acc_cw = pa.acc_yv + NREP * 2 * PW;
    pa.acc_cv = pa.acc_cw + NREP * 2 * PW * PW;
    pa.cv_out = Sall_ + (size_t)(p0 / PW) * 2 * PW * PW;
    pa.d = d_; pa.e = e_; pa.taus = taus_;
    const int nstrips = (n - p0 + RS - 1) / RS;
    static const int env_cap = [] {
      const char* e = getenv("TEVO_PANEL_MAX_CTAS");  // development knob: forces the multi-strip work split
      return e ? atoi(e) : 0;
    }();
    int cap = (max_ctas_ > 0) ? std::min(max_ctas_, num_sms_) : num_sms_;
    if (env_cap > 0) cap = std::min(cap, env_cap);
    const int capc = std::max(1, std::min(cap, MAXCTA));
    int grid = capc, chunks = 1;
    if (nstrips < capc) {  // fewer strips than SMs: split the columns of a strip (>= 8 groups of 16 columns per chunk)
      chunks = std::max(1, std::min(capc / nstrips, nstrips / 8));
      grid = nstrips * chunks;
    }
    pa.chunks = chunks;
    void* args[] = {&pa};
    const size_t sv_bytes = (size_t)(((n - p0 + 1) & ~1) + 32) * sizeof(cplx);
    const size_t strip_bytes = (size_t)(2 * PW * RS + 2 * RS) * sizeof(cplx);
    const int per_cta = (nstrips + grid - 1) / grid;
    static const size_t panel_smem = [] {
      const char* e = getenv("TEVO_PANEL_SMEM");  // development knob: shared-memory budget in bytes
      return e ? std::min<size_t>(PANEL_SMEM, (size_t)atol(e)) : (size_t)PANEL_SMEM;
    }();
    pa.smax = std::min(per_cta, (int)((panel_smem - sv_bytes) / strip_bytes));
    size_t smem = sv_bytes + (size_t)pa.smax * strip_bytes;
    pa.ncache = 0;
    if (per_cta == 1) {
      const int groups_per_cta = (nstrips + chunks - 1) / chunks;
      pa.ncache = (panel_smem > smem) ? std::min(groups_per_cta, (int)((panel_smem - smem) / (NT * sizeof(cplx)))) : 0;
      smem += (size_t)pa.ncache * NT * sizeof(cplx);
    }
    {
      ProfScope ps(st, P_TRI_PANEL);
      TEVO_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)tridiag_panel_kernel, dim3(grid), dim3(NT), args, smem, st));
      TEVO_LAUNCHED();
      g_panel_launches.fetch_add(1, std::memory_order_relaxed);
      double by = 0.0;
      for (int i = 0; i < pw; ++i) by += 16.0 * (double)(n - p0 - i - 1) * (double)(n - p0 - i - 1);
      double cur = g_panel_alg_bytes.load(std::memory_order_relaxed);
      while (!g_panel_alg_bytes.compare_exchange_weak(cur, cur + by, std::memory_order_relaxed)) {
      }
    }
    const int q0 = p0 + pw;
    // trailing rank-2k update  A22 -= V W^H + W V^H = [V|W] [W|V]^H
    ProfScope ps2(st, P_TRI_HER2K);
    if (pw == PW) {
      zgemm_herm(st, OP_N, OP_C, n - q0, 2 * PW, MONE, P_ + q0, n, Q_ + q0, n, ONE, A + (long)q0 * lda + q0, lda, true);
    } else {
      zgemm(st, OP_N, OP_C, n - q0, n - q0, pw, MONE, P_ + q0, n, Q_ + q0, n, ONE, A + (long)q0 * lda + q0, lda);
      zgemm(st, OP_N, OP_C, n - q0, n - q0, pw, MONE, P_ + (long)PW * n + q0, n, Q_ + (long)PW * n + q0, n, ONE,
            A + (long)q0 * lda + q0, lda);
    }
    p0 = q0;
  }
  if (!two_stage_) {
    last_diag_kernel<<<1, 256, 0, st>>>(n, A, lda, d_);
    TEVO_LAUNCHED();
  }
  // The compact-WY set-up below (T factors, V*T, pair factors: ~0.9 ms of small GEMMs at n = 2048) only needs the
  // reflectors; the divide & conquer (~3.1 ms of latency-bound stages) only needs d and e, so they run side by side
  // (WY work on a second stream, joined by vectors()).  Measured: -3.8 % per sequential TEBD4 step at chi = 1024
  // (profiles/r02_wy_overlap.log), on by default since the end of round 2 for the single-lane configuration
  // (max_ctas_ == 0: sequential TEBD, TDVP, centre moves).  The lanes of the layer-parallel mode never fork: their
  // concurrency comes from the other lanes, and an early version that forked inside the lanes hit an illegal memory
  // access that was not tracked down.  TEVO_WY_OVERLAP=0 switches the fork off.
  static const bool wy_overlap = [] {
    const char* e = getenv("TEVO_WY_OVERLAP");
    return e ? atoi(e) != 0 : true;
  }();
  cudaStream_t main_st = st;
  if (wy_overlap && aux_ && max_ctas_ == 0 && n >= 256) {
    TEVO_CUDA_CHECK(cudaEventRecord(ev_fork_, main_st));
    TEVO_CUDA_CHECK(cudaStreamWaitEvent(aux_, ev_fork_, 0));
    st = aux_;
  }
  if (tail_t0 >= 0) {
    // the tail kernel keeps no panel dots: V^H V of its panels (the input of the compact-WY factors) by GEMM
    const cplx ONE1 = mk(1, 0), ZERO1 = mk(0, 0);
    cplx* S = reinterpret_cast<cplx*>(Sall_) + (size_t)(tail_t0 / PW) * PW * PW;
    const cplx* Vt = A + (long)tail_t0 * lda;
    const int nfull_tail = (n - tail_t0) / PW, wlast = (n - tail_t0) % PW;
    if (nfull_tail > 0)
      zgemm_batched(st, OP_C, OP_N, PW, PW, n, ONE1, Vt, lda, (long)PW * lda, Vt, lda, (long)PW * lda, ZERO1, S, PW,
                    (long)PW * PW, nfull_tail);
    if (wlast > 0)
      zgemm(st, OP_C, OP_N, wlast, wlast, n, ONE1, Vt + (long)nfull_tail * PW * lda, lda, Vt + (long)nfull_tail * PW * lda,
            lda, ZERO1, S + (size_t)nfull_tail * PW * PW, PW);
  }
  if (n > 1) {
    // compact-WY factors of all panels in one launch, then V*T for the back-transformation as one batched GEMM over
    // the full-width panels (A holds V with explicit zeros above the reflectors) plus one for a narrower last panel
    ProfScope ps(st, P_TRI_WY);
    const int npanels = (n - 1 + PW - 1) / PW;
    wy_T_kernel<<<npanels, 1024, 2 * sizeof(cplx) * TRB * TRB, st>>>(n, taus_, Sall_, Tblk_);
    TEVO_LAUNCHED();
    const int nfull = n / PW;  // panels whose PW columns all exist
    if (nfull > 0)
      zgemm_batched(st, OP_N, OP_N, n, PW, PW, ONE, A, lda, (long)PW * lda, Tblk_, PW, (long)PW * PW, ZERO, VT_, n,
                    (long)PW * n, nfull);
    if (nfull < npanels) {
      const int q0 = nfull * PW;
      zgemm(st, OP_N, OP_N, n, n - q0, n - q0, ONE, A + (long)q0 * lda, lda, Tblk_ + (long)nfull * PW * PW, PW, ZERO,
            VT_ + (long)q0 * n, n);
    }
    // pairs of adjacent full-width panels become one 128-wide block reflector for the back-transformation:
    //   H_a H_b = I - [V_a V_b] [[T_a, T_ab], [0, T_b]] [V_a V_b]^H,   T_ab = -T_a (V_a^H V_b) T_b,
    // so (V T)_b gains V_a T_ab.  (A 64-row operand fills only half of the 128-row GEMM tile.)
    static const bool no_pairs = getenv("TEVO_BACK64") != nullptr;
    npairs_ = no_pairs ? 0 : nfull / 2;
    if (npairs_ > 0) {
      const long bsA = 2L * PW * lda, bsT = 2L * PW * PW, bsP = (long)PW * PW;
      cplx* S12 = pair_;
      cplx* W1 = pair_ + (size_t)npairs_ * PW * PW;
      cplx* T12 = pair_ + (size_t)2 * npairs_ * PW * PW;
      zgemm_batched(st, OP_C, OP_N, PW, PW, n, ONE, A, lda, bsA, A + (long)PW * lda, lda, bsA, ZERO, S12, PW, bsP, npairs_);
      zgemm_batched(st, OP_N, OP_N, PW, PW, PW, ONE, Tblk_, PW, bsT, S12, PW, bsP, ZERO, W1, PW, bsP, npairs_);
      zgemm_batched(st, OP_N, OP_N, PW, PW, PW, MONE, W1, PW, bsP, Tblk_ + bsP, PW, bsT, ZERO, T12, PW, bsP, npairs_);
      zgemm_batched(st, OP_N, OP_N, n, PW, PW, ONE, A, lda, bsA, T12, PW, bsP, ONE, VT_ + (long)PW * n, n, 2L * PW * n,
                    npairs_);
    }
  }
  if (st != main_st) {
    TEVO_CUDA_CHECK(cudaEventRecord(ev_join_, st));
    wy_pending_ = true;
    st = main_st;
  }
  {
    ProfScope ps(st, P_STEDC);
    stedc_.run(st, n, d_, e_, evals_, Z_);
  }
}

void Heig::vectors(cudaStream_t st, const int* perm_dev, int k, cplx* U, long ldu) {
  const int n = n_;
  if (wy_pending_) {  // join the compact-WY set-up that ran beside the divide & conquer
    TEVO_CUDA_CHECK(cudaStreamWaitEvent(st, ev_join_, 0));
    wy_pending_ = false;
  }
  if (k < 1) return;
  ProfScope ps(st, P_BACK);
  const cplx ONE = mk(1, 0), MONE = mk(-1, 0), ZERO = mk(0, 0);
  if (two_stage_) {
    ProfScope ps2(st, P_SBR_Q2);
    sbr_.apply_q2(st, n, k, Z_, n, perm_dev, U, ldu);
  } else {
    const long total = (long)n * k;
    int blocks = (int)std::min<long>((total + 255) / 256, 148 * 16);
    gather_real_cols_kernel<<<blocks, 256, 0, st>>>(n, k, Z_, n, perm_dev, U, ldu);
    TEVO_LAUNCHED();
  }
  if (n < 2) return;
  if ((size_t)k > xcap_) {
    if (X_) cudaFree(X_);
    X_ = nullptr;
    TEVO_CUDA_CHECK(cudaMalloc(&X_, sizeof(cplx) * (size_t)2 * PW * k));
    xcap_ = k;
  }
  // U = H_0 ... H_{n-2} Z : apply the panels last to first,  U -= (V T) (V^H U); the panels below 2*npairs_ go in
  // pairs (128 columns of V and V T at a time), the ones above singly
  int last = ((n - 2) / PW) * PW;
  for (int p0 = last; p0 >= 2 * npairs_ * PW; p0 -= PW) {
    const int pw = std::min(PW, n - 1 - p0);
    const int rows = n - p0;
    zgemm_splitk(st, OP_C, OP_N, pw, k, rows, ONE, A_ + (long)p0 * lda_ + p0, lda_, U + p0, ldu, ZERO, X_, PW, &part_,
                 &part_cap_);
    zgemm(st, OP_N, OP_N, rows, k, pw, MONE, VT_ + (long)p0 * n + p0, n, X_, PW, ONE, U + p0, ldu);
  }
  for (int q = npairs_ - 1; q >= 0; --q) {
    const int p0 = 2 * q * PW, rows = n - p0;
    zgemm_splitk(st, OP_C, OP_N, 2 * PW, k, rows, ONE, A_ + (long)p0 * lda_ + p0, lda_, U + p0, ldu, ZERO, X_, 2 * PW,
                 &part_, &part_cap_);
    zgemm(st, OP_N, OP_N, rows, k, 2 * PW, MONE, VT_ + (long)p0 * n + p0, n, X_, 2 * PW, ONE, U + p0, ldu);
  }
}

}  // namespace tevo
