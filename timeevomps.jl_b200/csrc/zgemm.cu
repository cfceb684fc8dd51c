// Complex-FP64 GEMM on the FP64 tensor pipe (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4), sm_100a.
//
// C(MxN, col-major, ldc) = alpha * op(A)(MxK) * op(B)(KxN) + beta * C,   op in {N, T, C}
//
// This is the ZGEMM-class workhorse behind theta = A_b * A_{b+1}  (src/bondgate.jl:49-50, src/tdvp.jl:60),
// the split-back GEMMs of replacebond! (src/bondgate.jl:51, src/tdvp.jl:64), the centre-move products of
// orthogonalize! (src/bondgate.jl:48) and the L*theta / T*R legs of ProjMPO's product (src/tdvp.jl:61,81).
//
// Design: interleaved (re,im) operands are staged by cp.async (16 B = one complex element per copy, zero-fill
// out of range) through a 3-stage shared-memory ring; one LDS.128 fetches the (re,im) pair a thread needs for
// an m8n8k4 fragment, so no planar repack is required.  A complex MAC is 4 real DMMAs:
//   Cre += Are*Bre ; Cre += (-Aim)*Bim ; Cim += Are*Bim ; Cim += Aim*Bre .
// CTA tile 128x64, 8 warps, warp tile 32x32 (16 DMMA blocks x re/im accumulators = 64 fp64 regs/thread); smaller
// problems use 64x32 or 32x32 tiles of the same kernel (pick_tile below).
#include <atomic>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "zgemm.h"

namespace tevo {

namespace {

constexpr int BM = 128, BN = 64, BK = 16, STAGES = 3, NTHREADS = 256;
constexpr int KPAD = 4;  // k-contiguous tiles: row stride (BK+4)*16 B = 320 B -> odd multiple of 64 B
constexpr int MPAD = 2;  // m-contiguous tiles: k stride (ROWS+2)*16 B == 32 B mod 128 B

template <int ROWS, bool KMAJOR>
struct Tile {
  static constexpr int elems = KMAJOR ? ROWS * (BK + KPAD) : BK * (ROWS + MPAD);
  __device__ static inline int idx(int r, int k) { return KMAJOR ? r * (BK + KPAD) + k : k * (ROWS + MPAD) + r; }
};

__device__ inline void cp_async16(void* smem_dst, const void* gsrc, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}
__device__ inline void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ inline void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---- operand staging by the TMA engine: cp.async.bulk (SASS UBLKCP) + mbarrier transaction counts ----
__device__ inline void mbar_init(unsigned long long* bar, unsigned count) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count));
}
__device__ inline void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ inline void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(a),
      "r"(parity)
      : "memory");
}
__device__ inline void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d),
               "l"(gsrc), "r"(bytes), "r"(b)
               : "memory");
}

// Stage a full (interior) ROWS x BK operand tile with bulk copies issued by one warp: one contiguous run per k
// (m-contiguous operands: ROWS*16 B) or per row (k-contiguous operands: BK*16 B) into the padded shared-memory rows.
template <int ROWS, bool KMAJOR>
__device__ inline void bulk_tile(cplx* s, const cplx* __restrict__ g, long ld, int r0, int k0, unsigned long long* bar,
                                 int lane) {
  typedef Tile<ROWS, KMAJOR> T;
  if (KMAJOR) {
    for (int r = lane; r < ROWS; r += 32) bulk_g2s(&s[T::idx(r, 0)], g + (long)(r0 + r) * ld + k0, BK * 16, bar);
  } else {
    for (int k = lane; k < BK; k += 32) bulk_g2s(&s[T::idx(0, k)], g + (long)(k0 + k) * ld + r0, ROWS * 16, bar);
  }
}

__device__ inline void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Stage a ROWS x BK tile of an operand.  KMAJOR: element (r,k) at g[k + r*ld]; else at g[r + k*ld].
template <int ROWS, bool KMAJOR>
__device__ inline void load_tile(cplx* s, const cplx* __restrict__ g, long ld, int r0, int k0, int R, int K) {
  typedef Tile<ROWS, KMAJOR> T;
#pragma unroll
  for (int e = threadIdx.x; e < ROWS * BK; e += NTHREADS) {
    int r, k;
    if (KMAJOR) {
      r = e / BK;
      k = e % BK;
    } else {
      k = e / ROWS;
      r = e % ROWS;
    }
    bool p = (r0 + r < R) && (k0 + k < K);
    const cplx* src = g;
    if (p) src = KMAJOR ? g + (long)(r0 + r) * ld + (k0 + k) : g + (long)(k0 + k) * ld + (r0 + r);
    cp_async16(&s[T::idx(r, k)], src, p);
  }
}

// CONJA / CONJB: 0 / 1 = compile-time conjugation (default since round 2: +4.6 % at the theta shape, tests green on the
// B200, profiles/r02_checks.log); -1 = sign taken from the run-time arguments conjA / conjB (TEVO_ZGEMM_CTSIGN=0, 128 x 64
// tile only).  With compile-time signs the signs of the imaginary parts fold into the operand-negation modifiers of DMMA instead of costing 12 FP64 multiplies/negations per
// k-slice on the pipe the DMMAs use (SASS: the DADD/FSEL/DMUL of the inner loop disappear).
// BULK: full interior tiles (and K a multiple of BK) are staged by the TMA engine (cp.async.bulk + mbarrier) instead
// of one 16-byte cp.async per element; edge tiles keep the zero-filling cp.async path.
// TBM x TBN: CTA tile (128 x 64 for problems that fill the GPU, 64 x 32 / 32 x 32 for small ones, see pick_tile): 8 warps
// as 4 x 2, warp tile (TBM/4) x (TBN/2).
// M3: three real products per complex one (see the kernel body); small tiles only.
template <int TBM, int TBN, bool A_KMAJOR, bool B_KMAJOR, int CONJA, int CONJB, bool BULK, bool M3>
__global__ void __launch_bounds__(NTHREADS, (TBM == 128 ? 1 : (TBM == 64 ? 2 : 3)))
zgemm_dmma_kernel(int M, int N, int Kfull, cplx alpha, const cplx* __restrict__ A, long lda, int conjA,
                  const cplx* __restrict__ B, long ldb, int conjB, cplx beta, cplx* __restrict__ C, long ldc, int Kc,
                  long part_stride, int nsplit, long bsA, long bsB, long bsC, int lower_only) {
  // blockIdx.z = batch * nsplit + split.  split-K: split s handles k in [s*Kc, min(Kfull, (s+1)*Kc)) and writes its
  // partial product to C + s*part_stride; batched: operands of batch b are offset by b*bsA / b*bsB / b*bsC
  const int zsplit = blockIdx.z % nsplit, zbatch = blockIdx.z / nsplit;
  A += (long)zbatch * bsA;
  B += (long)zbatch * bsB;
  C += (long)zbatch * bsC;
  const int kb = zsplit * Kc;
  const int K = min(Kfull - kb, Kc);
  A += A_KMAJOR ? (long)kb : (long)kb * lda;
  B += B_KMAJOR ? (long)kb : (long)kb * ldb;
  C += (long)zsplit * part_stride;
  typedef Tile<TBM, A_KMAJOR> TA;
  typedef Tile<TBN, B_KMAJOR> TB;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* sA = reinterpret_cast<cplx*>(smem_raw);
  cplx* sB = sA + STAGES * TA::elems;

  constexpr int MI = TBM / 32, NJ = TBN / 16;  // 8 x 8 DMMA blocks per warp tile
  const int m0 = blockIdx.x * TBM, n0 = blockIdx.y * TBN;
  if (lower_only && n0 >= m0 + TBM) return;  // Hermitian result: tiles strictly above the diagonal are mirrored later
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = (warp & 3) * (TBM / 4), wn = (warp >> 2) * (TBN / 2);
  const int g = lane >> 2, tq = lane & 3;

  // 4M (default for the big tile): cre += Are Bre - Aim Bim, cim += Are Bim + Aim Bre - four DMMAs per complex MAC.
  // 3M (M3): t1 = Are Bre, t2 = Aim Bim, t3 = (Are + Aim)(Bre + Bim); Re = t1 - t2, Im = t3 - t1 - t2 - three DMMAs and
  // two FP64 adds per fragment.  The kernel is bound by DMMA issue (ncu: DMMA sub-pipe 85 % busy, math-pipe throttle
  // the top stall), so the product count is what sets its time.  Error: normwise like 4M, eps |A||B| with a constant
  // about twice as large for the imaginary part (Higham, "Stability of a method for multiplying complex matrices with
  // three real matrix multiplications"); the three sums are combined once, after the k loop.
  double cre[MI][NJ][2], cim[MI][NJ][2], ct3[M3 ? MI : 1][M3 ? NJ : 1][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      cre[i][j][0] = cre[i][j][1] = 0.0;
      cim[i][j][0] = cim[i][j][1] = 0.0;
      if (M3) ct3[i][j][0] = ct3[i][j][1] = 0.0;
    }

  const int nk = (K + BK - 1) / BK;
  __shared__ unsigned long long full_bar[STAGES];
  // operands whose tile rows are long contiguous runs (m- / n-contiguous: 2 KB / 1 KB per k) go through the TMA
  // engine; k-contiguous operands (256-byte runs) stay on cp.async, where bulk copies measured slower
  const bool interior = BULK && (m0 + TBM <= M) && (n0 + TBN <= N) && (K % BK == 0);
  const bool bulkA = interior && !A_KMAJOR, bulkB = interior && !B_KMAJOR;
  const bool bulk = bulkA || bulkB;
  const unsigned stage_bytes = (unsigned)(((bulkA ? TBM : 0) + (bulkB ? TBN : 0)) * BK * sizeof(cplx));
  if (bulk) {
    if (threadIdx.x == 0) {
#pragma unroll
      for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
  }
  // fills stage st with k-tile nx (bulk part: warp 0 issues; the stage was last read before the preceding CTA barrier)
  auto fill = [&](int st, int nx) {
    if (bulk && warp == 0) {
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      if (lane == 0) mbar_expect_tx(&full_bar[st], stage_bytes);
      __syncwarp();
      if (bulkA) bulk_tile<TBM, A_KMAJOR>(sA + st * TA::elems, A, lda, m0, nx * BK, &full_bar[st], lane);
      if (bulkB) bulk_tile<TBN, B_KMAJOR>(sB + st * TB::elems, B, ldb, n0, nx * BK, &full_bar[st], lane);
    }
    if (!bulkA) load_tile<TBM, A_KMAJOR>(sA + st * TA::elems, A, lda, m0, nx * BK, M, K);
    if (!bulkB) load_tile<TBN, B_KMAJOR>(sB + st * TB::elems, B, ldb, n0, nx * BK, N, K);
  };
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) fill(s, s);
    cp_async_commit();
  }

  const double sa = conjA ? -1.0 : 1.0, sb = conjB ? -1.0 : 1.0;
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    if (bulk) mbar_wait(&full_bar[kt % STAGES], (unsigned)((kt / STAGES) & 1));
    __syncthreads();
    {
      const int nx = kt + STAGES - 1;
      if (nx < nk) fill(nx % STAGES, nx);
      cp_async_commit();
    }
    const cplx* a_s = sA + (kt % STAGES) * TA::elems;
    const cplx* b_s = sB + (kt % STAGES) * TB::elems;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double are[MI], aim[MI], nai[MI], bre[NJ], bim[NJ];
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        cplx a = a_s[TA::idx(wm + i * 8 + g, kk + tq)];
        are[i] = a.x;
        aim[i] = (CONJA < 0) ? sa * a.y : (CONJA ? -a.y : a.y);
        nai[i] = -aim[i];
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        cplx b = b_s[TB::idx(wn + j * 8 + g, kk + tq)];
        bre[j] = b.x;
        bim[j] = (CONJB < 0) ? sb * b.y : (CONJB ? -b.y : b.y);
      }
      if (M3) {
        // cre <- t1, cim <- t2, ct3 <- t3
        double asum[MI], bsum[NJ];
#pragma unroll
        for (int i = 0; i < MI; ++i) asum[i] = are[i] + aim[i];
#pragma unroll
        for (int j = 0; j < NJ; ++j) bsum[j] = bre[j] + bim[j];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j) dmma(cre[i][j][0], cre[i][j][1], are[i], bre[j]);
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j) dmma(cim[i][j][0], cim[i][j][1], aim[i], bim[j]);
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j) dmma(ct3[M3 ? i : 0][M3 ? j : 0][0], ct3[M3 ? i : 0][M3 ? j : 0][1], asum[i], bsum[j]);
      } else {
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j) dmma(cre[i][j][0], cre[i][j][1], are[i], bre[j]);
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j) dmma(cim[i][j][0], cim[i][j][1], are[i], bim[j]);
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j) dmma(cre[i][j][0], cre[i][j][1], nai[i], bim[j]);
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j) dmma(cim[i][j][0], cim[i][j][1], aim[i], bre[j]);
      }
    }
  }
  cp_async_wait<0>();

  const bool has_beta = (beta.x != 0.0 || beta.y != 0.0);
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int row = m0 + wm + i * 8 + g;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = n0 + wn + j * 8 + tq * 2 + e;
        if (col >= N) continue;
        const double pre = M3 ? cre[i][j][e] - cim[i][j][e] : cre[i][j][e];
        const double pim = M3 ? (ct3[M3 ? i : 0][M3 ? j : 0][e] - cre[i][j][e]) - cim[i][j][e] : cim[i][j][e];
        cplx v = cmul(alpha, mk(pre, pim));
        cplx* dst = C + (long)col * ldc + row;
        if (has_beta) v = cadd(v, cmul(beta, *dst));
        *dst = v;
      }
    }
  }
}

template <int TBM, int TBN, bool AK, bool BK_, int CA, int CB, bool BULK, bool M3 = false>
void launch_cb(cudaStream_t st, int M, int N, int K, cplx alpha, const cplx* A, long lda, int conjA, const cplx* B,
              long ldb, int conjB, cplx beta, cplx* C, long ldc, int splits, int Kc, long part_stride, int batch,
              long bsA, long bsB, long bsC, int lower_only) {
  typedef Tile<TBM, AK> TA;
  typedef Tile<TBN, BK_> TB;
  const size_t smem = (size_t)STAGES * (TA::elems + TB::elems) * sizeof(cplx);
  static PerDeviceOnce once;
  once([&]() {
    TEVO_CUDA_CHECK(cudaFuncSetAttribute(zgemm_dmma_kernel<TBM, TBN, AK, BK_, CA, CB, BULK, M3>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  });
  dim3 grid((M + TBM - 1) / TBM, (N + TBN - 1) / TBN, splits * batch);
  zgemm_dmma_kernel<TBM, TBN, AK, BK_, CA, CB, BULK, M3><<<grid, NTHREADS, smem, st>>>(
      M, N, K, alpha, A, lda, conjA, B, ldb, conjB, beta, C, ldc, splits > 1 ? Kc : K, part_stride, splits, bsA, bsB, bsC,
      lower_only);
  TEVO_LAUNCHED();
}

// operand staging: 1 = TMA-engine bulk copies for interior tiles (TEVO_ZGEMM_BULK, see zgemm_bulk_default), 0 = cp.async
int zgemm_bulk_default() {
  static const int v = [] {
    const char* e = getenv("TEVO_ZGEMM_BULK");
    return e ? atoi(e) : 1;
  }();
  return v;
}

// CTA tile for a problem.  One SM delivers 1/148 of the FP64 tensor throughput, so a GEMM with a handful of 128 x 64
// tiles is bound by the time ONE SM needs for its tile (17 us at K = 64), however small the product is - the panel /
// trailing GEMMs of the blocked Cholesky chain, everything at chi <= 256.  And the smaller tiles are not slower per
// flop where the big one has few waves: 2 - 3 CTAs per SM (116 / 80 registers, 92 / 61 KB of shared memory) hide the
// DMMA and shared-memory latencies that 8 warps per SM cannot, and the last wave is fuller.  Measured on the B200
// (tools/dev_zgemm.py with each tile forced, tools/dev_zgemm_tiles.py; profiles/r02_zgemm_tiles.json), TF/s for
// 128x64 / 64x32 / 32x32:  theta NN 2048x2048x1024 28.1 / 31.3 / 30.2;  R = U^H theta CN 27.6 / 30.6 / 29.5;  rho
// Hermitian NC 2048^2 x 2048 29.3 / 27.0 / 29.2;  rank-2k Hermitian 1920^2 x 128 16.4 / 19.6 / 21.9;  back-transformation
// NN 2048x1024x128 20.9 / 25.3 / 24.3;  4096^3 32.5 / 31.9 / 31.1;  Cholesky panel NC 960x64x64 24.6 / 10.3 / 6.2 us;
// chi = 256 theta NN 512x512x256 82 / 27 / 26 us.  Hence:
//   Hermitian results (lower tiles only): 32 x 32 - the diagonal tiles waste the least;
//   general results: 128 x 64 from 1024 such tiles on, else 64 x 32 if that gives a full wave at 2 CTAs per SM, else 32 x 32.
// Returns 0: 128 x 64, 1: 64 x 32, 2: 32 x 32.  zgemm_force_tile(0|1|2) (tests; TEVO_ZGEMM_TILE at start-up) forces one.
std::atomic<int> g_forced_tile{[] {
  const char* e = getenv("TEVO_ZGEMM_TILE");
  return e ? atoi(e) : -1;
}()};
// three-multiplication complex products in the small-tile kernels: on by default (TEVO_ZGEMM_3M=0 / zgemm_set_3m(0):
// four multiplications everywhere); a ZgemmPrecise scope on the calling thread forces four (the Rayleigh-quotient
// GEMM of the split, whose small rows carry the relative accuracy of the small Schmidt weights)
std::atomic<int> g_zgemm_3m{[] {
  const char* e = getenv("TEVO_ZGEMM_3M");
  return e ? atoi(e) : 1;
}()};
thread_local int tl_zgemm_precise = 0;
int pick_tile(int M, int N, int zdim, int lower_only, bool m3) {
  const int forced = g_forced_tile.load(std::memory_order_relaxed);
  if (forced >= 0 && forced <= 2) return forced;
  auto tiles = [&](int bm, int bn) { return (long)((M + bm - 1) / bm) * ((N + bn - 1) / bn) * zdim; };
  // with three multiplications per product the small tiles win at every size (theta shape 36.6 TF/s against 32.4 for
  // the 128 x 64 kernel at 4096^3); the big tile is for four-multiplication products of at least 1024 tiles
  if (!m3 && tiles(BM, BN) >= 1024) return 0;
  if (lower_only) return 2;
  if (tiles(64, 32) >= 296) return 1;
  return 2;
}

template <bool AK, bool BK_, int CA, int CB>
void launch_c(cudaStream_t st, int M, int N, int K, cplx alpha, const cplx* A, long lda, int conjA, const cplx* B,
              long ldb, int conjB, cplx beta, cplx* C, long ldc, int splits, int Kc, long part_stride, int batch,
              long bsA, long bsB, long bsC, int lower_only) {
#define TEVO_ZG_ARGS st, M, N, K, alpha, A, lda, conjA, B, ldb, conjB, beta, C, ldc, splits, Kc, part_stride, batch, bsA, \
                     bsB, bsC, lower_only
  // the small tiles exist with compile-time conjugation signs and cp.async staging only (their operand rows are too
  // short for the bulk copies to pay)
  const bool m3 = g_zgemm_3m.load(std::memory_order_relaxed) != 0 && tl_zgemm_precise == 0;
  const int tile = (CA < 0) ? 0 : pick_tile(M, N, splits * batch, lower_only, m3);
  if (tile == 1) {
    if (m3)
      launch_cb<64, 32, AK, BK_, CA < 0 ? 0 : CA, CB < 0 ? 0 : CB, false, true>(TEVO_ZG_ARGS);
    else
      launch_cb<64, 32, AK, BK_, CA < 0 ? 0 : CA, CB < 0 ? 0 : CB, false>(TEVO_ZG_ARGS);
  } else if (tile == 2) {
    if (m3)
      launch_cb<32, 32, AK, BK_, CA < 0 ? 0 : CA, CB < 0 ? 0 : CB, false, true>(TEVO_ZG_ARGS);
    else
      launch_cb<32, 32, AK, BK_, CA < 0 ? 0 : CA, CB < 0 ? 0 : CB, false>(TEVO_ZG_ARGS);
  } else if (zgemm_bulk_default()) {
    launch_cb<BM, BN, AK, BK_, CA, CB, true>(TEVO_ZG_ARGS);
  } else {
    launch_cb<BM, BN, AK, BK_, CA, CB, false>(TEVO_ZG_ARGS);
  }
#undef TEVO_ZG_ARGS
}

template <bool AK, bool BK_>
void launch(cudaStream_t st, int M, int N, int K, cplx alpha, const cplx* A, long lda, int conjA, const cplx* B,
            long ldb, int conjB, cplx beta, cplx* C, long ldc, int splits = 1, int Kc = 0, long part_stride = 0,
            int batch = 1, long bsA = 0, long bsB = 0, long bsC = 0, int lower_only = 0) {
  // compile-time conjugation signs (fold into DMMA operand modifiers): +4.6 % at the theta shape, tests green on the
  // B200 (profiles/r02_checks.log); TEVO_ZGEMM_CTSIGN=0 selects the run-time-sign instantiation
  static const bool ctsign = [] {
    const char* e = getenv("TEVO_ZGEMM_CTSIGN");
    return e ? atoi(e) != 0 : true;
  }();
#define TEVO_ZG_ARGS st, M, N, K, alpha, A, lda, conjA, B, ldb, conjB, beta, C, ldc, splits, Kc, part_stride, batch, bsA, \
                     bsB, bsC, lower_only
  if (!ctsign) {
    launch_c<AK, BK_, -1, -1>(TEVO_ZG_ARGS);
  } else if (conjA) {
    if (conjB)
      launch_c<AK, BK_, 1, 1>(TEVO_ZG_ARGS);
    else
      launch_c<AK, BK_, 1, 0>(TEVO_ZG_ARGS);
  } else {
    if (conjB)
      launch_c<AK, BK_, 0, 1>(TEVO_ZG_ARGS);
    else
      launch_c<AK, BK_, 0, 0>(TEVO_ZG_ARGS);
  }
#undef TEVO_ZG_ARGS
}

__global__ void zscale_matrix_kernel(int M, int N, cplx beta, cplx* C, long ldc) {
  long total = (long)M * N;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    int i = (int)(e % M);
    long j = e / M;
    cplx* p = C + j * ldc + i;
    *p = (beta.x == 0.0 && beta.y == 0.0) ? mk(0, 0) : cmul(beta, *p);
  }
}

// C = alpha * sum_z part_z + beta * C
__global__ void splitk_reduce_kernel(int M, int N, int splits, const cplx* __restrict__ part, long part_stride,
                                     cplx alpha, cplx beta, cplx* __restrict__ C, long ldc) {
  const long total = (long)M * N;
  const bool has_beta = (beta.x != 0.0 || beta.y != 0.0);
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    cplx s = mk(0, 0);
    for (int z = 0; z < splits; ++z) s = cadd(s, part[(long)z * part_stride + e]);
    const int i = (int)(e % M);
    const long j = e / M;
    cplx v = cmul(alpha, s);
    cplx* dst = C + j * ldc + i;
    if (has_beta) v = cadd(v, cmul(beta, *dst));
    *dst = v;
  }
}

template <bool AK, bool BK_>
void launch_split(cudaStream_t st, int M, int N, int K, const cplx* A, long lda, int conjA, const cplx* B, long ldb,
                  int conjB, cplx* part, int splits, int Kc) {
  launch<AK, BK_>(st, M, N, K, mk(1, 0), A, lda, conjA, B, ldb, conjB, mk(0, 0), part, M, splits, Kc, (long)M * N);
}

}  // namespace

long g_zgemm_launches = 0;
std::atomic<long long> g_launches{0};
void zgemm_force_tile(int tile) { g_forced_tile.store(tile); }
void zgemm_set_3m(int on) { g_zgemm_3m.store(on); }
ZgemmPrecise::ZgemmPrecise() { ++tl_zgemm_precise; }
ZgemmPrecise::~ZgemmPrecise() { --tl_zgemm_precise; }

void zgemm(cudaStream_t st, Op ta, Op tb, int M, int N, int K, cplx alpha, const cplx* A, long lda, const cplx* B,
           long ldb, cplx beta, cplx* C, long ldc) {
  if (M <= 0 || N <= 0) return;
  ++g_zgemm_launches;
  if (K <= 0) {
    long total = (long)M * N;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    zscale_matrix_kernel<<<blocks, 256, 0, st>>>(M, N, beta, C, ldc);
    TEVO_LAUNCHED();
    return;
  }
  const bool ak = (ta != OP_N);  // op(A)(i,k) = A[k + i*lda]  -> k contiguous
  const bool bk = (tb == OP_N);  // op(B)(k,j) = B[k + j*ldb]  -> k contiguous
  const int ca = (ta == OP_C), cb = (tb == OP_C);
  if (ak && bk)
    launch<true, true>(st, M, N, K, alpha, A, lda, ca, B, ldb, cb, beta, C, ldc);
  else if (ak && !bk)
    launch<true, false>(st, M, N, K, alpha, A, lda, ca, B, ldb, cb, beta, C, ldc);
  else if (!ak && bk)
    launch<false, true>(st, M, N, K, alpha, A, lda, ca, B, ldb, cb, beta, C, ldc);
  else
    launch<false, false>(st, M, N, K, alpha, A, lda, ca, B, ldb, cb, beta, C, ldc);
}


// Hermitian-result variant (C = alpha op(A) op(B) + beta C is known to be Hermitian, M == N): only the CTA tiles that
// touch the lower triangle are computed; mirror != 0 then writes conj(lower) into the strict upper triangle.
__global__ void mirror_lower_kernel(int n, cplx* __restrict__ C, long ldc) {
  __shared__ cplx tile[32][33];
  const int bi = blockIdx.x, bj = blockIdx.y;  // tile (bi, bj) of the lower triangle, bi >= bj
  if (bi < bj) return;
  const int r0 = bi * 32, c0 = bj * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + threadIdx.x, c = c0 + j;
    if (r < n && c < n) tile[j][threadIdx.x] = C[(long)c * ldc + r];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    // element (c0 + threadIdx.x, r0 + j) of the upper triangle = conj(C(r0 + j, c0 + threadIdx.x))
    const int rr = c0 + threadIdx.x, cc = r0 + j;
    if (rr < n && cc < n && rr < cc) {
      const cplx v = tile[threadIdx.x][j];
      C[(long)cc * ldc + rr] = mk(v.x, -v.y);
    }
  }
}

void zgemm_herm(cudaStream_t st, Op ta, Op tb, int N, int K, cplx alpha, const cplx* A, long lda, const cplx* B, long ldb,
                cplx beta, cplx* C, long ldc, bool mirror) {
  if (N <= 0) return;
  if (K <= 0) {
    zgemm(st, ta, tb, N, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
    return;
  }
  ++g_zgemm_launches;
  const bool ak = (ta != OP_N), bk = (tb == OP_N);
  const int ca = (ta == OP_C), cb = (tb == OP_C);
  if (ak && bk)
    launch<true, true>(st, N, N, K, alpha, A, lda, ca, B, ldb, cb, beta, C, ldc, 1, 0, 0, 1, 0, 0, 0, 1);
  else if (ak && !bk)
    launch<true, false>(st, N, N, K, alpha, A, lda, ca, B, ldb, cb, beta, C, ldc, 1, 0, 0, 1, 0, 0, 0, 1);
  else if (!ak && bk)
    launch<false, true>(st, N, N, K, alpha, A, lda, ca, B, ldb, cb, beta, C, ldc, 1, 0, 0, 1, 0, 0, 0, 1);
  else
    launch<false, false>(st, N, N, K, alpha, A, lda, ca, B, ldb, cb, beta, C, ldc, 1, 0, 0, 1, 0, 0, 0, 1);
  if (mirror && N > 1) {
    dim3 grid((N + 31) / 32, (N + 31) / 32), block(32, 8);
    mirror_lower_kernel<<<grid, block, 0, st>>>(N, C, ldc);
    TEVO_LAUNCHED();
  }
}

// Strided-batched variant: `batch` independent products with operand strides bsA / bsB / bsC (elements).
void zgemm_batched(cudaStream_t st, Op ta, Op tb, int M, int N, int K, cplx alpha, const cplx* A, long lda, long bsA,
                   const cplx* B, long ldb, long bsB, cplx beta, cplx* C, long ldc, long bsC, int batch) {
  if (M <= 0 || N <= 0 || K <= 0 || batch <= 0) return;
  ++g_zgemm_launches;
  const bool ak = (ta != OP_N), bk = (tb == OP_N);
  const int ca = (ta == OP_C), cb = (tb == OP_C);
  if (ak && bk)
    launch<true, true>(st, M, N, K, alpha, A, lda, ca, B, ldb, cb, beta, C, ldc, 1, 0, 0, batch, bsA, bsB, bsC);
  else if (ak && !bk)
    launch<true, false>(st, M, N, K, alpha, A, lda, ca, B, ldb, cb, beta, C, ldc, 1, 0, 0, batch, bsA, bsB, bsC);
  else if (!ak && bk)
    launch<false, true>(st, M, N, K, alpha, A, lda, ca, B, ldb, cb, beta, C, ldc, 1, 0, 0, batch, bsA, bsB, bsC);
  else
    launch<false, false>(st, M, N, K, alpha, A, lda, ca, B, ldb, cb, beta, C, ldc, 1, 0, 0, batch, bsA, bsB, bsC);
}

// Split-K variant for skinny outputs (few CTA tiles, long K): deterministic two-pass (partials + reduce).
void zgemm_splitk(cudaStream_t st, Op ta, Op tb, int M, int N, int K, cplx alpha, const cplx* A, long lda, const cplx* B,
                  long ldb, cplx beta, cplx* C, long ldc, cplx** part, size_t* part_cap) {
  if (M <= 0 || N <= 0) return;
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  int splits = 148 / (tiles > 0 ? tiles : 1);
  const int maxs = (K + 4 * BK - 1) / (4 * BK);
  if (splits > maxs) splits = maxs;
  if (splits > 32) splits = 32;
  if (splits <= 1) {
    zgemm(st, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
    return;
  }
  int Kc = (K + splits - 1) / splits;
  Kc = ((Kc + BK - 1) / BK) * BK;
  splits = (K + Kc - 1) / Kc;
  const size_t need = (size_t)splits * M * N;
  if (*part_cap < need) {
    if (*part) TEVO_CUDA_CHECK(cudaFree(*part));
    *part = nullptr;
    *part_cap = 0;
    TEVO_CUDA_CHECK(cudaMalloc(part, need * sizeof(cplx)));
    *part_cap = need;
  }
  ++g_zgemm_launches;
  const bool ak = (ta != OP_N), bk = (tb == OP_N);
  const int ca = (ta == OP_C), cb = (tb == OP_C);
  if (ak && bk)
    launch_split<true, true>(st, M, N, K, A, lda, ca, B, ldb, cb, *part, splits, Kc);
  else if (ak && !bk)
    launch_split<true, false>(st, M, N, K, A, lda, ca, B, ldb, cb, *part, splits, Kc);
  else if (!ak && bk)
    launch_split<false, true>(st, M, N, K, A, lda, ca, B, ldb, cb, *part, splits, Kc);
  else
    launch_split<false, false>(st, M, N, K, A, lda, ca, B, ldb, cb, *part, splits, Kc);
  const long total = (long)M * N;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  splitk_reduce_kernel<<<blocks, 256, 0, st>>>(M, N, splits, *part, (long)M * N, alpha, beta, C, ldc);
  TEVO_LAUNCHED();
}

}  // namespace tevo
