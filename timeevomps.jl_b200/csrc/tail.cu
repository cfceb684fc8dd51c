// The last columns of the Householder tridiagonalisation inside ONE thread-block cluster (default since round 2 in
// the single-lane configuration, 16-CTA cluster; validated on the B200 at cluster sizes 1 / 8 / 16 and measured at
// -5 % / -14 % of the panel time at n = 2048 / 1024: profiles/r02_deferred_checks.log; TEVO_CLUSTER_TAIL=0 disables it).
//
// The cooperative panel kernel (heig.cu) pays ~8 us per column for two grid barriers and the dependent L2 round
// trips between them, whatever the size of the trailing matrix.  Once the trailing matrix fits the shared memory of
// one cluster (m <= 384 rows over 16 CTAs, m <= 96 in a single CTA) the whole rest of the factorisation can run
// without global synchronisation: cluster.sync() costs ~0.2 us and the vectors travel through distributed shared
// memory (proto/cluster_bar_bench.cu).  With the matrix on chip the rank-2 updates are applied eagerly, so there is
// no panel bookkeeping (V, W, the replicated dot accumulators) at all:
//
//   rows of the trailing matrix are dealt cyclically to the CTAs (row i -> CTA i % C, full rows, Hermitian storage);
//   per column c:  (A) every CTA forms the reflector from the gathered column x (identical arithmetic everywhere),
//                  (B) one fused pass over its rows: apply the pending update  A -= v_p w_p^H + w_p v_p^H  of the
//                      previous column, then y_i = A(i, c+1:) v ; y_i is stored into every CTA      -> cluster.sync
//                  (C) every CTA forms  w = tau y - (tau conj(tau) y^H v / 2) v  and the next column of the updated
//                      matrix for its rows, stored into every CTA                                     -> cluster.sync
//
// Output convention identical to the panel kernel: d, e (= beta), taus, reflectors in the columns of A (unit element
// stored, zeros above), so the compact-WY factors follow from V^H V computed by a GEMM afterwards (heig.cu).
// The data flow was checked against LAPACK on the CPU (Q^H A Q tridiagonal to 1e-14) before it was written down here.
#include <cooperative_groups.h>

#include "common.cuh"
#include "heig.h"

namespace cg = cooperative_groups;

namespace tevo {
namespace {

constexpr int TNT = 256;

struct TailArgs {
  int n, t0;
  cplx* A;
  long lda;
  double* d;
  double* e;
  cplx* taus;
};

template <int C>
__device__ inline cplx* peer_ptr(cplx* p, unsigned peer) {
  if constexpr (C > 1) {
    return cg::this_cluster().map_shared_rank(p, peer);
  } else {
    return p;
  }
}
template <int C>
__device__ inline void cluster_barrier() {
  if constexpr (C > 1) {
    cg::this_cluster().sync();
  } else {
    __syncthreads();
  }
}

template <int C>
__global__ void __launch_bounds__(TNT, 1) tridiag_tail_kernel(TailArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int n = a.n, t0 = a.t0, m = n - t0;
  const int ldr = (m + 1) & ~1;            // row length in shared memory
  const int slots = (m + C - 1) / C;       // rows per CTA
  cplx* Ar = reinterpret_cast<cplx*>(smraw);  // slots x ldr : Ar[s][j] = A(s*C + rank, j)
  cplx* xs = Ar + (size_t)slots * ldr;     // gathered column c (written by all CTAs)
  cplx* ys = xs + ldr;                     // gathered y = A v
  cplx* vb = ys + ldr;                     // 2 x ldr : v of the current / previous column
  cplx* wb = vb + 2 * ldr;                 // 2 x ldr : w of the current / previous column
  __shared__ double sred[3 * 32];
  unsigned rank = 0;
  if constexpr (C > 1) rank = cg::this_cluster().block_rank();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- prologue: own rows (row i = conj of column i: the matrix is stored fully), column 0, no pending update
  for (int s = 0; s < slots; ++s) {
    const int i = s * C + (int)rank;
    if (i >= m) break;
    const cplx* col = a.A + (long)(t0 + i) * a.lda + t0;
    for (int j = tid; j < m; j += TNT) Ar[(size_t)s * ldr + j] = cconj(col[j]);
  }
  for (int j = tid; j < m; j += TNT) {
    xs[j] = a.A[(long)t0 * a.lda + t0 + j];
    vb[j] = vb[ldr + j] = mk(0, 0);
    wb[j] = wb[ldr + j] = mk(0, 0);
  }
  cluster_barrier<C>();  // every CTA has read its part of the global matrix before reflectors overwrite its columns

  for (int c = 0; c + 1 < m; ++c) {
    cplx* vs = vb + (c & 1) * ldr;
    cplx* ws = wb + (c & 1) * ldr;
    const cplx* vp = vb + ((c + 1) & 1) * ldr;
    const cplx* wp = wb + ((c + 1) & 1) * ldr;
    // ---- (A) reflector from the gathered column (same arithmetic in every CTA)
    double part[1] = {0.0};
    for (int i = c + 2 + tid; i < m; i += TNT) part[0] += cabs2(xs[i]);
    block_sum<1>(part, sred);
    const double xn2 = part[0];
    const cplx alpha = xs[c + 1];
    const double s2 = cabs2(alpha) + xn2;
    double beta;
    cplx tau, scal;
    if ((xn2 == 0.0 && alpha.y == 0.0) || !(s2 > 1e-290)) {
      beta = alpha.x;
      tau = mk(0, 0);
      scal = mk(0, 0);
    } else {
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
    if (rank == 0 && tid == 0) {
      a.d[t0 + c] = xs[c].x;
      a.e[t0 + c] = beta;
      a.taus[t0 + c] = tau;
    }
    for (int i = tid; i < m; i += TNT) vs[i] = (i <= c) ? mk(0, 0) : (i == c + 1 ? mk(1, 0) : cmul(xs[i], scal));
    __syncthreads();
    // the reflector goes to column t0+c of A (zeros on and above the diagonal row), spread over the cluster
    for (int i = (int)rank * TNT + tid; i < m; i += C * TNT) a.A[(long)(t0 + c) * a.lda + t0 + i] = vs[i];
    // ---- (B) fused pass over the own rows: a warp works on up to four of its rows at a time so that v, v_p, w_p of a
    // column are read from shared memory once per four matrix elements
    for (int q0 = 0; warp + (TNT / 32) * q0 < slots; q0 += 4) {
      cplx* row[4];
      bool on[4];
      int ri[4];
      cplx vpi[4], wpi[4], acc[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int s = warp + (TNT / 32) * (q0 + q);
        const int i = s * C + (int)rank;
        on[q] = (s < slots) && (i < m) && (i > c);  // warp-uniform
        ri[q] = i;
        row[q] = Ar + (size_t)(on[q] ? s : 0) * ldr;
        vpi[q] = on[q] ? vp[i] : mk(0, 0);
        wpi[q] = on[q] ? wp[i] : mk(0, 0);
        acc[q] = mk(0, 0);
      }
      for (int j = c + 1 + lane; j < m; j += 32) {
        const cplx wj = wp[j], vj = vp[j], vsj = vs[j];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (!on[q]) continue;
          cplx aij = row[q][j];
          // A(i,j) -= v_p(i) conj(w_p(j)) + w_p(i) conj(v_p(j))
          aij.x -= vpi[q].x * wj.x + vpi[q].y * wj.y + wpi[q].x * vj.x + wpi[q].y * vj.y;
          aij.y -= vpi[q].y * wj.x - vpi[q].x * wj.y + wpi[q].y * vj.x - wpi[q].x * vj.y;
          row[q][j] = aij;
          cfma(acc[q], aij, vsj);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (!on[q]) continue;
        acc[q].x = warp_sum(acc[q].x);
        acc[q].y = warp_sum(acc[q].y);
        if (lane < C) peer_ptr<C>(ys, lane)[ri[q]] = acc[q];
      }
    }
    cluster_barrier<C>();
    // ---- (C) w and the next column
    double p2[2] = {0.0, 0.0};
    for (int i = c + 1 + tid; i < m; i += TNT) {  // y^H v
      const cplx y = ys[i], v = vs[i];
      p2[0] += y.x * v.x + y.y * v.y;
      p2[1] += y.x * v.y - y.y * v.x;
    }
    block_sum<2>(p2, sred);
    const cplx yv = mk(p2[0], p2[1]);
    // w = tau y + alpha_w v,  alpha_w = -1/2 tau (conj(tau) y^H v)
    const cplx aw = cscale(-0.5, cmul(tau, cmul(cconj(tau), yv)));
    for (int i = tid; i < m; i += TNT)
      ws[i] = (i <= c) ? mk(0, 0) : cadd(cmul(tau, ys[i]), cmul(aw, vs[i]));
    __syncthreads();
    // column c+1 of the matrix after the update (v, w), own rows >= c+1; v(c+1) = 1
    {
      const cplx wc1 = ws[c + 1];
      for (int s = tid; s < slots; s += TNT) {
        const int i = s * C + (int)rank;
        if (i >= m || i < c + 1) continue;
        cplx x = Ar[(size_t)s * ldr + c + 1];
        const cplx vi = vs[i], wi = ws[i];
        // x -= v(i) conj(w(c+1)) + w(i) conj(v(c+1))
        x.x -= vi.x * wc1.x + vi.y * wc1.y + wi.x;
        x.y -= vi.y * wc1.x - vi.x * wc1.y + wi.y;
#pragma unroll
        for (int p = 0; p < C; ++p) peer_ptr<C>(xs, p)[i] = x;
      }
    }
    cluster_barrier<C>();
  }
  // last diagonal element back to the global matrix (last_diag_kernel reads it), explicit zeros above the reflectors
  if (rank == 0 && tid == 0) a.A[(long)(n - 1) * a.lda + n - 1] = mk(xs[m - 1].x, 0.0);
  for (long e = (long)rank * TNT + tid; e < (long)t0 * (m - 1); e += (long)C * TNT) {
    const int r = (int)(e % t0);
    const int k = (int)(e / t0);
    a.A[(long)(t0 + k) * a.lda + r] = mk(0, 0);
  }
}

template <int C>
void launch_tail(cudaStream_t st, const TailArgs& ta, size_t smem) {
  TEVO_CUDA_CHECK(cudaFuncSetAttribute(tridiag_tail_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (C > 8)
    TEVO_CUDA_CHECK(cudaFuncSetAttribute(tridiag_tail_kernel<C>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C);
  cfg.blockDim = dim3(TNT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (C > 1) ? 1 : 0;
  TEVO_CUDA_CHECK(cudaLaunchKernelEx(&cfg, tridiag_tail_kernel<C>, ta));
  TEVO_LAUNCHED();
}

size_t tail_smem(int cluster, int m) {
  const size_t ldr = (size_t)((m + 1) & ~1);
  const size_t slots = (size_t)((m + cluster - 1) / cluster);
  return (slots * ldr + 6 * ldr) * sizeof(cplx);
}

}  // namespace

// largest trailing dimension the tail kernel takes with this cluster size (shared-memory budget 200 KB per CTA)
int tridiag_tail_max_m(int cluster) {
  int m = 0;
  for (int cand = 64; cand <= 1024; cand += 64)
    if (tail_smem(cluster, cand) <= 200 * 1024) m = cand;
  return m;
}

void tridiag_tail(cudaStream_t st, int cluster, int n, int t0, cplx* A, long lda, double* d, double* e, cplx* taus) {
  TailArgs ta;
  ta.n = n; ta.t0 = t0; ta.A = A; ta.lda = lda; ta.d = d; ta.e = e; ta.taus = taus;
  const size_t smem = tail_smem(cluster, n - t0);
  switch (cluster) {
    case 1: launch_tail<1>(st, ta, smem); break;
    case 2: launch_tail<2>(st, ta, smem); break;
    case 4: launch_tail<4>(st, ta, smem); break;
    case 8: launch_tail<8>(st, ta, smem); break;
    case 16: launch_tail<16>(st, ta, smem); break;
    default: throw TevoError(1, "tridiag_tail: cluster size must be 1, 2, 4, 8 or 16");
  }
}

}  // namespace tevo
