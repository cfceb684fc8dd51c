// First-generation truncated split: one-sided (Hestenes) Jacobi SVD with accumulated rotations.
//   left : X = theta^+ (n x m); rotate columns of X until orthogonal, accumulate in V (m x m, starts as I):
//          theta^+ V = R^+  =>  U = V,  S V^+ = X^+   -> no Gram matrix, no division by singular values,
//          U is unitary to machine precision whatever the rank of theta.
//   right: X = theta (m x n), V (n x n):  theta V = U S  =>  L = X, R = V^+.
// Truncation (upstream truncate!) and normalisation are done on the device (kernels.cu).
#include "split.h"

#include <cmath>
#include <cstdlib>
#include <cstring>

#include "prof.h"
#include "zgemm.h"

namespace tevo {

void Split::init(cudaStream_t) {
  heig_.init();
  cholqr_.init();
  {
    const char* e2 = getenv("TEVO_NO_CHOLQR");
    use_cholqr_ = !(e2 && e2[0] == '1');
  }
  TEVO_CUDA_CHECK(cudaMalloc(&d_scal_, sizeof(double) * 8));
  const char* env = getenv("TEVO_SPLIT_JACOBI");
  use_jacobi = (env && env[0] == '1');
  TEVO_CUDA_CHECK(cudaMalloc(&d_info_, sizeof(TruncInfo)));
  TEVO_CUDA_CHECK(cudaMalloc(&d_rot_, sizeof(int)));
  TEVO_CUDA_CHECK(cudaMallocHost(&h_info_, sizeof(TruncInfo)));
  TEVO_CUDA_CHECK(cudaMallocHost(&h_rot_, sizeof(int)));
}

void Split::destroy() {
  heig_.destroy();
  cholqr_.destroy();
  if (d_scal_) cudaFree(d_scal_);
  for (auto& b : ws_)
    if (b.p) cudaFree(b.p);
  if (d_lam_) cudaFree(d_lam_);
  if (d_sorted_) cudaFree(d_sorted_);
  if (d_perm_) cudaFree(d_perm_);
  if (d_info_) cudaFree(d_info_);
  if (d_rot_) cudaFree(d_rot_);
  if (h_info_) cudaFreeHost(h_info_);
  if (h_rot_) cudaFreeHost(h_rot_);
  if (h_sorted_) cudaFreeHost(h_sorted_);
}

cplx* Split::ws(int slot, size_t nelem) {
  DevBuf& b = ws_[slot];
  if (b.cap < nelem) {
    if (b.p) TEVO_CUDA_CHECK(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = nelem + nelem / 8 + 64;
    TEVO_CUDA_CHECK(cudaMalloc(&b.p, want * sizeof(cplx)));
    b.cap = want;
  }
  return b.p;
}

void Split::ensure_vec(size_t n) {
  if (vcap_ >= n) return;
  if (d_lam_) cudaFree(d_lam_);
  if (d_sorted_) cudaFree(d_sorted_);
  if (d_perm_) cudaFree(d_perm_);
  if (h_sorted_) cudaFreeHost(h_sorted_);
  size_t want = n * 2 + 64;
  TEVO_CUDA_CHECK(cudaMalloc(&d_lam_, want * sizeof(double)));
  TEVO_CUDA_CHECK(cudaMalloc(&d_sorted_, want * sizeof(double)));
  TEVO_CUDA_CHECK(cudaMalloc(&d_perm_, want * sizeof(int)));
  TEVO_CUDA_CHECK(cudaMallocHost(&h_sorted_, want * sizeof(double)));
  vcap_ = want;
}

SplitResult Split::run_jacobi(cudaStream_t st, const cplx* theta, int m, int n, bool left, const TruncParams& tp,
                              bool normalize, double* eigs, int eigs_cap,
                              const std::function<DevBuf(size_t)>& alloc) {
  if (m < 1 || n < 1) throw TevoError(1, "split: empty matrix");
  const int nrx = left ? n : m;  // rows of X
  const int nc = left ? m : n;   // columns of X = size of V
  cplx* X = ws(0, (size_t)nrx * nc);
  cplx* V = ws(1, (size_t)nc * nc);
  ensure_vec((size_t)nc);
  if (left)
    conj_transpose(st, m, n, theta, m, X, n, true);
  else
    TEVO_CUDA_CHECK(cudaMemcpyAsync(X, theta, (size_t)m * n * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
  set_identity(st, nc, V, nc);

  const double tol = std::max(4.0 * 2.220446049250313e-16 * std::sqrt((double)nrx), 1e-15);
  const int nceven = (nc + 1) & ~1;
  last_sweeps = 0;
  if (nc > 1) {
    for (int sweep = 0; sweep < 60; ++sweep) {
      TEVO_CUDA_CHECK(cudaMemsetAsync(d_rot_, 0, sizeof(int), st));
      for (int r = 0; r < nceven - 1; ++r) jacobi_round(st, X, nrx, nrx, V, nc, nc, nc, r, tol, d_rot_);
      TEVO_CUDA_CHECK(cudaMemcpyAsync(h_rot_, d_rot_, sizeof(int), cudaMemcpyDeviceToHost, st));
      TEVO_CUDA_CHECK(cudaStreamSynchronize(st));
      ++last_sweeps;
      if (*h_rot_ == 0) break;
    }
  }
  colnorm2(st, X, nrx, nrx, nc, d_lam_);
  const int nfull = m < n ? m : n;
  sort_truncate(st, d_lam_, nc, nfull, tp, d_sorted_, d_perm_, d_info_);
  TEVO_CUDA_CHECK(cudaMemcpyAsync(h_info_, d_info_, sizeof(TruncInfo), cudaMemcpyDeviceToHost, st));
  if (eigs) TEVO_CUDA_CHECK(cudaMemcpyAsync(h_sorted_, d_sorted_, sizeof(double) * nfull, cudaMemcpyDeviceToHost, st));
  TEVO_CUDA_CHECK(cudaStreamSynchronize(st));
  SplitResult res;
  res.info = *h_info_;
  if (res.info.bad) throw TevoError(5, "split: non-finite singular values (NaN/Inf in the wavefunction)");
  const int k = res.info.nkeep;
  res.k = k;
  if (eigs) {
    if (eigs_cap < k) throw TevoError(1, "split: eigs buffer too small for the kept spectrum");
    memcpy(eigs, h_sorted_, sizeof(double) * k);
  }
  res.L = alloc((size_t)m * k);
  res.R = alloc((size_t)k * n);
  const double* nrm = normalize ? &d_info_->sum_kept : nullptr;
  if (left) {
    gather_cols(st, m, k, V, nc, d_perm_, res.L.p, m, nullptr);
    gather_cols_conjT(st, n, k, X, nrx, d_perm_, res.R.p, k, nrm);
  } else {
    gather_cols(st, m, k, X, nrx, d_perm_, res.L.p, m, nrm);
    gather_cols_conjT(st, n, k, V, nc, d_perm_, res.R.p, k, nullptr);
  }
  return res;
}


// ---------------------------------------------------------------------------------------------------------
// Second-generation split: density-matrix route.
//   left : rho = theta theta^+ (m x m) -> leading eigenvectors U (unitary to machine precision, no division by
//          singular values), R = U^+ theta; kept spectrum refined as |R_i|^2 (Rayleigh quotients, second-order accurate)
//   right: rho = theta^+ theta (n x n) -> V, L = theta V, R = V^+
// Discarded weight = |theta - L R|_F^2, i.e. the error of the state actually kept.
// ---------------------------------------------------------------------------------------------------------
namespace {
// rows of R (k x n, ld): out[i] = sum_j |R(i,j)|^2
__global__ void rownorm2_kernel(int k, int n, const cplx* __restrict__ R, long ld, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  double s = 0.0;
  for (int j = 0; j < n; ++j) s += cabs2(R[(long)j * ld + i]);
  out[i] = s;
}

// column scaling used by ortho_factor: Q(:, i) = B(:, i) / s_i, s_i = sqrt(norm2[i]) (zero column when s_i == 0)
__global__ void col_normalize_kernel(int nr, int nc, cplx* __restrict__ B, long ld, const double* __restrict__ norm2) {
  const long total = (long)nr * nc;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int r = (int)(e % nr);
    const long c = e / nr;
    const double s2 = norm2[c];
    const double inv = s2 > 0.0 ? rsqrt(s2) : 0.0;
    cplx* p = B + c * ld + r;
    *p = cscale(inv, *p);
  }
}
// C(i, :) = s_i * conj(V(:, i))^T   (k x n from V n x k), or row-normalise in place
__global__ void scaled_conjT_kernel(int n, int k, const cplx* __restrict__ V, long ldv, const double* __restrict__ norm2,
                                    cplx* __restrict__ C, long ldc) {
  const long total = (long)n * k;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int i = (int)(e % k);
    const long r = e / k;
    const double s = sqrt(fmax(norm2[i], 0.0));
    const cplx v = V[(long)i * ldv + r];
    C[r * ldc + i] = mk(s * v.x, -s * v.y);
  }
}
__global__ void row_normalize_kernel(int nr, int nc, cplx* __restrict__ B, long ld, const double* __restrict__ norm2) {
  const long total = (long)nr * nc;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int r = (int)(e % nr);
    const long c = e / nr;
    const double s2 = norm2[r];
    const double inv = s2 > 0.0 ? rsqrt(s2) : 0.0;
    cplx* p = B + c * ld + r;
    *p = cscale(inv, *p);
  }
}
// C(:, i) = s_i * U(:, i)
__global__ void col_scale_sqrt_kernel(int nr, int nc, const cplx* __restrict__ U, long ldu, const double* __restrict__ norm2,
                                      cplx* __restrict__ C, long ldc) {
  const long total = (long)nr * nc;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int r = (int)(e % nr);
    const long c = e / nr;
    const double s = sqrt(fmax(norm2[c], 0.0));
    C[c * ldc + r] = cscale(s, U[c * ldu + r]);
  }
}
inline int gridfor(long total) {
  long b = (total + 255) / 256;
  if (b < 1) b = 1;
  if (b > 148 * 16) b = 148 * 16;
  return (int)b;
}
}  // namespace

SplitResult Split::run(cudaStream_t st, const cplx* theta, int m, int n, bool left, const TruncParams& tp,
                       bool normalize, double* eigs, int eigs_cap, const std::function<DevBuf(size_t)>& alloc) {
  if (use_jacobi) return run_jacobi(st, theta, m, n, left, tp, normalize, eigs, eigs_cap, alloc);
  if (m < 1 || n < 1) throw TevoError(1, "split: empty matrix");
  const cplx ONE = mk(1, 0), ZERO = mk(0, 0), MONE = mk(-1, 0);
  const int N = left ? m : n;
  const int nfull = m < n ? m : n;
  cplx* rho = ws(0, (size_t)N * N);
  ensure_vec((size_t)N);
  zdotc(st, (long)m * n, theta, theta, d_scal_);  // |theta|_F^2
  {
    ProfScope ps(st, P_RHO);
    if (left)
      zgemm_herm(st, OP_N, OP_C, m, n, ONE, theta, m, theta, m, ZERO, rho, m, true);
    else
      zgemm_herm(st, OP_C, OP_N, n, m, ONE, theta, m, theta, m, ZERO, rho, n, true);
  }
  heig_.factor(st, N, rho, N);
  // stage 1: candidate count k1 from the Gram spectrum, with a safety margin of the size of its rounding noise so
  // that nothing the accurate (stage-2) weights would keep is dropped here
  TruncParams tp1 = tp;
  tp1.inflate_rel = 8.0 * 2.220446049250313e-16 * (double)N;
  {
    ProfScope ps(st, P_TRUNC);
    sort_truncate(st, heig_.evals(), N, nfull, tp1, d_sorted_, d_perm_, d_info_);
    TEVO_CUDA_CHECK(cudaMemcpyAsync(h_info_, d_info_, sizeof(TruncInfo), cudaMemcpyDeviceToHost, st));
  }
  TEVO_CUDA_CHECK(cudaStreamSynchronize(st));
  SplitResult res;
  res.info = *h_info_;
  if (res.info.bad) throw TevoError(5, "split: non-finite spectrum (NaN/Inf in the wavefunction)");
  const int k1 = res.info.nkeep;
  DevBuf L1 = alloc((size_t)m * k1), R1 = alloc((size_t)k1 * n);
  double* lam2 = d_lam_;
  // The rows of R = U^+ theta (columns of L = theta V) carry the refined Schmidt weights |R_i|^2 down to ~1e-9 of the
  // largest one: this product keeps four real multiplications per complex one (ZgemmPrecise), every other GEMM of the
  // path may use three (zgemm.cu)
  if (left) {
    heig_.vectors(st, d_perm_, k1, L1.p, m);
    ProfScope ps(st, P_SPLIT_GEMM);
    ZgemmPrecise precise_rows;
    zgemm(st, OP_C, OP_N, k1, n, m, ONE, L1.p, m, theta, m, ZERO, R1.p, k1);
    rownorm2_kernel<<<(k1 + 127) / 128, 128, 0, st>>>(k1, n, R1.p, k1, lam2);
    TEVO_LAUNCHED();
  } else {
    cplx* V = ws(1, (size_t)n * k1);
    heig_.vectors(st, d_perm_, k1, V, n);
    ProfScope ps(st, P_SPLIT_GEMM);
    ZgemmPrecise precise_cols;
    zgemm(st, OP_N, OP_N, m, k1, n, ONE, theta, m, V, n, ZERO, L1.p, m);
    conj_transpose(st, n, k1, V, n, R1.p, k1, true);
    colnorm2(st, L1.p, m, m, k1, lam2);
  }
  // discarded weight: when it is a sizeable fraction of the total (stage-1 estimate >= 1e-5 |theta|^2) the difference
  // |theta|^2 - sum(kept) is accurate to ~1e-11 relative; only small discarded weights need the residual GEMM
  const bool by_difference = (k1 < nfull) && (res.info.sum_all > 0.0) &&
                             ((res.info.sum_all - res.info.sum_kept) >= 1e-5 * res.info.sum_all);
  if (k1 < nfull && !by_difference) {  // accurately known discarded weight = error of the kept state
    ProfScope ps(st, P_RESID);
    cplx* resid = ws(2, (size_t)m * n);
    TEVO_CUDA_CHECK(cudaMemcpyAsync(resid, theta, (size_t)m * n * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
    zgemm(st, OP_N, OP_N, m, n, k1, MONE, L1.p, m, R1.p, k1, ONE, resid, m);
    zdotc(st, (long)m * n, resid, resid, d_scal_ + 2);
  }
  // stage 2: upstream's truncate! loop continued on the refined (Rayleigh-quotient) weights
  refine_truncate(st, lam2, k1, nfull, d_scal_, d_scal_ + 2, by_difference, tp, d_sorted_, d_perm_, d_info_);
  TEVO_CUDA_CHECK(cudaMemcpyAsync(h_info_, d_info_, sizeof(TruncInfo), cudaMemcpyDeviceToHost, st));
  if (eigs) TEVO_CUDA_CHECK(cudaMemcpyAsync(h_sorted_, d_sorted_, sizeof(double) * k1, cudaMemcpyDeviceToHost, st));
  TEVO_CUDA_CHECK(cudaStreamSynchronize(st));
  res.info = *h_info_;
  if (res.info.bad) throw TevoError(5, "split: non-finite spectrum (NaN/Inf in the wavefunction)");
  const int k = res.info.nkeep;
  res.k = k;
  const double* nrm = normalize ? &d_info_->sum_kept : nullptr;
  if (k == k1 && res.info.identity) {
    res.L = L1;
    res.R = R1;
    if (normalize) {
      if (left)
        zscal_rsqrt(st, (long)k * n, nrm, res.R.p);
      else
        zscal_rsqrt(st, (long)m * k, nrm, res.L.p);
    }
  } else {
    res.L = alloc((size_t)m * k);
    res.R = alloc((size_t)k * n);
    gather_cols(st, m, k, L1.p, m, d_perm_, res.L.p, m, left ? nullptr : nrm);
    gather_rows(st, k, n, R1.p, k1, d_perm_, res.R.p, k, left ? nrm : nullptr);
    pending_release_.push_back(L1);
    pending_release_.push_back(R1);
  }
  if (eigs) {
    if (eigs_cap < k) throw TevoError(1, "split: eigs buffer too small for the kept spectrum");
    memcpy(eigs, h_sorted_, sizeof(double) * k);
  }
  return res;
}

SplitResult Split::ortho_factor(cudaStream_t st, const cplx* A, int m, int n, bool left,
                                const std::function<DevBuf(size_t)>& alloc) {
  TruncParams tp;
  memset(&tp, 0, sizeof(tp));
  tp.mindim = 1;
  tp.do_rel_cutoff = 1;
  // the Gram matrix must be the small one; otherwise (or on the Jacobi path) use the general split
  if (use_jacobi || (left && m < n) || (!left && m > n)) return run(st, A, m, n, left, tp, false, nullptr, 0, alloc);
  const cplx ONE = mk(1, 0), ZERO = mk(0, 0);
  const int k = left ? n : m;
  if (use_cholqr_ && k >= 2) {
    // fast path: CholeskyQR2 (well-conditioned centre tensors); the right case works on A^H
    SplitResult res;
    res.k = k;
    memset(&res.info, 0, sizeof(res.info));
    res.info.nkeep = res.info.nfull = k;
    res.L = alloc((size_t)m * k);
    res.R = alloc((size_t)k * n);
    bool ok;
    if (left) {
      ok = cholqr_.factor(st, A, m, n, res.L.p, res.R.p);
    } else {
      cplx* At = ws(0, (size_t)n * m);
      cplx* Qt = ws(1, (size_t)n * m);
      cplx* Ct = ws(2, (size_t)m * m);
      conj_transpose(st, m, n, A, m, At, n, true);
      ok = cholqr_.factor(st, At, n, m, Qt, Ct);
      if (ok) {
        conj_transpose(st, n, m, Qt, n, res.R.p, m, true);  // Q (m x n) = Qt^H
        conj_transpose(st, m, m, Ct, m, res.L.p, m, true);  // C (m x m) = Ct^H
      }
    }
    if (ok) {
      ++cholqr_used;
      return res;
    }
    ++cholqr_fallback;
    pending_release_.push_back(res.L);
    pending_release_.push_back(res.R);
  }
  cplx* G = ws(0, (size_t)k * k);
  ensure_vec((size_t)k);
  {
    ProfScope ps(st, P_RHO);
    if (left)
      zgemm_herm(st, OP_C, OP_N, n, m, ONE, A, m, A, m, ZERO, G, n, true);
    else
      zgemm_herm(st, OP_N, OP_C, m, n, ONE, A, m, A, m, ZERO, G, m, true);
  }
  heig_.factor(st, k, G, k);
  sort_truncate(st, heig_.evals(), k, k, tp, d_sorted_, d_perm_, d_info_);
  cplx* V = ws(1, (size_t)k * k);
  heig_.vectors(st, d_perm_, k, V, k);
  SplitResult res;
  res.k = k;
  memset(&res.info, 0, sizeof(res.info));
  res.info.nkeep = res.info.nfull = k;
  res.L = alloc((size_t)m * k);
  res.R = alloc((size_t)k * n);
  double* s2 = d_lam_;
  ProfScope psh(st, P_HOP_GEMM);
  if (left) {
    // B = A V ; s_i = |B_i| ; Q = B / s ; C = diag(s) V^+
    zgemm(st, OP_N, OP_N, m, k, n, ONE, A, m, V, n, ZERO, res.L.p, m);
    colnorm2(st, res.L.p, m, m, k, s2);
    col_normalize_kernel<<<gridfor((long)m * k), 256, 0, st>>>(m, k, res.L.p, m, s2);
    TEVO_LAUNCHED();
    scaled_conjT_kernel<<<gridfor((long)n * k), 256, 0, st>>>(n, k, V, k, s2, res.R.p, k);
    TEVO_LAUNCHED();
  } else {
    // B = U^+ A ; s_i = |B_i| ; Q = B / s (rows) ; C = U diag(s)
    zgemm(st, OP_C, OP_N, k, n, m, ONE, V, m, A, m, ZERO, res.R.p, k);
    rownorm2_kernel<<<(k + 127) / 128, 128, 0, st>>>(k, n, res.R.p, k, s2);
    TEVO_LAUNCHED();
    row_normalize_kernel<<<gridfor((long)k * n), 256, 0, st>>>(k, n, res.R.p, k, s2);
    TEVO_LAUNCHED();
    col_scale_sqrt_kernel<<<gridfor((long)m * k), 256, 0, st>>>(m, k, V, m, s2, res.L.p, m);
    TEVO_LAUNCHED();
  }
  return res;
}

}  // namespace tevo
