// First-generation truncated split: one-sided (Hestenes) Jacobi SVD with accumulated rotations.
//   left : X = theta^+ (n x m); rotate columns of X until orthogonal, accumulate in V (m x m, starts as I):
//          theta^+ V = R^+  =>  U = V,  S V^+ = X^+   -> no Gram matrix, no division by singular values,
//          U is unitary to machine precision whatever the rank of theta.
//   right: X = theta (m x n), V (n x n):  theta V = U S  =>  L = X, R = V^+.
// Truncation (upstream truncate!) and normalisation are done on the device (kernels.cu).
#include "split.h"

#include <cmath>
#include <cstring>

namespace tevo {

void Split::init(cudaStream_t) {
  TEVO_CUDA_CHECK(cudaMalloc(&d_info_, sizeof(TruncInfo)));
  TEVO_CUDA_CHECK(cudaMalloc(&d_rot_, sizeof(int)));
  TEVO_CUDA_CHECK(cudaMallocHost(&h_info_, sizeof(TruncInfo)));
  TEVO_CUDA_CHECK(cudaMallocHost(&h_rot_, sizeof(int)));
}

void Split::destroy() {
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

SplitResult Split::run(cudaStream_t st, const cplx* theta, int m, int n, bool left, const TruncParams& tp,
                       bool normalize, double* eigs, int eigs_cap, const std::function<DevBuf(size_t)>& alloc) {
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

}  // namespace tevo
