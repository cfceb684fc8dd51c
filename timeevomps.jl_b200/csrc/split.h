// Truncated split of a two-site wavefunction: theta (m x n) -> L (m x k) * R (k x n), the arithmetic of
// ITensors.replacebond!/factorize/svd/truncate! as called from src/bondgate.jl:51 and src/tdvp.jl:64.
#pragma once
#include <functional>
#include <vector>

#include "common.cuh"
#include "cholqr.h"
#include "heig.h"
#include "kernels.h"

namespace tevo {

struct DevBuf {
  cplx* p = nullptr;
  size_t cap = 0;  // elements
};

struct SplitResult {
  DevBuf L, R;
  int k = 0;
  TruncInfo info;
};

class Split {
 public:
  void init(cudaStream_t st);
  void destroy();
  // left:  L = U (isometry),      R = S V^+ (centre; divided by its norm when normalize)
  // right: L = U S (centre),      R = V^+ (isometry)
  // eigs (host, optional): kept sigma^2, descending.  alloc: allocator for the two outputs.
  SplitResult run(cudaStream_t st, const cplx* theta, int m, int n, bool left, const TruncParams& tp, bool normalize,
                  double* eigs, int eigs_cap, const std::function<DevBuf(size_t)>& alloc);
  // orthogonal factorisation used by the centre moves of orthogonalize! (no truncation, no normalisation):
  //   left : A (m x n) = Q (m x k) * C (k x n),  Q^H Q = 1      right: A = C (m x k) * Q (k x n),  Q Q^H = 1
  SplitResult ortho_factor(cudaStream_t st, const cplx* A, int m, int n, bool left,
                           const std::function<DevBuf(size_t)>& alloc);
  // buffers replaced by a stage-2 re-gather; the owner returns them to its pool with take_released()
  std::vector<DevBuf> take_released() {
    std::vector<DevBuf> r;
    r.swap(pending_release_);
    return r;
  }
  CholQR& cholqr() { return cholqr_; }
  Heig& heig() { return heig_; }
  int* perm_dev() { return d_perm_; }
  double* sorted_dev() { return d_sorted_; }
  void reserve_vec(size_t n) { ensure_vec(n); }
  TruncInfo* info_dev() { return d_info_; }
  long cholqr_used = 0, cholqr_fallback = 0;
  int last_sweeps = 0;
  bool use_jacobi = false;  // TEVO_SPLIT_JACOBI=1: first-generation Hestenes path (slow, kept as a cross-check)

 private:
  SplitResult run_jacobi(cudaStream_t st, const cplx* theta, int m, int n, bool left, const TruncParams& tp,
                         bool normalize, double* eigs, int eigs_cap, const std::function<DevBuf(size_t)>& alloc);
  Heig heig_;
  CholQR cholqr_;
  bool use_cholqr_ = true;  // TEVO_NO_CHOLQR=1 disables the fast centre-move path
  std::vector<DevBuf> pending_release_;
  double* d_scal_ = nullptr;  // [0..1] |theta|^2, [2..3] |residual|^2
  cplx* ws(int slot, size_t nelem);
  DevBuf ws_[4];
  double* d_lam_ = nullptr;
  double* d_sorted_ = nullptr;
  int* d_perm_ = nullptr;
  size_t vcap_ = 0;
  TruncInfo* d_info_ = nullptr;
  int* d_rot_ = nullptr;
  // pinned
  TruncInfo* h_info_ = nullptr;
  int* h_rot_ = nullptr;
  double* h_sorted_ = nullptr;
  void ensure_vec(size_t n);
};

}  // namespace tevo
