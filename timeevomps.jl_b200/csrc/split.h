// Truncated split of a two-site wavefunction: theta (m x n) -> L (m x k) * R (k x n), the arithmetic of
// ITensors.replacebond!/factorize/svd/truncate! as called from src/bondgate.jl:51 and src/tdvp.jl:64.
#pragma once
#include <functional>

#include "common.cuh"
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
  int last_sweeps = 0;

 private:
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
