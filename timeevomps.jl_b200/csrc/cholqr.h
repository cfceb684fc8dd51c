// CholeskyQR2 (see cholqr.cu)
#pragma once
#include "common.cuh"

namespace tevo {

class CholQR {
 public:
  void init();
  void destroy();
  // A (m x n, ld m, m >= n) = Q (m x n, ld m) * C (n x n, ld n).  Returns false (outputs undefined) when A is too
  // ill-conditioned for the method; synchronises the stream twice (pivot checks).
  bool factor(cudaStream_t st, const cplx* A, int m, int n, cplx* Q, cplx* C);
  // one diagonal-block step on device buffers (tests / timing): variant 0 = panel kernel (default), 1 = register-tiled,
  // 2 = shared-memory kernel.  G, L, X: nb x nb column-major (ld), stats[2]; cycles: 8 clock64 phase sums of
  // thread 0 (variant 0 only) or nullptr.
  static void block_step(cudaStream_t st, int variant, int nb, const cplx* G, cplx* L, cplx* X, long ld, double* stats,
                         long long* cycles);

 private:
  void ensure(int m, int n);
  void pass(cudaStream_t st, const cplx* A, int m, int n, cplx* L, cplx* Q);
  cplx *G_ = nullptr, *L1_ = nullptr, *L2_ = nullptr, *Linv_ = nullptr, *T_ = nullptr, *Q1_ = nullptr;
  double *d_stats_ = nullptr, *h_stats_ = nullptr;
  double last_ratio_ = 0.0;
  // side stream of the factorisation (pass()): the rows of inv(L) are assembled beside the Cholesky chain
  cudaStream_t side_ = nullptr;
  cudaEvent_t ev_diag_ = nullptr, ev_inv_ = nullptr;
  cplx* part_ = nullptr;  // split-K partial sums of the inverse rows
  size_t part_cap_ = 0;
 public:
  long n_single = 0, n_double = 0, n_fail = 0, n_series = 0;
  double ratio_hist[8] = {0};  // counts of pivot ratio in [1,2),[2,4),[4,8),[8,16),[16,64),[64,1e3),[1e3,1e6),fail
 private:
  static constexpr double kSinglePassRatio = 8.0;
  int ncap_ = 0;
  size_t mcap_ = 0;
};

}  // namespace tevo
