// Two-stage Hermitian tridiagonalisation (dense -> band of width 8 -> tridiagonal): see sbr.cu.
#pragma once
#include "common.cuh"

namespace tevo {

constexpr int SBR_BAND = 8;                   // bandwidth after stage 1
constexpr int SBR_LDB = 2 * SBR_BAND;         // diagonals stored per column (the bulge of stage 2 included)
constexpr int SBR_MAX_N = 2048 + SBR_BAND;    // the panel QR keeps 8 rows x 8 columns per thread in registers
constexpr int SBR_SMALL_DOUBLES = 1024;       // per parity: S accumulator (8x8 complex) and T factor (8x8 complex) + barrier

class Sbr {
 public:
  void init(int num_sms) { num_sms_ = num_sms; }
  void destroy();
  // stage 1: A (n x n Hermitian, full storage, column-major) -> band matrix in Bd_; on return A holds the Householder
  // vectors of the block reflectors (column c: unit entry at row c + 8, explicit zeros above) and taus[c] their
  // scalars (0 for the columns without reflector).  max_ctas caps the cooperative grid (0 = all SMs).
  void reduce_to_band(cudaStream_t st, int n, cplx* A, long lda, cplx* taus, int max_ctas);
  // stage 2: band -> real symmetric tridiagonal (d: n, e: n-1), reflectors kept for apply_q2
  void chase(cudaStream_t st, int n, double* d, double* e);
  // U(:, i) = Q2 * Z(:, perm[i]),  i < k  (Z real n x n with leading dimension ldz; perm on the device, may be null)
  void apply_q2(cudaStream_t st, int n, int k, const double* Z, long ldz, const int* perm_dev, cplx* U, long ldu);
  // test access
  cplx* band() { return Bd_; }
  void ensure(int n);

 private:
  int num_sms_ = 148;
  int cap_ = 0, ks_ = 0;
  cplx *Bd_ = nullptr, *RV_ = nullptr, *Vb_ = nullptr, *Wb_ = nullptr, *Yb_ = nullptr;
  double* small_ = nullptr;
  int* prog_ = nullptr;
};

}  // namespace tevo
