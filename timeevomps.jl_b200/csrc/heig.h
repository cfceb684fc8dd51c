// Hermitian eigensolver (device): see heig.cu.
#pragma once
#include "common.cuh"
#include "sbr.h"
#include "stedc.h"

namespace tevo {

// cluster-resident tail of the tridiagonalisation (tail.cu; on by default with a 16-CTA cluster in the single-lane
// configuration, TEVO_CLUSTER_TAIL=<0 | cluster size>)
int tridiag_tail_max_m(int cluster);
void tridiag_tail(cudaStream_t st, int cluster, int n, int t0, cplx* A, long lda, double* d, double* e, cplx* taus);

// cumulative counters of the tridiagonalisation panels (process-wide): see heig.cu
void heig_counters(double* calls, double* bytes, double* launches);
// two-stage path: factorisations and algorithmic bytes of the stage-1 passes (read + write of the trailing matrix per
// block of 8 columns)
void sbr_counters(double* calls, double* bytes);
// 0: default (one-stage; two-stage if TEVO_HEIG_TWOSTAGE is set); 1: one-stage blocked Householder reduction;
// 2: two-stage reduction dense -> band -> tridiagonal where it applies (n <= SBR_MAX_N)
void heig_set_mode(int mode);

class Heig {
 public:
  void init();
  void destroy();
  // Factor the n x n Hermitian matrix A (full storage, column-major; overwritten by the Householder reflectors).
  // Afterwards evals() (ascending, device) is valid and vectors() can be called.  No host synchronisation.
  void factor(cudaStream_t st, int n, cplx* A, long lda);
  const double* evals() const { return evals_; }
  // cap on the CTAs of the panel kernel (two lanes share the SMs in the layer-parallel mode); 0 = all SMs
  void set_max_ctas(int n) { max_ctas_ = n; }
  int max_ctas() const { return max_ctas_; }
  // U(:, i) = eigenvector of column perm[i] (indices into the ascending order), i < k.  U is n x k, ld = ldu.
  void vectors(cudaStream_t st, const int* perm_dev, int k, cplx* U, long ldu);

 private:
  void ensure(int n);
  Stedc stedc_;
  Sbr sbr_;                // two-stage reduction (dense -> band -> tridiagonal), opt-in (see factor())
  bool two_stage_ = false;  // which reduction produced the current factorisation
  cudaStream_t aux_ = nullptr;  // compact-WY set-up beside the divide & conquer
  cudaEvent_t ev_fork_ = nullptr, ev_join_ = nullptr;
  bool wy_pending_ = false;
  int num_sms_ = 148;
  int max_ctas_ = 0;
  int cap_ = 0, n_ = 0;
  cplx* A_ = nullptr;
  long lda_ = 0;
  cplx *P_ = nullptr, *Q_ = nullptr, *Y_ = nullptr, *VT_ = nullptr, *taus_ = nullptr, *Tblk_ = nullptr, *X_ = nullptr, *part_ = nullptr;
  cplx* pair_ = nullptr;  // per pair of panels: V1^H V2, T1 V1^H V2, T12 (64 x 64 each)
  int npairs_ = 0;        // pairs of adjacent full-width panels applied as one 128-wide block reflector
  size_t xcap_ = 0, part_cap_ = 0;
  double *Sall_ = nullptr, *d_ = nullptr, *e_ = nullptr, *evals_ = nullptr, *Z_ = nullptr, *acc_ = nullptr;
};

}  // namespace tevo
