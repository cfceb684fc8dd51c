// Divide-and-conquer eigensolver for a real symmetric tridiagonal matrix, entirely on the device
// (Cuppen tearing, Gu-Eisenstat stable eigenvectors, DMMA GEMM back-multiplication).  Part of the Hermitian
// eigensolver behind the truncated split (replacebond!, src/bondgate.jl:51, src/tdvp.jl:64).
#pragma once
#include <map>
#include <vector>

#include "common.cuh"

namespace tevo {

struct MergeDesc {
  int lo, mid, hi;
};

void dgemm_merge(cudaStream_t st, const MergeDesc* descs_dev, const int* K_dev, int nmerge, int maxN, const double* A,
                 const double* B, double* C, long ld);

class Stedc {
 public:
  void destroy();
  // d (n), e (n-1) on the device.  On return: evals (n, ascending) and Z (n x n column-major, ld = n) on the
  // device hold the eigen-decomposition.  d/e are not modified.  No host synchronisation.
  void run(cudaStream_t st, int n, const double* d, const double* e, double* evals, double* Z);
  static constexpr int LEAF = 32;

 private:
  struct Plan {
    std::vector<std::vector<MergeDesc>> levels;
    MergeDesc* dev = nullptr;  // all levels concatenated
    std::vector<int> level_off;
    std::vector<int> level_maxN;
  };
  std::map<int, Plan> plans_;
  Plan& plan_for(int n);
  void ensure(int n);
  int cap_ = 0;
  // workspaces (cap_ sized)
  double *Zt_ = nullptr, *Um_ = nullptr, *Cw_ = nullptr;
  double *zz_ = nullptr, *dd_ = nullptr, *ddefl_ = nullptr, *mu_ = nullptr, *zhat_ = nullptr, *lam_ = nullptr;
  double *rot_c_ = nullptr, *rot_s_ = nullptr, *scale_ = nullptr, *rho_ = nullptr;
  int *org_ = nullptr, *gsrc_ = nullptr, *rot_a_ = nullptr, *rot_b_ = nullptr, *K_ = nullptr, *nrot_ = nullptr;
};

}  // namespace tevo
