// Optional per-phase device timing (CUDA events on the library stream), used by bench.py for the roofline object.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

namespace tevo {

enum ProfId {
  P_THETA = 0,      // theta = A_b * A_{b+1}
  P_GATE,           // two-site gate application
  P_RHO,            // density / Gram matrix GEMM
  P_TRI_PANEL,      // tridiagonalisation panel kernel (HEMV-bound)
  P_TRI_HER2K,      // trailing rank-2k update GEMM
  P_TRI_WY,         // compact-WY T factor and V*T
  P_STEDC,          // tridiagonal divide & conquer
  P_BACK,           // eigenvector back-transformation
  P_SPLIT_GEMM,     // R = U^+ theta / L = theta V
  P_RESID,          // discarded-weight residual
  P_TRUNC,          // sort / truncate / spectrum kernels + readback
  P_HOP_GEMM,       // centre-move products
  P_HEFF,           // effective Hamiltonian application
  P_KRYLOV_BLAS1,   // Lanczos orthogonalisation / linear combinations
  P_ENV,            // environment updates
  P_MEASURE,        // local expectation values
  P_CHOL_GRAM,      // CholeskyQR: Gram matrix GEMM
  P_CHOL_FACTOR,    // CholeskyQR: blocked Cholesky (diagonal blocks + panel/trailing GEMMs)
  P_CHOL_INV,       // CholeskyQR: triangular inverse assembly
  P_CHOL_Q,         // CholeskyQR: Q = A L^-H
  P_SBR_STAGE1,     // two-stage tridiagonalisation: dense -> band (persistent kernel)
  P_SBR_CHASE,      // two-stage tridiagonalisation: band -> tridiagonal (bulge chasing)
  P_SBR_Q2,         // back-transformation with the stage-2 reflectors
  P_COUNT
};

struct Prof {
  bool on = false;
  struct Rec {
    int id;
    cudaEvent_t a, b;
  };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  double ms[P_COUNT] = {0};
  long count[P_COUNT] = {0};
  cudaEvent_t get();
  void flush();  // synchronises the events and accumulates
  std::string json(bool reset);
};
extern Prof g_prof;
extern thread_local bool tl_prof_thread;  // only the main lane's host thread records

struct ProfScope {
  int id;
  cudaStream_t st;
  cudaEvent_t a = nullptr;
  ProfScope(cudaStream_t s, int i) : id(i), st(s) {
    if (g_prof.on && tl_prof_thread) {
      a = g_prof.get();
      cudaEventRecord(a, st);
    }
  }
  ~ProfScope() {
    if (a) {
      cudaEvent_t b = g_prof.get();
      cudaEventRecord(b, st);
      g_prof.recs.push_back({id, a, b});
      if (g_prof.recs.size() > 4096) g_prof.flush();
    }
  }
};

}  // namespace tevo
