#pragma once
#include "common.cuh"
namespace tevo {
// C = alpha*op(A)*op(B) + beta*C, column-major, complex FP64 on the DMMA pipe.
void zgemm(cudaStream_t st, Op ta, Op tb, int M, int N, int K, cplx alpha, const cplx* A, long lda, const cplx* B,
           long ldb, cplx beta, cplx* C, long ldc);
void zgemm_splitk(cudaStream_t st, Op ta, Op tb, int M, int N, int K, cplx alpha, const cplx* A, long lda, const cplx* B,
                  long ldb, cplx beta, cplx* C, long ldc, cplx** part, size_t* part_cap);
// C (N x N) Hermitian: lower-triangle tiles only (+ optional mirror of conj(lower) into the strict upper part)
void zgemm_herm(cudaStream_t st, Op ta, Op tb, int N, int K, cplx alpha, const cplx* A, long lda, const cplx* B, long ldb,
                cplx beta, cplx* C, long ldc, bool mirror);
void zgemm_batched(cudaStream_t st, Op ta, Op tb, int M, int N, int K, cplx alpha, const cplx* A, long lda, long bsA,
                   const cplx* B, long ldb, long bsB, cplx beta, cplx* C, long ldc, long bsC, int batch);
extern long g_zgemm_launches;
// CTA tile selection (tests): -1 = automatic (by problem size), 0 = 128 x 64, 1 = 64 x 32, 2 = 32 x 32
void zgemm_force_tile(int tile);
// three-multiplication complex products in the small-tile kernels (default on); 0 = four multiplications everywhere
void zgemm_set_3m(int on);
// while alive on a thread, that thread's products use four multiplications
struct ZgemmPrecise {
  ZgemmPrecise();
  ~ZgemmPrecise();
};
}  // namespace tevo
