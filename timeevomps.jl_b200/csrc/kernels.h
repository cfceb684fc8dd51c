#pragma once
#include "common.cuh"
namespace tevo {

struct TruncParams {
  int do_truncate;  // upstream svd(): truncate only if maxdim or cutoff was supplied
  int has_maxdim, maxdim, mindim;
  double cutoff;
  int use_absolute_cutoff, do_rel_cutoff;
  double inflate_rel;  // stage-1 (Gram spectrum) safety margin: every weight is raised by inflate_rel * largest weight
};
struct TruncInfo {
  int nkeep, nfull;
  double truncerr, entropy, sum_all, sum_kept;
  int bad;
  int identity;  // stage 2: kept rows are already in descending order
};

void gate_apply(cudaStream_t st, cplx* theta, int chil, int d, int chir, const cplx* G_dev);
void conj_transpose(cudaStream_t st, int m, int n, const cplx* in, long ldi, cplx* out, long ldo, bool do_conj);
void set_identity(cudaStream_t st, int n, cplx* V, long ld);
void jacobi_round(cudaStream_t st, cplx* X, long ldx, int nrx, cplx* V, long ldv, int nrv, int nc, int round,
                  double tol, int* rotcount);
void colnorm2(cudaStream_t st, const cplx* X, long ld, int nr, int nc, double* out);
void sort_truncate(cudaStream_t st, const double* lam, int nc, int nfull, const TruncParams& tp, double* sorted,
                   int* perm, TruncInfo* info);
// stage 2 of the split: sort the Rayleigh-refined weights of the k1 candidate rows, continue upstream's truncate! loop
// from n = k1 with the accurately known discarded weight *resid2 (absolute), fill the final Spectrum summary
void refine_truncate(cudaStream_t st, const double* lam2, int k1, int nfull, const double* sum_all_dev,
                     const double* resid2_dev, bool resid_by_difference, const TruncParams& tp, double* sorted, int* perm,
                     TruncInfo* info);
void gather_rows(cudaStream_t st, int k, int n, const cplx* src, long lds, const int* perm, cplx* dst, long ldd,
                 const double* norm2);
void gather_cols(cudaStream_t st, int nr, int k, const cplx* src, long lds, const int* perm, cplx* dst, long ldd,
                 const double* norm2);
void gather_cols_conjT(cudaStream_t st, int nr, int k, const cplx* src, long lds, const int* perm, cplx* dst,
                       long ldd, const double* norm2);
void zscal(cudaStream_t st, long n, cplx a, cplx* x);
void zscal_rsqrt(cudaStream_t st, long n, const double* norm2_dev, cplx* x);
void zdotc(cudaStream_t st, long n, const cplx* x, const cplx* y, double* out2_dev);
void expect2(cudaStream_t st, const cplx* theta, int chil, int d, int chir, int nops, const cplx* ops_dev,
             double* out_dev);
void mid_contract(cudaStream_t st, int na, int P, int Q, int nc, const cplx* Mx_dev, const long* oq_dev, long oc,
                  const cplx* in, cplx* out);
void multi_dotc(cudaStream_t st, long n, int k, const cplx* V, long ldv, const cplx* r, double* out_dev);
void multi_axpy_sub(cudaStream_t st, long n, int k, const cplx* V, long ldv, const double* h_dev, cplx* r);
void lincomb(cudaStream_t st, long n, int k, const cplx* V, long ldv, const double* y_dev, double scale, cplx* out);
void copy_scale(cudaStream_t st, long n, double s, const cplx* in, cplx* out);
void scale_rows(cudaStream_t st, int m, long n, int chil, const double* lam, const cplx* in, cplx* out);
void sqrt_scaled(cudaStream_t st, int n, const double* in, const double* denom, double* out);
void expect1(cudaStream_t st, const cplx* B, int chil, int d, int chir, const double* lam, int nops, const cplx* ops_dev,
             double* out_dev);

}  // namespace tevo
