/* libtevo - B200-native (sm_100a) two-site MPS update: TEBD gate-apply + truncated SVD split, two-site TDVP
 * (H_eff apply, Lanczos exp, one-site backward step), behind a plain C ABI.
 *
 * This header is the drop-in boundary for orialb/TimeEvoMPS.jl.  The reference has no FFI: its boundary is the
 * set of Julia call sites where src/ hands arrays to ITensors.jl / KrylovKit.jl.  Every entry point below names
 * the reference call site(s) it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - all indices (sites, bonds) are 0-based; bond b joins sites b and b+1   (reference: 1-based)
 *   - complex numbers are (re, im) pairs of doubles = C `double _Complex` = Julia `ComplexF64`
 *   - a site tensor is column-major (l, s, r): element (l,s,r) at  l + chil*(s + d*r)
 *   - a two-site gate is a row-major d^2 x d^2 matrix, G[(s1',s2'),(s1,s2)], s1 the slow index
 *   - an MPO tensor is column-major (wl, s_out, s_in, wr)
 *   - every function returns an int status (TEVO_OK == 0); tevo_last_error() gives the message
 *   - host pointers belong to the caller and are never retained after return; device memory belongs to the handle
 *   - a ctx and the handles created from it must be used from one host thread at a time; there is no CPU fallback
 */
#ifndef TEVO_H
#define TEVO_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define TEVO_OK 0
#define TEVO_EINVAL 1
#define TEVO_ECUDA 2
#define TEVO_ENOMEM 3
#define TEVO_ENOTCONVERGED 4 /* "exponentiate did not converge" (src/tdvp.jl:63,83) */
#define TEVO_EINTERNAL 5

typedef struct tevo_ctx tevo_ctx;
typedef struct tevo_mps tevo_mps;
typedef struct tevo_mpo tevo_mpo;
typedef struct tevo_env tevo_env; /* ITensors.ProjMPO: MPO + cached left/right environments */

typedef struct {
  double re, im;
} tevo_complex;

/* truncation keywords of ITensors.replacebond! as forwarded by apply_gate! / tdvp! (src/bondgate.jl:32-44,51;
 * src/tdvp.jl:14-22,64).  Truncation happens only if has_maxdim or has_cutoff is set (upstream svd()). */
typedef struct {
  int has_maxdim;
  int maxdim;
  int mindim; /* < 1 is treated as 1 */
  int has_cutoff;
  double cutoff;
  int use_absolute_cutoff; /* `use_absolute_cutoff` (bondgate.jl:44) a.k.a. `absoluteCutoff` (tdvp.jl:22) */
  int do_rel_cutoff;       /* upstream `doRelCutoff`, default 1 */
} tevo_trunc;

/* ITensors.Spectrum of one update plus what SpecCallback reads from it (src/callback.jl:139-160) */
typedef struct {
  int nkeep;       /* length(eigs(spec)) = new bond dimension */
  int nfull;       /* number of singular values before truncation */
  double truncerr; /* truncerror(spec) */
  double entropy;  /* entropy(spec) = -sum p ln p over kept p > 1e-13, p un-renormalised */
  double sum_all;  /* sum of all sigma^2 before truncation */
  double sum_kept; /* sum of kept sigma^2 (the state is divided by its square root when normalize != 0) */
} tevo_spec;

/* ---------------------------------------------------------------- context ------------------------------- */
int tevo_ctx_create(int device, tevo_ctx** out);
int tevo_ctx_destroy(tevo_ctx* ctx);
const char* tevo_last_error(tevo_ctx* ctx);
int tevo_ctx_synchronize(tevo_ctx* ctx);
/* the CUDA stream (cudaStream_t) all work of this ctx is issued on - for event timing by the caller */
int tevo_ctx_stream(tevo_ctx* ctx, void** stream_out);
/* counters: kernels launched by this library on this ctx since creation / since last reset */
int tevo_ctx_launch_count(tevo_ctx* ctx, long long* out, int reset);
const char* tevo_version(void);

/* ---------------------------------------------------------------- MPS state ----------------------------- */
/* ITensors MPS marshalling: the shim permutes each site ITensor to (l,s,r) and uploads it. */
int tevo_mps_create(tevo_ctx* ctx, int nsites, int d, tevo_mps** out);
int tevo_mps_free(tevo_mps* psi);
int tevo_mps_copy(tevo_mps* src, tevo_mps** out); /* deepcopy(psi) */
int tevo_mps_upload_site(tevo_mps* psi, int j, int chil, int chir, const tevo_complex* host);
int tevo_mps_download_site(tevo_mps* psi, int j, tevo_complex* host, size_t capacity_elems);
int tevo_mps_site_dims(tevo_mps* psi, int j, int* chil, int* chir);
int tevo_mps_bonddims(tevo_mps* psi, int* out /* nsites-1 */); /* maxlinkdim(psi) (src/tebd.jl:151) */
int tevo_mps_get_limits(tevo_mps* psi, int* llim, int* rlim);  /* ITensors leftlim/rightlim, 0-based */
int tevo_mps_set_limits(tevo_mps* psi, int llim, int rlim);
/* random saturated MPS generated on the device (bench input): iid complex Gaussian sites,
 * chi_j = min(d^j, d^(N-j), chi), then right-canonicalised to site 0 and normalised */
int tevo_mps_random(tevo_mps* psi, int chi, unsigned long long seed);

/* orthogonalize!(psi, j)  (src/bondgate.jl:48, src/tdvp.jl:51, src/itensor.jl:31) */
int tevo_mps_orthogonalize(tevo_mps* psi, int j);
/* reorthogonalize!(psi)   (src/itensor.jl:28-33; called at src/tebd.jl:174) */
int tevo_mps_reorthogonalize(tevo_mps* psi);
/* inner(a, b) = <a|b>     (test/test_tebd.jl:16,28,46) */
int tevo_mps_inner(tevo_mps* a, tevo_mps* b, tevo_complex* out);
/* measure!(psi, op, j): moves the centre to j, returns <psi|O_j|psi>  (src/testutils.jl:55-58). op: row-major dxd */
int tevo_mps_expect_site(tevo_mps* psi, int j, const tevo_complex* op, tevo_complex* out);

/* ---------------------------------------------------------------- TEBD ---------------------------------- */
/* apply_gate!(psi, G; ortho, normalize, maxdim, mindim, cutoff, ...) -> Spectrum   (src/bondgate.jl:46-53):
 * centre -> b, theta = psi[b]*psi[b+1], theta' = G*theta, truncated-SVD split back (replacebond!).
 * ortho_left != 0: psi[b] = U, psi[b+1] = S*V^dagger, centre ends at b+1 (upstream ortho="left").
 * eigs (optional, capacity eigs_cap doubles) receives the kept sigma^2 in descending order. */
int tevo_apply_gate(tevo_mps* psi, int b, const tevo_complex* G, const tevo_trunc* trunc, int ortho_left,
                    int normalize, tevo_spec* spec, double* eigs, int eigs_cap);

/* apply_gates!(psi, Gs; reversed, cb, ...)   (src/bondgate.jl:55-63): applies the ngates gates of one Trotter
 * layer SEQUENTIALLY (forward order, or reversed), moving the orthogonality centre between gates, with no host
 * round trip between gates except the bond-dimension readback.
 *   bonds[g], Gs[g*d^4 ...]: gate g;   specs[g], eigs[g*eigs_stride ...]: its Spectrum (in application order
 *   index g refers to the g-th entry of bonds[], not to the order of application)
 *   nops > 0: after gate g, the local expectation values of the nops single-site operators ops[o] (row-major
 *   dxd) on sites bonds[g] and bonds[g]+1 are computed from the updated two-site wavefunction exactly where
 *   LocalMeasurementCallback does (src/callback.jl:89-95) and stored as
 *   meas[(g*nops + o)*2 + {0: site b, 1: site b+1}].
 *   keep_ortho_left != 0 reproduces the reference literally (ortho="left" for every gate, src/bondgate.jl:58-61);
 *   0 lets reversed layers use ortho="right" (one centre hop per gate instead of three; same state). */
int tevo_apply_layer(tevo_mps* psi, int ngates, const int* bonds, const tevo_complex* Gs, int reversed,
                     const tevo_trunc* trunc, int normalize, int keep_ortho_left, tevo_spec* specs, double* eigs,
                     int eigs_stride, int nops, const tevo_complex* ops, tevo_complex* meas);

/* <theta|O_o|theta> for theta = psi[b]*psi[b+1] on sites b and b+1 (src/callback.jl:65-73,89-95).  The centre
 * must already be at b or b+1 (it is, right after an update of bond b).  out[o*2 + {0,1}]. */
int tevo_measure_bond(tevo_mps* psi, int b, int nops, const tevo_complex* ops, tevo_complex* out);
/* scalar(dag(wf)*noprime(h*wf)) with the centre moved to b: one term of measure(H::GateList, psi)
 * (src/bondgate.jl:71-79).  h: row-major d^2 x d^2. */
int tevo_gate_expect(tevo_mps* psi, int b, const tevo_complex* h, tevo_complex* out);

/* ------------------------------------------------ layer-parallel TEBD (multi-GPU building block) ------- */
/* NOT the reference's algorithm: the reference applies the gates of a layer sequentially with exact centre moves
 * (src/bondgate.jl:48,55-63).  Chain-segment sharding needs independent gates, i.e. the canonical form
 * B_j (right-orthonormal) + Schmidt values on every bond with Hastings' inverse-free update.  Identical to the
 * sequential sweep without truncation; O(discarded weight) apart with it. */
int tevo_mps_to_bform(tevo_mps* psi);
int tevo_apply_layer_bform(tevo_mps* psi, int ngates, const int* bonds, const tevo_complex* Gs, const tevo_trunc* trunc,
                           int normalize, tevo_spec* specs, double* eigs, int eigs_stride);
int tevo_bform_expect_site(tevo_mps* psi, int j, int nops, const tevo_complex* ops, tevo_complex* out);
int tevo_mps_get_lambda(tevo_mps* psi, int j, double* host, int cap, int* n_out);
int tevo_mps_set_lambda(tevo_mps* psi, int j, int n, const double* host);
int tevo_mps_mark_bform(tevo_mps* psi, int on);
/* raw device views / device-side setters for the NCCL peer-to-peer exchange of boundary tensors */
int tevo_mps_site_devptr(tevo_mps* psi, int j, void** ptr, int* chil, int* chir);
int tevo_mps_lambda_devptr(tevo_mps* psi, int j, void** ptr, int* n);
int tevo_mps_set_site_from_device(tevo_mps* psi, int j, int chil, int chir, const void* dev_site);
int tevo_mps_set_lambda_from_device(tevo_mps* psi, int j, int n, const void* dev_lam);

/* ---------------------------------------------------------------- TDVP ---------------------------------- */
int tevo_mpo_create(tevo_ctx* ctx, int nsites, int d, tevo_mpo** out);
int tevo_mpo_free(tevo_mpo* H);
int tevo_mpo_upload_site(tevo_mpo* H, int j, int wl, int wr, const tevo_complex* host /* (wl,s_out,s_in,wr) */);
/* inner(psi, H, phi) = <a|H|b>   (test/test_tdvp.jl:29-33) */
int tevo_mps_inner_mpo(tevo_mps* a, tevo_mpo* H, tevo_mps* b, tevo_complex* out);

/* PH = ProjMPO(H)  (src/tdvp.jl:52) */
int tevo_env_create(tevo_mpo* H, tevo_env** out);
int tevo_env_free(tevo_env* env);
/* nsite=1|2 then position!(PH, psi, pos)  (src/tdvp.jl:53,58-59,79-80) */
int tevo_env_position(tevo_env* env, tevo_mps* psi, int pos, int nsite);
/* (PH)(v): the effective Hamiltonian applied to a host vector at the current position - exposed for tests.
 * v/out: (chil, d, d, chir) for nsite=2, (chil, d, chir) for nsite=1, column-major. */
int tevo_env_apply_host(tevo_env* env, tevo_mps* psi, int pos, int nsite, const tevo_complex* v, tevo_complex* out);

typedef struct {
  int hermitian;    /* `hermitian` kw (src/tdvp.jl:40): Lanczos if != 0 (Arnoldi otherwise) */
  double tol;       /* `exp_tol` (src/tdvp.jl:41) = KrylovKit tol, error per unit time */
  int krylovdim;    /* src/tdvp.jl:42 */
  int maxiter;      /* src/tdvp.jl:43 */
} tevo_krylov;

typedef struct {
  int converged; /* info.converged */
  int numops;    /* operator applications */
  int numiter;   /* restarts */
} tevo_krylov_info;

/* one forward two-site step of the sweep (src/tdvp.jl:58-64): position!(PH,psi,b) with nsite=2,
 * wf = psi[b]*psi[b+1], wf = exponentiate(PH, t, wf), replacebond!(psi,b,wf; ortho, normalize, trunc...).
 * t is the full exponent prefactor (the caller passes -tau/2).  Returns TEVO_ENOTCONVERGED like the throw at :63 */
int tevo_tdvp_twosite(tevo_env* env, tevo_mps* psi, int b, tevo_complex t, const tevo_krylov* kp,
                      const tevo_trunc* trunc, int ortho_left, int normalize, tevo_spec* spec, double* eigs,
                      int eigs_cap, tevo_krylov_info* info);
/* the one-site backward step (src/tdvp.jl:77-83): position!(PH,psi,i) with nsite=1,
 * psi[i] = exponentiate(PH, t, psi[i]) (the caller passes +tau/2) */
int tevo_tdvp_onesite(tevo_env* env, tevo_mps* psi, int i, tevo_complex t, const tevo_krylov* kp,
                      tevo_krylov_info* info);
/* psi[i] /= norm(psi[i])  (src/tdvp.jl:84-87) */
int tevo_mps_normalize_site(tevo_mps* psi, int i);

/* ---------------------------------------------------------------- test hooks ---------------------------- */
/* C = alpha*op(A)*op(B) + beta*C on host buffers through the DMMA ZGEMM kernel (op: 0=N, 1=T, 2=C) */
int tevo_test_zgemm(tevo_ctx* ctx, int ta, int tb, int M, int N, int K, tevo_complex alpha, const tevo_complex* A,
                    int lda, const tevo_complex* B, int ldb, tevo_complex beta, tevo_complex* C, int ldc);
/* device-resident timing of the DMMA ZGEMM: average milliseconds per launch over `reps` launches of
 * C(MxN) = op(A) op(B) (+ C if accumulate); kind 0 = general, 1 = Hermitian result (N ignored, lower tiles + mirror) */
int tevo_test_zgemm_time(tevo_ctx* ctx, int kind, int ta, int tb, int M, int N, int K, int accumulate, int reps,
                         double* ms_out);
/* truncated split of a host m x n matrix: returns L (m x k) and R (k x n) with L*R ~ theta */
int tevo_test_split(tevo_ctx* ctx, int m, int n, const tevo_complex* theta, const tevo_trunc* trunc, int ortho_left,
                    int normalize, tevo_spec* spec, double* eigs, int eigs_cap, tevo_complex* L, tevo_complex* R,
                    int kcap);

/* Hermitian eigen-decomposition of a host n x n matrix (full storage, column-major): eigenvalues in descending
 * order and the matching eigenvectors (columns of U) - the core of the split, exposed for tests */
int tevo_test_heig(tevo_ctx* ctx, int n, const tevo_complex* A, double* evals_desc, tevo_complex* U);
/* same with the tridiagonalisation kernel limited to max_ctas CTAs (0 = all SMs): the configuration the lanes of the
 * layer-parallel mode run in (several 16-row strips per CTA) */
int tevo_test_heig_ctas(tevo_ctx* ctx, int n, const tevo_complex* A, double* evals_desc, tevo_complex* U,
                        int max_ctas);

/* the two-stage tridiagonalisation (dense -> band of width 8 -> tridiagonal) stage by stage on host buffers:
 * band_out (16 x n, column-major: B(c+d, c) at [d + 16 c]), v_out (n x n Householder vectors of stage 1), taus_out (n),
 * d_out (n) / e_out (n-1) the tridiagonal matrix, and optionally x_out (n x k) = Q2 * x_real (n x k real) */
int tevo_test_sbr(tevo_ctx* ctx, int n, const tevo_complex* A, tevo_complex* band_out, tevo_complex* v_out,
                  tevo_complex* taus_out, double* d_out, double* e_out, int k, const double* x_real, tevo_complex* x_out);
/* one diagonal-block step of the CholeskyQR factorisation (csrc/cholqr.cu) on host buffers: G (nb x nb, nb <= 64,
 * column-major, Hermitian positive definite, lower part read) -> L_out = chol(G) (lower, real diagonal), X_out =
 * inv(L), stats_out[0..1] = smallest / largest pivot (stats_out[0] < 0: not positive definite).  variant 0 = panel
 * kernel (default), 1 = register-tiled kernel, 2 = shared-memory kernel.  reps > 0: us_out = time per launch over
 * reps back-to-back launches; cycles_out (8 doubles or NULL): clock64 phase sums of variant 0 */
int tevo_test_potrf_block(tevo_ctx* ctx, int variant, int nb, const tevo_complex* G, tevo_complex* L_out,
                          tevo_complex* X_out, double* stats_out, int reps, double* us_out, double* cycles_out);
/* CTA tile of the DMMA ZGEMM (process-wide; tests): -1 = automatic by problem size (default), 0 = 128 x 64,
 * 1 = 64 x 32, 2 = 32 x 32 (csrc/zgemm.cu, pick_tile) */
int tevo_debug_set_zgemm_tile(int tile);
/* complex products of the 64 x 32 / 32 x 32 ZGEMM kernels (process-wide; tests): 1 = three real multiplications per
 * complex one (default), 0 = four.  The Rayleigh-quotient GEMM of the split always uses four. */
int tevo_debug_set_zgemm_3m(int on);
/* Hermitian eigensolver variant used by the split (process-wide): 0 = default (one-stage blocked Householder
 * reduction; the two-stage one if the environment sets TEVO_HEIG_TWOSTAGE), 1 = one-stage, 2 = two-stage reduction
 * dense -> band of width 8 -> tridiagonal (n <= 2056; parity-tested, not yet faster: DESIGN.md) */
int tevo_debug_set_heig_mode(int mode);
/* cumulative counters (development / bench.py): out[0..2] CholeskyQR single / two-pass / eigen-fallback centre moves,
 * [3..10] pivot-ratio histogram, [11] H_eff flops, [12..14] one-stage tridiagonalisation: factorisations, algorithmic
 * bytes, panel launches, [15..16] two-stage: factorisations, algorithmic bytes of the stage-1 passes */
int tevo_debug_counters(tevo_ctx* ctx, double* out, int n);

/* per-phase device timing with CUDA events on the library stream (bench.py's roofline object).  Off by default. */
int tevo_profile_enable(int on);
int tevo_profile_read(char* buf, size_t cap, int reset);

#ifdef __cplusplus
}
#endif
#endif /* TEVO_H */
