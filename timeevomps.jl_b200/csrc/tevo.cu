// libtevo: C ABI (include/tevo.h) + host-side orchestration of the two-site update on one B200.
// All arithmetic runs in the CUDA kernels of zgemm.cu / kernels.cu / heig.cu; there is no CPU fallback.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <exception>
#include <memory>
#include <thread>
#include <vector>

#include "../../include/tevo.h"
#include "common.cuh"
#include "kernels.h"
#include "prof.h"
#include "sbr.h"
#include "split.h"
#include "zgemm.h"

using namespace tevo;

// =====================================================================================================
// context
// =====================================================================================================
struct tevo_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  Split split;  // owns workspaces of the truncated split
  std::vector<DevBuf> pool;
  // generic workspaces
  DevBuf ws[8];
  unsigned char* d_small = nullptr;  // staging for gates / operators / offsets (device)
  size_t small_cap = 0;
  unsigned char* d_heff = nullptr;  // H_eff operator block (P x Q complex + Q offsets), sized to need (make_heff)
  size_t heff_cap = 0;
  void* heff_block(size_t bytes) {
    if (heff_cap < bytes) {
      if (d_heff) {
        TEVO_CUDA_CHECK(cudaStreamSynchronize(stream));
        TEVO_CUDA_CHECK(cudaFree(d_heff));
      }
      d_heff = nullptr;
      heff_cap = 0;
      const size_t want = std::max<size_t>(bytes + bytes / 2, 1 << 16);
      TEVO_CUDA_CHECK(cudaMalloc(&d_heff, want));
      heff_cap = want;
    }
    return d_heff;
  }
  double* d_scal = nullptr;  // device scalars
  double* h_scal = nullptr;  // pinned mirror
  static constexpr int NSCAL = 1024;
  DevBuf ones;  // 1x1x1 identity environment
  std::vector<tevo_ctx*> aux;  // extra lanes (own stream / workspaces) for the layer-parallel mode, created lazily
  bool is_aux = false;

  cplx* get_ws(int slot, size_t nelem) {
    DevBuf& b = ws[slot];
    if (b.cap < nelem) {
      if (b.p) TEVO_CUDA_CHECK(cudaFree(b.p));
      b.p = nullptr;
      b.cap = 0;
      size_t want = nelem + nelem / 8 + 64;
      TEVO_CUDA_CHECK(cudaMalloc(&b.p, want * sizeof(cplx)));
      b.cap = want;
    }
    return b.p;
  }
  void* small(size_t bytes) {
    if (small_cap < bytes) {
      if (d_small) TEVO_CUDA_CHECK(cudaFree(d_small));
      d_small = nullptr;
      small_cap = 0;
      size_t want = std::max<size_t>(bytes * 2, 1 << 16);
      TEVO_CUDA_CHECK(cudaMalloc(&d_small, want));
      small_cap = want;
    }
    return d_small;
  }
  DevBuf alloc(size_t nelem) {
    if (nelem == 0) nelem = 1;
    int best = -1;
    for (int i = 0; i < (int)pool.size(); ++i)
      if (pool[i].cap >= nelem && pool[i].cap <= 2 * nelem + 1024 && (best < 0 || pool[i].cap < pool[best].cap)) best = i;
    if (best >= 0) {
      DevBuf b = pool[best];
      pool.erase(pool.begin() + best);
      return b;
    }
    DevBuf b;
    TEVO_CUDA_CHECK(cudaMalloc(&b.p, nelem * sizeof(cplx)));
    b.cap = nelem;
    return b;
  }
  void release(DevBuf& b) {
    if (!b.p) return;
    if (pool.size() >= 24) {
      // drop the smallest pooled buffer to bound the pool
      int s = 0;
      for (int i = 1; i < (int)pool.size(); ++i)
        if (pool[i].cap < pool[s].cap) s = i;
      if (pool[s].cap < b.cap) {
        cudaStreamSynchronize(stream);
        cudaFree(pool[s].p);
        pool[s] = b;
      } else {
        cudaStreamSynchronize(stream);
        cudaFree(b.p);
      }
    } else {
      pool.push_back(b);
    }
    b.p = nullptr;
    b.cap = 0;
  }
};

enum { WS_THETA = 0, WS_T1 = 1, WS_T2 = 2, WS_KRYLOV = 3, WS_W = 4, WS_E0 = 5, WS_E1 = 6, WS_MISC = 7 };

struct tevo_mps {
  tevo_ctx* ctx;
  int N, d;
  std::vector<DevBuf> site;
  std::vector<int> chi;  // N+1
  int llim, rlim;
  // layer-parallel (B-form) mode: Schmidt values on the bond left of site j (device), valid when bform
  bool bform = false;
  std::vector<double*> lam;  // N+1 device arrays (lam[j] has chi[j] entries)
  std::vector<int> lam_cap;
};

struct tevo_mpo {
  tevo_ctx* ctx;
  int N, d;
  std::vector<int> w;                               // N+1
  std::vector<std::vector<std::complex<double>>> hW;  // host copies (wl, so, si, wr) column-major
};

struct tevo_env {
  tevo_mpo* H;
  tevo_ctx* ctx;
  int N;
  std::vector<DevBuf> LR;
  std::vector<int> dim_bra, dim_w, dim_ket;
  int lpos, rpos;
};

static thread_local std::string g_static_err;

#define TEVO_TRY(ctxptr) \
  tevo_ctx* _ctx = (ctxptr); \
  try { \
    if (_ctx) TEVO_CUDA_CHECK(cudaSetDevice(_ctx->device)); /* a process may hold contexts on several GPUs */
#define TEVO_CATCH                                                                                       \
  }                                                                                                      \
  catch (const tevo::CudaError& e) {                                                                     \
    std::string m = std::string("CUDA error: ") + cudaGetErrorString(e.err) + " at " + e.file + ":" +    \
                    std::to_string(e.line);                                                              \
    if (_ctx) _ctx->err = m; else g_static_err = m;                                                      \
    return e.err == cudaErrorMemoryAllocation ? TEVO_ENOMEM : TEVO_ECUDA;                                \
  }                                                                                                      \
  catch (const tevo::TevoError& e) {                                                                     \
    if (_ctx) _ctx->err = e.msg; else g_static_err = e.msg;                                              \
    return e.code;                                                                                       \
  }                                                                                                      \
  catch (const std::bad_alloc&) {                                                                        \
    if (_ctx) _ctx->err = "host out of memory"; else g_static_err = "host out of memory";                \
    return TEVO_ENOMEM;                                                                                  \
  }                                                                                                      \
  catch (...) {                                                                                          \
    if (_ctx) _ctx->err = "internal error"; else g_static_err = "internal error";                        \
    return TEVO_EINTERNAL;                                                                               \
  }                                                                                                      \
  return TEVO_OK;

static inline void require(bool ok, const char* msg) {
  if (!ok) throw TevoError(TEVO_EINVAL, msg);
}

static TruncParams to_params(const tevo_trunc* t) {
  TruncParams p;
  memset(&p, 0, sizeof(p));
  p.mindim = 1;
  p.do_rel_cutoff = 1;
  if (t) {
    p.do_truncate = (t->has_maxdim || t->has_cutoff) ? 1 : 0;
    p.has_maxdim = t->has_maxdim;
    p.maxdim = t->maxdim;
    p.mindim = t->mindim < 1 ? 1 : t->mindim;
    p.cutoff = t->has_cutoff ? t->cutoff : 0.0;
    p.use_absolute_cutoff = t->use_absolute_cutoff;
    p.do_rel_cutoff = t->do_rel_cutoff;
  }
  return p;
}

static void fill_spec(const TruncInfo& ti, tevo_spec* s) {
  if (!s) return;
  s->nkeep = ti.nkeep;
  s->nfull = ti.nfull;
  s->truncerr = ti.truncerr;
  s->entropy = ti.entropy;
  s->sum_all = ti.sum_all;
  s->sum_kept = ti.sum_kept;
}

static const cplx ONE = {1.0, 0.0}, ZERO = {0.0, 0.0};

// =====================================================================================================
// MPS internals
// =====================================================================================================
static void set_site(tevo_mps* psi, int j, DevBuf nb) {
  psi->ctx->release(psi->site[j]);
  psi->site[j] = nb;
  for (auto& b : psi->ctx->split.take_released()) psi->ctx->release(b);
}

// centre j -> j+1 (QR-equivalent hop of orthogonalize!; here an untruncated split U * (S V^+))
static void hop_right(tevo_mps* psi, int j) {
  tevo_ctx* c = psi->ctx;
  psi->bform = false;
  const int d = psi->d, m = psi->chi[j] * d, n = psi->chi[j + 1];
  SplitResult r = c->split.ortho_factor(c->stream, psi->site[j].p, m, n, /*left=*/true,
                                        [&](size_t ne) { return c->alloc(ne); });
  const int k = r.k, n2 = d * psi->chi[j + 2];
  DevBuf nb = c->alloc((size_t)k * n2);
  {
    ProfScope ps(c->stream, P_HOP_GEMM);
    zgemm(c->stream, OP_N, OP_N, k, n2, n, ONE, r.R.p, k, psi->site[j + 1].p, n, ZERO, nb.p, k);
  }
  c->release(r.R);
  set_site(psi, j, r.L);
  set_site(psi, j + 1, nb);
  psi->chi[j + 1] = k;
}

// centre j -> j-1
static void hop_left(tevo_mps* psi, int j) {
  tevo_ctx* c = psi->ctx;
  psi->bform = false;
  const int d = psi->d, m = psi->chi[j], n = d * psi->chi[j + 1];
  SplitResult r = c->split.ortho_factor(c->stream, psi->site[j].p, m, n, /*left=*/false,
                                        [&](size_t ne) { return c->alloc(ne); });
  const int k = r.k, m2 = psi->chi[j - 1] * d;
  DevBuf nb = c->alloc((size_t)m2 * k);
  {
    ProfScope ps(c->stream, P_HOP_GEMM);
    zgemm(c->stream, OP_N, OP_N, m2, k, m, ONE, psi->site[j - 1].p, m2, r.L.p, m, ZERO, nb.p, m2);
  }
  c->release(r.L);
  set_site(psi, j, r.R);
  set_site(psi, j - 1, nb);
  psi->chi[j] = k;
}

static void orthogonalize(tevo_mps* psi, int j) {
  require(j >= 0 && j < psi->N, "orthogonalize: site out of range");
  while (psi->llim < j - 1) {
    hop_right(psi, psi->llim + 1);
    psi->llim += 1;
    if (psi->rlim < psi->llim + 2) psi->rlim = psi->llim + 2;
  }
  while (psi->rlim > j + 1) {
    hop_left(psi, psi->rlim - 1);
    psi->rlim -= 1;
    if (psi->llim > psi->rlim - 2) psi->llim = psi->rlim - 2;
  }
}

// centre into {b, b+1}: sites < b left-orthonormal, sites > b+1 right-orthonormal
static void center_into_pair(tevo_mps* psi, int b) {
  while (psi->llim < b - 1) {
    hop_right(psi, psi->llim + 1);
    psi->llim += 1;
    if (psi->rlim < psi->llim + 2) psi->rlim = psi->llim + 2;
  }
  while (psi->rlim > b + 2) {
    hop_left(psi, psi->rlim - 1);
    psi->rlim -= 1;
    if (psi->llim > psi->rlim - 2) psi->llim = psi->rlim - 2;
  }
}

static cplx* form_theta(tevo_mps* psi, int b, int slot = WS_THETA) {
  tevo_ctx* c = psi->ctx;
  const int d = psi->d, m = psi->chi[b] * d, k = psi->chi[b + 1], n = d * psi->chi[b + 2];
  cplx* th = c->get_ws(slot, (size_t)m * n);
  ProfScope ps(c->stream, P_THETA);
  zgemm(c->stream, OP_N, OP_N, m, n, k, ONE, psi->site[b].p, m, psi->site[b + 1].p, k, ZERO, th, m);
  return th;
}

// replacebond!(psi, b, theta; ortho, normalize, trunc...) on a device theta (m x n)
static TruncInfo replace_bond(tevo_mps* psi, int b, const cplx* theta, bool left, const TruncParams& tp, bool normalize,
                              double* eigs, int eigs_cap) {
  tevo_ctx* c = psi->ctx;
  psi->bform = false;
  const int d = psi->d, m = psi->chi[b] * d, n = d * psi->chi[b + 2];
  SplitResult r = c->split.run(c->stream, theta, m, n, left, tp, normalize, eigs, eigs_cap,
                               [&](size_t ne) { return c->alloc(ne); });
  set_site(psi, b, r.L);
  set_site(psi, b + 1, r.R);
  psi->chi[b + 1] = r.k;
  if (left) {
    if (psi->llim >= b - 1) psi->llim = b;
    if (psi->rlim <= b + 2) psi->rlim = b + 2;
  } else {
    if (psi->rlim <= b + 2) psi->rlim = b + 1;
    if (psi->llim >= b - 1) psi->llim = b - 1;
  }
  return r.info;
}

static TruncInfo apply_gate_dev(tevo_mps* psi, int b, const cplx* G_dev, const TruncParams& tp, bool left,
                                bool normalize, bool literal_center, double* eigs, int eigs_cap) {
  require(b >= 0 && b < psi->N - 1, "apply_gate: bond out of range");
  tevo_ctx* c = psi->ctx;
  if (literal_center)
    orthogonalize(psi, b);  // src/bondgate.jl:48
  else
    center_into_pair(psi, b);
  cplx* th = form_theta(psi, b);  // src/bondgate.jl:50
  {
    ProfScope ps(c->stream, P_GATE);
    gate_apply(c->stream, th, psi->chi[b], psi->d, psi->chi[b + 2], G_dev);  // G * theta, noprime
  }
  return replace_bond(psi, b, th, left, tp, normalize, eigs, eigs_cap);    // src/bondgate.jl:51
}

static void download_scalars(tevo_ctx* c, const double* dev, int n, double* host_out) {
  require(n <= tevo_ctx::NSCAL, "too many scalars");
  TEVO_CUDA_CHECK(cudaMemcpyAsync(c->h_scal, dev, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  TEVO_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  memcpy(host_out, c->h_scal, sizeof(double) * n);
}

// =====================================================================================================
// C ABI: context
// =====================================================================================================
extern "C" const char* tevo_version(void) { return "libtevo 0.1 (sm_100a)"; }

extern "C" int tevo_ctx_create(int device, tevo_ctx** out) {
  TEVO_TRY(nullptr)
  require(out != nullptr, "tevo_ctx_create: null output");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    throw TevoError(TEVO_ECUDA, "no CUDA device available: libtevo has no CPU fallback");
  require(device >= 0 && device < ndev, "tevo_ctx_create: device out of range");
  TEVO_CUDA_CHECK(cudaSetDevice(device));
  std::unique_ptr<tevo_ctx> c(new tevo_ctx());
  c->device = device;
  TEVO_CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  TEVO_CUDA_CHECK(cudaMalloc(&c->d_scal, sizeof(double) * tevo_ctx::NSCAL));
  TEVO_CUDA_CHECK(cudaMallocHost(&c->h_scal, sizeof(double) * tevo_ctx::NSCAL));
  c->split.init(c->stream);
  c->ones = c->alloc(1);
  cplx one = ONE;
  TEVO_CUDA_CHECK(cudaMemcpyAsync(c->ones.p, &one, sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
  TEVO_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  *out = c.release();
  TEVO_CATCH
}

extern "C" int tevo_ctx_destroy(tevo_ctx* ctx) {
  if (!ctx) return TEVO_OK;
  for (auto* a : ctx->aux) tevo_ctx_destroy(a);
  ctx->aux.clear();
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& b : ctx->pool) cudaFree(b.p);
  for (auto& b : ctx->ws)
    if (b.p) cudaFree(b.p);
  if (ctx->ones.p) cudaFree(ctx->ones.p);
  if (ctx->d_small) cudaFree(ctx->d_small);
  if (ctx->d_heff) cudaFree(ctx->d_heff);
  if (ctx->d_scal) cudaFree(ctx->d_scal);
  if (ctx->h_scal) cudaFreeHost(ctx->h_scal);
  ctx->split.destroy();
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return TEVO_OK;
}

extern "C" const char* tevo_last_error(tevo_ctx* ctx) { return ctx ? ctx->err.c_str() : g_static_err.c_str(); }

extern "C" int tevo_ctx_synchronize(tevo_ctx* ctx) {
  TEVO_TRY(ctx)
  TEVO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  TEVO_CATCH
}

extern "C" int tevo_ctx_stream(tevo_ctx* ctx, void** stream_out) {
  TEVO_TRY(ctx)
  require(stream_out, "null");
  *stream_out = (void*)ctx->stream;
  TEVO_CATCH
}

extern "C" int tevo_ctx_launch_count(tevo_ctx* ctx, long long* out, int reset) {
  TEVO_TRY(ctx)
  if (out) *out = tevo::g_launches;
  if (reset) tevo::g_launches = 0;
  TEVO_CATCH
}

// =====================================================================================================
// C ABI: MPS
// =====================================================================================================
extern "C" int tevo_mps_create(tevo_ctx* ctx, int nsites, int d, tevo_mps** out) {
  TEVO_TRY(ctx)
  require(ctx && out, "tevo_mps_create: null argument");
  require(nsites >= 2, "tevo_mps_create: need at least 2 sites");
  require(d >= 2 && d <= 4, "tevo_mps_create: physical dimension must be 2, 3 or 4");
  tevo_mps* p = new tevo_mps();
  p->ctx = ctx;
  p->N = nsites;
  p->d = d;
  p->site.resize(nsites);
  p->chi.assign(nsites + 1, 1);
  p->llim = -1;
  p->rlim = nsites;
  // default content: |0...0> product state
  std::vector<cplx> t(d, ZERO);
  t[0] = ONE;
  for (int j = 0; j < nsites; ++j) {
    p->site[j] = ctx->alloc(d);
    TEVO_CUDA_CHECK(cudaMemcpyAsync(p->site[j].p, t.data(), sizeof(cplx) * d, cudaMemcpyHostToDevice, ctx->stream));
  }
  TEVO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  p->rlim = 1;
  *out = p;
  TEVO_CATCH
}

extern "C" int tevo_mps_free(tevo_mps* psi) {
  if (!psi) return TEVO_OK;
  for (auto& b : psi->site) psi->ctx->release(b);
  for (auto p : psi->lam)
    if (p) cudaFree(p);
  delete psi;
  return TEVO_OK;
}

extern "C" int tevo_mps_copy(tevo_mps* src, tevo_mps** out) {
  TEVO_TRY(src ? src->ctx : nullptr)
  require(src && out, "tevo_mps_copy: null argument");
  tevo_ctx* c = src->ctx;
  tevo_mps* p = new tevo_mps();
  p->ctx = c;
  p->N = src->N;
  p->d = src->d;
  p->chi = src->chi;
  p->llim = src->llim;
  p->rlim = src->rlim;
  p->site.resize(src->N);
  for (int j = 0; j < src->N; ++j) {
    size_t ne = (size_t)src->chi[j] * src->d * src->chi[j + 1];
    p->site[j] = c->alloc(ne);
    TEVO_CUDA_CHECK(cudaMemcpyAsync(p->site[j].p, src->site[j].p, ne * sizeof(cplx), cudaMemcpyDeviceToDevice, c->stream));
  }
  p->bform = src->bform;
  if (!src->lam.empty()) {
    p->lam.assign(src->N + 1, nullptr);
    p->lam_cap.assign(src->N + 1, 0);
    for (int j = 0; j <= src->N; ++j) {
      if (!src->lam[j]) continue;
      const int n = src->chi[j];
      TEVO_CUDA_CHECK(cudaMalloc(&p->lam[j], sizeof(double) * n));
      p->lam_cap[j] = n;
      TEVO_CUDA_CHECK(cudaMemcpyAsync(p->lam[j], src->lam[j], sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
    }
  }
  *out = p;
  TEVO_CATCH
}

extern "C" int tevo_mps_upload_site(tevo_mps* psi, int j, int chil, int chir, const tevo_complex* host) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi && host, "tevo_mps_upload_site: null argument");
  require(j >= 0 && j < psi->N && chil >= 1 && chir >= 1, "tevo_mps_upload_site: bad index or dimension");
  tevo_ctx* c = psi->ctx;
  size_t ne = (size_t)chil * psi->d * chir;
  DevBuf nb = c->alloc(ne);
  TEVO_CUDA_CHECK(cudaMemcpyAsync(nb.p, host, ne * sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
  TEVO_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  set_site(psi, j, nb);
  psi->chi[j] = chil;
  psi->chi[j + 1] = chir;
  psi->bform = false;
  // unknown gauge after an upload
  if (psi->llim >= j) psi->llim = j - 1;
  if (psi->rlim <= j) psi->rlim = j + 1;
  TEVO_CATCH
}

extern "C" int tevo_mps_download_site(tevo_mps* psi, int j, tevo_complex* host, size_t capacity_elems) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi && host, "tevo_mps_download_site: null argument");
  require(j >= 0 && j < psi->N, "tevo_mps_download_site: site out of range");
  size_t ne = (size_t)psi->chi[j] * psi->d * psi->chi[j + 1];
  require(capacity_elems >= ne, "tevo_mps_download_site: buffer too small");
  TEVO_CUDA_CHECK(cudaMemcpyAsync(host, psi->site[j].p, ne * sizeof(cplx), cudaMemcpyDeviceToHost, psi->ctx->stream));
  TEVO_CUDA_CHECK(cudaStreamSynchronize(psi->ctx->stream));
  TEVO_CATCH
}

extern "C" int tevo_mps_site_dims(tevo_mps* psi, int j, int* chil, int* chir) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi && j >= 0 && j < psi->N, "tevo_mps_site_dims: bad argument");
  if (chil) *chil = psi->chi[j];
  if (chir) *chir = psi->chi[j + 1];
  TEVO_CATCH
}

extern "C" int tevo_mps_bonddims(tevo_mps* psi, int* out) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi && out, "tevo_mps_bonddims: null argument");
  for (int b = 0; b < psi->N - 1; ++b) out[b] = psi->chi[b + 1];
  TEVO_CATCH
}

extern "C" int tevo_mps_get_limits(tevo_mps* psi, int* llim, int* rlim) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi, "null");
  if (llim) *llim = psi->llim;
  if (rlim) *rlim = psi->rlim;
  TEVO_CATCH
}

extern "C" int tevo_mps_set_limits(tevo_mps* psi, int llim, int rlim) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi && llim >= -1 && rlim <= psi->N && llim < rlim, "tevo_mps_set_limits: bad limits");
  psi->llim = llim;
  psi->rlim = rlim;
  TEVO_CATCH
}

__global__ void fill_gaussian_kernel(long n, unsigned long long seed, cplx* out) {
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(2 * e + 1);
    auto mix = [](unsigned long long x) {
      x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
      x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
      return x ^ (x >> 31);
    };
    unsigned long long a = mix(z), b = mix(z + 0x9E3779B97F4A7C15ull);
    double u1 = ((a >> 11) + 1.0) * (1.0 / 9007199254740993.0);
    double u2 = (b >> 11) * (1.0 / 9007199254740992.0);
    double r = sqrt(-2.0 * log(u1));
    double s, co;
    sincospi(2.0 * u2, &s, &co);
    out[e] = mk(r * co, r * s);
  }
}

extern "C" int tevo_mps_random(tevo_mps* psi, int chi, unsigned long long seed) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi && chi >= 1, "tevo_mps_random: bad argument");
  tevo_ctx* c = psi->ctx;
  const int N = psi->N, d = psi->d;
  std::vector<int> dims(N + 1);
  for (int j = 0; j <= N; ++j) {
    int e = std::min(j, N - j);
    double cap = std::pow((double)d, (double)std::min(e, 40));
    dims[j] = cap < (double)chi ? (int)cap : chi;
  }
  for (int j = 0; j < N; ++j) {
    size_t ne = (size_t)dims[j] * d * dims[j + 1];
    DevBuf nb = c->alloc(ne);
    long blocks = (long)((ne + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    fill_gaussian_kernel<<<(int)blocks, 256, 0, c->stream>>>((long)ne, seed * 1000003ull + 7919ull * (unsigned long long)j, nb.p);
    TEVO_LAUNCHED();
    set_site(psi, j, nb);
  }
  psi->chi = dims;
  psi->llim = -1;
  psi->rlim = N;
  // right-canonicalise towards site 0; the receiving tensor is renormalised after every hop (the norm of a chain
  // of N(0,1) tensors grows like (d*chi)^(N/2) and would overflow for long chains)
  for (int j = N - 1; j >= 1; --j) {
    hop_left(psi, j);
    psi->rlim = j;
    size_t ne = (size_t)psi->chi[j - 1] * d * psi->chi[j];
    zdotc(c->stream, (long)ne, psi->site[j - 1].p, psi->site[j - 1].p, c->d_scal);
    zscal_rsqrt(c->stream, (long)ne, c->d_scal, psi->site[j - 1].p);
  }
  psi->llim = -1;
  psi->rlim = 1;
  TEVO_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  TEVO_CATCH
}

extern "C" int tevo_mps_orthogonalize(tevo_mps* psi, int j) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi, "null");
  orthogonalize(psi, j);
  TEVO_CATCH
}

static void inner_dev(tevo_mps* a, tevo_mps* b, double* out2_host) {
  tevo_ctx* c = a->ctx;
  require(a->N == b->N && a->d == b->d && a->ctx == b->ctx, "inner: incompatible MPS");
  require(a->chi[0] == 1 && b->chi[0] == 1, "inner: open boundary expected");
  const int d = a->d;
  int cur = WS_E0;
  cplx* E = c->get_ws(cur, 1);
  TEVO_CUDA_CHECK(cudaMemcpyAsync(E, c->ones.p, sizeof(cplx), cudaMemcpyDeviceToDevice, c->stream));
  for (int j = 0; j < a->N; ++j) {
    const int al = a->chi[j], ar = a->chi[j + 1], bl = b->chi[j], br = b->chi[j + 1];
    cplx* T = c->get_ws(WS_T1, (size_t)al * d * br);
    zgemm(c->stream, OP_N, OP_N, al, d * br, bl, ONE, E, al, b->site[j].p, bl, ZERO, T, al);
    int nxt = (cur == WS_E0) ? WS_E1 : WS_E0;
    cplx* En = c->get_ws(nxt, (size_t)ar * br);
    zgemm(c->stream, OP_C, OP_N, ar, br, al * d, ONE, a->site[j].p, al * d, T, al * d, ZERO, En, ar);
    E = En;
    cur = nxt;
  }
  TEVO_CUDA_CHECK(cudaMemcpyAsync(c->h_scal, E, sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
  TEVO_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  out2_host[0] = c->h_scal[0];
  out2_host[1] = c->h_scal[1];
}

extern "C" int tevo_mps_inner(tevo_mps* a, tevo_mps* b, tevo_complex* out) {
  TEVO_TRY(a ? a->ctx : nullptr)
  require(a && b && out, "tevo_mps_inner: null argument");
  double r[2];
  inner_dev(a, b, r);
  out->re = r[0];
  out->im = r[1];
  TEVO_CATCH
}

extern "C" int tevo_mps_reorthogonalize(tevo_mps* psi) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi, "null");
  tevo_ctx* c = psi->ctx;
  psi->llim = -1;
  psi->rlim = psi->N;
  orthogonalize(psi, 0);
  double r[2];
  inner_dev(psi, psi, r);  // psi[1] /= sqrt(inner(psi,psi))  (src/itensor.jl:32)
  require(r[0] > 0.0, "reorthogonalize: state has zero norm");
  size_t ne0 = (size_t)psi->chi[0] * psi->d * psi->chi[1];
  zscal(c->stream, (long)ne0, mk(1.0 / std::sqrt(r[0]), 0.0), psi->site[0].p);
  TEVO_CATCH
}

static void upload_small(tevo_ctx* c, void* dst, const void* src, size_t bytes) {
  TEVO_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
}

extern "C" int tevo_measure_bond(tevo_mps* psi, int b, int nops, const tevo_complex* ops, tevo_complex* out) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi && ops && out && nops >= 1, "tevo_measure_bond: bad argument");
  require(b >= 0 && b < psi->N - 1, "tevo_measure_bond: bond out of range");
  require(4 * nops <= tevo_ctx::NSCAL, "tevo_measure_bond: too many operators");
  require(psi->llim >= b - 1 && psi->rlim <= b + 2, "tevo_measure_bond: orthogonality centre is not on the bond");
  tevo_ctx* c = psi->ctx;
  const int d = psi->d;
  cplx* dops = (cplx*)c->small(sizeof(cplx) * d * d * nops);
  upload_small(c, dops, ops, sizeof(cplx) * d * d * nops);
  cplx* th = form_theta(psi, b);
  expect2(c->stream, th, psi->chi[b], d, psi->chi[b + 2], nops, dops, c->d_scal);
  download_scalars(c, c->d_scal, 4 * nops, (double*)out);
  TEVO_CATCH
}

extern "C" int tevo_mps_expect_site(tevo_mps* psi, int j, const tevo_complex* op, tevo_complex* out) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi && op && out, "tevo_mps_expect_site: null argument");
  require(j >= 0 && j < psi->N, "tevo_mps_expect_site: site out of range");
  tevo_ctx* c = psi->ctx;
  orthogonalize(psi, j);  // src/testutils.jl:56
  const int d = psi->d;
  const int b = (j < psi->N - 1) ? j : j - 1;
  cplx* dops = (cplx*)c->small(sizeof(cplx) * d * d);
  upload_small(c, dops, op, sizeof(cplx) * d * d);
  cplx* th = form_theta(psi, b);
  expect2(c->stream, th, psi->chi[b], d, psi->chi[b + 2], 1, dops, c->d_scal);
  double r[4];
  download_scalars(c, c->d_scal, 4, r);
  const int which = (j == b) ? 0 : 1;
  out->re = r[2 * which];
  out->im = r[2 * which + 1];
  TEVO_CATCH
}

extern "C" int tevo_gate_expect(tevo_mps* psi, int b, const tevo_complex* h, tevo_complex* out) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi && h && out, "tevo_gate_expect: null argument");
  require(b >= 0 && b < psi->N - 1, "tevo_gate_expect: bond out of range");
  tevo_ctx* c = psi->ctx;
  const int d = psi->d;
  orthogonalize(psi, b);  // src/bondgate.jl:74
  cplx* dG = (cplx*)c->small(sizeof(cplx) * d * d * d * d);
  upload_small(c, dG, h, sizeof(cplx) * d * d * d * d);
  const size_t ne = (size_t)psi->chi[b] * d * d * psi->chi[b + 2];
  cplx* th = form_theta(psi, b, WS_THETA);
  cplx* hth = c->get_ws(WS_T1, ne);
  TEVO_CUDA_CHECK(cudaMemcpyAsync(hth, th, ne * sizeof(cplx), cudaMemcpyDeviceToDevice, c->stream));
  gate_apply(c->stream, hth, psi->chi[b], d, psi->chi[b + 2], dG);
  zdotc(c->stream, (long)ne, th, hth, c->d_scal);
  double r[2];
  download_scalars(c, c->d_scal, 2, r);
  out->re = r[0];
  out->im = r[1];
  TEVO_CATCH
}

// =====================================================================================================
// C ABI: TEBD
// =====================================================================================================
extern "C" int tevo_apply_gate(tevo_mps* psi, int b, const tevo_complex* G, const tevo_trunc* trunc, int ortho_left,
                               int normalize, tevo_spec* spec, double* eigs, int eigs_cap) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi && G, "tevo_apply_gate: null argument");
  tevo_ctx* c = psi->ctx;
  const int d4 = psi->d * psi->d * psi->d * psi->d;
  cplx* dG = (cplx*)c->small(sizeof(cplx) * d4);
  upload_small(c, dG, G, sizeof(cplx) * d4);
  TruncInfo ti = apply_gate_dev(psi, b, dG, to_params(trunc), ortho_left != 0, normalize != 0, true, eigs, eigs_cap);
  fill_spec(ti, spec);
  TEVO_CATCH
}

extern "C" int tevo_apply_layer(tevo_mps* psi, int ngates, const int* bonds, const tevo_complex* Gs, int reversed,
                                const tevo_trunc* trunc, int normalize, int keep_ortho_left, tevo_spec* specs,
                                double* eigs, int eigs_stride, int nops, const tevo_complex* ops, tevo_complex* meas) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi && (ngates == 0 || (bonds && Gs)), "tevo_apply_layer: null argument");
  require(nops == 0 || (ops && meas), "tevo_apply_layer: measurement buffers missing");
  tevo_ctx* c = psi->ctx;
  const int d = psi->d, d4 = d * d * d * d;
  const size_t gbytes = sizeof(cplx) * d4 * (size_t)ngates, obytes = sizeof(cplx) * d * d * (size_t)nops;
  const size_t mbytes = sizeof(double) * 4 * (size_t)nops * (size_t)ngates;
  unsigned char* base = (unsigned char*)c->small(gbytes + obytes + mbytes + 64);
  cplx* dG = (cplx*)base;
  cplx* dops = (cplx*)(base + gbytes);
  double* dmeas = (double*)(base + gbytes + obytes);
  if (ngates) upload_small(c, dG, Gs, gbytes);
  if (nops) upload_small(c, dops, ops, obytes);
  const TruncParams tp = to_params(trunc);
  for (int i = 0; i < ngates; ++i) {
    const int g = reversed ? ngates - 1 - i : i;
    const int b = bonds[g];
    const bool left = keep_ortho_left ? true : !reversed;
    TruncInfo ti = apply_gate_dev(psi, b, dG + (size_t)g * d4, tp, left, normalize != 0, keep_ortho_left != 0,
                                  eigs ? eigs + (size_t)g * eigs_stride : nullptr, eigs ? eigs_stride : 0);
    if (specs) fill_spec(ti, &specs[g]);
    if (nops) {
      cplx* th = form_theta(psi, b);
      expect2(c->stream, th, psi->chi[b], d, psi->chi[b + 2], nops, dops, dmeas + (size_t)4 * nops * g);
    }
  }
  if (nops && ngates) {
    TEVO_CUDA_CHECK(cudaMemcpyAsync(meas, dmeas, mbytes, cudaMemcpyDeviceToHost, c->stream));
    TEVO_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  }
  TEVO_CATCH
}

// =====================================================================================================
// MPO / environments / H_eff
// =====================================================================================================
extern "C" int tevo_mpo_create(tevo_ctx* ctx, int nsites, int d, tevo_mpo** out) {
  TEVO_TRY(ctx)
  require(ctx && out && nsites >= 2 && d >= 2 && d <= 4, "tevo_mpo_create: bad argument");
  tevo_mpo* H = new tevo_mpo();
  H->ctx = ctx;
  H->N = nsites;
  H->d = d;
  H->w.assign(nsites + 1, 1);
  H->hW.resize(nsites);
  *out = H;
  TEVO_CATCH
}

extern "C" int tevo_mpo_free(tevo_mpo* H) {
  delete H;
  return TEVO_OK;
}

extern "C" int tevo_mpo_upload_site(tevo_mpo* H, int j, int wl, int wr, const tevo_complex* host) {
  TEVO_TRY(H ? H->ctx : nullptr)
  require(H && host && j >= 0 && j < H->N && wl >= 1 && wr >= 1, "tevo_mpo_upload_site: bad argument");
  const size_t ne = (size_t)wl * H->d * H->d * wr;
  H->hW[j].assign((const std::complex<double>*)host, (const std::complex<double>*)host + ne);
  H->w[j] = wl;
  H->w[j + 1] = wr;
  TEVO_CATCH
}

static inline std::complex<double> Wat(const tevo_mpo* H, int j, int wl_i, int so, int si, int wr_i) {
  const int wl = H->w[j], d = H->d;
  return H->hW[j][wl_i + (size_t)wl * (so + (size_t)d * (si + (size_t)d * wr_i))];
}

// uploads a Q x P operator block and Q offsets; returns device pointers
struct MidOp {
  cplx* M;
  long* oq;
};
static MidOp upload_mid(tevo_ctx* c, const std::vector<std::complex<double>>& M, const std::vector<long>& oq) {
  size_t mb = M.size() * sizeof(cplx), ob = oq.size() * sizeof(long);
  size_t mba = (mb + 63) & ~(size_t)63;
  unsigned char* base = (unsigned char*)c->small(mba + ob + 64);
  upload_small(c, base, M.data(), mb);
  upload_small(c, base + mba, oq.data(), ob);
  MidOp r;
  r.M = (cplx*)base;
  r.oq = (long*)(base + mba);
  return r;
}

// L_new(b', y, b) from L(a', w, a), ket site Ak (cl, d, cr), W_j, bra site Ab (bl, d, br)
static DevBuf env_left_update(tevo_ctx* c, const tevo_mpo* H, int j, const cplx* L, int bl, int wl, int cl,
                              const cplx* Ak, int cr, const cplx* Ab, int br) {
  const int d = H->d, wr = H->w[j + 1];
  require(wl == H->w[j], "env_left_update: MPO bond mismatch");
  cplx* T1 = c->get_ws(WS_T1, (size_t)bl * wl * d * cr);
  zgemm(c->stream, OP_N, OP_N, bl * wl, d * cr, cl, ONE, L, bl * wl, Ak, cl, ZERO, T1, bl * wl);
  // mid: (a'; p=(w,s); r) -> (a'; q=(s',y); r)
  const int P = wl * d, Q = d * wr;
  std::vector<std::complex<double>> M((size_t)P * Q);
  std::vector<long> oq(Q);
  for (int sp = 0; sp < d; ++sp)
    for (int y = 0; y < wr; ++y) {
      const int q = sp + d * y;
      oq[q] = (long)bl * (sp + (long)d * y);
      for (int w = 0; w < wl; ++w)
        for (int s = 0; s < d; ++s) M[(size_t)q * P + (w + wl * s)] = Wat(H, j, w, sp, s, y);
    }
  MidOp mo = upload_mid(c, M, oq);
  cplx* T2 = c->get_ws(WS_T2, (size_t)bl * d * wr * cr);
  mid_contract(c->stream, bl, P, Q, cr, mo.M, mo.oq, (long)bl * d * wr, T1, T2);
  DevBuf out = c->alloc((size_t)br * wr * cr);
  zgemm(c->stream, OP_C, OP_N, br, wr * cr, bl * d, ONE, Ab, bl * d, T2, bl * d, ZERO, out.p, br);
  return out;
}

// R_new(a, x, a') from R(c, y, c'), ket site Ak (cl, d, cr), W_j, bra site Ab (bl, d, br)
static DevBuf env_right_update(tevo_ctx* c, const tevo_mpo* H, int j, const cplx* R, int cr, int wr, int br,
                               const cplx* Ak, int cl, const cplx* Ab, int bl) {
  const int d = H->d, wl = H->w[j];
  require(wr == H->w[j + 1], "env_right_update: MPO bond mismatch");
  cplx* T1 = c->get_ws(WS_T1, (size_t)cl * d * wr * br);
  zgemm(c->stream, OP_N, OP_N, cl * d, wr * br, cr, ONE, Ak, cl * d, R, cr, ZERO, T1, cl * d);
  // mid: (a; p=(s,y); c') -> (a; q=(x,s'); c')
  const int P = d * wr, Q = wl * d;
  std::vector<std::complex<double>> M((size_t)P * Q);
  std::vector<long> oq(Q);
  for (int x = 0; x < wl; ++x)
    for (int sp = 0; sp < d; ++sp) {
      const int q = x + wl * sp;
      oq[q] = (long)cl * (x + (long)wl * sp);
      for (int s = 0; s < d; ++s)
        for (int y = 0; y < wr; ++y) M[(size_t)q * P + (s + d * y)] = Wat(H, j, x, sp, s, y);
    }
  MidOp mo = upload_mid(c, M, oq);
  cplx* T2 = c->get_ws(WS_T2, (size_t)cl * wl * d * br);
  mid_contract(c->stream, cl, P, Q, br, mo.M, mo.oq, (long)cl * wl * d, T1, T2);
  DevBuf out = c->alloc((size_t)cl * wl * bl);
  zgemm(c->stream, OP_N, OP_C, cl * wl, bl, d * br, ONE, T2, cl * wl, Ab, bl, ZERO, out.p, cl * wl);
  return out;
}

extern "C" int tevo_mps_inner_mpo(tevo_mps* a, tevo_mpo* H, tevo_mps* b, tevo_complex* out) {
  TEVO_TRY(a ? a->ctx : nullptr)
  require(a && b && H && out, "tevo_mps_inner_mpo: null argument");
  require(a->N == b->N && a->N == H->N && a->d == H->d && b->d == H->d, "tevo_mps_inner_mpo: size mismatch");
  tevo_ctx* c = a->ctx;
  DevBuf E = c->alloc(1);
  TEVO_CUDA_CHECK(cudaMemcpyAsync(E.p, c->ones.p, sizeof(cplx), cudaMemcpyDeviceToDevice, c->stream));
  for (int j = 0; j < a->N; ++j) {
    DevBuf En = env_left_update(c, H, j, E.p, a->chi[j], H->w[j], b->chi[j], b->site[j].p, b->chi[j + 1],
                                a->site[j].p, a->chi[j + 1]);
    c->release(E);
    E = En;
  }
  TEVO_CUDA_CHECK(cudaMemcpyAsync(c->h_scal, E.p, sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
  TEVO_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  out->re = c->h_scal[0];
  out->im = c->h_scal[1];
  c->release(E);
  TEVO_CATCH
}

extern "C" int tevo_env_create(tevo_mpo* H, tevo_env** out) {
  TEVO_TRY(H ? H->ctx : nullptr)
  require(H && out, "tevo_env_create: null argument");
  for (int j = 0; j < H->N; ++j) require(!H->hW[j].empty(), "tevo_env_create: MPO site not uploaded");
  require(H->w[0] == 1 && H->w[H->N] == 1, "tevo_env_create: MPO boundary bonds must have dimension 1");
  tevo_env* e = new tevo_env();
  e->H = H;
  e->ctx = H->ctx;
  e->N = H->N;
  e->LR.resize(H->N);
  e->dim_bra.assign(H->N, 0);
  e->dim_w.assign(H->N, 0);
  e->dim_ket.assign(H->N, 0);
  e->lpos = -1;
  e->rpos = H->N;
  *out = e;
  TEVO_CATCH
}

extern "C" int tevo_env_free(tevo_env* env) {
  if (!env) return TEVO_OK;
  for (auto& b : env->LR) env->ctx->release(b);
  delete env;
  return TEVO_OK;
}

static void env_position(tevo_env* e, tevo_mps* psi, int pos, int nsite) {
  tevo_ctx* c = e->ctx;
  require(psi->N == e->N && psi->d == e->H->d, "position!: MPS/MPO size mismatch");
  require(nsite == 1 || nsite == 2, "position!: nsite must be 1 or 2");
  require(pos >= 0 && pos + nsite - 1 < e->N, "position!: position out of range");
  while (e->lpos < pos - 1) {
    const int j = e->lpos + 1;
    const cplx* L = (j == 0) ? c->ones.p : e->LR[j - 1].p;
    DevBuf n = env_left_update(c, e->H, j, L, psi->chi[j], e->H->w[j], psi->chi[j], psi->site[j].p, psi->chi[j + 1],
                               psi->site[j].p, psi->chi[j + 1]);
    c->release(e->LR[j]);
    e->LR[j] = n;
    e->dim_bra[j] = psi->chi[j + 1];
    e->dim_w[j] = e->H->w[j + 1];
    e->dim_ket[j] = psi->chi[j + 1];
    e->lpos = j;
  }
  e->lpos = pos - 1;
  const int end = pos + nsite - 1;
  while (e->rpos > end + 1) {
    const int j = e->rpos - 1;
    const cplx* R = (j == e->N - 1) ? c->ones.p : e->LR[j + 1].p;
    DevBuf n = env_right_update(c, e->H, j, R, psi->chi[j + 1], e->H->w[j + 1], psi->chi[j + 1], psi->site[j].p,
                                psi->chi[j], psi->site[j].p, psi->chi[j]);
    c->release(e->LR[j]);
    e->LR[j] = n;
    e->dim_ket[j] = psi->chi[j];
    e->dim_w[j] = e->H->w[j];
    e->dim_bra[j] = psi->chi[j];
    e->rpos = j;
  }
  e->rpos = end + 1;
}

extern "C" int tevo_env_position(tevo_env* env, tevo_mps* psi, int pos, int nsite) {
  TEVO_TRY(env ? env->ctx : nullptr)
  require(env && psi, "tevo_env_position: null argument");
  env_position(env, psi, pos, nsite);
  TEVO_CATCH
}

static double g_heff_flops = 0.0;  // algorithmic flops of all H_eff applications (two GEMM legs + operator block)

// H_eff application descriptor at a position (operator block uploaded once, reused by every Krylov matvec)
struct Heff {
  tevo_ctx* c;
  const cplx *L, *R;
  int cl, wl, cr, wr, d, nsite;
  cplx* dM;
  long* doq;
  int P, Q;
  long n;  // vector length
  void apply(const cplx* v, cplx* out) const {
    ProfScope ps(c->stream, P_HEFF);
    {
      const double dd0 = (nsite == 2) ? (double)d * d : (double)d;
      g_heff_flops += 8.0 * ((double)cl * wl * cl * dd0 * cr + (double)cl * cr * (double)P * Q +
                             (double)cl * dd0 * cr * wr * cr);
    }
    const int dd = (nsite == 2) ? d * d : d;
    cplx* T1 = c->get_ws(WS_T1, (size_t)cl * wl * dd * cr);
    zgemm(c->stream, OP_N, OP_N, cl * wl, dd * cr, cl, ONE, L, cl * wl, v, cl, ZERO, T1, cl * wl);
    cplx* T2 = c->get_ws(WS_T2, (size_t)cl * dd * cr * wr);
    mid_contract(c->stream, cl, P, Q, cr, dM, doq, (long)cl * dd, T1, T2);
    zgemm(c->stream, OP_N, OP_N, cl * dd, cr, cr * wr, ONE, T2, cl * dd, R, cr * wr, ZERO, out, cl * dd);
  }
};

static Heff make_heff(tevo_env* e, tevo_mps* psi, int pos, int nsite) {
  tevo_ctx* c = e->ctx;
  const tevo_mpo* H = e->H;
  const int d = H->d;
  Heff h;
  h.c = c;
  h.d = d;
  h.nsite = nsite;
  h.cl = psi->chi[pos];
  h.cr = psi->chi[pos + nsite];
  h.wl = H->w[pos];
  h.wr = H->w[pos + nsite];
  require(e->lpos == pos - 1 && e->rpos == pos + nsite, "H_eff: environment is not positioned here");
  h.L = (pos == 0) ? c->ones.p : e->LR[pos - 1].p;
  h.R = (pos + nsite == e->N) ? c->ones.p : e->LR[pos + nsite].p;
  if (pos > 0) require(e->dim_ket[pos - 1] == h.cl && e->dim_bra[pos - 1] == h.cl, "H_eff: stale left environment");
  if (pos + nsite < e->N)
    require(e->dim_ket[pos + nsite] == h.cr && e->dim_bra[pos + nsite] == h.cr, "H_eff: stale right environment");
  std::vector<std::complex<double>> M;
  std::vector<long> oq;
  if (nsite == 2) {
    const int wm = H->w[pos + 1];
    h.P = h.wl * d * d;
    h.Q = d * d * h.wr;
    M.assign((size_t)h.P * h.Q, 0.0);
    oq.resize(h.Q);
    for (int s1p = 0; s1p < d; ++s1p)
      for (int s2p = 0; s2p < d; ++s2p)
        for (int y = 0; y < h.wr; ++y) {
          const int q = s1p + d * (s2p + d * y);
          oq[q] = (long)h.cl * (s1p + (long)d * s2p) + (long)h.cl * d * d * h.cr * y;
          for (int w = 0; w < h.wl; ++w)
            for (int s1 = 0; s1 < d; ++s1)
              for (int s2 = 0; s2 < d; ++s2) {
                std::complex<double> acc = 0.0;
                for (int x = 0; x < wm; ++x) acc += Wat(H, pos, w, s1p, s1, x) * Wat(H, pos + 1, x, s2p, s2, y);
                M[(size_t)q * h.P + (w + h.wl * (s1 + d * s2))] = acc;
              }
        }
    h.n = (long)h.cl * d * d * h.cr;
  } else {
    h.P = h.wl * d;
    h.Q = d * h.wr;
    M.assign((size_t)h.P * h.Q, 0.0);
    oq.resize(h.Q);
    for (int sp = 0; sp < d; ++sp)
      for (int y = 0; y < h.wr; ++y) {
        const int q = sp + d * y;
        oq[q] = (long)h.cl * sp + (long)h.cl * d * h.cr * y;
        for (int w = 0; w < h.wl; ++w)
          for (int s = 0; s < d; ++s) M[(size_t)q * h.P + (w + h.wl * s)] = Wat(H, pos, w, sp, s, y);
      }
    h.n = (long)h.cl * d * h.cr;
  }
  // device copy in the context's dedicated block, sized from P and Q (wl*wr*d^4 complex for nsite = 2: beyond 64 KB
  // already for d = 2 with w >= 16), so that no other staging use of small() can alias or overrun it
  size_t mb = M.size() * sizeof(cplx);
  size_t mba = (mb + 63) & ~(size_t)63;
  cplx* dev_block = (cplx*)c->heff_block(mba + oq.size() * sizeof(long));
  upload_small(c, dev_block, M.data(), mb);
  upload_small(c, (unsigned char*)dev_block + mba, oq.data(), oq.size() * sizeof(long));
  h.dM = dev_block;
  h.doq = (long*)((unsigned char*)dev_block + mba);
  return h;
}

extern "C" int tevo_env_apply_host(tevo_env* env, tevo_mps* psi, int pos, int nsite, const tevo_complex* v,
                                   tevo_complex* out) {
  TEVO_TRY(env ? env->ctx : nullptr)
  require(env && psi && v && out, "tevo_env_apply_host: null argument");
  tevo_ctx* c = env->ctx;
  env_position(env, psi, pos, nsite);
  Heff h = make_heff(env, psi, pos, nsite);
  cplx* dv = c->get_ws(WS_W, (size_t)2 * h.n);
  TEVO_CUDA_CHECK(cudaMemcpyAsync(dv, v, h.n * sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
  h.apply(dv, dv + h.n);
  TEVO_CUDA_CHECK(cudaMemcpyAsync(out, dv + h.n, h.n * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
  TEVO_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  TEVO_CATCH
}

// =====================================================================================================
// Krylov exponentiate (KrylovKit.exponentiate semantics; src/tdvp.jl:61,81)
// =====================================================================================================
// real symmetric tridiagonal eigen-decomposition (implicit QL), host; k <= ~100
static void tridiag_eigh(int k, const std::vector<double>& alpha, const std::vector<double>& beta,
                         std::vector<double>& D, std::vector<double>& U /* k x k row-major: U[i*k+j] = comp i of vec j */) {
  D.assign(alpha.begin(), alpha.begin() + k);
  std::vector<double> e(k, 0.0);
  for (int i = 0; i + 1 < k; ++i) e[i] = beta[i];
  U.assign((size_t)k * k, 0.0);
  for (int i = 0; i < k; ++i) U[(size_t)i * k + i] = 1.0;
  for (int l = 0; l < k; ++l) {
    int iter = 0, m;
    do {
      for (m = l; m < k - 1; ++m) {
        double dd = std::fabs(D[m]) + std::fabs(D[m + 1]);
        if (std::fabs(e[m]) <= 2.3e-16 * dd) break;
      }
      if (m != l) {
        if (iter++ == 200) throw TevoError(TEVO_EINTERNAL, "tridiagonal QL did not converge");
        double g = (D[l + 1] - D[l]) / (2.0 * e[l]);
        double r = std::hypot(g, 1.0);
        g = D[m] - D[l] + e[l] / (g + (g >= 0 ? std::fabs(r) : -std::fabs(r)));
        double s = 1.0, cc = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; --i) {
          double f = s * e[i], b = cc * e[i];
          r = std::hypot(f, g);
          e[i + 1] = r;
          if (r == 0.0) {
            D[i + 1] -= p;
            e[m] = 0.0;
            break;
          }
          s = f / r;
          cc = g / r;
          g = D[i + 1] - p;
          r = (D[i] - g) * s + 2.0 * cc * b;
          p = s * r;
          D[i + 1] = g + p;
          g = cc * r - b;
          for (int kk = 0; kk < k; ++kk) {
            double f2 = U[(size_t)kk * k + i + 1];
            U[(size_t)kk * k + i + 1] = s * U[(size_t)kk * k + i] + cc * f2;
            U[(size_t)kk * k + i] = cc * U[(size_t)kk * k + i] - s * f2;
          }
        }
        if (r == 0.0 && i >= l) continue;
        D[l] -= p;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
}

// first column of exp(x * H) for a small dense k x k complex matrix (row-major): scaling and squaring with a Taylor
// series; used by the Arnoldi (hermitian=false) branch of exponentiate
static std::vector<std::complex<double>> small_expm_col0(int k, const std::vector<std::complex<double>>& H,
                                                         std::complex<double> x) {
  typedef std::complex<double> cd;
  std::vector<cd> A((size_t)k * k);
  double nrm = 0.0;
  for (int i = 0; i < k; ++i) {
    double row = 0.0;
    for (int j = 0; j < k; ++j) {
      A[(size_t)i * k + j] = x * H[(size_t)i * k + j];
      row += std::abs(A[(size_t)i * k + j]);
    }
    nrm = std::max(nrm, row);
  }
  int s = 0;
  while (nrm > 0.25 && s < 60) {
    nrm *= 0.5;
    ++s;
  }
  const double sc = std::ldexp(1.0, -s);
  for (auto& v : A) v *= sc;
  std::vector<cd> E((size_t)k * k, cd(0.0)), T((size_t)k * k, cd(0.0)), W((size_t)k * k);
  for (int i = 0; i < k; ++i) E[(size_t)i * k + i] = T[(size_t)i * k + i] = 1.0;
  auto matmul = [&](const std::vector<cd>& X, const std::vector<cd>& Y, std::vector<cd>& Z) {
    for (int i = 0; i < k; ++i)
      for (int j = 0; j < k; ++j) {
        cd acc = 0.0;
        for (int l = 0; l < k; ++l) acc += X[(size_t)i * k + l] * Y[(size_t)l * k + j];
        Z[(size_t)i * k + j] = acc;
      }
  };
  for (int j = 1; j <= 20; ++j) {
    matmul(T, A, W);
    for (size_t i = 0; i < W.size(); ++i) {
      T[i] = W[i] / (double)j;
      E[i] += T[i];
    }
  }
  for (int q = 0; q < s; ++q) {
    matmul(E, E, W);
    E.swap(W);
  }
  std::vector<cd> col(k);
  for (int i = 0; i < k; ++i) col[i] = E[(size_t)i * k];
  return col;
}

struct KrylovParams {
  bool hermitian;
  double tol;
  int krylovdim, maxiter;
};

// w <- exp(t * Heff) w  on the device.  Returns info; throws on invalid input.
static tevo_krylov_info krylov_expv(const Heff& h, std::complex<double> t, cplx* w, const KrylovParams& kp) {
  tevo_ctx* c = h.c;
  tevo_krylov_info info = {0, 0, 0};
  const long n = h.n;
  const double tau_total = std::abs(t);
  if (tau_total == 0.0) {
    info.converged = 1;
    return info;
  }
  const std::complex<double> sgn = t / tau_total;
  int kd = kp.krylovdim;
  if ((long)kd > n) kd = (int)n;
  require(kd >= 1 && 6 * kd + 8 <= tevo_ctx::NSCAL, "exponentiate: krylovdim out of range");
  require(kp.tol > 0.0 && std::isfinite(kp.tol), "exponentiate: tol must be positive and finite");
  require(kp.maxiter >= 1, "exponentiate: maxiter must be at least 1");
  cplx* V = c->get_ws(WS_KRYLOV, (size_t)(kd + 1) * n);
  double* dh1 = c->d_scal;                 // 2*(kd+1)
  double* dh2 = c->d_scal + 2 * (kd + 1);  // 2*(kd+1)
  double* dnb = c->d_scal + 4 * (kd + 1);  // 2
  double* dy = c->d_scal + 4 * (kd + 1) + 2;
  std::vector<double> hs(4 * (kd + 1) + 2);
  double tau_done = 0.0;
  const double eta = kp.tol;
  std::vector<double> alpha(kd), beta(kd), D, U;
  const bool herm = kp.hermitian;
  std::vector<std::complex<double>> Hess;  // Arnoldi: (kd+1) x kd upper Hessenberg, row-major with row stride kd
  if (!herm) Hess.assign((size_t)(kd + 1) * kd, 0.0);
  auto hess_block = [&](int kk) {
    std::vector<std::complex<double>> Hk((size_t)kk * kk);
    for (int i = 0; i < kk; ++i)
      for (int j = 0; j < kk; ++j) Hk[(size_t)i * kk + j] = Hess[(size_t)i * kd + j];
    return Hk;
  };
  while (true) {
    info.numiter += 1;
    // beta0 = |w|
    zdotc(c->stream, n, w, w, dnb);
    double nb2[2];
    download_scalars(c, dnb, 2, nb2);
    const double beta0 = std::sqrt(nb2[0]);
    if (!std::isfinite(beta0)) throw TevoError(TEVO_EINVAL, "exponentiate: the input vector is not finite");
    if (beta0 == 0.0) {  // exp(tA) 0 = 0
      info.converged = 1;
      return info;
    }
    copy_scale(c->stream, n, 1.0 / beta0, w, V);
    double dtau = tau_total - tau_done;
    int k = 0;
    bool done_sub = false;
    double nb = 0.0;
    if (!herm) std::fill(Hess.begin(), Hess.end(), std::complex<double>(0.0));
    auto estimate = [&](int kk, double dt, double nbv) {
      std::complex<double> e1 = 0.0, e2 = 0.0;
      if (!herm) {
        const auto Hk = hess_block(kk);
        e1 = small_expm_col0(kk, Hk, sgn * (dt / 2))[kk - 1];
        e2 = small_expm_col0(kk, Hk, sgn * dt)[kk - 1];
        return nbv * (2.0 * std::abs(e1) / 3.0 + std::abs(e2) / 6.0);
      }
      for (int j = 0; j < kk; ++j) {
        const double ukj = U[(size_t)(kk - 1) * kk + j], u1j = U[j];
        e1 += ukj * std::exp(sgn * (dt / 2) * D[j]) * u1j;
        e2 += ukj * std::exp(sgn * dt * D[j]) * u1j;
      }
      return nbv * (2.0 * std::abs(e1) / 3.0 + std::abs(e2) / 6.0);
    };
    while (k < kd) {
      cplx* r = V + (size_t)(k + 1) * n;
      h.apply(V + (size_t)k * n, r);
      info.numops += 1;
      multi_dotc(c->stream, n, k + 1, V, n, r, dh1);
      multi_axpy_sub(c->stream, n, k + 1, V, n, dh1, r);
      multi_dotc(c->stream, n, k + 1, V, n, r, dh2);
      multi_axpy_sub(c->stream, n, k + 1, V, n, dh2, r);
      zdotc(c->stream, n, r, r, dnb);
      download_scalars(c, c->d_scal, 4 * (kd + 1) + 2, hs.data());
      alpha[k] = hs[2 * k] + hs[2 * (kd + 1) + 2 * k];
      nb = std::sqrt(std::max(0.0, hs[4 * (kd + 1)]));
      beta[k] = nb;
      if (!herm) {
        for (int i = 0; i <= k; ++i)
          Hess[(size_t)i * kd + k] = std::complex<double>(hs[2 * i] + hs[2 * (kd + 1) + 2 * i],
                                                          hs[2 * i + 1] + hs[2 * (kd + 1) + 2 * i + 1]);
        Hess[(size_t)(k + 1) * kd + k] = nb;
      }
      k += 1;
      if (herm) tridiag_eigh(k, alpha, beta, D, U);
      const double eps = estimate(k, dtau, nb);
      if (eps < eta || (long)k == n) {
        done_sub = true;
        break;
      }
      if (nb > 0.0) zscal(c->stream, n, mk(1.0 / nb, 0.0), r);
    }
    if (!done_sub) {
      // largest sub-step the full Krylov space supports (KrylovKit: dtau *= (eta/eps)^(1/krylovdim)); bounded, and a
      // non-finite estimate is an error instead of an endless loop
      for (int it = 0;; ++it) {
        const double eps = estimate(k, dtau, nb);
        if (!std::isfinite(eps)) throw TevoError(TEVO_ENOTCONVERGED, "exponentiate: non-finite error estimate");
        if (eps < eta || info.numiter == kp.maxiter) break;
        if (it >= 200 || !(dtau > 1e-300)) throw TevoError(TEVO_ENOTCONVERGED, "exponentiate did not converge");
        dtau = dtau * std::pow(eta / eps, 1.0 / kd);
      }
    }
    // y = U exp(sgn*dtau*D) U^T e1
    std::vector<double> y(2 * k, 0.0);
    if (!herm) {
      const auto col = small_expm_col0(k, hess_block(k), sgn * dtau);
      for (int i = 0; i < k; ++i) {
        y[2 * i] = col[i].real();
        y[2 * i + 1] = col[i].imag();
      }
    } else
    for (int i = 0; i < k; ++i) {
      std::complex<double> acc = 0.0;
      for (int j = 0; j < k; ++j) acc += U[(size_t)i * k + j] * std::exp(sgn * dtau * D[j]) * U[j];
      y[2 * i] = acc.real();
      y[2 * i + 1] = acc.imag();
    }
    upload_small(c, dy, y.data(), sizeof(double) * 2 * k);
    lincomb(c->stream, n, k, V, n, dy, beta0, w);
    tau_done += dtau;
    if (tau_done >= tau_total * (1 - 1e-15)) {
      info.converged = 1;
      break;
    }
    if (info.numiter >= kp.maxiter) break;
  }
  return info;
}

static KrylovParams to_kp(const tevo_krylov* k) {
  KrylovParams p;
  p.hermitian = k ? (k->hermitian != 0) : true;
  p.tol = k ? k->tol : 1e-14;
  p.krylovdim = k ? k->krylovdim : 30;
  p.maxiter = k ? k->maxiter : 100;
  return p;
}

extern "C" int tevo_tdvp_twosite(tevo_env* env, tevo_mps* psi, int b, tevo_complex t, const tevo_krylov* kp,
                                 const tevo_trunc* trunc, int ortho_left, int normalize, tevo_spec* spec, double* eigs,
                                 int eigs_cap, tevo_krylov_info* info) {
  TEVO_TRY(env ? env->ctx : nullptr)
  require(env && psi, "tevo_tdvp_twosite: null argument");
  require(b >= 0 && b < psi->N - 1, "tevo_tdvp_twosite: bond out of range");
  tevo_ctx* c = env->ctx;
  env_position(env, psi, b, 2);  // src/tdvp.jl:58-59
  Heff h = make_heff(env, psi, b, 2);
  cplx* w = c->get_ws(WS_W, (size_t)h.n);
  {
    const int d = psi->d, m = psi->chi[b] * d, k = psi->chi[b + 1], n = d * psi->chi[b + 2];
    zgemm(c->stream, OP_N, OP_N, m, n, k, ONE, psi->site[b].p, m, psi->site[b + 1].p, k, ZERO, w, m);  // :60
  }
  tevo_krylov_info ki = krylov_expv(h, std::complex<double>(t.re, t.im), w, to_kp(kp));  // :61
  if (info) *info = ki;
  if (!ki.converged) throw TevoError(TEVO_ENOTCONVERGED, "exponentiate did not converge");  // :63
  TruncInfo ti = replace_bond(psi, b, w, ortho_left != 0, to_params(trunc), normalize != 0, eigs, eigs_cap);  // :64
  fill_spec(ti, spec);
  TEVO_CATCH
}

extern "C" int tevo_tdvp_onesite(tevo_env* env, tevo_mps* psi, int i, tevo_complex t, const tevo_krylov* kp,
                                 tevo_krylov_info* info) {
  TEVO_TRY(env ? env->ctx : nullptr)
  require(env && psi, "tevo_tdvp_onesite: null argument");
  require(i >= 0 && i < psi->N, "tevo_tdvp_onesite: site out of range");
  tevo_ctx* c = env->ctx;
  env_position(env, psi, i, 1);  // src/tdvp.jl:79-80
  Heff h = make_heff(env, psi, i, 1);
  tevo_krylov_info ki = krylov_expv(h, std::complex<double>(t.re, t.im), psi->site[i].p, to_kp(kp));  // :81
  if (info) *info = ki;
  if (!ki.converged) throw TevoError(TEVO_ENOTCONVERGED, "exponentiate did not converge");  // :83
  TEVO_CATCH
}

extern "C" int tevo_mps_normalize_site(tevo_mps* psi, int i) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi && i >= 0 && i < psi->N, "tevo_mps_normalize_site: bad argument");
  tevo_ctx* c = psi->ctx;
  size_t ne = (size_t)psi->chi[i] * psi->d * psi->chi[i + 1];
  zdotc(c->stream, (long)ne, psi->site[i].p, psi->site[i].p, c->d_scal);
  zscal_rsqrt(c->stream, (long)ne, c->d_scal, psi->site[i].p);
  TEVO_CATCH
}

// =====================================================================================================
// test hooks
// =====================================================================================================
extern "C" int tevo_test_zgemm(tevo_ctx* ctx, int ta, int tb, int M, int N, int K, tevo_complex alpha,
                               const tevo_complex* A, int lda, const tevo_complex* B, int ldb, tevo_complex beta,
                               tevo_complex* C, int ldc) {
  TEVO_TRY(ctx)
  require(ctx && A && B && C, "tevo_test_zgemm: null argument");
  const size_t na = (size_t)lda * (ta == 0 ? K : M), nb = (size_t)ldb * (tb == 0 ? N : K), nc = (size_t)ldc * N;
  DevBuf dA = ctx->alloc(na), dB = ctx->alloc(nb), dC = ctx->alloc(nc);
  TEVO_CUDA_CHECK(cudaMemcpyAsync(dA.p, A, na * sizeof(cplx), cudaMemcpyHostToDevice, ctx->stream));
  TEVO_CUDA_CHECK(cudaMemcpyAsync(dB.p, B, nb * sizeof(cplx), cudaMemcpyHostToDevice, ctx->stream));
  TEVO_CUDA_CHECK(cudaMemcpyAsync(dC.p, C, nc * sizeof(cplx), cudaMemcpyHostToDevice, ctx->stream));
  zgemm(ctx->stream, (Op)ta, (Op)tb, M, N, K, mk(alpha.re, alpha.im), dA.p, lda, dB.p, ldb, mk(beta.re, beta.im), dC.p,
        ldc);
  TEVO_CUDA_CHECK(cudaMemcpyAsync(C, dC.p, nc * sizeof(cplx), cudaMemcpyDeviceToHost, ctx->stream));
  TEVO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ctx->release(dA);
  ctx->release(dB);
  ctx->release(dC);
  TEVO_CATCH
}

/* device-resident timing of the DMMA ZGEMM (no host copies in the timed region): reps launches of
 * C = op(A) op(B) (+ C if accumulate) between two CUDA events on the library stream; kind 0 = zgemm, 1 = zgemm_herm
 * (M x M result, lower tiles + mirror).  ms_out = average milliseconds per launch. */
extern "C" int tevo_test_zgemm_time(tevo_ctx* ctx, int kind, int ta, int tb, int M, int N, int K, int accumulate, int reps,
                                    double* ms_out) {
  TEVO_TRY(ctx)
  require(ctx && ms_out && reps > 0 && M > 0 && N > 0 && K > 0, "tevo_test_zgemm_time: bad argument");
  const long lda = (ta == 0) ? M : K, ldb = (tb == 0) ? K : N;
  const size_t na = (size_t)lda * (ta == 0 ? K : M), nb = (size_t)ldb * (tb == 0 ? N : K), nc = (size_t)M * N;
  DevBuf dA = ctx->alloc(na), dB = ctx->alloc(nb), dC = ctx->alloc(nc);
  // finite, non-trivial contents: every byte 0x3f gives doubles near 4.7e-4
  TEVO_CUDA_CHECK(cudaMemsetAsync(dA.p, 0x3f, na * sizeof(cplx), ctx->stream));
  TEVO_CUDA_CHECK(cudaMemsetAsync(dB.p, 0x3f, nb * sizeof(cplx), ctx->stream));
  TEVO_CUDA_CHECK(cudaMemsetAsync(dC.p, 0, nc * sizeof(cplx), ctx->stream));
  const cplx beta = accumulate ? mk(1, 0) : mk(0, 0);
  auto run = [&]() {
    if (kind == 1)
      zgemm_herm(ctx->stream, (Op)ta, (Op)tb, M, K, ONE, dA.p, lda, dB.p, ldb, beta, dC.p, M, true);
    else
      zgemm(ctx->stream, (Op)ta, (Op)tb, M, N, K, ONE, dA.p, lda, dB.p, ldb, beta, dC.p, M);
  };
  for (int i = 0; i < 3; ++i) run();
  cudaEvent_t e0, e1;
  TEVO_CUDA_CHECK(cudaEventCreate(&e0));
  TEVO_CUDA_CHECK(cudaEventCreate(&e1));
  TEVO_CUDA_CHECK(cudaEventRecord(e0, ctx->stream));
  for (int i = 0; i < reps; ++i) run();
  TEVO_CUDA_CHECK(cudaEventRecord(e1, ctx->stream));
  TEVO_CUDA_CHECK(cudaEventSynchronize(e1));
  float ms = 0.f;
  TEVO_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_out = (double)ms / reps;
  ctx->release(dA);
  ctx->release(dB);
  ctx->release(dC);
  TEVO_CATCH
}

extern "C" int tevo_test_split(tevo_ctx* ctx, int m, int n, const tevo_complex* theta, const tevo_trunc* trunc,
                               int ortho_left, int normalize, tevo_spec* spec, double* eigs, int eigs_cap,
                               tevo_complex* L, tevo_complex* R, int kcap) {
  TEVO_TRY(ctx)
  require(ctx && theta && L && R, "tevo_test_split: null argument");
  DevBuf dT = ctx->alloc((size_t)m * n);
  TEVO_CUDA_CHECK(cudaMemcpyAsync(dT.p, theta, (size_t)m * n * sizeof(cplx), cudaMemcpyHostToDevice, ctx->stream));
  SplitResult r = ctx->split.run(ctx->stream, dT.p, m, n, ortho_left != 0, to_params(trunc), normalize != 0, eigs,
                                 eigs_cap, [&](size_t ne) { return ctx->alloc(ne); });
  fill_spec(r.info, spec);
  require(r.k <= kcap, "tevo_test_split: output capacity too small");
  TEVO_CUDA_CHECK(cudaMemcpyAsync(L, r.L.p, (size_t)m * r.k * sizeof(cplx), cudaMemcpyDeviceToHost, ctx->stream));
  TEVO_CUDA_CHECK(cudaMemcpyAsync(R, r.R.p, (size_t)r.k * n * sizeof(cplx), cudaMemcpyDeviceToHost, ctx->stream));
  TEVO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ctx->release(r.L);
  ctx->release(r.R);
  ctx->release(dT);
  TEVO_CATCH
}

extern "C" int tevo_test_heig_ctas(tevo_ctx* ctx, int n, const tevo_complex* A, double* evals_desc, tevo_complex* U,
                                   int max_ctas) {
  if (!ctx || max_ctas < 0) return TEVO_EINVAL;
  tevo::Heig& h = ctx->split.heig();
  const int saved = h.max_ctas();
  h.set_max_ctas(max_ctas);
  const int rc = tevo_test_heig(ctx, n, A, evals_desc, U);
  h.set_max_ctas(saved);
  return rc;
}

extern "C" int tevo_test_heig(tevo_ctx* ctx, int n, const tevo_complex* A, double* evals_desc, tevo_complex* U) {
  TEVO_TRY(ctx)
  require(ctx && A && evals_desc && U && n >= 1, "tevo_test_heig: bad argument");
  DevBuf dA = ctx->alloc((size_t)n * n), dU = ctx->alloc((size_t)n * n);
  TEVO_CUDA_CHECK(cudaMemcpyAsync(dA.p, A, (size_t)n * n * sizeof(cplx), cudaMemcpyHostToDevice, ctx->stream));
  Split& sp = ctx->split;
  sp.reserve_vec((size_t)n);
  sp.heig().factor(ctx->stream, n, dA.p, n);
  TruncParams tp = to_params(nullptr);
  sort_truncate(ctx->stream, sp.heig().evals(), n, n, tp, sp.sorted_dev(), sp.perm_dev(), sp.info_dev());
  sp.heig().vectors(ctx->stream, sp.perm_dev(), n, dU.p, n);
  TEVO_CUDA_CHECK(cudaMemcpyAsync(evals_desc, sp.sorted_dev(), sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  TEVO_CUDA_CHECK(cudaMemcpyAsync(U, dU.p, (size_t)n * n * sizeof(cplx), cudaMemcpyDeviceToHost, ctx->stream));
  TEVO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  ctx->release(dA);
  ctx->release(dU);
  TEVO_CATCH
}

extern "C" int tevo_profile_enable(int on) {
  tevo::g_prof.on = (on != 0);
  return TEVO_OK;
}

/* JSON {"phase": {"ms": total device time, "count": n}, ...} of the phases timed since the last reset */
extern "C" int tevo_profile_read(char* buf, size_t cap, int reset) {
  if (!buf || cap == 0) return TEVO_EINVAL;
  std::string s = tevo::g_prof.json(reset != 0);
  if (s.size() + 1 > cap) return TEVO_EINVAL;
  memcpy(buf, s.c_str(), s.size() + 1);
  return TEVO_OK;
}

/* development counters: out[0..2] = CholQR single-pass / double-pass / failed, out[3..10] = pivot-ratio histogram */
/* two-stage tridiagonalisation, stage by stage, on host buffers (tests / development):
 *   band_out (16 x n, column-major): the band matrix after stage 1, B(c+d, c) at [d + 16*c], d <= 8
 *   v_out (n x n): A after stage 1 = Householder vectors (explicit zeros / unit entries), taus_out (n)
 *   d_out (n), e_out (n-1): the tridiagonal matrix after stage 2
 *   x (n x k, in/out, optional): overwritten by Q2 * x_real where x_real (n x k doubles, column-major) is the input */
extern "C" int tevo_test_sbr(tevo_ctx* ctx, int n, const tevo_complex* A, tevo_complex* band_out, tevo_complex* v_out,
                             tevo_complex* taus_out, double* d_out, double* e_out, int k, const double* x_real,
                             tevo_complex* x_out) {
  TEVO_TRY(ctx)
  require(ctx && A && n >= 2, "tevo_test_sbr: bad argument");
  Sbr sbr;
  int sms = 148;
  TEVO_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
  sbr.init(sms);
  cplx* dA = ctx->get_ws(WS_T1, (size_t)n * n);
  cplx* dt = ctx->get_ws(WS_T2, (size_t)n + 64 + (size_t)n * std::max(k, 1));
  double* dd = reinterpret_cast<double*>(ctx->get_ws(WS_W, (size_t)2 * n + (size_t)n * std::max(k, 1)));
  TEVO_CUDA_CHECK(cudaMemcpyAsync(dA, A, sizeof(cplx) * (size_t)n * n, cudaMemcpyHostToDevice, ctx->stream));
  sbr.reduce_to_band(ctx->stream, n, dA, n, dt, 0);
  if (band_out)
    TEVO_CUDA_CHECK(cudaMemcpyAsync(band_out, sbr.band(), sizeof(cplx) * (size_t)SBR_LDB * n, cudaMemcpyDeviceToHost,
                                    ctx->stream));
  if (v_out) TEVO_CUDA_CHECK(cudaMemcpyAsync(v_out, dA, sizeof(cplx) * (size_t)n * n, cudaMemcpyDeviceToHost, ctx->stream));
  if (taus_out) TEVO_CUDA_CHECK(cudaMemcpyAsync(taus_out, dt, sizeof(cplx) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  TEVO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  sbr.chase(ctx->stream, n, dd, dd + n);
  if (d_out) TEVO_CUDA_CHECK(cudaMemcpyAsync(d_out, dd, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  if (e_out) TEVO_CUDA_CHECK(cudaMemcpyAsync(e_out, dd + n, sizeof(double) * (n - 1), cudaMemcpyDeviceToHost, ctx->stream));
  if (k > 0 && x_real && x_out) {
    double* dz = dd + 2 * n;
    cplx* dx = dt + n + 64;
    TEVO_CUDA_CHECK(cudaMemcpyAsync(dz, x_real, sizeof(double) * (size_t)n * k, cudaMemcpyHostToDevice, ctx->stream));
    sbr.apply_q2(ctx->stream, n, k, dz, n, nullptr, dx, n);
    TEVO_CUDA_CHECK(cudaMemcpyAsync(x_out, dx, sizeof(cplx) * (size_t)n * k, cudaMemcpyDeviceToHost, ctx->stream));
  }
  TEVO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  sbr.destroy();
  TEVO_CATCH
}

extern "C" int tevo_test_potrf_block(tevo_ctx* ctx, int variant, int nb, const tevo_complex* G, tevo_complex* L_out,
                                     tevo_complex* X_out, double* stats_out, int reps, double* us_out,
                                     double* cycles_out) {
  TEVO_TRY(ctx)
  require(ctx && G && L_out && X_out && stats_out && nb >= 1 && nb <= 64 && variant >= 0 && variant <= 2,
          "tevo_test_potrf_block: bad argument");
  const size_t nn = (size_t)nb * nb;
  DevBuf dG = ctx->alloc(nn), dL = ctx->alloc(nn), dX = ctx->alloc(nn), dS = ctx->alloc(8);
  double* d_stats = reinterpret_cast<double*>(dS.p);
  long long* d_cyc = reinterpret_cast<long long*>(dS.p + 1);  // 8 x 8 bytes after the two doubles
  TEVO_CUDA_CHECK(cudaMemcpyAsync(dG.p, G, sizeof(cplx) * nn, cudaMemcpyHostToDevice, ctx->stream));
  TEVO_CUDA_CHECK(cudaMemsetAsync(dL.p, 0, sizeof(cplx) * nn, ctx->stream));
  TEVO_CUDA_CHECK(cudaMemsetAsync(dX.p, 0, sizeof(cplx) * nn, ctx->stream));
  TEVO_CUDA_CHECK(cudaMemsetAsync(dS.p, 0, sizeof(cplx) * 8, ctx->stream));
  for (int i = 0; i < 2; ++i)  // the second launch is the one reported: instruction cache warm for the cycle counts
    CholQR::block_step(ctx->stream, variant, nb, dG.p, dL.p, dX.p, nb, d_stats, cycles_out ? d_cyc : nullptr);
  TEVO_CUDA_CHECK(cudaMemcpyAsync(L_out, dL.p, sizeof(cplx) * nn, cudaMemcpyDeviceToHost, ctx->stream));
  TEVO_CUDA_CHECK(cudaMemcpyAsync(X_out, dX.p, sizeof(cplx) * nn, cudaMemcpyDeviceToHost, ctx->stream));
  TEVO_CUDA_CHECK(cudaMemcpyAsync(stats_out, d_stats, sizeof(double) * 2, cudaMemcpyDeviceToHost, ctx->stream));
  long long cyc[8] = {0};
  if (cycles_out) TEVO_CUDA_CHECK(cudaMemcpyAsync(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost, ctx->stream));
  TEVO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  if (cycles_out)
    for (int i = 0; i < 8; ++i) cycles_out[i] = (double)cyc[i];
  if (reps > 0 && us_out) {
    cudaEvent_t e0, e1;
    TEVO_CUDA_CHECK(cudaEventCreate(&e0));
    TEVO_CUDA_CHECK(cudaEventCreate(&e1));
    TEVO_CUDA_CHECK(cudaEventRecord(e0, ctx->stream));
    for (int i = 0; i < reps; ++i) CholQR::block_step(ctx->stream, variant, nb, dG.p, dL.p, dX.p, nb, d_stats, nullptr);
    TEVO_CUDA_CHECK(cudaEventRecord(e1, ctx->stream));
    TEVO_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    TEVO_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *us_out = 1e3 * (double)ms / reps;  // back-to-back launches of a single-CTA kernel: kernel time + launch gap
  }
  ctx->release(dG);
  ctx->release(dL);
  ctx->release(dX);
  ctx->release(dS);
  TEVO_CATCH
}

extern "C" int tevo_debug_set_zgemm_tile(int tile) {
  if (tile < -1 || tile > 2) return TEVO_EINVAL;
  zgemm_force_tile(tile);
  return TEVO_OK;
}

extern "C" int tevo_debug_set_zgemm_3m(int on) {
  zgemm_set_3m(on != 0);
  return TEVO_OK;
}

extern "C" int tevo_debug_set_heig_mode(int mode) {
  heig_set_mode(mode);
  return TEVO_OK;
}

extern "C" int tevo_debug_counters(tevo_ctx* ctx, double* out, int n) {
  if (!ctx || !out || n < 11) return TEVO_EINVAL;
  CholQR& q = ctx->split.cholqr();
  out[0] = (double)q.n_single;
  out[1] = (double)(q.n_double + q.n_series);
  out[2] = (double)q.n_fail;
  for (int i = 0; i < 8; ++i) out[3 + i] = q.ratio_hist[i];
  if (n >= 12) out[11] = g_heff_flops;
  if (n >= 15) heig_counters(&out[12], &out[13], &out[14]);
  if (n >= 17) sbr_counters(&out[15], &out[16]);
  return TEVO_OK;
}

// =====================================================================================================
// Layer-parallel TEBD in canonical B-form (Vidal / Hastings).  NOT the reference's sequential algorithm: the gates of
// a layer are independent here, which is what chain-segment sharding over GPUs needs; identical to
// the sequential sweep without truncation, O(discarded weight) apart with it.
// =====================================================================================================
static double* lam_buf(tevo_mps* psi, int j, int n) {
  if (psi->lam.empty()) {
    psi->lam.assign(psi->N + 1, nullptr);
    psi->lam_cap.assign(psi->N + 1, 0);
  }
  if (psi->lam_cap[j] < n) {
    if (psi->lam[j]) TEVO_CUDA_CHECK(cudaFree(psi->lam[j]));
    psi->lam[j] = nullptr;
    int want = n + n / 4 + 8;
    TEVO_CUDA_CHECK(cudaMalloc(&psi->lam[j], sizeof(double) * want));
    psi->lam_cap[j] = want;
  }
  return psi->lam[j];
}

static void set_lam_ones(tevo_mps* psi, int j) {
  double one = 1.0;
  double* p = lam_buf(psi, j, 1);
  TEVO_CUDA_CHECK(cudaMemcpyAsync(p, &one, sizeof(double), cudaMemcpyHostToDevice, psi->ctx->stream));
}

/* psi -> right-orthonormal B tensors + Schmidt values on every bond (centre ends on site 0) */
extern "C" int tevo_mps_to_bform(tevo_mps* psi) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi, "null");
  tevo_ctx* c = psi->ctx;
  const int N = psi->N, d = psi->d;
  orthogonalize(psi, N - 1);
  {
    size_t ne = (size_t)psi->chi[N - 1] * d * psi->chi[N];
    zdotc(c->stream, (long)ne, psi->site[N - 1].p, psi->site[N - 1].p, c->d_scal);
    zscal_rsqrt(c->stream, (long)ne, c->d_scal, psi->site[N - 1].p);
  }
  TruncParams tp = to_params(nullptr);
  for (int j = N - 1; j >= 1; --j) {
    const int m = psi->chi[j], n = d * psi->chi[j + 1];
    SplitResult r = c->split.run(c->stream, psi->site[j].p, m, n, /*left=*/false, tp, false, nullptr, 0,
                                 [&](size_t ne) { return c->alloc(ne); });
    const int k = r.k, m2 = psi->chi[j - 1] * d;
    sqrt_scaled(c->stream, k, c->split.sorted_dev(), nullptr, lam_buf(psi, j, k));
    DevBuf nb = c->alloc((size_t)m2 * k);
    zgemm(c->stream, OP_N, OP_N, m2, k, m, ONE, psi->site[j - 1].p, m2, r.L.p, m, ZERO, nb.p, m2);
    c->release(r.L);
    set_site(psi, j, r.R);
    set_site(psi, j - 1, nb);
    psi->chi[j] = k;
  }
  set_lam_ones(psi, 0);
  set_lam_ones(psi, N);
  psi->llim = -1;
  psi->rlim = 1;
  psi->bform = true;
  TEVO_CATCH
}

// one Hastings update of bond b, issued on lane context c (its stream, workspaces, split and buffer pool)
static void bform_gate(tevo_mps* psi, tevo_ctx* c, int b, const cplx* dG_gate, const TruncParams& tp, bool normalize,
                       tevo_spec* spec, double* eigs, int eigs_cap) {
  require(b >= 0 && b < psi->N - 1, "tevo_apply_layer_bform: bond out of range");
  const int d = psi->d;
  const int chil = psi->chi[b], m = chil * d, kk = psi->chi[b + 1], n = d * psi->chi[b + 2];
  cplx* C = c->get_ws(WS_THETA, (size_t)m * n);
  {
    ProfScope ps(c->stream, P_THETA);
    zgemm(c->stream, OP_N, OP_N, m, n, kk, ONE, psi->site[b].p, m, psi->site[b + 1].p, kk, ZERO, C, m);
  }
  gate_apply(c->stream, C, chil, d, psi->chi[b + 2], dG_gate);
  cplx* th = c->get_ws(WS_T1, (size_t)m * n);
  scale_rows(c->stream, m, n, chil, psi->lam[b], C, th);
  SplitResult r = c->split.run(c->stream, th, m, n, /*left=*/false, tp, false, eigs, eigs_cap,
                               [&](size_t ne) { return c->alloc(ne); });
  const int k = r.k;
  DevBuf nb = c->alloc((size_t)m * k);
  {
    ProfScope ps(c->stream, P_SPLIT_GEMM);
    zgemm(c->stream, OP_N, OP_C, m, k, n, ONE, C, m, r.R.p, k, ZERO, nb.p, m);  // B'_b = C V
  }
  const double* denom = normalize ? &c->split.info_dev()->sum_kept : nullptr;
  if (normalize) zscal_rsqrt(c->stream, (long)m * k, denom, nb.p);
  sqrt_scaled(c->stream, k, c->split.sorted_dev(), denom, lam_buf(psi, b + 1, k));
  c->release(r.L);
  c->release(psi->site[b]);
  psi->site[b] = nb;
  c->release(psi->site[b + 1]);
  psi->site[b + 1] = r.R;
  for (auto& x : c->split.take_released()) c->release(x);
  psi->chi[b + 1] = k;
  if (spec) fill_spec(r.info, spec);
}

/* Hastings' inverse-free update of every gate of one layer.  The gates are independent (any order), so two host
 * threads drive two lanes (streams with their own workspaces) that split the layer between them: the eigensolver's
 * panel kernel is latency-bound (12 % of the warp slots busy), two of them on half the SMs each overlap well.
 * TEVO_LANES=1 forces the single-lane path. */
extern "C" int tevo_apply_layer_bform(tevo_mps* psi, int ngates, const int* bonds, const tevo_complex* Gs,
                                      const tevo_trunc* trunc, int normalize, tevo_spec* specs, double* eigs,
                                      int eigs_stride) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi && (ngates == 0 || (bonds && Gs)), "tevo_apply_layer_bform: null argument");
  require(psi->bform, "tevo_apply_layer_bform: state is not in B-form (call tevo_mps_to_bform)");
  tevo_ctx* c = psi->ctx;
  const int d = psi->d, d4 = d * d * d * d;
  cplx* dG = (cplx*)c->small(sizeof(cplx) * d4 * (size_t)std::max(ngates, 1));
  if (ngates) upload_small(c, dG, Gs, sizeof(cplx) * d4 * (size_t)ngates);
  const TruncParams tp = to_params(trunc);
  // all bonds must be distinct and non-adjacent (one even/odd layer): the gates then touch disjoint data
  bool disjoint = true;
  for (int g = 0; g < ngates && disjoint; ++g)
    for (int h = g + 1; h < ngates; ++h)
      if (std::abs(bonds[g] - bonds[h]) < 2) {
        disjoint = false;
        break;
      }
  static const int nlanes_cfg = [] {
    const char* e = getenv("TEVO_LANES");
    int v = e ? atoi(e) : 3;
    return v < 1 ? 1 : (v > 8 ? 8 : v);
  }();
  const int nl = std::min(nlanes_cfg, ngates);
  if (!disjoint || nl < 2) {
    for (int g = 0; g < ngates; ++g)
      bform_gate(psi, c, bonds[g], dG + (size_t)g * d4, tp, normalize != 0, specs ? &specs[g] : nullptr,
                 eigs ? eigs + (size_t)g * eigs_stride : nullptr, eigs ? eigs_stride : 0);
  } else {
    while ((int)c->aux.size() < nl - 1) {
      tevo_ctx* aux = nullptr;
      int st = tevo_ctx_create(c->device, &aux);
      if (st != TEVO_OK) throw TevoError(st, std::string("cannot create an extra lane: ") + tevo_last_error(nullptr));
      aux->is_aux = true;
      c->aux.push_back(aux);
    }
    std::vector<tevo_ctx*> lanes;
    lanes.push_back(c);
    for (int i = 0; i < nl - 1; ++i) lanes.push_back(c->aux[i]);
    TEVO_CUDA_CHECK(cudaStreamSynchronize(c->stream));  // gates uploaded, previous layer complete
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device);
    for (auto* l : lanes) l->split.heig().set_max_ctas(std::max(8, nsm / nl));
    if (psi->lam.empty()) throw TevoError(TEVO_EINTERNAL, "B-form state without Schmidt values");
    std::vector<std::exception_ptr> err(nl, nullptr);
    auto work = [&](int lane) {
      try {
        TEVO_CUDA_CHECK(cudaSetDevice(c->device));
        tevo_ctx* lc = lanes[lane];
        for (int g = lane; g < ngates; g += nl)
          bform_gate(psi, lc, bonds[g], dG + (size_t)g * d4, tp, normalize != 0, specs ? &specs[g] : nullptr,
                     eigs ? eigs + (size_t)g * eigs_stride : nullptr, eigs ? eigs_stride : 0);
        TEVO_CUDA_CHECK(cudaStreamSynchronize(lc->stream));
      } catch (...) {
        err[lane] = std::current_exception();
      }
    };
    std::vector<std::thread> ths;
    for (int i = 1; i < nl; ++i)
      ths.emplace_back([&, i] {
        tevo::tl_prof_thread = false;
        work(i);
      });
    work(0);
    for (auto& th : ths) th.join();
    for (auto* l : lanes) l->split.heig().set_max_ctas(0);
    for (int i = 0; i < nl; ++i)
      if (err[i]) std::rethrow_exception(err[i]);
  }
  psi->llim = -1;
  psi->rlim = 1;
  TEVO_CATCH
}

/* <O_o> on site j of a B-form state: Tr(T^+ O T), T = diag(lam_j) B_j */
extern "C" int tevo_bform_expect_site(tevo_mps* psi, int j, int nops, const tevo_complex* ops, tevo_complex* out) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi && ops && out && nops >= 1 && 2 * nops <= tevo_ctx::NSCAL, "tevo_bform_expect_site: bad argument");
  require(psi->bform, "tevo_bform_expect_site: state is not in B-form");
  require(j >= 0 && j < psi->N, "tevo_bform_expect_site: site out of range");
  tevo_ctx* c = psi->ctx;
  const int d = psi->d;
  cplx* dops = (cplx*)c->small(sizeof(cplx) * d * d * nops);
  upload_small(c, dops, ops, sizeof(cplx) * d * d * nops);
  expect1(c->stream, psi->site[j].p, psi->chi[j], d, psi->chi[j + 1], psi->lam[j], nops, dops, c->d_scal);
  download_scalars(c, c->d_scal, 2 * nops, (double*)out);
  TEVO_CATCH
}

/* Schmidt values on the bond left of site j (host copies) */
extern "C" int tevo_mps_get_lambda(tevo_mps* psi, int j, double* host, int cap, int* n_out) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi && psi->bform && j >= 0 && j <= psi->N, "tevo_mps_get_lambda: bad argument or not in B-form");
  const int n = psi->chi[j];
  if (n_out) *n_out = n;
  if (host) {
    require(cap >= n, "tevo_mps_get_lambda: buffer too small");
    TEVO_CUDA_CHECK(cudaMemcpyAsync(host, psi->lam[j], sizeof(double) * n, cudaMemcpyDeviceToHost, psi->ctx->stream));
    TEVO_CUDA_CHECK(cudaStreamSynchronize(psi->ctx->stream));
  }
  TEVO_CATCH
}

extern "C" int tevo_mps_set_lambda(tevo_mps* psi, int j, int n, const double* host) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi && host && j >= 0 && j <= psi->N && n >= 1, "tevo_mps_set_lambda: bad argument");
  double* p = lam_buf(psi, j, n);
  TEVO_CUDA_CHECK(cudaMemcpyAsync(p, host, sizeof(double) * n, cudaMemcpyHostToDevice, psi->ctx->stream));
  TEVO_CUDA_CHECK(cudaStreamSynchronize(psi->ctx->stream));
  TEVO_CATCH
}

/* declare a state to be in B-form (all Schmidt vectors set by the caller, e.g. a segment received from a neighbour) */
extern "C" int tevo_mps_mark_bform(tevo_mps* psi, int on) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi, "null");
  if (on) {
    require(!psi->lam.empty(), "tevo_mps_mark_bform: no Schmidt values set");
    for (int j = 0; j <= psi->N; ++j) require(psi->lam[j] != nullptr, "tevo_mps_mark_bform: missing Schmidt values");
  }
  psi->bform = (on != 0);
  TEVO_CATCH
}

/* raw device views for peer-to-peer exchange of boundary data (NCCL send/recv from the host layer) */
extern "C" int tevo_mps_site_devptr(tevo_mps* psi, int j, void** ptr, int* chil, int* chir) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi && ptr && j >= 0 && j < psi->N, "tevo_mps_site_devptr: bad argument");
  *ptr = (void*)psi->site[j].p;
  if (chil) *chil = psi->chi[j];
  if (chir) *chir = psi->chi[j + 1];
  TEVO_CATCH
}

extern "C" int tevo_mps_lambda_devptr(tevo_mps* psi, int j, void** ptr, int* n) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi && ptr && j >= 0 && j <= psi->N && !psi->lam.empty() && psi->lam[j], "tevo_mps_lambda_devptr: bad argument");
  *ptr = (void*)psi->lam[j];
  if (n) *n = psi->chi[j];
  TEVO_CATCH
}

/* replace site j by a copy of a device buffer (column-major (l,s,r)); lam != NULL also sets lam[j] (n = chil) */
extern "C" int tevo_mps_set_site_from_device(tevo_mps* psi, int j, int chil, int chir, const void* dev_site) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi && dev_site && j >= 0 && j < psi->N && chil >= 1 && chir >= 1, "tevo_mps_set_site_from_device: bad argument");
  tevo_ctx* c = psi->ctx;
  const size_t ne = (size_t)chil * psi->d * chir;
  DevBuf nb = c->alloc(ne);
  TEVO_CUDA_CHECK(cudaMemcpyAsync(nb.p, dev_site, ne * sizeof(cplx), cudaMemcpyDeviceToDevice, c->stream));
  set_site(psi, j, nb);
  psi->chi[j] = chil;
  psi->chi[j + 1] = chir;
  TEVO_CATCH
}

extern "C" int tevo_mps_set_lambda_from_device(tevo_mps* psi, int j, int n, const void* dev_lam) {
  TEVO_TRY(psi ? psi->ctx : nullptr)
  require(psi && dev_lam && j >= 0 && j <= psi->N && n >= 1, "tevo_mps_set_lambda_from_device: bad argument");
  double* p = lam_buf(psi, j, n);
  TEVO_CUDA_CHECK(cudaMemcpyAsync(p, dev_lam, sizeof(double) * n, cudaMemcpyDeviceToDevice, psi->ctx->stream));
  TEVO_CATCH
}
