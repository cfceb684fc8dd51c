#include "prof.h"

#include <cstdio>

namespace tevo {

Prof g_prof;
thread_local bool tl_prof_thread = true;

static const char* kNames[P_COUNT] = {"theta_gemm", "gate_apply", "rho_gemm",  "tridiag_panel", "tridiag_her2k", "tridiag_wy",
                                      "stedc",      "backtransform", "split_gemm", "residual",    "truncate",      "hop_gemm",
                                      "heff_apply", "krylov_blas1",  "env_update", "measure",
                                      "cholqr_gram", "cholqr_factor", "cholqr_inverse", "cholqr_q",
                                      "sbr_stage1", "sbr_chase", "sbr_q2"};

cudaEvent_t Prof::get() {
  if (!pool.empty()) {
    cudaEvent_t e = pool.back();
    pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

void Prof::flush() {
  for (auto& r : recs) {
    cudaEventSynchronize(r.b);
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    ms[r.id] += t;
    count[r.id] += 1;
    pool.push_back(r.a);
    pool.push_back(r.b);
  }
  recs.clear();
}

std::string Prof::json(bool reset) {
  flush();
  std::string s = "{";
  bool first = true;
  char buf[256];
  for (int i = 0; i < P_COUNT; ++i) {
    if (count[i] == 0) continue;
    snprintf(buf, sizeof(buf), "%s\"%s\": {\"ms\": %.6f, \"count\": %ld}", first ? "" : ", ", kNames[i], ms[i], count[i]);
    s += buf;
    first = false;
  }
  s += "}";
  if (reset)
    for (int i = 0; i < P_COUNT; ++i) {
      ms[i] = 0;
      count[i] = 0;
    }
  return s;
}

}  // namespace tevo
