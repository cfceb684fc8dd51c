// Common device/host helpers for libtevo (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <atomic>
#include <mutex>
#include <string>

namespace tevo {

typedef double2 cplx;  // (re, im), 16-byte aligned, bit-compatible with C `double _Complex` / Julia ComplexF64

__host__ __device__ inline cplx mk(double re, double im) { return make_double2(re, im); }
__host__ __device__ inline cplx cadd(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
__host__ __device__ inline cplx csub(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
__host__ __device__ inline cplx cmul(cplx a, cplx b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__host__ __device__ inline cplx cconj(cplx a) { return mk(a.x, -a.y); }
__host__ __device__ inline cplx cscale(double s, cplx a) { return mk(s * a.x, s * a.y); }
__host__ __device__ inline double cabs2(cplx a) { return a.x * a.x + a.y * a.y; }
// a += conj(x) * y
__host__ __device__ inline void cfma_conj(cplx& a, cplx x, cplx y) {
  a.x += x.x * y.x + x.y * y.y;
  a.y += x.x * y.y - x.y * y.x;
}
__host__ __device__ inline void cfma(cplx& a, cplx x, cplx y) {
  a.x += x.x * y.x - x.y * y.y;
  a.y += x.x * y.y + x.y * y.x;
}

enum Op { OP_N = 0, OP_T = 1, OP_C = 2 };

#define TEVO_CUDA_CHECK(expr)                                                                  \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      throw tevo::CudaError(_e, __FILE__, __LINE__);                                           \
    }                                                                                          \
  } while (0)

extern std::atomic<long long> g_launches;  // kernels launched by the library (all contexts)
#define TEVO_LAUNCHED()                 \
  do {                                  \
    ++tevo::g_launches;                 \
    TEVO_CUDA_CHECK(cudaGetLastError()); \
  } while (0)

struct CudaError {
  cudaError_t err;
  const char* file;
  int line;
  CudaError(cudaError_t e, const char* f, int l) : err(e), file(f), line(l) {}
};

struct TevoError {
  int code;
  std::string msg;
  TevoError(int c, std::string m) : code(c), msg(std::move(m)) {}
};

// cudaFuncSetAttribute and friends are per DEVICE: run a set-up lambda once for every device a launch path is used
// on (a process may hold contexts on several GPUs through tevo_ctx_create(device)).
struct PerDeviceOnce {
  static constexpr int MAXDEV = 64;
  std::mutex mu;
  std::atomic<bool> done[MAXDEV];
  PerDeviceOnce() {
    for (auto& d : done) d.store(false);
  }
  template <class F>
  void operator()(F&& f) {
    int dev = 0;
    TEVO_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAXDEV) {
      f();
      return;
    }
    if (done[dev].load(std::memory_order_acquire)) return;
    std::lock_guard<std::mutex> g(mu);
    if (!done[dev].load(std::memory_order_relaxed)) {
      f();
      done[dev].store(true, std::memory_order_release);
    }
  }
};

// warp / block reductions (double)
__device__ inline double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of up to 3 doubles at once; result valid in all threads. blockDim.x multiple of 32, <= 1024
template <int NV>
__device__ inline void block_sum(double (&v)[NV], double* smem /* >= NV*32 doubles */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) smem[i * 32 + wid] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double x = (lane < nw) ? smem[i * 32 + lane] : 0.0;
    v[i] = warp_sum(x);
  }
}

}  // namespace tevo
