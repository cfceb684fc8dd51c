// Dependent-issue latency of the FP64 operations the eigensolver's critical paths are made of (one warp, one SM):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_latency proto/fp64_latency.cu && /tmp/fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

__global__ void chain(int mode, int iters, double seed, double* out, long long* cycles) {
  double x = seed + threadIdx.x * 1e-9, y = 1.0000001, acc = 0.0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (mode == 0) {            // dependent DFMA
#pragma unroll
      for (int u = 0; u < 16; ++u) x = fma(x, y, 1e-9);
    } else if (mode == 1) {     // dependent DADD
#pragma unroll
      for (int u = 0; u < 16; ++u) x = x + y;
    } else if (mode == 2) {     // dependent DMUL
#pragma unroll
      for (int u = 0; u < 16; ++u) x = x * y;
    } else if (mode == 3) {     // shuffle + DADD (one stage of a warp reduction)
#pragma unroll
      for (int u = 0; u < 16; ++u) x = x + __shfl_xor_sync(0xffffffffu, x, 1);
    } else if (mode == 4) {     // sqrt
#pragma unroll
      for (int u = 0; u < 16; ++u) x = sqrt(x + 2.0);
    } else if (mode == 5) {     // division
#pragma unroll
      for (int u = 0; u < 16; ++u) x = 1.0 / (x + 2.0);
    } else if (mode == 6) {     // rsqrt
#pragma unroll
      for (int u = 0; u < 16; ++u) x = rsqrt(x + 2.0);
    } else if (mode == 7) {     // __drcp_rn
#pragma unroll
      for (int u = 0; u < 16; ++u) x = __drcp_rn(x + 2.0);
    } else if (mode == 8) {     // 4 independent DFMA chains (ILP)
      double a = x, b = x + 1, c = x + 2, d = x + 3;
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        a = fma(a, y, 1e-9); b = fma(b, y, 1e-9); c = fma(c, y, 1e-9); d = fma(d, y, 1e-9);
      }
      x = a + b + c + d;
    } else if (mode == 9) {     // dependent FFMA for comparison
      float f = (float)x;
#pragma unroll
      for (int u = 0; u < 16; ++u) f = fmaf(f, 1.0000001f, 1e-9f);
      x = f;
    }
    acc += x;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + x;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, sizeof(double) * 148 * 1024);
  cudaMalloc(&cyc, sizeof(long long));
  const char* names[] = {"DFMA dependent", "DADD dependent", "DMUL dependent", "SHFL+DADD", "sqrt(double)", "1.0/x",
                         "rsqrt(double)", "__drcp_rn", "4 independent DFMA chains (per 4 ops)", "FFMA dependent"};
  for (int warps = 1; warps <= 16; warps *= 4) {
    for (int mode = 0; mode < 10; ++mode) {
      const int iters = 256;
      chain<<<1, 32 * warps>>>(mode, iters, 1.0, out, cyc);
      chain<<<1, 32 * warps>>>(mode, iters, 1.0, out, cyc);
      long long c = 0;
      cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
      printf("%2d warp(s)  %-40s %7.1f cycles per op\n", warps, names[mode], (double)c / (iters * 16.0));
    }
  }
  return 0;
}
