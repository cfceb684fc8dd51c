// micro-benchmark (development aid, not part of the library): what would one column of the tridiagonalisation cost if
// the CTAs of a panel were one thread-block cluster?  Measures, per iteration,
//   (a) cluster.sync() alone,
//   (b) an all-gather of one complex value per CTA through distributed shared memory + cluster.sync(),
//   (c) an all-reduce of a 16-entry vector per CTA through DSMEM (every CTA reads every other CTA's partials),
// for cluster sizes 2..16 (16 is non-portable).  It answers whether the last columns of a factorisation (trailing
// matrix small enough for the shared memory of one cluster) should leave the cooperative-grid kernel: the grid
// barrier costs ~1.2 us on B200, two of them per column.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_bar_bench cluster_bar_bench.cu && ./cluster_bar_bench
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

template <int MODE>
__global__ void k(int iters, double* out, long long* cycles) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ double2 mine[16];
  __shared__ double2 gathered[16 * 16];
  const unsigned rank = cluster.block_rank(), csize = cluster.num_blocks();
  double acc = 0.0;
  if (threadIdx.x < 16) mine[threadIdx.x] = make_double2(rank + threadIdx.x, 1.0);
  cluster.sync();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE >= 1) {
      // every CTA publishes, every CTA reads all peers (one value in mode 1, 16 values in mode 2)
      const int per = (MODE == 1) ? 1 : 16;
      if (threadIdx.x < per) mine[threadIdx.x].x += 1.0;
      cluster.sync();
      for (unsigned e = threadIdx.x; e < csize * per; e += blockDim.x) {
        const unsigned peer = e / per, idx = e % per;
        const double2* remote = cluster.map_shared_rank(mine, peer);
        gathered[e] = remote[idx];
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        for (unsigned e = 0; e < csize * per; ++e) acc += gathered[e].x;
      }
    }
    cluster.sync();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  if (acc == 123.456) out[0] = acc;
}

template <int MODE>
static void run(int csize, int iters, double* out, long long* cycles_dev) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(csize);
  cfg.blockDim = dim3(256);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (csize > 8) cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaError_t e = cudaLaunchKernelEx(&cfg, k<MODE>, iters, out, cycles_dev);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("mode %d cluster %2d: %s\n", MODE, csize, cudaGetErrorString(e));
    cudaGetLastError();
    return;
  }
  long long c = 0;
  cudaMemcpy(&c, cycles_dev, sizeof(c), cudaMemcpyDeviceToHost);
  printf("mode %d cluster %2d: %.0f cycles per iteration\n", MODE, csize, (double)c / iters);
}

int main() {
  double* out;
  long long* cycles;
  cudaMalloc(&out, 8);
  cudaMalloc(&cycles, 8);
  const int iters = 2000;
  for (int csize : {1, 2, 4, 8, 16}) {
    run<0>(csize, iters, out, cycles);
    run<1>(csize, iters, out, cycles);
    run<2>(csize, iters, out, cycles);
  }
  return 0;
}
