// Microbenchmark: cost of the per-tile bookkeeping around tcgen05.mma in a single-thread issue loop: a
// tcgen05.fence::after_thread_sync, a tcgen05.commit to an mbarrier, an already-complete mbarrier.try_wait.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I vlnce_b200/csrc -o tools/micro/mma_loop tools/micro/mma_loop.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"

using namespace vlnce::ptx;

template <int MODE>
__global__ void __launch_bounds__(128, 1) bench(int n_mma, int N, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bar, ready, sink;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1); mbar_init(&ready, 1); mbar_init(&sink, 1);
    fence_barrier_init();
    mbar_arrive(&ready);      // phase 0 of `ready` is complete from now on
  }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = slot;
  const uint32_t idesc = make_idesc_f16(128, (uint32_t)N, 0u);
  const uint64_t a0 = make_smem_desc(base, 128), b0 = make_smem_desc(base + 16384, 128);
  if (warp == 1 && lane == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 4) {
      if (MODE == 1 || MODE == 4) tc_fence_after_sync();
      if (MODE == 3 || MODE == 4) { while (!mbar_try_wait(&ready, 0)) {} }
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem, a0 + 2 * k, b0 + 2 * k, idesc, 1u);
      if (MODE == 2 || MODE == 4) umma_commit(&sink);
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0, 1);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  const int smem = 64 * 1024;
  cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bench<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bench<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int n = 2048;
  const char* names[5] = {"4 MMAs", "fence::after_thread_sync + 4 MMAs", "4 MMAs + commit", "try_wait(complete) + 4 MMAs",
                          "fence + try_wait + 4 MMAs + commit"};
  for (int N : {16, 64, 128}) {
    for (int mode = 0; mode < 5; ++mode) {
      long long h[2] = {0, 0};
      for (int rep = 0; rep < 3; ++rep) {
        if (mode == 0) bench<0><<<1, 128, smem>>>(n, N, d);
        if (mode == 1) bench<1><<<1, 128, smem>>>(n, N, d);
        if (mode == 2) bench<2><<<1, 128, smem>>>(n, N, d);
        if (mode == 3) bench<3><<<1, 128, smem>>>(n, N, d);
        if (mode == 4) bench<4><<<1, 128, smem>>>(n, N, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      }
      printf("N=%3d %-38s: cycles per group of 4 MMAs: issue %.0f  complete %.0f\n", N, names[mode], 4.0 * h[0] / n, 4.0 * h[1] / n);
    }
  }
  return 0;
}
