// Microbenchmark: does a crowd of warps spinning on mbarrier.try_wait (what the epilogue warps of the conv kernels do
// while the accumulators fill) slow down the single thread that issues tcgen05.mma?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I vlnce_b200/csrc -o tools/micro/mma_spin tools/micro/mma_spin.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"

using namespace vlnce::ptx;

// SPIN: 0 = other warps parked at a named barrier, 1 = spin on try_wait, 2 = try_wait + nanosleep(200) backoff,
//       3 = one poller per warp (lane 0), the other lanes parked at __syncwarp
template <int SPIN>
__global__ void __launch_bounds__(352, 1) bench(int n_mma, int N, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bar, done;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&done, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = slot;
  const uint32_t idesc = make_idesc_f16(128, (uint32_t)N, 0u);
  const uint64_t a0 = make_smem_desc(base, 128), b0 = make_smem_desc(base + 16384, 128);
  if (warp == 2) {
    if (lane == 0) {
      const long long t0 = clock64();
      for (int i = 0; i < n_mma; i += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem, a0 + 2 * k, b0 + 2 * k, idesc, 1u);
      }
      const long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, 0, 1);
      const long long t2 = clock64();
      out[0] = t1 - t0;
      out[1] = t2 - t0;
      mbar_arrive(&done);
    }
    __syncwarp();
  } else if (warp >= 3) {
    if (SPIN == 0) {
      // parked below
    } else if (SPIN == 1) {
      while (!mbar_try_wait(&done, 0)) {}
    } else if (SPIN == 2) {
      while (!mbar_try_wait(&done, 0)) __nanosleep(200);
    } else {
      if (lane == 0) { while (!mbar_try_wait(&done, 0)) {} }
      __syncwarp();
    }
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
  const int n = 2048;
  const char* names[4] = {"8 warps parked at a barrier", "8 warps spinning on try_wait", "8 warps try_wait + nanosleep(200)",
                          "8 single-lane pollers"};
  for (int N : {16, 64, 128, 256}) {
    for (int mode = 0; mode < 4; ++mode) {
      long long h[2] = {0, 0};
      for (int rep = 0; rep < 3; ++rep) {
        if (mode == 0) bench<0><<<1, 352, smem>>>(n, N, d);
        if (mode == 1) bench<1><<<1, 352, smem>>>(n, N, d);
        if (mode == 2) bench<2><<<1, 352, smem>>>(n, N, d);
        if (mode == 3) bench<3><<<1, 352, smem>>>(n, N, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      }
      printf("N=%3d %-36s: cycles per MMA issue %.1f  complete %.1f  (tensor floor %.1f)\n", N, names[mode],
             (double)h[0] / n, (double)h[1] / n, 128.0 * N / 256.0);
    }
  }
  return 0;
}
