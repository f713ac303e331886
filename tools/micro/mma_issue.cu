// Microbenchmark: how fast can ONE thread issue tcgen05.mma (128 x N x 16, kind::f16) and how long do
// dependent / independent accumulation chains take?  Variants: issuing thread inside an `if (lane == 0)` branch
// (what the conv kernels do) vs a converged warp with elect.sync; rolled vs unrolled k-steps; 1 vs 4 accumulators.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I vlnce_b200/csrc -o tools/micro/mma_issue tools/micro/mma_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"

using namespace vlnce::ptx;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}

template <int MODE>
__global__ void __launch_bounds__(128, 1) bench(int n_mma, int N, int chains, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = slot;
  const uint32_t idesc = make_idesc_f16(128, (uint32_t)N, 0u);
  const uint64_t a0 = make_smem_desc(base, 128), b0 = make_smem_desc(base + 16384, 128);
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp == 1) {
    if (MODE == 0) {          // divergent single thread, rolled loop, runtime trip count
      if (lane == 0) {
        t0 = clock64();
        for (int i = 0; i < n_mma; ++i)
          umma_bf16(tmem + (uint32_t)((i & (chains - 1)) * N), a0 + 2 * (i & 3), b0 + 2 * (i & 3), idesc, 1u);
        t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0, 1);
        t2 = clock64();
      }
    } else if (MODE == 1) {   // divergent single thread, k-steps unrolled by 4
      if (lane == 0) {
        t0 = clock64();
        for (int i = 0; i < n_mma; i += 4) {
          const uint32_t d = tmem + (uint32_t)(((i >> 2) & (chains - 1)) * N);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d, a0 + 2 * k, b0 + 2 * k, idesc, 1u);
        }
        t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0, 1);
        t2 = clock64();
      }
    } else {                  // converged warp, elect.sync around the issue, unrolled by 4
      t0 = clock64();
      for (int i = 0; i < n_mma; i += 4) {
        const uint32_t d = tmem + (uint32_t)(((i >> 2) & (chains - 1)) * N);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d, a0 + 2 * k, b0 + 2 * k, idesc, 1u);
        }
        __syncwarp();
      }
      t1 = clock64();
      if (elect_one()) umma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, 0, 1);
      t2 = clock64();
    }
    if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
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
  const int n = 512;
  printf("n_mma %d per run; cycles per MMA: issue-only / until complete\n", n);
  for (int N : {16, 32, 64, 128, 256}) {
    for (int chains : {1, 4}) {
      if (chains * N > 512) continue;
      for (int mode = 0; mode < 3; ++mode) {
        long long h[2] = {0, 0};
        for (int rep = 0; rep < 3; ++rep) {
          if (mode == 0) bench<0><<<1, 128, smem>>>(n, N, chains, d);
          if (mode == 1) bench<1><<<1, 128, smem>>>(n, N, chains, d);
          if (mode == 2) bench<2><<<1, 128, smem>>>(n, N, chains, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        }
        printf("N=%3d chains=%d mode=%d (%s): issue %.1f  complete %.1f  (tensor floor %.1f)\n", N, chains, mode,
               mode == 0 ? "lane0 rolled" : mode == 1 ? "lane0 unroll4" : "elect unroll4", (double)h[0] / n, (double)h[1] / n,
               128.0 * N / 256.0);
      }
    }
  }
  return 0;
}
