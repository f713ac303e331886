// Shared helpers for the vlnce_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/vlnce_b200.h"

#include <stdlib.h>

namespace vlnce {

// ---- error plumbing: integer status across the C ABI, message kept per thread ----
void set_error(const char* fmt, ...);
int cuda_status(cudaError_t e, const char* what);
void count_launch();

#define VLNCE_CHECK_ARG(cond, ...)            \
  do {                                        \
    if (!(cond)) {                            \
      ::vlnce::set_error(__VA_ARGS__);        \
      return VLNCE_ERR_INVALID_ARGUMENT;      \
    }                                         \
  } while (0)

#define VLNCE_LAUNCH_CHECK(what)                                   \
  do {                                                             \
    ::vlnce::count_launch();                                       \
    cudaError_t e__ = cudaGetLastError();                          \
    if (e__ != cudaSuccess) return ::vlnce::cuda_status(e__, what); \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// The step is a chain of ~370 short dependent kernels; with the programmatic-stream-serialization launch
// attribute kernel i+1 is scheduled while kernel i still runs and parks at griddepcontrol.wait (which
// returns once kernel i has completed and its memory is visible), so launch latency and the prologue
// (barrier init, tensor-map prefetch, TMEM allocation) overlap the predecessor's tail.  Every kernel
// launched through launch_pdl() executes pdl_wait() before its first global-memory access and
// pdl_trigger() only after it owns its on-chip resources (TMEM).  VLNCE_PDL=0 turns the attribute off.
bool pdl_enabled();
// Pin the L1/shared-memory split of every kernel launched through launch_pdl() to "max shared" (once per
// kernel): the trunks alternate tensor-core tiles (>100 KB dynamic smem) with small statistics / pointwise
// kernels, and an SM has to drain before it can change its carve-out.  VLNCE_CARVEOUT=0 leaves the driver default.
bool carveout_first_use(const void* kernel);
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
  if (carveout_first_use(reinterpret_cast<const void*>(kernel)))
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ---- tiny device helpers ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum for blockDim.x <= 1024 (multiple of 32); `scratch` = 32 floats of smem
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float sigmoid_precise(float x) { return 1.f / (1.f + expf(-x)); }

template <typename T>
__device__ __forceinline__ float to_float(T v);
template <>
__device__ __forceinline__ float to_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_float<uint8_t>(uint8_t v) { return (float)v; }
template <>
__device__ __forceinline__ float to_float<__half>(__half v) { return __half2float(v); }

template <typename T>
__device__ __forceinline__ T from_float(float v);
template <>
__device__ __forceinline__ float from_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }

// Rows of the instruction batch that one CTA of the persistent LSTM kernels carries (forward and backward agree).
// Default 1: with four rows per CTA the recurrence holds a quarter of the SMs (R/4 instead of R CTAs) but each step
// takes ~2.4x as long; measured on the B200 at 64 rows the training step does not get faster (2.98 vs 2.94 ms) and the
// act latency gets worse, so the 4-row kernels are opt-in: VLNCE_LSTM_NR=4 (covered by tests/test_gpu_head_ops.py,
// which forces both variants).
int lstm_rows_per_cta(int R);   // common.cu; vlnce_set_lstm_rows() / VLNCE_LSTM_NR

}  // namespace vlnce
