// Device-resident trajectory cache: batch assembly of the DAgger / recollect trainers straight from
// trajectories that live in HBM.
//
//   reference: dagger_trainer.py:39-114 (collate_fn: pad every trajectory to the batch's longest, stack on a new
//   batch axis, flatten time-major), :189-211 (per-item tensors + inflection weights), :560-583 (every observation is
//   cast to float32 on the device; records written with IL.DAGGER.lmdb_fp16 arrive as float16, :342-346).
//
// One launch per observation key: out[(t*N + n), :] = float(arena[offsets[n] + t*D + :]) for t < lengths[n], else the
// pad value (1.0 for observations — SURVEY Q7 —, 0 for actions).  A second tiny kernel assembles the three (T,N)
// integer / weight matrices and the not-done mask.  HBM-bound copies: rows are D contiguous elements, read and written
// with 16-byte accesses when the row size and alignment allow.
#include "common.cuh"

#include <algorithm>
#include <cstdint>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace vlnce {
namespace {

template <typename TS>
__device__ __forceinline__ float to_f32(TS v) { return (float)v; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// grid: (chunks of a row, T*N rows); each thread converts VEC consecutive elements
template <typename TS, int VEC>
__global__ void collate_rows_kernel(const TS* __restrict__ arena, const long long* __restrict__ offsets,
                                    const int* __restrict__ lengths, float* __restrict__ out, int T, int N, long long D,
                                    float pad) {
  for (long long row = blockIdx.y; row < (long long)T * N; row += gridDim.y) {  // row = t*N + n
  const int t = (int)(row / N), n = (int)(row % N);
  const bool live = t < lengths[n];
  const TS* src = arena + offsets[n] + (long long)t * D;
  float* dst = out + row * D;
  if constexpr (VEC == 4) {
    // a trajectory that does not start on a 4-element boundary (never produced by TrajectoryCache: D % 4 == 0 there)
    // takes the scalar route for its rows
    if (live && (reinterpret_cast<uintptr_t>(src) % (4 * sizeof(TS))) != 0) {
      for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < D; i += (long long)gridDim.x * blockDim.x)
        dst[i] = to_f32(src[i]);
      continue;
    }
  }
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC; i < D;
       i += (long long)gridDim.x * blockDim.x * VEC) {
    if constexpr (VEC == 4) {
      float4 v;
      if (live) {
        if constexpr (sizeof(TS) == 4) {
          v = *reinterpret_cast<const float4*>(src + i);
        } else if constexpr (sizeof(TS) == 2) {
          const uint2 raw = *reinterpret_cast<const uint2*>(src + i);
          const TS* h = reinterpret_cast<const TS*>(&raw);
          v = make_float4(to_f32(h[0]), to_f32(h[1]), to_f32(h[2]), to_f32(h[3]));
        } else {
          v = make_float4(to_f32(src[i]), to_f32(src[i + 1]), to_f32(src[i + 2]), to_f32(src[i + 3]));
        }
      } else {
        v = make_float4(pad, pad, pad, pad);
      }
      *reinterpret_cast<float4*>(dst + i) = v;
    } else {
      dst[i] = live ? to_f32(src[i]) : pad;
    }
  }
  }
}

// (T,N) matrices of one batch: prev_actions / oracle_actions padded with 0, inflection weights
// (dagger_trainer.py:199-211: w = [1, coef][a_t != a_{t-1}], first step counts as an inflection; 0 in the padding),
// not_done_masks = 1 except the first row block (collate_fn :106-110).
__global__ void collate_actions_kernel(const long long* __restrict__ prev_arena, const long long* __restrict__ act_arena,
                                       const long long* __restrict__ offsets, const int* __restrict__ lengths,
                                       long long* __restrict__ prev_out, long long* __restrict__ act_out,
                                       float* __restrict__ weights, unsigned char* __restrict__ masks, int T, int N,
                                       float coef) {
  const int total = T * N;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int t = idx / N, n = idx % N;
    const bool live = t < lengths[n];
    const long long o = offsets[n] + t;
    const long long a = live ? act_arena[o] : 0;
    prev_out[idx] = live ? prev_arena[o] : 0;
    act_out[idx] = a;
    float w = 0.f;
    if (live) w = (t == 0 || a != act_arena[o - 1]) ? coef : 1.f;
    weights[idx] = w;
    masks[idx] = t == 0 ? 0 : 1;
  }
}

}  // namespace
}  // namespace vlnce

using namespace vlnce;

extern "C" int vlnce_collate_rows(const void* arena, int dtype, const long long* offsets, const int* lengths, float* out,
                                  int T, int N, long long D, float pad, cudaStream_t stream) {
  VLNCE_CHECK_ARG(arena && offsets && lengths && out && T > 0 && N > 0 && D > 0, "collate_rows: bad arguments");
  // vector path: row size a multiple of 4 elements and 16-byte aligned bases; rows of a trajectory whose offset is not a
  // multiple of 4 elements fall back to scalar accesses inside the kernel
  const bool vec = D % 4 == 0 && (reinterpret_cast<uintptr_t>(arena) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(out) % 16 == 0) && dtype != VLNCE_I64 && dtype != VLNCE_U8;
  const long long per_thread = vec ? 4 : 1;
  const unsigned bx = (unsigned)std::min<long long>(cdiv(D, 256 * per_thread), 64);
  dim3 grid(bx, (unsigned)std::min<long long>((long long)T * N, 65535));
#define GO(TS)                                                                                                       \
  do {                                                                                                               \
    if (vec) collate_rows_kernel<TS, 4><<<grid, 256, 0, stream>>>((const TS*)arena, offsets, lengths, out, T, N, D, pad); \
    else collate_rows_kernel<TS, 1><<<grid, 256, 0, stream>>>((const TS*)arena, offsets, lengths, out, T, N, D, pad);     \
  } while (0)
  switch (dtype) {
    case VLNCE_F32: GO(float); break;
    case VLNCE_F16: GO(__half); break;
    case VLNCE_BF16: GO(__nv_bfloat16); break;
    case VLNCE_U8: collate_rows_kernel<unsigned char, 1><<<grid, 256, 0, stream>>>((const unsigned char*)arena, offsets, lengths, out, T, N, D, pad); break;
    case VLNCE_I64: collate_rows_kernel<long long, 1><<<grid, 256, 0, stream>>>((const long long*)arena, offsets, lengths, out, T, N, D, pad); break;
    default:
      set_error("collate_rows: unsupported dtype %d", dtype);
      return VLNCE_ERR_INVALID_ARGUMENT;
  }
#undef GO
  VLNCE_LAUNCH_CHECK("collate_rows_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_collate_actions(const long long* prev_arena, const long long* act_arena, const long long* offsets,
                                     const int* lengths, long long* prev_out, long long* act_out, float* weights,
                                     unsigned char* masks, int T, int N, float coef, cudaStream_t stream) {
  VLNCE_CHECK_ARG(prev_arena && act_arena && offsets && lengths && prev_out && act_out && weights && masks && T > 0 &&
                      N > 0,
                  "collate_actions: bad arguments");
  collate_actions_kernel<<<(unsigned)std::min<long long>(cdiv((long long)T * N, 256), 1184), 256, 0, stream>>>(prev_arena, act_arena, offsets, lengths,
                                                                             prev_out, act_out, weights, masks, T, N,
                                                                             coef);
  VLNCE_LAUNCH_CHECK("collate_actions_kernel");
  return VLNCE_OK;
}
