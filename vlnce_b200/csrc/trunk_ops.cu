// CUDA-core kernels around the tensor-core conv tiles: the fp32-accumulate reference conv
// (fp32-exact mode + on-device cross-check), stem operand builders, pooling, GroupNorm /
// train-mode BatchNorm statistics and the fused normalise(+residual)(+ReLU) pass.
// All HBM-bound: coalesced channel-fastest accesses, 16-byte vectors where the layout allows.
#include "common.cuh"

namespace vlnce {
namespace {

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&f)[8]) {
#pragma unroll
  for (int e = 0; e < 8; ++e) f[e] = to_float<T>(p[e]);
}
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__half>(const __half* p, float (&f)[8]) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
    f[2 * e] = h.x; f[2 * e + 1] = h.y;
  }
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    f[2 * e] = __uint_as_float(w[e] << 16);
    f[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
  }
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&f)[8]) {
#pragma unroll
  for (int e = 0; e < 8; ++e) p[e] = from_float<T>(f[e]);
}
template <>
__device__ __forceinline__ void store8<float>(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
template <>
__device__ __forceinline__ void store8<__half>(__half* p, const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __half2 h = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
    w[e] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    w[e] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}


// ------------------------------------------------------------------------------------
// Reference conv: one thread per output element, fp32 accumulation in K order (r, s, c).
// ------------------------------------------------------------------------------------
template <typename TX, typename TW, typename TY, typename TR>
__global__ void conv_ref_kernel(const TX* __restrict__ x, const TW* __restrict__ w,
                                const float* __restrict__ scale, const float* __restrict__ shift,
                                const TR* __restrict__ residual, TY* __restrict__ y, int N, int H,
                                int W, int C, int K, int R, int S, int stride, int pad, int P, int Q,
                                int relu) {
  const long long total = (long long)N * P * Q * K;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % K);
    const long long m = idx / K;
    const int q = (int)(m % Q);
    const int p = (int)((m / Q) % P);
    const int n = (int)(m / ((long long)P * Q));
    float acc = 0.f;
    for (int r = 0; r < R; ++r) {
      const int ih = p * stride + r - pad;
      if ((unsigned)ih >= (unsigned)H) continue;
      for (int s = 0; s < S; ++s) {
        const int iw = q * stride + s - pad;
        if ((unsigned)iw >= (unsigned)W) continue;
        const TX* xp = x + ((size_t)(n * H + ih) * W + iw) * C;
        const TW* wp = w + ((size_t)(k * R + r) * S + s) * C;
        for (int c = 0; c < C; ++c) acc = fmaf(to_float<TX>(xp[c]), to_float<TW>(wp[c]), acc);
      }
    }
    if (scale != nullptr) acc = fmaf(acc, scale[k], shift[k]);
    if (residual != nullptr) acc += to_float<TR>(residual[idx]);
    if (relu) acc = fmaxf(acc, 0.f);
    y[idx] = from_float<TY>(acc);
  }
}

// Warp-cooperative fp32 convolution for the layers of the tf32 mode that the TMA tiles do not cover (4x4 images,
// odd sizes): one warp owns PIXW output pixels x KW filters, the lanes split the input channels (coalesced 128-byte
// reads of both operands), 32 partial sums per lane are reduced with shuffles.  fp32 products and sums.
template <int PIXW, int KW>
__global__ void __launch_bounds__(256)
conv_warp_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale,
                 const float* __restrict__ shift, const float* __restrict__ residual, float* __restrict__ y, int N,
                 int H, int W, int C, int K, int R, int S, int stride, int pad, int P, int Q, int relu) {
  const int lane = threadIdx.x & 31;
  const long long M = (long long)N * P * Q;
  const long long mt = (M + PIXW - 1) / PIXW;
  const int kt = (K + KW - 1) / KW;
  const long long tasks = mt * kt;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long task = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; task < tasks; task += warps) {
    const int k0 = (int)(task % kt) * KW;           // neighbouring warps share the pixel tile (L1 reuse of x)
    const long long m0 = (task / kt) * PIXW;
    float acc[PIXW][KW];
#pragma unroll
    for (int i = 0; i < PIXW; ++i)
#pragma unroll
      for (int j = 0; j < KW; ++j) acc[i][j] = 0.f;
    int pn[PIXW], pp[PIXW], pq[PIXW];
#pragma unroll
    for (int i = 0; i < PIXW; ++i) {
      const long long m = m0 + i < M ? m0 + i : M - 1;
      pq[i] = (int)(m % Q);
      pp[i] = (int)((m / Q) % P);
      pn[i] = (int)(m / ((long long)P * Q));
    }
    for (int r = 0; r < R; ++r) {
      for (int s = 0; s < S; ++s) {
        const float* xp[PIXW];
        bool any = false;
#pragma unroll
        for (int i = 0; i < PIXW; ++i) {
          const int ih = pp[i] * stride + r - pad, iw = pq[i] * stride + s - pad;
          const bool ok = (unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W;
          xp[i] = ok ? x + ((size_t)(pn[i] * H + ih) * W + iw) * C : nullptr;
          any |= ok;
        }
        if (!any) continue;
        const float* wp = w + ((size_t)(k0 * R + r) * S + s) * C;
        const size_t wstride = (size_t)R * S * C;
        for (int c = lane; c < C; c += 32) {
          float xv[PIXW], wv[KW];
#pragma unroll
          for (int i = 0; i < PIXW; ++i) xv[i] = xp[i] != nullptr ? xp[i][c] : 0.f;
#pragma unroll
          for (int j = 0; j < KW; ++j) wv[j] = k0 + j < K ? __ldg(wp + j * wstride + c) : 0.f;
#pragma unroll
          for (int i = 0; i < PIXW; ++i)
#pragma unroll
            for (int j = 0; j < KW; ++j) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < PIXW; ++i)
#pragma unroll
      for (int j = 0; j < KW; ++j) {
        float v = acc[i][j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[i][j] = v;
      }
    // lane l writes output (pixel l / KW, filter l % KW) — PIXW * KW <= 32
#pragma unroll
    for (int i = 0; i < PIXW; ++i)
#pragma unroll
      for (int j = 0; j < KW; ++j) {
        if (lane == i * KW + j && m0 + i < M && k0 + j < K) {
          float v = acc[i][j];
          const int k = k0 + j;
          const size_t o = (size_t)(m0 + i) * K + k;
          if (scale != nullptr) v = fmaf(v, scale[k], shift[k]);
          if (residual != nullptr) v += residual[o];
          if (relu) v = fmaxf(v, 0.f);
          y[o] = v;
        }
      }
  }
}

// ------------------------------------------------------------------------------------
// Stem operand builders (explicit im2col of the two 7x7/2 stems only; every other conv is
// implicit).  One thread writes one 16-byte chunk (8 bf16) of an operand row.
// ------------------------------------------------------------------------------------
template <typename TX, typename TA>
__global__ void im2col_rgb_stem_kernel(const TX* __restrict__ x, TA* __restrict__ a,
                                       int N, int H, int W, int P, int Q) {
  constexpr int KP = 192, CHUNKS = KP / 8;
  const long long total = (long long)N * P * Q * CHUNKS;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(idx % CHUNKS);
    const long long m = idx / CHUNKS;
    const int q = (int)(m % Q);
    const int p = (int)((m / Q) % P);
    const int n = (int)(m / ((long long)P * Q));
    __align__(16) TA v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int kk = ch * 8 + e;
      float f = 0.f;
      if (kk < 147) {
        const int c = kk % 3, tap = kk / 3;
        const int s = tap % 7, r = tap / 7;
        const int ih = p * 2 + r - 3, iw = q * 2 + s - 3;
        if ((unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W)
          f = to_float<TX>(__ldg(x + ((size_t)(n * H + ih) * W + iw) * 3 + c));
      }
      v[e] = from_float<TA>(f);
    }
    *reinterpret_cast<uint4*>(a + m * KP + ch * 8) = *reinterpret_cast<const uint4*>(v);
  }
}

template <typename TA>
__global__ void im2col_depth_stem_kernel(const float* __restrict__ x, TA* __restrict__ a,
                                         int N, int H, int W, int P, int Q) {
  pdl_trigger();  // PDL (common.cuh)
  pdl_wait();
  // x: (N,H,W,1); pooled image is (H/2, W/2); output pixel (p,q) of the 7x7/2 pad-3 conv
  constexpr int KP = 64, CHUNKS = KP / 8;
  const int Hp = H / 2, Wp = W / 2;
  const long long total = (long long)N * P * Q * CHUNKS;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(idx % CHUNKS);
    const long long m = idx / CHUNKS;
    const int q = (int)(m % Q);
    const int p = (int)((m / Q) % P);
    const int n = (int)(m / ((long long)P * Q));
    __align__(16) TA v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int kk = ch * 8 + e;
      float f = 0.f;
      if (kk < 49) {
        const int s = kk % 7, r = kk / 7;
        const int ih = p * 2 + r - 3, iw = q * 2 + s - 3;
        if ((unsigned)ih < (unsigned)Hp && (unsigned)iw < (unsigned)Wp) {
          const float* b = x + ((size_t)(n * H + 2 * ih) * W + 2 * iw);
          f = ((__ldg(b) + __ldg(b + 1)) + (__ldg(b + W) + __ldg(b + W + 1))) * 0.25f;
        }
      }
      v[e] = from_float<TA>(f);
    }
    *reinterpret_cast<uint4*>(a + m * KP + ch * 8) = *reinterpret_cast<const uint4*>(v);
  }
}

__global__ void avgpool2x2_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H,
                                  int W, int C) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * Ho * Wo * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long long m = idx / C;
    const int q = (int)(m % Wo);
    const int p = (int)((m / Wo) % Ho);
    const int n = (int)(m / ((long long)Ho * Wo));
    const float* b = x + ((size_t)(n * H + 2 * p) * W + 2 * q) * C + c;
    y[idx] = ((b[0] + b[C]) + (b[(size_t)W * C] + b[(size_t)W * C + C])) * 0.25f;
  }
}

template <typename T>
__global__ void maxpool3x3s2_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W,
                                    int C, int P, int Q) {
  pdl_trigger();  // PDL (common.cuh)
  pdl_wait();
  // 8 channels per thread (C % 8 == 0): one 16-byte load per window tap
  const int CV = C / 8;
  const long long total = (long long)N * P * Q * CV;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % CV);
    const long long m = idx / CV;
    const int q = (int)(m % Q);
    const int p = (int)((m / Q) % P);
    const int n = (int)(m / ((long long)P * Q));
    float best[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) best[e] = -INFINITY;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ih = p * 2 + r - 1;
      if ((unsigned)ih >= (unsigned)H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int iw = q * 2 + s - 1;
        if ((unsigned)iw >= (unsigned)W) continue;
        float v[8];
        load8<T>(x + ((size_t)(n * H + ih) * W + iw) * C + cv * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) best[e] = fmaxf(best[e], v[e]);
      }
    }
    store8<T>(y + m * C + cv * 8, best);
  }
}

// ------------------------------------------------------------------------------------
// GroupNorm statistics: one CTA per (sample, group).  fp32 partials, fp64 combine.
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void groupnorm_stats_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float* __restrict__ scale,
                                       float* __restrict__ shift, int HW, int C, int groups, float eps) {
  const int n = blockIdx.x / groups, g = blockIdx.x % groups;
  const int cpg = C / groups;
  const long long count = (long long)HW * cpg;
  const T* base = x + (size_t)n * HW * C + g * cpg;
  float s = 0.f, ss = 0.f;
  for (long long e = threadIdx.x; e < count; e += blockDim.x) {
    const int pix = (int)(e / cpg), ch = (int)(e % cpg);
    const float v = to_float<T>(base[(size_t)pix * C + ch]);
    s += v;
    ss = fmaf(v, v, ss);
  }
  __shared__ double sh_s[32], sh_ss[32];
  double ds = s, dss = ss;
  for (int o = 16; o > 0; o >>= 1) {
    ds += __shfl_xor_sync(0xffffffffu, ds, o);
    dss += __shfl_xor_sync(0xffffffffu, dss, o);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) { sh_s[wid] = ds; sh_ss[wid] = dss; }
  __syncthreads();
  if (wid == 0) {
    ds = lane < nw ? sh_s[lane] : 0.0;
    dss = lane < nw ? sh_ss[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
      ds += __shfl_xor_sync(0xffffffffu, ds, o);
      dss += __shfl_xor_sync(0xffffffffu, dss, o);
    }
    if (lane == 0) { sh_s[0] = ds; sh_ss[0] = dss; }
  }
  __syncthreads();
  const double mean = sh_s[0] / (double)count;
  double var = sh_ss[0] / (double)count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  for (int ch = threadIdx.x; ch < cpg; ch += blockDim.x) {
    const int c = g * cpg + ch;
    const float sc = rstd * gamma[c];
    scale[(size_t)n * C + c] = sc;
    shift[(size_t)n * C + c] = beta[c] - (float)mean * sc;
  }
}

// Coalesced two-stage GroupNorm statistics for fp32 raw conv output (C % 4 == 0, C/4 <= 256,
// 256 % (C/4) == 0): stage 1 = grid (N, S) CTAs stream float4 rows and bin (sum, sumsq) per
// group in shared memory; stage 2 combines the S slices in fp64 and writes scale/shift.
template <typename TI>
__device__ __forceinline__ float4 load4f(const TI* p);
template <>
__device__ __forceinline__ float4 load4f<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <>
__device__ __forceinline__ float4 load4f<__half>(const __half* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

template <typename TI>
__global__ void __launch_bounds__(256)
gn_partial_kernel(const TI* __restrict__ x, float* __restrict__ part /*[N][S][G][2]*/, int HW, int C,
                  int groups) {
  pdl_trigger();  // PDL (common.cuh)
  pdl_wait();
  __shared__ float bins[2 * 64];
  const int n = blockIdx.x, s = blockIdx.y, S = gridDim.y;
  const int c4n = C / 4;
  const int lane_c = threadIdx.x % c4n, lane_r = threadIdx.x / c4n, rstep = 256 / c4n;
  const int cpg = C / groups;
  for (int i = threadIdx.x; i < 2 * groups; i += 256) bins[i] = 0.f;
  __syncthreads();
  const int rows_per = (HW + S - 1) / S;
  const int r0 = s * rows_per, r1 = min(HW, r0 + rows_per);
  float s01 = 0.f, q01 = 0.f, s23 = 0.f, q23 = 0.f;
  const TI* base = x + (size_t)n * HW * C + lane_c * 4;
  for (int r = r0 + lane_r; r < r1; r += rstep) {
    const float4 v = load4f<TI>(base + (size_t)r * C);
    s01 += v.x + v.y; q01 = fmaf(v.x, v.x, fmaf(v.y, v.y, q01));
    s23 += v.z + v.w; q23 = fmaf(v.z, v.z, fmaf(v.w, v.w, q23));
  }
  const int g01 = (lane_c * 4) / cpg, g23 = (lane_c * 4 + 2) / cpg;
  if (g01 == g23) {
    atomicAdd(&bins[2 * g01], s01 + s23);
    atomicAdd(&bins[2 * g01 + 1], q01 + q23);
  } else {
    atomicAdd(&bins[2 * g01], s01); atomicAdd(&bins[2 * g01 + 1], q01);
    atomicAdd(&bins[2 * g23], s23); atomicAdd(&bins[2 * g23 + 1], q23);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * groups; i += 256)
    part[((size_t)(n * S + s) * groups) * 2 + i] = bins[i];
}

__global__ void gn_finalize_kernel(const float* __restrict__ part, int S, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ scale,
                                   float* __restrict__ shift, int HW, int C, int groups, float eps) {
  const int n = blockIdx.x;
  const int cpg = C / groups;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    double sum = 0.0, sq = 0.0;
    for (int s = 0; s < S; ++s) {
      const float* p = part + ((size_t)(n * S + s) * groups + g) * 2;
      sum += p[0];
      sq += p[1];
    }
    const double cnt = (double)HW * cpg;
    const double mean = sum / cnt;
    double var = sq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = rstd * gamma[c];
    scale[(size_t)n * C + c] = sc;
    shift[(size_t)n * C + c] = beta[c] - (float)mean * sc;
  }
}

// ------------------------------------------------------------------------------------
// Train-mode BatchNorm statistics: per-channel sums over all rows.
// Pass 1: grid (row splits) accumulates fp64 partial sums with atomics into workspace[2*C]
// (zeroed by the caller-visible memset below); pass 2 finalises + updates running stats.
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void bn_partial_kernel(const T* __restrict__ x, double* __restrict__ ws, long long rows,
                                  int C) {
  // blockDim.x threads walk channels (coalesced), blockIdx.y splits rows
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const long long per = (rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = per * blockIdx.y, r1 = min(rows, r0 + per);
  float s = 0.f, ss = 0.f;
  double ds = 0.0, dss = 0.0;
  int cnt = 0;
  for (long long r = r0; r < r1; ++r) {
    const float v = to_float<T>(x[(size_t)r * C + c]);
    s += v;
    ss = fmaf(v, v, ss);
    if (++cnt == 256) { ds += s; dss += ss; s = 0.f; ss = 0.f; cnt = 0; }
  }
  ds += s; dss += ss;
  atomicAdd(&ws[c], ds);
  atomicAdd(&ws[C + c], dss);
}

__global__ void bn_finalize_kernel(const double* __restrict__ ws, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float momentum,
                                   float* __restrict__ scale, float* __restrict__ shift,
                                   long long rows, int C, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = ws[c] / (double)rows;
  double var = ws[C + c] / (double)rows - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = rstd * gamma[c];
  scale[c] = sc;
  shift[c] = beta[c] - (float)mean * sc;
  if (running_mean != nullptr) {
    const double unbiased = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// ------------------------------------------------------------------------------------
// y = act(x * scale[n|0, c] + shift (+ residual)); 8 channels (16 B of bf16) per thread.
// ------------------------------------------------------------------------------------
template <typename T, typename TO>
__global__ void affine_apply_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                                    const float* __restrict__ shift, int scale_stride_n,
                                    const TO* __restrict__ residual, TO* __restrict__ y, int HW, int C,
                                    long long total_vec, int relu) {
  constexpr int V = 8;  // 8 channels per thread: 16/32-byte vector accesses
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total_vec;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long e0 = idx * V;
    const int c0 = (int)(e0 % C);
    const long long row = e0 / C;
    const int n = (int)(row / HW);
    float f[V], sc[V], sh[V];
    load8<T>(x + e0, f);
    load8<float>(scale + (size_t)n * scale_stride_n + c0, sc);
    load8<float>(shift + (size_t)n * scale_stride_n + c0, sh);
#pragma unroll
    for (int e = 0; e < V; ++e) f[e] = fmaf(f[e], sc[e], sh[e]);
    if (residual != nullptr) {
      float r[V];
      load8<TO>(residual + e0, r);
#pragma unroll
      for (int e = 0; e < V; ++e) f[e] += r[e];
    }
    if (relu) {
#pragma unroll
      for (int e = 0; e < V; ++e) f[e] = fmaxf(f[e], 0.f);
    }
    store8<TO>(y + e0, f);
  }
}

// GroupNorm finalize + apply in one pass: every CTA first rebuilds the per-channel scale/shift of
// its image from the stage-1 partial sums (S x G pairs, fp64 combine), then normalises its slice
// of rows: y = act(raw * scale[c] + shift[c] (+ residual)).  grid (N, slices).
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
gn_apply_kernel(const TI* __restrict__ x, const float* __restrict__ part, int S,
                const float* __restrict__ gamma, const float* __restrict__ beta, int groups, float eps,
                const TO* __restrict__ residual, TO* __restrict__ y, int HW, int C, int relu) {
  pdl_trigger();  // PDL (common.cuh)
  pdl_wait();
  extern __shared__ float gn_sm[];
  float* sc = gn_sm;
  float* sh = gn_sm + C;
  __shared__ float g_mean[64], g_rstd[64];
  const int n = blockIdx.x;
  const int cpg = C / groups;
  // one thread per GROUP combines the S slices (fp64), then every channel takes its group's moments
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    double sum = 0.0, sq = 0.0;
    const float2* p = reinterpret_cast<const float2*>(part) + (size_t)n * S * groups + g;
    for (int s = 0; s < S; ++s) {
      const float2 v = __ldg(p + (size_t)s * groups);
      sum += v.x;
      sq += v.y;
    }
    const double cnt = (double)HW * cpg;
    const double mean = sum / cnt;
    double var = sq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    g_mean[g] = (float)mean;
    g_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float a = g_rstd[g] * gamma[c];
    sc[c] = a;
    sh[c] = beta[c] - g_mean[g] * a;
  }
  __syncthreads();
  const int cv = C / 8;
  const long long vec_per_img = (long long)HW * cv;
  const long long per = (vec_per_img + gridDim.y - 1) / gridDim.y;
  const long long v0 = per * blockIdx.y, v1 = min(vec_per_img, v0 + per);
  for (long long v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
    const int c0 = (int)(v % cv) * 8;
    const long long e0 = ((long long)n * HW) * C + v * 8;
    float f[8];
    load8<TI>(x + e0, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = fmaf(f[e], sc[c0 + e], sh[c0 + e]);
    if (residual != nullptr) {
      float r[8];
      load8<TO>(residual + e0, r);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] += r[e];
    }
    if (relu & 1) {
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
    }
    if (relu & 2) {   // fp32 activations feeding kind::tf32 tiles: round to nearest TF32 (the MMA would truncate)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        uint32_t t;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(f[e]));
        f[e] = __uint_as_float(t);
      }
    }
    store8<TO>(y + e0, f);
  }
}

// adaptive_avg_pool2d -> (O,O): bin i covers [floor(i*H/O), ceil((i+1)*H/O))
template <typename T>
__global__ void adaptive_avgpool_kernel(const T* __restrict__ x, float* __restrict__ y, int ldy,
                                        int N, int H, int W, int C, int O) {
  pdl_trigger();  // PDL (common.cuh)
  pdl_wait();
  const int OO = O * O;
  const long long total = (long long)N * OO * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int pos = (int)((idx / C) % OO);
    const int n = (int)(idx / ((long long)OO * C));
    const int oi = pos / O, oj = pos % O;
    const int h0 = (oi * H) / O, h1 = ((oi + 1) * H + O - 1) / O;
    const int w0 = (oj * W) / O, w1 = ((oj + 1) * W + O - 1) / O;
    float acc = 0.f;
    for (int ih = h0; ih < h1; ++ih)
      for (int iw = w0; iw < w1; ++iw) acc += to_float<T>(x[((size_t)(n * H + ih) * W + iw) * C + c]);
    y[((size_t)n * OO + pos) * ldy + c] = acc / (float)((h1 - h0) * (w1 - w0));
  }
}

// Same, eight channels per thread (16-byte loads of the 16-bit activations, 2 x 16-byte fp32 stores): C % 8 == 0,
// ldy % 4 == 0 and 16-byte aligned tensors.  Same summation order per channel as the scalar kernel.
template <typename T>
__global__ void adaptive_avgpool_vec8_kernel(const T* __restrict__ x, float* __restrict__ y, int ldy,
                                             int N, int H, int W, int C, int O) {
  pdl_trigger();  // PDL (common.cuh)
  pdl_wait();
  const int OO = O * O, C8 = C / 8;
  const int total = N * OO * C8;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int c = (idx % C8) * 8;
    const int pos = (idx / C8) % OO;
    const int n = idx / (OO * C8);
    const int oi = pos / O, oj = pos % O;
    const int h0 = (oi * H) / O, h1 = ((oi + 1) * H + O - 1) / O;
    const int w0 = (oj * W) / O, w1 = ((oj + 1) * W + O - 1) / O;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int ih = h0; ih < h1; ++ih)
      for (int iw = w0; iw < w1; ++iw) {
        float f[8];
        load8<T>(x + ((size_t)(n * H + ih) * W + iw) * C + c, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e];
      }
    const float d = (float)((h1 - h0) * (w1 - w0));
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = acc[e] / d;
    store8<float>(y + ((size_t)n * OO + pos) * ldy + c, acc);
  }
}

// (N, C, P) -> (N, P, ldy) slice.  32x32 smem transpose per (n, tile).
__global__ void nchw_to_positions_kernel(const float* __restrict__ x, float* __restrict__ y, int ldy,
                                         int C, int P) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int c0 = blockIdx.x * 32, p0 = blockIdx.y * 32;
  const float* xb = x + (size_t)n * C * P;
  float* yb = y + (size_t)n * P * ldy;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < P) ? xb[(size_t)c * P + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (c < C && p < P) yb[(size_t)p * ldy + c] = tile[threadIdx.x][i];
  }
}

inline unsigned grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148ll * 32;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

template <typename TX, typename TW>
int conv_ref_dispatch_y(const void* x, const void* w, const float* scale, const float* shift,
                        const void* residual, int res_dtype, void* y, int y_dtype, int N, int H, int W,
                        int C, int K, int R, int S, int stride, int pad, int P, int Q, int relu,
                        cudaStream_t st) {
  const long long total = (long long)N * P * Q * K;
  const unsigned grid = grid_for(total, 256);
#define GO(TY, TR)                                                                              \
  conv_ref_kernel<TX, TW, TY, TR><<<grid, 256, 0, st>>>(                                        \
      (const TX*)x, (const TW*)w, scale, shift, (const TR*)residual, (TY*)y, N, H, W, C, K, R, S, \
      stride, pad, P, Q, relu)
  const bool rb = residual != nullptr && res_dtype == VLNCE_BF16;
  if (residual != nullptr && res_dtype == VLNCE_F16) {
    if (y_dtype != VLNCE_F32) {
      set_error("conv2d_ref: f16 residual needs f32 output");
      return VLNCE_ERR_INVALID_ARGUMENT;
    }
    GO(float, __half);
  } else if (y_dtype == VLNCE_F32) {
    if (rb) GO(float, __nv_bfloat16); else GO(float, float);
  } else {
    if (rb || residual == nullptr) GO(__nv_bfloat16, __nv_bfloat16); else GO(__nv_bfloat16, float);
  }
#undef GO
  VLNCE_LAUNCH_CHECK("conv_ref_kernel");
  return VLNCE_OK;
}

}  // namespace
}  // namespace vlnce

using namespace vlnce;

extern "C" int vlnce_conv2d_ref(const void* x, int x_dtype, const void* w, int w_dtype,
                                const float* scale, const float* shift, const void* residual,
                                int res_dtype, void* y, int y_dtype, int N, int H, int W, int C, int K,
                                int R, int S, int stride, int pad, int relu, cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && w && y, "conv2d_ref: null tensor");
  VLNCE_CHECK_ARG((scale == nullptr) == (shift == nullptr), "conv2d_ref: scale/shift must come together");
  VLNCE_CHECK_ARG(y_dtype == VLNCE_F32 || y_dtype == VLNCE_BF16, "conv2d_ref: y dtype");
  VLNCE_CHECK_ARG(w_dtype == VLNCE_F32 || w_dtype == VLNCE_BF16 || w_dtype == VLNCE_F16, "conv2d_ref: w dtype");
  if (w_dtype == VLNCE_F16) {
    VLNCE_CHECK_ARG(x_dtype == VLNCE_F16, "conv2d_ref: f16 weights need f16 input");
    const int P_ = (H + 2 * pad - R) / stride + 1, Q_ = (W + 2 * pad - S) / stride + 1;
    return conv_ref_dispatch_y<__half, __half>(x, w, scale, shift, residual, res_dtype, y, y_dtype, N, H, W, C, K,
                                               R, S, stride, pad, P_, Q_, relu, stream);
  }
  const int P = (H + 2 * pad - R) / stride + 1, Q = (W + 2 * pad - S) / stride + 1;
  VLNCE_CHECK_ARG(P > 0 && Q > 0, "conv2d_ref: empty output");
#define ARGS x, w, scale, shift, residual, res_dtype, y, y_dtype, N, H, W, C, K, R, S, stride, pad, P, Q, relu, stream
  if (w_dtype == VLNCE_F32) {
    if (x_dtype == VLNCE_F32) return conv_ref_dispatch_y<float, float>(ARGS);
    if (x_dtype == VLNCE_U8) return conv_ref_dispatch_y<uint8_t, float>(ARGS);
    if (x_dtype == VLNCE_BF16) return conv_ref_dispatch_y<__nv_bfloat16, float>(ARGS);
  } else {
    if (x_dtype == VLNCE_F32) return conv_ref_dispatch_y<float, __nv_bfloat16>(ARGS);
    if (x_dtype == VLNCE_U8) return conv_ref_dispatch_y<uint8_t, __nv_bfloat16>(ARGS);
    if (x_dtype == VLNCE_BF16) return conv_ref_dispatch_y<__nv_bfloat16, __nv_bfloat16>(ARGS);
  }
#undef ARGS
  set_error("conv2d_ref: unsupported x dtype %d", x_dtype);
  return VLNCE_ERR_INVALID_ARGUMENT;
}

extern "C" int vlnce_conv2d_warp(const float* x, const float* w, const float* scale, const float* shift,
                                 const float* residual, float* y, int N, int H, int W, int C, int K, int R, int S,
                                 int stride, int pad, int relu, cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && w && y, "conv2d_warp: null tensor");
  VLNCE_CHECK_ARG((scale == nullptr) == (shift == nullptr), "conv2d_warp: scale/shift must come together");
  VLNCE_CHECK_ARG(N > 0 && C > 0 && K > 0 && R > 0 && S > 0 && stride > 0 && pad >= 0, "conv2d_warp: sizes");
  const int P = (H + 2 * pad - R) / stride + 1, Q = (W + 2 * pad - S) / stride + 1;
  VLNCE_CHECK_ARG(P > 0 && Q > 0, "conv2d_warp: empty output");
  constexpr int PIXW = 8, KW = 4;
  const long long M = (long long)N * P * Q;
  const long long tasks = ((M + PIXW - 1) / PIXW) * ((K + KW - 1) / KW);
  const long long blocks = (tasks + 7) / 8;   // 8 warps per block
  const unsigned grid = (unsigned)(blocks < 1 ? 1 : (blocks > 148ll * 64 ? 148ll * 64 : blocks));
  conv_warp_kernel<PIXW, KW><<<grid, 256, 0, stream>>>(x, w, scale, shift, residual, y, N, H, W, C, K, R, S, stride,
                                                        pad, P, Q, relu);
  VLNCE_LAUNCH_CHECK("conv_warp_kernel");
  return VLNCE_OK;
}

// ------------------------------------------------------------------------------------
// MODEL.normalize_rgb = True (resnet_encoders.py:171-192): (x/255 - mean_c) / std_c.  The per-channel 1/std is
// folded into the stem weights' input channels and the 1/255 into the stem epilogue scale, so only the mean has
// to be removed ahead of the zero-padded stem conv: y = x - 255 mean_c (fp32; the padding stays an exact 0 as in
// the reference, where it is applied after the normalisation).
// ------------------------------------------------------------------------------------
template <typename TX>
__global__ void rgb_center_kernel(const TX* __restrict__ x, float* __restrict__ y, float m0, float m1, float m2,
                                  long long pixels) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pixels;
       i += (long long)gridDim.x * blockDim.x) {
    const TX* p = x + 3 * i;
    float* o = y + 3 * i;
    o[0] = (float)p[0] - m0;
    o[1] = (float)p[1] - m1;
    o[2] = (float)p[2] - m2;
  }
}

extern "C" int vlnce_rgb_center(const void* x, int x_dtype, float* y, float m0, float m1, float m2, long long pixels,
                                cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && y && pixels > 0, "rgb_center: bad arguments");
  VLNCE_CHECK_ARG(x_dtype == VLNCE_U8 || x_dtype == VLNCE_F32, "rgb_center: rgb must be u8 or f32");
  const unsigned grid = grid_for(pixels, 256);
  if (x_dtype == VLNCE_U8) rgb_center_kernel<uint8_t><<<grid, 256, 0, stream>>>((const uint8_t*)x, y, m0, m1, m2, pixels);
  else rgb_center_kernel<float><<<grid, 256, 0, stream>>>((const float*)x, y, m0, m1, m2, pixels);
  VLNCE_LAUNCH_CHECK("rgb_center_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_im2col_rgb_stem(const void* x, int x_dtype, void* a, int a_dtype, int N, int H, int W,
                                     cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && a && N > 0 && H > 0 && W > 0, "im2col_rgb_stem: bad arguments");
  VLNCE_CHECK_ARG(a_dtype == VLNCE_BF16 || a_dtype == VLNCE_F16, "im2col_rgb_stem: operand dtype must be bf16 or f16");
  VLNCE_CHECK_ARG(x_dtype == VLNCE_U8 || x_dtype == VLNCE_F32, "im2col_rgb_stem: rgb must be u8 or f32");
  const int P = (H + 6 - 7) / 2 + 1, Q = (W + 6 - 7) / 2 + 1;
  const long long total = (long long)N * P * Q * 24;
  const unsigned grid = grid_for(total, 256);
#define GO(TX, TA) im2col_rgb_stem_kernel<TX, TA><<<grid, 256, 0, stream>>>((const TX*)x, (TA*)a, N, H, W, P, Q)
  if (x_dtype == VLNCE_U8) {
    if (a_dtype == VLNCE_BF16) GO(uint8_t, __nv_bfloat16); else GO(uint8_t, __half);
  } else {
    if (a_dtype == VLNCE_BF16) GO(float, __nv_bfloat16); else GO(float, __half);
  }
#undef GO
  VLNCE_LAUNCH_CHECK("im2col_rgb_stem_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_im2col_depth_stem(const float* x, void* a, int a_dtype, int N, int H, int W,
                                       cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && a && N > 0 && H > 1 && W > 1 && H % 2 == 0 && W % 2 == 0,
                  "im2col_depth_stem: bad arguments");
  const int Hp = H / 2, Wp = W / 2;
  const int P = (Hp + 6 - 7) / 2 + 1, Q = (Wp + 6 - 7) / 2 + 1;
  const long long total = (long long)N * P * Q * 8;
  if (a_dtype == VLNCE_F16)
    im2col_depth_stem_kernel<__half><<<grid_for(total, 256), 256, 0, stream>>>(x, (__half*)a, N, H, W, P, Q);
  else if (a_dtype == VLNCE_BF16)
    im2col_depth_stem_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, stream>>>(x, (__nv_bfloat16*)a, N, H,
                                                                                       W, P, Q);
  else {
    set_error("im2col_depth_stem: operand dtype must be bf16 or f16");
    return VLNCE_ERR_INVALID_ARGUMENT;
  }
  VLNCE_LAUNCH_CHECK("im2col_depth_stem_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_avgpool2x2(const float* x, float* y, int N, int H, int W, int C,
                                cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && y && H % 2 == 0 && W % 2 == 0, "avgpool2x2: bad arguments");
  const long long total = (long long)N * (H / 2) * (W / 2) * C;
  avgpool2x2_kernel<<<grid_for(total, 256), 256, 0, stream>>>(x, y, N, H, W, C);
  VLNCE_LAUNCH_CHECK("avgpool2x2_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_maxpool3x3s2(const void* x, void* y, int dtype, int N, int H, int W, int C,
                                  cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && y && C % 8 == 0, "maxpool3x3s2: null tensor or C %% 8 != 0");
  const int P = (H + 2 - 3) / 2 + 1, Q = (W + 2 - 3) / 2 + 1;
  const long long total = (long long)N * P * Q * (C / 8);
  if (dtype == VLNCE_BF16)
    maxpool3x3s2_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, stream>>>(
        (const __nv_bfloat16*)x, (__nv_bfloat16*)y, N, H, W, C, P, Q);
  else if (dtype == VLNCE_F16)
    maxpool3x3s2_kernel<__half><<<grid_for(total, 256), 256, 0, stream>>>((const __half*)x, (__half*)y, N, H, W, C,
                                                                          P, Q);
  else if (dtype == VLNCE_F32)
    maxpool3x3s2_kernel<float><<<grid_for(total, 256), 256, 0, stream>>>((const float*)x, (float*)y, N,
                                                                         H, W, C, P, Q);
  else {
    set_error("maxpool3x3s2: dtype");
    return VLNCE_ERR_INVALID_ARGUMENT;
  }
  VLNCE_LAUNCH_CHECK("maxpool3x3s2_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_groupnorm_stats(const void* x, int dtype, const float* gamma, const float* beta,
                                     float* scale, float* shift, float* workspace, size_t workspace_bytes,
                                     int N, int HW, int C, int groups, float eps, cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && gamma && beta && scale && shift && groups > 0 && C % groups == 0,
                  "groupnorm_stats: bad arguments");
  const int c4n = C / 4;
  if (dtype == VLNCE_F32 && workspace != nullptr && C % 4 == 0 && c4n <= 256 && 256 % c4n == 0 &&
      (C / groups) % 2 == 0 && groups <= 64) {
    int S = cdiv((long long)HW * C, 4096);  // ~16 KB of raw data per CTA: enough CTAs to cover the SMs
    if (S > HW) S = HW;
    while (S > 1 && (size_t)N * S * groups * 2 * sizeof(float) > workspace_bytes) --S;
    if ((size_t)N * S * groups * 2 * sizeof(float) <= workspace_bytes) {
      gn_partial_kernel<float><<<dim3(N, S), 256, 0, stream>>>((const float*)x, workspace, HW, C, groups);
      VLNCE_LAUNCH_CHECK("gn_partial_kernel");
      gn_finalize_kernel<<<N, 256, 0, stream>>>(workspace, S, gamma, beta, scale, shift, HW, C, groups, eps);
      VLNCE_LAUNCH_CHECK("gn_finalize_kernel");
      return VLNCE_OK;
    }
  }
  if (dtype == VLNCE_BF16)
    groupnorm_stats_kernel<__nv_bfloat16><<<N * groups, 256, 0, stream>>>(
        (const __nv_bfloat16*)x, gamma, beta, scale, shift, HW, C, groups, eps);
  else if (dtype == VLNCE_F32)
    groupnorm_stats_kernel<float><<<N * groups, 256, 0, stream>>>((const float*)x, gamma, beta, scale,
                                                                  shift, HW, C, groups, eps);
  else {
    set_error("groupnorm_stats: dtype");
    return VLNCE_ERR_INVALID_ARGUMENT;
  }
  VLNCE_LAUNCH_CHECK("groupnorm_stats_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_batchnorm_train_stats(const void* x, int dtype, const float* gamma,
                                           const float* beta, float* running_mean, float* running_var,
                                           float momentum, float* scale, float* shift,
                                           double* workspace, long long rows, int C, float eps,
                                           cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && gamma && beta && scale && shift && workspace && rows > 0 && C > 0,
                  "batchnorm_train_stats: bad arguments");
  cudaError_t e = cudaMemsetAsync(workspace, 0, sizeof(double) * 2 * C, stream);
  if (e != cudaSuccess) return cuda_status(e, "batchnorm_train_stats memset");
  const int splits = (int)(rows / 512 < 1 ? 1 : (rows / 512 > 512 ? 512 : rows / 512));
  dim3 grid(cdiv(C, 128), splits);
  if (dtype == VLNCE_BF16)
    bn_partial_kernel<__nv_bfloat16><<<grid, 128, 0, stream>>>((const __nv_bfloat16*)x, workspace, rows, C);
  else if (dtype == VLNCE_F32)
    bn_partial_kernel<float><<<grid, 128, 0, stream>>>((const float*)x, workspace, rows, C);
  else {
    set_error("batchnorm_train_stats: dtype");
    return VLNCE_ERR_INVALID_ARGUMENT;
  }
  VLNCE_LAUNCH_CHECK("bn_partial_kernel");
  bn_finalize_kernel<<<cdiv(C, 128), 128, 0, stream>>>(workspace, gamma, beta, running_mean, running_var,
                                                      momentum, scale, shift, rows, C, eps);
  VLNCE_LAUNCH_CHECK("bn_finalize_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_affine_apply(const void* x, int x_dtype, const float* scale, const float* shift,
                                  int scale_stride_n, const void* residual, void* y, int y_dtype, int N,
                                  int HW, int C, int relu, cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && y && scale && shift && C % 8 == 0, "affine_apply: bad arguments (C %% 8)");
  const long long total_vec = (long long)N * HW * C / 8;
  const unsigned grid = grid_for(total_vec, 256);
  if (x_dtype == VLNCE_BF16 && y_dtype == VLNCE_BF16)
    affine_apply_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, stream>>>(
        (const __nv_bfloat16*)x, scale, shift, scale_stride_n, (const __nv_bfloat16*)residual,
        (__nv_bfloat16*)y, HW, C, total_vec, relu);
  else if (x_dtype == VLNCE_F32 && y_dtype == VLNCE_BF16)
    affine_apply_kernel<float, __nv_bfloat16><<<grid, 256, 0, stream>>>(
        (const float*)x, scale, shift, scale_stride_n, (const __nv_bfloat16*)residual, (__nv_bfloat16*)y,
        HW, C, total_vec, relu);
  else if (x_dtype == VLNCE_F32 && y_dtype == VLNCE_F16)
    affine_apply_kernel<float, __half><<<grid, 256, 0, stream>>>(
        (const float*)x, scale, shift, scale_stride_n, (const __half*)residual, (__half*)y, HW, C, total_vec,
        relu);
  else if (x_dtype == VLNCE_F32 && y_dtype == VLNCE_F32)
    affine_apply_kernel<float, float><<<grid, 256, 0, stream>>>(
        (const float*)x, scale, shift, scale_stride_n, (const float*)residual, (float*)y, HW, C,
        total_vec, relu);
  else {
    set_error("affine_apply: unsupported dtype pair %d -> %d", x_dtype, y_dtype);
    return VLNCE_ERR_INVALID_ARGUMENT;
  }
  VLNCE_LAUNCH_CHECK("affine_apply_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_adaptive_avgpool(const void* x, int dtype, float* y, int ldy, int N, int H, int W,
                                      int C, int O, cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && y && ldy >= C && (O == 1 || O == 4), "adaptive_avgpool: bad arguments");
  const long long total = (long long)N * O * O * C;
  const bool vec = C % 8 == 0 && ldy % 4 == 0 && total < (1ll << 31) && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (dtype == VLNCE_BF16 || dtype == VLNCE_F16);
  if (vec && dtype == VLNCE_BF16)
    adaptive_avgpool_vec8_kernel<__nv_bfloat16><<<grid_for(total / 8, 256), 256, 0, stream>>>(
        (const __nv_bfloat16*)x, y, ldy, N, H, W, C, O);
  else if (vec)
    adaptive_avgpool_vec8_kernel<__half><<<grid_for(total / 8, 256), 256, 0, stream>>>((const __half*)x, y, ldy, N, H, W,
                                                                                     C, O);
  else if (dtype == VLNCE_BF16)
    adaptive_avgpool_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, stream>>>(
        (const __nv_bfloat16*)x, y, ldy, N, H, W, C, O);
  else if (dtype == VLNCE_F16)
    adaptive_avgpool_kernel<__half><<<grid_for(total, 256), 256, 0, stream>>>((const __half*)x, y, ldy, N, H, W, C,
                                                                              O);
  else if (dtype == VLNCE_F32)
    adaptive_avgpool_kernel<float><<<grid_for(total, 256), 256, 0, stream>>>((const float*)x, y, ldy, N,
                                                                             H, W, C, O);
  else {
    set_error("adaptive_avgpool: dtype");
    return VLNCE_ERR_INVALID_ARGUMENT;
  }
  VLNCE_LAUNCH_CHECK("adaptive_avgpool_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_nchw_to_positions(const float* x, float* y, int ldy, int N, int C, int P,
                                       cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && y && ldy >= C, "nchw_to_positions: bad arguments");
  dim3 grid(cdiv(C, 32), cdiv(P, 32), N), block(32, 8);
  nchw_to_positions_kernel<<<grid, block, 0, stream>>>(x, y, ldy, C, P);
  VLNCE_LAUNCH_CHECK("nchw_to_positions_kernel");
  return VLNCE_OK;
}


// GroupNorm(+residual)(+ReLU) of fp32 raw conv output in two launches (partial statistics, then
// finalize+apply); falls back to the three-kernel path when the layout preconditions fail.
extern "C" int vlnce_groupnorm_fused(const float* x, const float* gamma, const float* beta, const void* residual,
                                     void* y, int y_dtype, float* workspace, size_t workspace_bytes,
                                     float* scale_scratch, float* shift_scratch, int N, int HW, int C, int groups,
                                     float eps, int relu, cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && gamma && beta && y && workspace && groups > 0 && C % groups == 0 && C % 8 == 0,
                  "groupnorm_fused: bad arguments");
  const int c4n = C / 4;
  const bool fast = c4n <= 256 && 256 % c4n == 0 && (C / groups) % 2 == 0 && groups <= 64 &&
                    (y_dtype == VLNCE_F16 || y_dtype == VLNCE_BF16 || y_dtype == VLNCE_F32);
  if (!fast) {
    int rc = vlnce_groupnorm_stats(x, VLNCE_F32, gamma, beta, scale_scratch, shift_scratch, workspace,
                                   workspace_bytes, N, HW, C, groups, eps, stream);
    if (rc != VLNCE_OK) return rc;
    return vlnce_affine_apply(x, VLNCE_F32, scale_scratch, shift_scratch, C, residual, y, y_dtype, N, HW, C, relu & 1,
                              stream);
  }
  int S = cdiv((long long)HW * C, 4096);
  if (S > HW) S = HW;
  if (S > 8) S = 8;  // N x S CTAs already fill the chip at N >= 19; fewer slices = a shorter combine in apply
  while (S > 1 && (size_t)N * S * groups * 2 * sizeof(float) > workspace_bytes) --S;
  VLNCE_CHECK_ARG((size_t)N * S * groups * 2 * sizeof(float) <= workspace_bytes, "groupnorm_fused: workspace too small");
  launch_pdl(gn_partial_kernel<float>, dim3(N, S), dim3(256), 0, stream, x, workspace, HW, C, groups);
  VLNCE_LAUNCH_CHECK("gn_partial_kernel");
  int slices = cdiv((long long)HW * C, 8192);
  if (slices > 64) slices = 64;
  const size_t smem = sizeof(float) * 2 * C;
  dim3 grid(N, slices);
  if (y_dtype == VLNCE_F16)
    launch_pdl(gn_apply_kernel<float, __half>, grid, dim3(256), smem, stream, x, (const float*)workspace, S, gamma, beta, groups,
               eps, (const __half*)residual, (__half*)y, HW, C, relu);
  else if (y_dtype == VLNCE_BF16)
    launch_pdl(gn_apply_kernel<float, __nv_bfloat16>, grid, dim3(256), smem, stream, x, (const float*)workspace, S, gamma, beta,
               groups, eps, (const __nv_bfloat16*)residual, (__nv_bfloat16*)y, HW, C, relu);
  else
    launch_pdl(gn_apply_kernel<float, float>, grid, dim3(256), smem, stream, x, (const float*)workspace, S, gamma, beta, groups,
               eps, (const float*)residual, (float*)y, HW, C, relu);
  VLNCE_LAUNCH_CHECK("gn_apply_kernel");
  return VLNCE_OK;
}


// Same two-launch GroupNorm for an fp16 raw conv output (x and y fp16): halves the bytes the statistics and the
// apply pass stream compared with the fp32 raw tensor (the depth trunk is bound by that traffic).  The statistics
// are those of the fp16-rounded values (fp32 bins, fp64 combine).  Layout preconditions as the fast path above.
extern "C" int vlnce_groupnorm_fused16(const void* x, const float* gamma, const float* beta, const void* residual,
                                       void* y, float* workspace, size_t workspace_bytes, int N, int HW, int C,
                                       int groups, float eps, int relu, cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && gamma && beta && y && workspace && groups > 0 && C % groups == 0 && C % 8 == 0,
                  "groupnorm_fused16: bad arguments");
  const int c4n = C / 4;
  VLNCE_CHECK_ARG(c4n <= 256 && 256 % c4n == 0 && (C / groups) % 2 == 0 && groups <= 64,
                  "groupnorm_fused16: layout not supported (C=%d groups=%d)", C, groups);
  int S = cdiv((long long)HW * C, 4096);
  if (S > HW) S = HW;
  if (S > 8) S = 8;
  while (S > 1 && (size_t)N * S * groups * 2 * sizeof(float) > workspace_bytes) --S;
  VLNCE_CHECK_ARG((size_t)N * S * groups * 2 * sizeof(float) <= workspace_bytes, "groupnorm_fused16: workspace too small");
  launch_pdl(gn_partial_kernel<__half>, dim3(N, S), dim3(256), 0, stream, (const __half*)x, workspace, HW, C, groups);
  VLNCE_LAUNCH_CHECK("gn_partial_kernel");
  int slices = cdiv((long long)HW * C, 8192);
  if (slices > 64) slices = 64;
  const size_t smem = sizeof(float) * 2 * C;
  launch_pdl(gn_apply_kernel<__half, __half>, dim3(N, slices), dim3(256), smem, stream, (const __half*)x,
             (const float*)workspace, S, gamma, beta, groups, eps, (const __half*)residual, (__half*)y, HW, C, relu);
  VLNCE_LAUNCH_CHECK("gn_apply_kernel");
  return VLNCE_OK;
}


// ------------------------------------------------------------------------------------------------
// SURVEY 8(f) N1 — observation transforms in front of the policy: ResizeShortestEdge (habitat_baselines v0.1.7
// obs_transformers.py / utils/common.py image_resize_shortest_edge: F.interpolate(mode="area") = adaptive average
// pooling, result cast back to the input dtype) fused with CenterCropperPerSensor
// (habitat_extensions/obs_transformers.py:21-87, hb center_crop): only the cropped window of the resized image is
// ever computed.  NHWC in, NHWC out; uint8 stays uint8 (truncating cast, bit-exact: the window sums are exact
// integers), fp32 stays fp32.  One thread per output element, channel-fastest (coalesced stores; the loads of
// neighbouring threads hit the same 128-B lines).  HBM-bound: reads each source pixel of the crop window once.
// ------------------------------------------------------------------------------------------------
template <typename T, int CC>
__global__ void resize_area_crop_kernel(const T* __restrict__ x, T* __restrict__ y, long long y_frame_stride, int N,
                                        int H, int W, int RH, int RW, int y0, int x0, int CH, int CW) {
  // one thread per output PIXEL (all CC channels): the source window is (h1-h0) runs of (w1-w0)*CC contiguous elements
  const long long total = (long long)N * CH * CW;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % CW);
    const int oy = (int)((idx / CW) % CH);
    const int n = (int)(idx / ((long long)CW * CH));
    const int ry = oy + y0, rx = ox + x0;  // position in the (virtual) resized image
    const int h0 = (int)(((long long)ry * H) / RH), h1 = (int)((((long long)ry + 1) * H + RH - 1) / RH);
    const int w0 = (int)(((long long)rx * W) / RW), w1 = (int)((((long long)rx + 1) * W + RW - 1) / RW);
    float acc[CC];
#pragma unroll
    for (int c = 0; c < CC; ++c) acc[c] = 0.f;
    for (int ih = h0; ih < h1; ++ih) {
      const T* row = x + (((size_t)n * H + ih) * W + w0) * CC;
      for (int k = 0; k < (w1 - w0); ++k) {
#pragma unroll
        for (int c = 0; c < CC; ++c) acc[c] += (float)row[k * CC + c];
      }
    }
    const float cnt = (float)((h1 - h0) * (w1 - w0));
    T* o = y + (size_t)n * y_frame_stride + ((size_t)oy * CW + ox) * CC;
#pragma unroll
    for (int c = 0; c < CC; ++c) o[c] = (T)(acc[c] / cnt);
  }
}

// Row-staged variant (the default): one CTA per (image, group of RB output rows).  The source rows under those output
// rows — only the column span the crop window maps back to — are copied to shared memory with 16-byte loads (the span
// starts at an arbitrary byte, so the copy starts at the enclosing aligned address and the consumers skip `mis`
// bytes), then one thread per output COLUMN computes its column window once and sums the RB windows out of shared
// memory in the same order as the element-wise kernel above.  All index arithmetic is 32-bit (the host checks that
// the products fit) and the per-column divisions are amortised over RB outputs: ncu showed the element-wise kernel
// and the first row-staged version issue-bound (80 % issue-active at 8-13 % of DRAM throughput), seven emulated
// 64-bit divisions per output pixel, not memory.
constexpr int kRacRB = 4;

template <typename T, int CC, int WMAX>   // WMAX: compile-time bound of the window width (0 = run-time loop)
__global__ void __launch_bounds__(256)
resize_area_crop_rows_kernel(const T* __restrict__ x, const T* __restrict__ x_end, T* __restrict__ y,
                             long long y_frame_stride, int H, int W, int RH, int RW, int y0, int x0, int CH, int CW,
                             int row_bytes /* smem bytes per staged row, multiple of 16 */, int groups /* per image */) {
  extern __shared__ uint4 smem_rac[];
  unsigned char* sm = reinterpret_cast<unsigned char*>(smem_rac);
  constexpr int RB = kRacRB;
  __shared__ int s_geo[4];             // n, first output row, wa, wb
  __shared__ int s_h[RB][2];           // source row range of each output row, relative to the first staged row
  __shared__ int s_hA, s_rows;
  __shared__ int s_mis[48];
  if (threadIdx.x == 0) {
    const unsigned g_ = blockIdx.x % (unsigned)groups, n_ = blockIdx.x / (unsigned)groups;
    const unsigned oy0 = g_ * RB;
    s_geo[0] = (int)n_; s_geo[1] = (int)oy0;
    s_geo[2] = (int)(((unsigned)x0 * (unsigned)W) / (unsigned)RW);
    s_geo[3] = (int)((((unsigned)x0 + (unsigned)CW) * (unsigned)W + (unsigned)RW - 1u) / (unsigned)RW);
    int hA = 0, hB = 0;
    for (int b = 0; b < RB; ++b) {
      const unsigned ry = oy0 + b + (unsigned)y0;
      const int h0 = (int)((ry * (unsigned)H) / (unsigned)RH);
      const int h1 = (int)(((ry + 1u) * (unsigned)H + (unsigned)RH - 1u) / (unsigned)RH);
      if (b == 0) hA = h0;
      if ((int)oy0 + b < CH) hB = h1;
      s_h[b][0] = h0 - hA; s_h[b][1] = h1 - hA;
    }
    s_hA = hA; s_rows = hB - hA;
  }
  __syncthreads();
  const int n = s_geo[0], oy0 = s_geo[1], wa = s_geo[2], wb = s_geo[3], hA = s_hA, rows = s_rows;
  const int span_bytes = (wb - wa) * CC * (int)sizeof(T);
  const unsigned char* xb = reinterpret_cast<const unsigned char*>(x);
  const unsigned char* xe = reinterpret_cast<const unsigned char*>(x_end);
  for (int r = 0; r < rows; ++r) {
    const unsigned char* src = reinterpret_cast<const unsigned char*>(x + (((size_t)n * H + hA + r) * W + wa) * CC);
    const int mis = (int)(reinterpret_cast<uintptr_t>(src) & 15);
    const unsigned char* base = src - mis;
    if (threadIdx.x == 0) s_mis[r] = mis;
    const int chunks = (span_bytes + mis + 15) >> 4;
    for (int i = threadIdx.x; i < chunks; i += blockDim.x) {
      const unsigned char* p = base + i * 16;
      uint4 v;
      if (p >= xb && p + 16 <= xe) {
        v = __ldg(reinterpret_cast<const uint4*>(p));
      } else {   // the first / last bytes of the tensor: never read outside it
        unsigned char b[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) b[e] = (p + e >= xb && p + e < xe) ? p[e] : (unsigned char)0;
        v = *reinterpret_cast<uint4*>(b);
      }
      *reinterpret_cast<uint4*>(sm + r * row_bytes + i * 16) = v;
    }
  }
  __syncthreads();
  for (int ox = threadIdx.x; ox < CW; ox += blockDim.x) {
    const unsigned rx = (unsigned)(ox + x0);
    const int w0 = (int)((rx * (unsigned)W) / (unsigned)RW) - wa;
    const int w1 = (int)(((rx + 1u) * (unsigned)W + (unsigned)RW - 1u) / (unsigned)RW) - wa;
    const int wn = w1 - w0;
#pragma unroll
    for (int b = 0; b < RB; ++b) {
      const int oy = oy0 + b;
      if (oy >= CH) break;
      const int r0 = s_h[b][0], r1 = s_h[b][1];
      float acc[CC];
#pragma unroll
      for (int c = 0; c < CC; ++c) acc[c] = 0.f;
      for (int r = r0; r < r1; ++r) {   // (CTA-uniform bounds)
        const T* row = reinterpret_cast<const T*>(sm + r * row_bytes + s_mis[r]) + w0 * CC;
        if constexpr (WMAX > 0) {
          // windows are 1..WMAX columns wide: a fully unrolled, predicated column loop instead of a per-thread trip
          // count (the divergent loops' branch / reconvergence instructions were 3/4 of the instruction stream);
          // adding the masked 0.f terms leaves the sum of the real terms, in the same order, unchanged
#pragma unroll
          for (int k = 0; k < WMAX; ++k) {
            const bool on = k < wn;
#pragma unroll
            for (int c = 0; c < CC; ++c) acc[c] += on ? (float)row[k * CC + c] : 0.f;
          }
        } else {
          for (int k = 0; k < wn; ++k) {
#pragma unroll
            for (int c = 0; c < CC; ++c) acc[c] += (float)row[k * CC + c];
          }
        }
      }
      const float cnt = (float)((r1 - r0) * wn);
      T* o = y + (size_t)n * y_frame_stride + ((size_t)oy * CW + ox) * CC;
#pragma unroll
      for (int c = 0; c < CC; ++c) o[c] = (T)(acc[c] / cnt);
    }
  }
}

extern "C" int vlnce_resize_area_crop(const void* x, void* y, int dtype, long long y_frame_stride, int N, int H, int W,
                                      int C, int RH, int RW, int y0, int x0, int CH, int CW, cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && y && N > 0 && H > 0 && W > 0 && C > 0 && RH > 0 && RW > 0, "resize_area_crop: bad arguments");
  VLNCE_CHECK_ARG(y0 >= 0 && x0 >= 0 && CH > 0 && CW > 0 && y0 + CH <= RH && x0 + CW <= RW,
                  "resize_area_crop: crop window outside the resized image");
  if (y_frame_stride == 0) y_frame_stride = (long long)CH * CW * C;
  VLNCE_CHECK_ARG(C >= 1 && C <= 4, "resize_area_crop: 1..4 channels (C=%d)", C);
  const long long total = (long long)N * CH * CW;
  // row-staged kernel when the source rows under one output row fit a modest shared-memory budget (always, for the
  // sensor sizes of the reference: 480x640 -> 256 -> 224 / 256); VLNCE_RESIZE_ROWS=0 keeps the element-wise kernel
  {
    static const bool rows_on = [] { const char* e = getenv("VLNCE_RESIZE_ROWS"); return !(e && e[0] == '0'); }();
    const int esz = dtype == VLNCE_U8 ? 1 : 4;
    const int wa = (int)(((long long)x0 * W) / RW), wb = (int)((((long long)x0 + CW) * W + RW - 1) / RW);
    const int row_bytes = (((wb - wa) * C * esz + 15 + 15) / 16) * 16;
    const int max_rows = kRacRB * ((H + RH - 1) / RH) + 2;   // source rows under kRacRB consecutive output rows
    const long long smem = (long long)row_bytes * max_rows + 64;   // + slack for the predicated-off reads of the last row
    const int wmax = (W + RW - 1) / RW + 1;                         // widest window of any output column
    const int groups = (CH + kRacRB - 1) / kRacRB;
    if (rows_on && (dtype == VLNCE_U8 || dtype == VLNCE_F32) && max_rows <= 48 && smem <= 48 * 1024 &&
        (long long)N * groups < (1ll << 31) && ((long long)RH + 1) * H + RH < (1ll << 32) &&
        ((long long)RW + 1) * W + RW < (1ll << 32)) {
#define VLNCE_RACR(T, CC)                                                                                              \
  if (wmax <= 3)                                                                                                       \
    resize_area_crop_rows_kernel<T, CC, 3><<<(unsigned)(N * groups), 256, (size_t)smem, stream>>>(                     \
        (const T*)x, (const T*)x + (size_t)N * H * W * C, (T*)y, y_frame_stride, H, W, RH, RW, y0, x0, CH, CW,         \
        row_bytes, groups);                                                                                            \
  else                                                                                                                 \
    resize_area_crop_rows_kernel<T, CC, 0><<<(unsigned)(N * groups), 256, (size_t)smem, stream>>>(                     \
        (const T*)x, (const T*)x + (size_t)N * H * W * C, (T*)y, y_frame_stride, H, W, RH, RW, y0, x0, CH, CW,         \
        row_bytes, groups)
#define VLNCE_RACR_C(T)                       \
  switch (C) {                                \
    case 1: VLNCE_RACR(T, 1); break;          \
    case 2: VLNCE_RACR(T, 2); break;          \
    case 3: VLNCE_RACR(T, 3); break;          \
    default: VLNCE_RACR(T, 4); break;         \
  }
      if (dtype == VLNCE_U8) {
        VLNCE_RACR_C(unsigned char)
      } else {
        VLNCE_RACR_C(float)
      }
#undef VLNCE_RACR_C
#undef VLNCE_RACR
      VLNCE_LAUNCH_CHECK("resize_area_crop_rows_kernel");
      return VLNCE_OK;
    }
  }
#define VLNCE_RAC(T, CC)                                                                                      \
  resize_area_crop_kernel<T, CC><<<grid_for(total, 128), 128, 0, stream>>>((const T*)x, (T*)y, y_frame_stride, N, H, W, \
                                                                           RH, RW, y0, x0, CH, CW)
#define VLNCE_RAC_C(T)                        \
  switch (C) {                                \
    case 1: VLNCE_RAC(T, 1); break;           \
    case 2: VLNCE_RAC(T, 2); break;           \
    case 3: VLNCE_RAC(T, 3); break;           \
    default: VLNCE_RAC(T, 4); break;          \
  }
  if (dtype == VLNCE_U8) {
    VLNCE_RAC_C(unsigned char)
  } else if (dtype == VLNCE_F32) {
    VLNCE_RAC_C(float)
  } else {
    set_error("resize_area_crop: dtype must be u8 or f32");
    return VLNCE_ERR_INVALID_ARGUMENT;
  }
#undef VLNCE_RAC_C
#undef VLNCE_RAC
  VLNCE_LAUNCH_CHECK("resize_area_crop_kernel");
  return VLNCE_OK;
}
