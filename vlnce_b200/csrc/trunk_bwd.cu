// Backward of the conv trunks (RGB_ENCODER.trainable / DEPTH_ENCODER.trainable = True,
// config/default.py:242,253; wired at cma_policy.py:80,94, seq2seq_policy.py:78,92) — the correctness-first fp32 path:
// CUDA-core kernels that mirror vlnce_conv2d_ref (the fp32 forward mode) layer by layer.  No shipped recipe trains the
// trunks; the tcgen05 dgrad / wgrad tiles are the follow-up once this path is the checker.
//
//   conv:       dx[n,h,w,c] = sum_{r,s,k} dy[n,p,q,k] w[k,r,s,c]   (p*stride + r - pad = h, q likewise)
//               dw[k,r,s,c] = alpha * sum_{n,p,q} x[n,p*stride+r-pad,q*stride+s-pad,c] dy[n,p,q,k]
//   GroupNorm:  per (image, group): dx = rstd (g - mean(g) - xhat mean(g xhat)), g = dy*gamma;
//               per-image partials of dgamma = sum dy xhat, dbeta = sum dy (summed over images by vlnce_colsum)
//   BatchNorm:  per channel over all rows: batch statistics (train mode) or running ones (eval mode)
//   max pool 3x3/2 pad 1: gradient to the first maximum of each window (torch's tie rule); adaptive average pool.
#include "common.cuh"

#include <algorithm>
#include <cstdint>

namespace vlnce {
namespace {

inline unsigned grid_1d(long long total, int block) {
  return (unsigned)std::max<long long>(1, std::min<long long>(cdiv(total, block), 148LL * 16));
}

__device__ __forceinline__ double block_sum_d(double v, double* scratch) {  // scratch: 32 doubles
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  double r = lane < nw ? scratch[lane] : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  return r;
}

__global__ void conv_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx,
                                  int N, int H, int W, int C, int K, int R, int S, int stride, int pad, int P, int Q,
                                  int accumulate) {
  const long long total = (long long)N * H * W * C;
  const long long kstride = (long long)R * S * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long long m = idx / C;
    const int iw = (int)(m % W);
    const int ih = (int)((m / W) % H);
    const int n = (int)(m / ((long long)H * W));
    float acc = 0.f;
    for (int r = 0; r < R; ++r) {
      const int th = ih + pad - r;
      if (th < 0 || th % stride != 0) continue;
      const int p = th / stride;
      if (p >= P) continue;
      for (int s = 0; s < S; ++s) {
        const int tw = iw + pad - s;
        if (tw < 0 || tw % stride != 0) continue;
        const int q = tw / stride;
        if (q >= Q) continue;
        const float* dyp = dy + ((size_t)(n * P + p) * Q + q) * K;
        const float* wp = w + ((size_t)r * S + s) * C + c;
        for (int k = 0; k < K; ++k) acc = fmaf(dyp[k], wp[(size_t)k * kstride], acc);
      }
    }
    dx[idx] = accumulate ? dx[idx] + acc : acc;
  }
}

template <typename TX>
__global__ void conv_wgrad_kernel(const TX* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, int N,
                                  int H, int W, int C, int K, int R, int S, int stride, int pad, int P, int Q,
                                  float alpha, int accumulate) {
  __shared__ float red[32];
  const long long elems = (long long)K * R * S * C;
  const long long M = (long long)N * P * Q;
  for (long long e = blockIdx.x; e < elems; e += gridDim.x) {
    const int c = (int)(e % C);
    const int s = (int)((e / C) % S);
    const int r = (int)((e / ((long long)C * S)) % R);
    const int k = (int)(e / ((long long)C * S * R));
    float acc = 0.f;
    for (long long m = threadIdx.x; m < M; m += blockDim.x) {
      const int q = (int)(m % Q);
      const int p = (int)((m / Q) % P);
      const int n = (int)(m / ((long long)P * Q));
      const int ih = p * stride + r - pad, iw = q * stride + s - pad;
      if ((unsigned)ih >= (unsigned)H || (unsigned)iw >= (unsigned)W) continue;
      acc = fmaf((float)x[((size_t)(n * H + ih) * W + iw) * C + c], dy[(size_t)m * K + k], acc);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) dw[e] = accumulate ? dw[e] + alpha * acc : alpha * acc;
    __syncthreads();
  }
}

// one CTA per (image, group); x = raw conv output (N, HW, C), dy = gradient w.r.t. the GroupNorm output
__global__ void groupnorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                     const float* __restrict__ gamma, float* __restrict__ dx,
                                     float* __restrict__ partial,  // (N, 2C): [dgamma | dbeta] per image
                                     int HW, int C, int groups, float eps) {
  __shared__ double scratch[32];
  extern __shared__ float chan[];  // 2*Cg: per-channel dgamma, dbeta of this (image, group)
  const int n = blockIdx.x / groups, g = blockIdx.x % groups;
  const int Cg = C / groups;
  const long long cnt = (long long)HW * Cg;
  const float* xb = x + (size_t)n * HW * C + (size_t)g * Cg;
  const float* dyb = dy + (size_t)n * HW * C + (size_t)g * Cg;
  float* dxb = dx + (size_t)n * HW * C + (size_t)g * Cg;
  for (int i = threadIdx.x; i < 2 * Cg; i += blockDim.x) chan[i] = 0.f;
  double s1 = 0.0, s2 = 0.0;
  for (long long i = threadIdx.x; i < cnt; i += blockDim.x) {
    const double v = xb[(i / Cg) * C + (i % Cg)];
    s1 += v;
    s2 += v * v;
  }
  s1 = block_sum_d(s1, scratch);
  s2 = block_sum_d(s2, scratch);
  const double mean = s1 / (double)cnt;
  const double var = fmax(s2 / (double)cnt - mean * mean, 0.0);
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float mu = (float)mean;
  double a = 0.0, b = 0.0;  // a = sum g*xhat, b = sum g  with g = dy*gamma
  for (long long i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int cc = (int)(i % Cg);
    const size_t o = (size_t)(i / Cg) * C + cc;
    const float xhat = (xb[o] - mu) * rstd;
    const float d = dyb[o];
    const float gm = d * gamma[g * Cg + cc];
    a += (double)gm * xhat;
    b += (double)gm;
    atomicAdd(&chan[cc], d * xhat);
    atomicAdd(&chan[Cg + cc], d);
  }
  a = block_sum_d(a, scratch);
  b = block_sum_d(b, scratch);
  const float am = (float)(a / (double)cnt), bm = (float)(b / (double)cnt);
  for (long long i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int cc = (int)(i % Cg);
    const size_t o = (size_t)(i / Cg) * C + cc;
    const float xhat = (xb[o] - mu) * rstd;
    dxb[o] = rstd * (dyb[o] * gamma[g * Cg + cc] - bm - xhat * am);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cg; i += blockDim.x) {
    partial[(size_t)n * 2 * C + g * Cg + i] = chan[i];
    partial[(size_t)n * 2 * C + C + g * Cg + i] = chan[Cg + i];
  }
}

// one CTA per channel (grid-stride); x = raw conv output (M, C)
__global__ void batchnorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                     const float* __restrict__ gamma, const float* __restrict__ rmean,
                                     const float* __restrict__ rvar, float* __restrict__ dx, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, long long M, int C, float eps, int training) {
  __shared__ double scratch[32];
  for (int c = blockIdx.x; c < C; c += gridDim.x) {
    double mean, var;
    if (training) {
      double s1 = 0.0, s2 = 0.0;
      for (long long m = threadIdx.x; m < M; m += blockDim.x) {
        const double v = x[(size_t)m * C + c];
        s1 += v;
        s2 += v * v;
      }
      s1 = block_sum_d(s1, scratch);
      s2 = block_sum_d(s2, scratch);
      mean = s1 / (double)M;
      var = fmax(s2 / (double)M - mean * mean, 0.0);
    } else {
      mean = rmean[c];
      var = rvar[c];
    }
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float mu = (float)mean;
    double sd = 0.0, sdx = 0.0;
    for (long long m = threadIdx.x; m < M; m += blockDim.x) {
      const float d = dy[(size_t)m * C + c];
      sd += d;
      sdx += (double)d * ((x[(size_t)m * C + c] - mu) * rstd);
    }
    sd = block_sum_d(sd, scratch);
    sdx = block_sum_d(sdx, scratch);
    if (threadIdx.x == 0) {
      dgamma[c] = (float)sdx;
      dbeta[c] = (float)sd;
    }
    const float gr = gamma[c] * rstd;
    const float md = training ? (float)(sd / (double)M) : 0.f, mdx = training ? (float)(sdx / (double)M) : 0.f;
    for (long long m = threadIdx.x; m < M; m += blockDim.x) {
      const size_t o = (size_t)m * C + c;
      dx[o] = gr * (dy[o] - md - (x[o] - mu) * rstd * mdx);
    }
    __syncthreads();
  }
}

__global__ void maxpool3x3s2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                                        int N, int H, int W, int C, int P, int Q) {
  const long long total = (long long)N * P * Q * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long long m = idx / C;
    const int q = (int)(m % Q);
    const int p = (int)((m / Q) % P);
    const int n = (int)(m / ((long long)P * Q));
    float best = -INFINITY;
    long long arg = -1;
    for (int r = 0; r < 3; ++r) {
      const int ih = 2 * p + r - 1;
      if ((unsigned)ih >= (unsigned)H) continue;
      for (int s = 0; s < 3; ++s) {
        const int iw = 2 * q + s - 1;
        if ((unsigned)iw >= (unsigned)W) continue;
        const long long o = ((long long)(n * H + ih) * W + iw) * C + c;
        const float v = x[o];
        if (v > best || arg < 0) {
          best = v;
          arg = o;
        }
      }
    }
    atomicAdd(&dx[arg], dy[idx]);
  }
}

// dout (N, O*O, ld): gradient w.r.t. the pooled positions (channels [0, C) of each row); dx (N, H, W, C)
__global__ void adaptive_avgpool_bwd_kernel(const float* __restrict__ dout, long long ld, float* __restrict__ dx, int N,
                                            int H, int W, int C, int O) {
  const long long total = (long long)N * H * W * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long long m = idx / C;
    const int iw = (int)(m % W);
    const int ih = (int)((m / W) % H);
    const int n = (int)(m / ((long long)H * W));
    float acc = 0.f;
    for (int oh = 0; oh < O; ++oh) {
      const int h0 = (oh * H) / O, h1 = ((oh + 1) * H + O - 1) / O;
      if (ih < h0 || ih >= h1) continue;
      for (int ow = 0; ow < O; ++ow) {
        const int w0 = (ow * W) / O, w1 = ((ow + 1) * W + O - 1) / O;
        if (iw < w0 || iw >= w1) continue;
        acc += dout[((size_t)n * O * O + oh * O + ow) * ld + c] / (float)((h1 - h0) * (w1 - w0));
      }
    }
    dx[idx] = acc;
  }
}

}  // namespace
}  // namespace vlnce

using namespace vlnce;

extern "C" int vlnce_conv2d_ref_dgrad(const float* dy, const float* w, float* dx, int N, int H, int W, int C, int K, int R,
                                      int S, int stride, int pad, int accumulate, cudaStream_t stream) {
  VLNCE_CHECK_ARG(dy && w && dx && N > 0 && H > 0 && W > 0 && C > 0 && K > 0 && R > 0 && S > 0 && stride > 0 && pad >= 0,
                  "conv2d_ref_dgrad: bad arguments");
  const int P = (H + 2 * pad - R) / stride + 1, Q = (W + 2 * pad - S) / stride + 1;
  VLNCE_CHECK_ARG(P > 0 && Q > 0, "conv2d_ref_dgrad: empty output");
  conv_dgrad_kernel<<<grid_1d((long long)N * H * W * C, 256), 256, 0, stream>>>(dy, w, dx, N, H, W, C, K, R, S, stride,
                                                                               pad, P, Q, accumulate);
  VLNCE_LAUNCH_CHECK("conv_dgrad_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_conv2d_ref_wgrad(const void* x, int x_dtype, const float* dy, float* dw, int N, int H, int W, int C,
                                      int K, int R, int S, int stride, int pad, float alpha, int accumulate,
                                      cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && dy && dw && N > 0 && H > 0 && W > 0 && C > 0 && K > 0 && R > 0 && S > 0 && stride > 0 && pad >= 0,
                  "conv2d_ref_wgrad: bad arguments");
  VLNCE_CHECK_ARG(x_dtype == VLNCE_F32 || x_dtype == VLNCE_U8, "conv2d_ref_wgrad: x must be f32 or u8");
  const int P = (H + 2 * pad - R) / stride + 1, Q = (W + 2 * pad - S) / stride + 1;
  VLNCE_CHECK_ARG(P > 0 && Q > 0, "conv2d_ref_wgrad: empty output");
  const unsigned grid = (unsigned)std::min<long long>((long long)K * R * S * C, 1 << 20);
  if (x_dtype == VLNCE_F32)
    conv_wgrad_kernel<float><<<grid, 128, 0, stream>>>((const float*)x, dy, dw, N, H, W, C, K, R, S, stride, pad, P, Q,
                                                       alpha, accumulate);
  else
    conv_wgrad_kernel<uint8_t><<<grid, 128, 0, stream>>>((const uint8_t*)x, dy, dw, N, H, W, C, K, R, S, stride, pad, P,
                                                         Q, alpha, accumulate);
  VLNCE_LAUNCH_CHECK("conv_wgrad_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_groupnorm_bwd(const float* x, const float* dy, const float* gamma, float* dx, float* partial, int N,
                                   int HW, int C, int groups, float eps, cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && dy && gamma && dx && partial && N > 0 && HW > 0 && C > 0 && groups > 0 && C % groups == 0,
                  "groupnorm_bwd: bad arguments");
  const size_t smem = sizeof(float) * 2 * (C / groups);
  groupnorm_bwd_kernel<<<N * groups, 256, smem, stream>>>(x, dy, gamma, dx, partial, HW, C, groups, eps);
  VLNCE_LAUNCH_CHECK("groupnorm_bwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_batchnorm_bwd(const float* x, const float* dy, const float* gamma, const float* running_mean,
                                   const float* running_var, float* dx, float* dgamma, float* dbeta, long long rows, int C,
                                   float eps, int training, cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && dy && gamma && dx && dgamma && dbeta && rows > 0 && C > 0, "batchnorm_bwd: bad arguments");
  VLNCE_CHECK_ARG(training || (running_mean && running_var), "batchnorm_bwd: eval mode needs the running statistics");
  batchnorm_bwd_kernel<<<std::min(C, 148 * 8), 256, 0, stream>>>(x, dy, gamma, running_mean, running_var, dx, dgamma,
                                                                dbeta, rows, C, eps, training);
  VLNCE_LAUNCH_CHECK("batchnorm_bwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_maxpool3x3s2_bwd(const float* x, const float* dy, float* dx, int N, int H, int W, int C,
                                      cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && dy && dx && N > 0 && H > 0 && W > 0 && C > 0, "maxpool3x3s2_bwd: bad arguments");
  const int P = (H + 2 - 3) / 2 + 1, Q = (W + 2 - 3) / 2 + 1;
  cudaError_t e = cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)N * H * W * C, stream);
  if (e != cudaSuccess) return cuda_status(e, "maxpool3x3s2_bwd memset");
  maxpool3x3s2_bwd_kernel<<<grid_1d((long long)N * P * Q * C, 256), 256, 0, stream>>>(x, dy, dx, N, H, W, C, P, Q);
  VLNCE_LAUNCH_CHECK("maxpool3x3s2_bwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_adaptive_avgpool_bwd(const float* dout, long long ld, float* dx, int N, int H, int W, int C, int O,
                                          cudaStream_t stream) {
  VLNCE_CHECK_ARG(dout && dx && N > 0 && H > 0 && W > 0 && C > 0 && O > 0 && ld >= C, "adaptive_avgpool_bwd: bad arguments");
  adaptive_avgpool_bwd_kernel<<<grid_1d((long long)N * H * W * C, 256), 256, 0, stream>>>(dout, ld, dx, N, H, W, C, O);
  VLNCE_LAUNCH_CHECK("adaptive_avgpool_bwd_kernel");
  return VLNCE_OK;
}
