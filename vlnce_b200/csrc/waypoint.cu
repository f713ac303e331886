// Kernels of the waypoint path: the pieces of WaypointPredictionNet
// (models/waypoint_predictors.py:286-625) that the CMA head kernels do not already cover, the
// action post-processing of WaypointPolicy.act / evaluate_actions (models/waypoint_policy.py:129-333)
// with the TruncatedNormal closed forms (models/utils.py:24-152), the fused WDDPPO mini-batch loss
// and its gradients (common/ddppo_alg.py:56-124), gradient-norm clipping (habitat_baselines v0.1.7
// PPO.before_step) and the GAE recursion of the rollout storage (common/rollout_storage.py:127-152).
// All fp32 / integer, HBM- or latency-bound: coalesced channel-fastest accesses, one CTA per row for
// the row reductions, no floating-point atomics (deterministic).
#include "common.cuh"

namespace vlnce {
namespace {

inline unsigned grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148ll * 16;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

constexpr float kInvSqrt2Pi = 0.3989422804014327f;
constexpr float kSqrt2Pi = 2.5066282746310002f;
constexpr float kInvSqrt2 = 0.7071067811865476f;
constexpr float kTwoPi = 6.283185307179586f;
constexpr float kHalfLog2PiE = 1.4189385332046727f;  // 0.5 * log(2*pi*e)

// ------------------------------------------------------------------------------------
// waypoint_predictors.py:327-366 — the 12 pano frames followed by the history frame (zeroed
// where masks == 0) as ONE trunk batch.  Frame order inside the batch is [all pano frames
// (b-major) | one history frame per env]; both trunks normalise per sample (GroupNorm) or with
// folded statistics (eval-mode BN), so the order is free.  Pure byte movement.
// ------------------------------------------------------------------------------------
template <typename V>
__global__ void stack_frames_kernel(const V* __restrict__ cur, const V* __restrict__ hist,
                                    const float* __restrict__ masks, V* __restrict__ out, long long cur_vec,
                                    long long frame_vec, int B) {
  const long long total = cur_vec + (long long)B * frame_vec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    if (i < cur_vec) {
      out[i] = cur[i];
    } else {
      const long long j = i - cur_vec;
      const int b = (int)(j / frame_vec);
      V v = hist[j];
      if (masks[b] == 0.f) memset(&v, 0, sizeof(V));
      out[i] = v;
    }
  }
}

// E[r, p, 0:C] = E[r, 0, 0:C] for p >= 1 (the globally pooled RGB vector shown at all
// 16 positions when single_spatial_filter=False, resnet_encoders.py:154-162)
__global__ void broadcast_positions_kernel(float* __restrict__ E, long long lde_r, long long lde_p, int R, int P,
                                           int C) {
  const long long total = (long long)R * (P - 1) * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int p = (int)((idx / C) % (P - 1)) + 1;
    const long long r = idx / ((long long)C * (P - 1));
    E[r * lde_r + p * lde_p + c] = E[r * lde_r + c];
  }
}

// waypoint_predictors.py:268-275,372-384: [sin h, cos h, offset_c, distance_c] * mask, h = pano * 2pi/P
__global__ void wp_prev_action_kernel(const long long* __restrict__ pano, const float* __restrict__ offset,
                                      const float* __restrict__ distance, const float* __restrict__ masks,
                                      float* __restrict__ out, long long ldo, float* __restrict__ out2,
                                      long long ldo2, int rows, float rad_per_pano, float off_a, float off_b,
                                      float dist_a, float dist_b) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float h = (float)pano[r] * rad_per_pano;
  const float m = masks[r];
  const float v[4] = {sinf(h) * m, cosf(h) * m, fmaf(off_b, offset[r], off_a) * m, fmaf(dist_b, distance[r], dist_a) * m};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    out[r * ldo + k] = v[k];
    if (out2 != nullptr) out2[r * ldo2 + k] = v[k];
  }
}

// ------------------------------------------------------------------------------------
// Instruction attention of the waypoint net (waypoint_predictors.py:419-424 + utils.py:155-178),
// SURVEY Q4 kept: the mask (True = padding) MULTIPLIES the energies, so every real token has
// energy 0 and every padded position (zero embedding -> key = bias b_k) keeps q.b_k.  With
// len valid tokens and npad = Lmax - len padded ones (Lmax = longest row of the batch, the
// length pad_packed_sequence pads to):
//   a_valid = 1 / (len + npad * exp(c)),  c = scale * q.b_k,  out = a_valid * sum_{p<len} V_p
// (padded values are zero).  save[r] = {a_valid, total probability mass on the padding}.
// One CTA per row.
// ------------------------------------------------------------------------------------
__global__ void wp_inst_attn_fwd_kernel(const float* __restrict__ q, long long ldq, const float* __restrict__ bk,
                                        const float* __restrict__ instr, const int* __restrict__ lengths,
                                        float scale, float* __restrict__ out, long long ldo,
                                        float* __restrict__ save, int R, int L, int D, int Dk) {
  __shared__ float red[32];
  const int r = blockIdx.x;
  float mx = 0.f;
  for (int i = threadIdx.x; i < R; i += blockDim.x) mx = fmaxf(mx, (float)min(lengths[i], L));
  const int lmax = (int)block_max(mx, red);
  __syncthreads();
  float part = 0.f;
  for (int d = threadIdx.x; d < Dk; d += blockDim.x) part = fmaf(q[r * ldq + d], bk[d], part);
  const float c = block_sum(part, red) * scale;
  const int len = min(lengths[r], L);
  const int npad = lmax - len;
  const float m = (npad > 0 && (c > 0.f || len == 0)) ? c : 0.f;  // softmax shift
  const float ev = len > 0 ? expf(0.f - m) : 0.f;
  const float ep = npad > 0 ? expf(c - m) : 0.f;
  const float z = (float)len * ev + (float)npad * ep;
  const float a_v = z > 0.f ? ev / z : 0.f;
  const float a_pad = z > 0.f ? (float)npad * ep / z : 0.f;
  const float* base = instr + (size_t)r * L * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < len; ++p) s += base[(size_t)p * D + d];
    out[r * ldo + d] = a_v * s;
  }
  if (threadIdx.x == 0) { save[2 * r] = a_v; save[2 * r + 1] = a_pad; }
}

// d_instr[r,p,:] = a_valid * dout (p < len, else 0);  dc = -(out . dout) * a_pad;  dcs[r] = dc * scale
// (the caller turns dcs into dq = dcs b_k^T and db_k = dcs^T q with two skinny GEMMs; dW_k is exactly 0).
__global__ void wp_inst_attn_bwd_kernel(const float* __restrict__ save, const float* __restrict__ out, long long ldo,
                                        const float* __restrict__ dout, long long lddo,
                                        const int* __restrict__ lengths, float scale, float* __restrict__ dcs,
                                        float* __restrict__ d_instr, int L, int D) {
  __shared__ float red[32];
  const int r = blockIdx.x;
  float part = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) part = fmaf(out[r * ldo + d], dout[r * lddo + d], part);
  const float dot = block_sum(part, red);
  const float a_v = save[2 * r], a_pad = save[2 * r + 1];
  if (threadIdx.x == 0) dcs[r] = -dot * a_pad * scale;
  const int len = min(lengths[r], L);
  float* base = d_instr + (size_t)r * L * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float g = a_v * dout[r * lddo + d];
    for (int p = 0; p < L; ++p) base[(size_t)p * D + d] = p < len ? g : 0.f;
  }
}

// waypoint_predictors.py:459-464 (quirk W2): repeat_interleave without a dim repeats ELEMENTS:
// out.flat[m] = q.flat[m / P]; viewed (B*P, D) by the caller.
__global__ void wp_scramble_query_kernel(const float* __restrict__ q, float* __restrict__ out, long long n_in, int P) {
  const long long total = n_in * P;
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < total;
       m += (long long)gridDim.x * blockDim.x)
    out[m] = q[m / P];
}
__global__ void wp_scramble_query_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dq, long long n_in,
                                             int P) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_in;
       i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < P; ++j) s += dout[i * P + j];
    dq[i] = s;
  }
}

// ------------------------------------------------------------------------------------
// LayerNorm(eps) over the last dimension (utils.py:214-215,263-265), one CTA per row.
// ------------------------------------------------------------------------------------
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float* __restrict__ y, long long ldy,
                                     float* __restrict__ save, int D, float eps) {
  __shared__ float red[32];
  const int r = blockIdx.x;
  float s = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) s += x[r * ldx + d];
  const float mean = block_sum(s, red) / (float)D;
  __syncthreads();
  float v = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) { const float t = x[r * ldx + d] - mean; v = fmaf(t, t, v); }
  const float rstd = rsqrtf(block_sum(v, red) / (float)D + eps);
  for (int d = threadIdx.x; d < D; d += blockDim.x)
    y[r * ldy + d] = fmaf((x[r * ldx + d] - mean) * rstd, gamma[d], beta[d]);
  if (threadIdx.x == 0 && save != nullptr) { save[2 * r] = mean; save[2 * r + 1] = rstd; }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma;  prod = dy * xhat (its column sum is
// d_gamma; the column sum of dy is d_beta — both taken by the caller with vlnce_colsum).
__global__ void layernorm_bwd_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ x,
                                     long long ldx, const float* __restrict__ gamma, const float* __restrict__ save,
                                     float* __restrict__ dx, long long lddx, float* __restrict__ prod, long long ldp,
                                     int D) {
  __shared__ float red[32];
  const int r = blockIdx.x;
  const float mean = save[2 * r], rstd = save[2 * r + 1];
  float s1 = 0.f, s2 = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float xh = (x[r * ldx + d] - mean) * rstd;
    const float g = dy[r * lddy + d] * gamma[d];
    s1 += g;
    s2 = fmaf(g, xh, s2);
  }
  const float m1 = block_sum(s1, red) / (float)D;
  __syncthreads();
  const float m2 = block_sum(s2, red) / (float)D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float xh = (x[r * ldx + d] - mean) * rstd;
    const float g = dy[r * lddy + d] * gamma[d];
    dx[r * lddx + d] = rstd * (g - m1 - xh * m2);
    prod[r * ldp + d] = dy[r * lddy + d] * xh;
  }
}

// ------------------------------------------------------------------------------------
// pano logits (waypoint_predictors.py:556-559): out[b, i] = F[b, i, :] . xs[b, :]; one warp per (b, i).
// ------------------------------------------------------------------------------------
__global__ void wp_dot_frames_fwd_kernel(const float* __restrict__ F, long long ldf_r, long long ldf_p,
                                         const float* __restrict__ xs, long long ldx, float* __restrict__ out,
                                         long long ldo, int B, int P, int D) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= B * P) return;
  const int b = w / P, i = w % P;
  const float* f = F + b * ldf_r + i * ldf_p;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) s = fmaf(f[d], xs[b * ldx + d], s);
  s = warp_sum(s);
  if (lane == 0) out[b * ldo + i] = s;
}
// dxs[b,d] = sum_i dout[b,i] F[b,i,d];  dF[b,i,d] (+)= dout[b,i] xs[b,d]
__global__ void wp_dot_frames_bwd_kernel(const float* __restrict__ F, long long ldf_r, long long ldf_p,
                                         const float* __restrict__ xs, long long ldx, const float* __restrict__ dout,
                                         long long lddo, float* __restrict__ dF, long long lddf_r, long long lddf_p,
                                         int accumulate, float* __restrict__ dxs, long long lddx, int B, int P, int D) {
  const long long total = (long long)B * D;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / D), d = (int)(idx % D);
    const float xv = xs[b * ldx + d];
    float acc = 0.f;
    for (int i = 0; i < P; ++i) {
      const float g = dout[b * lddo + i];
      acc = fmaf(g, F[b * ldf_r + i * ldf_p + d], acc);
      float* o = dF + b * lddf_r + i * lddf_p + d;
      *o = accumulate ? *o + g * xv : g * xv;
    }
    dxs[b * lddx + d] = acc;
  }
}

// out[b, i, :] = x[b, :] (x broadcast over the pano frames into its concat slice, :561-567)
__global__ void broadcast_rows_kernel(const float* __restrict__ x, long long ldx, float* __restrict__ out,
                                      long long ldo_r, long long ldo_p, int B, int P, int D) {
  const long long total = (long long)B * P * D;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(idx % D);
    const int i = (int)((idx / D) % P);
    const int b = (int)(idx / ((long long)D * P));
    out[b * ldo_r + i * ldo_p + d] = x[b * ldx + d];
  }
}
// out[r, d] (+)= sum_p E[r, p, d]
__global__ void sum_positions_kernel(const float* __restrict__ E, long long lde_r, long long lde_p,
                                     float* __restrict__ out, long long ldo, int accumulate, int R, int P, int D) {
  const long long total = (long long)R * D;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / D), d = (int)(idx % D);
    float acc = 0.f;
    for (int p = 0; p < P; ++p) acc += E[r * lde_r + p * lde_p + d];
    out[r * ldo + d] = accumulate ? out[r * ldo + d] + acc : acc;
  }
}

// ------------------------------------------------------------------------------------
// Output activations of the four continuous heads (waypoint_predictors.py:573-613).  raw (n,4) columns:
// 0 distance, 1 distance variance, 2 offset, 3 offset variance.  c = {d_span, d_min, dv_span, dv_min,
// o_scale, 1/T, ov_span, ov_min}.
// ------------------------------------------------------------------------------------
struct HeadConsts { float d_span, d_min, dv_span, dv_min, o_scale, inv_t, ov_span, ov_min; };

__global__ void wp_head_act_fwd_kernel(const float* __restrict__ raw, float* __restrict__ d1, float* __restrict__ d2,
                                       float* __restrict__ o1, float* __restrict__ o2, long long n, HeadConsts c) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float4 r = *reinterpret_cast<const float4*>(raw + 4 * i);
    d1[i] = fmaf(c.d_span, sigmoid_precise(r.x), c.d_min);
    d2[i] = fmaf(c.dv_span, sigmoid_precise(r.y), c.dv_min);
    o1[i] = c.o_scale * tanhf(r.z * c.inv_t);
    o2[i] = fmaf(c.ov_span, sigmoid_precise(r.w), c.ov_min);
  }
}
__global__ void wp_head_act_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ g_d1,
                                       const float* __restrict__ g_d2, const float* __restrict__ g_o1,
                                       const float* __restrict__ g_o2, float* __restrict__ g_raw, long long n,
                                       HeadConsts c) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float4 r = *reinterpret_cast<const float4*>(raw + 4 * i);
    const float s0 = sigmoid_precise(r.x), s1 = sigmoid_precise(r.y), s3 = sigmoid_precise(r.w);
    const float t = tanhf(r.z * c.inv_t);
    float4 g;
    g.x = g_d1[i] * c.d_span * s0 * (1.f - s0);
    g.y = g_d2[i] * c.dv_span * s1 * (1.f - s1);
    g.z = g_o1[i] * c.o_scale * (1.f - t * t) * c.inv_t;
    g.w = g_o2[i] * c.ov_span * s3 * (1.f - s3);
    *reinterpret_cast<float4*>(g_raw + 4 * i) = g;
  }
}

// ------------------------------------------------------------------------------------
// TruncatedNormal closed forms (utils.py:65-96,126-146): location mu, scale sigma, support [lo, hi].
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float phi_(float x) { return expf(-0.5f * x * x) * kInvSqrt2Pi; }
__device__ __forceinline__ float Phi_(float x) { return 0.5f * (1.f + erff(x * kInvSqrt2)); }

struct TN {
  float a, b, pa, pb, Z, mean, var, ent;
};
__device__ __forceinline__ TN tn_stats(float mu, float sigma, float lo, float hi) {
  TN t;
  t.a = (lo - mu) / sigma;
  t.b = (hi - mu) / sigma;
  t.pa = phi_(t.a);
  t.pb = phi_(t.b);
  t.Z = Phi_(t.b) - Phi_(t.a);
  const float dz = (t.pb - t.pa) / t.Z;
  t.mean = mu - sigma * dz;
  t.var = sigma * sigma * (1.f - (t.b * t.pb - t.a * t.pa) / t.Z - dz * dz);
  t.ent = kHalfLog2PiE + logf(sigma * t.Z) + (t.a * t.pa - t.b * t.pb) / (2.f * t.Z);
  return t;
}
__device__ __forceinline__ float tn_log_prob(float y, float mu, float sigma, float Z) {
  const float u = (y - mu) / sigma;
  return logf(expf(-0.5f * u * u) / (sigma * kSqrt2Pi) / Z);
}
// gradients of log_prob(y) and of the entropy w.r.t. (mu, variance v = sigma^2)
__device__ __forceinline__ void tn_grads(const TN& t, float y, float mu, float sigma, float* dlp_dmu, float* dlp_dv,
                                         float* dH_dmu, float* dH_dv) {
  const float inv_s = 1.f / sigma;
  const float u = (y - mu) * inv_s;
  const float dZ_dmu = (t.pa - t.pb) * inv_s;
  const float N = t.a * t.pa - t.b * t.pb;
  const float dZ_ds = N * inv_s;
  *dlp_dmu = u * inv_s - dZ_dmu / t.Z;
  const float dlp_ds = u * u * inv_s - inv_s - dZ_ds / t.Z;
  const float dN_dmu = ((1.f - t.b * t.b) * t.pb - (1.f - t.a * t.a) * t.pa) * inv_s;
  const float dN_ds = (t.b * (1.f - t.b * t.b) * t.pb - t.a * (1.f - t.a * t.a) * t.pa) * inv_s;
  *dH_dmu = dZ_dmu / t.Z + dN_dmu / (2.f * t.Z) - N * dZ_dmu / (2.f * t.Z * t.Z);
  const float dH_ds = inv_s + dZ_ds / t.Z + dN_ds / (2.f * t.Z) - N * dZ_ds / (2.f * t.Z * t.Z);
  const float ds_dv = 0.5f * inv_s;
  *dlp_dv = dlp_ds * ds_dv;
  *dH_dv = dH_ds * ds_dv;
}

// counter-based generator for the sampling path (not part of any parity claim: the reference draws from
// torch's global generator)
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ float uniform01(unsigned long long seed, unsigned long long ctr) {
  return ((float)(mix64(seed ^ mix64(ctr)) >> 40) + 0.5f) * (1.f / 16777216.f);
}
__device__ __forceinline__ float tn_sample(float mu, float sigma, float lo, float hi, unsigned long long seed,
                                           unsigned long long ctr) {
  for (int k = 0; k < 10000; ++k) {  // rejection sampling with the reference's resample limit (utils.py:106-124)
    const float u1 = uniform01(seed, ctr + 2ull * k), u2 = uniform01(seed, ctr + 2ull * k + 1);
    const float s = fmaf(sigma, sqrtf(-2.f * logf(u1)) * cosf(kTwoPi * u2), mu);
    if (s >= lo && s <= hi) return s;
  }
  return fminf(fmaxf(mu, lo), hi);
}

struct ActConsts {
  float d_lo, d_hi, o_lim;
  int predict_distance, predict_offset;
  int nd, no;                            // classes of a discrete distance / offset head (0 = continuous TruncatedNormal)
  float dist_a, dist_b, off_a, off_b;    // {distance,offset}_to_continuous = a + b * value (waypoint_predictors.py:184-209)
};

// One action component (distance or offset) of one row at the selected pano frame, either a two-sided
// TruncatedNormal(loc = v1[idx], scale = sqrt(v2[idx])) (waypoint_policy.py:48-55) or a Categorical over ncls
// logits v1[idx, :] (:56-59, batched_index_select).
struct Comp {
  int ncls;
  float mu, sg;      // continuous
  TN t;
  const float* lg;   // discrete: raw logits of the selected frame
  float lse, ent;    // discrete: logsumexp, entropy
};
__device__ __forceinline__ Comp comp_load(int ncls, const float* v1, const float* v2, size_t row_frame, float lo, float hi) {
  Comp c;
  c.ncls = ncls;
  if (ncls > 0) {
    c.lg = v1 + row_frame * ncls;
    float mx = -INFINITY;
    for (int j = 0; j < ncls; ++j) mx = fmaxf(mx, c.lg[j]);
    float s = 0.f;
    for (int j = 0; j < ncls; ++j) s += expf(c.lg[j] - mx);
    c.lse = mx + logf(s);
    float h = 0.f;
    for (int j = 0; j < ncls; ++j) { const float l = c.lg[j] - c.lse; h -= expf(l) * l; }
    c.ent = h;
    c.mu = c.sg = 0.f;
  } else {
    c.lg = nullptr;
    c.mu = v1[row_frame];
    c.sg = sqrtf(v2[row_frame]);
    c.t = tn_stats(c.mu, c.sg, lo, hi);
    c.ent = c.t.ent;
    c.lse = 0.f;
  }
  return c;
}
__device__ __forceinline__ float comp_log_prob(const Comp& c, float y) {
  if (c.ncls > 0) {
    int k = (int)y;
    k = k < 0 ? 0 : (k >= c.ncls ? c.ncls - 1 : k);
    return c.lg[k] - c.lse;
  }
  return tn_log_prob(y, c.mu, c.sg, c.t.Z);
}
__device__ __forceinline__ float comp_mode(const Comp& c) {  // CustomFixedCategorical.mode(): first maximum
  if (c.ncls == 0) return c.mu;
  int best = 0;
  float bv = expf(c.lg[0] - c.lse);
  for (int j = 1; j < c.ncls; ++j) { const float v = expf(c.lg[j] - c.lse); if (v > bv) { bv = v; best = j; } }
  return (float)best;
}
// gradients w.r.t. the head outputs of this row's frames: g1 / g2 are (P, max(ncls,1)) blocks, zero except at idx
// (mask_lp gates the log-prob term, mask_h the entropy term: they differ for categorical heads, see W4 below)
__device__ __forceinline__ void comp_grads(const Comp& c, float y, int P, int idx, float mask_lp, float mask_h,
                                           float g_lp, float g_h, float* g1, float* g2) {
  const int w = c.ncls > 0 ? c.ncls : 1;
  for (int j = 0; j < P * w; ++j) g1[j] = 0.f;
  if (c.ncls > 0) {
    int k = (int)y;
    k = k < 0 ? 0 : (k >= c.ncls ? c.ncls - 1 : k);
    for (int j = 0; j < c.ncls; ++j) {
      const float l = c.lg[j] - c.lse, p = expf(l);
      g1[idx * w + j] = mask_lp * g_lp * ((j == k ? 1.f : 0.f) - p) + mask_h * g_h * (-p * (l + c.ent));
    }
    return;
  }
  for (int j = 0; j < P; ++j) g2[j] = 0.f;
  float dlp_dmu, dlp_dv, dH_dmu, dH_dv;
  tn_grads(c.t, y, c.mu, c.sg, &dlp_dmu, &dlp_dv, &dH_dmu, &dH_dv);
  g1[idx] = mask_lp * g_lp * dlp_dmu + mask_h * g_h * dH_dmu;
  g2[idx] = mask_lp * g_lp * dlp_dv + mask_h * g_h * dH_dv;
}

// ------------------------------------------------------------------------------------
// WaypointPolicy.act post-processing (waypoint_policy.py:129-247), one thread per env.
// fout (B, 10): offset, distance, action_log_prob, var_distance, var_offset, mode_offset, mode_distance, r, theta, stop
// (discrete heads: offset / distance / modes are class indices, variances NaN like torch's Categorical.variance)
// ------------------------------------------------------------------------------------
__global__ void wp_act_kernel(const float* __restrict__ logp, const float* __restrict__ o1, const float* __restrict__ o2,
                              const float* __restrict__ d1, const float* __restrict__ d2, long long* __restrict__ pano_stop,
                              float* __restrict__ fout, int B, int P, int deterministic, unsigned long long seed,
                              unsigned long long offset, ActConsts c) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* lp = logp + (size_t)b * (P + 1);
  const unsigned long long ctr = offset + (unsigned long long)b * 65536ull;
  int ps = 0;
  if (deterministic) {  // CustomFixedCategorical.mode(): first maximum of the probabilities
    float bv = expf(lp[0]);
    for (int j = 1; j <= P; ++j) { const float v = expf(lp[j]); if (v > bv) { bv = v; ps = j; } }
  } else {
    const float u = uniform01(seed, ctr);
    float acc = 0.f;
    ps = P;
    for (int j = 0; j <= P; ++j) { acc += expf(lp[j]); if (u < acc) { ps = j; break; } }
  }
  const int pano = ps % P;
  const float go = ps != P ? 1.f : 0.f;
  const size_t rf = (size_t)b * P + pano;
  const Comp cd = comp_load(c.nd, d1, d2, rf, c.d_lo, c.d_hi);
  const Comp co = comp_load(c.no, o1, o2, rf, -c.o_lim, c.o_lim);
  float dist, ofs;
  if (c.nd > 0) {
    dist = comp_mode(cd);
    if (!deterministic) {
      const float u = uniform01(seed, ctr + 8);
      float acc = 0.f;
      dist = (float)(c.nd - 1);
      for (int j = 0; j < c.nd; ++j) { acc += expf(cd.lg[j] - cd.lse); if (u < acc) { dist = (float)j; break; } }
    }
  } else {
    dist = deterministic ? cd.mu : tn_sample(cd.mu, cd.sg, c.d_lo, c.d_hi, seed, ctr + 16);
  }
  if (c.no > 0) {
    ofs = comp_mode(co);
    if (!deterministic) {
      const float u = uniform01(seed, ctr + 9);
      float acc = 0.f;
      ofs = (float)(c.no - 1);
      for (int j = 0; j < c.no; ++j) { acc += expf(co.lg[j] - co.lse); if (u < acc) { ofs = (float)j; break; } }
    }
  } else {
    ofs = deterministic ? co.mu : tn_sample(co.mu, co.sg, -c.o_lim, c.o_lim, seed, ctr + 32768);
  }
  const float lp_d = comp_log_prob(cd, dist), lp_o = comp_log_prob(co, ofs);
  float a_dist = fmaf(c.dist_b, dist, c.dist_a), a_ofs = fmaf(c.off_b, ofs, c.off_a);
  float var_d = c.nd > 0 ? NAN : cd.t.var, var_o = c.no > 0 ? NAN : co.t.var;
  const float mode_d = comp_mode(cd), mode_o = comp_mode(co);
  if (!c.predict_offset) { a_ofs = 0.f; ofs = 0.f; var_o = 0.f; }
  if (!c.predict_distance) { a_dist = 0.25f; dist = c.nd > 0 ? 0.f : 0.25f; var_d = 0.f; }
  float theta = (float)pano * (kTwoPi / (float)P) + a_ofs;
  theta -= floorf(theta / kTwoPi) * kTwoPi;  // torch `%`: sign of the divisor
  float alp = lp[ps];
  if (c.predict_distance) alp += go * lp_d;
  if (c.predict_offset) alp += go * lp_o;
  pano_stop[b] = ps;
  float* f = fout + (size_t)b * 10;
  f[0] = ofs; f[1] = dist; f[2] = alp; f[3] = var_d; f[4] = var_o; f[5] = mode_o; f[6] = mode_d;
  f[7] = go != 0.f ? a_dist : 0.f; f[8] = go != 0.f ? theta : 0.f; f[9] = 1.f - go;
}

// Gradients of one row of evaluate_actions' outputs w.r.t. the net outputs, given the upstream gradients of
// {action_log_prob, entropy_pano, entropy_offset, entropy_distance}.
__device__ __forceinline__ void wp_row_grads(const float* lp, int a, int idx, int P, float Hp, const Comp& cd,
                                             const Comp& co, float y_d, float y_o, float dm, float om, float dmh,
                                             float omh, float g_lp, float g_hp, float g_ho, float g_hd, float* d_raw,
                                             float* d_o1, float* d_o2, float* d_d1, float* d_d2) {
  for (int j = 0; j <= P; ++j) {
    const float p = expf(lp[j]);
    d_raw[j] = g_lp * ((j == a ? 1.f : 0.f) - p) + g_hp * (-p * (lp[j] + Hp));
  }
  comp_grads(cd, y_d, P, idx, dm, dmh, g_lp, g_hd, d_d1, d_d2);
  comp_grads(co, y_o, P, idx, om, omh, g_lp, g_ho, d_o1, d_o2);
}

// backward of evaluate_actions for arbitrary upstream gradients up (R,4) = d{action_log_prob, entropy_pano,
// entropy_offset, entropy_distance} (the autograd bridge; the fused PPO kernel below inlines its own).
// Gradient blocks: d_d1 (R, P, max(nd,1)), d_d2 (R, P) [continuous only]; likewise for the offset.
__global__ void wp_eval_bwd_kernel(const float* __restrict__ logp, const float* __restrict__ o1,
                                   const float* __restrict__ o2, const float* __restrict__ d1,
                                   const float* __restrict__ d2, const float* __restrict__ a_pano,
                                   const float* __restrict__ a_offset, const float* __restrict__ a_distance,
                                   const float* __restrict__ up, int R, int P, ActConsts ac, float* __restrict__ d_raw,
                                   float* __restrict__ d_o1, float* __restrict__ d_o2, float* __restrict__ d_d1,
                                   float* __restrict__ d_d2) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float* lp = logp + (size_t)r * (P + 1);
  const int a = (int)a_pano[r];
  const int idx = a % P;
  const float go = a != P ? 1.f : 0.f;
  const float dm = go * (float)ac.predict_distance, om = go * (float)ac.predict_offset;
  float Hp = 0.f;
  for (int j = 0; j <= P; ++j) Hp -= expf(lp[j]) * lp[j];
  const size_t rf = (size_t)r * P + idx;
  const Comp cd = comp_load(ac.nd, d1, d2, rf, ac.d_lo, ac.d_hi);
  const Comp co = comp_load(ac.no, o1, o2, rf, -ac.o_lim, ac.o_lim);
  const float* u = up + (size_t)r * 4;
  const int wd = ac.nd > 0 ? ac.nd : 1, wo = ac.no > 0 ? ac.no : 1;
  wp_row_grads(lp, a, idx, P, Hp, cd, co, a_distance[r], a_offset[r], dm, om, ac.nd > 0 ? 1.f : dm,
               ac.no > 0 ? 1.f : om, u[0], u[1], u[2], u[3], d_raw + (size_t)r * (P + 1), d_o1 + (size_t)r * P * wo, d_o2 + (size_t)r * P, d_d1 + (size_t)r * P * wd,
               d_d2 + (size_t)r * P);
}

struct PPOConsts {
  float clip, value_coef, entropy_coef, offset_reg, c_pano, c_offset, c_distance;
  int clipped_value;
};

// ------------------------------------------------------------------------------------
// evaluate_actions (waypoint_policy.py:265-333) fused with the WDDPPO mini-batch loss and its
// gradients (ddppo_alg.py:56-124), one CTA.  Rows r < R:
//   eval_out (R,4) = {action_log_prob, entropy_pano, entropy_offset, entropy_distance}
//   (W4: for a CATEGORICAL head the reference multiplies a (R,1) mask with the (R,) entropy, waypoint_policy.py:318-
//   325, i.e. builds an (R,R) matrix mask_i * H_j; every mean over it equals mean(mask) * mean(H).  eval_out then
//   carries the raw H_j and the caller forms the matrix; the loss below uses mean(mask) as the row weight.)
//   losses[8] = {loss, value_loss, action_loss, entropy_loss, offset_loss, mean entropy pano/offset/distance}
//   d_raw (R,P+1) w.r.t. the raw pano/stop logits, d_o1,d_o2,d_d1,d_d2 (non-zero only at the taken pano; layout as
//   in wp_eval_bwd_kernel), d_values (R).  Gradient outputs and the loss inputs may be NULL (plain evaluate_actions).
// ------------------------------------------------------------------------------------
__global__ void wp_ppo_loss_kernel(const float* __restrict__ logp, const float* __restrict__ o1,
                                   const float* __restrict__ o2, const float* __restrict__ d1,
                                   const float* __restrict__ d2, const float* __restrict__ values,
                                   const float* __restrict__ a_pano, const float* __restrict__ a_offset,
                                   const float* __restrict__ a_distance, const float* __restrict__ v_old,
                                   const float* __restrict__ ret, const float* __restrict__ lp_old,
                                   const float* __restrict__ adv, int R, int P, ActConsts ac, PPOConsts pc,
                                   float* __restrict__ eval_out, float* __restrict__ losses, float* __restrict__ d_raw,
                                   float* __restrict__ d_o1, float* __restrict__ d_o2, float* __restrict__ d_d1,
                                   float* __restrict__ d_d2, float* __restrict__ d_values) {
  __shared__ float red[32];
  const bool with_loss = losses != nullptr;
  const bool with_grad = d_raw != nullptr;
  const float invR = 1.f / (float)R;
  const int wd = ac.nd > 0 ? ac.nd : 1, wo = ac.no > 0 ? ac.no : 1;
  float s_go = 0.f;
  for (int r = threadIdx.x; r < R; r += blockDim.x) s_go += ((int)a_pano[r] != P) ? 1.f : 0.f;
  const float go_bar = block_sum(s_go, red) * invR;  // mean of the not-STOP mask (W4)
  __syncthreads();
  float s_val = 0.f, s_act = 0.f, s_ent = 0.f, s_off = 0.f, s_hp = 0.f, s_ho = 0.f, s_hd = 0.f;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    const float* lp = logp + (size_t)r * (P + 1);
    const int a = (int)a_pano[r];
    const int idx = a % P;
    const float go = a != P ? 1.f : 0.f;
    const float dm = go * (float)ac.predict_distance, om = go * (float)ac.predict_offset;
    const float dmh = ac.nd > 0 ? go_bar * (float)ac.predict_distance : dm;  // entropy weights of this row
    const float omh = ac.no > 0 ? go_bar * (float)ac.predict_offset : om;
    float Hp = 0.f;
    for (int j = 0; j <= P; ++j) Hp -= expf(lp[j]) * lp[j];
    const size_t rf = (size_t)r * P + idx;
    const Comp cd = comp_load(ac.nd, d1, d2, rf, ac.d_lo, ac.d_hi);
    const Comp co = comp_load(ac.no, o1, o2, rf, -ac.o_lim, ac.o_lim);
    const float y_d = a_distance[r], y_o = a_offset[r];
    const float alp = lp[a] + dm * comp_log_prob(cd, y_d) + om * comp_log_prob(co, y_o);
    const float Ho = omh * co.ent, Hd = dmh * cd.ent;
    if (eval_out != nullptr) {
      float* e = eval_out + (size_t)r * 4;
      e[0] = alp; e[1] = Hp; e[2] = ac.no > 0 ? co.ent : Ho; e[3] = ac.nd > 0 ? cd.ent : Hd;
    }
    if (!with_loss) continue;
    const float ratio = expf(alp - lp_old[r]);
    const float A = adv[r];
    const float rc = fminf(fmaxf(ratio, 1.f - pc.clip), 1.f + pc.clip);
    const float s1 = ratio * A, s2 = rc * A;
    s_act -= fminf(s1, s2);
    const float v = values[r], vo = v_old[r], rt = ret[r];
    const float diff = v - vo;
    const float vc = vo + fminf(fmaxf(diff, -pc.clip), pc.clip);
    const float l1 = (v - rt) * (v - rt), l2 = (vc - rt) * (vc - rt);
    s_val += pc.clipped_value ? 0.5f * fmaxf(l1, l2) : 0.5f * l1;
    s_ent += pc.c_pano * Hp + pc.c_offset * Ho + pc.c_distance * Hd;
    s_off += fabsf(fmaf(ac.off_b, y_o, ac.off_a));
    s_hp += Hp; s_ho += Ho; s_hd += Hd;
    if (!with_grad) continue;
    // d loss / d action_log_prob: torch.min routes to the smaller surrogate (ties inside the clip range
    // share one value); the clamp passes gradient only inside [1-clip, 1+clip]
    const bool inside = ratio >= 1.f - pc.clip && ratio <= 1.f + pc.clip;
    const float g_lp = (inside || s1 < s2) ? -invR * ratio * A : 0.f;
    const float g_ent = -pc.entropy_coef * invR;  // loss -= entropy_coef * mean(...)
    float dv;
    if (!pc.clipped_value || l1 >= l2) dv = v - rt;
    else dv = (diff > -pc.clip && diff < pc.clip) ? (vc - rt) : 0.f;
    d_values[r] = pc.value_coef * invR * dv;
    wp_row_grads(lp, a, idx, P, Hp, cd, co, y_d, y_o, dm, om, dmh, omh, g_lp, g_ent * pc.c_pano, g_ent * pc.c_offset,
                 g_ent * pc.c_distance, d_raw + (size_t)r * (P + 1), d_o1 + (size_t)r * P * wo, d_o2 + (size_t)r * P,
                 d_d1 + (size_t)r * P * wd, d_d2 + (size_t)r * P);
  }
  if (!with_loss) return;
  const float t_val = block_sum(s_val, red); __syncthreads();
  const float t_act = block_sum(s_act, red); __syncthreads();
  const float t_ent = block_sum(s_ent, red); __syncthreads();
  const float t_off = block_sum(s_off, red); __syncthreads();
  const float t_hp = block_sum(s_hp, red); __syncthreads();
  const float t_ho = block_sum(s_ho, red); __syncthreads();
  const float t_hd = block_sum(s_hd, red);
  if (threadIdx.x == 0) {
    const float value_loss = t_val * invR * pc.value_coef, action_loss = t_act * invR;
    const float entropy_loss = t_ent * invR * pc.entropy_coef, offset_loss = pc.offset_reg * t_off * invR;
    losses[0] = value_loss + action_loss + offset_loss - entropy_loss;
    losses[1] = value_loss; losses[2] = action_loss; losses[3] = entropy_loss; losses[4] = offset_loss;
    losses[5] = t_hp * invR; losses[6] = t_ho * invR; losses[7] = t_hd * invR;
  }
}

// ------------------------------------------------------------------------------------
// torch.nn.utils.clip_grad_norm_ (hb v0.1.7 PPO.before_step): g *= min(1, max_norm / (||g||_2 + 1e-6)).
// ws: >= 1026 floats; ws[1024] = ||g||, ws[1025] = the applied factor.
// ------------------------------------------------------------------------------------
__global__ void sumsq_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ ws) {
  __shared__ float red[32];
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    s = fmaf(g[i], g[i], s);
  s = block_sum(s, red);
  if (threadIdx.x == 0) ws[blockIdx.x] = s;
}
__global__ void clip_factor_kernel(float* __restrict__ ws, int parts, float max_norm, float pre_scale) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < parts; i += blockDim.x) s += ws[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const float norm = sqrtf(s) * pre_scale;
    ws[1024] = norm;
    ws[1025] = fminf(1.f, max_norm / (norm + 1e-6f)) * pre_scale;
  }
}
__global__ void scale_by_dev_kernel(float* __restrict__ g, long long n, const float* __restrict__ factor) {
  const float f = *factor;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    g[i] *= f;
}

// ------------------------------------------------------------------------------------
// ActionDictRolloutStorage.compute_returns (rollout_storage.py:127-152), one thread per env.
// rewards (T,N), value_preds/returns (T+1,N), masks (T+1,N) u8.
// ------------------------------------------------------------------------------------
__global__ void gae_returns_kernel(const float* __restrict__ rewards, float* __restrict__ value_preds,
                                   const unsigned char* __restrict__ masks, const float* __restrict__ next_value,
                                   float* __restrict__ returns, float* __restrict__ advantages, int T, int N,
                                   float gamma, float tau, int use_gae) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  if (use_gae) {
    value_preds[(size_t)T * N + n] = next_value[n];
    float gae = 0.f;
    for (int t = T - 1; t >= 0; --t) {
      const float m = (float)masks[(size_t)(t + 1) * N + n];
      const float delta = rewards[(size_t)t * N + n] + gamma * value_preds[(size_t)(t + 1) * N + n] * m -
                          value_preds[(size_t)t * N + n];
      gae = delta + gamma * tau * m * gae;
      returns[(size_t)t * N + n] = gae + value_preds[(size_t)t * N + n];
    }
  } else {
    returns[(size_t)T * N + n] = next_value[n];
    for (int t = T - 1; t >= 0; --t)
      returns[(size_t)t * N + n] = returns[(size_t)(t + 1) * N + n] * gamma * (float)masks[(size_t)(t + 1) * N + n] +
                                   rewards[(size_t)t * N + n];
  }
  if (advantages != nullptr)  // WDDPPO.get_advantages (ddppo_alg.py:30-32)
    for (int t = 0; t < T; ++t)
      advantages[(size_t)t * N + n] = returns[(size_t)t * N + n] - value_preds[(size_t)t * N + n];
}

}  // namespace
}  // namespace vlnce

using namespace vlnce;

extern "C" int vlnce_wp_stack_frames(const void* cur, const void* hist, const float* masks, void* out, int B, int P,
                                     long long frame_bytes, cudaStream_t stream) {
  VLNCE_CHECK_ARG(cur && hist && masks && out && B > 0 && P > 0 && frame_bytes > 0, "wp_stack_frames: bad arguments");
  const bool vec = frame_bytes % 16 == 0 && (((uintptr_t)cur | (uintptr_t)hist | (uintptr_t)out) & 15) == 0;
  if (vec) {
    const long long fv = frame_bytes / 16;
    stack_frames_kernel<uint4><<<grid_for((long long)B * (P + 1) * fv, 256), 256, 0, stream>>>(
        (const uint4*)cur, (const uint4*)hist, masks, (uint4*)out, (long long)B * P * fv, fv, B);
  } else {
    stack_frames_kernel<unsigned char><<<grid_for((long long)B * (P + 1) * frame_bytes, 256), 256, 0, stream>>>(
        (const unsigned char*)cur, (const unsigned char*)hist, masks, (unsigned char*)out,
        (long long)B * P * frame_bytes, frame_bytes, B);
  }
  VLNCE_LAUNCH_CHECK("stack_frames_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_broadcast_positions(float* E, long long lde_r, long long lde_p, int R, int P, int C,
                                         cudaStream_t stream) {
  VLNCE_CHECK_ARG(E && R > 0 && P > 0 && C > 0, "broadcast_positions: bad arguments");
  if (P == 1) return VLNCE_OK;
  broadcast_positions_kernel<<<grid_for((long long)R * (P - 1) * C, 256), 256, 0, stream>>>(E, lde_r, lde_p, R, P, C);
  VLNCE_LAUNCH_CHECK("broadcast_positions_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_wp_prev_action(const long long* pano, const float* offset, const float* distance,
                                    const float* masks, float* out, long long ldo, float* out2, long long ldo2,
                                    int rows, int panos, float off_a, float off_b, float dist_a, float dist_b,
                                    cudaStream_t stream) {
  VLNCE_CHECK_ARG(pano && offset && distance && masks && out && rows > 0 && panos > 0, "wp_prev_action: bad arguments");
  wp_prev_action_kernel<<<cdiv(rows, 128), 128, 0, stream>>>(pano, offset, distance, masks, out, ldo, out2, ldo2, rows,
                                                            kTwoPi / (float)panos, off_a, off_b, dist_a, dist_b);
  VLNCE_LAUNCH_CHECK("wp_prev_action_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_wp_inst_attn_fwd(const float* q, long long ldq, const float* bk, const float* instr,
                                      const int* lengths, float scale, float* out, long long ldo, float* save, int R,
                                      int L, int D, int Dk, cudaStream_t stream) {
  VLNCE_CHECK_ARG(q && bk && instr && lengths && out && save && R > 0 && L > 0 && D > 0 && Dk > 0,
                  "wp_inst_attn_fwd: bad arguments");
  wp_inst_attn_fwd_kernel<<<R, 256, 0, stream>>>(q, ldq, bk, instr, lengths, scale, out, ldo, save, R, L, D, Dk);
  VLNCE_LAUNCH_CHECK("wp_inst_attn_fwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_wp_inst_attn_bwd(const float* save, const float* out, long long ldo, const float* dout,
                                      long long lddo, const int* lengths, float scale, float* dcs, float* d_instr,
                                      int R, int L, int D, cudaStream_t stream) {
  VLNCE_CHECK_ARG(save && out && dout && lengths && dcs && d_instr && R > 0 && L > 0 && D > 0,
                  "wp_inst_attn_bwd: bad arguments");
  wp_inst_attn_bwd_kernel<<<R, 256, 0, stream>>>(save, out, ldo, dout, lddo, lengths, scale, dcs, d_instr, L, D);
  VLNCE_LAUNCH_CHECK("wp_inst_attn_bwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_wp_scramble_query(const float* q, float* out, long long n_in, int P, cudaStream_t stream) {
  VLNCE_CHECK_ARG(q && out && n_in > 0 && P > 0, "wp_scramble_query: bad arguments");
  wp_scramble_query_kernel<<<grid_for(n_in * P, 256), 256, 0, stream>>>(q, out, n_in, P);
  VLNCE_LAUNCH_CHECK("wp_scramble_query_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_wp_scramble_query_bwd(const float* dout, float* dq, long long n_in, int P, cudaStream_t stream) {
  VLNCE_CHECK_ARG(dout && dq && n_in > 0 && P > 0, "wp_scramble_query_bwd: bad arguments");
  wp_scramble_query_bwd_kernel<<<grid_for(n_in, 256), 256, 0, stream>>>(dout, dq, n_in, P);
  VLNCE_LAUNCH_CHECK("wp_scramble_query_bwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, float* y,
                                   long long ldy, float* save, int R, int D, float eps, cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && gamma && beta && y && R > 0 && D > 0, "layernorm_fwd: bad arguments");
  layernorm_fwd_kernel<<<R, 128, 0, stream>>>(x, ldx, gamma, beta, y, ldy, save, D, eps);
  VLNCE_LAUNCH_CHECK("layernorm_fwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_layernorm_bwd(const float* dy, long long lddy, const float* x, long long ldx, const float* gamma,
                                   const float* save, float* dx, long long lddx, float* prod, long long ldp, int R,
                                   int D, cudaStream_t stream) {
  VLNCE_CHECK_ARG(dy && x && gamma && save && dx && prod && R > 0 && D > 0, "layernorm_bwd: bad arguments");
  layernorm_bwd_kernel<<<R, 128, 0, stream>>>(dy, lddy, x, ldx, gamma, save, dx, lddx, prod, ldp, D);
  VLNCE_LAUNCH_CHECK("layernorm_bwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_wp_dot_frames_fwd(const float* F, long long ldf_r, long long ldf_p, const float* xs,
                                       long long ldx, float* out, long long ldo, int B, int P, int D,
                                       cudaStream_t stream) {
  VLNCE_CHECK_ARG(F && xs && out && B > 0 && P > 0 && D > 0, "wp_dot_frames_fwd: bad arguments");
  wp_dot_frames_fwd_kernel<<<cdiv((long long)B * P, 8), 256, 0, stream>>>(F, ldf_r, ldf_p, xs, ldx, out, ldo, B, P, D);
  VLNCE_LAUNCH_CHECK("wp_dot_frames_fwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_wp_dot_frames_bwd(const float* F, long long ldf_r, long long ldf_p, const float* xs,
                                       long long ldx, const float* dout, long long lddo, float* dF, long long lddf_r,
                                       long long lddf_p, int accumulate_dF, float* dxs, long long lddx, int B, int P,
                                       int D, cudaStream_t stream) {
  VLNCE_CHECK_ARG(F && xs && dout && dF && dxs && B > 0 && P > 0 && D > 0, "wp_dot_frames_bwd: bad arguments");
  wp_dot_frames_bwd_kernel<<<grid_for((long long)B * D, 256), 256, 0, stream>>>(
      F, ldf_r, ldf_p, xs, ldx, dout, lddo, dF, lddf_r, lddf_p, accumulate_dF, dxs, lddx, B, P, D);
  VLNCE_LAUNCH_CHECK("wp_dot_frames_bwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_broadcast_rows(const float* x, long long ldx, float* out, long long ldo_r, long long ldo_p, int B,
                                    int P, int D, cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && out && B > 0 && P > 0 && D > 0, "broadcast_rows: bad arguments");
  broadcast_rows_kernel<<<grid_for((long long)B * P * D, 256), 256, 0, stream>>>(x, ldx, out, ldo_r, ldo_p, B, P, D);
  VLNCE_LAUNCH_CHECK("broadcast_rows_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_sum_positions(const float* E, long long lde_r, long long lde_p, float* out, long long ldo,
                                   int accumulate, int R, int P, int D, cudaStream_t stream) {
  VLNCE_CHECK_ARG(E && out && R > 0 && P > 0 && D > 0, "sum_positions: bad arguments");
  sum_positions_kernel<<<grid_for((long long)R * D, 256), 256, 0, stream>>>(E, lde_r, lde_p, out, ldo, accumulate, R,
                                                                          P, D);
  VLNCE_LAUNCH_CHECK("sum_positions_kernel");
  return VLNCE_OK;
}

static HeadConsts head_consts(const float* c) {
  HeadConsts h;
  h.d_span = c[0]; h.d_min = c[1]; h.dv_span = c[2]; h.dv_min = c[3];
  h.o_scale = c[4]; h.inv_t = c[5]; h.ov_span = c[6]; h.ov_min = c[7];
  return h;
}

extern "C" int vlnce_wp_head_act_fwd(const float* raw, float* d1, float* d2, float* o1, float* o2, long long n,
                                     const float* consts_host, cudaStream_t stream) {
  VLNCE_CHECK_ARG(raw && d1 && d2 && o1 && o2 && n > 0 && consts_host, "wp_head_act_fwd: bad arguments");
  wp_head_act_fwd_kernel<<<grid_for(n, 256), 256, 0, stream>>>(raw, d1, d2, o1, o2, n, head_consts(consts_host));
  VLNCE_LAUNCH_CHECK("wp_head_act_fwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_wp_head_act_bwd(const float* raw, const float* g_d1, const float* g_d2, const float* g_o1,
                                     const float* g_o2, float* g_raw, long long n, const float* consts_host,
                                     cudaStream_t stream) {
  VLNCE_CHECK_ARG(raw && g_d1 && g_d2 && g_o1 && g_o2 && g_raw && n > 0 && consts_host, "wp_head_act_bwd: bad arguments");
  wp_head_act_bwd_kernel<<<grid_for(n, 256), 256, 0, stream>>>(raw, g_d1, g_d2, g_o1, g_o2, g_raw, n,
                                                              head_consts(consts_host));
  VLNCE_LAUNCH_CHECK("wp_head_act_bwd_kernel");
  return VLNCE_OK;
}

static ActConsts act_consts(const float* c) {
  ActConsts a;
  a.d_lo = c[0]; a.d_hi = c[1]; a.o_lim = c[2];
  a.predict_distance = c[3] != 0.f; a.predict_offset = c[4] != 0.f;
  a.nd = (int)c[5]; a.no = (int)c[6];
  a.dist_a = c[7]; a.dist_b = c[8]; a.off_a = c[9]; a.off_b = c[10];
  return a;
}

extern "C" int vlnce_wp_act(const float* logp, const float* o1, const float* o2, const float* d1, const float* d2,
                            long long* pano_stop, float* fout, int B, int P, int deterministic,
                            unsigned long long seed, unsigned long long offset, const float* act_consts_host,
                            cudaStream_t stream) {
  VLNCE_CHECK_ARG(logp && o1 && d1 && pano_stop && fout && B > 0 && P > 0 && act_consts_host, "wp_act: bad arguments");
  VLNCE_CHECK_ARG((act_consts_host[5] != 0.f || d2) && (act_consts_host[6] != 0.f || o2),
                  "wp_act: a continuous head needs its variance tensor");
  wp_act_kernel<<<cdiv(B, 128), 128, 0, stream>>>(logp, o1, o2, d1, d2, pano_stop, fout, B, P, deterministic, seed,
                                                 offset, act_consts(act_consts_host));
  VLNCE_LAUNCH_CHECK("wp_act_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_wp_ppo_loss(const float* logp, const float* o1, const float* o2, const float* d1, const float* d2,
                                 const float* values, const float* a_pano, const float* a_offset,
                                 const float* a_distance, const float* v_old, const float* ret, const float* lp_old,
                                 const float* adv, int R, int P, const float* act_consts_host,
                                 const float* ppo_consts_host, float* eval_out, float* losses, float* d_raw, float* d_o1,
                                 float* d_o2, float* d_d1, float* d_d2, float* d_values, cudaStream_t stream) {
  VLNCE_CHECK_ARG(logp && o1 && d1 && a_pano && a_offset && a_distance && R > 0 && P > 0 && act_consts_host,
                  "wp_ppo_loss: bad arguments");
  const bool cont_d = act_consts_host[5] == 0.f, cont_o = act_consts_host[6] == 0.f;
  VLNCE_CHECK_ARG((!cont_d || d2) && (!cont_o || o2), "wp_ppo_loss: a continuous head needs its variance tensor");
  VLNCE_CHECK_ARG(losses == nullptr || (values && v_old && ret && lp_old && adv && ppo_consts_host),
                  "wp_ppo_loss: the loss needs values, value_preds, returns, old log-probs, advantages and constants");
  VLNCE_CHECK_ARG(d_raw == nullptr || (losses && d_o1 && d_d1 && d_values && (!cont_d || d_d2) && (!cont_o || d_o2)),
                  "wp_ppo_loss: gradient outputs come as a full set and need the loss");
  PPOConsts pc;
  memset(&pc, 0, sizeof(pc));
  if (ppo_consts_host != nullptr) {
    const float* c = ppo_consts_host;
    pc.clip = c[0]; pc.value_coef = c[1]; pc.entropy_coef = c[2]; pc.offset_reg = c[3];
    pc.c_pano = c[4]; pc.c_offset = c[5]; pc.c_distance = c[6]; pc.clipped_value = c[7] != 0.f;
  }
  wp_ppo_loss_kernel<<<1, 256, 0, stream>>>(logp, o1, o2, d1, d2, values, a_pano, a_offset, a_distance, v_old, ret,
                                            lp_old, adv, R, P, act_consts(act_consts_host), pc, eval_out, losses, d_raw,
                                            d_o1, d_o2, d_d1, d_d2, d_values);
  VLNCE_LAUNCH_CHECK("wp_ppo_loss_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_wp_eval_bwd(const float* logp, const float* o1, const float* o2, const float* d1, const float* d2,
                                 const float* a_pano, const float* a_offset, const float* a_distance,
                                 const float* upstream, int R, int P, const float* act_consts_host, float* d_raw,
                                 float* d_o1, float* d_o2, float* d_d1, float* d_d2, cudaStream_t stream) {
  VLNCE_CHECK_ARG(logp && o1 && d1 && a_pano && a_offset && a_distance && upstream && R > 0 && P > 0 &&
                      act_consts_host && d_raw && d_o1 && d_d1,
                  "wp_eval_bwd: bad arguments");
  VLNCE_CHECK_ARG((act_consts_host[5] != 0.f || (d2 && d_d2)) && (act_consts_host[6] != 0.f || (o2 && d_o2)),
                  "wp_eval_bwd: a continuous head needs its variance tensor and gradient");
  wp_eval_bwd_kernel<<<cdiv(R, 128), 128, 0, stream>>>(logp, o1, o2, d1, d2, a_pano, a_offset, a_distance, upstream, R,
                                                      P, act_consts(act_consts_host), d_raw, d_o1, d_o2, d_d1, d_d2);
  VLNCE_LAUNCH_CHECK("wp_eval_bwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_clip_grad_norm(float* g, long long n, float max_norm, float pre_scale, float* ws,
                                    size_t ws_bytes, cudaStream_t stream) {
  VLNCE_CHECK_ARG(g && n > 0 && ws && ws_bytes >= 1026 * sizeof(float) && max_norm > 0.f, "clip_grad_norm: bad arguments");
  const int parts = (int)grid_for(n, 256) > 1024 ? 1024 : (int)grid_for(n, 256);
  sumsq_partial_kernel<<<parts, 256, 0, stream>>>(g, n, ws);
  VLNCE_LAUNCH_CHECK("sumsq_partial_kernel");
  clip_factor_kernel<<<1, 256, 0, stream>>>(ws, parts, max_norm, pre_scale);
  VLNCE_LAUNCH_CHECK("clip_factor_kernel");
  scale_by_dev_kernel<<<grid_for(n, 256), 256, 0, stream>>>(g, n, ws + 1025);
  VLNCE_LAUNCH_CHECK("scale_by_dev_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_gae_returns(const float* rewards, float* value_preds, const unsigned char* masks,
                                 const float* next_value, float* returns, float* advantages, int T, int N, float gamma,
                                 float tau, int use_gae, cudaStream_t stream) {
  VLNCE_CHECK_ARG(rewards && value_preds && masks && next_value && returns && T > 0 && N > 0, "gae_returns: bad arguments");
  gae_returns_kernel<<<cdiv(N, 128), 128, 0, stream>>>(rewards, value_preds, masks, next_value, returns, advantages, T,
                                                      N, gamma, tau, use_gae);
  VLNCE_LAUNCH_CHECK("gae_returns_kernel");
  return VLNCE_OK;
}
