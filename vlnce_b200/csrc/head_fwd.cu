// Forward kernels of the trainable CMA / Seq2Seq head (everything after the conv trunks):
// token embedding + lengths, persistent (bi)LSTM instruction encoder, GRU state-encoder
// cell, single-query attention pooling, spatial-embedding fill, small layout helpers.
// All fp32, HBM/latency-bound: coalesced channel-fastest loads, weights of the recurrent
// step held in registers across the whole sequence.
#include "common.cuh"

namespace vlnce {
namespace {

// ------------------------------------------------------------------------------------
// instruction_encoder.py:70-78 — embedding gather and per-row length
// ------------------------------------------------------------------------------------
__global__ void embedding_fwd_kernel(const long long* __restrict__ ids, const float* __restrict__ table,
                                     float* __restrict__ out, long long rows, int dim, int ldo,
                                     int vocab) {
  const long long total = rows * dim;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / dim;
    const int d = (int)(idx % dim);
    long long id = ids[r];
    if (id < 0 || id >= vocab) id = 0;
    out[r * ldo + d] = __ldg(table + id * dim + d);
  }
}

// prev-action index, Q6 (cma_policy.py:233-235): idx = ((a + 1) * mask), then gather 32-d rows
__global__ void prev_action_embed_kernel(const long long* __restrict__ prev, const float* __restrict__ masks,
                                         const float* __restrict__ table, float* __restrict__ out,
                                         float* __restrict__ out2, int rows, int dim, int ldo, int ldo2,
                                         int num_rows_table) {
  const int total = rows * dim;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int r = idx / dim, d = idx % dim;
    long long id = (long long)(((float)prev[r] + 1.f) * masks[r]);
    if (id < 0 || id >= num_rows_table) id = 0;
    const float v = __ldg(table + id * dim + d);
    out[(size_t)r * ldo + d] = v;
    if (out2 != nullptr) out2[(size_t)r * ldo2 + d] = v;
  }
}

__global__ void token_lengths_kernel(const long long* __restrict__ ids, int* __restrict__ lengths,
                                     int rows, int L) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  int cnt = 0;
  for (int t = threadIdx.x & 31; t < L; t += 32) cnt += ids[(size_t)r * L + t] != 0;
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) lengths[r] = cnt;
}

// rxr features (R,L,D): a position counts if its row has any non-zero entry
__global__ void feature_lengths_kernel(const float* __restrict__ x, int* __restrict__ lengths, int L,
                                       int D) {
  const int r = blockIdx.x;
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  int local = 0;
  for (int t = warp; t < L; t += nw) {
    const float* row = x + ((size_t)r * L + t) * D;
    int any = 0;
    for (int d = lane; d < D; d += 32) any |= (row[d] != 0.f);
    any = __any_sync(0xffffffffu, any);
    local += any;
  }
  if (lane == 0) atomicAdd(&cnt, local);
  __syncthreads();
  if (threadIdx.x == 0) lengths[r] = cnt;
}

// ------------------------------------------------------------------------------------
// Persistent LSTM recurrence (instruction_encoder.py:80-84, torch.nn.LSTM on a packed
// sequence): one CTA per (group of NR rows, direction), 4H threads; thread j keeps row j of
// W_hh on chip for all L steps and applies it to the NR rows of its group in every step, so
// the weights are read once per step for NR rows and the (latency-bound) recurrence occupies
// R/NR SMs instead of R.  gx = x W_ih^T + b_ih is precomputed by one GEMM; the value of the
// NEXT step is fetched while the current step computes.  Gate order i, f, g, o (PyTorch).
// Steps t >= len are skipped (outputs stay zero); the reverse direction starts at each row's
// own last token.  Per-row arithmetic (summation order) does not depend on NR.
// ------------------------------------------------------------------------------------
template <int H, int NR>
__global__ void __launch_bounds__(4 * H, 1)
lstm_seq_fwd_kernel(const float* __restrict__ gx,   // (dirs, R, L, 4H) = x W_ih^T + b_ih
                    const float* __restrict__ whh_f, const float* __restrict__ whh_b,  // (4H, H) per direction
                    const float* __restrict__ bhh_f, const float* __restrict__ bhh_b,  // (4H) per direction
                    const int* __restrict__ lengths, float* __restrict__ out,  // (R, L, dirs*H)
                    float* __restrict__ hn,          // (dirs, R, H) or null
                    float* __restrict__ save_gates,  // (dirs, R, L, 4H) activated gates or null
                    float* __restrict__ save_c,      // (dirs, R, L, H) or null
                    int R, int L, int dirs) {
  // Row j of W_hh: columns [0, KR) live in registers, [KR, H) in shared memory as float4
  // groups laid out [k/4][j] (conflict-free 16B reads) — 4H threads cannot hold H registers each.
  constexpr int KR = H / 2;
  constexpr int KS4 = (H - KR) / 4;
  extern __shared__ __align__(16) float lstm_smem[];
  float4* w_s = reinterpret_cast<float4*>(lstm_smem);           // KS4 * 4H float4
  float* h_s = lstm_smem + (size_t)KS4 * 4 * H * 4;              // NR x H
  float* g_s = h_s + NR * H;                                     // NR x 4H
  const int r0 = blockIdx.x * NR, dir = blockIdx.y;
  const int j = threadIdx.x;  // gate row
  float w[KR];
  {
    const float4* wp = reinterpret_cast<const float4*>((dir == 0 ? whh_f : whh_b) + (size_t)j * H);
#pragma unroll
    for (int k = 0; k < KR / 4; ++k) {
      const float4 t4 = __ldg(wp + k);
      w[4 * k] = t4.x; w[4 * k + 1] = t4.y; w[4 * k + 2] = t4.z; w[4 * k + 3] = t4.w;
    }
#pragma unroll
    for (int k = 0; k < KS4; ++k) w_s[k * 4 * H + j] = __ldg(wp + KR / 4 + k);
  }
  int len[NR];
  int maxlen = 0;
#pragma unroll
  for (int rr = 0; rr < NR; ++rr) {
    len[rr] = (r0 + rr < R) ? min(lengths[r0 + rr], L) : 0;
    maxlen = max(maxlen, len[rr]);
  }
  const float bh = __ldg((dir == 0 ? bhh_f : bhh_b) + j);
  float c[NR];
#pragma unroll
  for (int rr = 0; rr < NR; ++rr) {
    c[rr] = 0.f;
    if (j < H) h_s[rr * H + j] = 0.f;
  }
  // gx of (row rr, step): time index t = step (forward) | len-1-step (reverse)
  auto gx_at = [&](int rr, int step) -> float {
    if (step >= len[rr]) return 0.f;
    const int t = dir == 0 ? step : len[rr] - 1 - step;
    return __ldg(gx + (((size_t)dir * R + r0 + rr) * L + t) * 4 * H + j);
  };
  float gnext[NR];
#pragma unroll
  for (int rr = 0; rr < NR; ++rr) gnext[rr] = gx_at(rr, 0);
  __syncthreads();
  for (int step = 0; step < maxlen; ++step) {
    float a0[NR], a1[NR], a2[NR], a3[NR];
#pragma unroll
    for (int rr = 0; rr < NR; ++rr) {
      a0[rr] = gnext[rr] + bh; a1[rr] = 0.f; a2[rr] = 0.f; a3[rr] = 0.f;
      gnext[rr] = gx_at(rr, step + 1);   // in flight during this step's dot products
    }
#pragma unroll
    for (int k = 0; k < KR; k += 4) {
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) {
        const float4 h4 = *reinterpret_cast<const float4*>(&h_s[rr * H + k]);
        a0[rr] = fmaf(w[k], h4.x, a0[rr]);
        a1[rr] = fmaf(w[k + 1], h4.y, a1[rr]);
        a2[rr] = fmaf(w[k + 2], h4.z, a2[rr]);
        a3[rr] = fmaf(w[k + 3], h4.w, a3[rr]);
      }
    }
#pragma unroll
    for (int k = 0; k < KS4; ++k) {
      const float4 w4 = w_s[k * 4 * H + j];
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) {
        const float4 h4 = *reinterpret_cast<const float4*>(&h_s[rr * H + KR + 4 * k]);
        a0[rr] = fmaf(w4.x, h4.x, a0[rr]);
        a1[rr] = fmaf(w4.y, h4.y, a1[rr]);
        a2[rr] = fmaf(w4.z, h4.z, a2[rr]);
        a3[rr] = fmaf(w4.w, h4.w, a3[rr]);
      }
    }
#pragma unroll
    for (int rr = 0; rr < NR; ++rr) {
      if (step < len[rr]) {
        const float pre = (a0[rr] + a1[rr]) + (a2[rr] + a3[rr]);
        const float act = (j >= 2 * H && j < 3 * H) ? tanhf(pre) : sigmoid_precise(pre);
        g_s[rr * 4 * H + j] = act;
        if (save_gates != nullptr) {
          const int t = dir == 0 ? step : len[rr] - 1 - step;
          save_gates[(((size_t)dir * R + r0 + rr) * L + t) * 4 * H + j] = act;
        }
      }
    }
    __syncthreads();
    if (j < H) {
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) {
        if (step < len[rr]) {
          const float* g = g_s + rr * 4 * H;
          const float ig = g[j], fg = g[H + j], gg = g[2 * H + j], og = g[3 * H + j];
          c[rr] = fmaf(fg, c[rr], ig * gg);
          const float h = og * tanhf(c[rr]);
          h_s[rr * H + j] = h;
          const int t = dir == 0 ? step : len[rr] - 1 - step;
          out[((size_t)(r0 + rr) * L + t) * (dirs * H) + dir * H + j] = h;
          if (save_c != nullptr) save_c[(((size_t)dir * R + r0 + rr) * L + t) * H + j] = c[rr];
        }
      }
    }
    __syncthreads();
  }
  if (hn != nullptr && j < H) {
#pragma unroll
    for (int rr = 0; rr < NR; ++rr)
      if (r0 + rr < R) hn[((size_t)dir * R + r0 + rr) * H + j] = h_s[rr * H + j];
  }
}

// ------------------------------------------------------------------------------------
// Persistent GRU recurrence (instruction_encoder.py:27,80-84 with INSTRUCTION_ENCODER.rnn_type = "GRU":
// torch.nn.GRU on a packed sequence): one CTA per (row, direction), 3H threads; thread j owns row j of W_hh
// (half in registers, half in shared memory) for all L steps.  gx = x W_ih^T + b_ih comes from one GEMM; b_hh
// is added here because the candidate gate needs r * (W_hn h + b_hn).  Gate order r, z, n (PyTorch).
// Steps t >= len are skipped (outputs stay zero); the reverse direction starts at each row's last token.
// ------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(3 * H, 1)
gru_seq_fwd_kernel(const float* __restrict__ gx,   // (dirs, R, L, 3H)
                   const float* __restrict__ whh_f, const float* __restrict__ whh_b,  // (3H, H) per direction
                   const float* __restrict__ bhh_f, const float* __restrict__ bhh_b,  // (3H) per direction
                   const int* __restrict__ lengths, float* __restrict__ out,  // (R, L, dirs*H)
                   float* __restrict__ hn,     // (dirs, R, H) or null
                   float* __restrict__ save,   // (dirs, R, L, 4H): r, z, n, W_hn h + b_hn, or null
                   int R, int L, int dirs) {
  constexpr int G = 3 * H;
  constexpr int KR = H / 2;
  constexpr int KS4 = (H - KR) / 4;
  extern __shared__ __align__(16) float gru_smem[];
  float4* w_s = reinterpret_cast<float4*>(gru_smem);  // KS4 * G float4, [k/4][j]
  float* h_s = gru_smem + (size_t)KS4 * G * 4;         // H
  float* g_s = h_s + H;                                // G: W_hh h + b_hh of this step
  const int r = blockIdx.x, dir = blockIdx.y;
  const int j = threadIdx.x;
  float w[KR];
  {
    const float4* wp = reinterpret_cast<const float4*>((dir == 0 ? whh_f : whh_b) + (size_t)j * H);
#pragma unroll
    for (int k = 0; k < KR / 4; ++k) {
      const float4 t4 = __ldg(wp + k);
      w[4 * k] = t4.x; w[4 * k + 1] = t4.y; w[4 * k + 2] = t4.z; w[4 * k + 3] = t4.w;
    }
#pragma unroll
    for (int k = 0; k < KS4; ++k) w_s[k * G + j] = __ldg(wp + KR / 4 + k);
  }
  const int len = min(lengths[r], L);
  const float bh = __ldg((dir == 0 ? bhh_f : bhh_b) + j);
  if (j < H) h_s[j] = 0.f;
  __syncthreads();
  const size_t row_base = ((size_t)dir * R + r) * L;
  for (int step = 0; step < len; ++step) {
    const int t = dir == 0 ? step : len - 1 - step;
    float a0 = bh, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int k = 0; k < KR; k += 4) {
      const float4 h4 = *reinterpret_cast<const float4*>(&h_s[k]);
      a0 = fmaf(w[k], h4.x, a0);
      a1 = fmaf(w[k + 1], h4.y, a1);
      a2 = fmaf(w[k + 2], h4.z, a2);
      a3 = fmaf(w[k + 3], h4.w, a3);
    }
#pragma unroll
    for (int k = 0; k < KS4; ++k) {
      const float4 w4 = w_s[k * G + j];
      const float4 h4 = *reinterpret_cast<const float4*>(&h_s[KR + 4 * k]);
      a0 = fmaf(w4.x, h4.x, a0);
      a1 = fmaf(w4.y, h4.y, a1);
      a2 = fmaf(w4.z, h4.z, a2);
      a3 = fmaf(w4.w, h4.w, a3);
    }
    g_s[j] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (j < H) {
      const float* x = gx + (row_base + t) * G;
      const float rg = sigmoid_precise(__ldg(x + j) + g_s[j]);
      const float zg = sigmoid_precise(__ldg(x + H + j) + g_s[H + j]);
      const float ghn = g_s[2 * H + j];
      const float ng = tanhf(fmaf(rg, ghn, __ldg(x + 2 * H + j)));
      const float h = fmaf(zg, h_s[j] - ng, ng);  // (1-z) n + z h
      h_s[j] = h;
      out[((size_t)r * L + t) * (dirs * H) + dir * H + j] = h;
      if (save != nullptr) {
        float* sv = save + (row_base + t) * 4 * H;
        sv[j] = rg; sv[H + j] = zg; sv[2 * H + j] = ng; sv[3 * H + j] = ghn;
      }
    }
    __syncthreads();
  }
  if (hn != nullptr && j < H) hn[((size_t)dir * R + r) * H + j] = h_s[j];
}

// ------------------------------------------------------------------------------------
// GRU cell pointwise (habitat_baselines v0.1.7 rnn_state_encoder single step; torch.nn.GRU):
//   r = s(gi_r + gh_r), z = s(gi_z + gh_z), n = tanh(gi_n + r * gh_n), h' = (1-z) n + z h
// gi = x W_ih^T + b_ih, gh = h W_hh^T + b_hh (both (rows, 3H)) come from the GEMM; h is the
// already mask-reset hidden state.
// ------------------------------------------------------------------------------------
__global__ void gru_pointwise_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                                         const float* __restrict__ h, long long ldh,
                                         float* __restrict__ hout, long long ldo,
                                         float* __restrict__ hout2, long long ldo2,
                                         float* __restrict__ save,  // (rows, 4H): r, z, n, gh_n or null
                                         int rows, int H) {
  const int total = rows * H;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int b = idx / H, k = idx % H;
    const float* gib = gi + (size_t)b * 3 * H;
    const float* ghb = gh + (size_t)b * 3 * H;
    const float rg = sigmoid_precise(gib[k] + ghb[k]);
    const float zg = sigmoid_precise(gib[H + k] + ghb[H + k]);
    const float ghn = ghb[2 * H + k];
    const float ng = tanhf(fmaf(rg, ghn, gib[2 * H + k]));
    const float hp = h[b * ldh + k];
    const float hn = fmaf(zg, hp - ng, ng);  // (1-z) n + z h
    hout[b * ldo + k] = hn;
    if (hout2 != nullptr) hout2[b * ldo2 + k] = hn;
    if (save != nullptr) {
      float* s = save + (size_t)b * 4 * H;
      s[k] = rg; s[H + k] = zg; s[2 * H + k] = ng; s[3 * H + k] = ghn;
    }
  }
}

// ------------------------------------------------------------------------------------
// One GRU step of the state encoder in ONE launch (cma_policy.py:248-256, 287-294; seq2seq_policy.py:112-137 via
// habitat_baselines v0.1.7 RNNStateEncoder.single_forward): hidden reset (h * mask), input projection, hidden
// projection, gates and the state update.  CTA j owns UN hidden units = 3*UN rows of W_ih and of W_hh, staged once in
// shared memory; a warp takes 4 batch rows at a time, its lanes split the reduction dimension (coalesced reads of x and
// h, conflict-free reads of the weight slice), 16 partial sums per row are shuffled down and lanes 0..15 finish the
// cell.  Roles inside the CTA: all warps stage, then every warp is a compute warp over its row groups.
// ------------------------------------------------------------------------------------
constexpr int kGruUN = 4;        // hidden units per CTA
constexpr int kGruRows = 4;      // batch rows per warp pass

template <bool VEC>   // VEC: rows, leading dimensions and pointers are 16-byte aligned -> 128-bit loads, 4 k per lane
__global__ void __launch_bounds__(512)
gru_step_fused_kernel(const float* __restrict__ x, long long ldx, int Kx, const float* __restrict__ h0, long long ldh0,
                      const float* __restrict__ mask, const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                      const float* __restrict__ b_ih, const float* __restrict__ b_hh, float* __restrict__ hm,
                      float* __restrict__ out, long long ldo, float* __restrict__ out2, long long ldo2,
                      float* __restrict__ save, int rows, int H) {
  extern __shared__ __align__(16) float smem_gru[];
  constexpr int UN = kGruUN, G = 3 * kGruUN, RW = kGruRows;
  float* s_wi = smem_gru;                   // [G][Kx]   slice row g*UN + u = W_ih[g*H + u0 + u]
  float* s_wh = smem_gru + (size_t)G * Kx;  // [G][H]
  const int u0 = blockIdx.x * UN;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int j = warp; j < 2 * G; j += nwarps) {   // one warp per weight row: coalesced, no index division per element
    const int g = (j % G) / UN, u = (j % G) % UN;
    if (j < G) {
      const float* src = w_ih + (size_t)(g * H + u0 + u) * Kx;
      for (int k = lane; k < Kx; k += 32) s_wi[j * Kx + k] = __ldg(src + k);
    } else {
      const float* src = w_hh + (size_t)(g * H + u0 + u) * H;
      for (int k = lane; k < H; k += 32) s_wh[(j - G) * H + k] = __ldg(src + k);
    }
  }
  __syncthreads();
  const int groups = (rows + RW - 1) / RW;
  for (int grp = warp; grp < groups; grp += nwarps) {
    const int r0 = grp * RW;
    // per row: [u] r-sum (gi+gh), [UN+u] z-sum, [2UN+u] gi_n, [3UN+u] gh_n
    float acc[RW][4 * UN];
#pragma unroll
    for (int r = 0; r < RW; ++r)
#pragma unroll
      for (int j = 0; j < 4 * UN; ++j) acc[r][j] = 0.f;
    const float* xr[RW];
    const float* hr[RW];
    float mk[RW];
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      const int row = r0 + r < rows ? r0 + r : rows - 1;
      xr[r] = x + (size_t)row * ldx;
      hr[r] = h0 + (size_t)row * ldh0;
      mk[r] = mask[row];
    }
    if constexpr (VEC) {
#pragma unroll 1
      for (int k = 4 * lane; k < Kx; k += 128) {
        float4 xv[RW];
#pragma unroll
        for (int r = 0; r < RW; ++r) xv[r] = *reinterpret_cast<const float4*>(xr[r] + k);
#pragma unroll
        for (int j = 0; j < G; ++j) {
          const float4 w = *reinterpret_cast<const float4*>(s_wi + j * Kx + k);
#pragma unroll
          for (int r = 0; r < RW; ++r)
            acc[r][j] = fmaf(xv[r].x, w.x, fmaf(xv[r].y, w.y, fmaf(xv[r].z, w.z, fmaf(xv[r].w, w.w, acc[r][j]))));
        }
      }
#pragma unroll 1
      for (int k = 4 * lane; k < H; k += 128) {
        float4 hv[RW];
#pragma unroll
        for (int r = 0; r < RW; ++r) {
          hv[r] = *reinterpret_cast<const float4*>(hr[r] + k);
          hv[r].x *= mk[r]; hv[r].y *= mk[r]; hv[r].z *= mk[r]; hv[r].w *= mk[r];
        }
#pragma unroll
        for (int j = 0; j < G; ++j) {
          const float4 w = *reinterpret_cast<const float4*>(s_wh + j * H + k);
          const int a = j < 2 * UN ? j : j + UN;   // r, z accumulate onto the input sums; the n gate keeps gh_n apart
#pragma unroll
          for (int r = 0; r < RW; ++r)
            acc[r][a] = fmaf(hv[r].x, w.x, fmaf(hv[r].y, w.y, fmaf(hv[r].z, w.z, fmaf(hv[r].w, w.w, acc[r][a]))));
        }
      }
    } else {
      for (int k = lane; k < Kx; k += 32) {
        float xv[RW], wv[G];
#pragma unroll
        for (int r = 0; r < RW; ++r) xv[r] = xr[r][k];
#pragma unroll
        for (int j = 0; j < G; ++j) wv[j] = s_wi[j * Kx + k];
#pragma unroll
        for (int r = 0; r < RW; ++r)
#pragma unroll
          for (int j = 0; j < G; ++j) acc[r][j] = fmaf(xv[r], wv[j], acc[r][j]);   // r, z, gi_n
      }
      for (int k = lane; k < H; k += 32) {
        float hv[RW], wv[G];
#pragma unroll
        for (int r = 0; r < RW; ++r) hv[r] = hr[r][k] * mk[r];
#pragma unroll
        for (int j = 0; j < G; ++j) wv[j] = s_wh[j * H + k];
#pragma unroll
        for (int r = 0; r < RW; ++r) {
#pragma unroll
          for (int j = 0; j < 2 * UN; ++j) acc[r][j] = fmaf(hv[r], wv[j], acc[r][j]);                       // r, z
#pragma unroll
          for (int u = 0; u < UN; ++u) acc[r][3 * UN + u] = fmaf(hv[r], wv[2 * UN + u], acc[r][3 * UN + u]);  // gh_n
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RW; ++r)
#pragma unroll
      for (int j = 0; j < 4 * UN; ++j) {
        float v = acc[r][j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[r][j] = v;
      }
    // lane = r * UN + u finishes row r0 + r, unit u0 + u (RW * UN = 16 lanes)
#pragma unroll
    for (int r = 0; r < RW; ++r)
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        if (lane == r * UN + u && r0 + r < rows) {
          const int row = r0 + r, k = u0 + u;
          const float rg = sigmoid_precise(acc[r][u] + b_ih[k] + b_hh[k]);
          const float zg = sigmoid_precise(acc[r][UN + u] + b_ih[H + k] + b_hh[H + k]);
          const float ghn = acc[r][3 * UN + u] + b_hh[2 * H + k];
          const float ng = tanhf(fmaf(rg, ghn, acc[r][2 * UN + u] + b_ih[2 * H + k]));
          const float hp = h0[(size_t)row * ldh0 + k] * mk[r];
          const float hn = fmaf(zg, hp - ng, ng);
          out[(size_t)row * ldo + k] = hn;
          if (out2 != nullptr) out2[(size_t)row * ldo2 + k] = hn;
          if (hm != nullptr) hm[(size_t)row * H + k] = hp;
          if (save != nullptr) {
            float* sv = save + (size_t)row * 4 * H;
            sv[k] = rg; sv[H + k] = zg; sv[2 * H + k] = ng; sv[3 * H + k] = ghn;
          }
        }
      }
  }
}

// ------------------------------------------------------------------------------------
// LSTM cell pointwise (habitat_baselines v0.1.7 rnn_state_encoder with rnn_type="LSTM"; torch.nn.LSTM gate order
// i, f, g, o):  c' = s(f) c + s(i) tanh(g),  h' = s(o) tanh(c').  gi = x W_ih^T + b_ih, gh = h W_hh^T + b_hh
// (both (rows, 4H)); c is the already mask-reset cell state.
// ------------------------------------------------------------------------------------
__global__ void lstm_pointwise_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                                          const float* __restrict__ c, long long ldc,
                                          float* __restrict__ hout, long long ldo,
                                          float* __restrict__ cout, long long ldco,
                                          float* __restrict__ hout2, float* __restrict__ cout2, long long ldo2,
                                          float* __restrict__ save,  // (rows, 5H): i, f, g, o, tanh(c') or null
                                          int rows, int H) {
  const int total = rows * H;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int b = idx / H, k = idx % H;
    const float* gib = gi + (size_t)b * 4 * H;
    const float* ghb = gh + (size_t)b * 4 * H;
    const float ig = sigmoid_precise(gib[k] + ghb[k]);
    const float fg = sigmoid_precise(gib[H + k] + ghb[H + k]);
    const float gg = tanhf(gib[2 * H + k] + ghb[2 * H + k]);
    const float og = sigmoid_precise(gib[3 * H + k] + ghb[3 * H + k]);
    const float cn = fmaf(fg, c[b * ldc + k], ig * gg);
    const float tc = tanhf(cn);
    const float hn = og * tc;
    hout[b * ldo + k] = hn;
    cout[b * ldco + k] = cn;
    if (hout2 != nullptr) { hout2[b * ldo2 + k] = hn; cout2[b * ldo2 + k] = cn; }
    if (save != nullptr) {
      float* s = save + (size_t)b * 5 * H;
      s[k] = ig; s[H + k] = fg; s[2 * H + k] = gg; s[3 * H + k] = og; s[4 * H + k] = tc;
    }
  }
}

// out[b, :] = h[b, :] * mask[b]   (hidden-state reset where an episode starts)
__global__ void mask_rows_kernel(const float* __restrict__ h, long long ldh, const float* __restrict__ mask,
                                 float* __restrict__ out, long long ldo, int rows, int H) {
  const int total = rows * H;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int b = idx / H, k = idx % H;
    out[b * ldo + k] = h[b * ldh + k] * mask[b];
  }
}

// ------------------------------------------------------------------------------------
// Single-query attention pooling (cma_policy.py:207-217 `_attn`, re-associated):
//   a = softmax(scale * (E q - 1e8 * [p >= len])),  out = sum_p a_p E_p
// With keys k_p = W_k e_p + b_k the reference's logits q.k_p equal (W_k^T q).e_p + q.b_k; the
// constant q.b_k cancels in the softmax, so the caller passes q' = W_k^T q and the raw
// embedding E (read once per pass instead of materialising K and V).  One CTA per row.
// ------------------------------------------------------------------------------------
__global__ void attn_pool_fwd_kernel(const float* __restrict__ q, long long ldq,
                                     const float* __restrict__ E, long long lde_r, long long lde_p,
                                     const int* __restrict__ lengths, float scale,
                                     float* __restrict__ probs,  // (R, P)
                                     float* __restrict__ out, long long ldo, int P, int D) {
  extern __shared__ float sm[];
  float* logit = sm;           // P
  float* red = sm + P;         // 32
  const int r = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float* qr = q + r * ldq;
  const float* Er = E + r * lde_r;
  const int len = lengths != nullptr ? lengths[r] : P;
  for (int p = warp; p < P; p += nw) {
    const float* e = Er + p * lde_p;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s = fmaf(qr[d], e[d], s);
    s = warp_sum(s);
    if (lane == 0) logit[p] = (p < len ? s : s - 1e8f) * scale;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int p = threadIdx.x; p < P; p += blockDim.x) mx = fmaxf(mx, logit[p]);
  mx = block_max(mx, red);
  float sum = 0.f;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const float ex = expf(logit[p] - mx);
    logit[p] = ex;
    sum += ex;
  }
  sum = block_sum(sum, red);
  __syncthreads();
  const float inv = 1.f / sum;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const float a = logit[p] * inv;
    logit[p] = a;
    if (probs != nullptr) probs[(size_t)r * P + p] = a;
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int p = 0; p < P; ++p) acc = fmaf(logit[p], Er[p * lde_p + d], acc);
    out[r * ldo + d] = acc;
  }
}

// out[r, d] = mean_p E[r, p, d]   (AdaptiveAvgPool1d(1), cma_policy.py:105-108)
__global__ void mean_positions_kernel(const float* __restrict__ E, long long lde_r, long long lde_p,
                                      float* __restrict__ out, long long ldo, int R, int P, int D) {
  const long long total = (long long)R * D;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / D), d = (int)(idx % D);
    float acc = 0.f;
    for (int p = 0; p < P; ++p) acc += E[r * lde_r + p * lde_p + d];
    out[r * ldo + d] = acc / (float)P;
  }
}

// E[r, s, c_off + c'] = table_flat[c' * P + s]: the raw `.view(1, 64, 4, 4)` of the (16, 64)
// spatial-embedding table (resnet_encoders.py:97-113, 201-217).
__global__ void spatial_embed_fill_kernel(const float* __restrict__ table, float* __restrict__ E,
                                          long long lde_r, long long lde_p, int c_off, int R, int P,
                                          int Cs) {
  const long long total = (long long)R * P * Cs;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % Cs);
    const int s = (int)((idx / Cs) % P);
    const int r = (int)(idx / ((long long)Cs * P));
    E[r * lde_r + s * lde_p + c_off + c] = __ldg(table + c * P + s);
  }
}

// (R, P, C) channels-last -> (R, C*P) NCHW-flatten order (torch Flatten of (R,C,4,4))
__global__ void positions_to_nchw_kernel(const float* __restrict__ E, long long lde_r, long long lde_p,
                                         float* __restrict__ out, long long ldo, int R, int P, int C) {
  const long long total = (long long)R * P * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(idx % P);
    const int c = (int)((idx / P) % C);
    const int r = (int)(idx / ((long long)P * C));
    out[r * ldo + (long long)c * P + s] = E[r * lde_r + s * lde_p + c];
  }
}

// log-softmax over A <= 32 logits per row, in place (torch.distributions.Categorical(logits=.))
__global__ void log_softmax_rows_kernel(float* __restrict__ x, int rows, int A) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float* p = x + (size_t)r * A;
  float mx = -INFINITY;
  for (int a = 0; a < A; ++a) mx = fmaxf(mx, p[a]);
  float s = 0.f;
  for (int a = 0; a < A; ++a) s += expf(p[a] - mx);
  const float lse = mx + logf(s);
  for (int a = 0; a < A; ++a) p[a] -= lse;
}

// argmax over probs (CustomFixedCategorical.mode, models/utils.py:288-289): first maximum
__global__ void argmax_rows_kernel(const float* __restrict__ x, long long* __restrict__ out, int rows,
                                   int A) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* p = x + (size_t)r * A;
  int best = 0;
  float bv = p[0];
  for (int a = 1; a < A; ++a)
    if (p[a] > bv) { bv = p[a]; best = a; }
  out[r] = best;
}

inline unsigned grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148ll * 16;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace
}  // namespace vlnce

using namespace vlnce;

extern "C" int vlnce_embedding_fwd(const long long* ids, const float* table, float* out,
                                   long long rows, int dim, int ldo, int vocab, cudaStream_t stream) {
  VLNCE_CHECK_ARG(ids && table && out && rows > 0 && dim > 0 && ldo >= dim, "embedding_fwd: bad arguments");
  embedding_fwd_kernel<<<grid_for(rows * dim, 256), 256, 0, stream>>>(ids, table, out, rows, dim, ldo, vocab);
  VLNCE_LAUNCH_CHECK("embedding_fwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_prev_action_embed(const long long* prev_actions, const float* masks,
                                       const float* table, float* out, int ldo, float* out2, int ldo2,
                                       int rows, int dim, int table_rows, cudaStream_t stream) {
  VLNCE_CHECK_ARG(prev_actions && masks && table && out && rows > 0, "prev_action_embed: bad arguments");
  prev_action_embed_kernel<<<grid_for((long long)rows * dim, 256), 256, 0, stream>>>(
      prev_actions, masks, table, out, out2, rows, dim, ldo, ldo2, table_rows);
  VLNCE_LAUNCH_CHECK("prev_action_embed_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_token_lengths(const long long* ids, int* lengths, int rows, int L,
                                   cudaStream_t stream) {
  VLNCE_CHECK_ARG(ids && lengths && rows > 0 && L > 0, "token_lengths: bad arguments");
  token_lengths_kernel<<<cdiv(rows, 8), 256, 0, stream>>>(ids, lengths, rows, L);
  VLNCE_LAUNCH_CHECK("token_lengths_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_feature_lengths(const float* x, int* lengths, int rows, int L, int D,
                                     cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && lengths && rows > 0 && L > 0 && D > 0, "feature_lengths: bad arguments");
  feature_lengths_kernel<<<rows, 256, 0, stream>>>(x, lengths, L, D);
  VLNCE_LAUNCH_CHECK("feature_lengths_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_lstm_seq_fwd(const float* gx, const float* whh_f, const float* whh_b,
                                  const float* bhh_f, const float* bhh_b, const int* lengths,
                                  float* out, float* hn, float* save_gates, float* save_c, int R, int L,
                                  int H, int dirs, cudaStream_t stream) {
  VLNCE_CHECK_ARG(gx && whh_f && bhh_f && lengths && out && R > 0 && L > 0 && (dirs == 1 || dirs == 2),
                  "lstm_seq_fwd: bad arguments");
  VLNCE_CHECK_ARG(dirs == 1 || (whh_b && bhh_b), "lstm_seq_fwd: reverse-direction weights missing");
  VLNCE_CHECK_ARG(H == 128, "lstm_seq_fwd: hidden size %d not built (128 only)", H);
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)R * L * dirs * H, stream);
  if (e != cudaSuccess) return cuda_status(e, "lstm_seq_fwd memset");
  constexpr int H_ = 128;
  // NR rows per CTA: with a batch of rows the recurrence is latency-bound, so 4 rows share one SM (and one read of
  // the weights per step); single rows / tiny batches keep one row per CTA
  const int nr = lstm_rows_per_cta(R);
  dim3 grid((R + nr - 1) / nr, dirs);
  const size_t smem = sizeof(float) * ((size_t)(H_ / 2) * 4 * H_ + (size_t)nr * 5 * H_);
  static bool configured = false;
  if (!configured) {
    const int big = (int)(sizeof(float) * ((size_t)(H_ / 2) * 4 * H_ + 4 * 5 * H_));
    e = cudaFuncSetAttribute(lstm_seq_fwd_kernel<H_, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_seq_fwd_kernel<H_, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_seq_fwd_kernel<H_, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(lstm_seq_fwd)");
    configured = true;
  }
  if (nr == 4)
    lstm_seq_fwd_kernel<H_, 4><<<grid, 4 * H_, smem, stream>>>(gx, whh_f, whh_b, bhh_f, bhh_b, lengths, out, hn, save_gates,
                                                             save_c, R, L, dirs);
  else if (nr == 2)
    lstm_seq_fwd_kernel<H_, 2><<<grid, 4 * H_, smem, stream>>>(gx, whh_f, whh_b, bhh_f, bhh_b, lengths, out, hn, save_gates,
                                                             save_c, R, L, dirs);
  else
    lstm_seq_fwd_kernel<H_, 1><<<grid, 4 * H_, smem, stream>>>(gx, whh_f, whh_b, bhh_f, bhh_b, lengths, out, hn, save_gates,
                                                             save_c, R, L, dirs);
  VLNCE_LAUNCH_CHECK("lstm_seq_fwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_gru_seq_fwd(const float* gx, const float* whh_f, const float* whh_b, const float* bhh_f,
                                 const float* bhh_b, const int* lengths, float* out, float* hn, float* save, int R,
                                 int L, int H, int dirs, cudaStream_t stream) {
  VLNCE_CHECK_ARG(gx && whh_f && bhh_f && lengths && out && R > 0 && L > 0 && (dirs == 1 || dirs == 2),
                  "gru_seq_fwd: bad arguments");
  VLNCE_CHECK_ARG(dirs == 1 || (whh_b && bhh_b), "gru_seq_fwd: reverse-direction weights missing");
  VLNCE_CHECK_ARG(H == 128, "gru_seq_fwd: hidden size %d not built (128 only)", H);
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)R * L * dirs * H, stream);
  if (e != cudaSuccess) return cuda_status(e, "gru_seq_fwd memset");
  dim3 grid(R, dirs);
  constexpr int H_ = 128;
  const size_t smem = sizeof(float) * ((size_t)(H_ / 2) * 3 * H_ + H_ + 3 * H_);
  static bool configured = false;
  if (!configured) {
    e = cudaFuncSetAttribute(gru_seq_fwd_kernel<H_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(gru_seq_fwd)");
    configured = true;
  }
  gru_seq_fwd_kernel<H_><<<grid, 3 * H_, smem, stream>>>(gx, whh_f, whh_b, bhh_f, bhh_b, lengths, out, hn, save, R, L,
                                                       dirs);
  VLNCE_LAUNCH_CHECK("gru_seq_fwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_gru_pointwise_fwd(const float* gi, const float* gh, const float* h, long long ldh,
                                       float* hout, long long ldo, float* hout2, long long ldo2,
                                       float* save, int rows, int H, cudaStream_t stream) {
  VLNCE_CHECK_ARG(gi && gh && h && hout && rows > 0 && H > 0, "gru_pointwise_fwd: bad arguments");
  gru_pointwise_fwd_kernel<<<grid_for((long long)rows * H, 256), 256, 0, stream>>>(
      gi, gh, h, ldh, hout, ldo, hout2, ldo2, save, rows, H);
  VLNCE_LAUNCH_CHECK("gru_pointwise_fwd_kernel");
  return VLNCE_OK;
}

namespace vlnce {
namespace {
__device__ __forceinline__ float round_tf32_dev(float v) {
  uint32_t t;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
  return __uint_as_float(t);
}
__global__ void transpose_f32_kernel(const float* __restrict__ src, long long lds, float* __restrict__ dst, long long ldd,
                                     int rows, int cols, int round_tf32) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(size_t)r * lds + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) {
      const float v = tile[threadIdx.x][i];
      dst[(size_t)c * ldd + r] = round_tf32 ? round_tf32_dev(v) : v;
    }
  }
}
__global__ void round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = round_tf32_dev(src[i]);
}
}  // namespace
}  // namespace vlnce

extern "C" int vlnce_transpose_f32(const float* src, long long lds, float* dst, long long ldd, int rows, int cols,
                                   int round_tf32, cudaStream_t stream) {
  VLNCE_CHECK_ARG(src && dst && rows > 0 && cols > 0 && lds >= cols && ldd >= rows, "transpose_f32: bad arguments");
  VLNCE_CHECK_ARG((rows + 31) / 32 <= 65535, "transpose_f32: too many rows");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32);
  transpose_f32_kernel<<<grid, dim3(32, 8), 0, stream>>>(src, lds, dst, ldd, rows, cols, round_tf32);
  VLNCE_LAUNCH_CHECK("transpose_f32_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_round_tf32(const float* src, float* dst, long long n, cudaStream_t stream) {
  VLNCE_CHECK_ARG(src && dst && n > 0, "round_tf32: bad arguments");
  round_tf32_kernel<<<grid_for(n, 256), 256, 0, stream>>>(src, dst, n);
  VLNCE_LAUNCH_CHECK("round_tf32_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_gru_step_fwd(const float* x, long long ldx, int in_size, const float* h0, long long ldh0,
                                  const float* mask, const float* w_ih, const float* w_hh, const float* b_ih,
                                  const float* b_hh, float* hm, float* out, long long ldo, float* out2, long long ldo2,
                                  float* save, int rows, int H, cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && h0 && mask && w_ih && w_hh && b_ih && b_hh && out, "gru_step_fwd: null tensor");
  VLNCE_CHECK_ARG(rows > 0 && H > 0 && in_size > 0 && H % kGruUN == 0, "gru_step_fwd: rows=%d H=%d in=%d", rows, H, in_size);
  const size_t smem = sizeof(float) * 3 * kGruUN * ((size_t)in_size + H);
  VLNCE_CHECK_ARG(smem <= 200 * 1024, "gru_step_fwd: weight slice of %zu bytes does not fit shared memory", smem);
  const bool vec = in_size % 4 == 0 && ldx % 4 == 0 && ldh0 % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(h0) & 15) == 0;
  static bool attr_set = false;   // (idempotent; a racing second call sets the same value)
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gru_step_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gru_step_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(gru_step_fused)");
    attr_set = true;
  }
  const int threads = rows > 32 ? 512 : 256;   // one warp per 4 batch rows, up to 16 warps
  if (vec)
    gru_step_fused_kernel<true><<<H / kGruUN, threads, smem, stream>>>(x, ldx, in_size, h0, ldh0, mask, w_ih, w_hh, b_ih,
                                                                        b_hh, hm, out, ldo, out2, ldo2, save, rows, H);
  else
    gru_step_fused_kernel<false><<<H / kGruUN, threads, smem, stream>>>(x, ldx, in_size, h0, ldh0, mask, w_ih, w_hh, b_ih,
                                                                         b_hh, hm, out, ldo, out2, ldo2, save, rows, H);
  VLNCE_LAUNCH_CHECK("gru_step_fused_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_lstm_pointwise_fwd(const float* gi, const float* gh, const float* c, long long ldc,
                                        float* hout, long long ldo, float* cout, long long ldco,
                                        float* hout2, float* cout2, long long ldo2, float* save, int rows, int H,
                                        cudaStream_t stream) {
  VLNCE_CHECK_ARG(gi && gh && c && hout && cout && rows > 0 && H > 0, "lstm_pointwise_fwd: bad arguments");
  VLNCE_CHECK_ARG((hout2 == nullptr) == (cout2 == nullptr), "lstm_pointwise_fwd: hout2 and cout2 go together");
  lstm_pointwise_fwd_kernel<<<grid_for((long long)rows * H, 256), 256, 0, stream>>>(
      gi, gh, c, ldc, hout, ldo, cout, ldco, hout2, cout2, ldo2, save, rows, H);
  VLNCE_LAUNCH_CHECK("lstm_pointwise_fwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_mask_rows(const float* h, long long ldh, const float* mask, float* out,
                               long long ldo, int rows, int H, cudaStream_t stream) {
  VLNCE_CHECK_ARG(h && mask && out && rows > 0 && H > 0, "mask_rows: bad arguments");
  mask_rows_kernel<<<grid_for((long long)rows * H, 256), 256, 0, stream>>>(h, ldh, mask, out, ldo, rows, H);
  VLNCE_LAUNCH_CHECK("mask_rows_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_attn_pool_fwd(const float* q, long long ldq, const float* E, long long lde_r,
                                   long long lde_p, const int* lengths, float scale, float* probs,
                                   float* out, long long ldo, int R, int P, int D,
                                   cudaStream_t stream) {
  VLNCE_CHECK_ARG(q && E && out && R > 0 && P > 0 && D > 0, "attn_pool_fwd: bad arguments");
  VLNCE_CHECK_ARG(P <= 8192, "attn_pool_fwd: P=%d too long", P);
  const size_t smem = sizeof(float) * (P + 32);
  attn_pool_fwd_kernel<<<R, 256, smem, stream>>>(q, ldq, E, lde_r, lde_p, lengths, scale, probs, out, ldo, P, D);
  VLNCE_LAUNCH_CHECK("attn_pool_fwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_mean_positions(const float* E, long long lde_r, long long lde_p, float* out,
                                    long long ldo, int R, int P, int D, cudaStream_t stream) {
  VLNCE_CHECK_ARG(E && out && R > 0 && P > 0 && D > 0, "mean_positions: bad arguments");
  mean_positions_kernel<<<grid_for((long long)R * D, 256), 256, 0, stream>>>(E, lde_r, lde_p, out, ldo, R, P, D);
  VLNCE_LAUNCH_CHECK("mean_positions_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_spatial_embed_fill(const float* table, float* E, long long lde_r, long long lde_p,
                                        int c_off, int R, int P, int Cs, cudaStream_t stream) {
  VLNCE_CHECK_ARG(table && E && R > 0 && P > 0 && Cs > 0, "spatial_embed_fill: bad arguments");
  spatial_embed_fill_kernel<<<grid_for((long long)R * P * Cs, 256), 256, 0, stream>>>(table, E, lde_r, lde_p,
                                                                                     c_off, R, P, Cs);
  VLNCE_LAUNCH_CHECK("spatial_embed_fill_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_positions_to_nchw(const float* E, long long lde_r, long long lde_p, float* out,
                                       long long ldo, int R, int P, int C, cudaStream_t stream) {
  VLNCE_CHECK_ARG(E && out && R > 0 && P > 0 && C > 0, "positions_to_nchw: bad arguments");
  positions_to_nchw_kernel<<<grid_for((long long)R * P * C, 256), 256, 0, stream>>>(E, lde_r, lde_p, out, ldo,
                                                                                   R, P, C);
  VLNCE_LAUNCH_CHECK("positions_to_nchw_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_log_softmax_rows(float* x, int rows, int A, cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && rows > 0 && A > 0, "log_softmax_rows: bad arguments");
  log_softmax_rows_kernel<<<cdiv(rows, 128), 128, 0, stream>>>(x, rows, A);
  VLNCE_LAUNCH_CHECK("log_softmax_rows_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_argmax_rows(const float* x, long long* out, int rows, int A, cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && out && rows > 0 && A > 0, "argmax_rows: bad arguments");
  argmax_rows_kernel<<<cdiv(rows, 128), 128, 0, stream>>>(x, out, rows, A);
  VLNCE_LAUNCH_CHECK("argmax_rows_kernel");
  return VLNCE_OK;
}
