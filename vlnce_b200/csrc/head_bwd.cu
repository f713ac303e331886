// Backward kernels of the trainable head, the fused DAgger / teacher-forcing loss
// (common/base_il_trainer.py:159-172 + common/aux_losses.py:24-32 + cma_policy.py:296-307)
// and the fused Adam step (torch.optim.Adam semantics, base_il_trainer.py:174-176).
// All fp32; reductions are deterministic (no floating-point atomics) except the two
// embedding-table scatters, which sum a handful of rows per entry.
#include "common.cuh"

namespace vlnce {
namespace {

inline unsigned grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148ll * 16;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ------------------------------------------------------------------------------------
// elementwise pieces
// ------------------------------------------------------------------------------------
__global__ void relu_bwd_kernel(float* __restrict__ d, long long ldd, const float* __restrict__ y,
                                long long ldy, int rows, int cols) {
  const long long total = (long long)rows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / cols), c = (int)(idx % cols);
    if (y[r * ldy + c] <= 0.f) d[r * ldd + c] = 0.f;
  }
}

__global__ void tanh_bwd_kernel(float* __restrict__ d, const float* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    d[i] *= 1.f - y[i] * y[i];
}

// d_raw = d_lp - softmax * sum_a d_lp   (backward of lp = raw - logsumexp(raw))
__global__ void log_softmax_bwd_kernel(const float* __restrict__ lp, const float* __restrict__ dlp,
                                       float* __restrict__ draw, int rows, int A) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float s = 0.f;
  for (int a = 0; a < A; ++a) s += dlp[(size_t)r * A + a];
  for (int a = 0; a < A; ++a) draw[(size_t)r * A + a] = dlp[(size_t)r * A + a] - expf(lp[(size_t)r * A + a]) * s;
}

// dst[b, :] += src[b, :] * mask[b]
__global__ void mask_rows_add_kernel(const float* __restrict__ src, long long lds,
                                     const float* __restrict__ mask, float* __restrict__ dst,
                                     long long ldd, int rows, int H) {
  const int total = rows * H;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int b = idx / H, k = idx % H;
    dst[b * ldd + k] += src[b * lds + k] * mask[b];
  }
}

// ------------------------------------------------------------------------------------
// GRU cell backward (pointwise part).  save = [r, z, n, gh_n] from the forward.
// ------------------------------------------------------------------------------------
__global__ void gru_pointwise_bwd_kernel(const float* __restrict__ dh, long long lddh,
                                         const float* __restrict__ save, const float* __restrict__ hm,
                                         float* __restrict__ dgi, float* __restrict__ dgh,
                                         float* __restrict__ dhm, int rows, int H) {
  const int total = rows * H;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int b = idx / H, k = idx % H;
    const float* s = save + (size_t)b * 4 * H;
    const float r = s[k], z = s[H + k], n = s[2 * H + k], ghn = s[3 * H + k];
    const float g = dh[b * lddh + k];
    const float hp = hm[(size_t)b * H + k];
    const float dn = g * (1.f - z);
    const float dz = g * (hp - n);
    const float dpn = dn * (1.f - n * n);
    const float dpz = dz * z * (1.f - z);
    const float dpr = dpn * ghn * r * (1.f - r);
    float* gi = dgi + (size_t)b * 3 * H;
    float* gh = dgh + (size_t)b * 3 * H;
    gi[k] = dpr; gi[H + k] = dpz; gi[2 * H + k] = dpn;
    gh[k] = dpr; gh[H + k] = dpz; gh[2 * H + k] = dpn * r;
    dhm[(size_t)b * H + k] = g * z;
  }
}

// ------------------------------------------------------------------------------------
// LSTM cell backward (pointwise part).  save = [i, f, g, o, tanh(c')] from the forward, cm = the mask-reset cell
// state the step consumed.  dc (rows, H) carries the cell-state gradient between steps: read (unless null = 0),
// then overwritten with d c_prev * mask[row] (null mask = 1) for the step before.  The gate pre-activations are
// gi + gh, so one dgates (rows, 4H) buffer serves both weight matrices.
// ------------------------------------------------------------------------------------
__global__ void lstm_pointwise_bwd_kernel(const float* __restrict__ dh, long long lddh, const float* dc_in,
                                          const float* __restrict__ save, const float* __restrict__ cm,
                                          const float* __restrict__ mask, float* __restrict__ dgates,
                                          float* dc_out, int rows, int H) {
  const int total = rows * H;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int b = idx / H, k = idx % H;
    const float* s = save + (size_t)b * 5 * H;
    const float i = s[k], f = s[H + k], gg = s[2 * H + k], o = s[3 * H + k], tc = s[4 * H + k];
    const float g = dh[b * lddh + k];
    const float dct = (dc_in != nullptr ? dc_in[(size_t)b * H + k] : 0.f) + g * o * (1.f - tc * tc);
    float* d = dgates + (size_t)b * 4 * H;
    d[k] = dct * gg * i * (1.f - i);
    d[H + k] = dct * cm[(size_t)b * H + k] * f * (1.f - f);
    d[2 * H + k] = dct * i * (1.f - gg * gg);
    d[3 * H + k] = g * tc * o * (1.f - o);
    dc_out[(size_t)b * H + k] = dct * f * (mask != nullptr ? mask[b] : 1.f);
  }
}

// ------------------------------------------------------------------------------------
// Attention-pool backward.  out = sum_p a_p E_p, a = softmax(scale * (E q - mask)).
//   da_p = E_p . dout;  dl_p = a_p (da_p - sum_j a_j da_j) * scale
//   dq = sum_p dl_p E_p;  dE_p (+)= a_p dout + dl_p q
// One CTA per row.
// ------------------------------------------------------------------------------------
__global__ void attn_pool_bwd_kernel(const float* __restrict__ q, long long ldq,
                                     const float* __restrict__ E, long long lde_r, long long lde_p,
                                     const float* __restrict__ probs, float scale,
                                     const float* __restrict__ dout, long long lddo,
                                     float* __restrict__ dq, long long lddq, float* __restrict__ dE,
                                     long long ldde_r, long long ldde_p, int accumulate_dE, int P, int D) {
  extern __shared__ float sm[];
  float* dl = sm;       // P
  float* red = sm + P;  // 32
  const int r = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float* qr = q + r * ldq;
  const float* Er = E + r * lde_r;
  const float* dor = dout + r * lddo;
  const float* pr = probs + (size_t)r * P;
  for (int p = warp; p < P; p += nw) {
    const float* e = Er + p * lde_p;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s = fmaf(e[d], dor[d], s);
    s = warp_sum(s);
    if (lane == 0) dl[p] = s;  // da_p
  }
  __syncthreads();
  float part = 0.f;
  for (int p = threadIdx.x; p < P; p += blockDim.x) part = fmaf(pr[p], dl[p], part);
  const float dot = block_sum(part, red);
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += blockDim.x) dl[p] = pr[p] * (dl[p] - dot) * scale;
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    const float qd = qr[d], dod = dor[d];
    for (int p = 0; p < P; ++p) {
      acc = fmaf(dl[p], Er[p * lde_p + d], acc);
      if (dE != nullptr) {
        float* de = dE + r * ldde_r + p * ldde_p + d;
        const float v = fmaf(pr[p], dod, dl[p] * qd);
        *de = accumulate_dE ? *de + v : v;
      }
    }
    dq[r * lddq + d] = acc;
  }
}

// ------------------------------------------------------------------------------------
// LSTM backward through time, one CTA per (group of NR rows, direction), 4H threads.
// Thread t owns hidden unit k = t % H and the gate-row quarter part = t / H; it keeps the
// transposed weights W_hh[j, k], j in its quarter, half in registers / half in shared memory, and
// applies them to the NR rows of its group in every step (as the forward kernel: the weights are
// read once per step for NR rows, the recurrence occupies R/NR SMs).  The saved activations of the
// NEXT step are fetched while the current step computes.
// Emits d_gx (pre-activation gate grads, zero beyond len) and hprev (h fed to each step).
// ------------------------------------------------------------------------------------
template <int H, int NR>
__global__ void __launch_bounds__(4 * H, 1)
lstm_seq_bwd_kernel(const float* __restrict__ dout,   // (R, L, dirs*H)
                    const float* __restrict__ gates,  // (dirs, R, L, 4H) activated i,f,g,o
                    const float* __restrict__ cs,     // (dirs, R, L, H)
                    const float* __restrict__ out,    // (R, L, dirs*H) forward outputs
                    const float* __restrict__ whh_f, const float* __restrict__ whh_b,
                    const int* __restrict__ lengths, float* __restrict__ dgx,  // (dirs, R, L, 4H)
                    float* __restrict__ hprev,                                 // (dirs, R, L, H)
                    int R, int L, int dirs) {
  constexpr int JR = H / 2;        // weights in registers
  constexpr int JS = H - JR;       // weights in shared memory
  extern __shared__ __align__(16) float smem_b[];
  float* w_s = smem_b;                         // [JS][4H]: w_s[jj][t] = W[part*H + JR + jj, k]
  float* dp_s = w_s + (size_t)JS * 4 * H;      // NR x 4H pre-activation grads of this step
  float* red_s = dp_s + NR * 4 * H;            // NR x 4H partial sums
  const int r0 = blockIdx.x * NR, dir = blockIdx.y;
  const int t = threadIdx.x, k = t % H, part = t / H;
  const float* whh = dir == 0 ? whh_f : whh_b;
  float w[JR];
#pragma unroll
  for (int jj = 0; jj < JR; ++jj) w[jj] = __ldg(whh + (size_t)(part * H + jj) * H + k);
  for (int jj = 0; jj < JS; ++jj) w_s[(size_t)jj * 4 * H + t] = __ldg(whh + (size_t)(part * H + JR + jj) * H + k);
  int len[NR];
  int maxlen = 0;
#pragma unroll
  for (int rr = 0; rr < NR; ++rr) {
    len[rr] = (r0 + rr < R) ? min(lengths[r0 + rr], L) : 0;
    maxlen = max(maxlen, len[rr]);
  }
  const int DH = dirs * H;
  float dh_carry[NR], dc_carry[NR];  // meaningful for part == 0 threads (unit k)
#pragma unroll
  for (int rr = 0; rr < NR; ++rr) { dh_carry[rr] = 0.f; dc_carry[rr] = 0.f; }
  // saved values of (row rr, step) needed by unit k: the four gates, c, c_prev, h_prev, dout
  struct Saved { float ig, fg, gg, og, c, cp, hp, d; };
  auto fetch = [&](int rr, int step) -> Saved {
    Saved v = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (part != 0 || step < 0 || step >= len[rr]) return v;
    const size_t row_base = ((size_t)dir * R + r0 + rr) * L;
    const int tt = dir == 0 ? step : len[rr] - 1 - step;      // time index of this step
    const int tp = dir == 0 ? tt - 1 : tt + 1;                // time index of the previous step
    const float* g = gates + (row_base + tt) * 4 * H;
    v.ig = __ldg(g + k); v.fg = __ldg(g + H + k); v.gg = __ldg(g + 2 * H + k); v.og = __ldg(g + 3 * H + k);
    v.c = __ldg(cs + (row_base + tt) * H + k);
    if (step > 0) {
      v.cp = __ldg(cs + (row_base + tp) * H + k);
      v.hp = __ldg(out + ((size_t)(r0 + rr) * L + tp) * DH + dir * H + k);
    }
    v.d = __ldg(dout + ((size_t)(r0 + rr) * L + tt) * DH + dir * H + k);
    return v;
  };
  Saved nxt[NR];
#pragma unroll
  for (int rr = 0; rr < NR; ++rr) nxt[rr] = fetch(rr, maxlen - 1);   // (zeros for rows shorter than the longest)
  __syncthreads();
  for (int step = maxlen - 1; step >= 0; --step) {
    if (part == 0) {
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) {
        const Saved v = nxt[rr];
        nxt[rr] = fetch(rr, step - 1);   // in flight during this step
        float* dp = dp_s + rr * 4 * H;
        if (step < len[rr]) {
          const size_t row_base = ((size_t)dir * R + r0 + rr) * L;
          const int tt = dir == 0 ? step : len[rr] - 1 - step;
          hprev[(row_base + tt) * H + k] = v.hp;
          const float th = tanhf(v.c);
          const float dh = v.d + dh_carry[rr];
          const float dc = fmaf(dh * v.og, 1.f - th * th, dc_carry[rr]);
          const float dpi = dc * v.gg * v.ig * (1.f - v.ig);
          const float dpf = dc * v.cp * v.fg * (1.f - v.fg);
          const float dpg = dc * v.ig * (1.f - v.gg * v.gg);
          const float dpo = dh * th * v.og * (1.f - v.og);
          dc_carry[rr] = dc * v.fg;
          dp[k] = dpi; dp[H + k] = dpf; dp[2 * H + k] = dpg; dp[3 * H + k] = dpo;
          float* o = dgx + (row_base + tt) * 4 * H;
          o[k] = dpi; o[H + k] = dpf; o[2 * H + k] = dpg; o[3 * H + k] = dpo;
        } else {
          dp[k] = 0.f; dp[H + k] = 0.f; dp[2 * H + k] = 0.f; dp[3 * H + k] = 0.f;
        }
      }
    }
    __syncthreads();
    // dh_prev[k] = sum_j W[j, k] dp[j]; this thread covers j in [part*H, (part+1)*H)
    float a0[NR], a1[NR];
#pragma unroll
    for (int rr = 0; rr < NR; ++rr) { a0[rr] = 0.f; a1[rr] = 0.f; }
#pragma unroll
    for (int jj = 0; jj < JR; jj += 2) {
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) {
        const float* dp = dp_s + rr * 4 * H + part * H;
        a0[rr] = fmaf(w[jj], dp[jj], a0[rr]);
        a1[rr] = fmaf(w[jj + 1], dp[jj + 1], a1[rr]);
      }
    }
#pragma unroll 8
    for (int jj = 0; jj < JS; jj += 2) {
      const float w0 = w_s[(size_t)jj * 4 * H + t], w1 = w_s[(size_t)(jj + 1) * 4 * H + t];
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) {
        const float* dp = dp_s + rr * 4 * H + part * H;
        a0[rr] = fmaf(w0, dp[JR + jj], a0[rr]);
        a1[rr] = fmaf(w1, dp[JR + jj + 1], a1[rr]);
      }
    }
#pragma unroll
    for (int rr = 0; rr < NR; ++rr) red_s[rr * 4 * H + t] = a0[rr] + a1[rr];
    __syncthreads();
    if (part == 0) {
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) {
        const float* rs = red_s + rr * 4 * H;
        dh_carry[rr] = (rs[k] + rs[H + k]) + (rs[2 * H + k] + rs[3 * H + k]);
      }
    }
    // next iteration's first phase only touches dp_s after the barrier below
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------
// GRU backward through time (packed instruction encoder with rnn_type = "GRU"), one CTA per (row, direction),
// 3H threads.  Thread t owns hidden unit k = t % H and the gate-row third part = t / H with the transposed
// weights W_hh[j, k], j in its third (half registers / half shared memory).  Emits d_gx (gradient w.r.t.
// x W_ih^T + b_ih), d_gh (w.r.t. h W_hh^T + b_hh; differs in the candidate rows by the factor r) and hprev
// (h fed to each step) — all zero beyond len; the weight gradients are GEMMs over them.
// ------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(3 * H, 1)
gru_seq_bwd_kernel(const float* __restrict__ dout,   // (R, L, dirs*H)
                   const float* __restrict__ save,   // (dirs, R, L, 4H): r, z, n, gh_n
                   const float* __restrict__ out,    // (R, L, dirs*H) forward outputs
                   const float* __restrict__ whh_f, const float* __restrict__ whh_b,
                   const int* __restrict__ lengths, float* __restrict__ dgx, float* __restrict__ dgh,  // (dirs,R,L,3H)
                   float* __restrict__ hprev,                                                         // (dirs,R,L,H)
                   int R, int L, int dirs) {
  constexpr int G = 3 * H;
  constexpr int JR = H / 2;
  constexpr int JS = H - JR;
  extern __shared__ __align__(16) float smem_g[];
  float* w_s = smem_g;                      // [JS][G]: w_s[jj][t] = W[part*H + JR + jj, k]
  float* dp_s = w_s + (size_t)JS * G;       // G: d_gh of this step
  float* red_s = dp_s + G;                  // G partial sums
  const int r = blockIdx.x, dir = blockIdx.y;
  const int t = threadIdx.x, k = t % H, part = t / H;
  const float* whh = dir == 0 ? whh_f : whh_b;
  float w[JR];
#pragma unroll
  for (int jj = 0; jj < JR; ++jj) w[jj] = __ldg(whh + (size_t)(part * H + jj) * H + k);
  for (int jj = 0; jj < JS; ++jj) w_s[(size_t)jj * G + t] = __ldg(whh + (size_t)(part * H + JR + jj) * H + k);
  const int len = min(lengths[r], L);
  const size_t row_base = ((size_t)dir * R + r) * L;
  const int DH = dirs * H;
  float dh_carry = 0.f, dh_direct = 0.f;  // meaningful for part == 0 threads (unit k)
  __syncthreads();
  for (int step = len - 1; step >= 0; --step) {
    const int tt = dir == 0 ? step : len - 1 - step;
    const int tp = dir == 0 ? tt - 1 : tt + 1;
    if (part == 0) {
      const float* sv = save + (row_base + tt) * 4 * H;
      const float rg = sv[k], zg = sv[H + k], ng = sv[2 * H + k], ghn = sv[3 * H + k];
      const float hp = step > 0 ? out[((size_t)r * L + tp) * DH + dir * H + k] : 0.f;
      hprev[(row_base + tt) * H + k] = hp;
      const float dh = dout[((size_t)r * L + tt) * DH + dir * H + k] + dh_carry;
      const float dpn = dh * (1.f - zg) * (1.f - ng * ng);
      const float dpz = dh * (hp - ng) * zg * (1.f - zg);
      const float dpr = dpn * ghn * rg * (1.f - rg);
      dh_direct = dh * zg;
      float* ox = dgx + (row_base + tt) * G;
      ox[k] = dpr; ox[H + k] = dpz; ox[2 * H + k] = dpn;
      float* oh = dgh + (row_base + tt) * G;
      oh[k] = dpr; oh[H + k] = dpz; oh[2 * H + k] = dpn * rg;
      dp_s[k] = dpr; dp_s[H + k] = dpz; dp_s[2 * H + k] = dpn * rg;
    }
    __syncthreads();
    float a0 = 0.f, a1 = 0.f;
    const float* dp = dp_s + part * H;
#pragma unroll
    for (int jj = 0; jj < JR; jj += 2) {
      a0 = fmaf(w[jj], dp[jj], a0);
      a1 = fmaf(w[jj + 1], dp[jj + 1], a1);
    }
#pragma unroll 8
    for (int jj = 0; jj < JS; jj += 2) {
      a0 = fmaf(w_s[(size_t)jj * G + t], dp[JR + jj], a0);
      a1 = fmaf(w_s[(size_t)(jj + 1) * G + t], dp[JR + jj + 1], a1);
    }
    red_s[t] = a0 + a1;
    __syncthreads();
    if (part == 0) dh_carry = dh_direct + (red_s[k] + red_s[H + k]) + red_s[2 * H + k];
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------
// Gradient of a final-state-only sequence encoder (instruction_encoder.py:87-90, h_n of the packed
// sequence = the output at token lengths[r]-1): dout (R, L, H) = 0 except dout[r, lengths[r]-1, :] = dh[r].
// ------------------------------------------------------------------------------------
__global__ void scatter_last_token_kernel(const float* __restrict__ dh, long long lddh,
                                          const int* __restrict__ lengths, float* __restrict__ dout, int R,
                                          int L, int H) {
  const long long total = (long long)R * L * H;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % H);
    const int t = (int)((idx / H) % L);
    const int r = (int)(idx / ((long long)H * L));
    dout[idx] = (t == lengths[r] - 1) ? dh[r * lddh + c] : 0.f;
  }
}

// ------------------------------------------------------------------------------------
// embedding-table scatters (few rows per entry -> fp32 atomics)
// ------------------------------------------------------------------------------------
__global__ void embedding_bwd_kernel(const long long* __restrict__ ids, const float* __restrict__ d,
                                     long long ldd, float* __restrict__ dtable, long long rows, int dim,
                                     int vocab, int padding_idx) {
  const long long total = rows * dim;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / dim;
    const int c = (int)(idx % dim);
    const long long id = ids[r];
    if (id < 0 || id >= vocab || id == padding_idx) continue;
    const float v = d[r * ldd + c];
    if (v != 0.f) atomicAdd(dtable + id * dim + c, v);
  }
}

__global__ void prev_action_embed_bwd_kernel(const long long* __restrict__ prev,
                                             const float* __restrict__ masks, const float* __restrict__ d1,
                                             long long ld1, const float* __restrict__ d2, long long ld2,
                                             float* __restrict__ dtable, int rows, int dim, int table_rows) {
  const int total = rows * dim;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int r = idx / dim, c = idx % dim;
    long long id = (long long)(((float)prev[r] + 1.f) * masks[r]);
    if (id < 0 || id >= table_rows) id = 0;
    atomicAdd(dtable + id * dim + c, d1[r * ld1 + c] + (d2 != nullptr ? d2[r * ld2 + c] : 0.f));
  }
}

// d_table_flat[c*P + s] (+)= sum_r dE[r, s, c_off + c] + sum_r dflat[r, c*P + s] + sum_r dmean[r, c] / P
// (the three paths the 64-channel spatial embedding takes: attention K/V source,
//  depth_linear's NCHW flatten, rgb_linear's mean pool).  One warp per (c, s) table entry.
__global__ void spatial_embed_grad_kernel(const float* __restrict__ dE, long long lde_r, long long lde_p,
                                          int c_off, const float* __restrict__ dflat, long long ldf,
                                          const float* __restrict__ dmean, long long ldm,
                                          float* __restrict__ dtable, int accumulate, int R, int P, int Cs) {
  // one warp per table entry, the lanes split the rows (the three sources are strided in r: 64 dependent scattered
  // loads per thread made this the slowest small kernel of the backward chain), shuffle reduction
  const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (idx >= Cs * P) return;
  const int c = idx / P, s = idx % P;
  float acc = 0.f;
  for (int r = lane; r < R; r += 32) {
    if (dE != nullptr) acc += dE[r * lde_r + s * lde_p + c_off + c];
    if (dflat != nullptr) acc += dflat[r * ldf + idx];
    if (dmean != nullptr) acc += dmean[r * ldm + c] / (float)P;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) dtable[idx] = accumulate ? dtable[idx] + acc : acc;
}

// ------------------------------------------------------------------------------------
// Fused IL loss (one CTA): action loss = mean_n( sum_t w CE / sum_t w ), aux = alpha *
// mean over {valid} of the progress MSE in the reference's broadcast form (Q3), and the
// gradients w.r.t. the raw logits and progress_hat.
//   progress_mode 0: no progress monitor; 1: progress (R,) elementwise; 2: progress (R,1)
//   broadcast to (R,R): loss[i,j] = (hat_j - p_i)^2, masked over columns j (aux_losses.py:29).
// ------------------------------------------------------------------------------------
__global__ void il_loss_kernel(const float* __restrict__ lp,      // (R,A) log-probs
                               const long long* __restrict__ actions, const float* __restrict__ weights,
                               const float* __restrict__ hat, const float* __restrict__ progress,
                               int progress_mode, float alpha, float inv_accum, int T, int N, int A,
                               float* __restrict__ d_raw, float* __restrict__ d_hat,
                               float* __restrict__ losses /*3: total, action, aux*/) {
  extern __shared__ float sm[];
  float* wsum = sm;        // N
  float* red = sm + N;     // 32
  const int R = T * N;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += weights[t * N + n];
    wsum[n] = s;
  }
  __syncthreads();
  float a_part = 0.f, cnt = 0.f, sp = 0.f, spp = 0.f;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    const int n = r % N;
    const float w = weights[r];
    const long long a = actions[r];
    const float ce = -lp[(size_t)r * A + a];
    a_part += w * ce / wsum[n];
    cnt += w > 0.f ? 1.f : 0.f;
    if (progress_mode != 0) { const float p = progress[r]; sp += p; spp += p * p; }
  }
  const float action_loss = block_sum(a_part, red) / (float)N;
  const float M = block_sum(cnt, red);
  const float mean_p = block_sum(sp, red) / (float)R;
  const float mean_pp = block_sum(spp, red) / (float)R;
  float aux_part = 0.f;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    const int n = r % N;
    const float w = weights[r];
    const long long a = actions[r];
    const float coef = w / (wsum[n] * (float)N) * inv_accum;
    for (int k = 0; k < A; ++k) {
      const float p = expf(lp[(size_t)r * A + k]);
      d_raw[(size_t)r * A + k] = coef * (p - (k == a ? 1.f : 0.f));
    }
    if (progress_mode != 0) {
      const float h = hat[r];
      const float valid = w > 0.f ? 1.f : 0.f;
      float l, g;
      if (progress_mode == 1) {
        const float d = h - progress[r];
        l = d * d;
        g = 2.f * d;
      } else {
        l = h * h - 2.f * h * mean_p + mean_pp;  // mean_i (h - p_i)^2
        g = 2.f * (h - mean_p);
      }
      aux_part += valid * l;
      d_hat[r] = valid * alpha * g / M * inv_accum;
    }
  }
  float aux = 0.f;
  if (progress_mode != 0) aux = alpha * block_sum(aux_part, red) / M;
  if (threadIdx.x == 0) {
    losses[0] = (action_loss + aux) * inv_accum;
    losses[1] = action_loss;
    losses[2] = aux;
  }
}

// ------------------------------------------------------------------------------------
// Adam (torch.optim.Adam, amsgrad=False, weight_decay=0): one launch over the flat buffers.
// ------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                            float bc1, float bc2_sqrt, float grad_scale, const int* __restrict__ step_dev) {
  if (step_dev != nullptr) {  // CUDA-graph replays: the step count lives on the device
    const float s = (float)(*step_dev);
    bc1 = 1.f - powf(b1, s);
    bc2_sqrt = sqrtf(1.f - powf(b2, s));
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}

__global__ void counter_add_kernel(int* c, int v) { *c += v; }

// Inflection weights of teacher-forcing trajectories (dagger_trainer.py:199-211, recollection_dataset.py:251-263)
// on the padded (T, N) action matrix collate_fn builds: w[t,n] = coef where the oracle action changes (and at
// t = 0), 1 elsewhere, 0 on the padding (t >= lengths[n], what collate_fn's zero fill gives).
__global__ void inflection_weights_kernel(const long long* __restrict__ actions, const int* __restrict__ lengths,
                                          float* __restrict__ w, int T, int N, float coef) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * N) return;
  const int t = idx / N, n = idx % N;
  float v = 0.f;
  if (lengths == nullptr || t < lengths[n]) v = (t == 0 || actions[idx] != actions[idx - N]) ? coef : 1.f;
  w[idx] = v;
}

__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float alpha, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = fmaf(alpha, x[i], y[i]);
}

}  // namespace
}  // namespace vlnce

using namespace vlnce;

extern "C" int vlnce_inflection_weights(const long long* actions, const int* lengths, float* weights, int T, int N,
                                        float coef, cudaStream_t stream) {
  VLNCE_CHECK_ARG(actions && weights && T > 0 && N > 0, "inflection_weights: bad arguments");
  inflection_weights_kernel<<<cdiv((long long)T * N, 256), 256, 0, stream>>>(actions, lengths, weights, T, N, coef);
  VLNCE_LAUNCH_CHECK("inflection_weights_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_counter_add(int* counter, int value, cudaStream_t stream) {
  VLNCE_CHECK_ARG(counter, "counter_add: null");
  counter_add_kernel<<<1, 1, 0, stream>>>(counter, value);
  VLNCE_LAUNCH_CHECK("counter_add_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_axpy(float* y, const float* x, float alpha, long long n, cudaStream_t stream) {
  VLNCE_CHECK_ARG(y && x && n > 0, "axpy: bad arguments");
  axpy_kernel<<<grid_for(n, 256), 256, 0, stream>>>(y, x, alpha, n);
  VLNCE_LAUNCH_CHECK("axpy_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_relu_bwd(float* d, long long ldd, const float* y, long long ldy, int rows, int cols,
                              cudaStream_t stream) {
  VLNCE_CHECK_ARG(d && y && rows > 0 && cols > 0, "relu_bwd: bad arguments");
  relu_bwd_kernel<<<grid_for((long long)rows * cols, 256), 256, 0, stream>>>(d, ldd, y, ldy, rows, cols);
  VLNCE_LAUNCH_CHECK("relu_bwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_tanh_bwd(float* d, const float* y, long long n, cudaStream_t stream) {
  VLNCE_CHECK_ARG(d && y && n > 0, "tanh_bwd: bad arguments");
  tanh_bwd_kernel<<<grid_for(n, 256), 256, 0, stream>>>(d, y, n);
  VLNCE_LAUNCH_CHECK("tanh_bwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_log_softmax_bwd(const float* lp, const float* dlp, float* draw, int rows, int A,
                                     cudaStream_t stream) {
  VLNCE_CHECK_ARG(lp && dlp && draw && rows > 0 && A > 0, "log_softmax_bwd: bad arguments");
  log_softmax_bwd_kernel<<<cdiv(rows, 128), 128, 0, stream>>>(lp, dlp, draw, rows, A);
  VLNCE_LAUNCH_CHECK("log_softmax_bwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_mask_rows_add(const float* src, long long lds, const float* mask, float* dst,
                                   long long ldd, int rows, int H, cudaStream_t stream) {
  VLNCE_CHECK_ARG(src && mask && dst && rows > 0 && H > 0, "mask_rows_add: bad arguments");
  mask_rows_add_kernel<<<grid_for((long long)rows * H, 256), 256, 0, stream>>>(src, lds, mask, dst, ldd, rows, H);
  VLNCE_LAUNCH_CHECK("mask_rows_add_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_gru_pointwise_bwd(const float* dh, long long lddh, const float* save, const float* hm,
                                       float* dgi, float* dgh, float* dhm, int rows, int H,
                                       cudaStream_t stream) {
  VLNCE_CHECK_ARG(dh && save && hm && dgi && dgh && dhm && rows > 0 && H > 0, "gru_pointwise_bwd: bad arguments");
  gru_pointwise_bwd_kernel<<<grid_for((long long)rows * H, 256), 256, 0, stream>>>(dh, lddh, save, hm, dgi, dgh,
                                                                                  dhm, rows, H);
  VLNCE_LAUNCH_CHECK("gru_pointwise_bwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_lstm_pointwise_bwd(const float* dh, long long lddh, const float* dc_in, const float* save,
                                        const float* cm, const float* mask, float* dgates, float* dc_out, int rows,
                                        int H, cudaStream_t stream) {
  VLNCE_CHECK_ARG(dh && save && cm && dgates && dc_out && rows > 0 && H > 0, "lstm_pointwise_bwd: bad arguments");
  lstm_pointwise_bwd_kernel<<<grid_for((long long)rows * H, 256), 256, 0, stream>>>(dh, lddh, dc_in, save, cm, mask,
                                                                                   dgates, dc_out, rows, H);
  VLNCE_LAUNCH_CHECK("lstm_pointwise_bwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_attn_pool_bwd(const float* q, long long ldq, const float* E, long long lde_r,
                                   long long lde_p, const float* probs, float scale, const float* dout,
                                   long long lddo, float* dq, long long lddq, float* dE, long long ldde_r,
                                   long long ldde_p, int accumulate_dE, int R, int P, int D,
                                   cudaStream_t stream) {
  VLNCE_CHECK_ARG(q && E && probs && dout && dq && R > 0 && P > 0 && D > 0, "attn_pool_bwd: bad arguments");
  const size_t smem = sizeof(float) * (P + 32);
  attn_pool_bwd_kernel<<<R, 256, smem, stream>>>(q, ldq, E, lde_r, lde_p, probs, scale, dout, lddo, dq, lddq,
                                                 dE, ldde_r, ldde_p, accumulate_dE, P, D);
  VLNCE_LAUNCH_CHECK("attn_pool_bwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_lstm_seq_bwd(const float* dout, const float* gates, const float* cs, const float* out,
                                  const float* whh_f, const float* whh_b, const int* lengths, float* dgx,
                                  float* hprev, int R, int L, int H, int dirs, cudaStream_t stream) {
  VLNCE_CHECK_ARG(dout && gates && cs && out && whh_f && lengths && dgx && hprev && R > 0 && L > 0 &&
                      (dirs == 1 || dirs == 2),
                  "lstm_seq_bwd: bad arguments");
  VLNCE_CHECK_ARG(H == 128, "lstm_seq_bwd: hidden size %d not built (128 only)", H);
  cudaError_t e = cudaMemsetAsync(dgx, 0, sizeof(float) * (size_t)dirs * R * L * 4 * H, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(hprev, 0, sizeof(float) * (size_t)dirs * R * L * H, stream);
  if (e != cudaSuccess) return cuda_status(e, "lstm_seq_bwd memset");
  constexpr int H_ = 128;
  const int nr = lstm_rows_per_cta(R);   // rows per CTA (common.cuh; same choice as the forward kernel)
  const size_t smem = sizeof(float) * ((size_t)(H_ / 2) * 4 * H_ + (size_t)nr * 8 * H_);
  static bool configured = false;
  if (!configured) {
    const int big = (int)(sizeof(float) * ((size_t)(H_ / 2) * 4 * H_ + 4 * 8 * H_));
    e = cudaFuncSetAttribute(lstm_seq_bwd_kernel<H_, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_seq_bwd_kernel<H_, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_seq_bwd_kernel<H_, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(lstm_seq_bwd)");
    configured = true;
  }
  dim3 grid((R + nr - 1) / nr, dirs);
  if (nr == 4)
    lstm_seq_bwd_kernel<H_, 4><<<grid, 4 * H_, smem, stream>>>(dout, gates, cs, out, whh_f, whh_b, lengths, dgx, hprev,
                                                             R, L, dirs);
  else if (nr == 2)
    lstm_seq_bwd_kernel<H_, 2><<<grid, 4 * H_, smem, stream>>>(dout, gates, cs, out, whh_f, whh_b, lengths, dgx, hprev,
                                                             R, L, dirs);
  else
    lstm_seq_bwd_kernel<H_, 1><<<grid, 4 * H_, smem, stream>>>(dout, gates, cs, out, whh_f, whh_b, lengths, dgx, hprev,
                                                             R, L, dirs);
  VLNCE_LAUNCH_CHECK("lstm_seq_bwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_gru_seq_bwd(const float* dout, const float* save, const float* out, const float* whh_f,
                                 const float* whh_b, const int* lengths, float* dgx, float* dgh, float* hprev, int R,
                                 int L, int H, int dirs, cudaStream_t stream) {
  VLNCE_CHECK_ARG(dout && save && out && whh_f && lengths && dgx && dgh && hprev && R > 0 && L > 0 &&
                      (dirs == 1 || dirs == 2) && (dirs == 1 || whh_b),
                  "gru_seq_bwd: bad arguments");
  VLNCE_CHECK_ARG(H == 128, "gru_seq_bwd: hidden size %d not built (128 only)", H);
  cudaError_t e = cudaMemsetAsync(dgx, 0, sizeof(float) * (size_t)dirs * R * L * 3 * H, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(dgh, 0, sizeof(float) * (size_t)dirs * R * L * 3 * H, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(hprev, 0, sizeof(float) * (size_t)dirs * R * L * H, stream);
  if (e != cudaSuccess) return cuda_status(e, "gru_seq_bwd memset");
  constexpr int H_ = 128;
  const size_t smem = sizeof(float) * ((size_t)(H_ / 2) * 3 * H_ + 6 * H_);
  static bool configured = false;
  if (!configured) {
    e = cudaFuncSetAttribute(gru_seq_bwd_kernel<H_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(gru_seq_bwd)");
    configured = true;
  }
  dim3 grid(R, dirs);
  gru_seq_bwd_kernel<H_><<<grid, 3 * H_, smem, stream>>>(dout, save, out, whh_f, whh_b, lengths, dgx, dgh, hprev, R, L,
                                                       dirs);
  VLNCE_LAUNCH_CHECK("gru_seq_bwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_scatter_last_token(const float* dh, long long lddh, const int* lengths, float* dout, int R,
                                        int L, int H, cudaStream_t stream) {
  VLNCE_CHECK_ARG(dh && lengths && dout && R > 0 && L > 0 && H > 0, "scatter_last_token: bad arguments");
  scatter_last_token_kernel<<<grid_for((long long)R * L * H, 256), 256, 0, stream>>>(dh, lddh, lengths, dout, R, L, H);
  VLNCE_LAUNCH_CHECK("scatter_last_token_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_embedding_bwd(const long long* ids, const float* d, long long ldd, float* dtable,
                                   long long rows, int dim, int vocab, int padding_idx,
                                   cudaStream_t stream) {
  VLNCE_CHECK_ARG(ids && d && dtable && rows > 0 && dim > 0, "embedding_bwd: bad arguments");
  embedding_bwd_kernel<<<grid_for(rows * dim, 256), 256, 0, stream>>>(ids, d, ldd, dtable, rows, dim, vocab,
                                                                      padding_idx);
  VLNCE_LAUNCH_CHECK("embedding_bwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_prev_action_embed_bwd(const long long* prev_actions, const float* masks, const float* d1,
                                           long long ld1, const float* d2, long long ld2, float* dtable,
                                           int rows, int dim, int table_rows, cudaStream_t stream) {
  VLNCE_CHECK_ARG(prev_actions && masks && d1 && dtable && rows > 0, "prev_action_embed_bwd: bad arguments");
  prev_action_embed_bwd_kernel<<<grid_for((long long)rows * dim, 256), 256, 0, stream>>>(
      prev_actions, masks, d1, ld1, d2, ld2, dtable, rows, dim, table_rows);
  VLNCE_LAUNCH_CHECK("prev_action_embed_bwd_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_spatial_embed_grad(const float* dE, long long lde_r, long long lde_p, int c_off,
                                        const float* dflat, long long ldf, const float* dmean, long long ldm,
                                        float* dtable, int accumulate, int R, int P, int Cs,
                                        cudaStream_t stream) {
  VLNCE_CHECK_ARG(dtable && R > 0 && P > 0 && Cs > 0, "spatial_embed_grad: bad arguments");
  spatial_embed_grad_kernel<<<cdiv((long long)Cs * P * 32, 256), 256, 0, stream>>>(dE, lde_r, lde_p, c_off, dflat, ldf, dmean, ldm,
                                                                 dtable, accumulate, R, P, Cs);
  VLNCE_LAUNCH_CHECK("spatial_embed_grad_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_il_loss(const float* logp, const long long* actions, const float* weights,
                             const float* hat, const float* progress, int progress_mode, float alpha,
                             float inv_accum, int T, int N, int A, float* d_raw, float* d_hat, float* losses,
                             cudaStream_t stream) {
  VLNCE_CHECK_ARG(logp && actions && weights && d_raw && losses && T > 0 && N > 0 && A > 0, "il_loss: bad arguments");
  VLNCE_CHECK_ARG(progress_mode == 0 || (hat && progress && d_hat), "il_loss: progress tensors missing");
  VLNCE_CHECK_ARG(N <= 8192, "il_loss: N=%d too large for one CTA", N);
  const size_t smem = sizeof(float) * (N + 32);
  il_loss_kernel<<<1, 1024, smem, stream>>>(logp, actions, weights, hat, progress, progress_mode, alpha, inv_accum,
                                            T, N, A, d_raw, d_hat, losses);
  VLNCE_LAUNCH_CHECK("il_loss_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                               float beta2, float eps, int step, const int* step_dev, float grad_scale,
                               cudaStream_t stream) {
  VLNCE_CHECK_ARG(p && g && m && v && n > 0 && (step > 0 || step_dev), "adam_step: bad arguments");
  const float bc1 = step > 0 ? 1.f - powf(beta1, (float)step) : 1.f;
  const float bc2_sqrt = step > 0 ? sqrtf(1.f - powf(beta2, (float)step)) : 1.f;
  adam_kernel<<<grid_for(n, 256), 256, 0, stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1, bc2_sqrt, grad_scale,
                                                    step_dev);
  VLNCE_LAUNCH_CHECK("adam_kernel");
  return VLNCE_OK;
}
