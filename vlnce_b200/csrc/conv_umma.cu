// Implicit-GEMM (im2col-free) NHWC convolution on the 5th-gen tensor cores.
//
//   Y[m, k] = act( scale[k] * sum_{r,s,c} X[n, p*st+r-pad, q*st+s-pad, c] * W[k, r, s, c]
//                  + shift[k] (+ residual[m, k]) ),   m = (n, p, q)
//
// Replaces the cuDNN convolutions behind torchvision ResNet-50 (`resnet_encoders.py:198-199`)
// and the DD-PPO GroupNorm ResNet-50 (`resnet_encoders.py:95`, habitat_baselines v0.1.7
// `ResNetEncoder`) on the policy hot path.  GEMM view: M = N*P*Q pixels, N = K filters,
// K = R*S*C.  One CTA computes a 128 x BLOCK_N output tile:
//
//   warp 5     TMA producer: streams the weight tile (BLOCK_N x BLOCK_K of the [K, R*S*C]
//              matrix) and — for 1x1 stride-1 convolutions, where the activation IS a row-major
//              [M, C] matrix — the activation tile too; prefetches the residual tile.
//   warps 0-3  for every other geometry gather the A tile (128 pixels x BLOCK_K channels of one
//              filter tap) straight from the NHWC activation with 16-byte cp.async into the
//              swizzled K-major layout the MMA expects (zero-fill = padding); then the epilogue:
//              tcgen05.ld the fp32 accumulator rows from TMEM, folded-norm scale/shift, residual,
//              ReLU, pack, and hand the tile back to TMA as swizzled shared-memory boxes.
//   warp 4     allocates TMEM and issues tcgen05.mma (one elected thread), committing each smem
//              stage back to the producers through an mbarrier.
//
// An mbarrier ring of 1..4 stages (shallow K loops take less smem -> more CTAs per SM) decouples
// the roles.  Activations stay 16-bit NHWC in HBM; the accumulator never leaves TMEM until the
// fused epilogue; loads and stores of whole tiles are TMA (coalesced, no LSU address traffic).
#include "common.cuh"
#include "ptx_sm100.cuh"

#include <stdlib.h>

namespace vlnce {
namespace {

using namespace ptx;

constexpr int BLOCK_M = 128;
constexpr int kMaxStages = 4;
constexpr int kLag = 2;  // cp.async groups kept in flight before a stage is published
constexpr int kProducerThreads = 128;
constexpr int kThreads = 192;
constexpr int kSmemBudget = 227 * 1024;

struct ConvParams {
  const __nv_bfloat16* x;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  void* y;
  int N, H, W, C, K, R, S, stride, pad, P, Q;
  int M;            // N*P*Q
  int c_chunks;     // C / BLOCK_K
  int num_k_iters;  // R*S*c_chunks
  int n_tiles;      // K / BLOCK_N
  int relu;
  int stages;       // smem ring depth (1..4)
  int y_fp32;       // 1: store fp32 instead of the 16-bit activation type
  int ab_half;      // 1: operands / residual / 16-bit output are fp16, else bf16
  int a_tma;        // 1: activation tile comes by TMA (1x1 stride-1 pad-0)
  int stem;         // 1: x is the raw (N,H,W,3) u8 image, 2: f32 image — 7x7/2 pad-3 RGB stem, the
                    //    im2col row (r,s,c) is formed on the fly (K padded 147 -> 192)
  int stem_patch;   // byte offset of the input-patch scratch after the staging buffer (0: none)
  int out_tma;      // 1: epilogue stages the tile in smem and stores it with TMA
};

template <int BLOCK_N, int BLOCK_K>
struct Cfg {
  static constexpr int ROW_BYTES = BLOCK_K * 2;
  static constexpr int CPR = ROW_BYTES / 16;             // 16B chunks per operand row
  static constexpr int ROWS_PER_PASS = kProducerThreads / CPR;
  static constexpr int RPT = BLOCK_M / ROWS_PER_PASS;    // rows gathered per producer thread
  static constexpr int A_BYTES = BLOCK_M * ROW_BYTES;
  static constexpr int B_BYTES = BLOCK_N * ROW_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 1024;                 // barriers + tmem slot, keeps alignment
  static constexpr int staging_bytes(bool out_tma, bool y_fp32) {
    return out_tma ? BLOCK_M * BLOCK_N * (y_fp32 ? 4 : 2) : 0;
  }
  static constexpr int smem_bytes(int stages, bool out_tma, bool y_fp32, int extra = 0) {
    return stages * STAGE_BYTES + BAR_BYTES + staging_bytes(out_tma, y_fp32) + extra + 1024;  // + align slack
  }
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "operand tiles must stay 1024B aligned");
};

template <int BLOCK_N, int BLOCK_K>
__global__ void __launch_bounds__(kThreads, 3)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                 const __grid_constant__ CUtensorMap tmap_y, const __grid_constant__ CUtensorMap tmap_res,
                 const ConvParams p) {
  using C = Cfg<BLOCK_N, BLOCK_K>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int stages = p.stages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + stages * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* accum_bar = empty_bar + kMaxStages;
  uint64_t* res_bar = accum_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 1);
  const uint32_t smem_a0 = smem_u32(smem);
  const uint32_t smem_b0 = smem_a0 + stages * C::A_BYTES;
  const uint32_t stage_out = smem_a0 + stages * C::STAGE_BYTES + C::BAR_BYTES;  // 1024B aligned

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x % p.n_tiles;
  const int m_tile = blockIdx.x / p.n_tiles;
  const int m0 = m_tile * BLOCK_M;
  const int n0 = n_tile * BLOCK_N;
  const bool has_res = p.residual != nullptr;

  if (threadIdx.x == 5 * 32) {
    for (int s = 0; s < stages; ++s) {
      // gather mode: 128 producer threads + the TMA issuer; TMA mode: the issuer alone
      mbar_init(&full_bar[s], p.a_tma ? 1 : kProducerThreads + 1);
      mbar_init(&empty_bar[s], 1);  // one tcgen05.commit
    }
    mbar_init(accum_bar, 1);
    mbar_init(res_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_w);
    if (p.a_tma) tma_prefetch_desc(&tmap_x);
    if (p.out_tma) tma_prefetch_desc(&tmap_y);
  }
  if (warp == 4) {
    tmem_alloc(tmem_slot, BLOCK_N);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();  // PDL (common.cuh): TMEM is ours; dependents may be scheduled, operands are read after the wait
  pdl_wait();

  if (warp < 4) {
    if (p.stem) {
      // ---------------------------------------------------------- RGB stem: im2col on the fly
      // A row = the 7x7x3 window of one output pixel, K index (r*7+s)*3+c; this thread fills the
      // 8-element (16 B) chunk `chunk` of rows t/8 + 16 i.  Byte loads hit L1 (each input byte is
      // used by ~12 windows); values 0..255 are exact in bf16/fp16.
      const int t = threadIdx.x;
      const int chunk = t % 8;
      const uint8_t* xb = reinterpret_cast<const uint8_t*>(p.x);
      const float* xf = reinterpret_cast<const float*>(p.x);
      int ih0[8], iw0[8], img_off[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = m0 + t / 8 + i * 16;
        if (m < p.M) {
          const int pq = p.P * p.Q;
          const int n_img = m / pq;
          const int rem = m - n_img * pq;
          const int po = rem / p.Q;
          ih0[i] = po * 2 - 3;
          iw0[i] = (rem - po * p.Q) * 2 - 3;
          img_off[i] = n_img * p.H * p.W * 3;
        } else {
          ih0[i] = -(1 << 20);
          iw0[i] = 0;
          img_off[i] = 0;
        }
      }
      // Fast path (one tile = one output row, Q % 128 == 0): stage the 7 input rows the tile reads
      // in shared memory once (coalesced 16-byte loads, zero-filled padding), then build the
      // im2col chunks from shared memory — no per-element global loads or bounds checks.
      constexpr int PATCH_W = 788;  // (2*127 + 6 + 1 + 1) * 3 bytes rounded up; x-origin at iw = -4
      const bool patch = p.stem_patch != 0;
      uint8_t* pb = smem + stages * C::STAGE_BYTES + C::BAR_BYTES + p.stem_patch;
      if (patch) {
        const int pq = p.P * p.Q;
        const int n_img = m0 / pq;
        const int rem = m0 - n_img * pq;
        const int po = rem / p.Q, q0 = rem - po * p.Q;
        const int esz = p.stem == 1 ? 1 : 4;
        const int row_bytes = p.W * 3 * esz;                 // multiple of 16 (W % 128 == 0)
        const int x_off = (2 * q0 - 4) * 3 * esz;            // byte offset of the patch origin in a row
        // 7 rows x 197 four-byte granules (788 B) per row, 2 granules per thread and row: all 14
        // loads are issued before the first shared-memory store (one L2 round trip, not fourteen)
        constexpr int G = PATCH_W / 4;  // granules per row for u8; f32 rows are 4x longer
        if (esz == 1) {
          uint32_t v[7][2];
#pragma unroll
          for (int r = 0; r < 7; ++r) {
            const int ih = po * 2 - 3 + r;
            const bool row_ok = (unsigned)ih < (unsigned)p.H;
            const uint8_t* grow = xb + ((size_t)(n_img * p.H + ih) * p.W) * 3;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int w4 = t + h * kProducerThreads;
              const int gb = x_off + w4 * 4;
              v[r][h] = (row_ok && w4 < G && gb >= 0 && gb + 4 <= row_bytes)
                            ? __ldg(reinterpret_cast<const uint32_t*>(grow + gb)) : 0u;
            }
          }
#pragma unroll
          for (int r = 0; r < 7; ++r)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int w4 = t + h * kProducerThreads;
              if (w4 < G) *reinterpret_cast<uint32_t*>(pb + r * PATCH_W + w4 * 4) = v[r][h];
            }
        } else {
          for (int r = 0; r < 7; ++r) {
            const int ih = po * 2 - 3 + r;
            uint8_t* prow = pb + r * PATCH_W * 4;
            const bool row_ok = (unsigned)ih < (unsigned)p.H;
            const uint8_t* grow = xb + ((size_t)(n_img * p.H + ih) * p.W) * 3 * 4;
            uint32_t v[7];
#pragma unroll
            for (int h = 0; h < 7; ++h) {  // 788 granules per f32 row: 7 per thread
              const int w4 = t + h * kProducerThreads;
              const int gb = x_off + w4 * 4;
              v[h] = (row_ok && w4 < PATCH_W && gb >= 0 && gb + 4 <= row_bytes)
                         ? __ldg(reinterpret_cast<const uint32_t*>(grow + gb)) : 0u;
            }
#pragma unroll
            for (int h = 0; h < 7; ++h) {
              const int w4 = t + h * kProducerThreads;
              if (w4 < PATCH_W) *reinterpret_cast<uint32_t*>(prow + w4 * 4) = v[h];
            }
          }
        }
        named_barrier_sync(2, kProducerThreads);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < p.num_k_iters; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 1);
        const uint32_t a_stage = smem_a0 + stage * C::A_BYTES;
        int dr[8], ds[8], dc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int kk = it * 64 + chunk * 8 + e;
          const int tap = kk / 3;
          dc[e] = kk - tap * 3;
          dr[e] = kk < 147 ? tap / 7 : (1 << 20);  // padded K columns read nothing
          ds[e] = tap % 7;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = t / 8 + i * 16;
          float v[8];
          if (patch) {
            const int qrow = (m0 + row) % p.Q - (m0 % p.Q);  // pixel index inside the tile's row
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int off = dr[e] * PATCH_W + (2 * qrow + ds[e] + 1) * 3 + dc[e];
              v[e] = dr[e] >= 7 ? 0.f : (p.stem == 1 ? (float)pb[off] : reinterpret_cast<const float*>(pb)[off]);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int ih = ih0[i] + dr[e], iw = iw0[i] + ds[e];
              const bool ok = (unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W;
              const int off = img_off[i] + (ih * p.W + iw) * 3 + dc[e];
              v[e] = ok ? (p.stem == 1 ? (float)__ldg(xb + off) : __ldg(xf + off)) : 0.f;
            }
          }
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (p.ab_half) {
              __half2 h2 = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
              o[e] = *reinterpret_cast<uint32_t*>(&h2);
            } else {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
              o[e] = *reinterpret_cast<uint32_t*>(&h2);
            }
          }
          const uint32_t addr = a_stage + swizzled_offset(row, chunk, 128);
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[0]), "r"(o[1]), "r"(o[2]),
                       "r"(o[3])
                       : "memory");
        }
        fence_proxy_async_smem();
        mbar_arrive(&full_bar[stage]);
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    } else if (!p.a_tma) {
      // ---------------------------------------------------------- A gather (producers)
      const int t = threadIdx.x;
      const int chunk = t % C::CPR;
      int ih0[C::RPT], iw0[C::RPT], img_off[C::RPT];
      uint32_t dst_off[C::RPT];
#pragma unroll
      for (int i = 0; i < C::RPT; ++i) {
        const int row = t / C::CPR + i * C::ROWS_PER_PASS;
        const int m = m0 + row;
        dst_off[i] = swizzled_offset(row, chunk, C::ROW_BYTES);
        if (m < p.M) {
          const int pq = p.P * p.Q;
          const int n_img = m / pq;
          const int rem = m - n_img * pq;
          const int po = rem / p.Q;
          const int qo = rem - po * p.Q;
          ih0[i] = po * p.stride - p.pad;
          iw0[i] = qo * p.stride - p.pad;
          img_off[i] = n_img * p.H * p.W * p.C;
        } else {
          ih0[i] = -(1 << 20);  // always out of range -> zero rows
          iw0[i] = 0;
          img_off[i] = 0;
        }
      }
      int rf = 0, sf = 0, cc = 0;
      const int lag = min(kLag, stages - 1);
      int stage = 0, pub = 0;  // ring slot being filled / next slot to publish
      uint32_t phase = 0;
      for (int it = 0; it < p.num_k_iters; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 1);
        const uint32_t a_stage = smem_a0 + stage * C::A_BYTES;
        const int c_off = cc * BLOCK_K + chunk * 8;
#pragma unroll
        for (int i = 0; i < C::RPT; ++i) {
          const int ih = ih0[i] + rf, iw = iw0[i] + sf;
          const bool ok = (unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W;
          const __nv_bfloat16* src =
              ok ? p.x + (size_t)img_off[i] + (size_t)(ih * p.W + iw) * p.C + c_off : p.x;
          cp_async_16(a_stage + dst_off[i], src, ok ? 16u : 0u);
        }
        cp_async_commit();
        if (it >= lag) {
          if (lag == 2) cp_async_wait<2>();
          else if (lag == 1) cp_async_wait<1>();
          else cp_async_wait<0>();
          fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the MMA (async proxy)
          mbar_arrive(&full_bar[pub]);
          if (++pub == stages) pub = 0;
        }
        if (++cc == p.c_chunks) {
          cc = 0;
          if (++sf == p.S) { sf = 0; ++rf; }
        }
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      cp_async_wait<0>();
      fence_proxy_async_smem();
      for (int it = (p.num_k_iters > lag ? p.num_k_iters - lag : 0); it < p.num_k_iters; ++it) {
        mbar_arrive(&full_bar[pub]);
        if (++pub == stages) pub = 0;
      }
    }

    // ------------------------------------------------------------ epilogue
    mbar_wait(accum_bar, 0, 2);
    tc_fence_after_sync();
    const int row = warp * 32 + lane;
    const int m = m0 + row;
    const bool row_ok = m < p.M;
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    if (p.out_tma && has_res) mbar_wait(res_bar, 0, 5);
#pragma unroll 1
    for (int cb = 0; cb < BLOCK_N / 32; ++cb) {
      uint32_t v[32];
      tmem_ld_32x32(t_row + cb * 32, v);
      tmem_ld_wait();
      const int col0 = n0 + cb * 32;
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      if (p.scale != nullptr) {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + col0) + j4);
          const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + col0) + j4);
          f[j4 * 4 + 0] = fmaf(f[j4 * 4 + 0], sc.x, sh.x);
          f[j4 * 4 + 1] = fmaf(f[j4 * 4 + 1], sc.y, sh.y);
          f[j4 * 4 + 2] = fmaf(f[j4 * 4 + 2], sc.z, sh.z);
          f[j4 * 4 + 3] = fmaf(f[j4 * 4 + 3], sc.w, sh.w);
        }
      }
      if (p.out_tma) {
        // tile staged as 128-byte-row boxes (64 x 16-bit or 32 x fp32 columns), 128B-swizzled:
        // conflict-free 16-byte accesses, and exactly the layout the TMA store reads.
        if (p.y_fp32) {
          const uint32_t box = stage_out + (uint32_t)cb * (BLOCK_M * 128) + (uint32_t)row * 128;
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const uint32_t addr = box + (uint32_t)((q4 ^ (row & 7)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(f[q4 * 4]),
                         "f"(f[q4 * 4 + 1]), "f"(f[q4 * 4 + 2]), "f"(f[q4 * 4 + 3])
                         : "memory");
          }
        } else {
          const uint32_t box = stage_out + (uint32_t)(cb >> 1) * (BLOCK_M * 128) + (uint32_t)row * 128;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const uint32_t addr = box + (uint32_t)((((cb & 1) * 4 + q4) ^ (row & 7)) << 4);
            if (has_res) {
              uint32_t w0, w1, w2, w3;
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                           : "r"(addr)
                           : "memory");
              const uint32_t w[4] = {w0, w1, w2, w3};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (p.ab_half) {
                  const float2 h2 = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
                  f[q4 * 8 + e * 2 + 0] += h2.x;
                  f[q4 * 8 + e * 2 + 1] += h2.y;
                } else {
                  f[q4 * 8 + e * 2 + 0] += __uint_as_float(w[e] << 16);
                  f[q4 * 8 + e * 2 + 1] += __uint_as_float(w[e] & 0xffff0000u);
                }
              }
            }
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float a = f[q4 * 8 + e * 2], b = f[q4 * 8 + e * 2 + 1];
              if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
              if (p.ab_half) {
                __half2 h2 = __floats2half2_rn(a, b);
                o[e] = *reinterpret_cast<uint32_t*>(&h2);
              } else {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
                o[e] = *reinterpret_cast<uint32_t*>(&h2);
              }
            }
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[0]), "r"(o[1]),
                         "r"(o[2]), "r"(o[3])
                         : "memory");
          }
        }
      } else if (row_ok) {
        // direct path (narrow tiles): every thread stores its own row segment
        const size_t off = (size_t)m * p.K + col0;
        if (has_res) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.residual + off);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const uint4 rv = __ldg(rp + q4);
            const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (p.ab_half) {
                const float2 h2 = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
                f[q4 * 8 + e * 2 + 0] += h2.x;
                f[q4 * 8 + e * 2 + 1] += h2.y;
              } else {
                f[q4 * 8 + e * 2 + 0] += __uint_as_float(w[e] << 16);
                f[q4 * 8 + e * 2 + 1] += __uint_as_float(w[e] & 0xffff0000u);
              }
            }
          }
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        if (p.y_fp32) {
          float4* yp = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + off);
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4)
            yp[q4] = make_float4(f[q4 * 4], f[q4 * 4 + 1], f[q4 * 4 + 2], f[q4 * 4 + 3]);
        } else {
          uint4* yp = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y) + off);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (p.ab_half) {
                __half2 h2 = __floats2half2_rn(f[q4 * 8 + e * 2], f[q4 * 8 + e * 2 + 1]);
                w[e] = *reinterpret_cast<uint32_t*>(&h2);
              } else {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(f[q4 * 8 + e * 2], f[q4 * 8 + e * 2 + 1]);
                w[e] = *reinterpret_cast<uint32_t*>(&h2);
              }
            }
            yp[q4] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
    }
    if (p.out_tma) {
      fence_proxy_async_smem();                     // smem tile -> visible to the TMA engine
      named_barrier_sync(1, kProducerThreads);      // the four epilogue warps only
      if (threadIdx.x == 0) {
        const int boxes = p.y_fp32 ? BLOCK_N / 32 : BLOCK_N / 64;
        const int cols = p.y_fp32 ? 32 : 64;
        for (int b = 0; b < boxes; ++b)
          tma_store_2d(&tmap_y, stage_out + (uint32_t)b * (BLOCK_M * 128), n0 + b * cols, m0);
        tma_store_commit();
        tma_store_wait_read();                      // smem must outlive the bulk reads
      }
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_f16(BLOCK_M, BLOCK_N, p.ab_half ? 0u : 1u);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < p.num_k_iters; ++it) {
      mbar_wait(&full_bar[stage], phase, 3);
      tc_fence_after_sync();
      if (lane == 0) {
        const uint64_t adesc = make_smem_desc(smem_a0 + stage * C::A_BYTES, C::ROW_BYTES);
        const uint64_t bdesc = make_smem_desc(smem_b0 + stage * C::B_BYTES, C::ROW_BYTES);
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k) {
          // +32 bytes (two 16B units) per K=16 step inside the swizzled row
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)((it | k) != 0));
        }
        umma_commit(&empty_bar[stage]);  // stage reusable once these MMAs have read it
      }
      __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
    if (lane == 0) umma_commit(accum_bar);
    __syncwarp();
  } else {
    // ------------------------------------------------------------ TMA producer (warp 5)
    if (lane == 0) {
      if (p.out_tma && has_res) {  // residual tile -> staging buffer, overlapped with the main loop
        mbar_arrive_expect_tx(res_bar, BLOCK_M * BLOCK_N * 2);
        for (int b = 0; b < BLOCK_N / 64; ++b)
          tma_load_2d(stage_out + (uint32_t)b * (BLOCK_M * 128), &tmap_res, res_bar, n0 + b * 64, m0);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < p.num_k_iters; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 4);
        mbar_arrive_expect_tx(&full_bar[stage], p.a_tma ? C::STAGE_BYTES : C::B_BYTES);
        if (p.a_tma) tma_load_2d(smem_a0 + stage * C::A_BYTES, &tmap_x, &full_bar[stage], it * BLOCK_K, m0);
        tma_load_2d(smem_b0 + stage * C::B_BYTES, &tmap_w, &full_bar[stage], it * BLOCK_K, n0);
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, BLOCK_N);
}

// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

// row-major [rows, cols] matrix, box = box_cols x box_rows, inner box span = 64 or 128 bytes
int make_map_2d(EncodeTiledFn encode, CUtensorMap* map, const void* base, CUtensorMapDataType dt, int elem_bytes,
                long long rows, long long cols, int box_cols, int box_rows, const char* what) {
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)cols * elem_bytes};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const int span = box_cols * elem_bytes;
  CUresult cr = encode(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       span == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("conv2d_bf16: cuTensorMapEncodeTiled(%s) failed (%d)", what, (int)cr);
    return VLNCE_ERR_CUDA_BASE + 1;
  }
  return VLNCE_OK;
}

template <int BLOCK_N, int BLOCK_K>
int launch(const CUtensorMap& tw, const CUtensorMap& tx, const CUtensorMap& ty, const CUtensorMap& tr,
           ConvParams& p, int stages_cap, int exclusive, cudaStream_t stream) {
  using C = Cfg<BLOCK_N, BLOCK_K>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel<BLOCK_N, BLOCK_K>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(conv_umma)");
    configured = true;
  }
  // Ring depth: deep enough to cover the K loop, but small enough that TWO CTAs fit on an SM
  // (one CTA's epilogue then overlaps the other's main loop); VLNCE_CONV_SMEM_KB overrides the
  // per-CTA budget for experiments.
  static int budget = -1;
  if (budget < 0) {
    const char* e = getenv("VLNCE_CONV_SMEM_KB");
    budget = e ? atoi(e) * 1024 : 113 * 1024;
    if (budget > kSmemBudget) budget = kSmemBudget;
  }
  // stem fast path: 7 x 788 input elements of scratch (u8 or f32) behind the staging buffer
  int extra = 0;
  p.stem_patch = 0;
  if (p.stem && p.Q % BLOCK_M == 0 && p.W % 128 == 0) extra = 7 * 788 * (p.stem == 1 ? 1 : 4) + 64;
  int stages = p.num_k_iters < kMaxStages ? p.num_k_iters : kMaxStages;
  if (stages_cap > 0) {  // tuned by the caller: ring depth given, only the hardware limit applies
    if (stages > stages_cap) stages = stages_cap;
    while (stages > 1 && C::smem_bytes(stages, p.out_tma, p.y_fp32, extra) > kSmemBudget) --stages;
  } else {
    while (stages > 1 && C::smem_bytes(stages, p.out_tma, p.y_fp32, extra) > budget) --stages;
  }
  if (C::smem_bytes(stages, p.out_tma, p.y_fp32, extra) > kSmemBudget) {  // staged tile too large: store directly
    p.out_tma = 0;
    stages = p.num_k_iters < kMaxStages ? p.num_k_iters : kMaxStages;
  }
  p.stages = stages;
  if (extra) p.stem_patch = C::staging_bytes(p.out_tma, p.y_fp32) + 16;  // never 0 when enabled
  const long long tiles = (long long)cdiv(p.M, BLOCK_M) * p.n_tiles;
  // "exclusive": request just over half of the SM's shared memory so that exactly ONE of these CTAs
  // is resident per SM and the other half stays free for kernels of concurrent graph branches
  // (depth trunk / head run beside the RGB trunk instead of queueing behind it).
  int smem = C::smem_bytes(stages, p.out_tma, p.y_fp32, extra);
  if (exclusive && smem < 116 * 1024) smem = 116 * 1024;
  launch_pdl(conv_umma_kernel<BLOCK_N, BLOCK_K>, dim3((unsigned)tiles), dim3(kThreads), (size_t)smem, stream, tw, tx, ty,
             tr, p);
  VLNCE_LAUNCH_CHECK("conv_umma_kernel");
  return VLNCE_OK;
}

}  // namespace
}  // namespace vlnce

using namespace vlnce;

static int conv2d_impl(const void* x, int stem, const void* w, int ab_dtype, const float* scale,
                       const float* shift, const void* residual, void* y, int y_dtype, int N, int H, int W,
                       int C, int K, int R, int S, int stride, int pad, int relu, int block_n,
                       cudaStream_t stream);

extern "C" int vlnce_conv2d_bf16(const void* x, const void* w, int ab_dtype, const float* scale,
                                 const float* shift, const void* residual, void* y, int y_dtype,
                                 int N, int H, int W, int C, int K, int R, int S, int stride,
                                 int pad, int relu, int block_n, cudaStream_t stream) {
  return conv2d_impl(x, 0, w, ab_dtype, scale, shift, residual, y, y_dtype, N, H, W, C, K, R, S, stride, pad, relu,
                     block_n, stream);
}

extern "C" int vlnce_conv2d_stem_rgb(const void* x, int x_dtype, const void* w, int ab_dtype,
                                     const float* scale, const float* shift, void* y, int y_dtype, int N,
                                     int H, int W, int K, int relu, int block_n, cudaStream_t stream) {
  VLNCE_CHECK_ARG(x_dtype == VLNCE_U8 || x_dtype == VLNCE_F32, "conv2d_stem_rgb: image must be u8 or f32");
  // geometry handed to the tile kernel: an (N*P*Q) x 192 GEMM whose A rows are built from the image
  return conv2d_impl(x, x_dtype == VLNCE_U8 ? 1 : 2, w, ab_dtype, scale, shift, nullptr, y, y_dtype, N, H, W, 192,
                     K, 1, 1, 1, 0, relu, block_n, stream);
}

static int conv2d_impl(const void* x, int stem, const void* w, int ab_dtype, const float* scale,
                                 const float* shift, const void* residual, void* y, int y_dtype,
                                 int N, int H, int W, int C, int K, int R, int S, int stride,
                                 int pad, int relu, int block_n, cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && w && y, "conv2d_bf16: null tensor");
  VLNCE_CHECK_ARG(N > 0 && H > 0 && W > 0 && R > 0 && S > 0 && stride > 0 && pad >= 0,
                  "conv2d_bf16: bad geometry");
  VLNCE_CHECK_ARG(C % 32 == 0, "conv2d_bf16: C=%d must be a multiple of 32", C);
  VLNCE_CHECK_ARG(K % 32 == 0, "conv2d_bf16: K=%d must be a multiple of 32", K);
  VLNCE_CHECK_ARG((scale == nullptr) == (shift == nullptr), "conv2d_bf16: scale/shift must come together");
  VLNCE_CHECK_ARG(ab_dtype == VLNCE_BF16 || ab_dtype == VLNCE_F16, "conv2d_bf16: operand dtype must be bf16 or f16");
  VLNCE_CHECK_ARG(y_dtype == ab_dtype || y_dtype == VLNCE_F32, "conv2d_bf16: y dtype must be the operand type or f32");
  VLNCE_CHECK_ARG(!(residual && y_dtype == VLNCE_F32), "conv2d_bf16: residual needs a 16-bit output");
  VLNCE_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(residual) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(scale) & 15) == 0 && (reinterpret_cast<uintptr_t>(shift) & 15) == 0,
                  "conv2d_bf16: tensors must be 16-byte aligned");
  const int P = stem ? (H + 6 - 7) / 2 + 1 : (H + 2 * pad - R) / stride + 1;
  const int Q = stem ? (W + 6 - 7) / 2 + 1 : (W + 2 * pad - S) / stride + 1;
  VLNCE_CHECK_ARG(P > 0 && Q > 0, "conv2d_bf16: empty output");
  VLNCE_CHECK_ARG((long long)N * H * W * C < (1ll << 31) && (long long)N * P * Q * K < (1ll << 40),
                  "conv2d_bf16: tensor too large for 32-bit offsets");
  const int block_k = (C % 64 == 0) ? 64 : 32;
  const int m_tiles = cdiv((long long)N * P * Q, BLOCK_M);
  const int stages_cap = (block_n >> 16) & 0xf;  // optional tuned ring depth in bits [16,20)
  const int exclusive = (block_n >> 20) & 1;      // bit 20: one CTA per SM (see launch())
  block_n &= 0xffff;
  int bn = block_n;
  if (bn == 0) {
    // widest tile that still gives every SM (148) a CTA; otherwise the narrowest legal one >= 64
    static const int cand[4] = {256, 128, 64, 32};
    for (int i = 0; i < 4 && bn == 0; ++i)
      if (K % cand[i] == 0 && (long long)m_tiles * (K / cand[i]) >= 148) bn = cand[i];
    if (bn == 0) bn = (K % 64 == 0) ? 64 : 32;
  }
  VLNCE_CHECK_ARG((bn == 32 || bn == 64 || bn == 128 || bn == 256) && K % bn == 0,
                  "conv2d_bf16: block_n=%d does not tile K=%d", bn, K);

  EncodeTiledFn encode = get_encode_fn();
  if (!encode) {
    set_error("conv2d_bf16: cuTensorMapEncodeTiled unavailable (no driver)");
    return VLNCE_ERR_NO_DRIVER;
  }
  const CUtensorMapDataType dt16 = ab_dtype == VLNCE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const long long M = (long long)N * P * Q;

  ConvParams p;
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.scale = scale;
  p.shift = shift;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.y = y;
  p.N = N; p.H = H; p.W = W; p.C = C; p.K = K; p.R = R; p.S = S;
  p.stride = stride; p.pad = pad; p.P = P; p.Q = Q;
  p.M = (int)M;
  p.c_chunks = C / block_k;
  p.num_k_iters = R * S * p.c_chunks;
  p.n_tiles = K / bn;
  p.relu = relu;
  p.stages = 1;
  p.y_fp32 = (y_dtype == VLNCE_F32);
  p.ab_half = (ab_dtype == VLNCE_F16);
  p.stem = stem;
  p.a_tma = (!stem && R == 1 && S == 1 && stride == 1 && pad == 0) ? 1 : 0;
  p.out_tma = (bn >= 64) ? 1 : 0;

  CUtensorMap tw, tx, ty, tr;
  int rc = make_map_2d(encode, &tw, w, dt16, 2, K, (long long)R * S * C, block_k, bn, "weights");
  if (rc != VLNCE_OK) return rc;
  tx = tw; ty = tw; tr = tw;  // placeholders when a role is unused
  if (p.a_tma) {
    rc = make_map_2d(encode, &tx, x, dt16, 2, M, C, block_k, BLOCK_M, "activation");
    if (rc != VLNCE_OK) return rc;
  }
  if (p.out_tma) {
    if (p.y_fp32)
      rc = make_map_2d(encode, &ty, y, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, K, 32, BLOCK_M, "output");
    else
      rc = make_map_2d(encode, &ty, y, dt16, 2, M, K, 64, BLOCK_M, "output");
    if (rc != VLNCE_OK) return rc;
    if (residual) {
      rc = make_map_2d(encode, &tr, residual, dt16, 2, M, K, 64, BLOCK_M, "residual");
      if (rc != VLNCE_OK) return rc;
    }
  }

#define VLNCE_CONV_CASE(BN, BK) \
  if (bn == BN && block_k == BK) return launch<BN, BK>(tw, tx, ty, tr, p, stages_cap, exclusive, stream);
  VLNCE_CONV_CASE(256, 64)
  VLNCE_CONV_CASE(128, 64)
  VLNCE_CONV_CASE(64, 64)
  VLNCE_CONV_CASE(32, 64)
  VLNCE_CONV_CASE(256, 32)
  VLNCE_CONV_CASE(128, 32)
  VLNCE_CONV_CASE(64, 32)
  VLNCE_CONV_CASE(32, 32)
#undef VLNCE_CONV_CASE
  set_error("conv2d_bf16: no kernel for block_n=%d block_k=%d", bn, block_k);
  return VLNCE_ERR_INVALID_ARGUMENT;
}
