// Image-stationary GroupNorm-ResNet trunk on tcgen05: the whole depth encoder
// (habitat_baselines v0.1.7 ResNetEncoder behind resnet_encoders.py:85-115 — stem, max-pool, the
// bottleneck stages, the 3x3 compression, every GroupNorm / residual / ReLU) in ONE launch.
//
// GroupNorm normalises per sample, so an image never needs another image's data: every CTA owns
// whole images and walks the layer program ("steps") of its images by itself; there is no
// inter-CTA synchronisation anywhere.  Per conv step:
//
//   * both operands are K-major swizzled tiles delivered by TMA: weights from the packed [Cout, R*S*Cin]
//     matrix, activations as tap-shifted windows of the image (borders arrive as zeros; stride-2
//     convs read four parity views of their input), i.e. im2col by tensor map, no gather code;
//   * the accumulators of the WHOLE image stay in TMEM (512 columns): a "tiled" operand supplies the
//     128 MMA rows, a "slab" operand the N columns.  Large images (>= 128 pixels) tile the pixels
//     and slab the filters; small images (8x8, 4x4) swap the roles (tiles of 128 filters, the
//     image's pixels as N) so that no MMA row is padding;
//   * GroupNorm runs out of TMEM in two passes of tcgen05.ld: statistics (fp32 partials, fp64
//     combine), then normalise (+ residual)(+ ReLU) and store the 16-bit activation for the next
//     step.  The raw conv output never leaves the SM.
//
// Activations live in a caller-owned scratch that stays L2-resident (an image's largest tensor is
// 256 KB).  Roles: warp 0 / 1 = TMA producers of the tiled / slab rings, warp 2 = MMA issuer,
// (and TMEM owner), warps 3-10 = epilogue (two warps per TMEM lane quarter).  The weight producer
// runs ahead across step boundaries; activation loads wait for the step that produced them.
#include "common.cuh"
#include "ptx_sm100.cuh"

#include <stdlib.h>
#include <vector>

namespace vlnce {
namespace {

using namespace ptx;

constexpr int kT = 4;                  // tiled-operand ring: transactions of two 128-byte-row tiles or four 64-byte-row tiles
constexpr int kS = 3;                  // slab-operand ring: transactions of up to eight k-chunks (<= 16 KB)
constexpr int kTBytes = 32 * 1024;
constexpr int kSBytes = 16 * 1024;
constexpr int kEpiWarps = 8;
constexpr int kEpi = kEpiWarps * 32;   // 256 epilogue threads
constexpr int kThreads = 3 * 32 + kEpi;   // T producer, S producer, MMA issuer (+ TMEM owner), epilogue
constexpr int kMaxGroups = 16;         // groups per pass (non-swapped: per slab; swapped: per layer)
constexpr int kMaxCout = 1024;
constexpr int kMaxSlab = 128;          // MMA N per pass
constexpr int kStageBytes = 32 * 1024; // swapped epilogue: fp32 staging of one 128-filter tile x <= 64 pixels
                                       // (aliases the per-channel statistics chs[cout][2] float2 of pass 1)
constexpr int kSmemBytes = kT * kTBytes + kS * kSBytes + kStageBytes + 1024 /*stats+barriers*/ + 1024 /*align*/;

struct FStep {
  int kind;        // 0 conv + GroupNorm (+residual)(+ReLU), 1 maxpool 3x3/2 pad 1
  int swapped;     // 0: tiles = 128-pixel row blocks, slab = filters; 1: tiles = 128 filters, slab = the image's pixels
  int kc;          // K elements per chunk (64: 128-byte swizzle, 32: 64-byte swizzle)
  int taps;        // 1 | 9
  int cin_chunks;  // cin / kc
  int stride;      // 1 | 2
  int n_tiles;     // 128-row tiles of the tiled operand
  int slab_rows;   // MMA N
  int n_slabs;     // passes over the accumulator (filters split so that a pass fits the 512 TMEM columns)
  int cin, cout, hout, wout, px;
  int rows_per_tile;  // non-swapped: image rows per pixel tile
  int groups, gs, relu;
  int map_w;
  int map_x[4];
  int external_input;  // activation input produced before the launch (stem operand)
  int sper;            // k-chunks of the slab operand per ring transaction (= one TMA request)
  int tper;            // (k-chunk, tile) units of the tiled operand per ring transaction (= one TMA request)
  int w_pairs_k;       // swapped, one filter tile: a weight request covers two consecutive k-chunks (else two filter tiles)
  int hin, win;        // maxpool input size
  int ld_f32;
  float eps;
  const float* gamma;
  const float* beta;
  const __half* res;
  __half* out;
  float* out_f32;
  const __half* pool_in;
};

struct PlanHeader {
  int n_steps, n_maps, n_images, magic;
  long long steps_off, maps_off, total_bytes, pad;
};
constexpr int kMagic = 0x44465031;  // "DFP1"

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// generic-proxy global writes of this thread -> ordered before later async-proxy (TMA) reads of global memory
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1,
                                            int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst_smem),
      "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst_smem, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1,
                                            int32_t c2, int32_t c3, int32_t c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst_smem),
      "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
// Wait until `steps_done` >= want (monotonic counter: immune to the phase-parity aliasing a producer
// that runs several steps ahead of the epilogue would hit on an mbarrier).
__device__ __forceinline__ void wait_steps(const uint32_t* counter, uint32_t want, int tag) {
  if (ld_acquire_u32(counter) >= want) return;
  const uint64_t t0 = global_timer_ns();
  for (uint32_t i = 1;; ++i) {
    if (ld_acquire_u32(counter) >= want) return;
    __nanosleep(100);
    if ((i & 255u) == 0 && global_timer_ns() - t0 > 2000000000ull) break;
  }
  printf("vlnce: depth_fused step wait timed out (tag %d, block %d, want %u)\n", tag, blockIdx.x, want);
  __trap();
}

// Debug trace (vlnce_depth_fused_set_trace): CTA 0 / its first image records %globaltimer per (step, slab):
//   [0] MMA issuer starts the pass   [1] last MMA committed   [2] epilogue sees the accumulators
//   [3] statistics done   [4] pass stored   [5] activation producer released by the previous step
//   [6] / [7] ns the MMA issuer spent blocked on the slab / tile ring during the pass
constexpr int kTraceSlots = 8;

// Sum of v[k] over the 32 lanes for all 32 k at once: 31 shuffles, lane l returns the total of v[l].
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < off; ++k) {
      const float send = up ? v[k] : v[k + off];
      const float keep = up ? v[k + off] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

__device__ __forceinline__ void window_of_tap(int tap, int stride, int& mi, int& xo, int& yo) {
  const int r = tap / 3, q = tap - r * 3;
  if (stride == 1) { mi = 0; xo = q - 1; yo = r - 1; return; }
  // stride 2: input row 2y + r - 1 = parity (r != 1), row index y + (r == 0 ? -1 : 0) of that view
  mi = (r != 1 ? 2 : 0) + (q != 1 ? 1 : 0);
  yo = r == 0 ? -1 : 0;
  xo = q == 0 ? -1 : 0;
}

__global__ void __launch_bounds__(kThreads, 1)
depth_fused_kernel(const uint8_t* __restrict__ plan, int n_images, long long* __restrict__ trace) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const uint32_t t_ring = smem_u32(smem);
  const uint32_t s_ring = t_ring + kT * kTBytes;
  uint8_t* scratch = smem + kT * kTBytes + kS * kSBytes;
  float* stage = reinterpret_cast<float*>(scratch);                 // [pixel][128 filters] fp32 | chs[channel][half] float2
  uint8_t* tail = scratch + kStageBytes;
  float2* stat = reinterpret_cast<float2*>(tail);                   // (mean, rstd) per group          [16]
  float* part = reinterpret_cast<float*>(tail + 128);               // [group][quarter][sum|sq] floats [16*8]
  double* red = reinterpret_cast<double*>(tail + 640);              // cross-warp scratch              [16]
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 768);
  uint64_t* t_full = bars;
  uint64_t* t_empty = t_full + kT;
  uint64_t* s_full = t_empty + kT;
  uint64_t* s_empty = s_full + kS;
  uint64_t* acc_full = s_empty + kS;
  uint64_t* tmem_free = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_free + 1);
  uint32_t* steps_done = tmem_slot + 1;

  const PlanHeader* hdr = reinterpret_cast<const PlanHeader*>(plan);
  const FStep* steps = reinterpret_cast<const FStep*>(plan + hdr->steps_off);
  const CUtensorMap* maps = reinterpret_cast<const CUtensorMap*>(plan + hdr->maps_off);
  const int n_steps = hdr->n_steps;
  if (blockIdx.x != 0) trace = nullptr;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kT; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 1); }
    for (int i = 0; i < kS; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 1); }
    mbar_init(acc_full, 1);
    mbar_init(tmem_free, kEpi);
    *steps_done = 0;
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ tiled-operand producer
    // A ring transaction = up to `tper` consecutive (k-chunk, tile) units behind one barrier (32 KB).
    if (lane == 0) {
      int ts = 0;
      uint32_t tph = 0, gstep = 0;
      for (int img = blockIdx.x; img < n_images; img += gridDim.x) {
        for (int si = 0; si < n_steps; ++si, ++gstep) {
          const FStep* sp = steps + si;
          if (sp->kind != 0) continue;
          const bool is_x = !sp->swapped;
          const int kc = sp->kc, cin = sp->cin, n_tiles = sp->n_tiles, rpt = sp->rows_per_tile;
          const int cin_chunks = sp->cin_chunks, n_slabs = sp->n_slabs, stride = sp->stride, taps = sp->taps;
          const int tper = sp->tper;
          const uint32_t tile_bytes = 128u * (uint32_t)kc * 2u;
          const CUtensorMap* mw = maps + sp->map_w;
          const CUtensorMap* mxs[4] = {maps + sp->map_x[0], maps + sp->map_x[1], maps + sp->map_x[2], maps + sp->map_x[3]};
          {  // descriptor prefetch for the next conv step (its first loads would otherwise pay the fetch latency)
            for (int sj = si + 1; sj < n_steps; ++sj) {
              const FStep* nx = steps + sj;
              if (nx->kind != 0) continue;
              tma_prefetch_desc(maps + nx->map_w);
              tma_prefetch_desc(maps + nx->map_x[0]);
              if (nx->map_x[1] != nx->map_x[0]) {
                tma_prefetch_desc(maps + nx->map_x[1]);
                tma_prefetch_desc(maps + nx->map_x[2]);
                tma_prefetch_desc(maps + nx->map_x[3]);
              }
              break;
            }
          }
          if (is_x && !sp->external_input) {
            wait_steps(steps_done, gstep, 1);
            if (trace && img == (int)blockIdx.x) trace[(si * 2) * kTraceSlots + 5] = (long long)global_timer_ns();
          }
          // ONE TMA request per ring transaction (the TMA unit serves ~4 requests / us whatever their size):
          // activations = tper consecutive pixel tiles (one taller box); weights = two filter tiles (256 rows) or,
          // with a single filter tile, two consecutive k-chunks (3-D box over the k-chunk dimension)
          const uint32_t tx_bytes = tile_bytes * (uint32_t)tper;
          const bool pairs_k = sp->w_pairs_k != 0;
          for (int slab = 0; slab < n_slabs; ++slab) {
            int kk = 0;
            for (int tap = 0; tap < taps; ++tap) {
              int mi = 0, xo = 0, yo = 0;
              if (taps == 9) window_of_tap(tap, stride, mi, xo, yo);
              const CUtensorMap* mx = mi == 0 ? mxs[0] : (mi == 1 ? mxs[1] : (mi == 2 ? mxs[2] : mxs[3]));
              for (int cc = 0; cc < cin_chunks; ++cc, ++kk) {
                if (!is_x && pairs_k) {
                  if (kk & 1) continue;     // the request issued at the even k-chunk covered this one
                  mbar_wait(&t_empty[ts], tph ^ 1, 40);
                  mbar_arrive_expect_tx(&t_full[ts], tx_bytes);
                  tma_load_3d(t_ring + ts * kTBytes, mw, &t_full[ts], 0, 0, kk);
                  if (++ts == kT) { ts = 0; tph ^= 1; }
                  continue;
                }
                for (int i = 0; i < n_tiles; i += tper) {
                  mbar_wait(&t_empty[ts], tph ^ 1, 40);
                  mbar_arrive_expect_tx(&t_full[ts], tx_bytes);
                  if (is_x) tma_load_4d(t_ring + ts * kTBytes, mx, &t_full[ts], cc * kc, xo, i * rpt + yo, img);
                  else tma_load_3d(t_ring + ts * kTBytes, mw, &t_full[ts], 0, i * 128, kk);
                  if (++ts == kT) { ts = 0; tph ^= 1; }
                }
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ slab-operand producer
    // A ring transaction = up to `sper` consecutive k-chunks of the slab operand behind one barrier (<= 16 KB).
    if (lane == 0) {
      int ss = 0;
      uint32_t sph = 0, gstep = 0;
      for (int img = blockIdx.x; img < n_images; img += gridDim.x) {
        for (int si = 0; si < n_steps; ++si, ++gstep) {
          const FStep* sp = steps + si;
          if (sp->kind != 0) continue;
          const bool is_x = sp->swapped != 0;
          const int kc = sp->kc, cin = sp->cin, slab_rows = sp->slab_rows;
          const int cin_chunks = sp->cin_chunks, n_slabs = sp->n_slabs, stride = sp->stride, taps = sp->taps;
          const int KK = taps * cin_chunks;
          const uint32_t chunk_bytes = (uint32_t)slab_rows * (uint32_t)kc * 2u;
          const int sper = sp->sper;
          const CUtensorMap* mw = maps + sp->map_w;
          const CUtensorMap* mxs[4] = {maps + sp->map_x[0], maps + sp->map_x[1], maps + sp->map_x[2], maps + sp->map_x[3]};
          if (is_x && !sp->external_input) {
            wait_steps(steps_done, gstep, 2);
            if (trace && img == (int)blockIdx.x) trace[(si * 2) * kTraceSlots + 5] = (long long)global_timer_ns();
          }
          const uint32_t tx_bytes = chunk_bytes * (uint32_t)sper;   // a box past the tensor's end is zero-filled but counts in full
          for (int slab = 0; slab < n_slabs; ++slab) {
            if (!is_x) {   // weights: 3-D box over sper consecutive k-chunks
              for (int kk = 0; kk < KK; kk += sper) {
                mbar_wait(&s_empty[ss], sph ^ 1, 41);
                mbar_arrive_expect_tx(&s_full[ss], tx_bytes);
                tma_load_3d(s_ring + ss * kSBytes, mw, &s_full[ss], 0, slab * slab_rows, kk);
                if (++ss == kS) { ss = 0; sph ^= 1; }
              }
              continue;
            }
            for (int tap = 0; tap < taps; ++tap) {   // activations: 5-D box over sper channel chunks of one tap window
              int mi = 0, xo = 0, yo = 0;
              if (taps == 9) window_of_tap(tap, stride, mi, xo, yo);
              const CUtensorMap* mx = mi == 0 ? mxs[0] : (mi == 1 ? mxs[1] : (mi == 2 ? mxs[2] : mxs[3]));
              for (int cc = 0; cc < cin_chunks; cc += sper) {
                mbar_wait(&s_empty[ss], sph ^ 1, 41);
                mbar_arrive_expect_tx(&s_full[ss], tx_bytes);
                tma_load_5d(s_ring + ss * kSBytes, mx, &s_full[ss], 0, xo, yo, img, cc);
                if (++ss == kS) { ss = 0; sph ^= 1; }
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      int ts = 0, ss = 0;
      uint32_t tph = 0, sph = 0, pass = 0;
      for (int img = blockIdx.x; img < n_images; img += gridDim.x) {
        for (int si = 0; si < n_steps; ++si) {
          const FStep* sp = steps + si;
          if (sp->kind != 0) continue;
          const int slab_rows = sp->slab_rows, n_tiles = sp->n_tiles, n_slabs = sp->n_slabs;
          const int KK = sp->taps * sp->cin_chunks, U = KK * n_tiles;
          const uint32_t idesc = make_idesc_f16(128, (uint32_t)slab_rows, 0u);
          const uint32_t row_bytes = (uint32_t)sp->kc * 2u;
          const bool k4 = sp->kc == 64;
          const int tper = sp->tper, sper = sp->sper;
          const uint32_t tile_step = (128u * row_bytes) >> 4;                  // descriptor units (16 B)
          const uint32_t chunk_step = ((uint32_t)slab_rows * row_bytes) >> 4;
          const uint64_t desc_hi = make_smem_desc(0, row_bytes);
          const bool tr = trace && img == (int)blockIdx.x;
          for (int slab = 0; slab < n_slabs; ++slab, ++pass) {
            if (tr) trace[(si * 2 + (slab & 1)) * kTraceSlots + 0] = (long long)global_timer_ns();
            mbar_wait(tmem_free, (pass & 1) ^ 1, 42);  // epilogue has drained the previous pass
            tc_fence_after_sync();
            int u = 0, tsub = 0, ssub = 0;
            uint64_t bdesc = 0, adesc = 0;
            for (int kk = 0; kk < KK; ++kk) {
              if (ssub == 0) {
                mbar_wait(&s_full[ss], sph, 43);
                tc_fence_after_sync();
                bdesc = desc_hi | (uint64_t)(((s_ring + ss * kSBytes) & 0x3FFFFu) >> 4);
              }
              const uint32_t acc = kk != 0 ? 1u : 0u;
              for (int i = 0; i < n_tiles; ++i, ++u) {
                if (tsub == 0) {
                  mbar_wait(&t_full[ts], tph, 44);
                  tc_fence_after_sync();
                  adesc = desc_hi | (uint64_t)(((t_ring + ts * kTBytes) & 0x3FFFFu) >> 4);
                }
                const uint32_t d = tmem_base + (uint32_t)(i * slab_rows);
                umma_bf16(d, adesc, bdesc, idesc, acc);
                umma_bf16(d, adesc + 2, bdesc + 2, idesc, 1u);
                if (k4) {
                  umma_bf16(d, adesc + 4, bdesc + 4, idesc, 1u);
                  umma_bf16(d, adesc + 6, bdesc + 6, idesc, 1u);
                }
                adesc += tile_step;
                if (++tsub == tper || u == U - 1) {
                  umma_commit(&t_empty[ts]);
                  tsub = 0;
                  if (++ts == kT) { ts = 0; tph ^= 1; }
                }
              }
              bdesc += chunk_step;
              if (++ssub == sper || kk == KK - 1) {
                umma_commit(&s_empty[ss]);
                ssub = 0;
                if (++ss == kS) { ss = 0; sph ^= 1; }
              }
            }
            umma_commit(acc_full);
            if (tr) trace[(si * 2 + (slab & 1)) * kTraceSlots + 1] = (long long)global_timer_ns();
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue (warps 3-10)
    const int ew = warp - 3;
    const int q = warp & 3;          // TMEM lane quarter this warp may read
    const int half = ew >> 2;        // two warps per quarter split the 16-column chunks
    const int r = q * 32 + lane;     // accumulator row (pixel in tile | filter in tile)
    const int tid = threadIdx.x - 96;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t pass = 0, gstep = 0;
    for (int img = blockIdx.x; img < n_images; img += gridDim.x) {
      for (int si = 0; si < n_steps; ++si, ++gstep) {
        const FStep* sp = steps + si;
        const bool tr = trace && img == (int)blockIdx.x && tid == 0;
        // all epilogue threads have finished writing the previous step (residual / pool inputs
        // are read below by other threads than the ones that wrote them)
        named_barrier_sync(1, kEpi);
        if (sp->kind == 1) {
          // ---- MaxPool2d(3, 2, 1) on the 16-bit NHWC image
          const int cin = sp->cin, hin = sp->hin, win = sp->win;
          const int C8 = cin / 8, P = sp->hout, Q = sp->wout;
          const __half* in = sp->pool_in + (size_t)img * hin * win * cin;
          __half* out = sp->out + (size_t)img * P * Q * cin;
          for (int idx = tid; idx < P * Q * C8; idx += kEpi) {
            const int c8 = idx % C8;
            const int m = idx / C8;
            const int qo = m % Q, po = m / Q;
            __half2 best[4];
            bool any = false;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const int ih = po * 2 + dy - 1;
              if ((unsigned)ih >= (unsigned)hin) continue;
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                const int iw = qo * 2 + dx - 1;
                if ((unsigned)iw >= (unsigned)win) continue;
                const uint4 v = *reinterpret_cast<const uint4*>(in + ((size_t)ih * win + iw) * cin + c8 * 8);
                const __half2* hv = reinterpret_cast<const __half2*>(&v);
                if (!any) {
#pragma unroll
                  for (int e = 0; e < 4; ++e) best[e] = hv[e];
                  any = true;
                } else {
#pragma unroll
                  for (int e = 0; e < 4; ++e) best[e] = __hmax2(best[e], hv[e]);
                }
              }
            }
            *reinterpret_cast<uint4*>(out + (size_t)m * cin + c8 * 8) = *reinterpret_cast<const uint4*>(best);
          }
        } else {
          const int slab_rows = sp->slab_rows, n_tiles = sp->n_tiles, n_slabs = sp->n_slabs;
          const int gs = sp->gs, cout = sp->cout, px = sp->px, relu = sp->relu, ld_f32 = sp->ld_f32;
          const int gs_shift = 31 - __clz(gs);
          const float eps = sp->eps;
          const float* gamma = sp->gamma;
          const float* beta = sp->beta;
          const __half* res = sp->res;
          __half* outp = sp->out;
          float* out_f32 = sp->out_f32;
          const int nchunks = slab_rows / 16;       // 16-column chunks of a tile
          for (int slab = 0; slab < n_slabs; ++slab, ++pass) {
            long long* trow = tr ? trace + (si * 2 + (slab & 1)) * kTraceSlots : nullptr;
            mbar_wait(acc_full, pass & 1, 45);
            tc_fence_after_sync();
            if (trow) trow[2] = (long long)global_timer_ns();
            if (!sp->swapped) {
              // ================= tiles = pixels (this thread: row r of every tile), columns = filters of the slab
              // ---- pass 1: per-column sums over the thread's pixels in registers (chunk outer, tile inner),
              // then one 31-shuffle transpose-reduce per chunk and a group fold
              for (int ch = half; ch < nchunks; ch += 2) {
                float v[32];
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = 0.f;
                uint32_t vb[2][16];
                const uint32_t col = lane_addr + (uint32_t)(ch * 16);
                tmem_ld_32x16(col, vb[0]);
                for (int i0 = 0; i0 < n_tiles; i0 += 2) {
#pragma unroll
                  for (int b = 0; b < 2; ++b) {
                    const int i = i0 + b;
                    if (i < n_tiles) {
                      tmem_ld_wait();
                      if (i + 1 < n_tiles) tmem_ld_32x16(col + (uint32_t)((i + 1) * slab_rows), vb[b ^ 1]);
#pragma unroll
                      for (int e = 0; e < 16; ++e) {
                        const float f = __uint_as_float(vb[b][e]);
                        v[e] += f;
                        v[16 + e] = fmaf(f, f, v[16 + e]);
                      }
                    }
                  }
                }
                float tot = warp_transpose_reduce32(v, lane);   // lane l < 16: sum of column l; else sumsq of column l - 16
                for (int off = 1; off < gs; off <<= 1) tot += __shfl_xor_sync(0xffffffffu, tot, off);
                if ((lane & (gs - 1)) == 0)
                  part[((ch * 16 + (lane & 15)) >> gs_shift) * 8 + q * 2 + (lane >> 4)] = tot;
              }
              named_barrier_sync(1, kEpi);
              const int ngs = slab_rows >> gs_shift;
              if (tid < ngs) {   // fp64 combine of the four lane quarters
                double ds = 0.0, dq = 0.0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  ds += (double)part[tid * 8 + k * 2];
                  dq += (double)part[tid * 8 + k * 2 + 1];
                }
                const double cnt = (double)px * (double)gs;
                const double mean = ds / cnt;
                double var = dq / cnt - mean * mean;
                if (var < 0.0) var = 0.0;
                stat[tid] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
              }
              named_barrier_sync(1, kEpi);
              if (trow) trow[3] = (long long)global_timer_ns();
              // ---- pass 2: normalise (+ residual)(+ ReLU), 16-bit store; the chunk's scale / shift stay in registers,
              // TMEM loads are double-buffered and the residual segment is requested three tiles ahead.  (A variant that
              // transposed the tile through shared memory for fully coalesced 16-byte stores measured 30 % slower.)
              for (int ch = half; ch < nchunks; ch += 2) {
                float sc[16], sh[16];
                const int cbase = slab * slab_rows + ch * 16;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                  const float2 st = stat[(ch * 16 + e) >> gs_shift];
                  sc[e] = st.y * __ldg(gamma + cbase + e);
                  sh[e] = __ldg(beta + cbase + e) - st.x * sc[e];
                }
                uint32_t vb[2][16];
                uint4 rb[4][2];
                const uint32_t col = lane_addr + (uint32_t)(ch * 16);
                const size_t o0 = ((size_t)img * px + (size_t)r) * cout + cbase;   // tile i: + i * 128 * cout
                const size_t tile_stride = (size_t)128 * cout;
                tmem_ld_32x16(col, vb[0]);
                if (res != nullptr) {
#pragma unroll
                  for (int k = 0; k < 3; ++k) {
                    if (k < n_tiles) {
                      rb[k][0] = *reinterpret_cast<const uint4*>(res + o0 + (size_t)k * tile_stride);
                      rb[k][1] = *reinterpret_cast<const uint4*>(res + o0 + (size_t)k * tile_stride + 8);
                    }
                  }
                }
                for (int i0 = 0; i0 < n_tiles; i0 += 4) {
#pragma unroll
                  for (int b = 0; b < 4; ++b) {
                    const int i = i0 + b;
                    if (i < n_tiles) {
                      const size_t o = o0 + (size_t)i * tile_stride;
                      tmem_ld_wait();
                      if (i + 1 < n_tiles) tmem_ld_32x16(col + (uint32_t)((i + 1) * slab_rows), vb[(b + 1) & 1]);
                      if (res != nullptr && i + 3 < n_tiles) {
                        rb[(b + 3) & 3][0] = *reinterpret_cast<const uint4*>(res + o + 3 * tile_stride);
                        rb[(b + 3) & 3][1] = *reinterpret_cast<const uint4*>(res + o + 3 * tile_stride + 8);
                      }
                      float f[16];
#pragma unroll
                      for (int e = 0; e < 16; ++e) f[e] = fmaf(__uint_as_float(vb[b & 1][e]), sc[e], sh[e]);
                      if (res != nullptr) {
                        const __half2* rh = reinterpret_cast<const __half2*>(&rb[b][0]);
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                          const float2 t = __half22float2(rh[e]);
                          f[2 * e] += t.x;
                          f[2 * e + 1] += t.y;
                        }
                      }
                      if (relu) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) f[e] = fmaxf(f[e], 0.f);
                      }
                      if (outp != nullptr) {
                        __align__(16) __half2 hv[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) hv[e] = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
                        *reinterpret_cast<uint4*>(outp + o) = *reinterpret_cast<const uint4*>(&hv[0]);
                        *reinterpret_cast<uint4*>(outp + o + 8) = *reinterpret_cast<const uint4*>(&hv[4]);
                      }
                      if (out_f32 != nullptr) {
                        float* of = out_f32 + ((size_t)img * px + (size_t)(i * 128 + r)) * ld_f32 + cbase;
#pragma unroll
                        for (int e = 0; e < 16; ++e) of[e] = f[e];
                      }
                    }
                  }
                }
              }
              tc_fence_before_sync();
              mbar_arrive(tmem_free);     // this thread is done with the accumulator
            } else {
              // ================= tiles = filters (this thread: filter r of every tile), columns = the image's pixels
              float2* chs = reinterpret_cast<float2*>(stage);  // [channel][half] partial (sum, sumsq)
              // ---- pass 1: this thread's filter summed over its pixel chunks
              for (int i = 0; i < n_tiles; ++i) {
                float sum = 0.f, sq = 0.f;
                for (int ch = half; ch < nchunks; ch += 2) {
                  uint32_t v[16];
                  tmem_ld_32x16(lane_addr + (uint32_t)(i * slab_rows + ch * 16), v);
                  tmem_ld_wait();
#pragma unroll
                  for (int e = 0; e < 16; ++e) {
                    const float f = __uint_as_float(v[e]);
                    sum += f;
                    sq = fmaf(f, f, sq);
                  }
                }
                chs[(i * 128 + r) * 2 + half] = make_float2(sum, sq);
              }
              named_barrier_sync(1, kEpi);
              {
                const int G = sp->groups;
                const int tpg = kEpi / G;          // threads per group (host: 256 % G == 0, tpg <= 32 or G == 1)
                const int g = tid / tpg, sub = tid - g * tpg;
                double ds = 0.0, dq = 0.0;
                for (int e = sub; e < 2 * gs; e += tpg) {
                  const float2 t = chs[g * gs * 2 + e];
                  ds += (double)t.x;
                  dq += (double)t.y;
                }
                const int width = tpg < 32 ? tpg : 32;
                for (int o = width >> 1; o > 0; o >>= 1) {
                  ds += __shfl_xor_sync(0xffffffffu, ds, o);
                  dq += __shfl_xor_sync(0xffffffffu, dq, o);
                }
                if (tpg > 32) {   // one group over all eight warps
                  if (lane == 0) { red[ew * 2] = ds; red[ew * 2 + 1] = dq; }
                  named_barrier_sync(1, kEpi);
                  ds = 0.0; dq = 0.0;
                  for (int w8 = 0; w8 < kEpiWarps; ++w8) { ds += red[w8 * 2]; dq += red[w8 * 2 + 1]; }
                }
                if (sub == 0) {
                  const double cnt = (double)px * (double)gs;
                  const double mean = ds / cnt;
                  double var = dq / cnt - mean * mean;
                  if (var < 0.0) var = 0.0;
                  stat[g] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
                }
              }
              named_barrier_sync(1, kEpi);
              if (trow) trow[3] = (long long)global_timer_ns();
              // ---- pass 2, tile by tile: normalised fp32 tile -> smem [pixel][128 filters] (transpose), then a
              // coalesced copy-out in 8-filter vectors that adds the residual, applies ReLU and rounds
              const int vecs = px * 16;
              for (int i = 0; i < n_tiles; ++i) {
                const int c = i * 128 + r;
                const float2 st = stat[c >> gs_shift];
                const float scv = st.y * __ldg(gamma + c);
                const float shv = __ldg(beta + c) - st.x * scv;
                for (int ch = half; ch < nchunks; ch += 2) {
                  uint32_t v[16];
                  tmem_ld_32x16(lane_addr + (uint32_t)(i * slab_rows + ch * 16), v);
                  tmem_ld_wait();
#pragma unroll
                  for (int e = 0; e < 16; ++e) stage[(ch * 16 + e) * 128 + r] = fmaf(__uint_as_float(v[e]), scv, shv);
                }
                if (i == n_tiles - 1) {   // last TMEM read of the pass: the MMA issuer may start the next one
                  tc_fence_before_sync();
                  mbar_arrive(tmem_free);
                }
                named_barrier_sync(1, kEpi);
                for (int t = tid; t < vecs; t += kEpi) {
                  const int p = t >> 4, c8 = (t & 15) * 8;
                  const float4 a = *reinterpret_cast<const float4*>(stage + p * 128 + c8);
                  const float4 b4 = *reinterpret_cast<const float4*>(stage + p * 128 + c8 + 4);
                  float f[8] = {a.x, a.y, a.z, a.w, b4.x, b4.y, b4.z, b4.w};
                  const size_t o = ((size_t)img * px + p) * cout + i * 128 + c8;
                  if (res != nullptr) {
                    const uint4 rv = *reinterpret_cast<const uint4*>(res + o);
                    const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      const float2 t2 = __half22float2(rh[e]);
                      f[2 * e] += t2.x;
                      f[2 * e + 1] += t2.y;
                    }
                  }
                  if (relu) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
                  }
                  if (outp != nullptr) {
                    __align__(16) __half2 hv[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) hv[e] = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
                    *reinterpret_cast<uint4*>(outp + o) = *reinterpret_cast<const uint4*>(hv);
                  }
                  if (out_f32 != nullptr) {
                    float* of = out_f32 + ((size_t)img * px + p) * ld_f32 + i * 128 + c8;
                    *reinterpret_cast<float4*>(of) = make_float4(f[0], f[1], f[2], f[3]);
                    *reinterpret_cast<float4*>(of + 4) = make_float4(f[4], f[5], f[6], f[7]);
                  }
                }
                named_barrier_sync(1, kEpi);   // the staging tile is free again
              }
            }
            if (trow) trow[4] = (long long)global_timer_ns();
          }
        }
        // publish the step: generic-proxy global writes -> visible to the TMA (async proxy) loads of later steps
        fence_proxy_async_all();
        __threadfence_block();
        named_barrier_sync(1, kEpi);
        if (tid == 0) st_release_u32(steps_done, gstep + 1);
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encoder() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

int encode(EncodeTiledFn enc, CUtensorMap* m, const void* base, int rank, const cuuint64_t* gdim,
           const cuuint64_t* gstr, const cuuint32_t* box, int kc, const char* what) {
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("depth_fused: cuTensorMapEncodeTiled(%s) failed (%d)", what, (int)r);
    return VLNCE_ERR_CUDA_BASE + 1;
  }
  return VLNCE_OK;
}

constexpr int kStepFields = 20;
long long* g_fused_trace = nullptr;

}  // namespace
}  // namespace vlnce

using namespace vlnce;

// ---- C ABI --------------------------------------------------------------------------------------
// A step table is `n_steps` rows of VLNCE_DEPTH_STEP_FIELDS (20) int64:
//   [0] kind (0 conv+GroupNorm, 1 maxpool3x3s2)   [1] x (N,hin,win,cin) f16   [2] w (cout, k*k*cin) f16
//   [3] gamma f32   [4] beta f32   [5] residual (N,hout,wout,cout) f16 | 0   [6] y f16 | 0
//   [7] y_f32 (N*hout*wout, ld) | 0   [8] ld of y_f32   [9] cin   [10] cout   [11] k (1|3)   [12] stride (1|2)
//   [13] hin   [14] win   [15] groups   [16] relu   [17] external input (x written before the launch)
//   [18] eps (float bits)   [19] reserved

extern "C" long long vlnce_depth_fused_plan_bytes(int n_steps) {
  if (n_steps <= 0) return 0;
  long long b = 128;                                                   // header
  b += (long long)((n_steps * (long long)sizeof(FStep) + 127) / 128) * 128;
  b += (long long)n_steps * 5 * 128;                                   // <= 5 tensor maps per step
  return b;
}

extern "C" int vlnce_depth_fused_plan(const long long* table, int n_steps, int n_images, void* plan_host,
                                      long long plan_bytes) {
  VLNCE_CHECK_ARG(table && plan_host && n_steps > 0 && n_images > 0, "depth_fused_plan: null table / plan");
  VLNCE_CHECK_ARG(plan_bytes >= vlnce_depth_fused_plan_bytes(n_steps), "depth_fused_plan: plan buffer too small");
  EncodeTiledFn enc = get_encoder();
  if (!enc) {
    set_error("depth_fused_plan: cuTensorMapEncodeTiled unavailable");
    return VLNCE_ERR_NO_DRIVER;
  }
  uint8_t* base = reinterpret_cast<uint8_t*>(plan_host);
  memset(base, 0, (size_t)plan_bytes);
  PlanHeader* hdr = reinterpret_cast<PlanHeader*>(base);
  hdr->n_steps = n_steps;
  hdr->n_images = n_images;
  hdr->magic = kMagic;
  hdr->steps_off = 128;
  hdr->maps_off = 128 + (long long)((n_steps * (long long)sizeof(FStep) + 127) / 128) * 128;
  FStep* steps = reinterpret_cast<FStep*>(base + hdr->steps_off);
  // tensor maps are staged in an aligned array (CUtensorMap wants 64-byte alignment) and copied into the blob
  std::vector<CUtensorMap> maps;
  maps.reserve((size_t)n_steps * 5);
  int rc;
  for (int si = 0; si < n_steps; ++si) {
    const long long* t = table + (size_t)si * kStepFields;
    FStep& s = steps[si];
    s.kind = (int)t[0];
    const int cin = (int)t[9], cout = (int)t[10], k = (int)t[11], stride = (int)t[12], hin = (int)t[13], win = (int)t[14];
    s.cin = cin; s.hin = hin; s.win = win;
    if (s.kind == 1) {
      VLNCE_CHECK_ARG(t[1] && t[6] && cin % 8 == 0 && hin % 2 == 0 && win % 2 == 0, "depth_fused_plan: step %d maxpool", si);
      s.pool_in = reinterpret_cast<const __half*>(t[1]);
      s.out = reinterpret_cast<__half*>(t[6]);
      s.hout = hin / 2; s.wout = win / 2; s.cout = cin; s.px = s.hout * s.wout;
      continue;
    }
    VLNCE_CHECK_ARG(s.kind == 0, "depth_fused_plan: step %d kind %d", si, s.kind);
    VLNCE_CHECK_ARG(t[1] && t[2] && t[3] && t[4] && (t[6] || t[7]), "depth_fused_plan: step %d null tensor", si);
    VLNCE_CHECK_ARG((k == 1 || k == 3) && (stride == 1 || stride == 2), "depth_fused_plan: step %d k=%d stride=%d", si, k, stride);
    VLNCE_CHECK_ARG(hin % stride == 0 && win % stride == 0, "depth_fused_plan: step %d odd size under stride 2", si);
    const int hout = hin / stride, wout = win / stride, px = hout * wout;
    const int groups = (int)t[15];
    VLNCE_CHECK_ARG(groups > 0 && cout % groups == 0, "depth_fused_plan: step %d groups", si);
    const int gs = cout / groups;
    VLNCE_CHECK_ARG((gs & (gs - 1)) == 0, "depth_fused_plan: step %d channels per group %d not a power of two", si, gs);
    const int kc = cin % 64 == 0 ? 64 : 32;
    VLNCE_CHECK_ARG(cin % kc == 0 && cout % 16 == 0 && cout <= kMaxCout, "depth_fused_plan: step %d channels %d -> %d", si, cin, cout);
    s.kc = kc; s.taps = k * k; s.cin_chunks = cin / kc; s.stride = stride;
    s.cout = cout; s.hout = hout; s.wout = wout; s.px = px;
    s.groups = groups; s.gs = gs; s.relu = (int)t[16]; s.external_input = (int)t[17];
    {
      const uint32_t bits = (uint32_t)t[18];
      memcpy(&s.eps, &bits, 4);
    }
    s.gamma = reinterpret_cast<const float*>(t[3]);
    s.beta = reinterpret_cast<const float*>(t[4]);
    s.res = reinterpret_cast<const __half*>(t[5]);
    s.out = reinterpret_cast<__half*>(t[6]);
    s.out_f32 = reinterpret_cast<float*>(t[7]);
    s.ld_f32 = (int)t[8];
    const long long ktot = (long long)k * k * cin;
    int box_y;
    if (px >= 128) {
      // tiles = pixel blocks of 128 (whole image rows), slab = filters
      VLNCE_CHECK_ARG(px % 128 == 0 && 128 % wout == 0 && wout <= 256, "depth_fused_plan: step %d image %dx%d", si, hout, wout);
      s.swapped = 0;
      s.n_tiles = px / 128;
      s.rows_per_tile = 128 / wout;
      int slab = cout;
      while (slab > kMaxSlab || s.n_tiles * slab > 512) slab /= 2;
      VLNCE_CHECK_ARG(slab >= 16 && slab % 16 == 0 && cout % slab == 0 && slab % gs == 0 && slab / gs <= kMaxGroups && gs <= 16,
                      "depth_fused_plan: step %d cannot slab %d filters over %d pixel tiles (group size %d)", si, cout, s.n_tiles, gs);
      s.slab_rows = slab;
      s.n_slabs = cout / slab;
      box_y = s.rows_per_tile;
    } else {
      // small image: tiles = 128 filters, slab = the image's pixels
      VLNCE_CHECK_ARG(cout % 128 == 0 && px % 16 == 0 && px * 128 * 4 <= kStageBytes && (cout / 128) * px <= 512,
                      "depth_fused_plan: step %d small image %dx%d with %d filters", si, hout, wout, cout);
      VLNCE_CHECK_ARG(groups <= kMaxGroups && kEpi % groups == 0 && (kEpi / groups <= 32 || groups == 1),
                      "depth_fused_plan: step %d group count %d", si, groups);
      s.swapped = 1;
      s.n_tiles = cout / 128;
      s.rows_per_tile = 0;
      s.slab_rows = px;
      s.n_slabs = 1;
      box_y = hout;
    }
    const int KK = k * k * (cin / kc);
    const int chunk_bytes = s.slab_rows * kc * 2;
    {
      int sper = kSBytes / chunk_bytes;
      if (sper > 8) sper = 8;
      if (sper < 1) sper = 1;
      if (s.swapped) {   // a request covers channel chunks of ONE tap window: sper must divide cin / kc
        while (s.cin_chunks % sper != 0) --sper;
      } else if (sper > KK) {
        sper = KK;
      }
      s.sper = sper;
    }
    s.w_pairs_k = 0;
    if (!s.swapped) {
      s.tper = kc == 32 ? 4 : 2;
      if (s.n_tiles % s.tper != 0) s.tper = 1;
    } else if (s.n_tiles % 2 == 0) {
      s.tper = 2;
    } else if (s.n_tiles == 1 && KK % 2 == 0) {
      s.tper = 2;
      s.w_pairs_k = 1;
    } else {
      s.tper = 1;
    }
    // weights as a 3-D tensor (k within chunk, filter, k-chunk): a box spans filters and / or consecutive k-chunks
    {
      s.map_w = (int)maps.size();
      maps.emplace_back();
      const cuuint64_t gd[3] = {(cuuint64_t)kc, (cuuint64_t)cout, (cuuint64_t)(ktot / kc)};
      const cuuint64_t gs_[2] = {(cuuint64_t)ktot * 2, (cuuint64_t)kc * 2};
      cuuint32_t bx[3] = {(cuuint32_t)kc, 0, 0};
      if (!s.swapped) { bx[1] = (cuuint32_t)s.slab_rows; bx[2] = (cuuint32_t)s.sper; }
      else if (s.w_pairs_k) { bx[1] = 128; bx[2] = 2; }
      else { bx[1] = (cuuint32_t)(128 * s.tper); bx[2] = 1; }
      if ((rc = encode(enc, &maps.back(), reinterpret_cast<const void*>(t[2]), 3, gd, gs_, bx, kc, "weights")) != VLNCE_OK) return rc;
    }
    // activation views: stride 1 -> one map; stride 2 -> four parity views (py, px).  Tiled operand (large images):
    // 4-D (cin, w, h, N), a box = tper pixel tiles of one 64-channel chunk.  Slab operand (small images): 5-D
    // (k within chunk, w, h, N, chunk), a box = the image window for sper channel chunks.
    const uint8_t* xb = reinterpret_cast<const uint8_t*>(t[1]);
    const int nviews = stride == 1 ? 1 : (k == 1 ? 1 : 4);
    for (int v = 0; v < 4; ++v) s.map_x[v] = -1;
    for (int v = 0; v < nviews; ++v) {
      const int py = v >> 1, pxo = v & 1;
      s.map_x[v] = (int)maps.size();
      maps.emplace_back();
      const void* vb = xb + ((size_t)py * win + pxo) * cin * 2;
      if (!s.swapped) {
        const cuuint64_t gd[4] = {(cuuint64_t)cin, (cuuint64_t)wout, (cuuint64_t)hout, (cuuint64_t)n_images};
        const cuuint64_t gstr[3] = {(cuuint64_t)stride * cin * 2, (cuuint64_t)stride * win * cin * 2,
                                    (cuuint64_t)hin * win * cin * 2};
        const cuuint32_t bx[4] = {(cuuint32_t)kc, (cuuint32_t)wout, (cuuint32_t)(box_y * s.tper), 1};
        if ((rc = encode(enc, &maps.back(), vb, 4, gd, gstr, bx, kc, "activation tiles")) != VLNCE_OK) return rc;
      } else {
        const cuuint64_t gd[5] = {(cuuint64_t)kc, (cuuint64_t)wout, (cuuint64_t)hout, (cuuint64_t)n_images,
                                  (cuuint64_t)(cin / kc)};
        const cuuint64_t gstr[4] = {(cuuint64_t)stride * cin * 2, (cuuint64_t)stride * win * cin * 2,
                                    (cuuint64_t)hin * win * cin * 2, (cuuint64_t)kc * 2};
        const cuuint32_t bx[5] = {(cuuint32_t)kc, (cuuint32_t)wout, (cuuint32_t)hout, 1, (cuuint32_t)s.sper};
        if ((rc = encode(enc, &maps.back(), vb, 5, gd, gstr, bx, kc, "activation window")) != VLNCE_OK) return rc;
      }
    }
    for (int v = nviews; v < 4; ++v) s.map_x[v] = s.map_x[0];
  }
  hdr->n_maps = (int)maps.size();
  hdr->total_bytes = hdr->maps_off + (long long)maps.size() * 128;
  VLNCE_CHECK_ARG(hdr->total_bytes <= plan_bytes, "depth_fused_plan: plan overflow");
  memcpy(base + hdr->maps_off, maps.data(), maps.size() * sizeof(CUtensorMap));
  return VLNCE_OK;
}

extern "C" int vlnce_depth_fused_run(const void* plan_dev, int n_images, int ctas, cudaStream_t stream) {
  VLNCE_CHECK_ARG(plan_dev && n_images > 0, "depth_fused_run: null plan / no images");
  VLNCE_CHECK_ARG((reinterpret_cast<uintptr_t>(plan_dev) & 127) == 0, "depth_fused_run: plan must be 128-byte aligned");
  static bool configured = false;
  static int num_sms = 0;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(depth_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(depth_fused)");
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
    configured = true;
  }
  int grid = ctas > 0 ? ctas : num_sms;
  if (grid > n_images) grid = n_images;
  if (grid > num_sms) grid = num_sms;
  depth_fused_kernel<<<grid, kThreads, kSmemBytes, stream>>>(reinterpret_cast<const uint8_t*>(plan_dev), n_images,
                                                             g_fused_trace);
  VLNCE_LAUNCH_CHECK("depth_fused_kernel");
  return VLNCE_OK;
}

// Debug: device buffer of n_steps * 2 * 8 int64 that CTA 0 fills with %globaltimer stamps per (step, slab) for its
// first image (see kTraceSlots); NULL switches tracing off.
extern "C" int vlnce_depth_fused_set_trace(void* buffer) {
  g_fused_trace = reinterpret_cast<long long*>(buffer);
  return VLNCE_OK;
}
