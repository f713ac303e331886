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
// warp 3 = TMEM owner, warps 4-11 = epilogue (two warps per TMEM lane quarter).  The weight producer
// runs ahead across step boundaries; activation loads wait for the step that produced them.
#include "common.cuh"
#include "ptx_sm100.cuh"

#include <stdlib.h>
#include <vector>

namespace vlnce {
namespace {

using namespace ptx;

constexpr int kT = 6;                  // tiled-operand ring slots (128 rows x 128 B)
constexpr int kS = 2;                  // slab-operand ring slots (<= 256 rows x 128 B)
constexpr int kTBytes = 16 * 1024;
constexpr int kSBytes = 32 * 1024;
constexpr int kEpiWarps = 8;
constexpr int kEpi = kEpiWarps * 32;   // 256 epilogue threads
constexpr int kThreads = 4 * 32 + kEpi;
constexpr int kMaxGroups = 16;         // groups per pass (non-swapped: per slab; swapped: per layer)
constexpr int kMaxCout = 1024;
constexpr int kAccBytes = kMaxGroups * kEpi * 8;  // float2 acc[group][thread]  (aliased by chs[cout][2] float2)
constexpr int kTabBytes = 256 * 8;                // float2 (scale, shift) per slab column
constexpr int kSmemBytes = kT * kTBytes + kS * kSBytes + kAccBytes + kTabBytes + 1024 /*stats+barriers*/ + 1024 /*align*/;

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
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
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
    if ((i & 255u) == 0) {
      __nanosleep(64);
      if (global_timer_ns() - t0 > 2000000000ull) break;
    }
  }
  printf("vlnce: depth_fused step wait timed out (tag %d, block %d, want %u)\n", tag, blockIdx.x, want);
  __trap();
}

// tap -> (parity view, x offset, y offset) of the activation window
__device__ __forceinline__ void tap_window(const FStep& s, int tap, int& mi, int& xo, int& yo) {
  mi = 0; xo = 0; yo = 0;
  if (s.taps == 1) return;
  const int r = tap / 3, q = tap - r * 3;
  if (s.stride == 1) { xo = q - 1; yo = r - 1; return; }
  // stride 2: input row 2y + r - 1 = parity (r != 1), row index y + (r == 0 ? -1 : 0) of that view
  mi = (r != 1 ? 2 : 0) + (q != 1 ? 1 : 0);
  yo = r == 0 ? -1 : 0;
  xo = q == 0 ? -1 : 0;
}

__global__ void __launch_bounds__(kThreads, 1)
depth_fused_kernel(const uint8_t* __restrict__ plan, int n_images) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const uint32_t t_ring = smem_u32(smem);
  const uint32_t s_ring = t_ring + kT * kTBytes;
  uint8_t* scratch = smem + kT * kTBytes + kS * kSBytes;
  float2* acc = reinterpret_cast<float2*>(scratch);                 // [group][thread] | chs[channel][half]
  float2* tab = reinterpret_cast<float2*>(scratch + kAccBytes);     // (scale, shift) per slab column
  uint8_t* tail = scratch + kAccBytes + kTabBytes;
  float2* stat = reinterpret_cast<float2*>(tail);                   // (mean, rstd) per group   [16]
  double* red = reinterpret_cast<double*>(tail + 128);              // cross-warp scratch        [16]
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 256);
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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kT; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 1); }
    for (int i = 0; i < kS; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 1); }
    mbar_init(acc_full, 1);
    mbar_init(tmem_free, kEpi);
    *steps_done = 0;
    fence_barrier_init();
  }
  if (warp == 3) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ tiled-operand producer
    if (lane == 0) {
      int ts = 0;
      uint32_t tph = 0, gstep = 0;
      for (int img = blockIdx.x; img < n_images; img += gridDim.x) {
        for (int si = 0; si < n_steps; ++si, ++gstep) {
          const FStep& s = steps[si];
          if (s.kind != 0) continue;
          const bool is_x = !s.swapped;
          if (is_x && !s.external_input) wait_steps(steps_done, gstep, 1);
          const CUtensorMap* mw = maps + s.map_w;
          const uint32_t bytes = 128u * (uint32_t)s.kc * 2u;
          for (int slab = 0; slab < s.n_slabs; ++slab) {
            for (int tap = 0; tap < s.taps; ++tap) {
              int mi, xo, yo;
              tap_window(s, tap, mi, xo, yo);
              const CUtensorMap* mx = maps + s.map_x[mi];
              for (int cc = 0; cc < s.cin_chunks; ++cc) {
                for (int i = 0; i < s.n_tiles; ++i) {
                  mbar_wait(&t_empty[ts], tph ^ 1, 40);
                  mbar_arrive_expect_tx(&t_full[ts], bytes);
                  const uint32_t dst = t_ring + ts * kTBytes;
                  if (is_x) tma_load_4d(dst, mx, &t_full[ts], cc * s.kc, xo, i * s.rows_per_tile + yo, img);
                  else tma_load_2d(dst, mw, &t_full[ts], tap * s.cin + cc * s.kc, i * 128);
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
    if (lane == 0) {
      int ss = 0;
      uint32_t sph = 0, gstep = 0;
      for (int img = blockIdx.x; img < n_images; img += gridDim.x) {
        for (int si = 0; si < n_steps; ++si, ++gstep) {
          const FStep& s = steps[si];
          if (s.kind != 0) continue;
          const bool is_x = s.swapped != 0;
          if (is_x && !s.external_input) wait_steps(steps_done, gstep, 2);
          const CUtensorMap* mw = maps + s.map_w;
          const uint32_t bytes = (uint32_t)s.slab_rows * (uint32_t)s.kc * 2u;
          for (int slab = 0; slab < s.n_slabs; ++slab) {
            for (int tap = 0; tap < s.taps; ++tap) {
              int mi, xo, yo;
              tap_window(s, tap, mi, xo, yo);
              const CUtensorMap* mx = maps + s.map_x[mi];
              for (int cc = 0; cc < s.cin_chunks; ++cc) {
                mbar_wait(&s_empty[ss], sph ^ 1, 41);
                mbar_arrive_expect_tx(&s_full[ss], bytes);
                const uint32_t dst = s_ring + ss * kSBytes;
                if (is_x) tma_load_4d(dst, mx, &s_full[ss], cc * s.kc, xo, yo, img);
                else tma_load_2d(dst, mw, &s_full[ss], tap * s.cin + cc * s.kc, slab * s.slab_rows);
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
          const FStep& s = steps[si];
          if (s.kind != 0) continue;
          const uint32_t idesc = make_idesc_f16(128, (uint32_t)s.slab_rows, 0u);
          const uint32_t row_bytes = (uint32_t)s.kc * 2u;
          const int ksteps = s.kc / 16;
          for (int slab = 0; slab < s.n_slabs; ++slab, ++pass) {
            mbar_wait(tmem_free, (pass & 1) ^ 1, 42);  // epilogue has drained the previous pass
            tc_fence_after_sync();
            uint32_t first = 0;
            for (int kk = 0; kk < s.taps * s.cin_chunks; ++kk) {
              mbar_wait(&s_full[ss], sph, 43);
              tc_fence_after_sync();
              const uint64_t bdesc = make_smem_desc(s_ring + ss * kSBytes, row_bytes);
              for (int i = 0; i < s.n_tiles; ++i) {
                mbar_wait(&t_full[ts], tph, 44);
                tc_fence_after_sync();
                const uint64_t adesc = make_smem_desc(t_ring + ts * kTBytes, row_bytes);
                const uint32_t d = tmem_base + (uint32_t)(i * s.slab_rows);
                for (int k = 0; k < ksteps; ++k)
                  umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc, (first | (uint32_t)k) != 0 ? 1u : 0u);
                umma_commit(&t_empty[ts]);
                if (++ts == kT) { ts = 0; tph ^= 1; }
              }
              first = 1;
              umma_commit(&s_empty[ss]);
              if (++ss == kS) { ss = 0; sph ^= 1; }
            }
            umma_commit(acc_full);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (warps 4-11)
    const int ew = warp - 4;
    const int q = warp & 3;          // TMEM lane quarter this warp may read
    const int half = ew >> 2;        // two warps per quarter split the work items
    const int r = q * 32 + lane;     // accumulator row (pixel in tile | filter in tile)
    const int tid = threadIdx.x - 128;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t pass = 0, gstep = 0;
    for (int img = blockIdx.x; img < n_images; img += gridDim.x) {
      for (int si = 0; si < n_steps; ++si, ++gstep) {
        const FStep& s = steps[si];
        // all epilogue threads have finished writing the previous step (residual / pool inputs
        // are read below by other threads than the ones that wrote them)
        named_barrier_sync(1, kEpi);
        if (s.kind == 1) {
          // ---- MaxPool2d(3, 2, 1) on the 16-bit NHWC image
          const int C8 = s.cin / 8, P = s.hout, Q = s.wout;
          const __half* in = s.pool_in + (size_t)img * s.hin * s.win * s.cin;
          __half* out = s.out + (size_t)img * P * Q * s.cin;
          for (int idx = tid; idx < P * Q * C8; idx += kEpi) {
            const int c8 = idx % C8;
            const int m = idx / C8;
            const int qo = m % Q, po = m / Q;
            __half2 best[4];
            bool any = false;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const int ih = po * 2 + dy - 1;
              if ((unsigned)ih >= (unsigned)s.hin) continue;
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                const int iw = qo * 2 + dx - 1;
                if ((unsigned)iw >= (unsigned)s.win) continue;
                const uint4 v = *reinterpret_cast<const uint4*>(in + ((size_t)ih * s.win + iw) * s.cin + c8 * 8);
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
            *reinterpret_cast<uint4*>(out + (size_t)m * s.cin + c8 * 8) = *reinterpret_cast<const uint4*>(best);
          }
        } else {
          const int cpt = s.slab_rows / 16;           // 16-column chunks per tile
          const int items = s.n_tiles * cpt;
          for (int slab = 0; slab < s.n_slabs; ++slab, ++pass) {
            if (!s.swapped) {
              // ================= tiles = pixels, columns = filters of this slab
              const int ngs = s.slab_rows / s.gs;
              for (int g = 0; g < ngs; ++g) acc[g * kEpi + tid] = make_float2(0.f, 0.f);
              mbar_wait(acc_full, pass & 1, 45);
              tc_fence_after_sync();
              // ---- pass 1: per-thread partial sums of every group
              for (int j = half; j < items; j += 2) {
                const int i = j / cpt, c0 = (j - i * cpt) * 16;
                uint32_t v[16];
                tmem_ld_32x16(lane_addr + (uint32_t)(i * s.slab_rows + c0), v);
                tmem_ld_wait();
                float sum = 0.f, sq = 0.f;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                  const float f = __uint_as_float(v[e]);
                  sum += f;
                  sq = fmaf(f, f, sq);
                  if (((e + 1) & (s.gs - 1)) == 0) {   // gs is a power of two <= 16
                    float2* a = &acc[((c0 + e) / s.gs) * kEpi + tid];
                    float2 t = *a;
                    t.x += sum; t.y += sq;
                    *a = t;
                    sum = 0.f; sq = 0.f;
                  }
                }
              }
              named_barrier_sync(1, kEpi);
              // ---- group statistics: fp64 combine of the 256 per-thread partials
              for (int g = ew; g < ngs; g += kEpiWarps) {
                double ds = 0.0, dq = 0.0;
#pragma unroll
                for (int m = 0; m < kEpi / 32; ++m) {
                  const float2 t = acc[g * kEpi + m * 32 + lane];
                  ds += (double)t.x;
                  dq += (double)t.y;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                  ds += __shfl_xor_sync(0xffffffffu, ds, o);
                  dq += __shfl_xor_sync(0xffffffffu, dq, o);
                }
                if (lane == 0) {
                  const double cnt = (double)s.px * (double)s.gs;
                  const double mean = ds / cnt;
                  double var = dq / cnt - mean * mean;
                  if (var < 0.0) var = 0.0;
                  stat[g] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)s.eps)));
                }
              }
              named_barrier_sync(1, kEpi);
              for (int t = tid; t < s.slab_rows; t += kEpi) {
                const int c = slab * s.slab_rows + t;
                const float2 st = stat[t / s.gs];
                const float sc = st.y * __ldg(s.gamma + c);
                tab[t] = make_float2(sc, __ldg(s.beta + c) - st.x * sc);
              }
              named_barrier_sync(1, kEpi);
              // ---- pass 2: normalise (+ residual)(+ ReLU), 16-bit store
              for (int j = half; j < items; j += 2) {
                const int i = j / cpt, c0 = (j - i * cpt) * 16;
                const size_t o = ((size_t)img * s.px + (size_t)(i * 128 + r)) * s.cout + slab * s.slab_rows + c0;
                uint4 rv[2];
                if (s.res != nullptr) {
                  rv[0] = *reinterpret_cast<const uint4*>(s.res + o);
                  rv[1] = *reinterpret_cast<const uint4*>(s.res + o + 8);
                }
                uint32_t v[16];
                tmem_ld_32x16(lane_addr + (uint32_t)(i * s.slab_rows + c0), v);
                tmem_ld_wait();
                float f[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                  const float2 t = tab[c0 + e];
                  f[e] = fmaf(__uint_as_float(v[e]), t.x, t.y);
                }
                if (s.res != nullptr) {
                  const __half2* rh = reinterpret_cast<const __half2*>(rv);
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    const float2 t = __half22float2(rh[e]);
                    f[2 * e] += t.x;
                    f[2 * e + 1] += t.y;
                  }
                }
                if (s.relu) {
#pragma unroll
                  for (int e = 0; e < 16; ++e) f[e] = fmaxf(f[e], 0.f);
                }
                if (s.out != nullptr) {
                  __align__(16) __half2 hv[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) hv[e] = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
                  *reinterpret_cast<uint4*>(s.out + o) = *reinterpret_cast<const uint4*>(&hv[0]);
                  *reinterpret_cast<uint4*>(s.out + o + 8) = *reinterpret_cast<const uint4*>(&hv[4]);
                }
                if (s.out_f32 != nullptr) {
                  float* of = s.out_f32 + ((size_t)img * s.px + (size_t)(i * 128 + r)) * s.ld_f32 + slab * s.slab_rows + c0;
#pragma unroll
                  for (int e = 0; e < 16; ++e) of[e] = f[e];
                }
              }
            } else {
              // ================= tiles = filters, columns = the image's pixels
              float2* chs = acc;  // [channel][half] partial (sum, sumsq)
              for (int t = tid; t < s.cout * 2; t += kEpi) chs[t] = make_float2(0.f, 0.f);
              named_barrier_sync(1, kEpi);
              mbar_wait(acc_full, pass & 1, 46);
              tc_fence_after_sync();
              for (int j = half; j < items; j += 2) {
                const int i = j / cpt, c0 = (j - i * cpt) * 16;
                uint32_t v[16];
                tmem_ld_32x16(lane_addr + (uint32_t)(i * s.slab_rows + c0), v);
                tmem_ld_wait();
                float sum = 0.f, sq = 0.f;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                  const float f = __uint_as_float(v[e]);
                  sum += f;
                  sq = fmaf(f, f, sq);
                }
                float2* a = &chs[(i * 128 + r) * 2 + half];
                float2 t = *a;
                t.x += sum; t.y += sq;
                *a = t;
              }
              named_barrier_sync(1, kEpi);
              {
                const int G = s.groups;
                const int tpg = kEpi / G;          // threads per group (host: 256 % G == 0, tpg <= 32 or G == 1)
                const int g = tid / tpg, sub = tid - g * tpg;
                double ds = 0.0, dq = 0.0;
                for (int e = sub; e < 2 * s.gs; e += tpg) {
                  const float2 t = chs[g * s.gs * 2 + e];
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
                  const double cnt = (double)s.px * (double)s.gs;
                  const double mean = ds / cnt;
                  double var = dq / cnt - mean * mean;
                  if (var < 0.0) var = 0.0;
                  stat[g] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)s.eps)));
                }
              }
              named_barrier_sync(1, kEpi);
              int cur_tile = -1;
              float sc = 0.f, sh = 0.f;
              for (int j = half; j < items; j += 2) {
                const int i = j / cpt, c0 = (j - i * cpt) * 16;
                const int c = i * 128 + r;
                if (i != cur_tile) {
                  cur_tile = i;
                  const float2 st = stat[c / s.gs];
                  sc = st.y * __ldg(s.gamma + c);
                  sh = __ldg(s.beta + c) - st.x * sc;
                }
                const size_t o = ((size_t)img * s.px + c0) * s.cout + c;
                float rf[16];
                if (s.res != nullptr) {
#pragma unroll
                  for (int e = 0; e < 16; ++e) rf[e] = __half2float(s.res[o + (size_t)e * s.cout]);
                }
                uint32_t v[16];
                tmem_ld_32x16(lane_addr + (uint32_t)(i * s.slab_rows + c0), v);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                  float f = fmaf(__uint_as_float(v[e]), sc, sh);
                  if (s.res != nullptr) f += rf[e];
                  if (s.relu) f = fmaxf(f, 0.f);
                  if (s.out != nullptr) s.out[o + (size_t)e * s.cout] = __float2half_rn(f);
                  if (s.out_f32 != nullptr) s.out_f32[((size_t)img * s.px + c0 + e) * s.ld_f32 + c] = f;
                }
              }
            }
            // this thread is done with the accumulator: hand TMEM back to the MMA issuer
            tc_fence_before_sync();
            mbar_arrive(tmem_free);
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
  if (warp == 3) tmem_dealloc(tmem_base, 512);
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
  const cuuint32_t estr[4] = {1, 1, 1, 1};
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
    int box_rows_w, box_y;
    if (px >= 128) {
      // tiles = pixel blocks of 128 (whole image rows), slab = filters
      VLNCE_CHECK_ARG(px % 128 == 0 && 128 % wout == 0 && wout <= 256, "depth_fused_plan: step %d image %dx%d", si, hout, wout);
      s.swapped = 0;
      s.n_tiles = px / 128;
      s.rows_per_tile = 128 / wout;
      int slab = cout;
      while (slab > 256 || s.n_tiles * slab > 512) slab /= 2;
      VLNCE_CHECK_ARG(slab >= 16 && slab % 16 == 0 && cout % slab == 0 && slab % gs == 0 && slab / gs <= kMaxGroups && gs <= 16,
                      "depth_fused_plan: step %d cannot slab %d filters over %d pixel tiles (group size %d)", si, cout, s.n_tiles, gs);
      s.slab_rows = slab;
      s.n_slabs = cout / slab;
      box_rows_w = slab;
      box_y = s.rows_per_tile;
    } else {
      // small image: tiles = 128 filters, slab = the image's pixels
      VLNCE_CHECK_ARG(cout % 128 == 0 && px % 16 == 0 && (cout / 128) * px <= 512,
                      "depth_fused_plan: step %d small image %dx%d with %d filters", si, hout, wout, cout);
      VLNCE_CHECK_ARG(groups <= kMaxGroups && kEpi % groups == 0 && (kEpi / groups <= 32 || groups == 1),
                      "depth_fused_plan: step %d group count %d", si, groups);
      s.swapped = 1;
      s.n_tiles = cout / 128;
      s.rows_per_tile = 0;
      s.slab_rows = px;
      s.n_slabs = 1;
      box_rows_w = 128;
      box_y = hout;
    }
    // weights [cout, k*k*cin]
    {
      s.map_w = (int)maps.size();
      maps.emplace_back();
      const cuuint64_t gd[2] = {(cuuint64_t)ktot, (cuuint64_t)cout};
      const cuuint64_t gs_[1] = {(cuuint64_t)ktot * 2};
      const cuuint32_t bx[2] = {(cuuint32_t)kc, (cuuint32_t)box_rows_w};
      if ((rc = encode(enc, &maps.back(), reinterpret_cast<const void*>(t[2]), 2, gd, gs_, bx, kc, "weights")) != VLNCE_OK) return rc;
    }
    // activation views: stride 1 -> one map over (cin, win, hin, N); stride 2 -> four parity views (py, px)
    const uint8_t* xb = reinterpret_cast<const uint8_t*>(t[1]);
    const int nviews = stride == 1 ? 1 : (k == 1 ? 1 : 4);
    for (int v = 0; v < 4; ++v) s.map_x[v] = -1;
    for (int v = 0; v < nviews; ++v) {
      const int py = v >> 1, pxo = v & 1;
      s.map_x[v] = (int)maps.size();
      maps.emplace_back();
      const cuuint64_t gd[4] = {(cuuint64_t)cin, (cuuint64_t)wout, (cuuint64_t)hout, (cuuint64_t)n_images};
      const cuuint64_t gstr[3] = {(cuuint64_t)stride * cin * 2, (cuuint64_t)stride * win * cin * 2,
                                  (cuuint64_t)hin * win * cin * 2};
      const cuuint32_t bx[4] = {(cuuint32_t)kc, (cuuint32_t)wout, (cuuint32_t)box_y, 1};
      const void* vb = xb + ((size_t)py * win + pxo) * cin * 2;
      if ((rc = encode(enc, &maps.back(), vb, 4, gd, gstr, bx, kc, "activation view")) != VLNCE_OK) return rc;
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
  depth_fused_kernel<<<grid, kThreads, kSmemBytes, stream>>>(reinterpret_cast<const uint8_t*>(plan_dev), n_images);
  VLNCE_LAUNCH_CHECK("depth_fused_kernel");
  return VLNCE_OK;
}
