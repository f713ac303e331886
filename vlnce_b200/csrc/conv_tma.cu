// Persistent, fully TMA-fed convolution tiles on tcgen05 — the main conv engine for every
// stride-1 layer of the two ResNet trunks:
//
//   mode ROWS  1x1 stride-1 conv: the NHWC activation is a row-major [M, C] matrix; a tile is
//              128 consecutive pixels x BLOCK_N filters, A tiles are 2-D TMA boxes.
//   mode HALO  3x3 stride-1 pad-1 conv: a tile is a 16 x 8 pixel patch; per 64-channel chunk ONE
//              4-D TMA box loads the (16+2) x (8+2) halo (image borders arrive as zeros) and the
//              nine taps are nine shifted UMMA descriptors into it (see conv3x3_halo.cu).
//
//   mode STEM  the 7x7/2 RGB stem as a stride-1 4x4 convolution over the 2x2 space-to-depth image
//              (12 real + 4 zero channels per cell, see stem_s2d_kernel): a tile is 128 consecutive
//              output pixels of one output row; K chunk r' (one of four cell rows) is ONE 4-D TMA box
//              whose pixel stride (one cell = 32 B) is smaller than its 128-byte extent, i.e. the
//              overlapping 4-cell windows are read straight from the s2d image — no im2col matrix.
//
// L2 -> SM bandwidth is what bounds these tiles (B200: ~42 B/clk/SM against 64 B/clk of weight
// traffic for a 128-row tile), so a scheduled tile may hold MT = 2 pixel sub-tiles that share every
// weight stage (two accumulators, weight bytes per MAC halved), and when one CTA's whole weight
// slice fits in shared memory next to the activation ring it is loaded ONCE ("stationary") and
// only activations stream afterwards.
//
// Persistent CTAs (one per SM) walk a static round-robin tile schedule.  Three roles run
// concurrently and never wait for each other's whole tile:
//
//   warp 9     activation producer (one thread): TMA ring requested across tile boundaries.
//   warp 10    weight producer (one thread): weight ring (or the stationary slice, once).
//   warp 11    store issuer (one thread): TMA store of every staged sub-tile; the residual of a later sub-tile is
//              prefetched into a staging buffer as soon as its last store has been read.
//   warp 8     MMA issuer (one thread): accumulates tile i into TMEM buffer i&1.
//   warps 0-7  epilogue of tile i-1 WHILE tile i is being accumulated: tcgen05.ld, scale/shift, residual, ReLU,
//              swizzled staging, then ONE mbarrier arrive (no CTA-wide barrier, no store issue on this path).
//
// Barrier protocol (all mbarriers, phases tracked per role): a_full/a_empty, b_full/b_empty rings (empty:
// tcgen05.commit); acc_full[2] (issuer -> epilogue), acc_empty[2] (epilogue -> issuer, 256 arrivals after
// the last tcgen05.ld of the tile); staged[s] (256 arrivals: sub-tile written to staging s); res_full[s] (residual
// landed in staging s); stg_free[s] (the TMA store that last read staging s has finished reading it).
#include "common.cuh"
#include "ptx_sm100.cuh"

#include <stdlib.h>

namespace vlnce {
namespace {

using namespace ptx;

constexpr int TH = 16, TW = 8;
constexpr int kEpiWarps = 8;                       // two groups of four: each TMEM lane quarter is read by 2 warps
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = kEpiThreads + 4 * 32;     // + MMA issuer, activation producer, weight producer, store issuer
constexpr int kMaxStg = 4;
constexpr int kSchedRing = 16;                     // tile ids in flight between the claiming thread and the slowest role
constexpr int kSmemBudget = 227 * 1024;
constexpr int kMaxA = 8, kMaxB = 18;

struct TmaConvParams {
  const float* scale;
  const float* shift;
  int mode;        // 0 rows (1x1), 1 halo (3x3), 2 stem (s2d 4x4, chunk = cell row), 3 strided 1x1 (stride 2: the TMA box
                   //   walks the input with element strides {1,2,2,1}, so only the sampled pixels are ever loaded)
  int img_per_tile;  // mode 3: images per 128-pixel tile (small outputs: 8x8 -> 2 images)
  int cpt;           // mode 3: 64-channel chunks per filter tap (K chunk q = tap * cpt + channel chunk); taps_s2 = 1 | 9
  int taps_s2;
  int box_stride;    // mode 3: 2 (stride-2 convs) | 1 (3x3 stride 1 on images too small for the halo tiles)
  float rcp_cpt;
  int M;           // rows mode: N*H*W
  int N, H, W, C, K;
  int c_chunks, n_tiles, m_tiles, tiles_x, tiles_y;
  float rcp_n_tiles, rcp_tiles_x, rcp_tiles_y;  // float reciprocals for the division-free tile decode
  int tw, th;      // output patch of one sub-tile (halo 8x16, stem 128x1)
  int w_cols;      // K extent of one filter tap in the weight matrix (C; stem: 0, chunks are consecutive)
  int b_stationary;  // 1: b_stages == taps*c_chunks and the weight slice is loaded once per CTA
  long long* trace;  // debug: per-role (event, clock64) records of CTA 0 (vlnce_debug_set_trace), else NULL
  int* sched;        // dynamic tile scheduler: {next tile, finished CTAs} of this launch (both zero before and after)
  int taps;        // 1 | 9
  int halo_w;      // 8 (rows mode: plain 8-row groups) | 10
  int a_bytes;     // ring slot size (1024-multiple), a_tx = bytes the TMA delivers
  int a_tx;
  int a_stages, b_stages, stg_bufs;
  int relu, y_fp32, ab_half, has_res;
  int tf32;        // 1: fp32 operands multiplied as TF32 (kind::tf32); a K chunk is then 32 channels (128 bytes)
  int kc;          // channels per K chunk: 64 (16-bit operands) | 32 (tf32)
};

__device__ __forceinline__ void tma_store_4d_(const void* tmap, uint32_t src_smem, int32_t c0, int32_t c1, int32_t c2,
                                              int32_t c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   (uint64_t)tmap),
               "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read_n() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint64_t desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Debug trace (CTA 0 only): role 0 activation producer / 1 MMA issuer / 2 epilogue thread 0 / 3 weight
// producer / 4 store issuer; up to kTraceCap
// records per role of {event << 32 | index, clock64}.
constexpr int kTraceCap = 768;
struct Trace {
  long long* base;
  int n;
  __device__ __forceinline__ Trace(long long* t, int role) : base(t ? t + (size_t)role * kTraceCap * 2 : nullptr), n(0) {
    if (blockIdx.x != 0) base = nullptr;
  }
  __device__ __forceinline__ void rec(int ev, int idx) {
    if (base != nullptr && n < kTraceCap) {
      base[2 * n] = ((long long)ev << 32) | (unsigned)idx;
      base[2 * n + 1] = clock64();
      ++n;
    }
  }
};

struct Tile {
  int n0;            // first filter
  int m0;            // rows mode: first pixel row
  int x0, y0, img;   // halo mode: patch origin
};

// q = a / d, r = a % d for 0 <= a < 2^24 without the ~100-cycle integer-division sequence (the
// producer / epilogue roles are single threads: their instruction latency IS the pipeline's period)
__device__ __forceinline__ int fast_divmod(int a, int d, float rcp, int& r) {
  int q = __float2int_rz(__int2float_rn(a) * rcp);
  r = a - q * d;
  if (r < 0) { --q; r += d; }
  else if (r >= d) { ++q; r -= d; }
  return q;
}

// scheduled tile t = (pixel super-tile, filter tile); sub-tile h of MT
__device__ __forceinline__ Tile decode_tile(const TmaConvParams& p, int t, int block_n, int mt, int h) {
  Tile ti;
  int n_tile, xi, yi;
  int m = fast_divmod(t, p.n_tiles, p.rcp_n_tiles, n_tile) * mt + h;
  ti.n0 = n_tile * block_n;
  ti.m0 = m * 128;
  m = fast_divmod(m, p.tiles_x, p.rcp_tiles_x, xi);
  ti.x0 = xi * p.tw;
  ti.img = fast_divmod(m, p.tiles_y, p.rcp_tiles_y, yi) * p.img_per_tile;
  ti.y0 = yi * p.th;
  return ti;
}

// Look-ahead barrier tests of the MMA issuer: mbarrier.test_wait writes a predicate that is only READ after the
// current stage's MMAs have been issued, so the ~150-cycle query overlaps their issue (an in-order thread stalls at
// the first reader of the predicate, not at the test).  The predicates live in function-scope PTX registers
// (declared once per kernel) because a predicate cannot leave an inline-asm block any other way.
#define LOOKAHEAD_DECLARE() asm volatile(".reg .pred vlnce_pa, vlnce_pb;")
__device__ __forceinline__ void lookahead_test_a(uint64_t* bar, uint32_t parity, bool enable) {
  asm volatile("setp.ne.u32 vlnce_pa, 0, 0;" ::: "memory");   // "not seen" unless the test below runs and succeeds
  if (enable) asm volatile("mbarrier.test_wait.parity.shared::cta.b64 vlnce_pa, [%0], %1;" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void lookahead_test_b(uint64_t* bar, uint32_t parity, bool enable) {
  asm volatile("setp.ne.u32 vlnce_pb, 0, 0;" ::: "memory");
  if (enable) asm volatile("mbarrier.test_wait.parity.shared::cta.b64 vlnce_pb, [%0], %1;" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool lookahead_read_a() {
  uint32_t ok;
  asm volatile("selp.u32 %0, 1, 0, vlnce_pa;" : "=r"(ok)::"memory");
  return ok != 0;
}
__device__ __forceinline__ bool lookahead_read_b() {
  uint32_t ok;
  asm volatile("selp.u32 %0, 1, 0, vlnce_pb;" : "=r"(ok)::"memory");
  return ok != 0;
}

template <int BLOCK_N, int MT>
__global__ void __launch_bounds__(kThreads, 1)
conv_tma_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                const __grid_constant__ CUtensorMap tmap_y, const __grid_constant__ CUtensorMap tmap_res,
                const TmaConvParams p) {
  static_assert(2 * MT * BLOCK_N <= 512, "two accumulator sets must fit the 512 TMEM columns");
  constexpr int B_BYTES = BLOCK_N * 128;
  constexpr int NI = BLOCK_N / 64;  // 32-column epilogue iterations per warp (each warp owns half the columns)
  LOOKAHEAD_DECLARE();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const uint32_t a_slot = (uint32_t)(MT * p.a_bytes);  // one ring slot = the MT sub-tiles of one K chunk
  const uint32_t smem_a0 = smem_u32(smem);
  const uint32_t smem_b0 = smem_a0 + p.a_stages * a_slot;
  uint8_t* bar_base = smem + p.a_stages * a_slot + p.b_stages * B_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* a_empty = a_full + kMaxA;
  uint64_t* b_full = a_empty + kMaxA;
  uint64_t* b_empty = b_full + kMaxB;
  uint64_t* acc_full = b_empty + kMaxB;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* res_full = acc_empty + 2;
  uint64_t* stg_free = res_full + kMaxStg;
  uint64_t* staged = stg_free + kMaxStg;
  uint64_t* sched_full = staged + kMaxStg;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sched_full + kSchedRing);
  volatile int* sched_tile = reinterpret_cast<volatile int*>(tmem_slot + 1);   // kSchedRing tile ids (-1 = no more tiles)
  int* sched_done = const_cast<int*>(sched_tile) + kSchedRing;                 // tiles the store issuer has finished
                                                                               // (one writer, one reader, atomics)
  const uint32_t stg_bytes = 128u * BLOCK_N * (p.y_fp32 ? 4u : 2u);
  const uint32_t stage_out = smem_b0 + p.b_stages * B_BYTES + 1024;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = (p.m_tiles / MT) * p.n_tiles;  // host guarantees m_tiles % MT == 0
  // Dynamic tile schedule: the activation producer claims tile ids from a global counter and publishes them to the
  // other roles through a small ring (sched_tile / sched_full); position `it` of a role's walk is ring slot it & 15,
  // phase (it >> 4) & 1.  A CTA that starts late (its SM was busy with another kernel of the step) simply claims
  // fewer tiles — with a static round-robin list the whole launch waited for it.
  auto tile_of = [&](int it) -> int {
    mbar_wait(&sched_full[it & (kSchedRing - 1)], (uint32_t)(it >> 4) & 1u, 40);
    return sched_tile[it & (kSchedRing - 1)];
  };

  if (threadIdx.x == (kEpiWarps + 1) * 32) {
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], kEpiThreads);
    }
    for (int s = 0; s < kMaxStg; ++s) {
      mbar_init(&res_full[s], 1);
      mbar_init(&stg_free[s], 1);
      mbar_init(&staged[s], kEpiThreads);
    }
    for (int s = 0; s < kSchedRing; ++s) mbar_init(&sched_full[s], 1);
    *sched_done = 0;
    fence_barrier_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_y);
  }
  if (warp == kEpiWarps) {
    tmem_alloc(tmem_slot, 2 * MT * BLOCK_N);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: this CTA owns its TMEM columns and barriers -> let the next kernel in the stream get scheduled; then
  // wait for the producer of our operands before the first global access (TMA loads / stores below)
  pdl_trigger();
  pdl_wait();

  // residual of sub-tile vt -> staging buffer sb (the buffer is free: nothing has used it yet / its last store has
  // been read); TMA store of the staged sub-tile vt out of buffer sb.  Issued by the store thread of the CTA.
  auto issue_res = [&](int t, int h, int sb) {
    const Tile ti = decode_tile(p, t, BLOCK_N, MT, h);
    // the residual has the output's element size: 64-column boxes of a 16-bit tensor, 32-column boxes of an fp32 one
    const int rcols = p.y_fp32 ? 32 : 64;
    mbar_arrive_expect_tx(&res_full[sb], stg_bytes);
    for (int b = 0; b < BLOCK_N / rcols; ++b) {
      const uint32_t dst = stage_out + sb * stg_bytes + (uint32_t)b * (128 * 128);
      if (p.mode == 0) tma_load_2d(dst, &tmap_res, &res_full[sb], ti.n0 + b * rcols, ti.m0);
      else tma_load_4d(dst, &tmap_res, &res_full[sb], ti.n0 + b * rcols, ti.x0, ti.y0, ti.img);
    }
  };
  auto store_subtile = [&](int t, int h, int sb) {
    const Tile ti = decode_tile(p, t, BLOCK_N, MT, h);
    const uint32_t stg = stage_out + sb * stg_bytes;
    const int boxes = p.y_fp32 ? BLOCK_N / 32 : BLOCK_N / 64;
    const int cols = p.y_fp32 ? 32 : 64;
    for (int b = 0; b < boxes; ++b) {
      if (ti.n0 + b * cols >= p.K) break;   // (K = 32 on a 64-wide tile: nothing to write past the last filter)
      const uint32_t src = stg + (uint32_t)b * (128 * 128);
      if (p.mode == 0) tma_store_2d(&tmap_y, src, ti.n0 + b * cols, ti.m0);
      else tma_store_4d_(&tmap_y, src, ti.n0 + b * cols, ti.x0, ti.y0, ti.img);
    }
    tma_store_commit();
  };

  if (warp == kEpiWarps + 1) {
    // ------------------------------------------------------------ activation producer (one thread)
    if (lane == 0) {
      int as = 0;
      uint32_t aph = 0;
      Trace tr(p.trace, 0);
      // every tile is claimed, the first one too: a CTA that starts late finds the counter past the end and leaves
      // (measured against a static first tile: 0.02 ms slower alone, 0.03 ms faster next to the other kernels of a step)
      int t = atomicAdd(p.sched, 1);
      if (t >= total_tiles) t = -1;
      for (int it = 0;; ++it) {
        // ring slot reuse: never more than kSchedRing tiles ahead of the store issuer (does not wait in practice)
        while (it - atomicAdd(sched_done, 0) >= kSchedRing) { }
        sched_tile[it & (kSchedRing - 1)] = t;
        mbar_arrive(&sched_full[it & (kSchedRing - 1)]);   // (release: the id is visible to whoever acquires the phase)
        if (t < 0) break;
        int t_next = atomicAdd(p.sched, 1);                // in flight while this tile's loads are issued
        Tile ti[MT];
#pragma unroll
        for (int h = 0; h < MT; ++h) ti[h] = decode_tile(p, t, BLOCK_N, MT, h);
        for (int cc = 0; cc < p.c_chunks; ++cc) {
          mbar_wait(&a_empty[as], aph ^ 1, 21);
          tr.rec(2, it * p.c_chunks + cc);
          mbar_arrive_expect_tx(&a_full[as], MT * p.a_tx);
#pragma unroll
          for (int h = 0; h < MT; ++h) {
            const uint32_t dst = smem_a0 + as * a_slot + h * p.a_bytes;
            if (p.mode == 0) tma_load_2d(dst, &tmap_x, &a_full[as], cc * p.kc, ti[h].m0);
            else if (p.mode == 1) tma_load_4d(dst, &tmap_x, &a_full[as], cc * p.kc, ti[h].x0 - 1, ti[h].y0 - 1, ti[h].img);
            else if (p.mode == 3) {
              // K chunk cc = (tap, channel chunk); a 3x3 tap (r, s) samples the input at (2y + r - 1, 2x + s - 1):
              // the same strided box, shifted — pixels left of / above the image arrive as zeros
              int ch, dx = 0, dy = 0;
              const int tap = fast_divmod(cc, p.cpt, p.rcp_cpt, ch);
              if (p.taps_s2 == 9) { dy = tap / 3 - 1; dx = tap - (tap / 3) * 3 - 1; }
              tma_load_4d(dst, &tmap_x, &a_full[as], ch * p.kc, p.box_stride * ti[h].x0 + dx, p.box_stride * ti[h].y0 + dy,
                          ti[h].img);
            }
            else tma_load_4d(dst, &tmap_x, &a_full[as], 0, ti[h].x0, ti[h].y0 + cc, ti[h].img);
          }
          tr.rec(3, it * p.c_chunks + cc);
          if (++as == p.a_stages) { as = 0; aph ^= 1; }
        }
        t = t_next < total_tiles ? t_next : -1;
      }
      // last CTA out re-arms the launch's scheduler slot (every CTA has made its final claim by then)
      if (atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) { p.sched[0] = 0; p.sched[1] = 0; }
    }
    __syncwarp();
  } else if (warp == kEpiWarps + 2) {
    {
      // ---------------------------------------------------------- weight producer (one thread)
      if (lane == 0) {
        int bs = 0;
        uint32_t bph = 0;
        Trace tr(p.trace, 3);
        for (int it = 0;; ++it) {
          if (p.b_stationary && it > 0) break;   // stationary: the slice of the (only) filter tile is loaded once
          const int tb = tile_of(it);
          if (tb < 0) break;
          int n_tile;
          fast_divmod(tb, p.n_tiles, p.rcp_n_tiles, n_tile);
          const int n0 = n_tile * BLOCK_N;
          for (int cc = 0; cc < p.c_chunks; ++cc) {
            int col = cc * p.kc;
            for (int tap = 0; tap < p.taps; ++tap, col += p.w_cols) {
              mbar_wait(&b_empty[bs], bph ^ 1, 22);
              tr.rec(4, it * p.c_chunks + cc);
              mbar_arrive_expect_tx(&b_full[bs], B_BYTES);
              tma_load_2d(smem_b0 + bs * B_BYTES, &tmap_w, &b_full[bs], col, n0);
              tr.rec(5, it * p.c_chunks + cc);
              if (++bs == p.b_stages) { bs = 0; bph ^= 1; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kEpiWarps || warp == kEpiWarps + 3) {
    if (warp == kEpiWarps + 3) {
      // ---------------------------------------------------------- store issuer (one thread): TMA store of each staged
      // sub-tile; the staging buffer is handed on (or refilled with its next residual) once the store has read it
      if (lane == 0) {
        Trace tr(p.trace, 4);
        const int S = p.stg_bufs;
        int sb = 0;
        uint32_t ph = 0;
        // residual prefetch walks the tile ring ahead of the stores: (rp_it, rp_h) = next sub-tile whose residual is
        // still to be requested, rp_t its tile id (-1 = past the last tile)
        int rp_it = 0, rp_h = 0, rp_t = p.has_res ? tile_of(0) : -1;
        auto res_next = [&](int buffer) {
          if (rp_t < 0) return;
          issue_res(rp_t, rp_h, buffer);
          if (++rp_h == MT) { rp_h = 0; rp_t = tile_of(++rp_it); }
        };
        if (p.has_res)
          for (int i = 0; i < S; ++i) res_next(i);
        int t_cur = -1;
        for (int vt = 0;; ++vt) {
          const int h = vt % MT;
          if (h == 0) {
            t_cur = tile_of(vt / MT);
            if (t_cur < 0) break;
          }
          mbar_wait(&staged[sb], ph, 30);
          store_subtile(t_cur, h, sb);
          tma_store_wait_read_n<0>();   // the store has left its staging buffer: hand the buffer on at once
          if (!p.has_res) mbar_arrive(&stg_free[sb]);
          else res_next(sb);
          tr.rec(24, vt);
          if (h == MT - 1) atomicExch(sched_done, vt / MT + 1);
          if (++sb == S) { sb = 0; ph ^= 1; }
        }
      }
    } else if (lane == 0) {
      // ---------------------------------------------------------- MMA issuer (one thread: no per-step warp
      // reconvergence).  Two issuing threads (one per sub-tile) were measured and bought nothing: the tensor pipe,
      // not the issue rate, paces these tiles (tools/micro/mma_issue.cu).
      const uint32_t idesc = p.tf32 ? make_idesc_tf32(128, BLOCK_N) : make_idesc_f16(128, BLOCK_N, p.ab_half ? 0u : 1u);
      const uint32_t sbo = (uint32_t)p.halo_w * 128u;
      // descriptor constants: only the 14-bit start-address field changes between MMAs
      const uint64_t a_hi = desc_sbo(0, sbo), b_hi = make_smem_desc(0, 128);
      const int rs = p.taps == 9 ? 3 : 1;
      const uint32_t row_step = (uint32_t)p.halo_w * 128u;
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      Trace tr(p.trace, 1);
      // The ring barriers are tested ONE stage ahead (mbar_test, non-blocking) before the current stage's MMAs
      // are issued, so the ~180-cycle round trip of the barrier query hides behind the MMA issue instead of
      // sitting between every two k-chunks (ptx_sm100.cuh: mbar_test).
      bool a_ok = false, b_ok = false;   // stage (as / bs) already seen full by a look-ahead test
      for (int it = 0;; ++it) {
        if (tile_of(it) < 0) break;
        const int buf = it & 1;
        const uint32_t u = (uint32_t)(it >> 1);
        mbar_wait(&acc_empty[buf], (u & 1) ^ 1, 24);  // epilogue has drained this accumulator set
        tr.rec(10, it);
        tc_fence_after_sync();
        const uint32_t tmem_acc = tmem_base + buf * (MT * BLOCK_N);
        const bool wait_b = !p.b_stationary || it == 0;
        uint32_t first = 0;  // 0 until the first MMA of the tile has been issued (it overwrites the accumulator)
        for (int cc = 0; cc < p.c_chunks; ++cc) {
          if (!a_ok) mbar_wait(&a_full[as], aph, 25);
          a_ok = false;
          tr.rec(11, it * p.c_chunks + cc);
          uint32_t a_row = smem_a0 + as * a_slot;
          // next activation stage (may belong to the next tile: the ring runs across tile boundaries)
          int as_n = as + 1;
          uint32_t aph_n = aph;
          if (as_n == p.a_stages) { as_n = 0; aph_n ^= 1; }
          const bool more_a = true;   // (a look-ahead test past the last tile just never succeeds)
          for (int r = 0; r < rs; ++r, a_row += row_step) {
            uint32_t a_tap = a_row;
            for (int sx = 0; sx < rs; ++sx, a_tap += 128) {
              if (wait_b && !b_ok) mbar_wait(&b_full[bs], bph, 26);
              b_ok = false;
              tc_fence_after_sync();
              const uint64_t bdesc = b_hi | (uint64_t)(((smem_b0 + bs * B_BYTES) & 0x3FFFFu) >> 4);
              // look one stage ahead before issuing this stage's MMAs
              int bs_n = bs + 1;
              uint32_t bph_n = bph;
              if (bs_n == p.b_stages) { bs_n = 0; bph_n ^= 1; }
              const bool last_tap = r == rs - 1 && sx == rs - 1;
              const bool last_of_tile = last_tap && cc == p.c_chunks - 1;
              const bool wait_b_next = last_of_tile ? !p.b_stationary : wait_b;
              lookahead_test_b(&b_full[bs_n], bph_n, wait_b_next);
              lookahead_test_a(&a_full[as_n], aph_n, last_tap && more_a);
#pragma unroll
              for (int h = 0; h < MT; ++h) {
                const uint64_t adesc = a_hi | (uint64_t)(((a_tap + h * p.a_bytes) & 0x3FFFFu) >> 4);
                if (p.tf32) {
                  umma_tf32(tmem_acc + h * BLOCK_N, adesc, bdesc, idesc, first);
#pragma unroll
                  for (int k = 1; k < 4; ++k) umma_tf32(tmem_acc + h * BLOCK_N, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
                } else {
                  umma_bf16(tmem_acc + h * BLOCK_N, adesc, bdesc, idesc, first);
#pragma unroll
                  for (int k = 1; k < 4; ++k) umma_bf16(tmem_acc + h * BLOCK_N, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
                }
              }
              first = 1;
              if (!p.b_stationary) umma_commit(&b_empty[bs]);
              if (last_tap) umma_commit(&a_empty[as]);
              tr.rec(13, it * p.c_chunks + cc);
              bs = bs_n; bph = bph_n;
              b_ok = lookahead_read_b();
              if (last_tap) a_ok = lookahead_read_a();
            }
          }
          as = as_n; aph = aph_n;
        }
        umma_commit(&acc_full[buf]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue (warps 0-7)
    // warp w reads TMEM lanes 32*(w&3).. (its hardware-accessible quarter) and the column half w>>2.
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;
    Trace tr(threadIdx.x == 0 ? p.trace : nullptr, 2);
    int sb = 0;            // staging buffer of the next sub-tile
    uint32_t sph = 0;      // parity of this buffer's use count
    for (int it = 0;; ++it) {
      const int t_epi = tile_of(it);
      if (t_epi < 0) break;
      const int buf = it & 1;
      const uint32_t u = (uint32_t)(it >> 1);
      int n_tile;
      fast_divmod(t_epi, p.n_tiles, p.rcp_n_tiles, n_tile);
      const int n0 = n_tile * BLOCK_N;
      // this warp's folded-norm constants for the tile, fetched while the accumulators are still filling
      tr.rec(20, it);
      mbar_wait(&acc_full[buf], u & 1, 27);
      tr.rec(21, it);
      tc_fence_after_sync();
#pragma unroll 1
      for (int h = 0; h < MT; ++h) {
        const int vt = it * MT + h;   // sub-tiles are the unit of staging / residual / store
        const uint32_t stg = stage_out + sb * stg_bytes;
        if (p.has_res) mbar_wait(&res_full[sb], sph, 28);     // residual landed (implies staging free)
        else mbar_wait(&stg_free[sb], sph ^ 1, 29);           // previous store has left the staging buffer
        tr.rec(25, vt);
        const uint32_t t_row = tmem_base + (buf * MT + h) * BLOCK_N + ((uint32_t)(q * 32) << 16);
        uint32_t vv[2][32];
        tmem_ld_32x32(t_row + (half * NI) * 32, vv[0]);
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const int cb = half * NI + i;
          const int col0 = n0 + cb * 32;
          float4 sc[8], sh[8];
          if (p.scale != nullptr) {   // issued before the TMEM wait: both latencies overlap
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              sc[j4] = __ldg(reinterpret_cast<const float4*>(p.scale + col0) + j4);
              sh[j4] = __ldg(reinterpret_cast<const float4*>(p.shift + col0) + j4);
            }
          }
          tmem_ld_wait();
          tr.rec(26, vt * 8 + i);
          if (i + 1 < NI) tmem_ld_32x32(t_row + (cb + 1) * 32, vv[(i + 1) & 1]);
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(vv[i & 1][j]);
          if (p.scale != nullptr) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              f[j4 * 4 + 0] = fmaf(f[j4 * 4 + 0], sc[j4].x, sh[j4].x);
              f[j4 * 4 + 1] = fmaf(f[j4 * 4 + 1], sc[j4].y, sh[j4].y);
              f[j4 * 4 + 2] = fmaf(f[j4 * 4 + 2], sc[j4].z, sh[j4].z);
              f[j4 * 4 + 3] = fmaf(f[j4 * 4 + 3], sc[j4].w, sh[j4].w);
            }
          }
          if (p.y_fp32) {
            const uint32_t box = stg + (uint32_t)cb * (128 * 128) + (uint32_t)row * 128;
            if (p.has_res) {   // fp32 residual: same swizzled staging position as the output it becomes
#pragma unroll
              for (int q4 = 0; q4 < 8; ++q4) {
                const uint32_t addr = box + (uint32_t)((q4 ^ (row & 7)) << 4);
                float r0, r1, r2, r3;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(r0), "=f"(r1), "=f"(r2), "=f"(r3)
                             : "r"(addr)
                             : "memory");
                f[q4 * 4] += r0; f[q4 * 4 + 1] += r1; f[q4 * 4 + 2] += r2; f[q4 * 4 + 3] += r3;
              }
            }
            if (p.relu & 1) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if (p.tf32 && !(p.relu & 2)) {   // (relu bit 1: GEMM use — the fp32 result is final, keep every bit)
              // the next layer's kind::tf32 MMA drops the low 13 mantissa bits of what it reads: round to nearest
              // here (cvt.rna.tf32.f32), so the activation chain carries unbiased TF32 values instead of truncated ones
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                uint32_t t;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(f[j]));
                f[j] = __uint_as_float(t);
              }
            }
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4) {
              const uint32_t addr = box + (uint32_t)((q4 ^ (row & 7)) << 4);
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(f[q4 * 4]), "f"(f[q4 * 4 + 1]),
                           "f"(f[q4 * 4 + 2]), "f"(f[q4 * 4 + 3])
                           : "memory");
            }
          } else {
            const uint32_t box = stg + (uint32_t)(cb >> 1) * (128 * 128) + (uint32_t)row * 128;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const uint32_t addr = box + (uint32_t)((((cb & 1) * 4 + q4) ^ (row & 7)) << 4);
              if (p.has_res) {
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
                if (p.relu & 1) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                if (p.ab_half) {
                  __half2 h2 = __floats2half2_rn(a, b);
                  o[e] = *reinterpret_cast<uint32_t*>(&h2);
                } else {
                  __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
                  o[e] = *reinterpret_cast<uint32_t*>(&h2);
                }
              }
              asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[0]), "r"(o[1]), "r"(o[2]),
                           "r"(o[3])
                           : "memory");
            }
          }
        }
        if (h == MT - 1) {  // this thread's share of both accumulators is in registers / staged: hand them back
          tc_fence_before_sync();
          mbar_arrive(&acc_empty[buf]);
        }
        tr.rec(22, vt);
        // this thread's part of the sub-tile is in the staging buffer: make it visible to the async proxy and
        // count in — the store issuer (warp 10) sends the TMA store once all 256 threads have arrived
        fence_proxy_async_smem();
        mbar_arrive(&staged[sb]);
        if (++sb == p.stg_bufs) { sb = 0; sph ^= 1; }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kEpiWarps) tmem_dealloc(tmem_base, 2 * MT * BLOCK_N);
}

// ---------------------------------------------------------------------------------------
// Space-to-depth operand of the RGB stem: cell (Y, X) of image n holds the 2x2 block of input
// pixels (2Y+dy-3, 2X+dx-3), dy,dx in {0,1}, as 12 channels (dy*2+dx)*3+c followed by 4 zeros
// (16 x 16-bit = 32 B per cell); pixels outside the image are zero (= the conv's padding of 3).
// With it the 7x7/2 conv is a stride-1 4x4 conv: tap (r', s') channel (dy,dx,c) carries the
// original weight (2r'+dy, 2s'+dx, c) (zero where that index is 7).
// ---------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void stem_s2d_kernel(const TI* __restrict__ x, TO* __restrict__ out, int N, int H, int W, int Ys,
                                int Xs) {
  const long long cells = (long long)N * Ys * Xs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < cells;
       i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % Xs);
    const int Y = (int)((i / Xs) % Ys);
    const int n = (int)(i / ((long long)Xs * Ys));
    float f[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) f[e] = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const int ih = 2 * Y + dy - 3;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int iw = 2 * X + dx - 3;
        if ((unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W) {
          const TI* px = x + ((size_t)(n * H + ih) * W + iw) * 3;
#pragma unroll
          for (int c = 0; c < 3; ++c) f[(dy * 2 + dx) * 3 + c] = to_float<TI>(px[c]);
        }
      }
    }
    __align__(16) TO o[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) o[e] = from_float<TO>(f[e]);
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)i * 16);
    dst[0] = *reinterpret_cast<const uint4*>(&o[0]);
    dst[1] = *reinterpret_cast<const uint4*>(&o[8]);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_enc() {
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

int map_nd(EncodeTiledFn enc, CUtensorMap* m, const void* base, CUtensorMapDataType dt, int rank,
           const cuuint64_t* gdim, const cuuint64_t* gstr, const cuuint32_t* box, const char* what,
           const cuuint32_t* estr_in = nullptr) {
  const cuuint32_t estr1[4] = {1, 1, 1, 1};
  const cuuint32_t* estr = estr_in ? estr_in : estr1;
  CUresult r = enc(m, dt, rank, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("conv_tma: cuTensorMapEncodeTiled(%s) failed (%d)", what, (int)r);
    return VLNCE_ERR_CUDA_BASE + 1;
  }
  return VLNCE_OK;
}

long long* g_trace = nullptr;
// Scheduler slots of the dynamic tile schedule: {next tile, finished CTAs} per launch, handed out round-robin (a slot
// is re-armed to zero by the last CTA of its launch; 1024 launches later it is long idle).
constexpr int kSchedSlots = 1024;
__device__ int g_sched_mem[2 * kSchedSlots];
int* g_sched = nullptr;
unsigned g_sched_seq = 0;
// Upper bound on the persistent grid (0 = one CTA per SM): leaving a few SMs free lets the short kernels of
// concurrent graph branches (depth trunk, head) run beside the RGB trunk instead of queueing behind it.
int g_cta_cap = -1;

template <int BLOCK_N, int MT>
int launch(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& ty, const CUtensorMap& tr,
           TmaConvParams& p, int b_stages_req, int no_stationary, cudaStream_t stream) {
  static bool configured = false;
  static int num_sms = 0;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tma_kernel<BLOCK_N, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kSmemBudget);
    if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(conv_tma)");
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
    configured = true;
  }
  const int stg = 128 * BLOCK_N * (p.y_fp32 ? 4 : 2);
  const int B_BYTES = BLOCK_N * 128;
  const int a_slot = MT * p.a_bytes;
  auto bytes = [&](int as, int bs, int sbufs) { return as * a_slot + bs * B_BYTES + 1024 + sbufs * stg + 1024; };
  int as, bs, sbufs;
  // A third staging buffer gives the residual prefetch (and the lagging buffer release of the store issuer) a whole
  // sub-tile of slack; it is given up before the rings drop below three stages, except under a residual.
  // (1) stationary weights: this CTA's whole weight slice (all taps x chunks) stays in shared memory
  const int b_all = p.taps * p.c_chunks;
  p.b_stationary = 0;
  if (!no_stationary && p.n_tiles == 1 && b_all <= kMaxB) {
    as = p.mode == 1 ? 3 : 4; sbufs = 3;
    while (bytes(as, b_all, sbufs) > kSmemBudget) {
      if (as > 3) --as;
      else if (sbufs > 2 && !p.has_res) --sbufs;
      else if (as > 2) --as;
      else if (sbufs > 1) --sbufs;
      else break;
    }
    if (bytes(as, b_all, sbufs) <= kSmemBudget) { p.b_stationary = 1; bs = b_all; }
  }
  if (!p.b_stationary) {
    // (2) streaming: deep rings, then staging; shrink until it fits
    as = p.mode == 1 ? 3 : 4; bs = b_stages_req > 0 ? b_stages_req : 4; sbufs = 3;
    if (bs > kMaxB) bs = kMaxB;
    while (bytes(as, bs, sbufs) > kSmemBudget) {
      if (bs > 4) --bs;
      else if (as > 3) --as;
      else if (sbufs > 2 && !p.has_res) --sbufs;
      else if (bs > 3) --bs;
      else if (sbufs > 2) --sbufs;
      else if (as > 2) --as;
      else if (sbufs > 1) --sbufs;
      else if (bs > 2) --bs;
      else {
        set_error("conv_tma: tile does not fit shared memory (BLOCK_N=%d, MT=%d)", BLOCK_N, MT);
        return VLNCE_ERR_INVALID_ARGUMENT;
      }
    }
  }
  // experiment overrides (ring depths / staging buffers / resident CTAs per SM), read once
  static int ov_as = -1, ov_bs = -1, ov_stg = -1, ov_ctas = -1;
  if (ov_as < 0) {
    const char* e;
    ov_as = (e = getenv("VLNCE_TMA_AS")) ? atoi(e) : 0;
    ov_bs = (e = getenv("VLNCE_TMA_BS")) ? atoi(e) : 0;
    ov_stg = (e = getenv("VLNCE_TMA_STG")) ? atoi(e) : 0;
    ov_ctas = (e = getenv("VLNCE_TMA_CTAS")) ? atoi(e) : 0;
  }
  if (ov_as > 0 && ov_as <= kMaxA) as = ov_as;
  if (ov_bs > 0 && ov_bs <= kMaxB && !p.b_stationary) bs = ov_bs;
  if (ov_stg > 0 && ov_stg <= kMaxStg) sbufs = ov_stg;
  while (bytes(as, bs, sbufs) > kSmemBudget) {  // overrides must still fit
    if (!p.b_stationary && bs > 2) --bs;
    else if (as > 2) --as;
    else if (sbufs > 1) --sbufs;
    else break;
  }
  p.a_stages = as; p.b_stages = bs; p.stg_bufs = sbufs;
  p.trace = g_trace;
  if (g_sched == nullptr) {
    void* ptr = nullptr;
    cudaError_t e = cudaGetSymbolAddress(&ptr, g_sched_mem);
    if (e != cudaSuccess) return cuda_status(e, "cudaGetSymbolAddress(conv_tma scheduler)");
    g_sched = reinterpret_cast<int*>(ptr);
  }
  p.sched = g_sched + 2 * (g_sched_seq++ % kSchedSlots);
  p.rcp_n_tiles = 1.0f / (float)p.n_tiles;
  p.rcp_tiles_x = 1.0f / (float)p.tiles_x;
  p.rcp_tiles_y = 1.0f / (float)p.tiles_y;
  const int total = (p.m_tiles / MT) * p.n_tiles;
  int ctas = num_sms;
  if (g_cta_cap < 0) {
    const char* e = getenv("VLNCE_TMA_CTA_CAP");
    g_cta_cap = e ? atoi(e) : 0;
  }
  if (g_cta_cap > 0 && g_cta_cap < ctas) ctas = g_cta_cap;
  if (ov_ctas > 1 && bytes(as, bs, sbufs) * ov_ctas <= 228 * 1024 - 1024 * ov_ctas && 2 * MT * BLOCK_N * ov_ctas <= 512)
    ctas = num_sms * ov_ctas;
  const int grid = total < ctas ? total : ctas;
  launch_pdl(conv_tma_kernel<BLOCK_N, MT>, dim3(grid), dim3(kThreads), (size_t)bytes(as, bs, sbufs), stream, tx, tw, ty,
             tr, p);
  VLNCE_LAUNCH_CHECK("conv_tma_kernel");
  return VLNCE_OK;
}

int dispatch(int bn, int mt, const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& ty,
             const CUtensorMap& tr, TmaConvParams& p, int b_stages, int no_stationary, cudaStream_t stream) {
  if (mt == 2 && (p.m_tiles % 2 != 0 || bn == 256)) mt = 1;  // pairs need an even tile count and 4 accumulators
  if (bn == 256) return launch<256, 1>(tx, tw, ty, tr, p, b_stages, no_stationary, stream);
  if (bn == 128) return mt == 2 ? launch<128, 2>(tx, tw, ty, tr, p, b_stages, no_stationary, stream)
                                : launch<128, 1>(tx, tw, ty, tr, p, b_stages, no_stationary, stream);
  return mt == 2 ? launch<64, 2>(tx, tw, ty, tr, p, b_stages, no_stationary, stream)
                 : launch<64, 1>(tx, tw, ty, tr, p, b_stages, no_stationary, stream);
}

}  // namespace
}  // namespace vlnce

using namespace vlnce;

// block_n argument of the conv_tma entry points: tile width (64/128/256, 0 = auto) | weight ring depth << 16
// | (1 << 21): two pixel sub-tiles per scheduled tile share each weight stage | (1 << 22): never keep the
// weight slice stationary in shared memory.
#define VLNCE_TMA_MT2 (1 << 21)
#define VLNCE_TMA_NO_STATIONARY (1 << 22)

// R = 1 (1x1, pad 0) or R = 3 (3x3, pad 1); stride 1.
extern "C" int vlnce_set_tma_cta_cap(int cap) {
  g_cta_cap = cap < 0 ? 0 : cap;
  return VLNCE_OK;
}

static int conv_box_tma(const void* x, const void* w, int ab_dtype, const float* scale, const float* shift, void* y,
                        int y_dtype, int N, int H, int W, int C, int K, int R, int stride, int relu, int block_n,
                        cudaStream_t stream);

extern "C" int vlnce_conv_tma(const void* x, const void* w, int ab_dtype, const float* scale, const float* shift,
                              const void* residual, void* y, int y_dtype, int N, int H, int W, int C, int K, int R,
                              int relu, int block_n, cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && w && y && N > 0 && H > 0 && W > 0, "conv_tma: null tensor / empty shape");
  VLNCE_CHECK_ARG(R == 1 || R == 3, "conv_tma: only 1x1 and 3x3 stride-1 convolutions");
  // 3x3 on images too small for the 16x8 halo tiles (8x8, 4x8 ...): tap-shifted boxes over 128/(8*H) images per tile
  if (R == 3 && H % TH != 0 && H < TH && W % TW == 0 && 128 % (8 * H) == 0 && N % (128 / (8 * H)) == 0 && residual == nullptr)
    return conv_box_tma(x, w, ab_dtype, scale, shift, y, y_dtype, N, H, W, C, K, 3, 1, relu, block_n, stream);
  // ab_dtype F32: the fp32 operands are multiplied as TF32 (kind::tf32), output / residual fp32, 32-channel K chunks
  const bool tf32 = ab_dtype == VLNCE_F32;
  const int kc = tf32 ? 32 : 64, eb = tf32 ? 4 : 2;
  // K = 32 (the first stage of the DD-PPO depth ResNet) rides on a 64-wide tile in the fp32 / raw-output mode: the
  // weight box reads rows 32..63 out of bounds (zeros), the store skips the columns past K
  const bool k32 = tf32 && K == 32 && scale == nullptr && residual == nullptr;
  VLNCE_CHECK_ARG(C % kc == 0 && (K % 64 == 0 || k32), "conv_tma: C must be a multiple of %d and K of 64", kc);
  // 3x3: 16x8-pixel tiles; an image whose height / width is not a multiple of the tile is covered by ragged border
  // tiles — the TMA unit zero-fills the part of a load box that lies outside the tensor (exactly the conv's zero padding)
  // and clips the store box, so nothing else in the kernel knows (224-pixel frames: 56 / 28 / 14-pixel feature maps)
  VLNCE_CHECK_ARG(R == 1 || (H >= 8 && W >= TW), "conv_tma: 3x3 needs H >= 8 and W >= 8 (smaller images: the box mode)");
  VLNCE_CHECK_ARG((scale == nullptr) == (shift == nullptr), "conv_tma: scale/shift must come together");
  VLNCE_CHECK_ARG(ab_dtype == VLNCE_BF16 || ab_dtype == VLNCE_F16 || tf32, "conv_tma: operand dtype");
  VLNCE_CHECK_ARG(y_dtype == ab_dtype || y_dtype == VLNCE_F32, "conv_tma: y dtype");
  VLNCE_CHECK_ARG(!(residual && y_dtype == VLNCE_F32 && !tf32), "conv_tma: an fp32 residual needs fp32 operands");
  VLNCE_CHECK_ARG((long long)N * H * W < (1ll << 31), "conv_tma: too many pixels");
  const int b_stages = (block_n >> 16) & 0xf;
  const int mt = (block_n & VLNCE_TMA_MT2) ? 2 : 1;
  const int no_stationary = (block_n & VLNCE_TMA_NO_STATIONARY) ? 1 : 0;
  int bn = block_n & 0xffff;
  if (bn == 0) bn = (K % 128 == 0 && !tf32) ? 128 : 64;
  if (k32) bn = 64;
  VLNCE_CHECK_ARG((bn == 64 || bn == 128 || bn == 256) && (K % bn == 0 || k32), "conv_tma: block_n=%d vs K=%d", bn, K);
  {  // the smallest configuration (2+2 ring slots, one staging buffer) must fit: narrow wide fp32 tiles
    const int a_b = (R == 1 ? 16 * 1024 : 23 * 1024) * mt;
    while (bn > 64 && 2 * a_b + 2 * bn * 128 + 2048 + 128 * bn * (y_dtype == VLNCE_F32 ? 4 : 2) > kSmemBudget) bn /= 2;
  }
  EncodeTiledFn enc = get_enc();
  if (!enc) {
    set_error("conv_tma: cuTensorMapEncodeTiled unavailable");
    return VLNCE_ERR_NO_DRIVER;
  }
  const CUtensorMapDataType dt16 = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                        : (ab_dtype == VLNCE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  const bool f32 = y_dtype == VLNCE_F32;
  const long long M = (long long)N * H * W;
  CUtensorMap tx, tw, ty, tr;
  int rc;
  {  // weights [K, R*R*C]
    const cuuint64_t gd[2] = {(cuuint64_t)R * R * C, (cuuint64_t)K};
    const cuuint64_t gs[1] = {(cuuint64_t)R * R * C * eb};
    const cuuint32_t bx[2] = {(cuuint32_t)kc, (cuuint32_t)bn};
    if ((rc = map_nd(enc, &tw, w, dt16, 2, gd, gs, bx, "weights")) != VLNCE_OK) return rc;
  }
  TmaConvParams p;
  if (R == 1) {
    const cuuint64_t gd[2] = {(cuuint64_t)C, (cuuint64_t)M};
    const cuuint64_t gs[1] = {(cuuint64_t)C * eb};
    const cuuint32_t bx[2] = {(cuuint32_t)kc, 128};
    if ((rc = map_nd(enc, &tx, x, dt16, 2, gd, gs, bx, "activation")) != VLNCE_OK) return rc;
    const cuuint64_t yd[2] = {(cuuint64_t)K, (cuuint64_t)M};
    const cuuint64_t ys[1] = {(cuuint64_t)K * (f32 ? 4 : 2)};
    const cuuint32_t yb[2] = {(cuuint32_t)(f32 ? 32 : 64), 128};
    if ((rc = map_nd(enc, &ty, y, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : dt16, 2, yd, ys, yb, "output")) != VLNCE_OK)
      return rc;
    tr = ty;
    if (residual) {
      const cuuint64_t rs[1] = {(cuuint64_t)K * eb};
      const cuuint32_t rb[2] = {(cuuint32_t)(tf32 ? 32 : 64), 128};
      if ((rc = map_nd(enc, &tr, residual, dt16, 2, yd, rs, rb, "residual")) != VLNCE_OK) return rc;
    }
    p.mode = 0; p.taps = 1; p.halo_w = 8; p.a_bytes = 16 * 1024; p.a_tx = 16 * 1024; p.img_per_tile = 1; p.cpt = 1; p.rcp_cpt = 1.f; p.taps_s2 = 1; p.box_stride = 1;
    p.m_tiles = cdiv(M, 128); p.tiles_x = 1; p.tiles_y = 1; p.tw = 128; p.th = 1;
  } else {
    const cuuint64_t gd[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t gs[3] = {(cuuint64_t)C * eb, (cuuint64_t)W * C * eb, (cuuint64_t)H * W * C * eb};
    const cuuint32_t bx[4] = {(cuuint32_t)kc, TW + 2, TH + 2, 1};
    if ((rc = map_nd(enc, &tx, x, dt16, 4, gd, gs, bx, "halo")) != VLNCE_OK) return rc;
    const int esz = f32 ? 4 : 2;
    const cuuint64_t yd[4] = {(cuuint64_t)K, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t ys[3] = {(cuuint64_t)K * esz, (cuuint64_t)W * K * esz, (cuuint64_t)H * W * K * esz};
    const cuuint32_t yb[4] = {(cuuint32_t)(f32 ? 32 : 64), TW, TH, 1};
    if ((rc = map_nd(enc, &ty, y, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : dt16, 4, yd, ys, yb, "output")) != VLNCE_OK)
      return rc;
    tr = ty;
    if (residual) {
      const cuuint64_t rs[3] = {(cuuint64_t)K * eb, (cuuint64_t)W * K * eb, (cuuint64_t)H * W * K * eb};
      const cuuint32_t rb[4] = {(cuuint32_t)(tf32 ? 32 : 64), TW, TH, 1};
      if ((rc = map_nd(enc, &tr, residual, dt16, 4, yd, rs, rb, "residual")) != VLNCE_OK) return rc;
    }
    p.mode = 1; p.taps = 9; p.halo_w = TW + 2; p.a_bytes = 23 * 1024; p.a_tx = (TW + 2) * (TH + 2) * 128; p.img_per_tile = 1; p.cpt = 1; p.rcp_cpt = 1.f; p.taps_s2 = 1; p.box_stride = 1;
    p.tiles_x = cdiv(W, TW); p.tiles_y = cdiv(H, TH); p.m_tiles = N * p.tiles_x * p.tiles_y; p.tw = TW; p.th = TH;
  }
  p.scale = scale; p.shift = shift;
  p.M = (int)M; p.N = N; p.H = H; p.W = W; p.C = C; p.K = K;
  p.c_chunks = C / kc;
  p.w_cols = C;
  p.n_tiles = cdiv(K, bn);
  p.tf32 = tf32 ? 1 : 0; p.kc = kc;
  p.relu = relu; p.y_fp32 = f32; p.ab_half = (ab_dtype == VLNCE_F16); p.has_res = residual != nullptr;
  p.a_stages = p.b_stages = p.stg_bufs = 0; p.b_stationary = 0;
  return dispatch(bn, mt, tx, tw, ty, tr, p, b_stages, no_stationary, stream);
}

// Stride-2 convolutions, 1x1 pad 0 (the ResNet downsample branches) and 3x3 pad 1 (the first block of a stage):
// x (N,H,W,C) -> y (N,H/2,W/2,K).  Every activation tile is ONE 4-D TMA box with element strides {1,2,2,1}: 128 sampled
// pixels x 64 channels land as a plain K-major tile; the 1x1 never loads the three quarters of the input it does not
// read, the 3x3 is nine such boxes per channel chunk, shifted by the tap (im2col by tensor map, borders zero-filled by
// the TMA unit).  A tile is 8 x bh output pixels of 128/(8*bh) images (bh = min(16, H/2)).  Needs H, W even,
// (W/2) % 8 == 0, (H/2) % bh == 0, N % images-per-tile == 0.
static int conv_box_tma(const void* x, const void* w, int ab_dtype, const float* scale, const float* shift, void* y,
                        int y_dtype, int N, int H, int W, int C, int K, int R, int stride, int relu, int block_n,
                        cudaStream_t stream) {
  VLNCE_CHECK_ARG(R == 1 || R == 3, "conv_s2_tma: 1x1 or 3x3");
  VLNCE_CHECK_ARG(x && w && y && N > 0 && H > 0 && W > 0, "conv_s2_tma: null tensor / empty shape");
  const bool tf32 = ab_dtype == VLNCE_F32;   // fp32 operands multiplied as TF32, fp32 output, 32-channel K chunks
  const int kc = tf32 ? 32 : 64, eb = tf32 ? 4 : 2;
  const bool k32 = tf32 && K == 32 && scale == nullptr;   // see vlnce_conv_tma
  VLNCE_CHECK_ARG(C % kc == 0 && (K % 64 == 0 || k32), "conv_s2_tma: C must be a multiple of %d and K of 64", kc);
  VLNCE_CHECK_ARG(stride == 1 || (H % 2 == 0 && W % 2 == 0), "conv_s2_tma: even input size");
  VLNCE_CHECK_ARG((scale == nullptr) == (shift == nullptr), "conv_s2_tma: scale/shift must come together");
  VLNCE_CHECK_ARG(ab_dtype == VLNCE_BF16 || ab_dtype == VLNCE_F16 || tf32, "conv_s2_tma: operand dtype");
  VLNCE_CHECK_ARG(y_dtype == ab_dtype || y_dtype == VLNCE_F32, "conv_s2_tma: y dtype");
  const int P = H / stride, Q = W / stride;
  const int bh = P >= 16 ? 16 : P;
  VLNCE_CHECK_ARG(Q % 8 == 0 && P % bh == 0 && 128 % (8 * bh) == 0, "conv_s2_tma: output %dx%d does not tile", P, Q);
  const int ipt = 128 / (8 * bh);
  VLNCE_CHECK_ARG(N % ipt == 0, "conv_s2_tma: %d images do not fill tiles of %d", N, ipt);
  const int b_stages = (block_n >> 16) & 0xf;
  const int mt = (block_n & VLNCE_TMA_MT2) ? 2 : 1;
  const int no_stationary = (block_n & VLNCE_TMA_NO_STATIONARY) ? 1 : 0;
  int bn = block_n & 0xffff;
  if (bn == 0) bn = (K % 128 == 0 && !tf32) ? 128 : 64;
  if (k32) bn = 64;
  VLNCE_CHECK_ARG((bn == 64 || bn == 128 || bn == 256) && (K % bn == 0 || k32), "conv_s2_tma: block_n=%d vs K=%d", bn, K);
  while (bn > 64 && 2 * 16 * 1024 * mt + 2 * bn * 128 + 2048 + 128 * bn * (y_dtype == VLNCE_F32 ? 4 : 2) > kSmemBudget) bn /= 2;
  EncodeTiledFn enc = get_enc();
  if (!enc) {
    set_error("conv_s2_tma: cuTensorMapEncodeTiled unavailable");
    return VLNCE_ERR_NO_DRIVER;
  }
  const CUtensorMapDataType dt16 = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                        : (ab_dtype == VLNCE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  const bool f32 = y_dtype == VLNCE_F32;
  CUtensorMap tx, tw, ty;
  int rc;
  {  // weights [K, R*R*C]
    const cuuint64_t gd[2] = {(cuuint64_t)R * R * C, (cuuint64_t)K};
    const cuuint64_t gs[1] = {(cuuint64_t)R * R * C * eb};
    const cuuint32_t bx[2] = {(cuuint32_t)kc, (cuuint32_t)bn};
    if ((rc = map_nd(enc, &tw, w, dt16, 2, gd, gs, bx, "weights")) != VLNCE_OK) return rc;
  }
  {  // input, sampled with stride 2 in W and H: the box spans 16 x 2*bh pixels and keeps every second one
    const cuuint64_t gd[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t gs[3] = {(cuuint64_t)C * eb, (cuuint64_t)W * C * eb, (cuuint64_t)H * W * C * eb};
    const cuuint32_t bx[4] = {(cuuint32_t)kc, (cuuint32_t)(8 * stride), (cuuint32_t)(stride * bh), (cuuint32_t)ipt};
    const cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    if ((rc = map_nd(enc, &tx, x, dt16, 4, gd, gs, bx, "strided input", es)) != VLNCE_OK) return rc;
  }
  {
    const int esz = f32 ? 4 : 2;
    const cuuint64_t yd[4] = {(cuuint64_t)K, (cuuint64_t)Q, (cuuint64_t)P, (cuuint64_t)N};
    const cuuint64_t ys[3] = {(cuuint64_t)K * esz, (cuuint64_t)Q * K * esz, (cuuint64_t)P * Q * K * esz};
    const cuuint32_t yb[4] = {(cuuint32_t)(f32 ? 32 : 64), 8, (cuuint32_t)bh, (cuuint32_t)ipt};
    if ((rc = map_nd(enc, &ty, y, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : dt16, 4, yd, ys, yb, "output")) != VLNCE_OK)
      return rc;
  }
  TmaConvParams p;
  p.mode = 3; p.taps = 1; p.halo_w = 8; p.a_bytes = 16 * 1024; p.a_tx = 16 * 1024; p.img_per_tile = ipt;
  p.cpt = C / kc; p.rcp_cpt = 1.0f / (float)p.cpt; p.taps_s2 = R * R; p.box_stride = stride;
  p.tw = 8; p.th = bh;
  p.tiles_x = Q / 8; p.tiles_y = P / bh; p.m_tiles = (N / ipt) * p.tiles_x * p.tiles_y;
  p.scale = scale; p.shift = shift;
  p.M = N * P * Q; p.N = N; p.H = P; p.W = Q; p.C = C; p.K = K;
  p.c_chunks = R * R * C / kc;   // K chunks in the weight matrix's column order: tap-major, channel chunk minor
  p.tf32 = tf32 ? 1 : 0; p.kc = kc;
  p.w_cols = C;
  p.n_tiles = cdiv(K, bn);
  p.relu = relu; p.y_fp32 = f32; p.ab_half = (ab_dtype == VLNCE_F16); p.has_res = 0;
  p.a_stages = p.b_stages = p.stg_bufs = 0; p.b_stationary = 0;
  return dispatch(bn, mt, tx, tw, ty, ty, p, b_stages, no_stationary, stream);
}

extern "C" int vlnce_conv_s2_tma(const void* x, const void* w, int ab_dtype, const float* scale, const float* shift,
                                 void* y, int y_dtype, int N, int H, int W, int C, int K, int R, int relu, int block_n,
                                 cudaStream_t stream) {
  return conv_box_tma(x, w, ab_dtype, scale, shift, y, y_dtype, N, H, W, C, K, R, 2, relu, block_n, stream);
}

// Debug: device buffer of 5 * 768 * 2 int64 that CTA 0 of every later conv_tma launch fills with
// (event, clock64) records per role (producer / MMA issuer / epilogue); NULL switches tracing off.
extern "C" int vlnce_debug_set_trace(void* buffer) {
  g_trace = reinterpret_cast<long long*>(buffer);
  return VLNCE_OK;
}

// (N,H,W,3) u8|f32 image -> (N, P+3, Q+3, 16) 16-bit space-to-depth cells (see stem_s2d_kernel);
// P = (H-1)/2+1, Q = (W-1)/2+1 are the stem's output size.
extern "C" int vlnce_stem_s2d(const void* x, int x_dtype, void* out, int out_dtype, int N, int H, int W,
                              cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && out && N > 0 && H > 0 && W > 0, "stem_s2d: null tensor / empty shape");
  VLNCE_CHECK_ARG(x_dtype == VLNCE_U8 || x_dtype == VLNCE_F32, "stem_s2d: image must be u8 or f32");
  VLNCE_CHECK_ARG(out_dtype == VLNCE_BF16 || out_dtype == VLNCE_F16, "stem_s2d: output must be bf16 or f16");
  VLNCE_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0, "stem_s2d: output must be 16-byte aligned");
  const int P = (H - 1) / 2 + 1, Q = (W - 1) / 2 + 1;
  const int Ys = P + 3, Xs = Q + 3;
  const long long cells = (long long)N * Ys * Xs;
  const int grid = (int)((cells + 255) / 256 < 148 * 16 ? (cells + 255) / 256 : 148 * 16);
#define S2D_CASE(TI, TO) \
  stem_s2d_kernel<TI, TO><<<grid, 256, 0, stream>>>((const TI*)x, (TO*)out, N, H, W, Ys, Xs)
  if (x_dtype == VLNCE_U8 && out_dtype == VLNCE_BF16) S2D_CASE(uint8_t, __nv_bfloat16);
  else if (x_dtype == VLNCE_U8) S2D_CASE(uint8_t, __half);
  else if (out_dtype == VLNCE_BF16) S2D_CASE(float, __nv_bfloat16);
  else S2D_CASE(float, __half);
#undef S2D_CASE
  VLNCE_LAUNCH_CHECK("stem_s2d_kernel");
  return VLNCE_OK;
}

// RGB stem on the s2d operand: xs (N, P+3, Q+3, 16), w (K, 256) = [k][r'][s'][16 cell channels],
// y (N, P, Q, K) ab_dtype | f32.  K % 64 == 0.
extern "C" int vlnce_conv_stem_tma(const void* xs, const void* w, int ab_dtype, const float* scale,
                                   const float* shift, void* y, int y_dtype, int N, int P, int Q, int K, int relu,
                                   int block_n, cudaStream_t stream) {
  VLNCE_CHECK_ARG(xs && w && y && N > 0 && P > 0 && Q > 0, "conv_stem_tma: null tensor / empty shape");
  VLNCE_CHECK_ARG(K % 64 == 0, "conv_stem_tma: K must be a multiple of 64");
  VLNCE_CHECK_ARG((scale == nullptr) == (shift == nullptr), "conv_stem_tma: scale/shift must come together");
  VLNCE_CHECK_ARG(ab_dtype == VLNCE_BF16 || ab_dtype == VLNCE_F16, "conv_stem_tma: operand dtype");
  VLNCE_CHECK_ARG(y_dtype == ab_dtype || y_dtype == VLNCE_F32, "conv_stem_tma: y dtype");
  const int b_stages = (block_n >> 16) & 0xf;
  const int mt = (block_n & VLNCE_TMA_MT2) ? 2 : 1;
  const int no_stationary = (block_n & VLNCE_TMA_NO_STATIONARY) ? 1 : 0;
  int bn = block_n & 0xffff;
  if (bn == 0) bn = 64;
  VLNCE_CHECK_ARG((bn == 64 || bn == 128 || bn == 256) && K % bn == 0, "conv_stem_tma: block_n=%d vs K=%d", bn, K);
  EncodeTiledFn enc = get_enc();
  if (!enc) {
    set_error("conv_stem_tma: cuTensorMapEncodeTiled unavailable");
    return VLNCE_ERR_NO_DRIVER;
  }
  const CUtensorMapDataType dt16 = ab_dtype == VLNCE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const bool f32 = y_dtype == VLNCE_F32;
  const int Ys = P + 3, Xs = Q + 3;
  CUtensorMap tx, tw, ty;
  int rc;
  {  // weights [K, 256]
    const cuuint64_t gd[2] = {256, (cuuint64_t)K};
    const cuuint64_t gs[1] = {512};
    const cuuint32_t bx[2] = {64, (cuuint32_t)bn};
    if ((rc = map_nd(enc, &tw, w, dt16, 2, gd, gs, bx, "stem weights")) != VLNCE_OK) return rc;
  }
  {  // overlapping 4-cell windows: dim0 = 64 elements from cell x, dim1 = x with a stride of ONE cell (32 B)
    const cuuint64_t gd[4] = {64, (cuuint64_t)Q, (cuuint64_t)Ys, (cuuint64_t)N};
    const cuuint64_t gs[3] = {32, (cuuint64_t)Xs * 32, (cuuint64_t)Ys * Xs * 32};
    const cuuint32_t bx[4] = {64, 128, 1, 1};
    if ((rc = map_nd(enc, &tx, xs, dt16, 4, gd, gs, bx, "s2d windows")) != VLNCE_OK) return rc;
  }
  {
    const int esz = f32 ? 4 : 2;
    const cuuint64_t yd[4] = {(cuuint64_t)K, (cuuint64_t)Q, (cuuint64_t)P, (cuuint64_t)N};
    const cuuint64_t ys[3] = {(cuuint64_t)K * esz, (cuuint64_t)Q * K * esz, (cuuint64_t)P * Q * K * esz};
    const cuuint32_t yb[4] = {(cuuint32_t)(f32 ? 32 : 64), 128, 1, 1};
    if ((rc = map_nd(enc, &ty, y, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : dt16, 4, yd, ys, yb, "stem output")) != VLNCE_OK)
      return rc;
  }
  TmaConvParams p;
  p.mode = 2; p.taps = 1; p.halo_w = 8; p.a_bytes = 16 * 1024; p.a_tx = 16 * 1024; p.img_per_tile = 1; p.cpt = 1; p.rcp_cpt = 1.f; p.taps_s2 = 1; p.box_stride = 1;
  p.tw = 128; p.th = 1;
  p.tiles_x = cdiv(Q, 128); p.tiles_y = P; p.m_tiles = N * p.tiles_x * p.tiles_y;
  p.scale = scale; p.shift = shift;
  p.M = N * P * Q; p.N = N; p.H = P; p.W = Q; p.C = 256; p.K = K;
  p.c_chunks = 4;
  p.tf32 = 0; p.kc = 64;
  p.w_cols = 0;
  p.n_tiles = K / bn;
  p.relu = relu; p.y_fp32 = f32; p.ab_half = (ab_dtype == VLNCE_F16); p.has_res = 0;
  p.a_stages = p.b_stages = p.stg_bufs = 0; p.b_stationary = 0;
  return dispatch(bn, mt, tx, tw, ty, ty, p, b_stages, no_stationary, stream);
}
