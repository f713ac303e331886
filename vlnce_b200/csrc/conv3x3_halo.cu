// 3x3 stride-1 pad-1 NHWC convolution on tcgen05 with TMA *halo* loads: im2col-free AND
// re-read-free.  One CTA computes a 16 x 8 pixel patch (128 GEMM rows) x BLOCK_N filters.
// Per 64-channel chunk ONE 4-D TMA box brings the (16+2) x (8+2) pixel halo of the patch into
// shared memory (out-of-image pixels arrive as zeros = the conv padding); the nine filter taps
// are nine UMMA descriptors into that same tile: tap (r, s) starts (r*10 + s) pixel rows in,
// every 8-pixel output row segment is one 8-row core group, consecutive output rows are one halo
// row (10 x 128 B) apart = the descriptor's stride-byte-offset.  The 128B swizzle is a function of
// the shared-memory address bits, so TMA's write pattern and the shifted reads agree.
// Operand bytes per MAC drop ~3-4x against the tap-by-tap gather (conv_umma.cu), which is what
// bounds that kernel (latency of operand delivery, see profiles/r01_conv_umma_ncu_full.md).
//
// Warp roles: warp 5 = TMA producer (halo ring + weight ring), warp 4 = MMA issuer + TMEM owner,
// warps 0-3 = epilogue (tcgen05.ld -> scale/shift (+residual) (+ReLU) -> swizzled smem -> TMA store
// of the 16 x 8 x BLOCK_N patch through a 4-D tensor map).
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace vlnce {
namespace {

using namespace ptx;

constexpr int TH = 16, TW = 8;                 // output patch
constexpr int HALO_W = TW + 2, HALO_H = TH + 2;
constexpr int HALO_ROWS = HALO_W * HALO_H;     // 180 pixel rows of 128 B
constexpr int A_BYTES = 23 * 1024;             // 180 * 128 = 23040, rounded to keep 1024 B alignment
constexpr int A_TX_BYTES = HALO_ROWS * 128;
constexpr int kAStages = 2;
constexpr int kThreads = 192;
constexpr int kSmemBudget = 227 * 1024;

struct HaloParams {
  const float* scale;
  const float* shift;
  int N, H, W, C, K;
  int c_chunks;   // C / 64
  int n_tiles;    // K / BLOCK_N
  int tiles_x, tiles_y;
  int relu, y_fp32, ab_half, has_res;
  int b_stages;
};

__device__ __forceinline__ void tma_store_4d(const void* tmap, uint32_t src_smem, int32_t c0, int32_t c1, int32_t c2,
                                             int32_t c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   (uint64_t)tmap),
               "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// K-major SW128 descriptor with an explicit stride-byte-offset between 8-row groups.
__device__ __forceinline__ uint64_t make_desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 2)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ CUtensorMap tmap_y, const __grid_constant__ CUtensorMap tmap_res,
                    const HaloParams p) {
  constexpr int B_BYTES = BLOCK_N * 128;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int bst = p.b_stages;
  const uint32_t smem_a0 = smem_u32(smem);
  const uint32_t smem_b0 = smem_a0 + kAStages * A_BYTES;
  uint8_t* bar_base = smem + kAStages * A_BYTES + bst * B_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* a_empty = a_full + kAStages;
  uint64_t* b_full = a_empty + kAStages;
  uint64_t* b_empty = b_full + 8;
  uint64_t* accum_bar = b_empty + 8;
  uint64_t* res_bar = accum_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 1);
  const uint32_t stage_out = smem_b0 + bst * B_BYTES + 1024;  // after the barrier block, 1024 B aligned

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x % p.n_tiles;
  int t = blockIdx.x / p.n_tiles;
  const int tx = t % p.tiles_x; t /= p.tiles_x;
  const int ty = t % p.tiles_y;
  const int img = t / p.tiles_y;
  const int x0 = tx * TW, y0 = ty * TH, n0 = n_tile * BLOCK_N;

  if (threadIdx.x == 5 * 32) {
    for (int s = 0; s < kAStages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < bst; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(accum_bar, 1);
    mbar_init(res_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_y);
  }
  if (warp == 4) {
    tmem_alloc(tmem_slot, BLOCK_N);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 5) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      if (p.has_res) {
        mbar_arrive_expect_tx(res_bar, 128 * BLOCK_N * 2);
        for (int b = 0; b < BLOCK_N / 64; ++b)
          tma_load_4d(stage_out + (uint32_t)b * (128 * 128), &tmap_res, res_bar, n0 + b * 64, x0, y0, img);
      }
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      auto load_halo = [&](int cc) {
        mbar_wait(&a_empty[as], aph ^ 1, 11);
        mbar_arrive_expect_tx(&a_full[as], A_TX_BYTES);
        tma_load_4d(smem_a0 + as * A_BYTES, &tmap_x, &a_full[as], cc * 64, x0 - 1, y0 - 1, img);
        if (++as == kAStages) { as = 0; aph ^= 1; }
      };
      load_halo(0);
      for (int cc = 0; cc < p.c_chunks; ++cc) {
        for (int tap = 0; tap < 9; ++tap) {
          // the next chunk's halo is requested half a chunk ahead (its slot was freed by chunk cc-1)
          if (tap == 4 && cc + 1 < p.c_chunks) load_halo(cc + 1);
          mbar_wait(&b_empty[bs], bph ^ 1, 12);
          mbar_arrive_expect_tx(&b_full[bs], B_BYTES);
          tma_load_2d(smem_b0 + bs * B_BYTES, &tmap_w, &b_full[bs], tap * p.C + cc * 64, n0);
          if (++bs == bst) { bs = 0; bph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 4) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_f16(128, BLOCK_N, p.ab_half ? 0u : 1u);
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    for (int cc = 0; cc < p.c_chunks; ++cc) {
      mbar_wait(&a_full[as], aph, 13);
      for (int tap = 0; tap < 9; ++tap) {
        mbar_wait(&b_full[bs], bph, 14);
        tc_fence_after_sync();
        if (lane == 0) {
          const int r = tap / 3, s = tap - 3 * r;
          const uint64_t adesc = make_desc_sbo(smem_a0 + as * A_BYTES + (r * HALO_W + s) * 128, HALO_W * 128);
          const uint64_t bdesc = make_smem_desc(smem_b0 + bs * B_BYTES, 128);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)((cc | tap | k) != 0));
          umma_commit(&b_empty[bs]);
          if (tap == 8) umma_commit(&a_empty[as]);
        }
        __syncwarp();
        if (++bs == bst) { bs = 0; bph ^= 1; }
      }
      if (++as == kAStages) { as = 0; aph ^= 1; }
    }
    if (lane == 0) umma_commit(accum_bar);
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue (warps 0-3)
    mbar_wait(accum_bar, 0, 15);
    tc_fence_after_sync();
    const int row = warp * 32 + lane;  // = y_local * 8 + x_local
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    if (p.has_res) mbar_wait(res_bar, 0, 16);
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
      if (p.y_fp32) {
        const uint32_t box = stage_out + (uint32_t)cb * (128 * 128) + (uint32_t)row * 128;
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const uint32_t addr = box + (uint32_t)((q4 ^ (row & 7)) << 4);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(f[q4 * 4]), "f"(f[q4 * 4 + 1]),
                       "f"(f[q4 * 4 + 2]), "f"(f[q4 * 4 + 3])
                       : "memory");
        }
      } else {
        const uint32_t box = stage_out + (uint32_t)(cb >> 1) * (128 * 128) + (uint32_t)row * 128;
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
            if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
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
    fence_proxy_async_smem();
    named_barrier_sync(1, 128);
    if (threadIdx.x == 0) {
      const int boxes = p.y_fp32 ? BLOCK_N / 32 : BLOCK_N / 64;
      const int cols = p.y_fp32 ? 32 : 64;
      for (int b = 0; b < boxes; ++b)
        tma_store_4d(&tmap_y, stage_out + (uint32_t)b * (128 * 128), n0 + b * cols, x0, y0, img);
      tma_store_commit();
      tma_store_wait_read();
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, BLOCK_N);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
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

// NHWC tensor (N, H, W, C) as a 4-D map (C, W, H, N); box = (box_c, box_w, box_h, 1)
int make_map_nhwc(EncodeTiledFn enc, CUtensorMap* m, const void* base, CUtensorMapDataType dt, int esz, int N, int H,
                  int W, int C, int box_c, int box_w, int box_h, const char* what) {
  const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t gstr[3] = {(cuuint64_t)C * esz, (cuuint64_t)W * C * esz, (cuuint64_t)H * W * C * esz};
  const cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, dt, 4, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("conv3x3_halo: cuTensorMapEncodeTiled(%s) failed (%d)", what, (int)r);
    return VLNCE_ERR_CUDA_BASE + 1;
  }
  return VLNCE_OK;
}

template <int BLOCK_N>
int launch(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& ty, const CUtensorMap& tr, HaloParams& p,
           int b_stages, int exclusive, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_halo_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kSmemBudget);
    if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(conv3x3_halo)");
    configured = true;
  }
  const int staging = 128 * BLOCK_N * (p.y_fp32 ? 4 : 2);
  auto bytes = [&](int bs) { return kAStages * A_BYTES + bs * BLOCK_N * 128 + 1024 + staging + 1024; };
  int bs = b_stages > 0 ? b_stages : 4;
  if (bs > 8) bs = 8;
  while (bs > 2 && bytes(bs) > kSmemBudget) --bs;
  if (bytes(bs) > kSmemBudget) {
    set_error("conv3x3_halo: tile does not fit shared memory (BLOCK_N=%d)", BLOCK_N);
    return VLNCE_ERR_INVALID_ARGUMENT;
  }
  p.b_stages = bs;
  const long long tiles = (long long)p.N * p.tiles_x * p.tiles_y * p.n_tiles;
  int smem = bytes(bs);
  if (exclusive && smem < 116 * 1024) smem = 116 * 1024;  // one CTA per SM, rest of the SM left to other branches
  conv3x3_halo_kernel<BLOCK_N><<<(unsigned)tiles, kThreads, smem, stream>>>(tx, tw, ty, tr, p);
  VLNCE_LAUNCH_CHECK("conv3x3_halo_kernel");
  return VLNCE_OK;
}

}  // namespace
}  // namespace vlnce

using namespace vlnce;

extern "C" int vlnce_conv3x3_halo(const void* x, const void* w, int ab_dtype, const float* scale, const float* shift,
                                  const void* residual, void* y, int y_dtype, int N, int H, int W, int C, int K,
                                  int relu, int block_n, cudaStream_t stream) {
  VLNCE_CHECK_ARG(x && w && y && N > 0, "conv3x3_halo: null tensor");
  VLNCE_CHECK_ARG(H % TH == 0 && W % TW == 0, "conv3x3_halo: H %% 16 and W %% 8 must be 0 (got %dx%d)", H, W);
  VLNCE_CHECK_ARG(C % 64 == 0 && K % 64 == 0, "conv3x3_halo: C and K must be multiples of 64");
  VLNCE_CHECK_ARG((scale == nullptr) == (shift == nullptr), "conv3x3_halo: scale/shift must come together");
  VLNCE_CHECK_ARG(ab_dtype == VLNCE_BF16 || ab_dtype == VLNCE_F16, "conv3x3_halo: operand dtype");
  VLNCE_CHECK_ARG(y_dtype == ab_dtype || y_dtype == VLNCE_F32, "conv3x3_halo: y dtype");
  VLNCE_CHECK_ARG(!(residual && y_dtype == VLNCE_F32), "conv3x3_halo: residual needs a 16-bit output");
  const int b_stages = (block_n >> 16) & 0xf;
  const int exclusive = (block_n >> 20) & 1;
  int bn = block_n & 0xffff;
  if (bn == 0) bn = (K % 256 == 0) ? 256 : (K % 128 == 0 ? 128 : 64);
  VLNCE_CHECK_ARG((bn == 64 || bn == 128 || bn == 256) && K % bn == 0, "conv3x3_halo: block_n=%d vs K=%d", bn, K);
  // fp32 staging of a 128 x 256 tile (128 KB) does not fit beside the rings: narrow the tile
  while (bn > 64 && kAStages * A_BYTES + 2 * bn * 128 + 2048 + 128 * bn * (y_dtype == VLNCE_F32 ? 4 : 2) > kSmemBudget)
    bn /= 2;
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("conv3x3_halo: cuTensorMapEncodeTiled unavailable");
    return VLNCE_ERR_NO_DRIVER;
  }
  const CUtensorMapDataType dt16 = ab_dtype == VLNCE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tx, tw, ty, tr;
  int rc = make_map_nhwc(enc, &tx, x, dt16, 2, N, H, W, C, 64, HALO_W, HALO_H, "halo");
  if (rc != VLNCE_OK) return rc;
  {  // weights [K, 9*C] row-major, box 64 x bn
    const cuuint64_t gdim[2] = {(cuuint64_t)9 * C, (cuuint64_t)K};
    const cuuint64_t gstr[1] = {(cuuint64_t)9 * C * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)bn};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tw, dt16, 2, const_cast<void*>(w), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv3x3_halo: cuTensorMapEncodeTiled(weights) failed (%d)", (int)r);
      return VLNCE_ERR_CUDA_BASE + 1;
    }
  }
  if (y_dtype == VLNCE_F32)
    rc = make_map_nhwc(enc, &ty, y, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, N, H, W, K, 32, TW, TH, "output");
  else
    rc = make_map_nhwc(enc, &ty, y, dt16, 2, N, H, W, K, 64, TW, TH, "output");
  if (rc != VLNCE_OK) return rc;
  tr = ty;
  if (residual) {
    rc = make_map_nhwc(enc, &tr, residual, dt16, 2, N, H, W, K, 64, TW, TH, "residual");
    if (rc != VLNCE_OK) return rc;
  }
  HaloParams p;
  p.scale = scale; p.shift = shift;
  p.N = N; p.H = H; p.W = W; p.C = C; p.K = K;
  p.c_chunks = C / 64;
  p.n_tiles = K / bn;
  p.tiles_x = W / TW; p.tiles_y = H / TH;
  p.relu = relu; p.y_fp32 = (y_dtype == VLNCE_F32); p.ab_half = (ab_dtype == VLNCE_F16);
  p.has_res = residual != nullptr;
  p.b_stages = 0;
  if (bn == 256) return launch<256>(tx, tw, ty, tr, p, b_stages, exclusive, stream);
  if (bn == 128) return launch<128>(tx, tw, ty, tr, p, b_stages, exclusive, stream);
  return launch<64>(tx, tw, ty, tr, p, b_stages, exclusive, stream);
}
