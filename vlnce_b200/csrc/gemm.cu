// fp32 strided GEMM on CUDA cores for the (small) trainable head:
//   C[m,n] = act( sum_k A[m*lda_m + k*lda_k] * B[k*ldb_k + n*ldb_n] + bias[n] + beta*C[m,n] )
// Arbitrary strides on both operands cover every forward/backward variant of the head's
// Linear / Conv1d(k=1) layers (x W^T, dY W, dY^T X) without transposed copies.  64x64x16
// tiles, 256 threads, 4x4 register micro-tiles, true fp32 (this is the 1e-3-parity path).
#include "common.cuh"

#include <stdlib.h>
#include <string.h>

namespace vlnce {
namespace {

constexpr int TM = 64, TN = 64, TK = 32;
constexpr int kMaxGroup = 24;          // problems per grouped launch (descriptor block stays < 4 KB of kernel parameters)
constexpr int kCounterBytes = 16384;   // head of the workspace: one arrival counter per split tile
constexpr int kMaxCounters = kCounterBytes / 4;

// One GEMM of a grouped launch.  CTAs [cta_begin, cta_begin + tiles * splits) of the 1-D grid belong to it:
// cta -> (tile, split) with the split index fastest, tile -> (tile_m, tile_n) with tile_n fastest.
struct Prob {
  const float* A;
  const float* B;   // NULL: B == 1 (column sums of A over k — the bias gradients)
  float* C;
  const float* bias;
  long long lda_m, lda_k, ldb_k, ldb_n, ldc;
  int M, N, K;
  float beta;
  int act;
  int tiles_n, tiles;
  int splits, k_per;
  int cta_begin;
  int ws_off;    // float offset of this problem's split-K partials behind the counters
  int ctr_off;   // first arrival counter
  int red_begin; // two-pass mode: first CTA of the reduction launch that belongs to this problem
};
// The descriptor block travels as a kernel parameter; its size is a launch cost (measured: ~1 us per launch for the
// 3 KB block of 24 problems in CUDA-graph replay), so single GEMMs and small groups use smaller blocks.
template <int NP>
struct GroupT {
  Prob p[NP];
  int count;
  int two_pass;  // 1: split-K partials are summed by a second (grouped) launch instead of the last CTA of a tile
};
using Group = GroupT<kMaxGroup>;

// fp32 64x64x32 tile; split-K partials are reduced IN THE SAME LAUNCH by the last CTA to arrive at the tile (fixed
// summation order z = 0..splits-1, so the result does not depend on the arrival order: deterministic).
template <int NP>
__global__ void __launch_bounds__(256)
sgemm_grouped_kernel(const __grid_constant__ GroupT<NP> g, float* __restrict__ ws, unsigned* __restrict__ counters) {
  pdl_trigger();  // PDL (common.cuh)
  pdl_wait();
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  __shared__ int s_last;
  const int tid = threadIdx.x;
  int pi = 0;
#pragma unroll 1
  while (pi + 1 < g.count && (int)blockIdx.x >= g.p[pi + 1].cta_begin) ++pi;
  const Prob& P = g.p[pi];
  const int local = (int)blockIdx.x - P.cta_begin;
  const int z = local % P.splits, tile = local / P.splits;
  const int m0 = (tile / P.tiles_n) * TM, n0 = (tile % P.tiles_n) * TN;
  const float* __restrict__ A = P.A;
  const float* __restrict__ B = P.B;
  const long long lda_m = P.lda_m, lda_k = P.lda_k, ldb_k = P.ldb_k, ldb_n = P.ldb_n;
  const int M = P.M, N = P.N, K = P.K;
  const int tx = tid % 16, ty = tid / 16;  // micro-tile coords: rows ty*4.., cols tx*4..
  float acc[4][4] = {};
  // load mappings: favour the contiguous dimension of each operand
  const bool a_kfast = (lda_k == 1);
  const bool b_nfast = (ldb_n == 1);
  const int kb = z * P.k_per;
  const int ke = min(K, kb + P.k_per);
  constexpr int LD = TM * TK / 256;  // 8 elements of each operand tile per thread
  float ra[LD], rb[LD];
  // global -> registers for the k-tile starting at k0 (issued one tile ahead of its use, so the
  // L2 latency of tile i+1 hides behind the FMAs of tile i)
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < LD; ++i) {
      const int e = tid + i * 256;
      int mm, kk;
      if (a_kfast) { kk = e % TK; mm = e / TK; } else { mm = e % TM; kk = e / TM; }
      const int gm = m0 + mm, gk = k0 + kk;
      ra[i] = (gm < M && gk < ke) ? __ldg(A + gm * lda_m + gk * lda_k) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < LD; ++i) {
      const int e = tid + i * 256;
      int nn, kk;
      if (b_nfast) { nn = e % TN; kk = e / TN; } else { kk = e % TK; nn = e / TK; }
      const int gn = n0 + nn, gk = k0 + kk;
      rb[i] = (gn < N && gk < ke) ? (B != nullptr ? __ldg(B + gk * ldb_k + gn * ldb_n) : 1.f) : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < LD; ++i) {
      const int e = tid + i * 256;
      int mm, kk;
      if (a_kfast) { kk = e % TK; mm = e / TK; } else { mm = e % TM; kk = e / TM; }
      As[kk][mm] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < LD; ++i) {
      const int e = tid + i * 256;
      int nn, kk;
      if (b_nfast) { nn = e % TN; kk = e / TN; } else { kk = e % TK; nn = e / TK; }
      Bs[kk][nn] = rb[i];
    }
  };
  if (kb < ke) fetch(kb);
  for (int k0 = kb; k0 < ke; k0 += TK) {
    stash();
    __syncthreads();
    if (k0 + TK < ke) fetch(k0 + TK);
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int splits = P.splits;
  if (splits > 1) {
    // (1) park this split's partial tile, (2) count the arrival, (3) the last CTA of the tile sums all partials
    // (z = 0..splits-1 in order for every element; 16-byte accesses and four splits in flight per thread, so the
    // one SM that reduces a tile pulls the partials at L2 bandwidth instead of one load latency per split)
    float* part = ws + P.ws_off;
    const int gn0 = n0 + tx * 4;
    const bool vec = (N & 3) == 0;   // (ws_off and the row offsets are multiples of 4 floats then)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gm = m0 + ty * 4 + i;
      if (gm >= M) continue;
      float* dst = part + ((size_t)z * M + gm) * N + gn0;
      if (vec) {
        if (gn0 < N) __stcg(reinterpret_cast<float4*>(dst), make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (gn0 + j < N) __stcg(dst + j, acc[i][j]);
      }
    }
    if (g.two_pass) return;   // grouped_reduce_kernel sums the partials
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const unsigned prev = atomicAdd(counters + P.ctr_off + tile, 1u);
      s_last = (prev == (unsigned)(splits - 1));
      if (s_last) counters[P.ctr_off + tile] = 0u;  // every split has arrived: re-arm for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    if (vec) {
      if (gn0 < N) {
        for (int zz = 0; zz < splits; zz += 4) {
          float4 v[4][4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int gm = m0 + ty * 4 + i;
              v[u][i] = (zz + u < splits && gm < M)
                            ? __ldcg(reinterpret_cast<const float4*>(part + ((size_t)(zz + u) * M + gm) * N + gn0))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              acc[i][0] += v[u][i].x; acc[i][1] += v[u][i].y; acc[i][2] += v[u][i].z; acc[i][3] += v[u][i].w;
            }
        }
      }
    } else {
      for (int zz = 0; zz < splits; ++zz) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int gm = m0 + ty * 4 + i;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (gm < M && gn0 + j < N) acc[i][j] += __ldcg(part + ((size_t)zz * M + gm) * N + gn0 + j);
        }
      }
    }
  }
  const float* __restrict__ bias = P.bias;
  float* __restrict__ C = P.C;
  const float beta = P.beta;
  const int act = P.act;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias != nullptr) v += bias[gn];
      float* cp = C + gm * P.ldc + gn;
      if (beta != 0.f) v += beta * (*cp);
      if (act == 1) v = fmaxf(v, 0.f);
      else if (act == 2) v = tanhf(v);
      *cp = v;
    }
  }
}

// Second pass of a grouped split-K launch: C = act(sum_z partial[z] + bias + beta*C) for every split problem of the
// group, one output element per thread, fixed summation order (deterministic).
template <int NP>
__global__ void __launch_bounds__(256)
grouped_reduce_kernel(const __grid_constant__ GroupT<NP> g, const float* __restrict__ ws) {
  pdl_trigger();  // PDL (common.cuh)
  pdl_wait();
  int pi = -1;
#pragma unroll 1
  for (int i = 0; i < g.count; ++i)
    if (g.p[i].splits > 1 && (int)blockIdx.x >= g.p[i].red_begin) pi = i;
  if (pi < 0) return;
  const Prob& P = g.p[pi];
  const int total = P.M * P.N;
  const float* part = ws + P.ws_off;
  const int idx = ((int)blockIdx.x - P.red_begin) * 256 + (int)threadIdx.x;   // one output element per thread
  if (idx >= total) return;
  const int m = idx / P.N, n = idx - m * P.N;
  float v = 0.f;
  for (int z = 0; z < P.splits; ++z) v += __ldcg(part + (size_t)z * total + idx);
  if (P.bias != nullptr) v += P.bias[n];
  float* cp = P.C + m * P.ldc + n;
  if (P.beta != 0.f) v += P.beta * (*cp);
  if (P.act == 1) v = fmaxf(v, 0.f);
  else if (P.act == 2) v = tanhf(v);
  *cp = v;
}

// out[n] = (accumulate ? out[n] : 0) + sum_m A[m*lda + n]   (bias gradients).  grid (N/32, S):
// each CTA sums a row slice of 32 columns (8 row lanes x coalesced 128B reads); S > 1 writes
// partial[S][N] for the fixed-order second stage below.
__global__ void colsum_kernel(const float* __restrict__ A, long long lda, float* __restrict__ out,
                              int M, int N, int accumulate, float* __restrict__ partial) {
  pdl_trigger();  // PDL (common.cuh)
  pdl_wait();
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int slice = threadIdx.x >> 5, nslices = blockDim.x >> 5;
  const int S = gridDim.y;
  const int rows_per = (M + S - 1) / S;
  const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  __shared__ float part[8][33];
  float s = 0.f;
  if (n < N)
    for (int m = r0 + slice; m < r1; m += nslices) s += A[m * lda + n];
  part[slice][threadIdx.x & 31] = s;
  __syncthreads();
  if (slice == 0 && n < N) {
    float t = 0.f;
    for (int i = 0; i < nslices; ++i) t += part[i][threadIdx.x & 31];
    if (partial != nullptr) partial[(size_t)blockIdx.y * N + n] = t;
    else out[n] = accumulate ? out[n] + t : t;
  }
}

__global__ void colsum_final_kernel(const float* __restrict__ partial, int S, float* __restrict__ out, int N,
                                    int accumulate) {
  pdl_trigger();  // PDL (common.cuh)
  pdl_wait();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float t = 0.f;
  for (int s = 0; s < S; ++s) t += partial[(size_t)s * N + n];
  out[n] = accumulate ? out[n] + t : t;
}

}  // namespace
}  // namespace vlnce

using namespace vlnce;

// Host side of a grouped launch: decomposition (tiles, split-K factors, CTA ranges, workspace slices) of up to
// kMaxGroup problems, then ONE launch.  Split-K: skinny problems (M = 64 policy rows) leave most of the 148 SMs
// idle, so K is split until the whole group covers ~2 waves.
template <int NP>
static int launch_sized(const Group& g, int ctas, int red_ctas, float* part, unsigned* counters, cudaStream_t stream) {
  GroupT<NP> s;
  for (int i = 0; i < g.count; ++i) s.p[i] = g.p[i];
  s.count = g.count;
  s.two_pass = g.two_pass;
  launch_pdl(sgemm_grouped_kernel<NP>, dim3(ctas), dim3(256), 0, stream, s, part, counters);
  VLNCE_LAUNCH_CHECK("sgemm_grouped_kernel");
  if (g.two_pass && red_ctas > 0) {
    launch_pdl(grouped_reduce_kernel<NP>, dim3(red_ctas), dim3(256), 0, stream, s, (const float*)part);
    VLNCE_LAUNCH_CHECK("grouped_reduce_kernel");
  }
  return VLNCE_OK;
}

static int launch_group(Group& g, float* workspace, size_t workspace_bytes, cudaStream_t stream) {
  // VLNCE_SPLITK=inkernel: partials summed by the last CTA of each tile (one launch); default: a second grouped launch
  // sums them with the whole GPU (a tile's 16-32 partials are ~0.5 MB for ONE SM to pull otherwise)
  static int two_pass = -1;
  if (two_pass < 0) {
    const char* e = getenv("VLNCE_SPLITK");
    two_pass = (e != nullptr && strcmp(e, "inkernel") == 0) ? 0 : 1;
  }
  g.two_pass = two_pass;
  const int max_splits = two_pass ? 32 : 16;
  const bool can_split = workspace != nullptr && workspace_bytes > (size_t)kCounterBytes;
  const size_t ws_floats = can_split ? (workspace_bytes - kCounterBytes) / sizeof(float) : 0;
  size_t ws_used = 0;
  int ctr_used = 0, ctas = 0, red_ctas = 0;
  // CTAs a launch may spread its split-K over: ~2 waves of the 148 SMs finishes a lone GEMM soonest; inside the pipelined
  // step (every SM already busy with the trunks) a smaller budget trades latency for less partial-sum traffic
  // (VLNCE_GEMM_BUDGET; measured in DESIGN 6)
  static int total_budget = -1;
  if (total_budget < 0) {
    const char* e = getenv("VLNCE_GEMM_BUDGET");
    total_budget = (e != nullptr && atoi(e) > 0) ? atoi(e) : 296;
  }
  const int budget = total_budget / g.count > 8 ? total_budget / g.count : 8;
  for (int i = 0; i < g.count; ++i) {
    Prob& P = g.p[i];
    P.tiles_n = cdiv(P.N, TN);
    P.tiles = P.tiles_n * cdiv(P.M, TM);
    int splits = 1;
    if (can_split && P.tiles < 96 && P.K >= 128) {
      splits = min(min(budget / P.tiles, P.K / 64), max_splits);
      if (splits < 1) splits = 1;
      while (splits > 1 && (ws_used + (size_t)splits * P.M * P.N > ws_floats || ctr_used + P.tiles > kMaxCounters)) --splits;
    }
    int k_per = P.K;
    if (splits > 1) {
      k_per = cdiv(cdiv(P.K, splits), TK) * TK;
      splits = cdiv(P.K, k_per);
    }
    P.splits = splits;
    P.k_per = k_per;
    P.cta_begin = ctas;
    P.ws_off = (int)ws_used;
    P.ctr_off = ctr_used;
    P.red_begin = red_ctas;
    if (splits > 1) red_ctas += cdiv(P.M * P.N, 256);
    if (splits > 1) {
      ws_used += ((size_t)splits * P.M * P.N + 3) & ~(size_t)3;
      ctr_used += P.tiles;
    }
    ctas += P.tiles * splits;
  }
  float* part = can_split ? reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + kCounterBytes) : nullptr;
  unsigned* counters = reinterpret_cast<unsigned*>(workspace);
  if (g.count == 1) return launch_sized<1>(g, ctas, red_ctas, part, counters, stream);
  if (g.count <= 4) return launch_sized<4>(g, ctas, red_ctas, part, counters, stream);
  if (g.count <= 8) return launch_sized<8>(g, ctas, red_ctas, part, counters, stream);
  return launch_sized<kMaxGroup>(g, ctas, red_ctas, part, counters, stream);
}

extern "C" int vlnce_sgemm(const float* A, long long lda_m, long long lda_k, const float* B,
                           long long ldb_k, long long ldb_n, float* C, long long ldc,
                           const float* bias, int M, int N, int K, float beta, int act,
                           float* workspace, size_t workspace_bytes, cudaStream_t stream) {
  VLNCE_CHECK_ARG(A && B && C, "sgemm: null operand");
  VLNCE_CHECK_ARG(M > 0 && N > 0 && K > 0, "sgemm: empty problem %dx%dx%d", M, N, K);
  VLNCE_CHECK_ARG(act >= 0 && act <= 2, "sgemm: act must be 0 (none), 1 (relu) or 2 (tanh)");
  Group gg;  // host staging only; the launch copies it by value
  gg.count = 1;
  Prob& P = gg.p[0];
  P.A = A; P.B = B; P.C = C; P.bias = bias;
  P.lda_m = lda_m; P.lda_k = lda_k; P.ldb_k = ldb_k; P.ldb_n = ldb_n; P.ldc = ldc;
  P.M = M; P.N = N; P.K = K; P.beta = beta; P.act = act;
  return launch_group(gg, workspace, workspace_bytes, stream);
}

// descs: count x 16 64-bit words {A, lda_m, lda_k, B, ldb_k, ldb_n, C, ldc, bias, M, N, K, beta (IEEE double), act,
// 0, 0} (include/vlnce_b200.h: vlnce_gemm_desc).  B == NULL: the problem is a column sum of A over k.
extern "C" int vlnce_sgemm_grouped(const long long* descs, int count, float* workspace, size_t workspace_bytes,
                                   cudaStream_t stream) {
  VLNCE_CHECK_ARG(descs != nullptr && count > 0, "sgemm_grouped: no problems");
  Group g;
  for (int base = 0; base < count; base += kMaxGroup) {
    g.count = min(kMaxGroup, count - base);
    for (int i = 0; i < g.count; ++i) {
      const long long* d = descs + (size_t)(base + i) * 16;
      Prob& P = g.p[i];
      P.A = reinterpret_cast<const float*>(d[0]); P.lda_m = d[1]; P.lda_k = d[2];
      P.B = reinterpret_cast<const float*>(d[3]); P.ldb_k = d[4]; P.ldb_n = d[5];
      P.C = reinterpret_cast<float*>(d[6]); P.ldc = d[7];
      P.bias = reinterpret_cast<const float*>(d[8]);
      P.M = (int)d[9]; P.N = (int)d[10]; P.K = (int)d[11];
      double beta;
      memcpy(&beta, &d[12], sizeof(double));
      P.beta = (float)beta;
      P.act = (int)d[13];
      VLNCE_CHECK_ARG(P.A && P.C, "sgemm_grouped: problem %d has a null operand", base + i);
      VLNCE_CHECK_ARG(P.M > 0 && P.N > 0 && P.K > 0, "sgemm_grouped: problem %d is empty (%dx%dx%d)", base + i, P.M, P.N, P.K);
      VLNCE_CHECK_ARG(P.act >= 0 && P.act <= 2, "sgemm_grouped: problem %d: act must be 0, 1 or 2", base + i);
    }
    const int rc = launch_group(g, workspace, workspace_bytes, stream);
    if (rc != VLNCE_OK) return rc;
  }
  return VLNCE_OK;
}

extern "C" int vlnce_colsum(const float* A, long long lda, float* out, int M, int N, int accumulate,
                            float* workspace, size_t workspace_bytes, cudaStream_t stream) {
  VLNCE_CHECK_ARG(A && out && M > 0 && N > 0, "colsum: bad arguments");
  int S = 1;
  // the head of the workspace holds the split-K arrival counters of sgemm_grouped_kernel: stay behind them
  if (workspace != nullptr && workspace_bytes > (size_t)kCounterBytes) {
    workspace = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + kCounterBytes);
    workspace_bytes -= kCounterBytes;
  } else {
    workspace = nullptr;
  }
  if (workspace != nullptr && M >= 256) {
    S = min(64, M / 128);
    while (S > 1 && (size_t)S * N * sizeof(float) > workspace_bytes) --S;
  }
  dim3 grid(cdiv(N, 32), S);
  launch_pdl(colsum_kernel, grid, dim3(256), 0, stream, A, lda, out, M, N, accumulate, S > 1 ? workspace : (float*)nullptr);
  VLNCE_LAUNCH_CHECK("colsum_kernel");
  if (S > 1) {
    launch_pdl(colsum_final_kernel, dim3(cdiv(N, 128)), dim3(128), 0, stream, (const float*)workspace, S, out, N, accumulate);
    VLNCE_LAUNCH_CHECK("colsum_final_kernel");
  }
  return VLNCE_OK;
}
