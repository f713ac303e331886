// fp32 strided GEMM on CUDA cores for the (small) trainable head:
//   C[m,n] = act( sum_k A[m*lda_m + k*lda_k] * B[k*ldb_k + n*ldb_n] + bias[n] + beta*C[m,n] )
// Arbitrary strides on both operands cover every forward/backward variant of the head's
// Linear / Conv1d(k=1) layers (x W^T, dY W, dY^T X) without transposed copies.  64x64x16
// tiles, 256 threads, 4x4 register micro-tiles, true fp32 (this is the 1e-3-parity path).
#include "common.cuh"

namespace vlnce {
namespace {

constexpr int TM = 64, TN = 64, TK = 32;

__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, long long lda_m, long long lda_k,
             const float* __restrict__ B, long long ldb_k, long long ldb_n, float* __restrict__ C,
             long long ldc, const float* __restrict__ bias, int M, int N, int K, float beta, int act,
             float* __restrict__ partial, int k_per_split) {
  pdl_trigger();  // PDL (common.cuh)
  pdl_wait();
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tx = tid % 16, ty = tid / 16;  // micro-tile coords: rows ty*4.., cols tx*4..
  float acc[4][4] = {};
  // load mappings: favour the contiguous dimension of each operand
  const bool a_kfast = (lda_k == 1);
  const bool b_nfast = (ldb_n == 1);
  // split-K: blockIdx.z owns K range [kb, ke); partial sums go to partial[z][M][N]
  const int kb = blockIdx.z * k_per_split;
  const int ke = min(K, kb + k_per_split);
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
      rb[i] = (gn < N && gk < ke) ? __ldg(B + gk * ldb_k + gn * ldb_n) : 0.f;
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
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (partial != nullptr) {
        partial[((size_t)blockIdx.z * M + gm) * N + gn] = v;
        continue;
      }
      if (bias != nullptr) v += bias[gn];
      float* cp = C + gm * ldc + gn;
      if (beta != 0.f) v += beta * (*cp);
      if (act == 1) v = fmaxf(v, 0.f);
      else if (act == 2) v = tanhf(v);
      *cp = v;
    }
  }
}

// C = act(sum_z partial[z] + bias + beta*C)  — second stage of the split-K GEMM (fixed order:
// deterministic)
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, float* __restrict__ C,
                                     long long ldc, const float* __restrict__ bias, int M, int N,
                                     float beta, int act) {
  pdl_trigger();  // PDL (common.cuh)
  pdl_wait();
  const long long total = (long long)M * N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(idx / N), n = (int)(idx % N);
    float v = 0.f;
    for (int z = 0; z < splits; ++z) v += partial[(size_t)z * total + idx];
    if (bias != nullptr) v += bias[n];
    float* cp = C + m * ldc + n;
    if (beta != 0.f) v += beta * (*cp);
    if (act == 1) v = fmaxf(v, 0.f);
    else if (act == 2) v = tanhf(v);
    *cp = v;
  }
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

extern "C" int vlnce_sgemm(const float* A, long long lda_m, long long lda_k, const float* B,
                           long long ldb_k, long long ldb_n, float* C, long long ldc,
                           const float* bias, int M, int N, int K, float beta, int act,
                           float* workspace, size_t workspace_bytes, cudaStream_t stream) {
  VLNCE_CHECK_ARG(A && B && C, "sgemm: null operand");
  VLNCE_CHECK_ARG(M > 0 && N > 0 && K > 0, "sgemm: empty problem %dx%dx%d", M, N, K);
  VLNCE_CHECK_ARG(act >= 0 && act <= 2, "sgemm: act must be 0 (none), 1 (relu) or 2 (tanh)");
  dim3 grid(cdiv(N, TN), cdiv(M, TM));
  // Skinny problems (M = 64 policy rows) leave most of the 148 SMs idle: split K across
  // blockIdx.z until the grid covers ~2 waves, partial sums reduced in a second launch.
  const int tiles = grid.x * grid.y;
  int splits = 1;
  if (workspace != nullptr && tiles < 96 && K >= 128) {
    splits = min(min(296 / tiles, K / 64), 32);
    while (splits > 1 && (size_t)splits * M * N * sizeof(float) > workspace_bytes) --splits;
  }
  if (splits <= 1) {
    launch_pdl(sgemm_kernel, grid, dim3(256), 0, stream, A, lda_m, lda_k, B, ldb_k, ldb_n, C, ldc, bias, M, N, K, beta, act,
               (float*)nullptr, K);
    VLNCE_LAUNCH_CHECK("sgemm_kernel");
    return VLNCE_OK;
  }
  int k_per = cdiv(cdiv(K, splits), TK) * TK;
  splits = cdiv(K, k_per);
  grid.z = splits;
  launch_pdl(sgemm_kernel, grid, dim3(256), 0, stream, A, lda_m, lda_k, B, ldb_k, ldb_n, C, ldc, bias, M, N, K, beta, act,
             workspace, k_per);
  VLNCE_LAUNCH_CHECK("sgemm_kernel(split-K)");
  const long long total = (long long)M * N;
  const int rgrid = (int)((total + 255) / 256 > 1184 ? 1184 : (total + 255) / 256);
  launch_pdl(splitk_reduce_kernel, dim3(rgrid), dim3(256), 0, stream, (const float*)workspace, splits, C, ldc, bias, M, N,
             beta, act);
  VLNCE_LAUNCH_CHECK("splitk_reduce_kernel");
  return VLNCE_OK;
}

extern "C" int vlnce_colsum(const float* A, long long lda, float* out, int M, int N, int accumulate,
                            float* workspace, size_t workspace_bytes, cudaStream_t stream) {
  VLNCE_CHECK_ARG(A && out && M > 0 && N > 0, "colsum: bad arguments");
  int S = 1;
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
