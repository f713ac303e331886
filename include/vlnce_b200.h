/* vlnce_b200 — C ABI of the B200 (sm_100a) kernels behind the VLN-CE policy hot path.
 *
 * Drop-in boundary: these entry points are what the reference's Python policy classes
 * (vlnce_baselines/models/{policy,cma_policy,seq2seq_policy}.py, encoders/*.py) and the
 * per-step loss (common/base_il_trainer.py:134-180) bind through ctypes in place of their
 * torch/cuDNN/cuBLAS calls (INTEGRATION.md shows the binding).  Each function cites the
 * reference lines whose arithmetic it replaces.
 *
 * Conventions
 *   - plain pointers + sizes; no torch types.  All pointers are DEVICE pointers unless noted.
 *   - the caller owns every tensor and workspace; the library allocates no device memory
 *     and keeps no reference after return (only cached function attributes).
 *   - every call enqueues on `stream` and returns immediately; 0 = ok, <0 = bad argument,
 *     >=1000 = 1000 + cudaError_t.  vlnce_last_error() gives the message (thread-local).
 *   - activations are channels-last: images NHWC, sequences (rows, positions, channels).
 *   - dtype codes: VLNCE_F32, VLNCE_BF16, VLNCE_U8, VLNCE_I64, VLNCE_F16.
 */
#ifndef VLNCE_B200_H_
#define VLNCE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define VLNCE_OK 0
#define VLNCE_ERR_INVALID_ARGUMENT (-1)
#define VLNCE_ERR_NO_DRIVER (-2)
#define VLNCE_ERR_CUDA_BASE 1000

#define VLNCE_F32 0
#define VLNCE_BF16 1
#define VLNCE_U8 2
#define VLNCE_I64 3
#define VLNCE_F16 4

/* ---- library ---------------------------------------------------------------------- */
const char* vlnce_version(void);
const char* vlnce_last_error(void);
/* number of kernel launches issued through this library since process start */
long long vlnce_launch_count(void);
/* Programmatic dependent launch of the chained trunk / GEMM kernels (launch-latency overlap; csrc/common.cuh).
 * Default: environment VLNCE_PDL (unset = off). */
int vlnce_set_pdl(int on);
long long vlnce_get_pdl(void);

/* Workspace sizes (host-only arithmetic; SURVEY 8b's vlnce_<op>_workspace_bytes).  Every entry point below that takes a
 * caller-owned workspace also accepts a smaller one (it narrows its split-K / split-row plan) or NULL; these return the
 * size with which the plan is never narrowed.  vlnce_sgemm / vlnce_sgemm_grouped / vlnce_colsum share one workspace whose
 * first 16 KB are arrival counters (zero before first use, left zero by every call). */
long long vlnce_sgemm_workspace_bytes(int M, int N, int K);
long long vlnce_colsum_workspace_bytes(int M, int N);
long long vlnce_groupnorm_workspace_bytes(int N, int HW, int C, int groups);
long long vlnce_clip_grad_norm_workspace_bytes(long long n);

/* ---- convolution trunks ------------------------------------------------------------ */
/* Implicit-GEMM NHWC conv on tcgen05 tensor cores (16-bit operands, fp32 accumulate in TMEM):
 *   y = act(scale[k] * conv(x, w) + shift[k] (+ residual)).
 * ab_dtype = VLNCE_BF16 or VLNCE_F16 is the type of x, w and residual (fp16 for the
 * GroupNorm depth trunk whose activations are O(1): 8x finer than bf16 at the same cost).
 * x (N,H,W,C), w (K,R,S,C), y (N,P,Q,K) ab_dtype|f32, residual (N,P,Q,K) ab_dtype or NULL,
 * scale/shift (K) f32 or both NULL (raw conv).  C % 32 == 0, K % 32 == 0.  block_n: 0 = auto,
 * else the output-tile width (32/64/128/256) | optional smem ring depth (1..4) << 16.
 * Replaces torch.nn.Conv2d + folded BatchNorm2d/ReLU/residual in torchvision ResNet
 * (resnet_encoders.py:198-199) and the convs of the DD-PPO GroupNorm ResNet
 * (resnet_encoders.py:95; habitat_baselines v0.1.7 rl/ddppo/policy/resnet.py). */
int vlnce_conv2d_bf16(const void* x, const void* w, int ab_dtype, const float* scale,
                      const float* shift, const void* residual, void* y, int y_dtype, int N, int H, int W, int C,
                      int K, int R, int S, int stride, int pad, int relu, int block_n,
                      cudaStream_t stream);

/* RGB stem (torchvision conv1: 7x7/2 pad 3 on the (N,H,W,3) u8|f32 image, resnet_encoders.py:182,198)
 * on the same tensor-core tiles with the im2col row formed on the fly inside the producer warps:
 * w is the packed (K,192) matrix (K index (r*7+s)*3+c, zero padded 147..191); y (N,P,Q,K). */
int vlnce_conv2d_stem_rgb(const void* x, int x_dtype, const void* w, int ab_dtype, const float* scale,
                          const float* shift, void* y, int y_dtype, int N, int H, int W, int K, int relu,
                          int block_n, cudaStream_t stream);

/* Persistent, fully TMA-fed tiles for stride-1 convolutions (R = 1: 1x1 pad 0 over the [M,C] row
 * matrix; R = 3: 3x3 pad 1 with halo boxes over 16x8-pixel tiles, H >= 8 and W >= 8 — sizes that are not multiples of
 * the tile get ragged border tiles, the TMA unit zero-fills / clips what lies outside the image): one CTA per SM claims
 * tiles from a per-launch counter, the epilogue of tile i-1 overlaps the accumulation of tile i (two TMEM
 * accumulators).  C % 64 == 0, K % 64 == 0; same epilogue contract as vlnce_conv2d_bf16.
 * ab_dtype VLNCE_F32 selects the fp32-storage mode: x, w, residual and y are fp32, the products run on
 * the tensor cores as TF32 (tcgen05 kind::tf32, fp32 accumulation) — what stock PyTorch / cuDNN do for
 * fp32 convolutions on this GPU; C % 32 == 0 then.  The fp32 outputs are rounded to the nearest TF32 value (they feed
 * the next layer's TF32 products) unless bit 1 of `relu` is set (relu = 2 | 3: plain fp32 result — the R = 1 mode used
 * as a GEMM  y[M,K] = x[M,C] w[K,C]^T * scale + shift  by the head's large projections, N*H*W = M).
 * block_n: 0 = auto, else width (64/128/256) | weight-ring depth << 16 | (1 << 21): two 128-pixel
 * sub-tiles per scheduled tile share every weight stage (halves the weight bytes per MAC pulled from
 * L2; needs an even tile count and width <= 128) | (1 << 22): never keep the CTA's weight slice
 * stationary in shared memory (by default it is loaded once whenever K == width and it fits). */
int vlnce_conv_tma(const void* x, const void* w, int ab_dtype, const float* scale, const float* shift,
                   const void* residual, void* y, int y_dtype, int N, int H, int W, int C, int K, int R,
                   int relu, int block_n, cudaStream_t stream);

/* RGB stem (torchvision conv1 7x7/2 pad 3, resnet_encoders.py:182,198) as a stride-1 4x4 convolution
 * over the 2x2 space-to-depth image, fully TMA-fed:
 *   vlnce_stem_s2d      x (N,H,W,3) u8|f32 0..255 -> xs (N, P+3, Q+3, 16) bf16|f16 cells, cell (Y,X) =
 *                       input pixels (2Y+dy-3, 2X+dx-3) as channels (dy*2+dx)*3+c, 4 zero channels,
 *                       zeros outside the image; P = (H-1)/2+1, Q = (W-1)/2+1.
 *   vlnce_conv_stem_tma y (N,P,Q,K) = act(scale * (xs (*) w) + shift), w (K,256) = [k][r'][s'][16] with
 *                       w[k][r'][s'][(dy*2+dx)*3+c] = conv1.weight[k][c][2r'+dy][2s'+dx] (0 where the index
 *                       is 7 or the channel is padding).  K % 64 == 0; block_n as vlnce_conv_tma. */
/* Debug aid of the persistent conv tiles: a device buffer of 5*768*2 int64 that CTA 0 of every later
 * vlnce_conv_tma / vlnce_conv_stem_tma launch fills with per-role (event, clock64) records; NULL = off. */
/* cap on the persistent conv_tma grid (0 = one CTA per SM; default: environment VLNCE_TMA_CTA_CAP) */
/* Stride-2 convolutions, R = 1 (pad 0: ResNet downsample branches) or R = 3 (pad 1: conv2 of a stage's first block) —
 * torchvision Bottleneck behind resnet_encoders.py:194-199: x (N,H,W,C) -> y (N,H/2,W/2,K) with the folded-norm epilogue
 * of vlnce_conv_tma.  The activation operand is a TMA box with element strides {1,2,2,1} (only the sampled pixels are
 * loaded; a 3x3 is nine tap-shifted boxes per channel chunk).  H, W even, (W/2) % 8 == 0; block_n as vlnce_conv_tma. */
int vlnce_conv_s2_tma(const void* x, const void* w, int ab_dtype, const float* scale, const float* shift, void* y,
                      int y_dtype, int N, int H, int W, int C, int K, int R, int relu, int block_n, cudaStream_t stream);
int vlnce_set_tma_cta_cap(int cap);
int vlnce_debug_set_trace(void* buffer);
/* rows of the instruction batch per CTA of the persistent LSTM kernels: 1 (default) | 4; 0 = back to the default /
 * environment VLNCE_LSTM_NR (see csrc/common.cuh) */
int vlnce_set_lstm_rows(int rows_per_cta);
/* Debug aid: *slot = %globaltimer (ns) when `stream` reaches this point (graph-capturable timeline probe). */
int vlnce_debug_stamp(long long* slot, cudaStream_t stream);
int vlnce_stem_s2d(const void* x, int x_dtype, void* out, int out_dtype, int N, int H, int W,
                   cudaStream_t stream);
int vlnce_conv_stem_tma(const void* xs, const void* w, int ab_dtype, const float* scale, const float* shift,
                        void* y, int y_dtype, int N, int P, int Q, int K, int relu, int block_n,
                        cudaStream_t stream);

/* 3x3 stride-1 pad-1 conv with TMA halo loads (one (16+2)x(8+2)-pixel box per 64-channel chunk,
 * nine shifted UMMA descriptors = the nine taps): same epilogue contract as vlnce_conv2d_bf16.
 * Needs H % 16 == 0, W % 8 == 0, C % 64 == 0, K % 64 == 0.  block_n: 0 = auto, else width
 * (64/128/256) | weight-ring depth << 16. */
int vlnce_conv3x3_halo(const void* x, const void* w, int ab_dtype, const float* scale, const float* shift,
                       const void* residual, void* y, int y_dtype, int N, int H, int W, int C, int K, int relu,
                       int block_n, cudaStream_t stream);

/* Same contract on CUDA cores with fp32 accumulation and arbitrary C/K: the fp32-exact mode
 * and the on-device cross-check of the tensor-core tiles.  dtypes per tensor. */
int vlnce_conv2d_ref(const void* x, int x_dtype, const void* w, int w_dtype, const float* scale,
                     const float* shift, const void* residual, int res_dtype, void* y,
                     int y_dtype, int N, int H, int W, int C, int K, int R, int S, int stride,
                     int pad, int relu, cudaStream_t stream);

/* fp32 NHWC / KRSC convolution on CUDA cores, warp-cooperative (lanes split the input channels, 8 pixels x 4
 * filters per warp): the fallback of the tf32 mode for geometries the TMA tiles do not cover (4x4 images, sizes that
 * are not multiples of 8).  Same epilogue contract as vlnce_conv2d_ref, fp32 tensors only; the summation order
 * differs from vlnce_conv2d_ref (per-lane partial sums), so the fp32-exact mode does not use it. */
int vlnce_conv2d_warp(const float* x, const float* w, const float* scale, const float* shift, const float* residual,
                      float* y, int N, int H, int W, int C, int K, int R, int S, int stride, int pad, int relu,
                      cudaStream_t stream);

/* RGB stem operand: x (N,H,W,3) u8|f32 0..255 -> im2col rows (N*P*Q, 192) bf16|f16 of the 7x7/2
 * pad-3 window, K index (r*7+s)*3+c, zero padded 147..191 (resnet_encoders.py:182,198:
 * permute + /255 + conv1; the /255 and optional mean/std live in the folded scale/shift). */
int vlnce_im2col_rgb_stem(const void* x, int x_dtype, void* a, int a_dtype, int N, int H, int W,
                          cudaStream_t stream);
/* MODEL.normalize_rgb = True (resnet_encoders.py:171-192, config/default.py:216): y (pixels,3) f32 = x - (m0,m1,m2),
 * m = 255 * ImageNet mean; 1/std is folded into the stem weights and 1/255 into the stem scale by the caller. */
int vlnce_rgb_center(const void* x, int x_dtype, float* y, float m0, float m1, float m2, long long pixels,
                     cudaStream_t stream);
/* Depth stem operand: depth (N,H,W,1) f32 -> avg_pool2d(2) -> 7x7/2 pad-3 im2col rows
 * (N*P*Q, 64) bf16|f16, K index r*7+s, zero padded 49..63 (habitat_baselines v0.1.7
 * ResNetEncoder.forward: permute, avg_pool2d(2), backbone.conv1). */
int vlnce_im2col_depth_stem(const float* x, void* a, int a_dtype, int N, int H, int W,
                            cudaStream_t stream);
/* avg_pool2d(2) on (N,H,W,C) f32 -> (N,H/2,W/2,C) f32 (fp32-exact mode of the depth stem) */
int vlnce_avgpool2x2(const float* x, float* y, int N, int H, int W, int C, cudaStream_t stream);
/* Image-stationary GroupNorm-ResNet trunk (the whole depth encoder of resnet_encoders.py:85-115 /
 * habitat_baselines v0.1.7 ResNetEncoder in ONE launch, csrc/depth_fused.cu): every CTA owns whole images and
 * walks the step table by itself - conv as TMA-fed tcgen05 tiles with the image's accumulators resident in
 * TMEM, GroupNorm (+residual)(+ReLU) applied out of TMEM, 16-bit activations in the caller's scratch.
 * table: n_steps rows of 20 int64 (host memory):
 *   [0] kind (0 conv+GroupNorm, 1 maxpool 3x3/2)  [1] x (N,hin,win,cin) f16  [2] w (cout, k*k*cin) f16
 *   [3] gamma f32  [4] beta f32  [5] residual (N,hout,wout,cout) f16 | 0  [6] y f16 | 0
 *   [7] y_f32 (N*hout*wout, ld) | 0  [8] ld  [9] cin  [10] cout  [11] k (1|3)  [12] stride (1|2)  [13] hin
 *   [14] win  [15] groups  [16] relu  [17] x was written before the launch  [18] eps (float bits)  [19] 0
 * vlnce_depth_fused_plan encodes the table (tensor maps included) into plan_host (host memory,
 * vlnce_depth_fused_plan_bytes(n_steps) bytes); the caller copies it to device memory (128-byte aligned) and
 * passes that to vlnce_depth_fused_run.  The plan holds device addresses: rebuild it when a buffer moves.
 * ctas: persistent grid (0 = one CTA per image up to the SM count); a CTA loops over images b, b+ctas, ... */
long long vlnce_depth_fused_plan_bytes(int n_steps);
int vlnce_depth_fused_plan(const long long* table, int n_steps, int n_images, void* plan_host, long long plan_bytes);
int vlnce_depth_fused_run(const void* plan_dev, int n_images, int ctas, cudaStream_t stream);
/* Debug aid: device buffer of n_steps * 2 * 8 int64; CTA 0 of later vlnce_depth_fused_run launches stamps
 * %globaltimer per (step, slab) of its first image: [0] MMA pass start, [1] last MMA committed, [2] epilogue sees the
 * accumulators, [3] statistics done, [4] pass stored, [5] activation producer released.  NULL = off. */
int vlnce_depth_fused_set_trace(void* buffer);

/* MaxPool2d(3, stride 2, pad 1) NHWC */
int vlnce_maxpool3x3s2(const void* x, void* y, int dtype, int N, int H, int W, int C,
                       cudaStream_t stream);
/* GroupNorm statistics of raw conv output x (N, HW, C): writes per-(n,c) scale/shift with the
 * affine folded: scale = rstd*gamma, shift = beta - mean*rstd*gamma (torch.nn.GroupNorm).
 * workspace (caller-owned, optional) enables the coalesced two-stage reduction. */
int vlnce_groupnorm_stats(const void* x, int dtype, const float* gamma, const float* beta,
                          float* scale, float* shift, float* workspace, size_t workspace_bytes, int N,
                          int HW, int C, int groups, float eps, cudaStream_t stream);
/* GroupNorm (+residual)(+ReLU) of the fp32 raw conv output in two launches: partial statistics,
 * then finalize+normalise (each CTA rebuilds its image's scale/shift from the partials).
 * scale/shift_scratch (N*C f32 each) are used only by the fallback path.  relu: bit 0 = ReLU, bit 1 = round the
 * output to the nearest TF32 value (fp32 activations that feed the kind::tf32 conv tiles, which would truncate). */
int vlnce_groupnorm_fused(const float* x, const float* gamma, const float* beta, const void* residual, void* y,
                          int y_dtype, float* workspace, size_t workspace_bytes, float* scale_scratch,
                          float* shift_scratch, int N, int HW, int C, int groups, float eps, int relu,
                          cudaStream_t stream);
/* the same for an fp16 raw conv output (x, residual, y all fp16): half the bytes per pass */
int vlnce_groupnorm_fused16(const void* x, const float* gamma, const float* beta, const void* residual, void* y,
                            float* workspace, size_t workspace_bytes, int N, int HW, int C, int groups, float eps,
                            int relu, cudaStream_t stream);
/* Train-mode BatchNorm2d statistics over all N*HW rows: per-channel scale/shift
 * from batch mean / biased var; updates running_mean/var (momentum, unbiased var) in place. */
int vlnce_batchnorm_train_stats(const void* x, int dtype, const float* gamma, const float* beta,
                                float* running_mean, float* running_var, float momentum,
                                float* scale, float* shift, double* workspace, long long rows,
                                int C, float eps, cudaStream_t stream);
/* y = act(x*scale[n or 0, c] + shift[..] (+ residual)); scale_stride_n = C (per sample) or 0.
 * x may be the fp32 raw conv output while y / residual are bf16 activations. */
int vlnce_affine_apply(const void* x, int x_dtype, const float* scale, const float* shift,
                       int scale_stride_n, const void* residual, void* y, int y_dtype, int N, int HW,
                       int C, int relu, cudaStream_t stream);
/* adaptive_avg_pool2d(x,(O,O)), O in {1,4}, on NHWC -> f32 (N,O*O,ldy) slice
 * (resnet_encoders.py:154-162 4x4 pool; torchvision's global pool for the non-spatial
 * encoder); with H=W=O it is a plain bf16->f32 layout copy. */
int vlnce_adaptive_avgpool(const void* x, int dtype, float* y, int ldy, int N, int H, int W, int C,
                           int O, cudaStream_t stream);
/* (N,C,4,4) f32 cached features (rgb_features / depth_features, resnet_encoders.py:92-93,
 * 194-195) -> (N,16,ldy) f32 channels-last slice */
int vlnce_nchw_to_positions(const float* x, float* y, int ldy, int N, int C, int P,
                            cudaStream_t stream);

/* ResizeShortestEdge (habitat_baselines v0.1.7 image_resize_shortest_edge, F.interpolate "area")
 * fused with the center crop of CenterCropperPerSensor (habitat_extensions/obs_transformers.py:21-87).
 * x (N,H,W,C) -> y (N,CH,CW,C), same dtype (VLNCE_U8 | VLNCE_F32); the virtual resized image is (RH,RW), the crop
 * window starts at (y0,x0); y_frame_stride (elements, 0 = dense) lets the caller write straight into an ObsStack /
 * pano-stack slot (obs_transformers.py:90-145). */
int vlnce_resize_area_crop(const void* x, void* y, int dtype, long long y_frame_stride, int N, int H, int W, int C,
                           int RH, int RW, int y0, int x0, int CH, int CW, cudaStream_t stream);

/* ---- trainable head: dense algebra --------------------------------------------------- */
/* C[m,n] = act(sum_k A[m*lda_m + k*lda_k] * B[k*ldb_k + n*ldb_n] + bias[n] + beta*C[m,n]),
 * fp32 on CUDA cores; act 0 none / 1 ReLU / 2 tanh.  One routine for every Linear / Conv1d(k=1) of
 * the head forward and backward (cma_policy.py:101-184: rgb_linear, depth_linear, state_q,
 * text_k, text_q, rgb_kv, depth_kv, second_state_compress, progress_monitor; policy.py:19-21
 * CategoricalNet; GRU / LSTM input and hidden projections).
 * workspace (optional, caller-owned, one per stream) enables a deterministic split-K for skinny problems, reduced
 * inside the same launch by the last CTA to reach a tile.  Its first 16 KB are arrival counters: they must be
 * ZERO before the first call and are left zero by every call. */
int vlnce_sgemm(const float* A, long long lda_m, long long lda_k, const float* B, long long ldb_k,
                long long ldb_n, float* C, long long ldc, const float* bias, int M, int N, int K,
                float beta, int act, float* workspace, size_t workspace_bytes, cudaStream_t stream);
/* `count` INDEPENDENT problems of the form above in one launch (no problem may read or accumulate into another's
 * output).  descs: count x 16 64-bit words per problem, in this order:
 *   { A, lda_m, lda_k, B, ldb_k, ldb_n, C, ldc, bias, M, N, K, beta (IEEE-754 double), act, 0, 0 }
 * (pointers as 64-bit addresses).  B == 0 means B == 1 everywhere, i.e. C[m,0..N) = sum_k A[m,k]: with lda_m = 1,
 * lda_k = ld, N = 1 that is the column sum of a row-major matrix — the bias gradients ride in the same launch as
 * the weight gradients.  More than 24 problems are issued as consecutive launches of up to 24. */
int vlnce_sgemm_grouped(const long long* descs, int count, float* workspace, size_t workspace_bytes,
                        cudaStream_t stream);
/* out[n] (+)= sum_m A[m*lda + n]  — bias gradients */
int vlnce_colsum(const float* A, long long lda, float* out, int M, int N, int accumulate,
                 float* workspace, size_t workspace_bytes, cudaStream_t stream);

/* ---- trainable head: forward -------------------------------------------------------- */
/* nn.Embedding gather (instruction_encoder.py:70-73): out[r, :dim] = table[ids[r]] */
int vlnce_embedding_fwd(const long long* ids, const float* table, float* out, long long rows,
                        int dim, int ldo, int vocab, cudaStream_t stream);
/* prev-action embedding with index ((a+1)*mask) (cma_policy.py:233-235); writes the
 * 32-d rows into two concat slices at once */
int vlnce_prev_action_embed(const long long* prev_actions, const float* masks, const float* table,
                            float* out, int ldo, float* out2, int ldo2, int rows, int dim,
                            int table_rows, cudaStream_t stream);
/* per-row instruction length (instruction_encoder.py:77-78), token ids / rxr features */
int vlnce_token_lengths(const long long* ids, int* lengths, int rows, int L, cudaStream_t stream);
int vlnce_feature_lengths(const float* x, int* lengths, int rows, int L, int D, cudaStream_t stream);
/* persistent (bi)LSTM over packed rows (instruction_encoder.py:80-94): gx (dirs,R,L,4H) = input
 * projection + b_ih (b_hh is added here); out (R,L,dirs*H) zero beyond each row's length; hn (dirs,R,H) final
 * states; save_* optional residuals for the backward. */
int vlnce_lstm_seq_fwd(const float* gx, const float* whh_f, const float* whh_b, const float* bhh_f,
                       const float* bhh_b, const int* lengths, float* out, float* hn,
                       float* save_gates, float* save_c, int R, int L, int H, int dirs,
                       cudaStream_t stream);
/* persistent (bi)GRU over packed rows (instruction_encoder.py:27,80-94 with INSTRUCTION_ENCODER.rnn_type = "GRU"):
 * gx (dirs,R,L,3H) = input projection + b_ih; gate order r,z,n; out / hn as vlnce_lstm_seq_fwd;
 * save (dirs,R,L,4H) = r, z, n, W_hn h + b_hn for the backward (optional). */
int vlnce_gru_seq_fwd(const float* gx, const float* whh_f, const float* whh_b, const float* bhh_f,
                      const float* bhh_b, const int* lengths, float* out, float* hn, float* save, int R, int L,
                      int H, int dirs, cudaStream_t stream);
/* GRU cell pointwise part (habitat_baselines v0.1.7 rnn_state_encoder / torch.nn.GRU) */
int vlnce_gru_pointwise_fwd(const float* gi, const float* gh, const float* h, long long ldh,
                            float* hout, long long ldo, float* hout2, long long ldo2, float* save,
                            int rows, int H, cudaStream_t stream);
/* One whole GRU step of the state encoder in ONE launch (cma_policy.py:248-256 and :287-294, seq2seq_policy.py:112-137
 * through habitat_baselines v0.1.7 RNNStateEncoder.single_forward): hm = h0 * mask[row]; gi = x W_ih^T + b_ih;
 * gh = hm W_hh^T + b_hh; gates r, z, n; out = (1 - z) n + z hm (optionally also into out2, the rnn_states_out slot).
 * hm (rows,H) and save (rows,4H) = r, z, n, gh_n are what vlnce_gru_pointwise_bwd and the weight-gradient GEMMs read
 * (either may be NULL for inference).  Replaces vlnce_mask_rows + two vlnce_sgemm + vlnce_gru_pointwise_fwd.
 * H % 4 == 0 and 12 * (in_size + H) floats of shared memory (<= 200 KB). */
int vlnce_gru_step_fwd(const float* x, long long ldx, int in_size, const float* h0, long long ldh0, const float* mask,
                       const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, float* hm, float* out,
                       long long ldo, float* out2, long long ldo2, float* save, int rows, int H, cudaStream_t stream);
/* LSTM cell pointwise part (rnn_state_encoder with STATE_ENCODER.rnn_type = "LSTM", cma_policy.py:129,175;
 * seq2seq_policy.py:112; waypoint_predictors.py:72,160): gates i,f,g,o from gi + gh (rows,4H), c = mask-reset cell
 * state; writes h', c' (and optional second copies into the rnn_states_out slots); save (rows,5H) for the backward. */
int vlnce_lstm_pointwise_fwd(const float* gi, const float* gh, const float* c, long long ldc, float* hout,
                             long long ldo, float* cout, long long ldco, float* hout2, float* cout2, long long ldo2,
                             float* save, int rows, int H, cudaStream_t stream);
/* out = h * mask[row]  (hidden reset at episode starts) */
int vlnce_mask_rows(const float* h, long long ldh, const float* mask, float* out, long long ldo,
                    int rows, int H, cudaStream_t stream);
/* single-query attention pooling (cma_policy.py:207-217, keys/values re-associated) */
int vlnce_attn_pool_fwd(const float* q, long long ldq, const float* E, long long lde_r,
                        long long lde_p, const int* lengths, float scale, float* probs, float* out,
                        long long ldo, int R, int P, int D, cudaStream_t stream);
int vlnce_mean_positions(const float* E, long long lde_r, long long lde_p, float* out, long long ldo,
                         int R, int P, int D, cudaStream_t stream);
int vlnce_spatial_embed_fill(const float* table, float* E, long long lde_r, long long lde_p,
                             int c_off, int R, int P, int Cs, cudaStream_t stream);
int vlnce_positions_to_nchw(const float* E, long long lde_r, long long lde_p, float* out,
                            long long ldo, int R, int P, int C, cudaStream_t stream);
/* dst[c*ldd + r] = src[r*lds + c] (fp32, 32x32 smem tiles), optionally rounded to the nearest TF32 value; and the
 * element-wise rounding alone (dst may equal src).  Operand preparation of the TF32 tensor-core GEMMs of the head
 * (instruction-feature input projection and its weight gradient, instruction_encoder.py:27,80-94): kind::tf32 would
 * truncate un-rounded operands, a systematic 2^-10 bias on every product. */
int vlnce_transpose_f32(const float* src, long long lds, float* dst, long long ldd, int rows, int cols, int round_tf32,
                        cudaStream_t stream);
int vlnce_round_tf32(const float* src, float* dst, long long n, cudaStream_t stream);
int vlnce_log_softmax_rows(float* x, int rows, int A, cudaStream_t stream);
int vlnce_argmax_rows(const float* x, long long* out, int rows, int A, cudaStream_t stream);

/* ---- trainable head: backward, loss, optimiser -------------------------------------- */
/* d[r,c] = 0 where y[r,c] <= 0 (ReLU backward, in place) */
int vlnce_relu_bwd(float* d, long long ldd, const float* y, long long ldy, int rows, int cols,
                   cudaStream_t stream);
/* d *= 1 - y^2 (tanh backward, in place) */
int vlnce_tanh_bwd(float* d, const float* y, long long n, cudaStream_t stream);
/* backward of lp = raw - logsumexp(raw): draw = dlp - exp(lp) * sum_a dlp */
int vlnce_log_softmax_bwd(const float* lp, const float* dlp, float* draw, int rows, int A,
                          cudaStream_t stream);
/* dst += src * mask[row]  (hidden-state gradient carried across a time step) */
int vlnce_mask_rows_add(const float* src, long long lds, const float* mask, float* dst, long long ldd,
                        int rows, int H, cudaStream_t stream);
int vlnce_gru_pointwise_bwd(const float* dh, long long lddh, const float* save, const float* hm,
                            float* dgi, float* dgh, float* dhm, int rows, int H, cudaStream_t stream);
/* LSTM cell backward: dgates (rows,4H) pre-activation gradients (shared by W_ih and W_hh); dc_in nullable (= 0),
 * dc_out = d c_prev * mask[row] (mask nullable = 1) carried to the step before */
int vlnce_lstm_pointwise_bwd(const float* dh, long long lddh, const float* dc_in, const float* save, const float* cm,
                             const float* mask, float* dgates, float* dc_out, int rows, int H, cudaStream_t stream);
int vlnce_attn_pool_bwd(const float* q, long long ldq, const float* E, long long lde_r, long long lde_p,
                        const float* probs, float scale, const float* dout, long long lddo, float* dq,
                        long long lddq, float* dE, long long ldde_r, long long ldde_p, int accumulate_dE,
                        int R, int P, int D, cudaStream_t stream);
/* BPTT of the packed (bi)GRU: dgx / dgh (dirs,R,L,3H) gradients w.r.t. the input- and hidden-side gate
 * pre-activations, hprev (dirs,R,L,H) the hidden state fed to every step; zero beyond each row's length. */
int vlnce_gru_seq_bwd(const float* dout, const float* save, const float* out, const float* whh_f,
                      const float* whh_b, const int* lengths, float* dgx, float* dgh, float* hprev, int R, int L,
                      int H, int dirs, cudaStream_t stream);
/* BPTT of the packed (bi)LSTM: emits pre-activation gate gradients dgx (dirs,R,L,4H) and the
 * hidden state fed to every step hprev (dirs,R,L,H); the weight gradients are then GEMMs. */
int vlnce_lstm_seq_bwd(const float* dout, const float* gates, const float* cs, const float* out,
                       const float* whh_f, const float* whh_b, const int* lengths, float* dgx,
                       float* hprev, int R, int L, int H, int dirs, cudaStream_t stream);
/* Gradient entering a final-state-only instruction encoder (instruction_encoder.py:87-90; Seq2Seq):
 * dout (R,L,H) = 0 except dout[r, lengths[r]-1, :] = dh[r, :] — then vlnce_lstm_seq_bwd as usual. */
int vlnce_scatter_last_token(const float* dh, long long lddh, const int* lengths, float* dout, int R,
                             int L, int H, cudaStream_t stream);
int vlnce_embedding_bwd(const long long* ids, const float* d, long long ldd, float* dtable,
                        long long rows, int dim, int vocab, int padding_idx, cudaStream_t stream);
int vlnce_prev_action_embed_bwd(const long long* prev_actions, const float* masks, const float* d1,
                                long long ld1, const float* d2, long long ld2, float* dtable, int rows,
                                int dim, int table_rows, cudaStream_t stream);
int vlnce_spatial_embed_grad(const float* dE, long long lde_r, long long lde_p, int c_off,
                             const float* dflat, long long ldf, const float* dmean, long long ldm,
                             float* dtable, int accumulate, int R, int P, int Cs, cudaStream_t stream);
/* Fused DAgger / teacher-forcing loss and its gradients (base_il_trainer.py:159-172,
 * aux_losses.py:24-32, cma_policy.py:296-307): logp (T*N,A) log-probs, actions/weights (T,N),
 * hat/progress (T*N); progress_mode 0 none / 1 elementwise / 2 the (R,R) broadcast of SURVEY Q3.
 * losses[3] = {loss/accum, action_loss, aux_loss}; d_raw (T*N,A) and d_hat (T*N) already carry
 * the 1/accum factor. */
int vlnce_il_loss(const float* logp, const long long* actions, const float* weights, const float* hat,
                  const float* progress, int progress_mode, float alpha, float inv_accum, int T, int N,
                  int A, float* d_raw, float* d_hat, float* losses, cudaStream_t stream);
/* torch.optim.Adam step (no weight decay / amsgrad) over flat fp32 buffers; grad_scale folds the
 * 1/world_size of the data-parallel gradient all-reduce.  step > 0: host step count;
 * step == 0: read it from step_dev (CUDA-graph replays). */
int vlnce_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                    float beta2, float eps, int step, const int* step_dev, float grad_scale,
                    cudaStream_t stream);
/* Inflection weights (dagger_trainer.py:199-211, recollection_dataset.py:251-263) on the padded time-major
 * (T,N) oracle-action matrix of collate_fn: coef where the action changes and at t = 0, 1 elsewhere, 0 where
 * t >= lengths[n] (lengths may be NULL: no padding). */
int vlnce_inflection_weights(const long long* actions, const int* lengths, float* weights, int T, int N, float coef,
                             cudaStream_t stream);
/* Device-resident trajectory cache — batch assembly of dagger_trainer.py:39-114 (collate_fn) + :189-211
 * + the float32 staging of :560-566 in one launch per observation key: arena holds the trajectories' rows back to
 * back (dtype VLNCE_F32 | F16 (IL.DAGGER.lmdb_fp16) | BF16 | U8 | I64), trajectory n starts at element offsets[n] and
 * has lengths[n] rows of D elements; out (T*N, D) f32 time-major, rows t >= lengths[n] = pad (1.0 for observations). */
int vlnce_collate_rows(const void* arena, int dtype, const long long* offsets, const int* lengths, float* out, int T,
                       int N, long long D, float pad, cudaStream_t stream);
/* ... and the (T,N) matrices of the same batch: prev / oracle actions padded with 0, inflection weights
 * ([1, coef][a_t != a_{t-1}], first step an inflection, 0 in the padding), not_done_masks (0 on the first row block). */
int vlnce_collate_actions(const long long* prev_arena, const long long* act_arena, const long long* offsets,
                          const int* lengths, long long* prev_out, long long* act_out, float* weights,
                          unsigned char* masks, int T, int N, float coef, cudaStream_t stream);
/* device-resident step counter for graph replays; y += alpha * x for gradient accumulation
 * (loss_accumulation_scalar, base_il_trainer.py:170-176) */
int vlnce_counter_add(int* counter, int value, cudaStream_t stream);
int vlnce_axpy(float* y, const float* x, float alpha, long long n, cudaStream_t stream);

/* ---- waypoint path --------------------------------------------- */
/* waypoint_predictors.py:327-366: out = [B*P pano frames | B history frames, zeroed where masks == 0],
 * one trunk batch; byte movement only (frame_bytes = H*W*C*sizeof(element)). */
int vlnce_wp_stack_frames(const void* cur, const void* hist, const float* masks, void* out, int B, int P,
                          long long frame_bytes, cudaStream_t stream);
/* E[r,p,0:C] = E[r,0,0:C], p >= 1 (single_spatial_filter=False keeps the global pool,
 * resnet_encoders.py:154-162) */
int vlnce_broadcast_positions(float* E, long long lde_r, long long lde_p, int R, int P, int C,
                              cudaStream_t stream);
/* waypoint_predictors.py:268-275,372-384: [sin h, cos h, off_a + off_b*offset, dist_a + dist_b*distance] * mask,
 * h = pano * 2pi / panos; written to two concat slices */
int vlnce_wp_prev_action(const long long* pano, const float* offset, const float* distance, const float* masks,
                         float* out, long long ldo, float* out2, long long ldo2, int rows, int panos, float off_a,
                         float off_b, float dist_a, float dist_b, cudaStream_t stream);
/* waypoint_predictors.py:419-424 with the multiplicative inverted mask of utils.py:173-175 in closed
 * form; instr (R,L,D) zero beyond each row's length; save (R,2) */
int vlnce_wp_inst_attn_fwd(const float* q, long long ldq, const float* bk, const float* instr, const int* lengths,
                           float scale, float* out, long long ldo, float* save, int R, int L, int D, int Dk,
                           cudaStream_t stream);
int vlnce_wp_inst_attn_bwd(const float* save, const float* out, long long ldo, const float* dout, long long lddo,
                           const int* lengths, float scale, float* dcs, float* d_instr, int R, int L, int D,
                           cudaStream_t stream);
/* waypoint_predictors.py:459-464: out.flat[m] = q.flat[m / P] (repeat_interleave of ELEMENTS) and its adjoint */
int vlnce_wp_scramble_query(const float* q, float* out, long long n_in, int P, cudaStream_t stream);
int vlnce_wp_scramble_query_bwd(const float* dout, float* dq, long long n_in, int P, cudaStream_t stream);
/* nn.LayerNorm over the last dimension (utils.py:214-215,263-265); save (R,2) = mean, rstd;
 * bwd also emits prod = dy * xhat whose column sum is d_gamma */
int vlnce_layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, float* y, long long ldy,
                        float* save, int R, int D, float eps, cudaStream_t stream);
int vlnce_layernorm_bwd(const float* dy, long long lddy, const float* x, long long ldx, const float* gamma,
                        const float* save, float* dx, long long lddx, float* prod, long long ldp, int R, int D,
                        cudaStream_t stream);
/* waypoint_predictors.py:556-559: out[b,i] = F[b,i,:] . xs[b,:] and its adjoints */
int vlnce_wp_dot_frames_fwd(const float* F, long long ldf_r, long long ldf_p, const float* xs, long long ldx,
                            float* out, long long ldo, int B, int P, int D, cudaStream_t stream);
int vlnce_wp_dot_frames_bwd(const float* F, long long ldf_r, long long ldf_p, const float* xs, long long ldx,
                            const float* dout, long long lddo, float* dF, long long lddf_r, long long lddf_p,
                            int accumulate_dF, float* dxs, long long lddx, int B, int P, int D, cudaStream_t stream);
/* out[b,i,:] = x[b,:] (waypoint_predictors.py:561-567) / out[r,:] (+)= sum_p E[r,p,:] */
int vlnce_broadcast_rows(const float* x, long long ldx, float* out, long long ldo_r, long long ldo_p, int B, int P,
                         int D, cudaStream_t stream);
int vlnce_sum_positions(const float* E, long long lde_r, long long lde_p, float* out, long long ldo, int accumulate,
                        int R, int P, int D, cudaStream_t stream);
/* Output activations of the continuous heads (waypoint_predictors.py:573-613).  raw (n,4): distance, distance
 * variance, offset, offset variance pre-activations.  consts_host: HOST pointer to 8 floats {d_span, d_min,
 * dv_span, dv_min, o_scale, 1/temperature, ov_span, ov_min}. */
int vlnce_wp_head_act_fwd(const float* raw, float* d1, float* d2, float* o1, float* o2, long long n,
                          const float* consts_host, cudaStream_t stream);
int vlnce_wp_head_act_bwd(const float* raw, const float* g_d1, const float* g_d2, const float* g_o1, const float* g_o2,
                          float* g_raw, long long n, const float* consts_host, cudaStream_t stream);
/* WaypointPolicy.act post-processing (waypoint_policy.py:129-247; TruncatedNormal utils.py:24-152).
 * logp (B,P+1) log-probs; o1,o2,d1,d2 (B,P).  pano_stop (B) i64; fout (B,10) = offset, distance,
 * action_log_prob, var_distance, var_offset, mode_offset, mode_distance, r, theta, stop.
 * act_consts_host: HOST pointer to 11 floats {d_lo, d_hi, offset_limit, predict_distance, predict_offset,
 * #distance classes, #offset classes (0 = continuous TruncatedNormal head), dist_a, dist_b, off_a, off_b
 * (to_continuous = a + b * value)}.  A discrete head passes its logits (B,P,classes) as d1 / o1 and NULL for
 * d2 / o2; fout then carries class indices and NaN variances (torch Categorical.variance). */
int vlnce_wp_act(const float* logp, const float* o1, const float* o2, const float* d1, const float* d2,
                 long long* pano_stop, float* fout, int B, int P, int deterministic, unsigned long long seed,
                 unsigned long long offset, const float* act_consts_host, cudaStream_t stream);
/* evaluate_actions (waypoint_policy.py:265-333) fused with the WDDPPO mini-batch loss and its gradients
 * (common/ddppo_alg.py:56-124).  eval_out (R,4) = action_log_prob, entropy pano/offset/distance; losses[8] =
 * loss, value_loss, action_loss, entropy_loss, offset_loss, mean entropies (pano, offset, distance).
 * losses and the gradient outputs may be NULL.  ppo_consts_host: HOST pointer to 8 floats {clip, value_coef,
 * entropy_coef, offset_regularize_coef, pano/offset/distance entropy coefs, clipped_value_loss}.  Gradient blocks
 * of a discrete head are (R,P,classes) in d_d1 / d_o1 (d_d2 / d_o2 unused). */
int vlnce_wp_ppo_loss(const float* logp, const float* o1, const float* o2, const float* d1, const float* d2,
                      const float* values, const float* a_pano, const float* a_offset, const float* a_distance,
                      const float* v_old, const float* ret, const float* lp_old, const float* adv, int R, int P,
                      const float* act_consts_host, const float* ppo_consts_host, float* eval_out, float* losses,
                      float* d_raw, float* d_o1, float* d_o2, float* d_d1, float* d_d2, float* d_values,
                      cudaStream_t stream);
/* backward of evaluate_actions for arbitrary upstream gradients (R,4) = d{action_log_prob, entropy_pano,
 * entropy_offset, entropy_distance} — what the reference's own loss.backward() (ddppo_alg.py:127) needs */
int vlnce_wp_eval_bwd(const float* logp, const float* o1, const float* o2, const float* d1, const float* d2,
                      const float* a_pano, const float* a_offset, const float* a_distance, const float* upstream,
                      int R, int P, const float* act_consts_host, float* d_raw, float* d_o1, float* d_o2,
                      float* d_d1, float* d_d2, cudaStream_t stream);
/* torch.nn.utils.clip_grad_norm_ (habitat_baselines v0.1.7 PPO.before_step; ddppo_alg.py:131): g *= pre_scale *
 * min(1, max_norm / (pre_scale * ||g|| + 1e-6)); ws >= 1026 floats, ws[1024] = norm, ws[1025] = factor. */
int vlnce_clip_grad_norm(float* g, long long n, float max_norm, float pre_scale, float* ws, size_t ws_bytes,
                         cudaStream_t stream);
/* ActionDictRolloutStorage.compute_returns (common/rollout_storage.py:127-152): rewards (T,N), value_preds and
 * returns (T+1,N), masks (T+1,N) u8, next_value (N); optional advantages (T,N) = returns - value_preds
 * (WDDPPO.get_advantages, common/ddppo_alg.py:30-32). */
int vlnce_gae_returns(const float* rewards, float* value_preds, const unsigned char* masks, const float* next_value,
                      float* returns, float* advantages, int T, int N, float gamma, float tau, int use_gae,
                      cudaStream_t stream);

/* ---- backward of the conv trunks (RGB_ENCODER.trainable / DEPTH_ENCODER.trainable, config/default.py:242,253;
 * cma_policy.py:80,94) — fp32 correctness-first path mirroring vlnce_conv2d_ref.  All tensors fp32 NHWC / KRSC. ---- */
/* dx (N,H,W,C) (+)= dgrad of dy (N,P,Q,K) through w (K,R,S,C) */
int vlnce_conv2d_ref_dgrad(const float* dy, const float* w, float* dx, int N, int H, int W, int C, int K, int R, int S,
                           int stride, int pad, int accumulate, cudaStream_t stream);
/* dw (K,R,S,C) (+)= alpha * wgrad of x (N,H,W,C; f32 or the u8 image of the RGB stem) and dy (N,P,Q,K) */
int vlnce_conv2d_ref_wgrad(const void* x, int x_dtype, const float* dy, float* dw, int N, int H, int W, int C, int K,
                           int R, int S, int stride, int pad, float alpha, int accumulate, cudaStream_t stream);
/* GroupNorm backward on the raw conv output x (N,HW,C): dx, and per-image partial (N,2C) = [dgamma | dbeta] */
int vlnce_groupnorm_bwd(const float* x, const float* dy, const float* gamma, float* dx, float* partial, int N, int HW,
                        int C, int groups, float eps, cudaStream_t stream);
/* BatchNorm2d backward on the raw conv output x (rows,C): batch statistics (training != 0) or the running ones */
int vlnce_batchnorm_bwd(const float* x, const float* dy, const float* gamma, const float* running_mean,
                        const float* running_var, float* dx, float* dgamma, float* dbeta, long long rows, int C,
                        float eps, int training, cudaStream_t stream);
/* max_pool2d(3, stride 2, pad 1) backward: dx zeroed, gradient routed to the first maximum of each window */
int vlnce_maxpool3x3s2_bwd(const float* x, const float* dy, float* dx, int N, int H, int W, int C, cudaStream_t stream);
/* adaptive_avg_pool2d backward from the positions layout dout (N, O*O, ld) (channels [0,C) of each row) */
int vlnce_adaptive_avgpool_bwd(const float* dout, long long ld, float* dx, int N, int H, int W, int C, int O,
                               cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VLNCE_B200_H_ */
