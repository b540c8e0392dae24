/*
 * include/memotr_b200.h -- C ABI of the B200-native MeMOTR hot path (libmemotr_b200.so).
 *
 * Plain pointers and sizes only: no torch / ATen types cross this boundary.  Every entry point
 *   - takes DEVICE pointers to caller-allocated, contiguous buffers (inputs are borrowed, never freed);
 *   - enqueues its kernels on the CUDA stream passed as `stream` (a cudaStream_t cast to void*; NULL = legacy
 *     default stream) and never synchronises, allocates device memory, or keeps global mutable state
 *     (re-entrant; safe to call from the autograd engine thread and under CUDA-graph capture);
 *   - returns 0 on success or a negative MEMOTR_E* code; memotr_last_error() then returns a thread-local
 *     message.  (The reference only printf()s launch errors -- ms_deform_im2col_cuda.cuh:948-952 -- its argument
 *     errors are AT_ASSERTM/AT_ERROR exceptions, ms_deform_attn_cuda.cu:28-52; the Python shim turns a
 *     non-zero return into the same RuntimeError.)
 *
 * Reference interfaces replaced (paths relative to /root/reference/models/ops/):
 *   memotr_msda_forward   <- ms_deform_attn_forward   src/ms_deform_attn.h:20-39,  src/vision.cpp:14
 *                            ms_deform_attn_cuda_forward  src/cuda/ms_deform_attn_cuda.cu:20-80
 *   memotr_msda_backward  <- ms_deform_attn_backward  src/ms_deform_attn.h:41-61,  src/vision.cpp:15
 *                            ms_deform_attn_cuda_backward src/cuda/ms_deform_attn_cuda.cu:83-153
 * The remaining entry points have no native counterpart in the reference (there the work is a chain of ATen
 * calls issued from Python); each cites the Python lines it stands in for.
 */
#ifndef MEMOTR_B200_H_
#define MEMOTR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: round 2 -- memotr_msda_forward_strided lost its layout flag, the pair / head-major / prologue-LN entry points are gone,
 * the windowed gather, the fused-LayerNorm variants, the pipelining / fp32x3 / input-projection entry points were added */
#define MEMOTR_ABI_VERSION 2

#if defined(__GNUC__)
#define MEMOTR_API __attribute__((visibility("default")))
#else
#define MEMOTR_API
#endif

/* element types of the floating-point buffers */
#define MEMOTR_F32 0
#define MEMOTR_F64 1
#define MEMOTR_BF16 2
#define MEMOTR_F16 3  /* value maps only: memotr_linear output, memotr_msda_forward_ex input */

/* return codes */
#define MEMOTR_OK 0
#define MEMOTR_EINVAL (-1)   /* bad argument (shape, dtype, null pointer, misalignment) */
#define MEMOTR_ECUDA (-2)    /* CUDA runtime / launch error                             */
#define MEMOTR_ENOSYS (-3)   /* combination not implemented                             */

MEMOTR_API int memotr_abi_version(void);
MEMOTR_API const char *memotr_last_error(void);

/*
 * Multi-scale deformable attention, forward.
 *   value            (B, S, H, D)        dtype     pixel-major, S = sum_l H_l*W_l
 *   spatial_shapes   (L, 2)              int64     (H_l, W_l), device memory
 *   level_start_idx  (L,)                int64     device memory
 *   sampling_loc     (B, Lq, H, L, K, 2) dtype     last dim (x, y) in [0,1] image coordinates
 *   attn_weight      (B, Lq, H, L, K)    dtype
 *   output           (B, Lq, H*D)        dtype     every element is written (no pre-zeroing needed)
 * dtype F32 with D == 32 reproduces the reference kernel's fp32 rounding sequence bit for bit.
 * The reference's `im2col_step` batching argument has no effect on the result and is not part of this ABI.
 */
MEMOTR_API int memotr_msda_forward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_idx,
                        const void *sampling_loc, const void *attn_weight, void *output, int B, int S, int H,
                        int D, int L, int Lq, int K, int dtype, void *stream);

/*
 * Multi-scale deformable attention, backward.
 *   grad_output      (B, Lq, H*D)  dtype
 *   grad_value       like value         -- MUST be zero-filled by the caller (accumulated with atomics)
 *   grad_sampling_loc like sampling_loc -- fully written
 *   grad_attn_weight like attn_weight   -- fully written
 * dtype F32 or F64.
 */
MEMOTR_API int memotr_msda_backward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_idx,
                         const void *sampling_loc, const void *attn_weight, const void *grad_output,
                         void *grad_value, void *grad_sampling_loc, void *grad_attn_weight, int B, int S, int H,
                         int D, int L, int Lq, int K, int dtype, void *stream);


/*
 * Engine variant of the forward op: value/output in `dtype` (F32 or BF16; F16 = fp16 value map with bf16 output), sampling locations and attention
 * weights always fp32 (the outputs of memotr_msda_prep), D == 32, and an explicit pixel stride (elements between
 * consecutive pixels of `value`, >= H*32) so that the value maps of all decoder layers can live interleaved in one
 * (S, n_layers*256) buffer written by a single GEMM.  Same arithmetic as memotr_msda_forward.
 */
MEMOTR_API int memotr_msda_forward_ex(const void *value, int value_pixel_stride, const int64_t *spatial_shapes,
                                      const int64_t *level_start_idx, const float *sampling_loc,
                                      const float *attn_weight, void *output, int B, int S, int H, int L, int Lq,
                                      int K, int dtype, void *stream);

/* memotr_msda_forward_ex for an fp16 value map (bf16 output) with strided sampling locations / attention weights */
MEMOTR_API int memotr_msda_forward_strided(const void *value, int value_pixel_stride, const int64_t *spatial_shapes,
                                           const int64_t *level_start_idx, const float *sampling_loc, int ld_loc,
                                           const float *attn_weight, int ld_attn, void *output, int B, int S, int H, int L,
                                           int Lq, int K, void *stream);

/* The offsets / attention-logits projection of MSDeformAttn with memotr_msda_prep (encoder mode) fused into the GEMM
 * epilogue (ms_deform_attn.py:104-120): out (M, 3*H*L*K) fp32 rows = [sampling locations (H, L*K, 2) | softmax weights
 * (H, L*K)]; A (M,K), W (3*H*L*K, K) bf16; shapes_hw (2L) / level_start (L) host arrays, valid_ratios (L,2) device.
 * Needs L*K == 16 and more 128 x 128 output tiles than SMs (the persistent tcgen05 kernel). */
MEMOTR_API int memotr_linear_msda_prep(const void *A, int lda, const void *W, int ldw, const float *bias, float *out, int ldo,
                                       int M, int K, int n_heads, int n_levels, int n_points, const int *shapes_hw,
                                       const int *level_start, const float *valid_ratios, void *stream);

/*
 * Encoder-shaped forward op (queries = the S pixels of the pyramid, models/deformable_encoder.py:124; batch 1) with
 * TMA-staged value-map windows in shared memory (csrc/msda_window.cu): replaces ms_deformable_im2col_gpu_kernel
 * (models/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299) for an fp16 pixel-major value map (pixel stride >= H*32) and a
 * bf16 output (S, H*32); sampling locations / attention weights fp32 with row strides ld_loc / ld_attn.  Bit-identical to
 * memotr_msda_forward_strided on the same inputs.
 *   shapes_hw (2L) / level_start (L): HOST copies of spatial_shapes / level_start_idx; valid_ratios (L,2) device.
 *   window_shift: host (H, L, 2) floats or NULL -- expected sampling offset (x, y) of each head on each level in pixels of
 *     that level (for MSDeformAttn: the mean over the K points of sampling_offsets.bias, ms_deform_attn.py:72-81);
 *   window_radius: spread of the samples around it, pixels.  Both only steer WHAT is staged: a tap outside its window is
 *     read from global memory with the same arithmetic.
 *   max_classes: 0 = default (the queries of up to four levels); n = stage windows only for the queries of levels < n,
 *     the rest reads global memory.
 *   stats: device, 2 x uint64 or NULL -- profiling counters += {sampling points served from windows, points left to
 *     global memory} over the window units of this launch.
 * L <= 5, H <= 16, K even.
 */
MEMOTR_API int memotr_msda_forward_window(const void *value, int value_pixel_stride, const int *shapes_hw, const int *level_start,
                                          const float *sampling_loc, int ld_loc, const float *attn_weight, int ld_attn,
                                          const float *valid_ratios, const float *window_shift, float window_radius,
                                          int max_classes, unsigned long long *stats, void *output, int S, int H, int L, int K,
                                          void *stream);
/* The staging plan of memotr_msda_forward_window as integers (host only, no GPU needed): info[0..7] = {classes, window
 * units (CTAs), global-memory CTAs, first global-memory query, dynamic shared memory bytes per CTA, 0, 0, 0}, then per
 * class 18 ints {query level, tile w, tile h, tiles_x, tiles_y, units, TMA bytes per unit, record stride, ww[5], wh[5]}.
 * `info` must hold 8 + 4 * 18 ints (up to four classes: the queries of levels 0 .. 3). */
MEMOTR_API int memotr_msda_window_plan(const int *shapes_hw, const int *level_start, int S, int H, int L, int K, float radius,
                                       int max_classes, int *info);

/*
 * Sampling locations + attention weights from the raw projections -- models/ops/modules/ms_deform_attn.py:108-120.
 *   ol            (Lq, ldol) fp32: per row [ offsets (H,L,K,2) | logits (H,L*K) ]  (one GEMM with the stacked
 *                 sampling_offsets / attention_weights weights)
 *   valid_ratios  (L,2) fp32 (w,h) per level -- models/deformable_transformer.py:175-190
 *   mode 0 (encoder self-attention): query q is the q-th pixel of the pyramid; its 2-d reference point is rebuilt from
 *          spatial_shapes / level_start_idx / valid_ratios  (models/deformable_encoder.py:29-40);  ref4 unused
 *   mode 1 (decoder cross-attention): ref4 (Lq,4) sigmoid-space boxes, scaled per level by valid_ratios
 *          (models/deformable_decoder.py:82-84)
 *   -> sampling_loc (Lq,H,L,K,2) fp32, attn_weight (Lq,H,L,K) fp32 (softmax over the joint L*K axis)
 */
MEMOTR_API int memotr_msda_prep(const float *ol, int ldol, const int64_t *spatial_shapes,
                                const int64_t *level_start_idx, const float *valid_ratios, const float *ref4, int mode,
                                float *sampling_loc, float *attn_weight, int Lq, int H, int L, int K, void *stream);

/*
 * (c_dtype MEMOTR_F16 is additionally accepted on the tensor-core path: the value maps the sampling kernel reads.)
 * C = epilogue(A . W^T):  v = acc + bias;  act (0 none, 1 ReLU, 2 sigmoid);  v *= mul;  v += add;  rows with
 * rowzero[m] != 0 are written as zeros.  A (M,K) lda, W (N,K) ldw, both `ab_dtype` (F32 or BF16); C (M,N) ldc in
 * `c_dtype` (F32, or BF16 when ab_dtype is BF16); mul/add (M,N) in ab_dtype; bias fp32; fp32 accumulation.
 * path 0 = auto (bf16: M <= 1024 rows -> latency-optimised mma.sync kernel; otherwise N % 64 == 0, K % 64 == 0 ->
 * tcgen05/TMA/TMEM kernel; otherwise CUDA-core kernel), 1 = force CUDA cores, 2 = force tcgen05, 3 = force mma.sync
 * (MEMOTR_ENOSYS if the shape is unsupported by a forced path).
 * Replaces torch.nn.functional.linear at every call site of the hot path: models/ops/modules/ms_deform_attn.py:104-129,
 * models/deformable_encoder.py:97-107, models/deformable_decoder.py:245-273, models/mlp.py:22-25, models/ffn.py:15-25,
 * models/query_updater.py:109-132, models/memotr.py:153-154 (and the padding-mask fill of ms_deform_attn.py:106).
 */
MEMOTR_API int memotr_linear(const void *A, int lda, const void *W, int ldw, const float *bias, const void *mul,
                             int ldmul, const void *add, int ldadd, const unsigned char *rowzero, void *C, int ldc,
                             int M, int N, int K, int ab_dtype, int c_dtype, int act, int path, void *stream);

/*
 * Fused two-layer MLP / FFN on the tensor cores (bf16 operands, fp32 accumulate):
 *   C = act2( relu(X . W1^T + b1) . W2^T + b2 ) [* mul],   X (M,256) ldx, W1 (Hd,256), W2 (256,Hd), Hd %% 128 == 0,
 * C (M,256) ldc in c_dtype (F32 or BF16); the (M,Hd) hidden activation stays in shared memory / TMEM.
 * Replaces linear2(activation(linear1(x))) of models/deformable_encoder.py:97-107, models/deformable_decoder.py:263-273,
 * models/ffn.py:15-22 and the 256-input two-layer MLPs (models/mlp.py:22-25).
 */
MEMOTR_API int memotr_mlp2(const void *X, int ldx, const void *W1, const float *b1, const void *W2, const float *b2,
                           const void *mul, int ldmul, void *C, int ldc, int M, int K1, int Hd, int N2, int c_dtype,
                           int act2, void *stream);

/* The encoder FFN with its LayerNorm in the epilogue (models/deformable_encoder.py:103-107 followed by :128-131):
 *   y = LayerNorm(res + relu(X W1^T + b1) W2^T + b2) * gamma + beta      (256 columns; the whole row sits in one CTA's TMEM)
 * written three ways, as the next layer reads it: y (M,256) bf16 ldy; y32 fp32 ld32 (NULL: skip); ypos = y + pos, bf16
 * (pos / ypos both NULL: skip).  X (M,256) bf16, res (M,256) fp32.  One launch of ceil(M/128) CTAs, one CTA per SM: callers
 * with more row tiles than SMs pass the first round here and run the remaining rows through memotr_mlp2 (which splits the
 * hidden dimension of few-tile launches over the idle SMs) + memotr_layernorm. */
MEMOTR_API int memotr_mlp2_lnout(const void *X, int ldx, const void *W1, const float *b1, const void *W2, const float *b2,
                                 const float *res, int ldres, const float *gamma, const float *beta, float eps, void *y, int ldy,
                                 float *y32, int ld32, const void *pos, int ldpos, void *ypos, int ldypos, int M, int Hd,
                                 void *stream);

/* The dense half of an encoder layer in ONE kernel per 128-row tile (models/deformable_encoder.py:124-131 with
 * models/ops/modules/ms_deform_attn.py:129):
 *   x = LayerNorm1(src32 + att Wout^T + bout)                      front GEMM + LayerNorm, x never leaves the SM as a GEMM operand
 *   y = LayerNorm2(x + relu(x W1^T + b1) W2^T + b2)                fused FFN + LayerNorm epilogue (memotr_mlp2_lnout)
 * att (M,256) bf16 = the gather's output rows; src32 (M,256) fp32 residual stream (may alias y32: a tile reads its rows
 * before it writes them); outputs x32 (fp32 x, residual of norm2), y (bf16), y32 (fp32), ypos = y + pos (bf16); pre (M,256)
 * fp32 scratch for the row tiles beyond one round of SMs, which run with the hidden dimension split and get LayerNorm2 from
 * memotr_layernorm.  Replaces output_proj + norm1 + linear1/relu/linear2 + norm2: four launches and ~140 MB of HBM round
 * trips per layer. */
MEMOTR_API int memotr_encoder_dense_block(const void *att, int ldatt, const void *Wout, const float *bout, const float *src32,
                                          int ldsrc, const float *gamma1, const float *beta1, float *x32, int ldx32,
                                          const void *W1, const float *b1, const void *W2, const float *b2, const float *gamma2,
                                          const float *beta2, const void *pos, int ldpos, void *y, int ldy, float *y32, int ld32,
                                          void *ypos, int ldypos, float *pre, int ldpre, int M, int Hd, float eps, void *stream);

/*
 * y = LayerNorm(res + A W^T + bias) * gamma + beta for a 256 x 256 projection (A, W bf16; res fp32): y bf16, y32 fp32.
 * `src = norm1(src + output_proj(attn))` of models/deformable_encoder.py:124-126 + models/ops/modules/ms_deform_attn.py:129
 * as one tcgen05 kernel per 128-row tile (the product never leaves the SM).  Replaces memotr_linear + memotr_layernorm on
 * that pair.
 */
MEMOTR_API int memotr_linear256_layernorm(const void *A, int lda, const void *W, const float *bias, const float *res, int ldres,
                                          const float *gamma, const float *beta, float eps, void *y, int ldy, float *y32, int ld32,
                                          int M, void *stream);

/*
 * C = act(A W^T + bias) with fp32 A (M,K), fp32 C and fp32-ACCURATE products on the tensor cores: both operands are split into
 * two fp16 terms and the three significant products are summed in fp32 (one fp16 GEMM with 3K; error ~2^-22 per product).
 * W3 (N, 3K) fp16 = [hi | lo | hi] of (2^s W) packed once by the host, w_scale_inv = 2^-s; scratch_a3: M x 3K fp16.
 * The nn.Linear call sites of the encoder in the engine's "fp32tc" mode (the reference runs them in fp32 with TF32 off,
 * main.py:96-97).  Needs N % 64 == 0 and K % 64 == 0.
 */
MEMOTR_API int memotr_linear_f32x3(const float *A /* NULL: scratch_a3 already holds the split operand */, int lda, const void *W3,
                                   const float *bias, const unsigned char *rowzero, float *C, int ldc, int M, int N, int K, int act,
                                   float w_scale_inv, void *scratch_a3,
                                   void *split_out /* NULL, or (M, 3N) fp16: the result as the split operand of the next call */,
                                   void *stream);

/*
 * Input projections in front of the transformer (models/memotr.py:66-78,107-123), batch 1, fp32, channel-major maps (C, H*W):
 *   memotr_conv_gemm     Y (M, N) = W (M, K) X (K, N) + bias[m]: a 1x1 convolution (M = C_out, K = C_in, N = pixels), or the
 *                        3x3 / stride 2 convolution over the im2col buffer (K = 9 C_in)
 *   memotr_im2col_3x3s2  col (9 C, Ho Wo) of x (C, H, W) for kernel 3, stride 2, padding 1; Ho = (H - 1) / 2 + 1
 *   memotr_groupnorm_cm  GroupNorm(groups, C) in place on a channel-major (C, P) map (biased variance, eps as given)
 */
MEMOTR_API int memotr_conv_gemm(const float *W, const float *X, const float *bias, float *Y, int M, int N, int K, void *stream);
MEMOTR_API int memotr_im2col_3x3s2(const float *x, int C, int H, int W, float *col, void *stream);
MEMOTR_API int memotr_groupnorm_cm(float *x, const float *gamma, const float *beta, int groups, int C, int P, float eps, void *stream);

/* Profiling hook (tools/micro_dense.py), not part of the reference surface: every later memotr_mlp2* / encoder_dense_block
 * launch writes 8 clock64 stamps per CTA into `buf` (device int64[8 x CTAs]); null switches it off again. */
MEMOTR_API int memotr_mlp2_debug_stamps(long long *buf);
/* same for the persistent tcgen05 GEMM (memotr_linear with more tiles than SMs, memotr_linear_msda_prep): 20 stamps per CTA */
MEMOTR_API int memotr_gemm_debug_stamps(long long *buf);

/*
 * y = LayerNorm(x [+ x2]) (C == 256, eps as given, affine fp32); optional ypos = y + pos and fp32 copy y32.
 * models/deformable_encoder.py:124-130, models/deformable_decoder.py:251-252,313-318, models/ffn.py:23-24,
 * models/query_updater.py:126-133.
 */
MEMOTR_API int memotr_layernorm(const void *x, int x_dtype, int ldx, const void *x2, int ldx2, const float *gamma,
                                const float *beta, float eps, void *y, int y_dtype, int ldy, const void *pos,
                                int ldpos, void *ypos, int ldypos, float *y32, int ldy32, int M, int C, void *stream);

/*
 * Multi-head attention core (head_dim 32): O = softmax(Q K^T / sqrt(32) + key_padding_mask) V per head.
 * Q (Nq, n_heads*32) ldq, K/V (Nk, n_heads*32) in in_dtype; O in out_dtype; key_padding_mask (Nk) uint8 or NULL
 * (1 = ignore key).  The bf16 engine keeps the projected Q/K/V in fp32 (in_dtype F32) and takes O as bf16.
 * nn.MultiheadAttention math path at models/deformable_decoder.py:245-252 and models/query_updater.py:125.
 */
MEMOTR_API int memotr_mha(const void *Q, int ldq, const void *K, int ldk, const void *V, int ldv,
                          const unsigned char *key_padding_mask, void *O, int ldo, int Nq, int Nk, int n_heads,
                          int head_dim, int in_dtype, int out_dtype, void *stream);

/* One pyramid level, (C,HW) fp32 maps -> token rows [row0,row0+HW): src^T, pos^T+level_embed, and their sum, in
 * `dtype`; src_tok32 (optional) additionally keeps src^T in fp32 (the residual stream of the bf16 engine).
 * models/deformable_transformer.py:200-216 (+ with_pos_embed, models/deformable_encoder.py:124). */
MEMOTR_API int memotr_tokens_from_nchw(const float *src, const float *pos, const float *level_embed, void *src_tok,
                                       void *pos_tok, void *q_tok, float *src_tok32, int C, int HW, int row0, int ld,
                                       int dtype, void *stream);

/* memotr_tokens_from_nchw with the position map evaluated in place: pos = PositionEmbeddingSine(mask) of this level
 * (models/position_embedding.py:23-49, see memotr_pos_embed_sine for dim_i / scale / scratch), never materialised */
MEMOTR_API int memotr_tokens_from_nchw_pe(const float *src, const unsigned char *mask, int H, int W, const float *dim_i,
                                          float scale, float *scratch, const float *level_embed, void *src_tok,
                                          void *pos_tok, void *q_tok, float *src_tok32, int C, int row0, int ld, int dtype,
                                          void *stream);

/* The two steps of memotr_tokens_from_nchw_pe separately: the normalised cumulative counts of ALL levels in one launch
 * (mask: the levels' masks concatenated, level l at level_start[l]; emb + 2 * level_start[l] receives its (2, H_l W_l) planes;
 * shapes_hw / level_start are host arrays), then one token kernel per level reading its planes. */
MEMOTR_API int memotr_pos_cumsum_levels(const unsigned char *mask, const int *shapes_hw, const int *level_start, int n_levels,
                                        float scale, float *emb, float *valid_ratios /* NULL or (L,2): as memotr_valid_ratio */,
                                        void *stream);
MEMOTR_API int memotr_tokens_from_nchw_emb(const float *src, const float *emb, const float *dim_i, const float *level_embed,
                                           void *src_tok, void *pos_tok, void *q_tok, float *src_tok32, int C, int HW,
                                           int row0, int ld, int dtype, void *stream);

/* memotr_tokens_from_nchw_emb for all levels in ONE launch (C % 64 == 0): srcs = host array of the levels' (C, H_l W_l) fp32
 * device pointers, level_embed (L, C); two channels (a sin / cos pair) per thread, 4-byte paired stores. */
MEMOTR_API int memotr_tokens_from_nchw_levels(const float *const *srcs, const float *emb, const float *dim_i, const float *level_embed,
                                              void *src_tok, void *pos_tok, void *q_tok, float *src_tok32, int C,
                                              const int *shapes_hw, const int *level_start, int n_levels, int ld, int dtype,
                                              void *stream);

/* valid ratio (w,h) of one level's (H,W) uint8 padding mask -- models/deformable_transformer.py:175-190 */
MEMOTR_API int memotr_valid_ratio(const unsigned char *mask, int H, int W, float *out2, void *stream);

/* PositionEmbeddingSine(normalize=True) of one level from its padding mask (models/position_embedding.py:23-43, as built by
 * :46-49 with num_pos_feats = hidden_dim / 2, temperature 20, scale 2 pi): out (2 * num_pos_feats, H * W) fp32, channels
 * [y features | x features]; dim_i (num_pos_feats) = temperature ** (2 * (i // 2) / num_pos_feats) from the host;
 * scratch (2 * H * W) fp32.  The reference runs this inside the backbone and, for the extra level, inside
 * MeMOTR.forward (models/memotr.py:121). */
MEMOTR_API int memotr_pos_embed_sine(const unsigned char *mask, int H, int W, const float *dim_i, int num_pos_feats,
                                     float scale, float *scratch, float *out, void *stream);

/* sine embedding of (N,4) boxes -> (N,512): models/utils.py:78-85; optional sigmoid first (query_updater.py:102) and
 * per-coordinate scale (deformable_decoder.py:82-91); dim_t = the 128 divisors 10000^(2*(i//2)/128). */
MEMOTR_API int memotr_sine_embed(const float *pts, int ldp, const float *scale4, int apply_sigmoid, const float *dim_t,
                                 void *out, int ldo, int N, int out_dtype, void *stream);

/* out = a + b, (M,N) with leading dimensions -- with_pos_embed (deformable_decoder.py:246,304), query_updater.py:121-122 */
MEMOTR_API int memotr_add(const void *a, int lda, const void *b, int ldb, void *out, int ldo, int M, int N, int dtype,
                          void *stream);

/* strided copy with dtype conversion (torch.cat / slicing / .to(dtype) on the hot path) */
MEMOTR_API int memotr_convert(const void *src, int src_dtype, int lds, void *dst, int dst_dtype, int ldd, int M, int N,
                              void *stream);

/* new = sigmoid(delta + inverse_sigmoid(ref)); ref_next rows [0,n_take) = new, others = ref
 * (models/deformable_decoder.py:139-159; the same expression yields pred_bboxes, models/memotr.py:147-160) */
MEMOTR_API int memotr_box_refine(const float *delta, const float *ref, float *new_ref, float *ref_next, int N,
                                 int n_take, void *stream);

/* op 0: sigmoid, op 1: inverse_sigmoid (utils/utils.py:61-74) */
MEMOTR_API int memotr_unary(const float *in, float *out, long n, int op, void *stream);

/* QueryUpdater first step: is_pos = max_c sigmoid(logits) > thr; ref_new = is_pos ? inverse_sigmoid(boxes) : ref_pts
 * (models/query_updater.py:84-85,99-102) */
MEMOTR_API int memotr_upd_prepare(const float *logits, int ncls, const float *boxes, const float *ref_pts, float thr,
                                  unsigned char *is_pos, float *ref_new, int Nt, void *stream);

/* QueryUpdater last step, in place on the fp32 track state (models/query_updater.py:135-147) */
MEMOTR_API int memotr_upd_finalize(const unsigned char *is_pos, const void *feat, int feat_dtype, int ldf,
                                   const float *out_e, float *query_embed, float *long_memory, float *last_output,
                                   float lam, int Nt, int C, void *stream);

/*
 * Tracker glue on the device (SURVEY.md section 8(f) "next" row 1).  The reference does this on the host with python
 * loops over TrackInstances (structures/track_instances.py:11-38).  A track table here is a fixed-capacity
 * structure-of-arrays in device memory; rows [0, *n_active) are live, in the reference's order.
 */
typedef struct memotr_track_table {
  long long *ids, *labels, *disappear_time;                      /* (capacity) int64, as TrackInstances */
  float *query_embed, *output_embed, *last_output, *long_memory; /* (capacity, C) */
  float *ref_pts, *boxes, *logits;                               /* (capacity, 4), (capacity, 4), (capacity, ncls) */
  int *n_active;                                                 /* device scalar */
} memotr_track_table;

/* the rows of MeMOTR.forward's output dict the tracker reads (models/memotr.py:180-195), (n_det + capacity) rows each:
 * detect queries first, then one row per track-table row */
typedef struct memotr_frame_outputs {
  const float *pred_logits, *pred_boxes, *outputs, *last_ref_pts, *aux_queries; /* aux_outputs[-1]["queries"] */
} memotr_frame_outputs;

/*
 * RuntimeTracker.update (models/runtime_tracker.py:29-101, use_motion=False) followed by
 * QueryUpdater.select_active_tracks in eval mode (models/query_updater.py:243-254), in place on `tracks`:
 *   previous track i: score = sigmoid(logits[n_det+i][labels[i]]); disappear_time = score < track_thresh ? +1 : 0;
 *                     id = -1 once disappear_time >= miss_tolerance; boxes/logits/output_embed <- this frame's rows;
 *   newborn: detect query j with max_c sigmoid(logits[j][c]) >= det_thresh: id = *max_obj_id + rank, label = argmax,
 *            ref_pts = last_ref_pts[j], output_embed = last_output = outputs[j], query_embed = long_memory = aux_queries[j];
 *   result = surviving previous tracks in order, then newborns in order; *max_obj_id advanced.
 * `scratch` is a second table of the same capacity (work space), `src_index` (capacity) int32 work space,
 * `track_pad` (capacity) receives 1 for rows >= *n_active (the key-padding mask of the track rows for the next frame).
 * Newborns that do not fit are dropped and counted in *overflow (device int, accumulated) -- the caller must check it.
 * No reference counterpart for the capacity: the reference's tensors grow.
 */
MEMOTR_API int memotr_tracker_update(const memotr_frame_outputs *frame, int n_det, int ncls, int C,
                                     const memotr_track_table *tracks, const memotr_track_table *scratch, int capacity,
                                     float det_thresh, float track_thresh, int miss_tolerance, long long *max_obj_id,
                                     int *src_index, unsigned char *track_pad, int *overflow, void *stream);

/* per-frame result rows (submit_engine.py:89-102): keep = live && max_c sigmoid(logits) > score_thresh &&
 * w*ori_w*h*ori_h > area_thresh; boxes cxcywh (normalised) -> xyxy in pixels.  All outputs have `capacity` rows. */
MEMOTR_API int memotr_tracker_results(const memotr_track_table *tracks, int capacity, int ncls, float score_thresh,
                                      float area_thresh, float ori_w, float ori_h, long long *ids, float *boxes_xyxy,
                                      float *scores, unsigned char *keep, void *stream);

/*
 * The whole DeformableDecoder (all layers) + per-layer box / class heads as one persistent kernel, bf16 engine
 * (models/deformable_decoder.py:56-160,276-319; models/memotr.py:147-162).  One CTA per block of 16 query rows, one grid
 * barrier per layer (keys / values of the self-attention), weights streamed from L2 in program order.
 * All pointers are device pointers; the struct itself lives on the host and is passed by value to the kernel.
 */
#define MEMOTR_DEC_MAX_LAYERS 8
typedef struct memotr_dec_gemm { /* one dense layer of the weight program, W (N, K) bf16 with N % 64 == 0, K % 256 == 0,
                                    packed as (N/64) x (K/256) slot images: K == 256: n-blocks in order; K > 256 (then
                                    N must be 256): k-slice-major, i.e. (k-slice, n-block) row-major;
                                    a slot image = 64 rows x 264 bf16 (256 weights of that row + 8 bf16 of padding) */
  const void *W;
  int ldw, N, K, pad_;             /* ldw: unused (kept for layout), N, K as above */
} memotr_dec_gemm;

typedef struct memotr_dec_layer {
  const float *qk_b, *v_b, *sao_b, *ol_b, *cao_b, *f1_b, *f2_b, *bb0_b, *bb1_b, *bb2_b, *cls_b; /* biases, fp32 */
  const float *n1_g, *n1_b, *n2_g, *n2_b, *n3_g, *n3_b;                                          /* LayerNorm */
  const void *bb2_w, *cls_w;  /* (4, 256) and (ncls, 256) bf16: the two skinny heads, read directly */
  const void *value;          /* this layer's value map: fp16, pixel-major, `value_stride` elements between pixels */
  float *tgt_out, *ref_out, *pred_box, *pred_logit; /* (nq,256) layer output, (nq,4) next reference, (nq,4), (nq,ncls) */
} memotr_dec_layer;

typedef struct memotr_dec_params {
  const memotr_dec_gemm *prog; /* device array, program order: per layer rph0, rph1, [qs0, qs1 if layer > 0], qk, v,
                                  sa_out, ol, ca_out, ffn1/ffn2 (in two halves of the hidden dimension when d_ffn > 1024), bb0, bb1 */
  int n_prog, n_layers, nq, nd, merge, ncls, n_levels, n_points, d_ffn, value_stride, np, pad_;
  const float *rph0_b, *rph1_b, *qs0_b, *qs1_b; /* biases of the shared ref_point_head / query_scale MLPs */
  const float *tgt_in, *ref_in;                 /* (nq,256) decoder input, (nq,4) initial reference boxes (sigmoid space) */
  const float *vr_scale4, *valid_ratios, *dim_t; /* unused, (n_levels,2), (128) sine temperatures */
  const unsigned char *query_pad;               /* (nq) key-padding mask or NULL */
  void *kbuf, *vbuf;                            /* scratch: 2 x (np,256) fp16 keys, 2 x (256,np) fp16 values (transposed) */
  unsigned int *barrier;                        /* scratch: one counter */
  long long *prof;                              /* NULL, or (blocks, n_layers, 16) clock64 phase stamps (measurement) */
  float *init_ref_out, *last_ref_out;           /* NULL or (nq,4): inverse_sigmoid of the references entering the first /
                                                   the last layer (memotr.py:183-187) */
  int shapes[16], lsi[8];                       /* (H, W) and first pixel of every level */
  memotr_dec_layer layers[MEMOTR_DEC_MAX_LAYERS];
} memotr_dec_params;

/*
 * The fused decoder: a 4-CTA thread-block cluster per 16-row block (csrc/decoder_cluster.cu): dense layers split over
 * output columns, attention / gather over heads, the FFN over the hidden dimension.  `prog` holds FOUR programs of
 * n_prog entries each (rank-major); rank r's entries are, per layer: ref_point_head.0 rows [64r,64r+64), .1 likewise,
 * [query_scale.0/.1 if layer > 0], q rows, k rows, v rows, self-attn out rows, sampling-offset rows [64r,+64), attention-
 * logit rows [32r,+32) padded to 64, cross-attn out rows, linear1 rows [F/4 r, +F/4), linear2 columns [F/4 r, +F/4)
 * (K padded to a multiple of 256), bbox_embed.0 / .1 rows.  Requires n_levels * n_points == 16.
 */
MEMOTR_API int memotr_decoder_forward_cluster(const memotr_dec_params *params, void *stream);

/*
 * The same decoder with ONE CTA per 16-row block (csrc/decoder_fused.cu): 25 CTAs at the DanceTrack size, ~600 us instead of
 * ~380 us, but 15 k instead of 38 k SM-microseconds -- the variant the frame-pipelined clip uses, where the decoder of frame k
 * shares the GPU with the encoder of frame k+1 and its latency hides behind it (FrameEngine.run_clip_pipelined).  `prog` is ONE
 * program in layer order: rph0, rph1, [qs0, qs1 if layer > 0], qk, v, sa_out, ol, ca_out, ffn1 / ffn2 (in two halves of the
 * hidden dimension when d_ffn > 1024), bb0, bb1; slot images as memotr_dec_gemm describes.
 */
MEMOTR_API int memotr_decoder_forward(const memotr_dec_params *params, void *stream);

/*
 * Cap the number of SMs the persistent kernels of LATER launches size their grids for (persistent tcgen05 GEMM, fused FFN,
 * windowed gather); 0 = all SMs.  Host-side state read at launch time (grids are baked into a captured graph): the
 * frame-pipelined clip launches the first encoder layers of frame k+1 with the SMs the decoder of frame k occupies left free.
 */
MEMOTR_API int memotr_set_sm_budget(int n_sm);

/*
 * QueryUpdater.update_tracks_embedding (models/query_updater.py:82-166, DAB branch) on a device-resident track table as one
 * persistent kernel (csrc/updater_cluster.cu; a 4-CTA cluster per 16 track rows, bf16 engine).  `prog`: FOUR programs of
 * n_prog = 14 entries (rank-major, slot images as in memotr_dec_gemm); rank r: confidence_weight_net.0/.1 rows [64r,+64),
 * short_memory_fusion.0 rows [128r,+128) (K = 512), .1 rows [64r,+64) (K = 512), query_pos_head.0 rows (K = 512), .1 rows,
 * memory_attn q / k / v / out_proj rows [64r,+64), memory_ffn.linear1 rows [F/4 r,+F/4), linear2 columns [F/4 r,+F/4),
 * query_feat_ffn.linear1 / linear2 likewise.  The table fields are updated in place; `feedback_*` (optional) receive the
 * next frame's track queries (ref_pts, query_embed).
 */
typedef struct memotr_upd_params {
  const memotr_dec_gemm *prog;
  int n_prog, nt, ncls, d_ffn, np, pad_;
  float update_thresh, long_memory_lambda;
  const float *conf0_b, *conf1_b, *fus0_b, *fus1_b, *ph0_b, *ph1_b, *q_b, *k_b, *v_b, *out_b, *mf1_b, *mf2_b, *ff1_b, *ff2_b;
  const float *mn_g, *mn_b, *mfn_g, *mfn_b, *fn_g, *fn_b, *ffn_g, *ffn_b; /* memory_norm, memory_ffn.norm, query_feat_norm, query_feat_ffn.norm */
  const float *dim_t;                  /* (128) sine temperatures, as memotr_sine_embed */
  const unsigned char *track_pad;      /* (nt) key-padding mask of the table rows, or NULL */
  const float *logits, *boxes, *output_embed;                  /* (nt,ncls), (nt,4), (nt,256): this frame's values */
  float *ref_pts, *query_embed, *long_memory, *last_output;    /* (nt,4), (nt,256) x3: updated in place */
  float *feedback_ref, *feedback_embed;                        /* NULL or (nt,4), (nt,256) */
  void *kbuf, *vbuf;                   /* scratch: (np,256) fp16, (256,np) fp16 */
  unsigned int *barrier;               /* scratch: one counter */
} memotr_upd_params;

MEMOTR_API int memotr_updater_forward_cluster(const memotr_upd_params *params, void *stream);

/*
 * Interval timer for measurement (bench.py): n CUDA events; memotr_timer_record enqueues event `idx` on `stream`
 * (as an external event-record node when the stream is being captured into a CUDA graph), memotr_timer_elapsed_ms
 * reads the time between two recorded events after the work has completed.  No reference counterpart (the reference
 * times with time.time(), train_engine.py:191,251-252).
 */
MEMOTR_API void *memotr_timer_create(int n_events);
MEMOTR_API void memotr_timer_destroy(void *timer);
MEMOTR_API int memotr_timer_record(void *timer, int idx, void *stream);
MEMOTR_API int memotr_timer_elapsed_ms(void *timer, int idx_start, int idx_stop, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* MEMOTR_B200_H_ */
