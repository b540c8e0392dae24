/*
 * oracle/msda_oracle.c -- CPU restatement of the reference multi-scale deformable attention operator.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's library.  Nothing under memotr_b200/ links, imports or executes it.
 *
 * What it restates (all paths relative to /root/reference/models/ops/src/cuda/):
 *   forward   ms_deformable_im2col_gpu_kernel           ms_deform_im2col_cuda.cuh:237-299
 *             ms_deform_attn_im2col_bilinear            ms_deform_im2col_cuda.cuh:33-84
 *   backward  ..._shm_blocksize_aware_reduce_v1         ms_deform_im2col_cuda.cuh:301-403   (the D=32 variant)
 *             ms_deform_attn_col2im_bilinear            ms_deform_im2col_cuda.cuh:87-159
 *   host      ms_deform_attn_cuda_forward / _backward   ms_deform_attn_cuda.cu:20-80 / 83-153 (zero-init outputs,
 *             im2col_step chunking is a no-op for the arithmetic and is not restated)
 *
 * Semantics kept (SURVEY.md section 8a checklist):
 *   - sampling_loc[...,0] is x (width), [...,1] is y; spatial_shapes[l] = (H_l, W_l)          (.cuh:276-282)
 *   - pixel coordinate = loc*size - 0.5                                                        (.cuh:285-286)
 *   - a point contributes iff  h>-1 && w>-1 && h<H && w<W  (strict)                            (.cuh:288)
 *   - each of the four corners is bounds-checked on its own => zero padding                    (.cuh:56-78)
 *   - value is pixel-major (B,S,H,D); corner address ((y*W+x)*H + m)*D + c past the level start (.cuh:47-53,278)
 *   - accumulation order: level outer, point inner                                             (.cuh:272-296)
 *   - backward: grad_sampling_loc scaled by W (x) and H (y); grad_attn_weight = top_grad * bilinear summed
 *     over the D channels of the head in channel order (thread 0's serial sum, .cuh:377-393)   (.cuh:156-158)
 *
 * Rounding.  The *_fma entry points reproduce the exact fp32/fp64 operation sequence the reference kernel
 * executes once nvcc has contracted it (read from the SASS of oracle/_ref, sm_100a, nvcc 12.9):
 *     h_im = fma(loc_h, H, -0.5)                       w_im = fma(loc_w, W, -0.5)
 *     val  = fma(w4,v4, fma(w3,v3, fma(w1,v1, w2*v2)))
 *     col  = fma(weight, val, col)
 * so the forward result is bit-identical to the reference CUDA kernel (checked on the GPU in tests/).
 * The plain entry points round every operation separately (C semantics with -ffp-contract=off).
 * grad_value in the reference is accumulated with float atomics in a non-deterministic order; here the order is
 * fixed (b, q, head, channel, level, point, corner), so backward parity is to a tolerance, never bit-wise.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define MSDA_ORACLE_API __attribute__((visibility("default")))

#define DEFINE_ORACLE(T, SFX, FLOORF, FMAF)                                                                   \
                                                                                                               \
  /* one sampling point: bilinear read of channel c of head m (im2col_bilinear, .cuh:33-84) */                \
  static T bilinear_##SFX(const T *lvl, int Hh, int Ww, int nheads, int ch, T h, T w, int m, int c,            \
                          int use_fma) {                                                                       \
    const int y0 = (int)FLOORF(h), x0 = (int)FLOORF(w);                                                        \
    const int y1 = y0 + 1, x1 = x0 + 1;                                                                        \
    const T lh = h - (T)y0, lw = w - (T)x0;                                                                    \
    const T hh = (T)1 - lh, hw = (T)1 - lw;                                                                    \
    const int xs = nheads * ch, ys = Ww * xs, base = m * ch + c;                                               \
    T v1 = 0, v2 = 0, v3 = 0, v4 = 0;                                                                          \
    if (y0 >= 0 && x0 >= 0) v1 = lvl[y0 * ys + x0 * xs + base];                                                \
    if (y0 >= 0 && x1 <= Ww - 1) v2 = lvl[y0 * ys + x1 * xs + base];                                           \
    if (y1 <= Hh - 1 && x0 >= 0) v3 = lvl[y1 * ys + x0 * xs + base];                                           \
    if (y1 <= Hh - 1 && x1 <= Ww - 1) v4 = lvl[y1 * ys + x1 * xs + base];                                      \
    const T w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;                                            \
    if (use_fma) return FMAF(w4, v4, FMAF(w3, v3, FMAF(w1, v1, w2 * v2)));                                     \
    return ((w1 * v1 + w2 * v2) + w3 * v3) + w4 * v4;                                                          \
  }                                                                                                            \
                                                                                                               \
  MSDA_ORACLE_API void msda_oracle_forward_##SFX(                                                              \
      const T *value, const int64_t *shapes, const int64_t *lsi, const T *loc, const T *attn, int B, int S,    \
      int Hn, int D, int L, int Lq, int K, int use_fma, T *out) {                                              \
    for (int b = 0; b < B; ++b)                                                                                \
      for (int q = 0; q < Lq; ++q)                                                                             \
        for (int m = 0; m < Hn; ++m) {                                                                         \
          const long qh = ((long)b * Lq + q) * Hn + m;                                                         \
          const T *aw = attn + qh * L * K;                                                                     \
          const T *sl = loc + qh * L * K * 2;                                                                  \
          for (int c = 0; c < D; ++c) {                                                                        \
            T col = 0;                                                                                         \
            for (int l = 0; l < L; ++l) {                                                                      \
              const int Hh = (int)shapes[2 * l], Ww = (int)shapes[2 * l + 1];                                  \
              const T *lvl = value + ((long)b * S + lsi[l]) * Hn * D;                                          \
              for (int p = 0; p < K; ++p) {                                                                    \
                const T lw_ = sl[(l * K + p) * 2], lh_ = sl[(l * K + p) * 2 + 1], a = aw[l * K + p];           \
                const T h_im = use_fma ? FMAF(lh_, (T)Hh, (T)-0.5) : lh_ * (T)Hh - (T)0.5;                     \
                const T w_im = use_fma ? FMAF(lw_, (T)Ww, (T)-0.5) : lw_ * (T)Ww - (T)0.5;                     \
                if (h_im > -1 && w_im > -1 && h_im < Hh && w_im < Ww) {                                        \
                  const T v = bilinear_##SFX(lvl, Hh, Ww, Hn, D, h_im, w_im, m, c, use_fma);                   \
                  col = use_fma ? FMAF(a, v, col) : col + v * a;                                               \
                }                                                                                              \
              }                                                                                                \
            }                                                                                                  \
            out[qh * D + c] = col;                                                                             \
          }                                                                                                    \
        }                                                                                                      \
  }                                                                                                            \
                                                                                                               \
  /* backward, following col2im_bilinear (.cuh:87-159) and the per-point channel sum (.cuh:377-393).          \
     All three gradient buffers are zero-filled here, as ms_deform_attn_cuda.cu:121-123 does. */               \
  MSDA_ORACLE_API void msda_oracle_backward_##SFX(                                                             \
      const T *value, const int64_t *shapes, const int64_t *lsi, const T *loc, const T *attn,                  \
      const T *grad_out, int B, int S, int Hn, int D, int L, int Lq, int K, T *grad_value, T *grad_loc,        \
      T *grad_attn) {                                                                                          \
    memset(grad_value, 0, sizeof(T) * (size_t)B * S * Hn * D);                                                 \
    memset(grad_loc, 0, sizeof(T) * (size_t)B * Lq * Hn * L * K * 2);                                          \
    memset(grad_attn, 0, sizeof(T) * (size_t)B * Lq * Hn * L * K);                                             \
    for (int b = 0; b < B; ++b)                                                                                \
      for (int q = 0; q < Lq; ++q)                                                                             \
        for (int m = 0; m < Hn; ++m) {                                                                         \
          const long qh = ((long)b * Lq + q) * Hn + m;                                                         \
          for (int l = 0; l < L; ++l) {                                                                        \
            const int Hh = (int)shapes[2 * l], Ww = (int)shapes[2 * l + 1];                                    \
            const long lvl_off = ((long)b * S + lsi[l]) * Hn * D;                                              \
            const T *lvl = value + lvl_off;                                                                    \
            T *glvl = grad_value + lvl_off;                                                                    \
            for (int p = 0; p < K; ++p) {                                                                      \
              const long pi = qh * L * K + l * K + p;                                                          \
              const T lw_ = loc[pi * 2], lh_ = loc[pi * 2 + 1], a = attn[pi];                                  \
              const T h = lh_ * (T)Hh - (T)0.5, w = lw_ * (T)Ww - (T)0.5;                                      \
              if (!(h > -1 && w > -1 && h < Hh && w < Ww)) continue;                                           \
              const int y0 = (int)FLOORF(h), x0 = (int)FLOORF(w), y1 = y0 + 1, x1 = x0 + 1;                    \
              const T lh = h - (T)y0, lw = w - (T)x0, hh = (T)1 - lh, hw = (T)1 - lw;                          \
              const int xs = Hn * D, ys = Ww * xs;                                                             \
              const T w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;                                  \
              T sum_gx = 0, sum_gy = 0, sum_ga = 0;                                                            \
              for (int c = 0; c < D; ++c) {                                                                    \
                const int base = m * D + c;                                                                    \
                const T top = grad_out[qh * D + c];                                                            \
                const T tgv = top * a;                                                                         \
                T gh = 0, gw = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0;                                              \
                if (y0 >= 0 && x0 >= 0) {                                                                      \
                  const int o = y0 * ys + x0 * xs + base;                                                      \
                  v1 = lvl[o]; gh -= hw * v1; gw -= hh * v1; glvl[o] += w1 * tgv;                              \
                }                                                                                              \
                if (y0 >= 0 && x1 <= Ww - 1) {                                                                 \
                  const int o = y0 * ys + x1 * xs + base;                                                      \
                  v2 = lvl[o]; gh -= lw * v2; gw += hh * v2; glvl[o] += w2 * tgv;                              \
                }                                                                                              \
                if (y1 <= Hh - 1 && x0 >= 0) {                                                                 \
                  const int o = y1 * ys + x0 * xs + base;                                                      \
                  v3 = lvl[o]; gh += hw * v3; gw -= lh * v3; glvl[o] += w3 * tgv;                              \
                }                                                                                              \
                if (y1 <= Hh - 1 && x1 <= Ww - 1) {                                                            \
                  const int o = y1 * ys + x1 * xs + base;                                                      \
                  v4 = lvl[o]; gh += lw * v4; gw += lh * v4; glvl[o] += w4 * tgv;                              \
                }                                                                                              \
                const T val = ((w1 * v1 + w2 * v2) + w3 * v3) + w4 * v4;                                       \
                sum_ga += top * val;                                                                           \
                sum_gx += (T)Ww * gw * tgv;                                                                    \
                sum_gy += (T)Hh * gh * tgv;                                                                    \
              }                                                                                                \
              grad_attn[pi] = sum_ga;                                                                          \
              grad_loc[pi * 2] = sum_gx;                                                                       \
              grad_loc[pi * 2 + 1] = sum_gy;                                                                   \
            }                                                                                                  \
          }                                                                                                    \
        }                                                                                                      \
  }

DEFINE_ORACLE(float, f32, floorf, fmaf)
DEFINE_ORACLE(double, f64, floor, fma)
