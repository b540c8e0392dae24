// msda_bwd.cu -- multi-scale deformable attention, backward, for sm_100a.
//
// Replaces ms_deformable_col2im_gpu_kernel_shm_blocksize_aware_reduce_v1<T,32> (the variant MeMOTR's D=32 hits),
// its siblings, ms_deform_attn_col2im_bilinear[_gm] and ms_deform_attn_cuda_backward
// (/root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:87-234, 301-920, 956-1327;
//  src/cuda/ms_deform_attn_cuda.cu:83-153).
//
// Gradient definitions (kept from the reference):
//   grad_value[corner]   += bilinear_w(corner) * grad_out * attn_weight                       (.cuh:125-152)
//   grad_attn_weight      = sum_c grad_out[c] * bilinear(value)[c]                            (.cuh:156, 377-393)
//   grad_sampling_loc.x   = sum_c W_l * d(bilinear)/dw * grad_out[c] * attn_weight            (.cuh:157)
//   grad_sampling_loc.y   = sum_c H_l * d(bilinear)/dh * grad_out[c] * attn_weight            (.cuh:158)
//
// B200 design for D == 32, fp32:  8 lanes x float4 own one (b,q,head) -- a warp covers 4 heads.  The reference
// runs 32-thread blocks, parks three scalars per thread in shared memory and lets thread 0 add 32 of them serially
// behind two __syncthreads per point; here the per-point sums are 3-step xor-shuffle reductions inside the 8-lane
// group (no shared memory, no barriers) and grad_value is accumulated with one 16-byte vector reduction
// (red.global.add.v4.f32, sm_90+) per corner per lane instead of four scalar atomics.
// grad_value accumulation order is not deterministic (it is not in the reference either).
#include "common.cuh"

namespace memotr {

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

__global__ void __launch_bounds__(256)
msda_bwd_f32_d32(const float *__restrict__ value, const int64_t *__restrict__ shapes,
                 const int64_t *__restrict__ lsi, const float *__restrict__ loc, const float *__restrict__ attn,
                 const float *__restrict__ grad_out, float *__restrict__ grad_value, float *__restrict__ grad_loc,
                 float *__restrict__ grad_attn, int S, int H, int L, int Lq, int K, long n_qh) {
  pdl_grid_sync();
  constexpr int D = 32;
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long qh_raw = tid >> 3;
  const bool active = qh_raw < n_qh;
  const long qh = active ? qh_raw : n_qh - 1;  // inactive tail lanes shadow a valid group (they never write)
  const int sub = (int)(tid & 7);
  const int m = (int)(qh % H);
  const int b = (int)((qh / H) / Lq);
  const int xs = H * D;
  const long pbase = qh * L * K;
  const float4 top = ldg_f4(grad_out + qh * D + sub * 4);

  for (int l = 0; l < L; ++l) {
    const int Hh = (int)__ldg(shapes + 2 * l), Ww = (int)__ldg(shapes + 2 * l + 1);
    const float Hf = (float)Hh, Wf = (float)Ww;
    const int ys = Ww * xs;
    const long lvl_off = ((long)b * S + __ldg(lsi + l)) * xs + m * D + sub * 4;
    const float *lvl = value + lvl_off;
    float *glvl = grad_value + lvl_off;
    for (int p = 0; p < K; ++p) {
      const long pi = pbase + l * K + p;
      const float2 xy = __ldg(reinterpret_cast<const float2 *>(loc) + pi);
      const float aw = __ldg(attn + pi);
      const float h_im = __fmaf_rn(xy.y, Hf, -0.5f), w_im = __fmaf_rn(xy.x, Wf, -0.5f);
      float ga = 0.f, gx = 0.f, gy = 0.f;
      if (h_im > -1.f && w_im > -1.f && h_im < Hf && w_im < Wf) {
        const float hfl = floorf(h_im), wfl = floorf(w_im);
        const int y0 = (int)hfl, x0 = (int)wfl;
        const float lh = h_im - hfl, lw = w_im - wfl, hh = 1.f - lh, hw = 1.f - lw;
        const bool y0ok = y0 >= 0, y1ok = y0 + 1 <= Hh - 1, x0ok = x0 >= 0, x1ok = x0 + 1 <= Ww - 1;
        const int o00 = y0 * ys + x0 * xs;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 v1 = (y0ok && x0ok) ? ldg_f4(lvl + o00) : z;
        const float4 v2 = (y0ok && x1ok) ? ldg_f4(lvl + o00 + xs) : z;
        const float4 v3 = (y1ok && x0ok) ? ldg_f4(lvl + o00 + ys) : z;
        const float4 v4 = (y1ok && x1ok) ? ldg_f4(lvl + o00 + ys + xs) : z;
        const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
        const float4 tg = make_float4(top.x * aw, top.y * aw, top.z * aw, top.w * aw);
        if (active) {
          if (y0ok && x0ok) red_add_v4(glvl + o00, w1 * tg.x, w1 * tg.y, w1 * tg.z, w1 * tg.w);
          if (y0ok && x1ok) red_add_v4(glvl + o00 + xs, w2 * tg.x, w2 * tg.y, w2 * tg.z, w2 * tg.w);
          if (y1ok && x0ok) red_add_v4(glvl + o00 + ys, w3 * tg.x, w3 * tg.y, w3 * tg.z, w3 * tg.w);
          if (y1ok && x1ok) red_add_v4(glvl + o00 + ys + xs, w4 * tg.x, w4 * tg.y, w4 * tg.z, w4 * tg.w);
        }
#define MSDA_BWD_CH(c)                                                              \
  {                                                                                 \
    const float gh_ = hw * (v3.c - v1.c) + lw * (v4.c - v2.c);                      \
    const float gw_ = hh * (v2.c - v1.c) + lh * (v4.c - v3.c);                      \
    const float val = w1 * v1.c + w2 * v2.c + w3 * v3.c + w4 * v4.c;                \
    ga += top.c * val;                                                              \
    gx += gw_ * tg.c;                                                               \
    gy += gh_ * tg.c;                                                               \
  }
        MSDA_BWD_CH(x) MSDA_BWD_CH(y) MSDA_BWD_CH(z) MSDA_BWD_CH(w)
#undef MSDA_BWD_CH
        gx *= Wf;
        gy *= Hf;
      }
#pragma unroll
      for (int s = 4; s >= 1; s >>= 1) {
        ga += __shfl_xor_sync(0xffffffffu, ga, s);
        gx += __shfl_xor_sync(0xffffffffu, gx, s);
        gy += __shfl_xor_sync(0xffffffffu, gy, s);
      }
      if (active && sub == 0) {
        grad_attn[pi] = ga;
        *reinterpret_cast<float2 *>(grad_loc + 2 * pi) = make_float2(gx, gy);
      }
    }
  }
}

// ---- generic: any D, float / double; one thread per (b,q,head,channel), global atomics ------------------------
template <typename T>
__global__ void __launch_bounds__(256)
msda_bwd_generic(const T *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
                 const T *__restrict__ loc, const T *__restrict__ attn, const T *__restrict__ grad_out,
                 T *__restrict__ grad_value, T *__restrict__ grad_loc, T *__restrict__ grad_attn, int S, int H, int D,
                 int L, int Lq, int K, long n_out) {
  pdl_grid_sync();
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < n_out; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % D);
    const long qh = idx / D;
    const int m = (int)(qh % H);
    const int b = (int)((qh / H) / Lq);
    const long xs = (long)H * D;
    const long pbase = qh * L * K;
    const T top = grad_out[idx];
    for (int l = 0; l < L; ++l) {
      const int Hh = (int)shapes[2 * l], Ww = (int)shapes[2 * l + 1];
      const long ys = Ww * xs;
      const long lvl_off = ((long)b * S + lsi[l]) * xs + m * D + c;
      const T *lvl = value + lvl_off;
      T *glvl = grad_value + lvl_off;
      for (int p = 0; p < K; ++p) {
        const long pi = pbase + l * K + p;
        const T lx = loc[pi * 2], ly = loc[pi * 2 + 1], aw = attn[pi];
        const T h_im = ly * (T)Hh - (T)0.5, w_im = lx * (T)Ww - (T)0.5;
        if (!(h_im > -1 && w_im > -1 && h_im < Hh && w_im < Ww)) continue;
        const T hfl = floor(h_im), wfl = floor(w_im);
        const int y0 = (int)hfl, x0 = (int)wfl;
        const T lh = h_im - hfl, lw = w_im - wfl, hh = (T)1 - lh, hw = (T)1 - lw;
        const bool y0ok = y0 >= 0, y1ok = y0 + 1 <= Hh - 1, x0ok = x0 >= 0, x1ok = x0 + 1 <= Ww - 1;
        const long o00 = y0 * ys + x0 * xs;
        const T v1 = (y0ok && x0ok) ? lvl[o00] : (T)0;
        const T v2 = (y0ok && x1ok) ? lvl[o00 + xs] : (T)0;
        const T v3 = (y1ok && x0ok) ? lvl[o00 + ys] : (T)0;
        const T v4 = (y1ok && x1ok) ? lvl[o00 + ys + xs] : (T)0;
        const T w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
        const T tg = top * aw;
        if (y0ok && x0ok) atomicAdd(glvl + o00, w1 * tg);
        if (y0ok && x1ok) atomicAdd(glvl + o00 + xs, w2 * tg);
        if (y1ok && x0ok) atomicAdd(glvl + o00 + ys, w3 * tg);
        if (y1ok && x1ok) atomicAdd(glvl + o00 + ys + xs, w4 * tg);
        const T gh_ = hw * (v3 - v1) + lw * (v4 - v2);
        const T gw_ = hh * (v2 - v1) + lh * (v4 - v3);
        const T val = w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;
        atomicAdd(grad_attn + pi, top * val);
        atomicAdd(grad_loc + 2 * pi, (T)Ww * gw_ * tg);
        atomicAdd(grad_loc + 2 * pi + 1, (T)Hh * gh_ * tg);
      }
    }
  }
}

}  // namespace memotr

using namespace memotr;

extern "C" int memotr_msda_backward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_idx,
                                    const void *sampling_loc, const void *attn_weight, const void *grad_output,
                                    void *grad_value, void *grad_sampling_loc, void *grad_attn_weight, int B, int S,
                                    int H, int D, int L, int Lq, int K, int dtype, void *stream) {
  MEMOTR_REQUIRE(B >= 0 && S > 0 && H > 0 && D > 0 && L > 0 && Lq >= 0 && K > 0, "msda_backward: bad sizes");
  MEMOTR_REQUIRE((long)B * S * H * D < (1L << 31), "msda_backward: value has >= 2^31 elements");
  if ((long)B * Lq == 0) return MEMOTR_OK;
  MEMOTR_REQUIRE(value && spatial_shapes && level_start_idx && sampling_loc && attn_weight && grad_output &&
                     grad_value && grad_sampling_loc && grad_attn_weight,
                 "msda_backward: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const long n_qh = (long)B * Lq * H;
  if (dtype == MEMOTR_F32 && D == 32 && aligned16(value) && aligned16(grad_value) && aligned16(grad_output) &&
      ((reinterpret_cast<uintptr_t>(sampling_loc) & 7u) == 0) &&
      ((reinterpret_cast<uintptr_t>(grad_sampling_loc) & 7u) == 0)) {
    const long threads = n_qh * 8;
    const int grid = (int)((threads + 255) / 256);
    MEMOTR_LAUNCH((msda_bwd_f32_d32), grid, 256, 0, st, (const float *)value, spatial_shapes, level_start_idx,
                                           (const float *)sampling_loc, (const float *)attn_weight,
                                           (const float *)grad_output, (float *)grad_value,
                                           (float *)grad_sampling_loc, (float *)grad_attn_weight, S, H, L, Lq, K,
                                           n_qh);
    return check_launch("msda_bwd_f32_d32");
  }
  const size_t esz = dtype == MEMOTR_F64 ? 8 : 4;
  MEMOTR_REQUIRE(dtype == MEMOTR_F32 || dtype == MEMOTR_F64, "msda_backward: dtype %d not supported (f32/f64 only)",
                 dtype);
  // the generic kernel accumulates the per-point gradients with atomics: clear them on the same stream first
  cudaError_t e = cudaMemsetAsync(grad_sampling_loc, 0, esz * n_qh * L * K * 2, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(grad_attn_weight, 0, esz * n_qh * L * K, st);
  if (e != cudaSuccess) return fail(MEMOTR_ECUDA, "msda_backward: memset: %s", cudaGetErrorString(e));
  const long n_out = n_qh * D;
  const int grid = (int)((n_out + 255) / 256 > (1L << 30) ? (1L << 30) : (n_out + 255) / 256);
  if (dtype == MEMOTR_F32)
    MEMOTR_LAUNCH((msda_bwd_generic<float>), grid, 256, 0, st, (const float *)value, spatial_shapes, level_start_idx,
                                                  (const float *)sampling_loc, (const float *)attn_weight,
                                                  (const float *)grad_output, (float *)grad_value,
                                                  (float *)grad_sampling_loc, (float *)grad_attn_weight, S, H, D, L,
                                                  Lq, K, n_out);
  else
    MEMOTR_LAUNCH((msda_bwd_generic<double>), grid, 256, 0, st, (const double *)value, spatial_shapes, level_start_idx,
                                                   (const double *)sampling_loc, (const double *)attn_weight,
                                                   (const double *)grad_output, (double *)grad_value,
                                                   (double *)grad_sampling_loc, (double *)grad_attn_weight, S, H, D,
                                                   L, Lq, K, n_out);
  return check_launch("msda_bwd_generic");
}
