// msda_fwd.cu -- multi-scale deformable attention, forward, for sm_100a.
//
// Replaces ms_deformable_im2col_gpu_kernel + ms_deform_attn_im2col_bilinear
// (/root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299 and :33-84) and the host wrapper
// ms_deform_attn_cuda_forward (src/cuda/ms_deform_attn_cuda.cu:20-80).
//
// Layout (unchanged from the reference): value (B,S,H,D) pixel-major; sampling_loc
// (B,Lq,H,L,K,2) with (x,y) last; attn_weight (B,Lq,H,L,K); output (B,Lq,H*D).
//
// Kernels
//   msda_fwd_vec<...>   D == 32.  A group of G lanes owns one (b,q,head): fp32 -> 8 lanes x float4, bf16 -> 4 lanes
//                       x 8 channels (one 16-byte load per lane per corner, i.e. one full 128 B / 64 B row of the
//                       head per corner).  All K points of a level are decoded first, then their 4K corner loads
//                       are issued back to back (memory-level parallelism), then the blend runs in the reference's
//                       order.  No shared memory, no atomics, every output element written exactly once.
//   msda_fwd_generic<T> any D, float or double: one thread per output scalar (the reference's mapping); used for
//                       odd channel counts and for the fp64 gradcheck path.
//
// fp32 rounding: the operation sequence below is the one the reference kernel executes after nvcc's FMA
// contraction (read from its sm_100a SASS):  h = fma(loc_y, H, -0.5);  val = fma(w4,v4, fma(w3,v3, fma(w1,v1, w2*v2)));
// col = fma(weight, val, col), levels outer / points inner.  Explicit __f*_rn intrinsics pin it, so the fp32
// output is bit-identical to the reference op (tests/test_msda_gpu.py checks this against oracle/_ref).
#include <cstdlib>

#include "common.cuh"

namespace memotr {

template <typename T>
struct LocIO;
template <>
struct LocIO<float> {
  static __device__ __forceinline__ float2 xy(const float *loc, long i) {
    return __ldg(reinterpret_cast<const float2 *>(loc) + i);
  }
  static __device__ __forceinline__ float w(const float *a, long i) { return __ldg(a + i); }
};
template <>
struct LocIO<__nv_bfloat16> {
  static __device__ __forceinline__ float2 xy(const __nv_bfloat16 *loc, long i) {
    const unsigned int raw = __ldg(reinterpret_cast<const unsigned int *>(loc) + i);
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&raw));
  }
  static __device__ __forceinline__ float w(const __nv_bfloat16 *a, long i) {
    const unsigned short raw = __ldg(reinterpret_cast<const unsigned short *>(a) + i);
    return __bfloat162float(*reinterpret_cast<const __nv_bfloat16 *>(&raw));
  }
};

// One lane's slice of a head row: CH channels held as floats.
template <typename T>
struct Row;
template <>
struct Row<float> {
  static constexpr int CH = 4;  // float4
  static __device__ __forceinline__ void load(const float *p, bool ok, float (&v)[4]) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) t = ldg_f4(p);
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float *p, const float (&v)[4]) {
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Row<__nv_bfloat16> {
  static constexpr int CH = 8;  // 8 x bf16 = 16 bytes
  static __device__ __forceinline__ void load(const __nv_bfloat16 *p, bool ok, float (&v)[8]) {
    uint4 t = make_uint4(0u, 0u, 0u, 0u);
    if (ok) t = __ldg(reinterpret_cast<const uint4 *>(p));
    bf16x8_to_f32(t, v);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float (&v)[8]) {
    *reinterpret_cast<uint4 *>(p) = f32x8_to_bf16(v);
  }
};

// One (b,q,head) group's work: `sub` is this lane's slice of the 32 channels.
// KT > 0: points per level known at compile time (fully unrolled, loads batched per level); KT == 0: runtime K.
// T: value/output element type; TL: sampling_loc / attn_weight element type; xs: elements between x-neighbouring
// pixels of `value` (H*D for the reference layout; larger when the value maps of several layers are interleaved).
template <typename T, typename TL, int KT>
__device__ __forceinline__ void msda_fwd_group(const T *__restrict__ value, const int64_t *__restrict__ shapes,
                                               const int64_t *__restrict__ lsi, const TL *__restrict__ loc,
                                               const TL *__restrict__ attn, T *__restrict__ out, int S, int H, int L,
                                               int Lq, int Kr, int xs, long qh, int sub) {
  constexpr int D = 32;
  constexpr int CH = Row<T>::CH;
  const int m = (int)(qh % H);
  const int b = (int)((qh / H) / Lq);
  const int K = KT ? KT : Kr;
  const long pbase = qh * L * K;

  float acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = 0.f;

  for (int l = 0; l < L; ++l) {
    const int Hh = (int)__ldg(shapes + 2 * l), Ww = (int)__ldg(shapes + 2 * l + 1);
    const float Hf = (float)Hh, Wf = (float)Ww;
    const int ys = Ww * xs;
    const T *lvl = value + ((long)b * S + __ldg(lsi + l)) * xs + m * D + sub * CH;

    if constexpr (KT > 0) {
      float w1[KT], w2[KT], w3[KT], w4[KT], aw[KT];
      int o00[KT];
      unsigned okm[KT];  // bit0..3: corner valid, bit4: point contributes
#pragma unroll
      for (int p = 0; p < KT; ++p) {
        const float2 xy = LocIO<TL>::xy(loc, pbase + l * KT + p);
        aw[p] = LocIO<TL>::w(attn, pbase + l * KT + p);
        const float h_im = __fmaf_rn(xy.y, Hf, -0.5f), w_im = __fmaf_rn(xy.x, Wf, -0.5f);
        const bool inside = h_im > -1.f && w_im > -1.f && h_im < Hf && w_im < Wf;
        const float hfl = floorf(h_im), wfl = floorf(w_im);
        const int y0 = (int)hfl, x0 = (int)wfl;
        const float lh = __fsub_rn(h_im, hfl), lw = __fsub_rn(w_im, wfl);
        const float hh = __fsub_rn(1.f, lh), hw = __fsub_rn(1.f, lw);
        w1[p] = __fmul_rn(hh, hw), w2[p] = __fmul_rn(hh, lw), w3[p] = __fmul_rn(lh, hw), w4[p] = __fmul_rn(lh, lw);
        const bool y0ok = y0 >= 0, y1ok = y0 + 1 <= Hh - 1, x0ok = x0 >= 0, x1ok = x0 + 1 <= Ww - 1;
        okm[p] = inside ? (16u | (y0ok && x0ok ? 1u : 0u) | (y0ok && x1ok ? 2u : 0u) | (y1ok && x0ok ? 4u : 0u) |
                           (y1ok && x1ok ? 8u : 0u))
                        : 0u;
        o00[p] = inside ? y0 * ys + x0 * xs : 0;
      }
      float v1[KT][CH], v2[KT][CH], v3[KT][CH], v4[KT][CH];
#pragma unroll
      for (int p = 0; p < KT; ++p) {
        const T *p00 = lvl + o00[p];
        Row<T>::load(p00, okm[p] & 1u, v1[p]);
        Row<T>::load(p00 + xs, okm[p] & 2u, v2[p]);
        Row<T>::load(p00 + ys, okm[p] & 4u, v3[p]);
        Row<T>::load(p00 + ys + xs, okm[p] & 8u, v4[p]);
      }
#pragma unroll
      for (int p = 0; p < KT; ++p) {
        const bool contributes = okm[p] & 16u;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const float val = __fmaf_rn(w4[p], v4[p][c],
                                      __fmaf_rn(w3[p], v3[p][c], __fmaf_rn(w1[p], v1[p][c], __fmul_rn(w2[p], v2[p][c]))));
          const float nxt = __fmaf_rn(aw[p], val, acc[c]);
          acc[c] = contributes ? nxt : acc[c];
        }
      }
    } else {
      for (int p = 0; p < K; ++p) {
        const float2 xy = LocIO<TL>::xy(loc, pbase + l * K + p);
        const float aw = LocIO<TL>::w(attn, pbase + l * K + p);
        const float h_im = __fmaf_rn(xy.y, Hf, -0.5f), w_im = __fmaf_rn(xy.x, Wf, -0.5f);
        if (h_im > -1.f && w_im > -1.f && h_im < Hf && w_im < Wf) {
          const float hfl = floorf(h_im), wfl = floorf(w_im);
          const int y0 = (int)hfl, x0 = (int)wfl;
          const float lh = __fsub_rn(h_im, hfl), lw = __fsub_rn(w_im, wfl);
          const float hh = __fsub_rn(1.f, lh), hw = __fsub_rn(1.f, lw);
          const float w1 = __fmul_rn(hh, hw), w2 = __fmul_rn(hh, lw), w3 = __fmul_rn(lh, hw), w4 = __fmul_rn(lh, lw);
          const bool y0ok = y0 >= 0, y1ok = y0 + 1 <= Hh - 1, x0ok = x0 >= 0, x1ok = x0 + 1 <= Ww - 1;
          const T *p00 = lvl + (y0 * ys + x0 * xs);
          float v1[CH], v2[CH], v3[CH], v4[CH];
          Row<T>::load(p00, y0ok && x0ok, v1);
          Row<T>::load(p00 + xs, y0ok && x1ok, v2);
          Row<T>::load(p00 + ys, y1ok && x0ok, v3);
          Row<T>::load(p00 + ys + xs, y1ok && x1ok, v4);
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const float val =
                __fmaf_rn(w4, v4[c], __fmaf_rn(w3, v3[c], __fmaf_rn(w1, v1[c], __fmul_rn(w2, v2[c]))));
            acc[c] = __fmaf_rn(aw, val, acc[c]);
          }
        }
      }
    }
  }
  Row<T>::store(out + qh * D + sub * CH, acc);
}

// Mapping A ("linear"): consecutive lane groups take consecutive (b,q,head).  Any Lq.
template <typename T, typename TL, int KT>
__global__ void __launch_bounds__(256)
msda_fwd_vec(const T *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
             const TL *__restrict__ loc, const TL *__restrict__ attn, T *__restrict__ out, int S, int H, int L, int Lq,
             int Kr, int xs, long n_qh) {
  pdl_grid_sync();
  constexpr int G = 32 / Row<T>::CH;  // lanes per (b,q,head)
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long qh = tid / G;
  if (qh >= n_qh) return;
  msda_fwd_group<T, TL, KT>(value, shapes, lsi, loc, attn, out, S, H, L, Lq, Kr, xs, qh, (int)(tid % G));
}

// ---- generic: any D, float / double; one thread per output scalar ------------------------------------------
template <typename T>
__device__ __forceinline__ T fma_t(T a, T b, T c);
template <>
__device__ __forceinline__ float fma_t<float>(float a, float b, float c) { return __fmaf_rn(a, b, c); }
template <>
__device__ __forceinline__ double fma_t<double>(double a, double b, double c) { return __fma_rn(a, b, c); }

template <typename T>
__global__ void __launch_bounds__(256)
msda_fwd_generic(const T *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
                 const T *__restrict__ loc, const T *__restrict__ attn, T *__restrict__ out, int S, int H, int D,
                 int L, int Lq, int K, long n_out) {
  pdl_grid_sync();
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < n_out; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % D);
    const long qh = idx / D;
    const int m = (int)(qh % H);
    const int b = (int)((qh / H) / Lq);
    const long xs = (long)H * D;
    const long pbase = qh * L * K;
    T col = 0;
    for (int l = 0; l < L; ++l) {
      const int Hh = (int)shapes[2 * l], Ww = (int)shapes[2 * l + 1];
      const long ys = Ww * xs;
      const T *lvl = value + ((long)b * S + lsi[l]) * xs + m * D + c;
      for (int p = 0; p < K; ++p) {
        const T lx = loc[(pbase + l * K + p) * 2], ly = loc[(pbase + l * K + p) * 2 + 1];
        const T aw = attn[pbase + l * K + p];
        const T h_im = fma_t<T>(ly, (T)Hh, (T)-0.5), w_im = fma_t<T>(lx, (T)Ww, (T)-0.5);
        if (h_im > -1 && w_im > -1 && h_im < Hh && w_im < Ww) {
          const T hfl = floor(h_im), wfl = floor(w_im);
          const int y0 = (int)hfl, x0 = (int)wfl;
          const T lh = h_im - hfl, lw = w_im - wfl, hh = (T)1 - lh, hw = (T)1 - lw;
          const bool y0ok = y0 >= 0, y1ok = y0 + 1 <= Hh - 1, x0ok = x0 >= 0, x1ok = x0 + 1 <= Ww - 1;
          const T *p00 = lvl + (y0 * ys + x0 * xs);
          const T v1 = (y0ok && x0ok) ? p00[0] : (T)0;
          const T v2 = (y0ok && x1ok) ? p00[xs] : (T)0;
          const T v3 = (y1ok && x0ok) ? p00[ys] : (T)0;
          const T v4 = (y1ok && x1ok) ? p00[ys + xs] : (T)0;
          const T w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
          const T val = fma_t<T>(w4, v4, fma_t<T>(w3, v3, fma_t<T>(w1, v1, w2 * v2)));
          col = fma_t<T>(aw, val, col);
        }
      }
    }
    out[idx] = col;
  }
}

template <typename T, typename TL>
static int launch_vec(const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc, const void *attn,
                      void *out, int B, int S, int H, int L, int Lq, int K, int xs, cudaStream_t st) {
  constexpr int G = 32 / Row<T>::CH;
  const long n_qh = (long)B * Lq * H;
  const long threads = n_qh * G;
  const int grid = (int)((threads + 255) / 256);
  auto a = [&](auto kern) {
    MEMOTR_LAUNCH((kern), grid, 256, 0, st, (const T *)value, shapes, lsi, (const TL *)loc, (const TL *)attn, (T *)out, S, H, L, Lq,
                               K, xs, n_qh);
  };
  switch (K) {
    case 1: a(msda_fwd_vec<T, TL, 1>); break;
    case 2: a(msda_fwd_vec<T, TL, 2>); break;
    case 4: a(msda_fwd_vec<T, TL, 4>); break;
    case 8: a(msda_fwd_vec<T, TL, 8>); break;
    default: a(msda_fwd_vec<T, TL, 0>); break;
  }
  return check_launch("msda_fwd_vec");
}

}  // namespace memotr

using namespace memotr;

extern "C" int memotr_msda_forward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_idx,
                                   const void *sampling_loc, const void *attn_weight, void *output, int B, int S,
                                   int H, int D, int L, int Lq, int K, int dtype, void *stream) {
  MEMOTR_REQUIRE(B >= 0 && S > 0 && H > 0 && D > 0 && L > 0 && Lq >= 0 && K > 0, "msda_forward: bad sizes");
  MEMOTR_REQUIRE((long)B * S * H * D < (1L << 31), "msda_forward: value has >= 2^31 elements");
  if ((long)B * Lq == 0) return MEMOTR_OK;
  MEMOTR_REQUIRE(value && spatial_shapes && level_start_idx && sampling_loc && attn_weight && output,
                 "msda_forward: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec_ok = (D == 32) && aligned16(value) && aligned16(output) &&
                      ((reinterpret_cast<uintptr_t>(sampling_loc) & 7u) == 0);
  if (dtype == MEMOTR_F32 && vec_ok)
    return launch_vec<float, float>(value, spatial_shapes, level_start_idx, sampling_loc, attn_weight, output, B, S, H,
                                    L, Lq, K, H * D, st);
  if (dtype == MEMOTR_BF16) {
    MEMOTR_REQUIRE(vec_ok, "msda_forward: bf16 requires D == 32 and 16-byte aligned buffers");
    return launch_vec<__nv_bfloat16, __nv_bfloat16>(value, spatial_shapes, level_start_idx, sampling_loc, attn_weight,
                                                    output, B, S, H, L, Lq, K, H * D, st);
  }
  const long n_out = (long)B * Lq * H * D;
  const int grid = (int)((n_out + 255) / 256 > (1L << 30) ? (1L << 30) : (n_out + 255) / 256);
  if (dtype == MEMOTR_F32) {
    MEMOTR_LAUNCH((msda_fwd_generic<float>), grid, 256, 0, st, (const float *)value, spatial_shapes, level_start_idx,
                                                  (const float *)sampling_loc, (const float *)attn_weight,
                                                  (float *)output, S, H, D, L, Lq, K, n_out);
  } else if (dtype == MEMOTR_F64) {
    MEMOTR_LAUNCH((msda_fwd_generic<double>), grid, 256, 0, st, (const double *)value, spatial_shapes, level_start_idx,
                                                   (const double *)sampling_loc, (const double *)attn_weight,
                                                   (double *)output, S, H, D, L, Lq, K, n_out);
  } else {
    return fail(MEMOTR_EINVAL, "msda_forward: unknown dtype %d", dtype);
  }
  return check_launch("msda_fwd_generic");
}

// =====================================================================================================================
// Engine entry points.  fp32 / bf16 value maps run the reference-order kernel above; fp16 value maps (the bf16 engine) run
// the packed-half gather below.
//
// msda_fwd_h16: fp16 value map, packed-half blending, bf16 output (arithmetic: msda_h16.cuh).  ncu on the fp32-widening
// kernels (profiles/r01_msda_variants_ncu.md): with a bf16 value map the gather is bound by the ALU pipe (72 % busy) --
// one integer op per element just to widen bf16 to fp32.  Storing the value map in fp16 (11-bit mantissa: MORE accurate
// than bf16 for this read-only intermediate; the value_proj GEMM epilogue writes it) lets the blend run as packed HFMA2 on
// 2 channels per instruction with no widening at all.  Used for decoder-shaped launches and for whatever the windowed
// encoder gather (msda_window.cu) cannot stage; every variant that lost to it in round 1 (decode-once records, 8x8 query
// tiles, pair-duplicated and head-major maps: profiles/r01_msda_variants_ncu.md) has been removed.
#include "msda_h16.cuh"

namespace memotr {

// SPLIT = 1: 4 lanes per (b,q,head) walk all levels (throughput shape).
// SPLIT = 4: 16 lanes per (b,q,head), each 4-lane subgroup takes every 4th level and the partial sums are combined with
//            two xor-shuffles per channel -- 4x the threads and a quarter of the dependent load->blend chain for the
//            decoder-shaped launch (400 queries: 50 CTAs of serial work otherwise, 12.4 us measured).
template <int KT, int SPLIT, bool H8>
__global__ void __launch_bounds__(256)
msda_fwd_h16(const __half *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
             const float *__restrict__ loc, const float *__restrict__ attn, __nv_bfloat16 *__restrict__ out, int S, int H,
             int L, int Lq, int Kr, int xs, int n_qh, int ld_loc, int ld_attn, int B) {
  // H8: eight heads (shifts instead of integer divisions -- this kernel is issue-sensitive); all index arithmetic is
  // 32-bit (the launcher checks n_qh * 16 < 2^31)
  // ld_loc / ld_attn: floats between consecutive queries of `loc` / `attn` (dense: H*L*K*2 and H*L*K; 3*H*L*K each when
  // both live in the [locations | weights] rows written by the prep epilogue of the projection GEMM)
  pdl_grid_sync();
  constexpr int D = 32, G = 4 * SPLIT;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int qh_raw = tid / G;
  const bool live = qh_raw < n_qh;
  if (SPLIT == 1 && !live) return;
  const int qh = live ? qh_raw : n_qh - 1;      // SPLIT > 1: tail lanes shadow a valid group (shuffles stay full-warp)
  const int sub = tid % 4, lvl0 = (tid % G) / 4;
  const int m = H8 ? (qh & 7) : qh % H;
  const int ql = H8 ? (qh >> 3) : qh / H;       // linear query index b * Lq + q
  const int b = B == 1 ? 0 : ql / Lq;
  const int K = KT ? KT : Kr;
  const float2 *locq = reinterpret_cast<const float2 *>(loc + (long)ql * ld_loc) + m * L * K;
  const float *attq = attn + (long)ql * ld_attn + m * L * K;
  const __half *vb = value + (long)b * S * xs + m * D + sub * 8;
  float acc0[8], acc1[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc0[c] = acc1[c] = 0.f;
  h16::gather_global<KT>(vb, h16::ShapesI64{shapes, lsi}, locq, attq, L, Kr, xs, lvl0, SPLIT, acc0, acc1);
#pragma unroll
  for (int c = 0; c < 8; ++c) acc0[c] += acc1[c];
  if constexpr (SPLIT > 1) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      acc0[c] += __shfl_xor_sync(0xffffffffu, acc0[c], 4);
      acc0[c] += __shfl_xor_sync(0xffffffffu, acc0[c], 8);
    }
    if (!live || lvl0 != 0) return;
  }
  *reinterpret_cast<uint4 *>(out + (long)qh * D + sub * 8) = f32x8_to_bf16(acc0);
}

int launch_h16(const void *value, const int64_t *shapes, const int64_t *lsi, const float *loc, const float *attn, void *out,
               int B, int S, int H, int L, int Lq, int K, int xs, cudaStream_t st, int ld_loc, int ld_attn) {
  if (!ld_loc) ld_loc = H * L * K * 2;
  if (!ld_attn) ld_attn = H * L * K;
  const long n_qh = (long)B * Lq * H;
  // too few groups to fill the GPU: spread the levels over lanes.  (The level-split sums round differently from the
  // level-pair order of msda_h16.cuh -- both within the bf16 bar; MEMOTR_MSDA_NO_SPLIT=1 keeps the sequential order, which is
  // what the bit-equality tests against the windowed gather use on small pyramids.)
  const char *ns = getenv("MEMOTR_MSDA_NO_SPLIT");
  const bool no_split = ns && ns[0] == '1';
  const bool split = !no_split && n_qh * 4 < (long)kNumSMs * 256 * 2;
  const int grid = (int)((n_qh * (split ? 16 : 4) + 255) / 256);
  const bool h8 = H == 8;
  if (n_qh * 16 >= (1L << 31)) return fail(MEMOTR_EINVAL, "msda_fwd_h16: more than 2^27 (query, head) pairs");
#define H16_ARGS (const __half *)value, shapes, lsi, loc, attn, (__nv_bfloat16 *)out, S, H, L, Lq, K, xs, (int)n_qh, ld_loc, \
                 ld_attn, B
#define H16_LAUNCH(KT_)                                                                       \
  if (split && h8) MEMOTR_LAUNCH((msda_fwd_h16<KT_, 4, true>), grid, 256, 0, st, H16_ARGS);      \
  else if (split) MEMOTR_LAUNCH((msda_fwd_h16<KT_, 4, false>), grid, 256, 0, st, H16_ARGS);      \
  else if (h8) MEMOTR_LAUNCH((msda_fwd_h16<KT_, 1, true>), grid, 256, 0, st, H16_ARGS);          \
  else MEMOTR_LAUNCH((msda_fwd_h16<KT_, 1, false>), grid, 256, 0, st, H16_ARGS)
  switch (K) {
    case 1: H16_LAUNCH(1); break;
    case 2: H16_LAUNCH(2); break;
    case 4: H16_LAUNCH(4); break;
    case 8: H16_LAUNCH(8); break;
    default: H16_LAUNCH(0); break;
  }
#undef H16_LAUNCH
#undef H16_ARGS
  return check_launch("msda_fwd_h16");
}

}  // namespace memotr

using namespace memotr;

// Engine variant: value/output in `dtype` (f32 or bf16; f16 value -> bf16 output), sampling locations and attention
// weights always fp32 (they come straight from memotr_msda_prep), D == 32, and an explicit pixel stride so the value maps
// of all decoder layers can live interleaved in one (S, n_layers*256) buffer written by a single GEMM.
extern "C" int memotr_msda_forward_ex(const void *value, int value_pixel_stride, const int64_t *spatial_shapes,
                                      const int64_t *level_start_idx, const float *sampling_loc,
                                      const float *attn_weight, void *output, int B, int S, int H, int L, int Lq, int K,
                                      int dtype, void *stream) {
  MEMOTR_REQUIRE(B >= 0 && S > 0 && H > 0 && L > 0 && Lq >= 0 && K > 0, "msda_forward_ex: bad sizes");
  MEMOTR_REQUIRE(value_pixel_stride >= H * 32, "msda_forward_ex: pixel stride < H*32");
  MEMOTR_REQUIRE((long)B * S * value_pixel_stride < (1L << 31), "msda_forward_ex: value spans >= 2^31 elements");
  if ((long)B * Lq == 0) return MEMOTR_OK;
  MEMOTR_REQUIRE(value && spatial_shapes && level_start_idx && sampling_loc && attn_weight && output,
                 "msda_forward_ex: null pointer");
  const int al = dtype == MEMOTR_F32 ? 4 : 8;
  MEMOTR_REQUIRE(aligned16(value) && aligned16(output) && value_pixel_stride % al == 0 &&
                     ((reinterpret_cast<uintptr_t>(sampling_loc) & 7u) == 0),
                 "msda_forward_ex: misaligned buffer");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MEMOTR_F16)  // fp16 value map -> bf16 output, packed-half blend
    return launch_h16(value, spatial_shapes, level_start_idx, sampling_loc, attn_weight, output, B, S, H, L, Lq, K,
                      value_pixel_stride, st, 0, 0);
  if (dtype == MEMOTR_F32)
    return launch_vec<float, float>(value, spatial_shapes, level_start_idx, sampling_loc, attn_weight, output, B, S, H,
                                    L, Lq, K, value_pixel_stride, st);
  if (dtype == MEMOTR_BF16)
    return launch_vec<__nv_bfloat16, float>(value, spatial_shapes, level_start_idx, sampling_loc, attn_weight, output,
                                            B, S, H, L, Lq, K, value_pixel_stride, st);
  return fail(MEMOTR_EINVAL, "msda_forward_ex: dtype must be f32, bf16 or f16");
}

// fp16-value-map gather with strided sampling locations / attention weights (both fp32): the rows written by
// memotr_linear_msda_prep hold [locations (H, L*K, 2) | weights (H, L*K)], ld_loc = ld_attn = 3*H*L*K.
extern "C" int memotr_msda_forward_strided(const void *value, int value_pixel_stride, const int64_t *spatial_shapes,
                                           const int64_t *level_start_idx, const float *sampling_loc, int ld_loc,
                                           const float *attn_weight, int ld_attn, void *output, int B, int S, int H, int L,
                                           int Lq, int K, void *stream) {
  MEMOTR_REQUIRE(B >= 0 && S > 0 && H > 0 && L > 0 && Lq >= 0 && K > 0, "msda_forward_strided: bad sizes");
  MEMOTR_REQUIRE(value_pixel_stride >= H * 32 && value_pixel_stride % 8 == 0, "msda_forward_strided: bad pixel stride");
  MEMOTR_REQUIRE((long)B * S * value_pixel_stride < (1L << 31), "msda_forward_strided: value spans >= 2^31 elements");
  if ((long)B * Lq == 0) return MEMOTR_OK;
  MEMOTR_REQUIRE(value && spatial_shapes && level_start_idx && sampling_loc && attn_weight && output,
                 "msda_forward_strided: null pointer");
  MEMOTR_REQUIRE(ld_loc >= H * L * K * 2 && ld_loc % 2 == 0 && ld_attn >= H * L * K && aligned16(value) && aligned16(output) &&
                     ((reinterpret_cast<uintptr_t>(sampling_loc) & 7u) == 0),
                 "msda_forward_strided: bad stride / alignment");
  return launch_h16(value, spatial_shapes, level_start_idx, sampling_loc, attn_weight, output, B, S, H, L, Lq, K,
                    value_pixel_stride, (cudaStream_t)stream, ld_loc, ld_attn);
}
