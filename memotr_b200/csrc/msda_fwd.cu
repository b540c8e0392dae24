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

// Mapping B ("tiled"), used when Lq == S, i.e. the queries ARE the pixels of the pyramid (encoder self-attention,
// deformable_encoder.py:124): a CTA takes an 8x8 patch of queries of one level for ONE head.  Neighbouring queries
// sample neighbouring pixels, so the patch's corner reads (64 queries x L*K points x 4 corners of 128 B) fall into a
// few-KB window of that head's value map and are served by L1 instead of L2.  Pure re-ordering of the work: every
// (b,q,head) is still computed exactly once by the same code, so results are identical to mapping A.
// Persistent-style grid: CTAs stride over (batch, head, tile); the tile table is derived on the device from
// spatial_shapes (the host never reads them -- no sync).
constexpr int kTile = 8;
template <typename T, typename TL, int KT>
__global__ void __launch_bounds__(kTile *kTile * (32 / Row<T>::CH))
msda_fwd_tiled(const T *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
               const TL *__restrict__ loc, const TL *__restrict__ attn, T *__restrict__ out, int B, int S, int H, int L,
               int Kr, int xs) {
  pdl_grid_sync();
  constexpr int G = 32 / Row<T>::CH;
  const int g = threadIdx.x / G, sub = threadIdx.x % G;  // g: query slot inside the 8x8 patch
  const int gy = g / kTile, gx = g % kTile;
  int n_tiles = 0;
  for (int l = 0; l < L; ++l)
    n_tiles += ceil_div((int)__ldg(shapes + 2 * l), kTile) * ceil_div((int)__ldg(shapes + 2 * l + 1), kTile);
  const long n_work = (long)B * H * n_tiles;
  for (long w = blockIdx.x; w < n_work; w += gridDim.x) {
    int t = (int)(w % n_tiles);
    const int m = (int)((w / n_tiles) % H);
    const int b = (int)(w / ((long)n_tiles * H));
    int l = 0, Hh = 0, Ww = 0, tx_n = 0;
    for (; l < L; ++l) {
      Hh = (int)__ldg(shapes + 2 * l), Ww = (int)__ldg(shapes + 2 * l + 1);
      tx_n = ceil_div(Ww, kTile);
      const int n = ceil_div(Hh, kTile) * tx_n;
      if (t < n) break;
      t -= n;
    }
    const int y = (t / tx_n) * kTile + gy, x = (t % tx_n) * kTile + gx;
    if (y < Hh && x < Ww) {
      const long q = __ldg(lsi + l) + (long)y * Ww + x;
      msda_fwd_group<T, TL, KT>(value, shapes, lsi, loc, attn, out, S, H, L, S, Kr, xs, ((long)b * S + q) * H + m, sub);
    }
  }
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
  // Mapping B measured SLOWER than mapping A on B200 (profiles/r01_micro_msda_v2_tiled.json: 125 vs 104 us on the
  // encoder-shaped call even with spatially coherent samples), i.e. the kernel is not bound by L1/L2 hit rates; it
  // stays available for experiments through MEMOTR_MSDA_MAPPING=tiled.
  const char *force = getenv("MEMOTR_MSDA_MAPPING");
  const bool tiled = force && force[0] == 't' && Lq == S;
  if (tiled) {
    // upper bound on the work items without reading the device-side shapes: every 8x8 patch holds >= 1 pixel
    const long max_work = (long)B * H * S;
    const int grid_t = (int)(max_work < (long)kNumSMs * 32 ? max_work : (long)kNumSMs * 32);
    auto at = [&](auto kern) {
      MEMOTR_LAUNCH((kern), grid_t, kTile * kTile * G, 0, st, (const T *)value, shapes, lsi, (const TL *)loc, (const TL *)attn,
                                                   (T *)out, B, S, H, L, K, xs);
    };
    switch (K) {
      case 1: at(msda_fwd_tiled<T, TL, 1>); break;
      case 2: at(msda_fwd_tiled<T, TL, 2>); break;
      case 4: at(msda_fwd_tiled<T, TL, 4>); break;
      case 8: at(msda_fwd_tiled<T, TL, 8>); break;
      default: at(msda_fwd_tiled<T, TL, 0>); break;
    }
    return check_launch("msda_fwd_tiled");
  }
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
// v2 gather ("decode once"), used by the engine entry point.
//
// ncu on v1 (profiles/r01_msda_fwd_v1_ncu.md): issue slots 80 % busy, L1 61 %, DRAM 7 % -- the kernel is bound by
// instruction issue, and more than half of its instructions are the per-point decode (floor, bilinear weights,
// bounds tests, addresses) that all G lanes of a group repeat.  Here a CTA first decodes every point of its
// (b,q,head) groups exactly once -- one point per thread, coalesced reads of sampling_loc / attn_weight -- into a
// shared-memory record of 4 clamped element offsets + 4 corner weights pre-multiplied by the attention weight (zero for
// out-of-range corners / non-contributing points, so the gather needs no predicates), then the lane groups stream the
// records (conflict-free 16-byte broadcasts) and do nothing but 16-byte loads and FMAs.
// Arithmetic differs from the reference order only by the pre-multiplication (<= 1 ulp per term); the bit-exact
// reference sequence stays in v1, which memotr_msda_forward (the MSDeformAttnFunction path) keeps using.
namespace memotr {

struct __align__(16) TapRec {
  int off[4];   // element offsets of the 4 corners from the (batch, head, lane) base pointer, always in range
  float c[4];   // bilinear weight x attention weight; 0 where the corner / point does not contribute
};

// raw 16-byte row slices as loaded (conversion to fp32 is deferred to the blend so that a batch in flight costs 4
// registers per corner for either element type)
template <typename T>
struct Raw;
template <>
struct Raw<float> {
  using type = float4;
  static __device__ __forceinline__ float4 ld(const float *p) { return ldg_f4(p); }
  static __device__ __forceinline__ void fma_into(float (&acc)[4], float w, const float4 &r) {
    acc[0] = fmaf(w, r.x, acc[0]), acc[1] = fmaf(w, r.y, acc[1]), acc[2] = fmaf(w, r.z, acc[2]), acc[3] = fmaf(w, r.w, acc[3]);
  }
};
template <>
struct Raw<__nv_bfloat16> {
  using type = uint4;
  static __device__ __forceinline__ uint4 ld(const __nv_bfloat16 *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
  static __device__ __forceinline__ void fma_into(float (&acc)[8], float w, const uint4 &r) {
    float f[8];
    bf16x8_to_f32(r, f);
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = fmaf(w, f[c], acc[c]);
  }
};

// U = points per batch.  The gather loop is software-pipelined by hand (ping-pong register batches A/B): the 4U
// 16-byte loads of batch i+1 are issued BEFORE the blend of batch i, so every warp keeps 4U loads in flight while it
// computes -- left to itself the compiler interleaves each load with its consumer (1-2 loads in flight per warp) and
// the kernel becomes latency-bound (measured: 106 us vs 78 us for v1 on the encoder-shaped bf16 launch).
template <typename T, int U>
__global__ void __launch_bounds__(256)
msda_fwd_v2(const T *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
            const float *__restrict__ loc, const float *__restrict__ attn, T *__restrict__ out, int S, int H, int L,
            int Lq, int K, int xs, long n_qh) {
  pdl_grid_sync();
  constexpr int D = 32, CH = Row<T>::CH, G = D / CH, GROUPS = 256 / G;
  using RawT = typename Raw<T>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int P = L * K;
  const int Pp = (P + U - 1) / U * U;   // records per group, padded with zero-weight taps to a multiple of U
  const int gstride = Pp * 8 + 4;       // words per group: records + 16 B pad => the groups of a warp hit disjoint banks
  float *recs = reinterpret_cast<float *>(smem_raw);
  const long qh0 = (long)blockIdx.x * GROUPS;

  // ---- phase 1: decode, one point per thread --------------------------------------------------------------------
  for (int idx = threadIdx.x; idx < GROUPS * Pp; idx += 256) {
    const int g = idx / Pp, i = idx - g * Pp;
    const long qh = qh0 + g;
    int o0 = 0, o1 = 0, o2 = 0, o3 = 0;
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
    if (qh < n_qh && i < P) {
      const int l = i / K;
      const int Hh = (int)__ldg(shapes + 2 * l), Ww = (int)__ldg(shapes + 2 * l + 1);
      const float2 xy = __ldg(reinterpret_cast<const float2 *>(loc) + qh * P + i);
      const float aw = __ldg(attn + qh * P + i);
      const float h_im = __fmaf_rn(xy.y, (float)Hh, -0.5f), w_im = __fmaf_rn(xy.x, (float)Ww, -0.5f);
      if (h_im > -1.f && w_im > -1.f && h_im < (float)Hh && w_im < (float)Ww) {
        const float hfl = floorf(h_im), wfl = floorf(w_im);
        const int y0 = (int)hfl, x0 = (int)wfl;
        const float lh = h_im - hfl, lw = w_im - wfl, hh = 1.f - lh, hw = 1.f - lw;
        const bool y0ok = y0 >= 0, y1ok = y0 + 1 <= Hh - 1, x0ok = x0 >= 0, x1ok = x0 + 1 <= Ww - 1;
        const int yc0 = max(y0, 0), yc1 = min(y0 + 1, Hh - 1), xc0 = max(x0, 0), xc1 = min(x0 + 1, Ww - 1);
        const int base = (int)__ldg(lsi + l) * xs;
        const int ys = Ww * xs;
        o0 = base + yc0 * ys + xc0 * xs;
        o1 = base + yc0 * ys + xc1 * xs;
        o2 = base + yc1 * ys + xc0 * xs;
        o3 = base + yc1 * ys + xc1 * xs;
        c0 = (y0ok && x0ok) ? hh * hw * aw : 0.f;
        c1 = (y0ok && x1ok) ? hh * lw * aw : 0.f;
        c2 = (y1ok && x0ok) ? lh * hw * aw : 0.f;
        c3 = (y1ok && x1ok) ? lh * lw * aw : 0.f;
      }
    }
    float4 *dst = reinterpret_cast<float4 *>(recs + g * gstride + i * 8);
    dst[0] = make_float4(__int_as_float(o0), __int_as_float(o1), __int_as_float(o2), __int_as_float(o3));
    dst[1] = make_float4(c0, c1, c2, c3);
  }
  __syncthreads();

  // ---- phase 2: gather + blend, ping-pong pipelined -------------------------------------------------------------
  const int g = threadIdx.x / G, sub = threadIdx.x % G;
  const long qh = qh0 + g;
  if (qh >= n_qh) return;
  const int m = (int)(qh % H);
  const int b = (int)((qh / H) / Lq);
  const T *vb = value + (long)b * S * xs + m * D + sub * CH;
  const float4 *rp = reinterpret_cast<const float4 *>(recs + g * gstride);
  float acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = 0.f;

  float4 wA[U], wB[U];
  RawT rA[U][4], rB[U][4];
#define MSDA_LOAD(W_, R_, i0)                                       \
  _Pragma("unroll") for (int u = 0; u < U; ++u) {                   \
    const float4 o = rp[((i0) + u) * 2];                            \
    W_[u] = rp[((i0) + u) * 2 + 1];                                 \
    R_[u][0] = Raw<T>::ld(vb + __float_as_int(o.x));                \
    R_[u][1] = Raw<T>::ld(vb + __float_as_int(o.y));                \
    R_[u][2] = Raw<T>::ld(vb + __float_as_int(o.z));                \
    R_[u][3] = Raw<T>::ld(vb + __float_as_int(o.w));                \
  }
#define MSDA_BLEND(W_, R_)                                          \
  _Pragma("unroll") for (int u = 0; u < U; ++u) {                   \
    Raw<T>::fma_into(acc, W_[u].x, R_[u][0]);                       \
    Raw<T>::fma_into(acc, W_[u].y, R_[u][1]);                       \
    Raw<T>::fma_into(acc, W_[u].z, R_[u][2]);                       \
    Raw<T>::fma_into(acc, W_[u].w, R_[u][3]);                       \
  }
  int i = 0;
  MSDA_LOAD(wA, rA, 0)
  while (true) {
    if (i + U < Pp) { MSDA_LOAD(wB, rB, i + U) }
    MSDA_BLEND(wA, rA)
    i += U;
    if (i >= Pp) break;
    if (i + U < Pp) { MSDA_LOAD(wA, rA, i + U) }
    MSDA_BLEND(wB, rB)
    i += U;
    if (i >= Pp) break;
  }
#undef MSDA_LOAD
#undef MSDA_BLEND
  Row<T>::store(out + qh * D + sub * CH, acc);
}

template <typename T, int U>
static int launch_v2u(const void *value, const int64_t *shapes, const int64_t *lsi, const float *loc, const float *attn,
                      void *out, int B, int S, int H, int L, int Lq, int K, int xs, cudaStream_t st, bool *handled) {
  constexpr int GROUPS = 256 / (32 / Row<T>::CH);
  const int Pp = (L * K + U - 1) / U * U;
  const size_t smem = (size_t)GROUPS * (Pp * 8 + 4) * sizeof(float);
  *handled = smem <= 160 * 1024;
  if (!*handled) return MEMOTR_OK;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(msda_fwd_v2<T, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) return fail(MEMOTR_ECUDA, "msda_fwd_v2: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const long n_qh = (long)B * Lq * H;
  const int grid = (int)((n_qh + GROUPS - 1) / GROUPS);
  MEMOTR_LAUNCH((msda_fwd_v2<T, U>), grid, 256, smem, st, (const T *)value, shapes, lsi, loc, attn, (T *)out, S, H, L, Lq, K, xs,
                                             n_qh);
  return check_launch("msda_fwd_v2");
}

template <typename T>
static int launch_v2(const void *value, const int64_t *shapes, const int64_t *lsi, const float *loc, const float *attn,
                     void *out, int B, int S, int H, int L, int Lq, int K, int xs, cudaStream_t st, bool *handled) {
  const char *u = getenv("MEMOTR_MSDA_U");  // tuning knob: points per in-flight batch (1, 2 or 4)
  const int U = u ? atoi(u) : 2;
  if (U == 1) return launch_v2u<T, 1>(value, shapes, lsi, loc, attn, out, B, S, H, L, Lq, K, xs, st, handled);
  if (U == 4) return launch_v2u<T, 4>(value, shapes, lsi, loc, attn, out, B, S, H, L, Lq, K, xs, st, handled);
  return launch_v2u<T, 2>(value, shapes, lsi, loc, attn, out, B, S, H, L, Lq, K, xs, st, handled);
}

}  // namespace memotr

namespace memotr {
static int launch_h16(const void *value, const int64_t *shapes, const int64_t *lsi, const float *loc, const float *attn,
                      void *out, int B, int S, int H, int L, int Lq, int K, int xs, cudaStream_t st, int ld_loc = 0,
                      int ld_attn = 0, long head_stride = 32);
}
using namespace memotr;

// Engine variant: value/output in `dtype` (f32 or bf16), sampling locations and attention weights always fp32 (they
// come straight from memotr_msda_prep), D == 32, and an explicit pixel stride so the value maps of all decoder layers
// can live interleaved in one (S, n_layers*256) buffer written by a single GEMM.
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
  if (dtype == MEMOTR_F16)  // fp16 value map -> bf16 output, packed-half blend (v4)
    return launch_h16(value, spatial_shapes, level_start_idx, sampling_loc, attn_weight, output, B, S, H, L, Lq, K,
                      value_pixel_stride, st);
  // v2 (decode-once) measured no faster than v1 on B200 (profiles/r01_micro_msda_v3_decode_once.json), so the
  // reference-order kernel stays the default; MEMOTR_MSDA_KERNEL=v2 [MEMOTR_MSDA_U=1|2|4] selects the experiment.
  const char *force = getenv("MEMOTR_MSDA_KERNEL");
  if (force && force[0] == 'v' && force[1] == '2') {
    bool handled = false;
    int rc = dtype == MEMOTR_F32
                 ? launch_v2<float>(value, spatial_shapes, level_start_idx, sampling_loc, attn_weight, output, B, S, H, L,
                                    Lq, K, value_pixel_stride, st, &handled)
                 : dtype == MEMOTR_BF16
                       ? launch_v2<__nv_bfloat16>(value, spatial_shapes, level_start_idx, sampling_loc, attn_weight,
                                                  output, B, S, H, L, Lq, K, value_pixel_stride, st, &handled)
                       : fail(MEMOTR_EINVAL, "msda_forward_ex: dtype must be f32 or bf16");
    if (rc != MEMOTR_OK || handled) return rc;
  }
  if (dtype == MEMOTR_F32)
    return launch_vec<float, float>(value, spatial_shapes, level_start_idx, sampling_loc, attn_weight, output, B, S, H,
                                    L, Lq, K, value_pixel_stride, st);
  if (dtype == MEMOTR_BF16)
    return launch_vec<__nv_bfloat16, float>(value, spatial_shapes, level_start_idx, sampling_loc, attn_weight, output,
                                            B, S, H, L, Lq, K, value_pixel_stride, st);
  return fail(MEMOTR_EINVAL, "msda_forward_ex: dtype must be f32 or bf16");
}

// fp16-value-map gather with strided sampling locations / attention weights (both fp32): the rows written by
// memotr_linear_msda_prep hold [locations (H, L*K, 2) | weights (H, L*K)], ld_loc = ld_attn = 3*H*L*K.
extern "C" int memotr_msda_forward_strided(const void *value, int value_pixel_stride, const int64_t *spatial_shapes,
                                           const int64_t *level_start_idx, const float *sampling_loc, int ld_loc,
                                           const float *attn_weight, int ld_attn, void *output, int B, int S, int H, int L,
                                           int Lq, int K, int head_major, void *stream) {
  MEMOTR_REQUIRE(B >= 0 && S > 0 && H > 0 && L > 0 && Lq >= 0 && K > 0, "msda_forward_strided: bad sizes");
  MEMOTR_REQUIRE(head_major ? (value_pixel_stride == 32 && B == 1) : (value_pixel_stride >= H * 32 && value_pixel_stride % 8 == 0),
                 "msda_forward_strided: bad pixel stride (head-major maps: 32, batch 1)");
  MEMOTR_REQUIRE((long)B * S * value_pixel_stride < (1L << 31), "msda_forward_strided: value spans >= 2^31 elements");
  if ((long)B * Lq == 0) return MEMOTR_OK;
  MEMOTR_REQUIRE(value && spatial_shapes && level_start_idx && sampling_loc && attn_weight && output,
                 "msda_forward_strided: null pointer");
  MEMOTR_REQUIRE(ld_loc >= H * L * K * 2 && ld_loc % 2 == 0 && ld_attn >= H * L * K && aligned16(value) && aligned16(output) &&
                     ((reinterpret_cast<uintptr_t>(sampling_loc) & 7u) == 0),
                 "msda_forward_strided: bad stride / alignment");
  return launch_h16(value, spatial_shapes, level_start_idx, sampling_loc, attn_weight, output, B, S, H, L, Lq, K,
                    value_pixel_stride, (cudaStream_t)stream, ld_loc, ld_attn, head_major ? (long)S * 32 : 32L);
}

// =====================================================================================================================
// v3 gather for the encoder-shaped launch (bf16): "pair-duplicated head-major" value map.
//
// The L1 cache serves one 128-byte line per clock per SM, and in the reference's pixel-major layout every corner of
// every head is its own line (4 lines per sampling point; profiles/r01_msda_fwd_v1_ncu.md).  memotr_msda_pairs_layout
// rewrites a value map once per layer as  pairs[h][s][2][32] (bf16):  entry s = pixel (y,x) of its level holds that
// pixel's 32 channels of head h followed by those of its RIGHT neighbour (y,x+1) (zeros at the end of a row), i.e. the
// two x-corners of a bilinear footprint are one aligned 128-byte line.  A group of 8 lanes then fetches BOTH corners of
// one footprint row with a single 16-byte load per lane: 2 lines and 2 load instructions per point instead of 4 and 4.
// Lanes 0-3 blend the x0 column, lanes 4-7 the x1 column; the halves are added with one xor-shuffle per channel at the end.
namespace memotr {

__global__ void __launch_bounds__(256)
msda_pairs_layout_kernel(const __nv_bfloat16 *__restrict__ value, int xs, const int64_t *__restrict__ shapes,
                         const int64_t *__restrict__ lsi, __nv_bfloat16 *__restrict__ pairs, int S, int H, int L) {
  pdl_grid_sync();
  // one thread per (s, head, half, 16-byte quarter): 8 threads move the 128-byte entry of one (s, head)
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)S * H * 8) return;
  const int part = (int)(idx & 7), h = (int)((idx >> 3) % H), s = (int)(idx / (8 * H));
  const int half = part >> 2, q = part & 3;
  int l = 0;
  for (int t = 1; t < L; ++t)
    if (s >= (int)lsi[t]) l = t;
  const int Ww = (int)shapes[2 * l + 1];
  const int x = (s - (int)lsi[l]) % Ww;
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (half == 0 || x + 1 < Ww) v = __ldg(reinterpret_cast<const uint4 *>(value + (long)(s + half) * xs + h * 32 + q * 8));
  *reinterpret_cast<uint4 *>(pairs + (((long)h * S + s) * 2 + half) * 32 + q * 8) = v;
}

template <int KT>
__global__ void __launch_bounds__(256)
msda_fwd_pairs(const __nv_bfloat16 *__restrict__ pairs, const int64_t *__restrict__ shapes,
               const int64_t *__restrict__ lsi, const float *__restrict__ loc, const float *__restrict__ attn,
               __nv_bfloat16 *__restrict__ out, int S, int H, int L, int Lq, long n_qh) {
  pdl_grid_sync();
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long qh_raw = tid >> 3;
  const bool live = qh_raw < n_qh;
  const long qh = live ? qh_raw : n_qh - 1;            // tail lanes shadow a valid group (shuffles stay full-warp)
  const int sub = (int)(tid & 7), xsel = sub >> 2, c8 = sub & 3;
  const int m = (int)(qh % H);
  const __nv_bfloat16 *plane = pairs + (long)m * S * 64 + c8 * 8;   // + entry * 64 + half * 32
  const long pbase = qh * L * KT;
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
  for (int l = 0; l < L; ++l) {
    const int Hh = (int)__ldg(shapes + 2 * l), Ww = (int)__ldg(shapes + 2 * l + 1);
    const float Hf = (float)Hh, Wf = (float)Ww;
    const int base = (int)__ldg(lsi + l);
    float w0[KT], w1[KT];
    int o0[KT], o1[KT];
#pragma unroll
    for (int p = 0; p < KT; ++p) {
      const float2 xy = __ldg(reinterpret_cast<const float2 *>(loc) + pbase + l * KT + p);
      const float aw = __ldg(attn + pbase + l * KT + p);
      const float h_im = __fmaf_rn(xy.y, Hf, -0.5f), w_im = __fmaf_rn(xy.x, Wf, -0.5f);
      const bool inside = h_im > -1.f && w_im > -1.f && h_im < Hf && w_im < Wf;
      const float hfl = floorf(h_im), wfl = floorf(w_im);
      const int y0 = (int)hfl, x0 = (int)wfl;
      const float lh = h_im - hfl, lw = w_im - wfl;
      const float wx = (xsel ? lw : 1.f - lw) * aw;          // this lane's column weight (x0 or x1)
      const int xc = x0 + xsel;                              // this lane's column
      const bool xok = inside && xc >= 0 && xc <= Ww - 1;
      // entry = pixel max(x0,0); its half 0 is column max(x0,0), half 1 the column to its right
      const int xe = max(x0, 0), hsel = xc - xe;             // hsel in {0,1} whenever xok
      const int yc0 = min(max(y0, 0), Hh - 1), yc1 = min(max(y0 + 1, 0), Hh - 1);
      const int xec = min(xe, Ww - 1), hs = xok ? hsel : 0;
      o0[p] = ((base + yc0 * Ww + xec) * 2 + hs) * 32;
      o1[p] = ((base + yc1 * Ww + xec) * 2 + hs) * 32;
      w0[p] = (xok && y0 >= 0) ? (1.f - lh) * wx : 0.f;
      w1[p] = (xok && y0 + 1 <= Hh - 1) ? lh * wx : 0.f;
    }
    uint4 r0[KT], r1[KT];
#pragma unroll
    for (int p = 0; p < KT; ++p) {
      r0[p] = __ldg(reinterpret_cast<const uint4 *>(plane + o0[p]));
      r1[p] = __ldg(reinterpret_cast<const uint4 *>(plane + o1[p]));
    }
#pragma unroll
    for (int p = 0; p < KT; ++p) {
      float f0[8], f1[8];
      bf16x8_to_f32(r0[p], f0);
      bf16x8_to_f32(r1[p], f1);
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = fmaf(w1[p], f1[c], fmaf(w0[p], f0[c], acc[c]));
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 4);   // x0 half + x1 half
  if (live && xsel == 0) *reinterpret_cast<uint4 *>(out + qh * 32 + c8 * 8) = f32x8_to_bf16(acc);
}

}  // namespace memotr

extern "C" int memotr_msda_pairs_layout(const void *value, int value_pixel_stride, const int64_t *spatial_shapes,
                                        const int64_t *level_start_idx, void *pairs, int S, int H, int L, void *stream) {
  MEMOTR_REQUIRE(value && spatial_shapes && level_start_idx && pairs && S > 0 && H > 0 && L > 0,
                 "msda_pairs_layout: bad arguments");
  MEMOTR_REQUIRE(value_pixel_stride >= H * 32 && value_pixel_stride % 8 == 0 && aligned16(value) && aligned16(pairs),
                 "msda_pairs_layout: misaligned buffer");
  const long n = (long)S * H * 8;
  MEMOTR_LAUNCH((msda_pairs_layout_kernel), (int)((n + 255) / 256), 256, 0, (cudaStream_t)stream,
                (const __nv_bfloat16 *)value, value_pixel_stride, spatial_shapes, level_start_idx, (__nv_bfloat16 *)pairs,
                S, H, L);
  return check_launch("msda_pairs_layout");
}

extern "C" int memotr_msda_forward_pairs(const void *pairs, const int64_t *spatial_shapes, const int64_t *level_start_idx,
                                         const float *sampling_loc, const float *attn_weight, void *output, int S, int H,
                                         int L, int Lq, int K, void *stream) {
  MEMOTR_REQUIRE(pairs && spatial_shapes && level_start_idx && sampling_loc && attn_weight && output && S > 0 && H > 0 &&
                     L > 0 && Lq >= 0,
                 "msda_forward_pairs: bad arguments");
  MEMOTR_REQUIRE((long)S * H * 64 < (1L << 31), "msda_forward_pairs: value map too large");
  MEMOTR_REQUIRE(aligned16(pairs) && aligned16(output) && ((reinterpret_cast<uintptr_t>(sampling_loc) & 7u) == 0),
                 "msda_forward_pairs: misaligned buffer");
  if (Lq == 0) return MEMOTR_OK;
  const long n_qh = (long)Lq * H;
  const int grid = (int)((n_qh * 8 + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  using bf = __nv_bfloat16;
#define PAIRS_LAUNCH(KT_)                                                                                       \
  MEMOTR_LAUNCH((msda_fwd_pairs<KT_>), grid, 256, 0, st, (const bf *)pairs, spatial_shapes, level_start_idx,   \
                sampling_loc, attn_weight, (bf *)output, S, H, L, Lq, n_qh)
  switch (K) {
    case 1: PAIRS_LAUNCH(1); break;
    case 2: PAIRS_LAUNCH(2); break;
    case 4: PAIRS_LAUNCH(4); break;
    case 8: PAIRS_LAUNCH(8); break;
    default: return fail(MEMOTR_ENOSYS, "msda_forward_pairs: K must be 1, 2, 4 or 8 (got %d)", K);
  }
#undef PAIRS_LAUNCH
  return check_launch("msda_fwd_pairs");
}

// =====================================================================================================================
// v4 gather: fp16 value map, packed-half blending.
//
// ncu on v1/v2/v3 (profiles/r01_msda_variants_ncu.md): with a bf16 value map the kernel is bound by the ALU pipe
// (72 % busy) -- one integer op per element just to widen bf16 to fp32 (32 per point per lane), on top of addressing --
// whatever the thread mapping or layout.  Storing the value map in fp16 (11-bit mantissa: MORE accurate than bf16 for
// this read-only intermediate; the value_proj GEMM epilogue writes it) lets the 4-corner blend run as packed HFMA2 on
// 2 channels per instruction with no widening at all; the fp16 partial sum of one level's K points is widened and
// added into the fp32 accumulator once per level.  Corner weights (bilinear x attention, <= 1) are rounded to fp16;
// out-of-range corners get weight 0 and a clamped address, so the loads are unpredicated.
namespace memotr {

// SPLIT = 1: 4 lanes per (b,q,head) walk all levels (throughput shape: encoder, Lq = S).
// SPLIT = 4: 16 lanes per (b,q,head), each 4-lane subgroup takes every 4th level and the partial sums are combined with
//            two xor-shuffles per channel -- 4x the threads and a quarter of the dependent load->blend chain for the
//            decoder-shaped launch (400 queries: 50 CTAs of serial work otherwise, 12.4 us measured).
#ifndef MEMOTR_H16_MINB
#define MEMOTR_H16_MINB 1   // (5 -> 48 registers, five CTAs per SM: measured neutral)
#endif
template <int KT, int SPLIT, bool HM, bool H8>
__global__ void __launch_bounds__(256, MEMOTR_H16_MINB)
msda_fwd_h16(const __half *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
             const float *__restrict__ loc, const float *__restrict__ attn, __nv_bfloat16 *__restrict__ out, int S, int H,
             int L, int Lq, int Kr, int xs, int n_qh, int ld_loc, int ld_attn, long head_stride, int B) {
  // H8: eight heads (shifts instead of integer divisions -- this kernel is issue-sensitive: a runtime head stride alone
  // cost 9 %); all index arithmetic is 32-bit (the launcher checks n_qh * 16 < 2^31)
  // head_stride: elements between the heads of one pixel (32 in the pixel-major map; S * 32 with xs = 32 in the head-major map)
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
  // HM: head-major map (head_stride = S * 32, xs = 32, batch 1); otherwise the pixel-major layout with compile-time head offset
  const __half *vb = HM ? value + m * head_stride + sub * 8 : value + (long)b * S * xs + m * D + sub * 8;
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;

  for (int l = lvl0; l < L; l += SPLIT) {
    const int Hh = (int)__ldg(shapes + 2 * l), Ww = (int)__ldg(shapes + 2 * l + 1);
    const float Hf = (float)Hh, Wf = (float)Ww;
    const int base = (int)__ldg(lsi + l) * xs, ys = Ww * xs;
    __half2 a[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = __float2half2_rn(0.f);

    auto point = [&](int p, __half2 (&w)[4], int (&o)[4]) {
      const float2 xy = __ldg(locq + l * K + p);
      const float aw = __ldg(attq + l * K + p);
      const float h_im = __fmaf_rn(xy.y, Hf, -0.5f), w_im = __fmaf_rn(xy.x, Wf, -0.5f);
      const bool inside = h_im > -1.f && w_im > -1.f && h_im < Hf && w_im < Wf;
      const float hfl = floorf(h_im), wfl = floorf(w_im);
      const int y0 = (int)hfl, x0 = (int)wfl;
      const float lh = h_im - hfl, lw = w_im - wfl, hh = 1.f - lh, hw = 1.f - lw;
      const bool y0ok = inside && y0 >= 0, y1ok = inside && y0 + 1 <= Hh - 1, x0ok = x0 >= 0, x1ok = x0 + 1 <= Ww - 1;
      const int yc0 = min(max(y0, 0), Hh - 1), yc1 = min(max(y0 + 1, 0), Hh - 1);
      const int xc0 = min(max(x0, 0), Ww - 1), xc1 = min(max(x0 + 1, 0), Ww - 1);
      o[0] = base + yc0 * ys + xc0 * xs;
      o[1] = base + yc0 * ys + xc1 * xs;
      o[2] = base + yc1 * ys + xc0 * xs;
      o[3] = base + yc1 * ys + xc1 * xs;
      w[0] = __float2half2_rn((y0ok && x0ok) ? hh * hw * aw : 0.f);
      w[1] = __float2half2_rn((y0ok && x1ok) ? hh * lw * aw : 0.f);
      w[2] = __float2half2_rn((y1ok && x0ok) ? lh * hw * aw : 0.f);
      w[3] = __float2half2_rn((y1ok && x1ok) ? lh * lw * aw : 0.f);
    };
    auto blend = [&](const __half2 (&w)[4], const uint4 (&r)[4]) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const __half2 *v = reinterpret_cast<const __half2 *>(&r[c]);
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = __hfma2(w[c], v[j], a[j]);
      }
    };
    if constexpr (KT > 0) {
      __half2 w[KT][4];
      int o[KT][4];
      uint4 r[KT][4];
#pragma unroll
      for (int p = 0; p < KT; ++p) point(p, w[p], o[p]);
#pragma unroll
      for (int p = 0; p < KT; ++p)
#pragma unroll
        for (int c = 0; c < 4; ++c) r[p][c] = __ldg(reinterpret_cast<const uint4 *>(vb + o[p][c]));
#pragma unroll
      for (int p = 0; p < KT; ++p) blend(w[p], r[p]);
    } else {
      for (int p = 0; p < K; ++p) {
        __half2 w[4];
        int o[4];
        uint4 r[4];
        point(p, w, o);
#pragma unroll
        for (int c = 0; c < 4; ++c) r[c] = __ldg(reinterpret_cast<const uint4 *>(vb + o[c]));
        blend(w, r);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(a[j]);
      acc[2 * j] += f.x;
      acc[2 * j + 1] += f.y;
    }
  }
  if constexpr (SPLIT > 1) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 4);
      acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 8);
    }
    if (!live || lvl0 != 0) return;
  }
  *reinterpret_cast<uint4 *>(out + (long)qh * D + sub * 8) = f32x8_to_bf16(acc);
}

static int launch_h16(const void *value, const int64_t *shapes, const int64_t *lsi, const float *loc, const float *attn,
                      void *out, int B, int S, int H, int L, int Lq, int K, int xs, cudaStream_t st, int ld_loc,
                      int ld_attn, long head_stride) {
  if (!ld_loc) ld_loc = H * L * K * 2;
  if (!ld_attn) ld_attn = H * L * K;
  const long n_qh = (long)B * Lq * H;
  const bool split = n_qh * 4 < (long)kNumSMs * 256 * 2;   // too few groups to fill the GPU: spread the levels over lanes
  const int grid = (int)((n_qh * (split ? 16 : 4) + 255) / 256);
  const bool hm = head_stride != 32, h8 = H == 8 && !hm;
  if (hm && split) return fail(MEMOTR_EINVAL, "msda_fwd_h16: head-major value maps need an encoder-sized launch");
  if (n_qh * 16 >= (1L << 31)) return fail(MEMOTR_EINVAL, "msda_fwd_h16: more than 2^27 (query, head) pairs");
#define H16_ARGS (const __half *)value, shapes, lsi, loc, attn, (__nv_bfloat16 *)out, S, H, L, Lq, K, xs, (int)n_qh, ld_loc, \
                 ld_attn, head_stride, B
#define H16_LAUNCH(KT_)                                                                       \
  if (split && h8) MEMOTR_LAUNCH((msda_fwd_h16<KT_, 4, false, true>), grid, 256, 0, st, H16_ARGS);      \
  else if (split) MEMOTR_LAUNCH((msda_fwd_h16<KT_, 4, false, false>), grid, 256, 0, st, H16_ARGS);      \
  else if (hm) MEMOTR_LAUNCH((msda_fwd_h16<KT_, 1, true, false>), grid, 256, 0, st, H16_ARGS);          \
  else if (h8) MEMOTR_LAUNCH((msda_fwd_h16<KT_, 1, false, true>), grid, 256, 0, st, H16_ARGS);          \
  else MEMOTR_LAUNCH((msda_fwd_h16<KT_, 1, false, false>), grid, 256, 0, st, H16_ARGS)
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
