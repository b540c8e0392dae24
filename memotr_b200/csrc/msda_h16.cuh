// msda_h16.cuh -- the fp16-value-map arithmetic shared by the two gathers of the bf16 engine:
//   msda_fwd_h16      (msda_fwd.cu)     taps straight from global memory (decoder-shaped launches, any sampling pattern)
//   msda_window_kernel (msda_window.cu) taps from TMA-staged shared-memory windows (encoder-shaped launches)
// Both replace ms_deformable_im2col_gpu_kernel + ms_deform_attn_im2col_bilinear
// (/root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299, :33-84) for a value map stored in fp16 and a
// bf16 output row, and both evaluate EXACTLY the same expression, so their outputs are bit-identical
// (tests/test_msda_gpu.py):
//
//   per sampling point: (x, y) -> pixel coordinate as the reference does (h = fma(loc_y, H, -0.5), .cuh:285-286), the
//   open-interval test of .cuh:288, floor, the four bilinear weights times the attention weight in fp32, each rounded
//   to fp16 (0 for a corner outside the image, .cuh:56-78);
//   per PAIR of levels (2j, 2j+1) and x-SIDE s in {x0, x0+1}:  a_s = sum over the pair's points, level-major, of
//   w(y0,s) v(y0,s) + w(y1,s) v(y1,s)  as packed-half FMAs in that order (two channels per instruction; 16 FMAs per
//   accumulator at K = 4, the length round 1 validated against the oracle);
//   acc_s += float(a_s) after every level pair (fp32);   out = bf16(acc_0 + acc_1).
//
// The split by x-side is what lets the shared-memory gather read both x-corners of a footprint row with ONE conflict-free
// 128-byte wavefront (eight lanes: four on the x0 pixel, four on the x0+1 pixel).
#pragma once
#include "common.cuh"

namespace memotr {
namespace h16 {

struct Point {
  int x0, y0;       // floor of the pixel coordinate; may lie outside the image
  int xc[2], yc[2]; // the two columns / rows clamped into the image (addresses always valid)
  __half2 ws[2];    // per x-side s: (w(y0, s), w(y1, s)) = bilinear x attention weight, 0 where the corner is outside
};

__device__ __forceinline__ Point decode(float2 xy, float aw, int Hh, int Ww) {
  const float Hf = (float)Hh, Wf = (float)Ww;
  const float h_im = __fmaf_rn(xy.y, Hf, -0.5f), w_im = __fmaf_rn(xy.x, Wf, -0.5f);
  const bool inside = h_im > -1.f && w_im > -1.f && h_im < Hf && w_im < Wf;
  const float hfl = floorf(h_im), wfl = floorf(w_im);
  Point p;
  // (float -> int conversion saturates on the GPU, NaN -> 0: a far-away or non-finite location is not `inside`, contributes
  //  nothing, and its clamped addresses below stay valid)
  p.y0 = __float2int_rd(h_im), p.x0 = __float2int_rd(w_im);
  const float lh = h_im - hfl, lw = w_im - wfl, hh = 1.f - lh, hw = 1.f - lw;
  const bool y0ok = inside && p.y0 >= 0, y1ok = inside && p.y0 + 1 <= Hh - 1, x0ok = p.x0 >= 0, x1ok = p.x0 + 1 <= Ww - 1;
  p.yc[0] = min(max(p.y0, 0), Hh - 1), p.yc[1] = min(max(p.y0 + 1, 0), Hh - 1);
  p.xc[0] = min(max(p.x0, 0), Ww - 1), p.xc[1] = min(max(p.x0 + 1, 0), Ww - 1);
  p.ws[0] = __floats2half2_rn((y0ok && x0ok) ? hh * hw * aw : 0.f, (y1ok && x0ok) ? lh * hw * aw : 0.f);
  p.ws[1] = __floats2half2_rn((y0ok && x1ok) ? hh * lw * aw : 0.f, (y1ok && x1ok) ? lh * lw * aw : 0.f);
  return p;
}

// a += w(y0) * row0 + w(y1) * row1 for the eight channels (16 bytes) a lane owns
__device__ __forceinline__ void blend(__half2 (&a)[4], __half2 w, const uint4 &r0, const uint4 &r1) {
  const __half2 *v0 = reinterpret_cast<const __half2 *>(&r0), *v1 = reinterpret_cast<const __half2 *>(&r1);
  const __half2 wy0 = __low2half2(w), wy1 = __high2half2(w);
#pragma unroll
  for (int j = 0; j < 4; ++j) a[j] = __hfma2(wy0, v0[j], a[j]);
#pragma unroll
  for (int j = 0; j < 4; ++j) a[j] = __hfma2(wy1, v1[j], a[j]);
}

__device__ __forceinline__ void widen_add(float (&acc)[8], const __half2 (&a)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(a[j]);
    acc[2 * j] += f.x, acc[2 * j + 1] += f.y;
  }
}

// level table read either from the op's int64 device tensors or from host-filled ints in the kernel parameters
struct ShapesI64 {
  const int64_t *shapes, *lsi;
  __device__ __forceinline__ int H(int l) const { return (int)__ldg(shapes + 2 * l); }
  __device__ __forceinline__ int W(int l) const { return (int)__ldg(shapes + 2 * l + 1); }
  __device__ __forceinline__ int start(int l) const { return (int)__ldg(lsi + l); }
};
struct ShapesI32 {
  const int *hw, *lsi;
  __device__ __forceinline__ int H(int l) const { return hw[2 * l]; }
  __device__ __forceinline__ int W(int l) const { return hw[2 * l + 1]; }
  __device__ __forceinline__ int start(int l) const { return lsi[l]; }
};

// One (query, head) from global memory with FOUR lanes (lane `sub` owns channels 8*sub .. 8*sub+7 of the head and both
// x-sides).  `vb` = value + head * 32 + sub * 8 (+ batch offset); levels l0, l0 + lstep, ... (lstep > 1: the levels are
// spread over several 4-lane subgroups and the caller adds the partial sums).  KT > 0: points per level at compile time.
template <int KT, typename SH>
__device__ __forceinline__ void gather_global(const __half *__restrict__ vb, const SH &sh, const float2 *__restrict__ locq,
                                              const float *__restrict__ attq, int L, int Kr, int xs, int l0, int lstep,
                                              float (&acc0)[8], float (&acc1)[8]) {
  const int K = KT ? KT : Kr;
  __half2 a0[4], a1[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) a0[j] = a1[j] = __float2half2_rn(0.f);
  auto level = [&](int l) {
    const int Hh = sh.H(l), Ww = sh.W(l);
    const int base = sh.start(l) * xs, ys = Ww * xs;
    if constexpr (KT > 0) {
      Point p[KT];
      uint4 r[KT][4];
#pragma unroll
      for (int i = 0; i < KT; ++i) p[i] = decode(__ldg(locq + l * K + i), __ldg(attq + l * K + i), Hh, Ww);
#pragma unroll
      for (int i = 0; i < KT; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c)
          r[i][c] = __ldg(reinterpret_cast<const uint4 *>(vb + base + p[i].yc[c >> 1] * ys + p[i].xc[c & 1] * xs));
#pragma unroll
      for (int i = 0; i < KT; ++i) {
        blend(a0, p[i].ws[0], r[i][0], r[i][2]);
        blend(a1, p[i].ws[1], r[i][1], r[i][3]);
      }
    } else {
      for (int i = 0; i < K; ++i) {
        const Point p = decode(__ldg(locq + l * K + i), __ldg(attq + l * K + i), Hh, Ww);
        uint4 r[4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
          r[c] = __ldg(reinterpret_cast<const uint4 *>(vb + base + p.yc[c >> 1] * ys + p.xc[c & 1] * xs));
        blend(a0, p.ws[0], r[0], r[2]);
        blend(a1, p.ws[1], r[1], r[3]);
      }
    }
  };
  auto widen = [&]() {                        // end of a level pair: into the fp32 sums
    widen_add(acc0, a0);
    widen_add(acc1, a1);
#pragma unroll
    for (int j = 0; j < 4; ++j) a0[j] = a1[j] = __float2half2_rn(0.f);
  };
  if (lstep == 1) {
    for (int l = l0; l < L; l += 2) {
      level(l);
      if (l + 1 < L) level(l + 1);
      widen();
    }
  } else {                                     // (levels spread over lane subgroups: each subgroup's levels are 'pairs' of one)
    for (int l = l0; l < L; l += lstep) {
      level(l);
      widen();
    }
  }
}

}  // namespace h16
}  // namespace memotr
