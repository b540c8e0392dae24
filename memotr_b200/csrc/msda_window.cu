// msda_window.cu -- EXPERIMENT (opt-in, MEMOTR_MSDA_WINDOW=1): encoder-shaped MSDA forward with TMA-staged value-map
// windows in shared memory (the structure BASELINE.json's north star names).
//
// The global-memory gather (msda_fwd.cu, msda_fwd_h16) is bound on the SM by the L1 line rate: every bilinear corner is its
// own 64-byte segment of a 128-byte line.  Encoder queries are pixels; the sampling points of a tile of neighbouring queries
// fall, on every level, into a small window around the tile's footprint.  Here a CTA owns a tile of 8 x 8 level-0 queries
// and a PAIR of heads, loads one window per level (footprint + halo, 64 channels = 128 B per pixel) with four 3-D TMA
// copies (out-of-image pixels arrive as zeros, which is exactly the bilinear zero padding), and takes the taps from shared
// memory; a sampling point whose 2 x 2 footprint leaves the window falls back to the global path of msda_fwd_h16 with its
// validity masks.  Queries of the coarser levels (25 % of the rows) go through the global kernel.
#include "tc_common.cuh"

extern "C" int memotr_msda_forward_strided(const void *value, int value_pixel_stride, const int64_t *spatial_shapes,
                                           const int64_t *level_start_idx, const float *sampling_loc, int ld_loc,
                                           const float *attn_weight, int ld_attn, void *output, int B, int S, int H, int L,
                                           int Lq, int K, int head_major, void *stream);

namespace memotr {
namespace win {

constexpr int TQ = 8, HALO = 3, NL = 4;

struct Levels {
  int hw[2 * NL], lsi[NL];      // (H, W) and first pixel of every level
  int ww[NL], wh[NL], off[NL];  // window width / height (pixels) and byte offset of the level's window in shared memory
  int bytes;                    // sum of the window bytes
};

__global__ void __launch_bounds__(256)
msda_window_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                   const __grid_constant__ CUtensorMap tm2, const __grid_constant__ CUtensorMap tm3,
                   const __half *__restrict__ value, int xs, Levels lv, const float *__restrict__ loc, int ld_loc,
                   const float *__restrict__ attn, int ld_attn, const float *__restrict__ vr, __nv_bfloat16 *__restrict__ out,
                   int H, int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  __shared__ uint64_t bar;
  __shared__ int worg[2 * NL];                       // window origin (x, y) per level
  const int tid = threadIdx.x;
  const int W0 = lv.hw[1], H0 = lv.hw[0];
  const int tiles_x = (W0 + TQ - 1) / TQ;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x, hp = blockIdx.y;
  pdl_grid_sync();
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // window origin per level: pixel coordinate of the tile's first query on that level, minus the halo
    const float vx0 = __ldg(vr), vy0 = __ldg(vr + 1);
    const float rx = ((float)(tx * TQ) + 0.5f) / (vx0 * (float)W0), ry = ((float)(ty * TQ) + 0.5f) / (vy0 * (float)H0);
    tc::mbar_expect_tx(&bar, (uint32_t)lv.bytes);
    for (int l = 0; l < NL; ++l) {
      const float px = rx * __ldg(vr + 2 * l) * (float)lv.hw[2 * l + 1] - 0.5f, py = ry * __ldg(vr + 2 * l + 1) * (float)lv.hw[2 * l] - 0.5f;
      const int ox = (int)floorf(px) - HALO, oy = (int)floorf(py) - HALO;
      worg[2 * l] = ox, worg[2 * l + 1] = oy;
      const CUtensorMap *tm = l == 0 ? &tm0 : l == 1 ? &tm1 : l == 2 ? &tm2 : &tm3;
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
              tc::smem_u32(smem + lv.off[l])),
          "l"(tm), "r"(tc::smem_u32(&bar)), "r"(hp * 64), "r"(ox), "r"(oy)
          : "memory");
    }
  }
  __syncthreads();
  tc::mbar_wait(&bar, 0);

  const int LK = NL * K;
  for (int pass = 0; pass < 2; ++pass) {
    const int task = pass * 256 + tid;
    const int ql = task >> 3, hl = (task >> 2) & 1, sub = task & 3;
    const int x = tx * TQ + (ql & 7), y = ty * TQ + (ql >> 3);
    if (x >= W0 || y >= H0) continue;                 // (whole 4-lane groups leave together)
    const int q = y * W0 + x, h = hp * 2 + hl;
    const float2 *locq = reinterpret_cast<const float2 *>(loc + (long)q * ld_loc) + h * LK;
    const float *attq = attn + (long)q * ld_attn + h * LK;
    const __half *vb = value + h * 32 + sub * 8;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    for (int l = 0; l < NL; ++l) {
      const int Hh = lv.hw[2 * l], Ww = lv.hw[2 * l + 1];
      const float Hf = (float)Hh, Wf = (float)Ww;
      const int ox = worg[2 * l], oy = worg[2 * l + 1], ww = lv.ww[l], wh = lv.wh[l];
      const uint8_t *wbase = smem + lv.off[l] + hl * 64 + sub * 16;
      const long gbase = (long)lv.lsi[l] * xs;
      const int ys = Ww * xs;
      __half2 a[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) a[j] = __float2half2_rn(0.f);
      for (int p = 0; p < K; ++p) {
        const float2 xy = __ldg(locq + l * K + p);
        const float aw = __ldg(attq + l * K + p);
        const float h_im = __fmaf_rn(xy.y, Hf, -0.5f), w_im = __fmaf_rn(xy.x, Wf, -0.5f);
        const float hfl = floorf(h_im), wfl = floorf(w_im);
        const int y0 = (int)hfl, x0 = (int)wfl;
        const float lh = h_im - hfl, lw = w_im - wfl, hh = 1.f - lh, hw = 1.f - lw;
        uint4 r[4];
        __half2 w[4];
        const int dx = x0 - ox, dy = y0 - oy;
        if (dx >= 0 && dx + 1 < ww && dy >= 0 && dy + 1 < wh) {
          // the 2 x 2 footprint lies inside the staged window: taps from shared memory, zero padding came with the TMA fill
          const uint8_t *t0 = wbase + (dy * ww + dx) * 128;
          r[0] = *reinterpret_cast<const uint4 *>(t0);
          r[1] = *reinterpret_cast<const uint4 *>(t0 + 128);
          r[2] = *reinterpret_cast<const uint4 *>(t0 + ww * 128);
          r[3] = *reinterpret_cast<const uint4 *>(t0 + ww * 128 + 128);
          w[0] = __float2half2_rn(hh * hw * aw), w[1] = __float2half2_rn(hh * lw * aw);
          w[2] = __float2half2_rn(lh * hw * aw), w[3] = __float2half2_rn(lh * lw * aw);
        } else {
          const bool inside = h_im > -1.f && w_im > -1.f && h_im < Hf && w_im < Wf;
          const bool y0ok = inside && y0 >= 0, y1ok = inside && y0 + 1 <= Hh - 1, x0ok = x0 >= 0, x1ok = x0 + 1 <= Ww - 1;
          const int yc0 = min(max(y0, 0), Hh - 1), yc1 = min(max(y0 + 1, 0), Hh - 1);
          const int xc0 = min(max(x0, 0), Ww - 1), xc1 = min(max(x0 + 1, 0), Ww - 1);
          r[0] = __ldg(reinterpret_cast<const uint4 *>(vb + gbase + (long)yc0 * ys + xc0 * xs));
          r[1] = __ldg(reinterpret_cast<const uint4 *>(vb + gbase + (long)yc0 * ys + xc1 * xs));
          r[2] = __ldg(reinterpret_cast<const uint4 *>(vb + gbase + (long)yc1 * ys + xc0 * xs));
          r[3] = __ldg(reinterpret_cast<const uint4 *>(vb + gbase + (long)yc1 * ys + xc1 * xs));
          w[0] = __float2half2_rn((y0ok && x0ok) ? hh * hw * aw : 0.f), w[1] = __float2half2_rn((y0ok && x1ok) ? hh * lw * aw : 0.f);
          w[2] = __float2half2_rn((y1ok && x0ok) ? lh * hw * aw : 0.f), w[3] = __float2half2_rn((y1ok && x1ok) ? lh * lw * aw : 0.f);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const __half2 *v = reinterpret_cast<const __half2 *>(&r[c]);
#pragma unroll
          for (int j = 0; j < 4; ++j) a[j] = __hfma2(w[c], v[j], a[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(a[j]);
        acc[2 * j] += f.x, acc[2 * j + 1] += f.y;
      }
    }
    *reinterpret_cast<uint4 *>(out + (long)q * (H * 32) + h * 32 + sub * 8) = f32x8_to_bf16(acc);
  }
}

// 3-D fp16 map of one level of a pixel-major value map: dims (channels, W, H), box (64 channels, ww, wh), no swizzle
static bool make_level_map(CUtensorMap *map, const void *base, int C, int xs, int Hh, int Ww, int ww, int wh) {
  tc::EncodeTiledFn fn = tc::encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)Ww, (cuuint64_t)Hh};
  cuuint64_t strides[2] = {(cuuint64_t)xs * 2, (cuuint64_t)Ww * xs * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)ww, (cuuint32_t)wh};
  cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace win
}  // namespace memotr

using namespace memotr;

extern "C" int memotr_msda_forward_window(const void *value, int value_pixel_stride, const int64_t *spatial_shapes,
                                          const int64_t *level_start_idx, const int *shapes_hw, const int *level_start,
                                          const float *sampling_loc, int ld_loc, const float *attn_weight, int ld_attn,
                                          const float *valid_ratios, void *output, int S, int H, int L, int K, void *stream) {
  MEMOTR_REQUIRE(value && spatial_shapes && level_start_idx && shapes_hw && level_start && sampling_loc && attn_weight &&
                     valid_ratios && output,
                 "msda_forward_window: null pointer");
  MEMOTR_REQUIRE(L == win::NL && H % 2 == 0 && H * 32 <= value_pixel_stride && value_pixel_stride % 8 == 0 && K >= 1,
                 "msda_forward_window: needs 4 levels, an even head count, pixel stride >= H*32");
  MEMOTR_REQUIRE(aligned16(value) && aligned16(output) && ((reinterpret_cast<uintptr_t>(sampling_loc) & 7u) == 0) && ld_loc % 2 == 0,
                 "msda_forward_window: misaligned buffer");
  MEMOTR_REQUIRE(tc::encode_fn() != nullptr, "msda_forward_window: cuTensorMapEncodeTiled unavailable");
  win::Levels lv;
  CUtensorMap tm[win::NL];
  const int W0 = shapes_hw[1], H0 = shapes_hw[0];
  int off = 0;
  for (int l = 0; l < win::NL; ++l) {
    const int Hh = shapes_hw[2 * l], Ww = shapes_hw[2 * l + 1];
    lv.hw[2 * l] = Hh, lv.hw[2 * l + 1] = Ww, lv.lsi[l] = level_start[l];
    // footprint of 8 level-0 pixels on this level (+5 % for differing valid ratios) + halo on both sides + bilinear + rounding
    lv.ww[l] = (int)((win::TQ - 1) * 1.05f * Ww / W0) + 2 * win::HALO + 3;
    lv.wh[l] = (int)((win::TQ - 1) * 1.05f * Hh / H0) + 2 * win::HALO + 3;
    lv.off[l] = off;
    off += lv.ww[l] * lv.wh[l] * 128;
    MEMOTR_REQUIRE(lv.ww[l] <= 256 && lv.wh[l] <= 256, "msda_forward_window: window too large");
    if (!win::make_level_map(&tm[l], reinterpret_cast<const __half *>(value) + (long)level_start[l] * value_pixel_stride, H * 32,
                             value_pixel_stride, Hh, Ww, lv.ww[l], lv.wh[l]))
      return fail(MEMOTR_ECUDA, "msda_forward_window: cuTensorMapEncodeTiled failed (level %d)", l);
  }
  lv.bytes = off;
  MEMOTR_REQUIRE(off + 256 <= 200 * 1024, "msda_forward_window: windows do not fit in shared memory");
  static int attr_bytes = 0;
  if (attr_bytes < off + 256) {
    const cudaError_t e = cudaFuncSetAttribute(win::msda_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, off + 256);
    if (e != cudaSuccess) return fail(MEMOTR_ECUDA, "msda_forward_window: smem attribute: %s", cudaGetErrorString(e));
    attr_bytes = off + 256;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles = ceil_div(W0, win::TQ) * ceil_div(H0, win::TQ);
  MEMOTR_LAUNCH((win::msda_window_kernel), dim3(tiles, H / 2), 256, (size_t)off + 256, st, tm[0], tm[1], tm[2], tm[3],
                (const __half *)value, value_pixel_stride, lv, sampling_loc, ld_loc, attn_weight, ld_attn, valid_ratios,
                (__nv_bfloat16 *)output, H, K);
  const int rc = check_launch("msda_window");
  if (rc != MEMOTR_OK) return rc;
  const int n0 = H0 * W0;                                   // the coarser levels' queries: global-memory gather
  if (S > n0)
    return memotr_msda_forward_strided(value, value_pixel_stride, spatial_shapes, level_start_idx,
                                       sampling_loc + (long)n0 * ld_loc, ld_loc, attn_weight + (long)n0 * ld_attn, ld_attn,
                                       reinterpret_cast<__nv_bfloat16 *>(output) + (long)n0 * H * 32, 1, S, H, L, S - n0, K, 0,
                                       stream);
  return MEMOTR_OK;
}
