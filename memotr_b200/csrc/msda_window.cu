// msda_window.cu -- encoder-shaped multi-scale deformable attention forward with TMA-staged value-map windows in shared
// memory (the structure BASELINE.json's north star names).  Replaces ms_deformable_im2col_gpu_kernel
// (/root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299, bilinear taps :33-84) for the encoder's
// self-attention (deformable_encoder.py:124: the queries ARE the pixels of the pyramid), fp16 value map, bf16 output.
// Arithmetic: msda_h16.cuh -- bit-identical to the global-memory gather msda_fwd_h16.
//
// Why: a gather over a pixel-major map from global memory pays one L1 tag cycle per 64-byte corner (the lane groups of a
// warp touch different lines; measured 2 cycles per line inside one LDG), 24 cycles per 4 sampling points, 61 us per encoder
// launch = 14 % of the HBM roofline with the DRAM pipe 7 % busy (profiles/r01_msda_fwd_h16_final_ncu.md).  Shared memory
// has no line granularity, only banks.  So:
//
//   unit  = (tile of neighbouring queries of one level, one head); one CTA of 256 threads per unit, two CTAs per SM
//   stage = per level one 3-D TMA copy of the window (32 channels x ww x wh pixels, 64 B per pixel, pixels outside the
//           image arrive as zeros = the bilinear zero padding) around where the tile's reference points land on that
//           level, shifted by the head's mean sampling offset (a host-side hint; correctness never depends on it)
//   decode= every sampling point of the tile is decoded ONCE (one point per thread, coalesced reads of the locations /
//           weights) into an 8-byte record per x-side: {shared-memory offset of (y0, x_side) | fallback pixel, w(y0), w(y1)}
//   taps  = eight lanes per (query, head): lanes 0-3 own the x0 pixel, lanes 4-7 the x0+1 pixel, 16 bytes = 8 channels
//           each, so the two x-corners of a footprint row are ONE contiguous 128-byte shared-memory read -- a single
//           conflict-free wavefront -- and a point costs two LDS.128 per lane; packed-half FMAs; one xor-shuffle per
//           channel at the end adds the two sides.  A point whose footprint leaves the window reads global memory instead
//           (same arithmetic), so any sampling pattern is computed correctly.
//   the queries of the coarse levels (6 % of the rows at 1333x800), whose windows would be larger than their taps, run
//   in the same launch on CTAs that take the global-memory path.
//
// Floors per encoder launch (22 323 queries x 8 heads x 16 points x 4 corners x 64 B = 731 MB of taps): shared-memory
// bandwidth 128 B/clk/SM x 148 SMs = 19.6 us at 1.965 GHz; algorithmic HBM bytes 57.1 MB = 8.7 us.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "msda_h16.cuh"
#include "tc_common.cuh"

namespace memotr {
namespace win {

constexpr int MAXL = 5, MAXC = 4, MAXH = 16;
constexpr int TAP_WARPS = 16, DEC_WARPS = 8, THREADS = (TAP_WARPS + DEC_WARPS) * 32;
constexpr int SMEM_BUDGET = 110 * 1024;      // per stage (windows + records of one unit); a CTA holds two stages + ~5 KB of static
                                             // shared memory (unit descriptors, barriers) in the 227 KB of an SM

struct ClassGeom {
  int lq;                       // the level this class's queries live on
  int tw, th, tiles_x, tiles_y; // tile of tw x th queries (powers of two, tw * th a multiple of 32); tiles over the level
  int tw_shift;
  int unit0, n_units;           // first unit of the class; units = tiles * heads (head fastest)
  int q0, Wq, Hq;               // first query row of the level, its extent
  int ww[MAXL], wh[MAXL];       // window extent in pixels per value level (0: no window, the level's taps read global memory)
  int off[MAXL];                // byte offset of the window in dynamic shared memory
  int win_bytes, rec_off, rec_stride;
  int flag_off;                 // one byte per (query of the tile, level): some point of that level reads global memory
  float tiles_x_rcp;
};

struct Params {
  ClassGeom cls[MAXC];
  int n_cls, n_glob_blocks, glob_q0;   // CTAs [0, n_glob_blocks) take the queries [glob_q0, S) from global memory
  int hw[2 * MAXL], lsi[MAXL];
  int L, K, H, S, xs, ld_loc, ld_attn;
  int stage_bytes;                     // bytes of one stage (windows + records); the CTA double-buffers stages
  int h_shift;                         // log2(H) when H is a power of two, else -1
  int debug;                           // MEMOTR_WINDOW_DEBUG (timing experiments, wrong results): 1 = the decode warps build records
                                       // for their first two units only (tap-bound time), 2 = the tap warps skip the taps (decode-bound)
  float radius;
  float shift[MAXH * MAXL * 2];        // per (head, level): expected sampling offset (x, y) in pixels of that level
};

struct Maps {
  CUtensorMap m[MAXC * MAXL];
};

// record of one sampling point and x-side
//   window tap:   off = byte offset of pixel (y0, x_side) in dynamic shared memory (the y1 row is ww_l * 64 B further)
//   global tap:   off = 1 << 31 | (y1 row == y0 row) << 30 | pixel index of (yc0, xc_side) in the whole value map
struct __align__(8) Rec {
  uint32_t off;
  __half2 w;                     // (w(y0, side), w(y1, side))
};

// window origin of level l for the tile whose first query has the normalised reference point (rx, ry): every thread that
// needs it evaluates exactly this sequence (explicitly rounded operations: no context-dependent contraction), so the copy
// the TMA issuer uses and the copies the decoding threads use are the same integers
__device__ __forceinline__ int2 window_origin(const Params &P, const ClassGeom &G, const float *__restrict__ vr, float rx, float ry,
                                              int h, int l) {
  const int Hh = P.hw[2 * l], Ww = P.hw[2 * l + 1];
  const float px = __fadd_rn(__fmaf_rn(__fmul_rn(rx, __ldg(vr + 2 * l)), (float)Ww, -0.5f), P.shift[(h * MAXL + l) * 2]);
  const float py = __fadd_rn(__fmaf_rn(__fmul_rn(ry, __ldg(vr + 2 * l + 1)), (float)Hh, -0.5f), P.shift[(h * MAXL + l) * 2 + 1]);
  int ox = (int)floorf(__fsub_rn(px, P.radius)), oy = (int)floorf(__fsub_rn(py, P.radius));
  // keep the window on the image (one zero column / row of border is all a bilinear footprint can use): border tiles and
  // levels smaller than their window then stage what can actually be sampled
  ox = max(-1, min(ox, Ww + 1 - G.ww[l])), oy = max(-1, min(oy, Hh + 1 - G.wh[l]));
  return make_int2(ox, oy);
}

struct Unit {
  int c, h, tx, ty;
};
__device__ __forceinline__ Unit unit_of(const Params &P, int u) {
  Unit t;
  t.c = 0;
  for (int c = 1; c < P.n_cls; ++c)
    if (u >= P.cls[c].unit0) t.c = c;
  const ClassGeom &G = P.cls[t.c];
  const int v = u - G.unit0;
  int tile;
  if (P.h_shift >= 0) tile = v >> P.h_shift, t.h = v & (P.H - 1);
  else tile = v / P.H, t.h = v - tile * P.H;
  int ty = (int)(((float)tile + 0.5f) * G.tiles_x_rcp);       // tile / tiles_x (tile < 2^20: exact after the fix-up)
  if (ty * G.tiles_x > tile) --ty;
  t.ty = ty, t.tx = tile - ty * G.tiles_x;
  return t;
}

// one sampling point -> its two records (x-side 0 / 1).  Same numbers as h16::decode; a tap that lands in its window needs
// neither clamped addresses nor zeroed weights (pixels outside the image arrive as zeros from the TMA fill).
__device__ __forceinline__ void make_records(const Params &P, const ClassGeom &G, int l, int2 org, uint32_t wbase, float2 xy,
                                             float aw, Rec &r0, Rec &r1, int &n_win, int &n_glob) {
  const int Hh = P.hw[2 * l], Ww = P.hw[2 * l + 1];
  const float Hf = (float)Hh, Wf = (float)Ww;
  const float h_im = __fmaf_rn(xy.y, Hf, -0.5f), w_im = __fmaf_rn(xy.x, Wf, -0.5f);
  r0.off = r1.off = wbase;
  r0.w = r1.w = __float2half2_rn(0.f);
  if (!(h_im > -1.f && w_im > -1.f && h_im < Hf && w_im < Wf)) return;        // contributes nothing (.cuh:288; NaN too)
  const int y0 = __float2int_rd(h_im), x0 = __float2int_rd(w_im);
  const int ww = G.ww[l], dx = x0 - org.x, dy = y0 - org.y;
  if (ww && (unsigned)dx < (unsigned)(ww - 1) && (unsigned)dy < (unsigned)(G.wh[l] - 1)) {
    const float lh = h_im - floorf(h_im), lw = w_im - floorf(w_im), hh = 1.f - lh, hw = 1.f - lw;
    r0.w = __floats2half2_rn(hh * hw * aw, lh * hw * aw);
    r1.w = __floats2half2_rn(hh * lw * aw, lh * lw * aw);
    r0.off = wbase + (uint32_t)(G.off[l] + (dy * ww + dx) * 64);
    r1.off = r0.off + 64u;
    ++n_win;
  } else {
    const h16::Point p = h16::decode(xy, aw, Hh, Ww);
    const uint32_t flags = 0x80000000u | (p.yc[1] == p.yc[0] ? 0x40000000u : 0u);
    const uint32_t row = (uint32_t)(P.lsi[l] + p.yc[0] * Ww);
    r0.w = p.ws[0], r1.w = p.ws[1];
    r0.off = flags | (row + (uint32_t)p.xc[0]);
    r1.off = flags | (row + (uint32_t)p.xc[1]);
    ++n_glob;
  }
}

// wait with a suspend-time hint: a warp that has to wait (the other role is behind) sleeps in the barrier unit instead of
// spinning through the issue slots the working warps of its scheduler need
__device__ __forceinline__ void mbar_wait_sleepy(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "SLEEPY_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra SLEEPY_DONE;\n\t"
      "bra SLEEPY_WAIT;\n\t"
      "SLEEPY_DONE:\n\t"
      "}" ::"r"(tc::smem_u32(bar)),
      "r"(parity), "r"(20000u)
      : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

// Persistent, warp-specialised: one CTA per SM walks units u = blockIdx, + stride, ...; the DECODE warps (8) stage unit n + 1
// (TMA windows + records, two buffers of each) while the TAP warps (16) consume unit n, so the shared-memory pipe -- the
// resource that bounds this kernel -- never waits for a decode phase.
//   win_full[b]  TMA bytes of window buffer b have landed           (decode thread 0 arms it, the copy engine completes it)
//   rec_full[b]  the records of buffer b are written                (one arrival per decode warp)
//   buf_free[b]  the tap warps are done with buffers b               (one arrival per tap warp)
template <int KT>
__global__ void __launch_bounds__(THREADS, 1)
msda_window_kernel(const __grid_constant__ Maps maps, const __grid_constant__ Params P, const __half *__restrict__ value,
                   const float *__restrict__ loc, const float *__restrict__ attn, const float *__restrict__ vr,
                   unsigned long long *__restrict__ stats, __nv_bfloat16 *__restrict__ out) {
  extern __shared__ __align__(128) uint8_t smem[];      // (declared alignment: the windows are TMA destinations)
  __shared__ uint64_t win_full[2], rec_full[2], buf_free[2];
  const int tid = threadIdx.x;
  const int K = KT ? KT : P.K, L = P.L, H = P.H, LK = L * K;

  // ------------------------------------------------------------------------------------------- global-memory role
  if ((int)blockIdx.x < P.n_glob_blocks) {
    pdl_grid_sync();
    const int n_qh = (P.S - P.glob_q0) * H;
    const int qh = (blockIdx.x * THREADS + tid) >> 2, sub = tid & 3;
    if (qh >= n_qh) return;
    const int q = P.glob_q0 + qh / H, m = qh % H;
    float acc0[8], acc1[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc0[c] = acc1[c] = 0.f;
    h16::gather_global<0>(value + m * 32 + sub * 8, h16::ShapesI32{P.hw, P.lsi},
                           reinterpret_cast<const float2 *>(loc + (long)q * P.ld_loc) + m * LK,
                           attn + (long)q * P.ld_attn + m * LK, L, P.K, P.xs, 0, 1, acc0, acc1);
#pragma unroll
    for (int c = 0; c < 8; ++c) acc0[c] += acc1[c];
    *reinterpret_cast<uint4 *>(out + ((long)q * H + m) * 32 + sub * 8) = f32x8_to_bf16(acc0);
    return;
  }

  // ------------------------------------------------------------------------------------------- window role (persistent)
  const int n_units = P.cls[P.n_cls - 1].unit0 + P.cls[P.n_cls - 1].n_units;
  const int stride = gridDim.x - P.n_glob_blocks;
  const int first = blockIdx.x - P.n_glob_blocks;
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) tc::mbar_init(win_full + b, 1), tc::mbar_init(rec_full + b, DEC_WARPS), tc::mbar_init(buf_free + b, TAP_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_grid_sync();
  const int xs = P.xs;

  if (tid >= TAP_WARPS * 32) {
    // =========================================================================================== decode warps
    const int dt = tid - TAP_WARPS * 32, lane = tid & 31;
    int n_win = 0, n_glob = 0;                  // (profiling only: points staged / points left to global memory)
    // KT == 4: a thread decodes the four points of (query, level) pairs -- pair p = dt + 256 i, query = p / L, level = p % L --
    // from registers that were loaded one unit ahead (the global-memory latency hides behind the previous unit's decode)
    constexpr int NPRE = KT == 4 ? 2 : 1;
    float4 pre_xy[NPRE][2], pre_w[NPRE];
    auto prefetch = [&](int u) {
      if constexpr (KT == 4) {
        const Unit t = unit_of(P, u);
        const ClassGeom &G = P.cls[t.c];
#pragma unroll
        for (int i = 0; i < NPRE; ++i) {
          const int p = dt + i * DEC_WARPS * 32, ql = p / L, l = p - ql * L;
          const int x = t.tx * G.tw + (ql & (G.tw - 1)), y = t.ty * G.th + (ql >> G.tw_shift);
          pre_w[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          pre_xy[i][0] = pre_xy[i][1] = pre_w[i];
          if (ql < G.tw * G.th && x < G.Wq && y < G.Hq) {
            const long q = G.q0 + y * G.Wq + x;
            const float4 *pl = reinterpret_cast<const float4 *>(loc + q * P.ld_loc + (t.h * LK + l * 4) * 2);
            pre_xy[i][0] = __ldg(pl), pre_xy[i][1] = __ldg(pl + 1);
            pre_w[i] = __ldg(reinterpret_cast<const float4 *>(attn + q * P.ld_attn + t.h * LK + l * 4));
          }
        }
      }
    };
    if (first < n_units) prefetch(first);
    int n = 0;
    for (int unit = first; unit < n_units; unit += stride, ++n) {
      const int b = n & 1;
      const Unit t = unit_of(P, unit);
      const int c = t.c, h = t.h, tx = t.tx, ty = t.ty;
      const ClassGeom &G = P.cls[c];
      const int TQ = G.tw * G.th;
      const uint32_t wbase = (uint32_t)(b * P.stage_bytes);
      uint8_t *recs = smem + wbase + G.rec_off;
      const float rx = __fdiv_rn(__fadd_rn((float)(tx * G.tw), 0.5f), __fmul_rn(__ldg(vr + 2 * G.lq), (float)G.Wq));
      const float ry = __fdiv_rn(__fadd_rn((float)(ty * G.th), 0.5f), __fmul_rn(__ldg(vr + 2 * G.lq + 1), (float)G.Hq));
      if (n >= 2) mbar_wait_sleepy(buf_free + b, ((n >> 1) - 1) & 1);       // the tap warps have finished unit n - 2
      if (dt == 0) {
        // where the tile's first reference point lands on every level (deformable_encoder.py:29-40), plus the head's
        // expected offset, minus the radius: the window origin.  Any origin is correct; a good one makes every tap a window tap.
        tc::mbar_expect_tx(win_full + b, (uint32_t)G.win_bytes);
        for (int l = 0; l < L; ++l) {
          if (!G.ww[l]) continue;
          const int2 org = window_origin(P, G, vr, rx, ry, h, l);
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                  tc::smem_u32(smem + wbase + G.off[l])),
              "l"(&maps.m[c * MAXL + l]), "r"(tc::smem_u32(win_full + b)), "r"(h * 32), "r"(org.x), "r"(org.y)
              : "memory");
        }
      }
      // ---- every sampling point of the tile once, into the records ------------------------------------------------
      if constexpr (KT == 4) {
#pragma unroll
        for (int i = 0; i < NPRE; ++i) {
          const int p = dt + i * DEC_WARPS * 32, ql = p / L, l = p - ql * L;
          if (ql >= TQ) continue;
          const int x = tx * G.tw + (ql & (G.tw - 1)), y = ty * G.th + (ql >> G.tw_shift);
          Rec r[4][2];
#pragma unroll
          for (int j = 0; j < 4; ++j) r[j][0].off = r[j][1].off = wbase, r[j][0].w = r[j][1].w = __float2half2_rn(0.f);
          if (x < G.Wq && y < G.Hq) {    // (a query outside the level: zero-weight taps on the first bytes of the buffer)
            const int2 org = window_origin(P, G, vr, rx, ry, h, l);
            const float4 a = pre_xy[i][0], bb = pre_xy[i][1], w = pre_w[i];
            make_records(P, G, l, org, wbase, make_float2(a.x, a.y), w.x, r[0][0], r[0][1], n_win, n_glob);
            make_records(P, G, l, org, wbase, make_float2(a.z, a.w), w.y, r[1][0], r[1][1], n_win, n_glob);
            make_records(P, G, l, org, wbase, make_float2(bb.x, bb.y), w.z, r[2][0], r[2][1], n_win, n_glob);
            make_records(P, G, l, org, wbase, make_float2(bb.z, bb.w), w.w, r[3][0], r[3][1], n_win, n_glob);
          }
          // records of a query: per pair of points a 32-byte block [side 0: point 2k, 2k+1 | side 1: point 2k, 2k+1]
          uint4 *dst = reinterpret_cast<uint4 *>(recs + ql * G.rec_stride + l * 64);
          dst[0] = make_uint4(r[0][0].off, *reinterpret_cast<uint32_t *>(&r[0][0].w), r[1][0].off, *reinterpret_cast<uint32_t *>(&r[1][0].w));
          dst[1] = make_uint4(r[0][1].off, *reinterpret_cast<uint32_t *>(&r[0][1].w), r[1][1].off, *reinterpret_cast<uint32_t *>(&r[1][1].w));
          dst[2] = make_uint4(r[2][0].off, *reinterpret_cast<uint32_t *>(&r[2][0].w), r[3][0].off, *reinterpret_cast<uint32_t *>(&r[3][0].w));
          dst[3] = make_uint4(r[2][1].off, *reinterpret_cast<uint32_t *>(&r[2][1].w), r[3][1].off, *reinterpret_cast<uint32_t *>(&r[3][1].w));
        }
        if (unit + stride < n_units) prefetch(unit + stride);
      } else {
        int ql = dt / LK, j = dt - ql * LK;
        const int dq = (DEC_WARPS * 32) / LK, dj = DEC_WARPS * 32 - dq * LK;
        while (ql < TQ) {
          const int l = j / K;
          const int x = tx * G.tw + (ql & (G.tw - 1)), y = ty * G.th + (ql >> G.tw_shift);
          Rec r0, r1;
          r0.off = r1.off = wbase;
          r0.w = r1.w = __float2half2_rn(0.f);
          if (x < G.Wq && y < G.Hq) {
            const long q = G.q0 + y * G.Wq + x;
            const float2 xy = __ldg(reinterpret_cast<const float2 *>(loc + q * P.ld_loc) + h * LK + j);
            const float aw = __ldg(attn + q * P.ld_attn + h * LK + j);
            make_records(P, G, l, window_origin(P, G, vr, rx, ry, h, l), wbase, xy, aw, r0, r1, n_win, n_glob);
          }
          uint8_t *dst = recs + ql * G.rec_stride + (j >> 1) * 32 + (j & 1) * 8;
          *reinterpret_cast<Rec *>(dst) = r0;
          *reinterpret_cast<Rec *>(dst + 16) = r1;
          ql += dq, j += dj;
          if (j >= LK) j -= LK, ++ql;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(rec_full + b);     // (release: this warp's record stores are visible to the waiting tap warps)
    }
    if (stats) atomicAdd(stats, (unsigned long long)n_win), atomicAdd(stats + 1, (unsigned long long)n_glob);
    return;
  }

  // ============================================================================================= tap warps
  const int grp = tid >> 3, side = (tid >> 2) & 1, sub = tid & 3, lane = tid & 31;
  const uint32_t smem_base = tc::smem_u32(smem) + sub * 16;
  int n = 0;
  for (int unit = first; unit < n_units; unit += stride, ++n) {
    const int b = n & 1;
    const Unit t = unit_of(P, unit);
    const int h = t.h, tx = t.tx, ty = t.ty;
    const ClassGeom &G = P.cls[t.c];
    const int TQ = G.tw * G.th;
    const uint8_t *recs = smem + b * P.stage_bytes + G.rec_off;
    mbar_wait_sleepy(rec_full + b, (n >> 1) & 1);
    mbar_wait_sleepy(win_full + b, (n >> 1) & 1);
    const __half *vb = value + h * 32 + sub * 8;
    for (int ql = grp; ql < TQ; ql += TAP_WARPS * 4) {
      const uint8_t *rq = recs + ql * G.rec_stride + side * 16;
      float acc[8];
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) acc[cc] = 0.f;
      __half2 a[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) a[jj] = __float2half2_rn(0.f);
      for (int l = 0; l < L; ++l) {
        const uint32_t rowb = (uint32_t)G.ww[l] * 64u;        // bytes between the y0 and y1 rows of this level's window
        auto window_tap = [&](uint32_t off, uint4 &r0, uint4 &r1) {
          const uint32_t a0 = smem_base + off, a1 = a0 + rowb;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r0.x), "=r"(r0.y), "=r"(r0.z), "=r"(r0.w) : "r"(a0));
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r1.x), "=r"(r1.y), "=r"(r1.z), "=r"(r1.w) : "r"(a1));
        };
        auto any_tap = [&](uint32_t off, uint4 &r0, uint4 &r1) {
          if (off & 0x80000000u) {
            const __half *p0 = vb + (long)(off & 0x3FFFFFFFu) * xs;
            const __half *p1 = (off & 0x40000000u) ? p0 : p0 + (long)P.hw[2 * l + 1] * xs;
            r0 = __ldg(reinterpret_cast<const uint4 *>(p0));
            r1 = __ldg(reinterpret_cast<const uint4 *>(p1));
          } else {
            window_tap(off, r0, r1);
          }
        };
        if constexpr (KT == 4) {
          const uint4 ra = *reinterpret_cast<const uint4 *>(rq + (l * 2) * 32);
          const uint4 rb = *reinterpret_cast<const uint4 *>(rq + (l * 2 + 1) * 32);
          uint4 v[4][2];
          // warp-uniform choice: the plain shared-memory loads unless some lane of the warp has a global-memory tap
          if (!__any_sync(0xffffffffu, (ra.x | ra.z | rb.x | rb.z) & 0x80000000u)) {
            window_tap(ra.x, v[0][0], v[0][1]);
            window_tap(ra.z, v[1][0], v[1][1]);
            window_tap(rb.x, v[2][0], v[2][1]);
            window_tap(rb.z, v[3][0], v[3][1]);
          } else {
            any_tap(ra.x, v[0][0], v[0][1]);
            any_tap(ra.z, v[1][0], v[1][1]);
            any_tap(rb.x, v[2][0], v[2][1]);
            any_tap(rb.z, v[3][0], v[3][1]);
          }
          h16::blend(a, *reinterpret_cast<const __half2 *>(&ra.y), v[0][0], v[0][1]);
          h16::blend(a, *reinterpret_cast<const __half2 *>(&ra.w), v[1][0], v[1][1]);
          h16::blend(a, *reinterpret_cast<const __half2 *>(&rb.y), v[2][0], v[2][1]);
          h16::blend(a, *reinterpret_cast<const __half2 *>(&rb.w), v[3][0], v[3][1]);
        } else {
          for (int pp = 0; pp < K / 2; ++pp) {
            const uint4 ra = *reinterpret_cast<const uint4 *>(rq + (l * (K / 2) + pp) * 32);
            uint4 v[2][2];
            any_tap(ra.x, v[0][0], v[0][1]);
            any_tap(ra.z, v[1][0], v[1][1]);
            h16::blend(a, *reinterpret_cast<const __half2 *>(&ra.y), v[0][0], v[0][1]);
            h16::blend(a, *reinterpret_cast<const __half2 *>(&ra.w), v[1][0], v[1][1]);
          }
        }
        if ((l & 1) || l + 1 == L) {             // end of a level pair: widen into the fp32 sums (msda_h16.cuh)
          h16::widen_add(acc, a);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) a[jj] = __float2half2_rn(0.f);
        }
      }
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) acc[cc] += __shfl_xor_sync(0xffffffffu, acc[cc], 4);
      const int x = tx * G.tw + (ql & (G.tw - 1)), y = ty * G.th + (ql >> G.tw_shift);
      if (side == 0 && x < G.Wq && y < G.Hq)
        *reinterpret_cast<uint4 *>(out + ((long)(G.q0 + y * G.Wq + x) * H + h) * 32 + sub * 8) = f32x8_to_bf16(acc);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(buf_free + b);
  }
}


// =====================================================================================================================
// K = 4 points x L = 4 levels (the configuration of every released MeMOTR / Deformable-DETR model): the same pipeline with
// everything the general kernel decides at run time decided here at compile time.  What the profile of the general kernel
// showed (profiles/r02_msda_window_v4_ncu.md): 37 M warp instructions per encoder launch of which 7 M spin in mbarrier
// try-wait loops, the tap warps waiting for records 16 % of their samples -- the DECODE side was the critical path (684
// instructions per warp and unit, low ILP, two warps per scheduler) -- and 81 instructions per (4 queries, level) in the tap
// loop against 42 that do work.  Here:
//   * a decode thread owns ONE level (pair index = thread + 256 i  ->  level = thread % 4): window origin, window bounds as
//     floats, the level's extents and the record base are per-unit registers; a point inside its window costs ~30
//     instructions (the window test in floats implies the reference's open-interval test because windows are clamped to
//     the image plus one zero border; everything else goes through slow_record);
//   * hand-offs are hardware named barriers in producer / consumer form (bar.arrive by the producers, bar.sync by the
//     consumers): a waiting warp is parked, not spinning;
//   * decode publishes one flag byte per (query, level); a tap warp reads the four bytes of a query as one word and, when no
//     lane of the warp has a global-memory tap (the common case), runs the four levels fully unrolled without a test.
// Results are bit-identical to the general kernel and to msda_fwd_h16 (tests/test_msda_window_gpu.py).
// =====================================================================================================================
__device__ __noinline__ uint4 slow_record(const int *hw, const int *lsi, int l, uint32_t wbase, float2 xy, float aw) {
  const int Hh = hw[2 * l], Ww = hw[2 * l + 1];
  const float Hf = (float)Hh, Wf = (float)Ww;
  const float h_im = __fmaf_rn(xy.y, Hf, -0.5f), w_im = __fmaf_rn(xy.x, Wf, -0.5f);
  if (!(h_im > -1.f && w_im > -1.f && h_im < Hf && w_im < Wf)) return make_uint4(wbase, 0u, wbase, 0u);   // (.cuh:288; NaN too)
  const h16::Point p = h16::decode(xy, aw, Hh, Ww);
  const uint32_t flags = 0x80000000u | (p.yc[1] == p.yc[0] ? 0x40000000u : 0u);
  const uint32_t row = (uint32_t)(lsi[l] + p.yc[0] * Ww);
  return make_uint4(flags | (row + (uint32_t)p.xc[0]), *reinterpret_cast<const uint32_t *>(&p.ws[0]), flags | (row + (uint32_t)p.xc[1]),
                    *reinterpret_cast<const uint32_t *>(&p.ws[1]));
}

// Per-unit descriptor, computed ONCE per CTA for all of its units before the pipeline starts (one thread per unit; the tap
// warps have nothing to do during the pipeline fill anyway).  The decode-only experiment (MEMOTR_WINDOW_DEBUG=2) showed the
// decode warps alone need 42 of the kernel's 58 us -- 2.8 us per unit, a third of it the serial head of every unit: unit
// index -> (class, head, tile), the fast division for the tile's reference point, the window origin, indexed constant loads.
struct __align__(16) UnitLevel {
  int ox, oy;            // window origin
  float oxf, oyf, xmaxf, ymaxf;   // in-window test on the floored sample position (xmaxf = -3e38: no window for this level)
  int vrel, ww;          // window base - (oy * ww + ox) * 64 relative to the stage; window width in pixels
};
struct __align__(16) UnitDesc {
  int c, h, tx, ty;
  UnitLevel lv[4];
};
constexpr int MAX_CTA_UNITS = 32;      // units per CTA covered by descriptors (more: computed on the fly, as before)

__device__ __forceinline__ UnitDesc describe_unit(const Params &P, const float *__restrict__ vr, int u) {
  UnitDesc d;
  const Unit t = unit_of(P, u);
  const ClassGeom &G = P.cls[t.c];
  d.c = t.c, d.h = t.h, d.tx = t.tx, d.ty = t.ty;
  // (any origin is correct as long as the copy and the records agree: both read this descriptor)
  const float rx = __fdividef((float)(t.tx * G.tw) + 0.5f, __ldg(vr + 2 * G.lq) * (float)G.Wq);
  const float ry = __fdividef((float)(t.ty * G.th) + 0.5f, __ldg(vr + 2 * G.lq + 1) * (float)G.Hq);
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const int2 org = window_origin(P, G, vr, rx, ry, t.h, l);
    const int ww = G.ww[l], wh = G.wh[l];
    UnitLevel &L = d.lv[l];
    L.ox = org.x, L.oy = org.y, L.oxf = (float)org.x, L.oyf = (float)org.y;
    L.xmaxf = ww ? (float)(org.x + ww - 2) : -3e38f, L.ymaxf = (float)(org.y + wh - 2);
    L.vrel = G.off[l] - (org.y * ww + org.x) * 64, L.ww = ww;
  }
  return d;
}

#define MEMOTR_BAR_ARRIVE(id) asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(THREADS) : "memory")
#define MEMOTR_BAR_SYNC(id) asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(THREADS) : "memory")

__global__ void __launch_bounds__(THREADS, 1)
msda_window_k4l4_kernel(const __grid_constant__ Maps maps, const __grid_constant__ Params P, const __half *__restrict__ value,
                        const float *__restrict__ loc, const float *__restrict__ attn, const float *__restrict__ vr,
                        unsigned long long *__restrict__ stats, __nv_bfloat16 *__restrict__ out) {
  constexpr int L = 4, LK = 16;
  constexpr int REC_FULL = 1, BUF_FREE = 3;       // named barriers REC_FULL + b, BUF_FREE + b (b = buffer 0 / 1)
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t win_full[2];
  const int tid = threadIdx.x, H = P.H;

  // ------------------------------------------------------------------------------------------- global-memory role
  if ((int)blockIdx.x < P.n_glob_blocks) {
    pdl_grid_sync();
    const int n_qh = (P.S - P.glob_q0) * H;
    const int qh = (blockIdx.x * THREADS + tid) >> 2, sub = tid & 3;
    if (qh >= n_qh) return;
    const int q = P.glob_q0 + qh / H, m = qh % H;
    float acc0[8], acc1[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc0[c] = acc1[c] = 0.f;
    h16::gather_global<0>(value + m * 32 + sub * 8, h16::ShapesI32{P.hw, P.lsi},
                           reinterpret_cast<const float2 *>(loc + (long)q * P.ld_loc) + m * LK,
                           attn + (long)q * P.ld_attn + m * LK, L, 4, P.xs, 0, 1, acc0, acc1);
#pragma unroll
    for (int c = 0; c < 8; ++c) acc0[c] += acc1[c];
    *reinterpret_cast<uint4 *>(out + ((long)q * H + m) * 32 + sub * 8) = f32x8_to_bf16(acc0);
    return;
  }

  // ------------------------------------------------------------------------------------------- window role (persistent)
  const int n_units = P.cls[P.n_cls - 1].unit0 + P.cls[P.n_cls - 1].n_units;
  const int stride = gridDim.x - P.n_glob_blocks;
  const int first = blockIdx.x - P.n_glob_blocks;
  __shared__ UnitDesc udesc[MAX_CTA_UNITS];
  if (tid == 0) {
    tc::mbar_init(win_full, 1), tc::mbar_init(win_full + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_grid_sync();
  if (tid < MAX_CTA_UNITS && first + tid * stride < n_units) udesc[tid] = describe_unit(P, vr, first + tid * stride);
  __syncthreads();

  if (tid >= TAP_WARPS * 32) {
    // =========================================================================================== decode warps
    const int dt = tid - TAP_WARPS * 32, l = dt & 3, qd = dt >> 2;      // this thread: level l of the queries qd, qd + 64
    const float Hf = (float)P.hw[2 * l], Wf = (float)P.hw[2 * l + 1];
    int n_win = 0, n_glob = 0;                   // (profiling only)
    float4 pre_xy[2][2], pre_w[2];
    auto prefetch = [&](const Unit &t) {        // locations / weights of the unit's points, one unit ahead of their use
      const ClassGeom &G = P.cls[t.c];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int ql = qd + 64 * i;
        const int x = t.tx * G.tw + (ql & (G.tw - 1)), y = t.ty * G.th + (ql >> G.tw_shift);
        pre_w[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        pre_xy[i][0] = pre_xy[i][1] = pre_w[i];
        if (ql < G.tw * G.th && x < G.Wq && y < G.Hq) {
          const long q = G.q0 + y * G.Wq + x;
          const float4 *pl = reinterpret_cast<const float4 *>(loc + q * P.ld_loc + (t.h * LK + l * 4) * 2);
          pre_xy[i][0] = __ldg(pl), pre_xy[i][1] = __ldg(pl + 1);
          pre_w[i] = __ldg(reinterpret_cast<const float4 *>(attn + q * P.ld_attn + t.h * LK + l * 4));
        }
      }
    };
    auto unit_at = [&](int n, int unit) {        // the n-th unit of this CTA: from its descriptor when there is one
      if (n < MAX_CTA_UNITS) {
        const int4 v = *reinterpret_cast<const int4 *>(&udesc[n]);
        Unit t;
        t.c = v.x, t.h = v.y, t.tx = v.z, t.ty = v.w;
        return t;
      }
      return unit_of(P, unit);
    };
    Unit t = unit_at(0, first < n_units ? first : 0);
    if (first < n_units) prefetch(t);
    int n = 0;
    for (int unit = first; unit < n_units; unit += stride, ++n) {
      const int b = n & 1;
      const ClassGeom &G = P.cls[t.c];
      const int TQ = G.tw * G.th;
      const uint32_t wbase = (uint32_t)(b * P.stage_bytes);
      UnitLevel UL;
      if (n < MAX_CTA_UNITS) {
        UL = udesc[n].lv[l];
      } else {
        UL = describe_unit(P, vr, unit).lv[l];
      }
      const int ww = UL.ww;
      const int2 org = make_int2(UL.ox, UL.oy);
      const float oxf = UL.oxf, oyf = UL.oyf, xmaxf = UL.xmaxf, ymaxf = UL.ymaxf;
      const int vbase = (int)wbase + UL.vrel;                                   // byte offset of the level's pixel (0, 0)
      if (n >= 2) MEMOTR_BAR_SYNC(BUF_FREE + b);                               // the tap warps have finished unit n - 2
      if (dt < L) {
        if (dt == 0) tc::mbar_expect_tx(win_full + b, (uint32_t)G.win_bytes);
        __syncwarp(0xfu);
        if (ww)
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                  tc::smem_u32(smem + wbase + G.off[l])),
              "l"(&maps.m[t.c * MAXL + l]), "r"(tc::smem_u32(win_full + b)), "r"(t.h * 32), "r"(org.x), "r"(org.y)
              : "memory");
      }
      uint8_t *recs = smem + wbase + G.rec_off + l * 64;
      uint8_t *flags = smem + wbase + G.flag_off + l;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int ql = qd + 64 * i;
        if (ql >= TQ || (P.debug == 1 && n >= 2)) continue;
        uint4 r[4];                                 // per point {off side 0, w side 0, off side 1, w side 1}
#pragma unroll
        for (int j = 0; j < 4; ++j) r[j] = make_uint4(wbase, 0u, wbase, 0u);
        const int x = t.tx * G.tw + (ql & (G.tw - 1)), y = t.ty * G.th + (ql >> G.tw_shift);
        if (x < G.Wq && y < G.Hq) {      // (a query outside the level: zero-weight taps on the first bytes of the buffer)
          const float px[4] = {pre_xy[i][0].x, pre_xy[i][0].z, pre_xy[i][1].x, pre_xy[i][1].z};
          const float py[4] = {pre_xy[i][0].y, pre_xy[i][0].w, pre_xy[i][1].y, pre_xy[i][1].w};
          const float pw[4] = {pre_w[i].x, pre_w[i].y, pre_w[i].z, pre_w[i].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float h_im = __fmaf_rn(py[j], Hf, -0.5f), w_im = __fmaf_rn(px[j], Wf, -0.5f);
            const float fy = floorf(h_im), fx = floorf(w_im);
            if (fx >= oxf && fx <= xmaxf && fy >= oyf && fy <= ymaxf) {
              const float lh = h_im - fy, lw = w_im - fx, hh = 1.f - lh, hw = 1.f - lw, aw = pw[j];
              const __half2 w0 = __floats2half2_rn(hh * hw * aw, lh * hw * aw), w1 = __floats2half2_rn(hh * lw * aw, lh * lw * aw);
              const uint32_t off = (uint32_t)(vbase + (__float2int_rz(fy) * ww + __float2int_rz(fx)) * 64);
              r[j] = make_uint4(off, *reinterpret_cast<const uint32_t *>(&w0), off + 64u, *reinterpret_cast<const uint32_t *>(&w1));
              ++n_win;
            } else {
              r[j] = slow_record(P.hw, P.lsi, l, wbase, make_float2(px[j], py[j]), pw[j]);
              n_glob += (int)(r[j].x >> 31);
            }
          }
        }
        // records of a query: per pair of points a 32-byte block [side 0: point 2k, 2k+1 | side 1: point 2k, 2k+1]
        uint4 *dst = reinterpret_cast<uint4 *>(recs + ql * G.rec_stride);
        dst[0] = make_uint4(r[0].x, r[0].y, r[1].x, r[1].y);
        dst[1] = make_uint4(r[0].z, r[0].w, r[1].z, r[1].w);
        dst[2] = make_uint4(r[2].x, r[2].y, r[3].x, r[3].y);
        dst[3] = make_uint4(r[2].z, r[2].w, r[3].z, r[3].w);
        flags[ql * 4] = (uint8_t)((r[0].x | r[1].x | r[2].x | r[3].x) >> 31);
      }
      if (unit + stride < n_units) {
        t = unit_at(n + 1, unit + stride);
        prefetch(t);
      }
      __syncwarp();
      MEMOTR_BAR_ARRIVE(REC_FULL + b);       // (the record stores above are ordered before the consumers' reads)
    }
    if (stats) atomicAdd(stats, (unsigned long long)n_win), atomicAdd(stats + 1, (unsigned long long)n_glob);
    return;
  }

  // ============================================================================================= tap warps
  const int grp = tid >> 3, side = (tid >> 2) & 1, sub = tid & 3;
  const uint32_t smem0 = tc::smem_u32(smem), lane_base = smem0 + sub * 16;
  const __half *vb0 = value + sub * 8;
  const int xs = P.xs;
  int n = 0;
  for (int unit = first; unit < n_units; unit += stride, ++n) {
    const int b = n & 1;
    Unit t;
    if (n < MAX_CTA_UNITS) {
      const int4 v = *reinterpret_cast<const int4 *>(&udesc[n]);
      t.c = v.x, t.h = v.y, t.tx = v.z, t.ty = v.w;
    } else {
      t = unit_of(P, unit);
    }
    const ClassGeom &G = P.cls[t.c];
    const int TQ = G.tw * G.th, rec_stride = G.rec_stride;
    const uint32_t stage = smem0 + (uint32_t)(b * P.stage_bytes);
    const uint32_t recs = stage + G.rec_off + side * 16, flags = stage + G.flag_off;
    const uint32_t rowb0 = (uint32_t)G.ww[0] * 64u, rowb1 = (uint32_t)G.ww[1] * 64u, rowb2 = (uint32_t)G.ww[2] * 64u,
                   rowb3 = (uint32_t)G.ww[3] * 64u;
    const int x = t.tx * G.tw + (grp & (G.tw - 1)), ystep = 64 >> G.tw_shift;
    int y = t.ty * G.th + (grp >> G.tw_shift);
    const bool x_ok = side == 0 && x < G.Wq;
    __nv_bfloat16 *optr = out + ((long)(G.q0 + y * G.Wq + x) * H + t.h) * 32 + sub * 8;
    const long ostep = (long)ystep * G.Wq * H * 32;
    const __half *vb = vb0 + t.h * 32;
    MEMOTR_BAR_SYNC(REC_FULL + b);
    mbar_wait_sleepy(win_full + b, (n >> 1) & 1);
    for (int ql = grp; ql < TQ; ql += TAP_WARPS * 4, y += ystep, optr += ostep) {
      if (P.debug == 2) continue;
      const uint32_t rq = recs + ql * rec_stride;
      uint32_t flag;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(flag) : "r"(flags + ql * 4));
      float acc[8];
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) acc[cc] = 0.f;
      __half2 a[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) a[jj] = __float2half2_rn(0.f);
      auto lds128 = [](uint32_t addr) {
        uint4 r;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
        return r;
      };
      if (!__any_sync(0xffffffffu, flag != 0u)) {
        // ---- every tap of these four queries comes from the windows -------------------------------------------------
#pragma unroll
        for (int l = 0; l < L; ++l) {
          const uint32_t rowb = l == 0 ? rowb0 : l == 1 ? rowb1 : l == 2 ? rowb2 : rowb3;
          const uint4 ra = lds128(rq + l * 64), rb = lds128(rq + l * 64 + 32);
          const uint32_t a0 = lane_base + ra.x, a1 = lane_base + ra.z, a2 = lane_base + rb.x, a3 = lane_base + rb.z;
          const uint4 v00 = lds128(a0), v01 = lds128(a0 + rowb), v10 = lds128(a1), v11 = lds128(a1 + rowb);
          const uint4 v20 = lds128(a2), v21 = lds128(a2 + rowb), v30 = lds128(a3), v31 = lds128(a3 + rowb);
          h16::blend(a, *reinterpret_cast<const __half2 *>(&ra.y), v00, v01);
          h16::blend(a, *reinterpret_cast<const __half2 *>(&ra.w), v10, v11);
          h16::blend(a, *reinterpret_cast<const __half2 *>(&rb.y), v20, v21);
          h16::blend(a, *reinterpret_cast<const __half2 *>(&rb.w), v30, v31);
          if (l & 1) {                             // end of a level pair: widen into the fp32 sums (msda_h16.cuh)
            h16::widen_add(acc, a);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) a[jj] = __float2half2_rn(0.f);
          }
        }
      } else {
        // ---- some lane has a global-memory tap: per-tap choice (same arithmetic) -------------------------------------
#pragma unroll 1
        for (int l = 0; l < L; ++l) {
          const uint32_t rowb = (uint32_t)G.ww[l] * 64u;
          const int Wl = P.hw[2 * l + 1];
          auto any_tap = [&](uint32_t off, uint4 &r0, uint4 &r1) {
            if (off & 0x80000000u) {
              const __half *p0 = vb + (long)(off & 0x3FFFFFFFu) * xs;
              const __half *p1 = (off & 0x40000000u) ? p0 : p0 + (long)Wl * xs;
              r0 = __ldg(reinterpret_cast<const uint4 *>(p0));
              r1 = __ldg(reinterpret_cast<const uint4 *>(p1));
            } else {
              r0 = lds128(lane_base + off), r1 = lds128(lane_base + off + rowb);
            }
          };
          const uint4 ra = lds128(rq + l * 64), rb = lds128(rq + l * 64 + 32);
          uint4 v[4][2];
          any_tap(ra.x, v[0][0], v[0][1]);
          any_tap(ra.z, v[1][0], v[1][1]);
          any_tap(rb.x, v[2][0], v[2][1]);
          any_tap(rb.z, v[3][0], v[3][1]);
          h16::blend(a, *reinterpret_cast<const __half2 *>(&ra.y), v[0][0], v[0][1]);
          h16::blend(a, *reinterpret_cast<const __half2 *>(&ra.w), v[1][0], v[1][1]);
          h16::blend(a, *reinterpret_cast<const __half2 *>(&rb.y), v[2][0], v[2][1]);
          h16::blend(a, *reinterpret_cast<const __half2 *>(&rb.w), v[3][0], v[3][1]);
          if (l & 1) {
            h16::widen_add(acc, a);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) a[jj] = __float2half2_rn(0.f);
          }
        }
      }
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) acc[cc] += __shfl_xor_sync(0xffffffffu, acc[cc], 4);
      if (x_ok && y < G.Hq) *reinterpret_cast<uint4 *>(optr) = f32x8_to_bf16(acc);
    }
    __syncwarp();
    MEMOTR_BAR_ARRIVE(BUF_FREE + b);
  }
}

// 3-D fp16 map of one level of a pixel-major value map: dims (channels, W, H), box (32 channels, ww, wh), no swizzle
static bool make_level_map(CUtensorMap *map, const void *base, int C, int xs, int Hh, int Ww, int ww, int wh) {
  tc::EncodeTiledFn fn = tc::encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)Ww, (cuuint64_t)Hh};
  cuuint64_t strides[2] = {(cuuint64_t)xs * 2, (cuuint64_t)Ww * xs * 2};
  cuuint32_t box[3] = {32, (cuuint32_t)ww, (cuuint32_t)wh};
  cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace win

int launch_h16(const void *value, const int64_t *shapes, const int64_t *lsi, const float *loc, const float *attn, void *out,
               int B, int S, int H, int L, int Lq, int K, int xs, cudaStream_t st, int ld_loc, int ld_attn);

}  // namespace memotr

using namespace memotr;

static int plan(win::Params &P, const int *shapes_hw, const int *level_start, int S, int H, int L, int K, int xs, int ld_loc,
                int ld_attn, const float *shift, float radius, int max_classes) {
  std::memset(&P, 0, sizeof(P));
  P.L = L, P.K = K, P.H = H, P.S = S, P.xs = xs, P.ld_loc = ld_loc, P.ld_attn = ld_attn, P.radius = radius;
  P.h_shift = -1;
  for (int b = 0; b < 5; ++b)
    if ((1 << b) == H) P.h_shift = b;
  for (int l = 0; l < L; ++l) P.hw[2 * l] = shapes_hw[2 * l], P.hw[2 * l + 1] = shapes_hw[2 * l + 1], P.lsi[l] = level_start[l];
  for (int h = 0; h < H; ++h)
    for (int l = 0; l < L; ++l)
      for (int d = 0; d < 2; ++d) P.shift[(h * win::MAXL + l) * 2 + d] = shift ? shift[(h * L + l) * 2 + d] : 0.f;
  const int LK = L * K;
  int units = 0, q_end = 0;
  for (int c = 0; c < std::min(std::min(max_classes, win::MAXC), L); ++c) {
    win::ClassGeom &G = P.cls[c];
    G.lq = c, G.Hq = shapes_hw[2 * c], G.Wq = shapes_hw[2 * c + 1], G.q0 = level_start[c];
    if (level_start[c] != q_end) break;                       // levels must be consecutive rows of the query table
    G.rec_stride = (LK / 2) * 32 + 32;                         // pad so that the four queries of a warp start in different banks
    while (G.rec_stride % 128 != 32) G.rec_stride += 32;
    // tile: 16 x 8 queries of level 0, 8 x 8 of level 1, 8 x 4 of the coarser levels (their reference points are 2x, 4x, ...
    // as far apart, so the windows of a tile grow with the level), halved while the records of the tile (one per sampling
    // point and x-side) would take more than 40 KB
    G.tw = c == 0 ? 16 : 8, G.th = c <= 1 ? 8 : 4, G.tw_shift = c == 0 ? 4 : 3;
    while (G.tw * G.th > 32 && (G.tw * G.th * G.rec_stride > 40 * 1024 || (K == 4 && G.tw * G.th * L > 2 * win::DEC_WARPS * 32))) {
      if (G.tw > G.th) G.tw >>= 1, --G.tw_shift; else G.th >>= 1;
    }
    if (G.tw * G.th * G.rec_stride > 64 * 1024) break;
    G.tiles_x = ceil_div(G.Wq, G.tw), G.tiles_y = ceil_div(G.Hq, G.th);
    G.tiles_x_rcp = 1.0f / (float)G.tiles_x;
    const int rec_bytes = G.tw * G.th * G.rec_stride;
    // window extents: footprint of the tile's reference points on the level (+5 % for differing valid ratios) + the radius
    // on both sides + the second bilinear column / row + rounding
    for (int l = 0; l < L; ++l) {
      const float sx = (float)(G.tw - 1) * 1.05f * (float)shapes_hw[2 * l + 1] / (float)G.Wq;
      const float sy = (float)(G.th - 1) * 1.05f * (float)shapes_hw[2 * l] / (float)G.Hq;
      G.ww[l] = std::min(std::min(255, shapes_hw[2 * l + 1] + 2), (int)std::ceil(sx + 2.f * radius) + 2);
      G.wh[l] = std::min(std::min(255, shapes_hw[2 * l] + 2), (int)std::ceil(sy + 2.f * radius) + 2);
    }
    // drop the largest windows until everything fits next to the records (their taps then come from global memory)
    for (;;) {
      int bytes = 0, big = -1;
      for (int l = 0; l < L; ++l) {
        if (!G.ww[l]) continue;
        if (big < 0 || G.ww[l] * G.wh[l] > G.ww[big] * G.wh[big]) big = l;
        bytes += (G.ww[l] * G.wh[l] * 64 + 127) / 128 * 128;
      }
      if (bytes + rec_bytes + G.tw * G.th * win::MAXL + 256 <= win::SMEM_BUDGET || big < 0) break;
      G.ww[big] = G.wh[big] = 0;
    }
    int off = 0, any = 0;
    for (int l = 0; l < L; ++l) {
      G.off[l] = off;
      if (G.ww[l]) off += (G.ww[l] * G.wh[l] * 64 + 127) / 128 * 128, G.win_bytes += G.ww[l] * G.wh[l] * 64, any = 1;
    }
    if (!any || off >= (1 << 20)) break;                       // nothing fits: leave this class to the global-memory role
    G.rec_off = off;
    G.flag_off = off + rec_bytes;
    G.unit0 = units, G.n_units = G.tiles_x * G.tiles_y * H;
    units += G.n_units;
    q_end = G.q0 + G.Hq * G.Wq;
    P.n_cls = c + 1;
  }
  P.glob_q0 = q_end;
  P.n_glob_blocks = ceil_div((S - q_end) * H * 4, win::THREADS);
  P.stage_bytes = 0;
  for (int c = 0; c < P.n_cls; ++c)
    P.stage_bytes = std::max(P.stage_bytes, (P.cls[c].flag_off + P.cls[c].tw * P.cls[c].th * win::MAXL + 127) / 128 * 128);
  return units;
}

// The staging plan memotr_msda_forward_window would use, as integers (host-only; tests and bench reporting):
// info[0..7] = {classes, window units, global-role CTAs, first global-role query, dynamic shared memory bytes, 0, 0, 0};
// then per class 8 + 2*MAXL ints: {query level, tw, th, tiles_x, tiles_y, units, window bytes, record stride, ww[5], wh[5]}.
extern "C" int memotr_msda_window_plan(const int *shapes_hw, const int *level_start, int S, int H, int L, int K, float radius,
                                       int max_classes, int *info) {
  MEMOTR_REQUIRE(shapes_hw && level_start && info && L >= 1 && L <= win::MAXL && H >= 1 && H <= win::MAXH && K >= 2 && K % 2 == 0,
                 "msda_window_plan: bad arguments");
  win::Params P;
  const int units = plan(P, shapes_hw, level_start, S, H, L, K, H * 32, 0, 0, nullptr, radius,
                         max_classes <= 0 ? win::MAXC : max_classes);
  const size_t smem = 2 * (size_t)P.stage_bytes;
  std::memset(info, 0, sizeof(int) * (8 + win::MAXC * (8 + 2 * win::MAXL)));
  info[0] = P.n_cls, info[1] = units, info[2] = P.n_glob_blocks, info[3] = P.glob_q0, info[4] = (int)smem;
  for (int c = 0; c < P.n_cls; ++c) {
    const win::ClassGeom &G = P.cls[c];
    int *o = info + 8 + c * (8 + 2 * win::MAXL);
    o[0] = G.lq, o[1] = G.tw, o[2] = G.th, o[3] = G.tiles_x, o[4] = G.tiles_y, o[5] = G.n_units, o[6] = G.win_bytes, o[7] = G.rec_stride;
    for (int l = 0; l < L; ++l) o[8 + l] = G.ww[l], o[8 + win::MAXL + l] = G.wh[l];
  }
  return MEMOTR_OK;
}

// Encoder-shaped gather (queries = the S pixels of the pyramid, batch 1) from TMA-staged windows.  `window_shift` (host
// memory, (H, L, 2) floats, may be null): expected sampling offset of each head on each level in pixels of that level --
// for MSDeformAttn the mean over the K points of sampling_offsets.bias; `window_radius`: how far around it the samples
// spread (pixels).  Both only steer what is staged: taps outside a window are read from global memory.  `stats` (device,
// 2 x u64, may be null): profiling counters += {taps taken from windows, taps left to global memory} of the window units.
extern "C" int memotr_msda_forward_window(const void *value, int value_pixel_stride, const int *shapes_hw, const int *level_start,
                                          const float *sampling_loc, int ld_loc, const float *attn_weight, int ld_attn,
                                          const float *valid_ratios, const float *window_shift, float window_radius,
                                          int max_classes, unsigned long long *stats, void *output, int S, int H, int L, int K,
                                          void *stream) {
  MEMOTR_REQUIRE(value && shapes_hw && level_start && sampling_loc && attn_weight && valid_ratios && output,
                 "msda_forward_window: null pointer");
  MEMOTR_REQUIRE(L >= 1 && L <= win::MAXL && H >= 1 && H <= win::MAXH && K >= 2 && K % 2 == 0 && K <= 64,
                 "msda_forward_window: needs 1..5 levels, 1..16 heads, an even number of points per level");
  MEMOTR_REQUIRE(H * 32 <= value_pixel_stride && value_pixel_stride % 8 == 0 && (long)S * value_pixel_stride < (1L << 30),
                 "msda_forward_window: bad pixel stride");
  MEMOTR_REQUIRE(aligned16(value) && aligned16(output) && ((reinterpret_cast<uintptr_t>(sampling_loc) & 7u) == 0) &&
                     ld_loc % 2 == 0 && ld_loc >= H * L * K * 2 && ld_attn >= H * L * K,
                 "msda_forward_window: misaligned buffer / bad row stride");
  MEMOTR_REQUIRE(K != 4 || (aligned16(sampling_loc) && aligned16(attn_weight) && ld_loc % 4 == 0 && ld_attn % 4 == 0 && (H * L * K) % 4 == 0),
                 "msda_forward_window: K == 4 needs 16-byte aligned location / weight rows");
  MEMOTR_REQUIRE(window_radius >= 0.f && window_radius <= 32.f, "msda_forward_window: radius out of range");
  MEMOTR_REQUIRE(tc::encode_fn() != nullptr, "msda_forward_window: cuTensorMapEncodeTiled unavailable");
  int total = 0;
  for (int l = 0; l < L; ++l) {
    MEMOTR_REQUIRE(level_start[l] == total && shapes_hw[2 * l] > 0 && shapes_hw[2 * l + 1] > 0,
                   "msda_forward_window: level_start must be the running sum of H_l * W_l");
    total += shapes_hw[2 * l] * shapes_hw[2 * l + 1];
  }
  MEMOTR_REQUIRE(total == S, "msda_forward_window: S != sum of the level sizes");
  win::Params P;
  win::Maps M;
  const int units = plan(P, shapes_hw, level_start, S, H, L, K, value_pixel_stride, ld_loc, ld_attn, window_shift, window_radius,
                         max_classes <= 0 ? win::MAXC : max_classes);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = 2 * (size_t)P.stage_bytes;       // two stages: the decode warps fill one while the tap warps drain the other
  std::memset(&M, 0, sizeof(M));
  for (int c = 0; c < P.n_cls; ++c) {
    const win::ClassGeom &G = P.cls[c];
    for (int l = 0; l < L; ++l)
      if (G.ww[l] && !win::make_level_map(&M.m[c * win::MAXL + l],
                                          reinterpret_cast<const __half *>(value) + (long)level_start[l] * value_pixel_stride, H * 32,
                                          value_pixel_stride, shapes_hw[2 * l], shapes_hw[2 * l + 1], G.ww[l], G.wh[l]))
        return fail(MEMOTR_ECUDA, "msda_forward_window: cuTensorMapEncodeTiled failed (class %d, level %d)", c, l);
  }
  // K = 4, L = 4: the specialised kernel (MEMOTR_WINDOW_GENERIC=1: the general one, for A/B and its tests)
  const char *gen = getenv("MEMOTR_WINDOW_GENERIC");
  const int which = (K == 4 && L == 4 && !(gen && gen[0] == '1')) ? 2 : K == 4 ? 1 : 0;
  auto kern = which == 2 ? win::msda_window_k4l4_kernel : which == 1 ? win::msda_window_kernel<4> : win::msda_window_kernel<0>;
  static bool attr_set[3] = {false, false, false};
  if (!attr_set[which]) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * win::SMEM_BUDGET);
    if (e != cudaSuccess) return fail(MEMOTR_ECUDA, "msda_forward_window: smem attribute: %s", cudaGetErrorString(e));
    attr_set[which] = true;
  }
  // persistent window CTAs: one per SM (two stages of shared memory each)
  int dev = 0, n_sm = kNumSMs;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const int win_ctas = std::min(units, sm_limit(n_sm));
  const char *dbg = getenv("MEMOTR_WINDOW_DEBUG");
  P.debug = dbg ? atoi(dbg) : 0;
  MEMOTR_LAUNCH((kern), P.n_glob_blocks + win_ctas, win::THREADS, smem, st, M, P, (const __half *)value, sampling_loc,
                attn_weight, valid_ratios, stats, (__nv_bfloat16 *)output);
  return check_launch("msda_window");
}
