// mlp_tc.cu -- fused two-layer MLP / FFN on the tensor cores:  C = epi( relu(X . W1^T + b1) . W2^T + b2 ),
// X (M,256) bf16, W1 (Hd,256), W2 (256,Hd), Hd a multiple of 128 (256 for the MLPs, 2048 for the FFNs).
//
// Replaces `linear2(dropout(activation(linear1(x))))` of the encoder / decoder FFNs
// (models/deformable_encoder.py:97-107, models/deformable_decoder.py:263-273, models/ffn.py:15-22) and the two-layer
// MLPs with 256-wide input (models/mlp.py:22-25 as used at deformable_decoder.py:93 query_scale, :140 bbox_embed layers
// 0-1, query_updater.py:109 confidence net).  The reference materialises the (M, Hd) hidden activation in HBM
// (91 MB per encoder layer in bf16, written once and read once); here it never leaves the SM.
//
// One CTA owns a 128-row tile of X (TMA-loaded once, 64 KB) and walks the hidden dimension in chunks of 128:
//     GEMM1(c): acc1[c&1] (TMEM, 128 cols) = X . W1[c]^T                 4 k-blocks, tcgen05.mma M128 N128 K16
//     epi1(c) : acc1 -> +b1 -> ReLU -> bf16 -> shared memory H[c&1]      written directly in the 128B-swizzled K-major
//                                                                        layout UMMA wants for an A operand
//     GEMM2(c): acc2 (TMEM, 256 cols) += H[c&1] . W2[:, c]^T             2 k-panels, tcgen05.mma M128 N256 K16
// Warp roles: 0 = TMA producer streaming W1/W2 chunks through a 3-slot x 32 KB ring, 1 = MMA issuer (GEMM1 runs one
// chunk ahead of GEMM2 so epi1(c) overlaps GEMM1(c+1)), 2-9 = epilogue (two warps per TMEM lane quarter, each taking half
// of the columns, two tcgen05.ld in flight per warp: with four warps the per-chunk epilogue, ~4000 clk, was the critical
// path -- 82 us per encoder FFN whatever the weight-load strategy).  acc1 and H are double-buffered; all hand-offs
// are mbarriers (TMA expect_tx, tcgen05.commit, and 128-thread arrives from the epilogue warps).  TMEM: 2x128 + 256 =
// 512 columns.  Shared memory: X 64 KB + ring 96 KB + H 64 KB = 224 KB (one CTA per SM).  Final epilogue as in
// gemm_tc.cu: bias / activation / multiplier -> swizzled panels in the (dead) X+ring memory -> TMA store.
//
// What bounds it (profiles/r01_mlp2_ncu.md): the 128 B/clk shared-memory pipe -- per 128-wide hidden chunk 224 KB of MMA
// operand reads + 128 KB of TMA writes + 32 KB of epilogue stores = ~3000 clk against 2048 clk of tcgen05 work; a tile
// costs 6.6 us fixed + 16 x 1.65 us.  One CTA per SM (224 KB), so 175 encoder row tiles on 148 SMs are two rounds: launches
// with few row tiles split the hidden dimension over gridDim.y (TMA reduce-add stores into the zeroed fp32 output), and the
// host runs the tiles beyond the first round that way: 73 -> 59 us per encoder FFN.
//
// LayerNorm epilogue (LnOut): for the encoder's `src = norm2(src + ffn(src))` (deformable_encoder.py:103-107,128-131) the
// final epilogue adds the fp32 residual, normalises the 256-wide row -- the whole row sits in this CTA's TMEM accumulator --
// and writes the three things the next layer reads: y (bf16, GEMM operand), y fp32 (residual master), y + pos (bf16, the
// query of the next layer's offset / weight projection).  That removes the stand-alone LayerNorm kernel and the fp32 round
// trip of the pre-norm sum (19 us + 46 MB per layer).  Variants measured slower in round 1 and removed: LayerNorm in the
// PROLOGUE, weight multicast in clusters, uniform split-K.
#include "tc_common.cuh"

namespace memotr {
namespace tc {

namespace mlp {
constexpr int K1 = 256, N2 = 256, HC = 128;
constexpr int XP = K1 / BK;                     // 4 X panels of 128 rows x 64 cols
constexpr int PANEL = BM * 128;                 // 16 KB: 128 rows x 128 bytes
constexpr int X_BYTES = XP * PANEL;             // 64 KB
constexpr int SLOT = 2 * PANEL;                 // 32 KB: two W1 k-blocks (128 rows) or one W2 k-panel (256 rows)
constexpr int NSLOT = 3;
constexpr int H_BYTES = 2 * PANEL;              // one hidden chunk as A operand: 2 panels of 64 columns
constexpr int OFF_RING = X_BYTES;
constexpr int OFF_H = OFF_RING + NSLOT * SLOT;
constexpr int OFF_BAR = OFF_H + 2 * H_BYTES;    // 224 KB
constexpr int TOTAL = OFF_BAR + 256 + 1024;
}  // namespace mlp

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// LayerNorm epilogue of the final GEMM (res == nullptr: off): out = LayerNorm(acc2 + b2 + res) * gamma + beta
struct LnOut {
  const float *res, *gamma, *beta;   // fp32 residual rows, LayerNorm affine
  float *y32;                        // fp32 copy of the result (may be null)
  const void *pos;                   // bf16 rows added to the result for the second bf16 output (null: no second output)
  int ldres, ld32, ldpos;
  float eps;
};

template <typename TC>
__global__ void __launch_bounds__(320, 1)
mlp2_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
               const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmC,
               const __grid_constant__ CUtensorMap tmQ, const float *__restrict__ b1, int M, int Hd, Epilogue ep, int tile0,
               LnOut ln) {
  // tile0: first row tile of this launch (the tail tiles of a GEMM are launched separately with a hidden-dimension split)
  // gridDim.y > 1: split-K over the hidden dimension -- CTA (x, y) handles hidden chunks [y*NC, (y+1)*NC) of row tile x
  // and ADDS its partial product into the (zero-initialised, fp32) output with a TMA reduce-store; bias from split 0.
  using namespace mlp;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + OFF_BAR);
  uint64_t *x_full = bars, *full = bars + 1, *empty = full + NSLOT, *acc1_full = empty + NSLOT, *acc1_empty = acc1_full + 2,
           *h_full = acc1_empty + 2, *h_empty = h_full + 2, *acc2_full = h_empty + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc2_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blk = blockIdx.x + tile0;
  const int NC = Hd / HC / (int)gridDim.y;          // hidden chunks of this CTA
  const int c_off = (int)blockIdx.y * NC;           // first hidden chunk of this CTA
  const bool split = gridDim.y > 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    mbar_init(x_full, 1);
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc1_full + b, 1);
      mbar_init(acc1_empty + b, 256);
      mbar_init(h_full + b, 256);
      mbar_init(h_empty + b, 1);
    }
    mbar_init(acc2_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_grid_sync();

  uint8_t *ring = smem + OFF_RING;
  uint8_t *hbuf = smem + OFF_H;

  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer: X once, then W1(0), [W1(c+1), W2(c)] ... in exactly the order the MMA warp consumes ----
      mbar_expect_tx(x_full, X_BYTES);
      for (int p = 0; p < XP; ++p) tma_load_2d(smem + p * PANEL, &tmX, x_full, p * BK, m_blk * BM);
      int t = 0;
      auto load_w1 = [&](int c) {          // slot = [k-block 2*half: 128 rows][k-block 2*half+1: 128 rows]
        for (int half = 0; half < 2; ++half, ++t) {
          const int s = t % NSLOT;
          mbar_wait(empty + s, ((t / NSLOT) & 1) ^ 1);
          mbar_expect_tx(full + s, SLOT);
          tma_load_2d(ring + s * SLOT, &tmW1, full + s, (2 * half) * BK, (c_off + c) * HC);
          tma_load_2d(ring + s * SLOT + PANEL, &tmW1, full + s, (2 * half + 1) * BK, (c_off + c) * HC);
        }
      };
      auto load_w2 = [&](int c) {          // slot = 256 output rows x 64 hidden columns
        for (int j = 0; j < 2; ++j, ++t) {
          const int s = t % NSLOT;
          mbar_wait(empty + s, ((t / NSLOT) & 1) ^ 1);
          mbar_expect_tx(full + s, SLOT);
          tma_load_2d(ring + s * SLOT, &tmW2, full + s, (c_off + c) * HC + j * BK, 0);
        }
      };
      load_w1(0);
      for (int c = 0; c < NC; ++c) {
        if (c + 1 < NC) load_w1(c + 1);
        load_w2(c);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---- MMA issuer ----
      constexpr uint32_t idesc1 = umma_idesc(HC), idesc2 = umma_idesc(N2);
      mbar_wait(x_full, 0);
      tcgen05_fence_after();
      const uint32_t x_addr = smem_u32(smem), ring_addr = smem_u32(ring), h_addr = smem_u32(hbuf);
      int t = 0;
      auto gemm1 = [&](int c) {
        const int b = c & 1;
        mbar_wait(acc1_empty + b, ((c >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        for (int half = 0; half < 2; ++half, ++t) {
          const int s = t % NSLOT;
          mbar_wait(full + s, (t / NSLOT) & 1);
          tcgen05_fence_after();
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const int kb = 2 * half + kk;
            const uint64_t adesc = umma_desc(x_addr + kb * PANEL), bdesc = umma_desc(ring_addr + s * SLOT + kk * PANEL);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(tmem_base + b * HC, adesc + 2 * k, bdesc + 2 * k, idesc1, (kb | k) != 0);
          }
          umma_commit(empty + s);
        }
        umma_commit(acc1_full + b);
      };
      auto gemm2 = [&](int c) {
        const int b = c & 1;
        mbar_wait(h_full + b, (c >> 1) & 1);
        tcgen05_fence_after();
        for (int j = 0; j < 2; ++j, ++t) {
          const int s = t % NSLOT;
          mbar_wait(full + s, (t / NSLOT) & 1);
          tcgen05_fence_after();
          const uint64_t adesc = umma_desc(h_addr + b * H_BYTES + j * PANEL), bdesc = umma_desc(ring_addr + s * SLOT);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(tmem_base + 2 * HC, adesc + 2 * k, bdesc + 2 * k, idesc2, (c | j | k) != 0);
          umma_commit(empty + s);
        }
        umma_commit(h_empty + b);
      };
      gemm1(0);
      for (int c = 0; c < NC; ++c) {
        if (c + 1 < NC) gemm1(c + 1);
        gemm2(c);
      }
      umma_commit(acc2_full);
    }
  } else {
    // ---- epilogue warps 2..9: lane quarter = warp % 4 (one accumulator row per thread), column half = (warp - 2) / 4 ----
    const int quarter = warp & 3, chalf = (warp - 2) >> 2;
    const int r_in = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    for (int c = 0; c < NC; ++c) {
      const int b = c & 1;
      mbar_wait(acc1_full + b, (c >> 1) & 1);
      tcgen05_fence_after();
      mbar_wait(h_empty + b, ((c >> 1) & 1) ^ 1);   // GEMM2(c-2) has finished reading this H buffer
      // this warp: hidden columns [chalf*64, chalf*64+64) of the chunk = H panel `chalf`, both 32-column halves in flight
      uint32_t r0[32], r1[32];
      tmem_ld32_issue(tmem_base + lane_off + (uint32_t)(b * HC + chalf * 64), r0);
      tmem_ld32_issue(tmem_base + lane_off + (uint32_t)(b * HC + chalf * 64 + 32), r1);
      tmem_ld_wait();
      uint8_t *prow = hbuf + b * H_BYTES + chalf * PANEL + r_in * 128;
      const float *bias = b1 + (c_off + c) * HC + chalf * 64;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 ba = __ldg(reinterpret_cast<const float4 *>(bias + 8 * k));
        const float4 bb = __ldg(reinterpret_cast<const float4 *>(bias + 8 * k + 4));
        float t8[8];
#define RV(i) __uint_as_float(k < 4 ? r0[8 * k + (i)] : r1[8 * (k - 4) + (i)])
        t8[0] = fmaxf(RV(0) + ba.x, 0.f);
        t8[1] = fmaxf(RV(1) + ba.y, 0.f);
        t8[2] = fmaxf(RV(2) + ba.z, 0.f);
        t8[3] = fmaxf(RV(3) + ba.w, 0.f);
        t8[4] = fmaxf(RV(4) + bb.x, 0.f);
        t8[5] = fmaxf(RV(5) + bb.y, 0.f);
        t8[6] = fmaxf(RV(6) + bb.z, 0.f);
        t8[7] = fmaxf(RV(7) + bb.w, 0.f);
#undef RV
        *reinterpret_cast<uint4 *>(prow + ((k ^ (r_in & 7)) << 4)) = f32x8_to_bf16(t8);
      }
      tcgen05_fence_before();                                        // TMEM reads ordered before the hand-off
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // st.shared -> visible to the UMMA (async proxy)
      mbar_arrive(acc1_empty + b);
      mbar_arrive(h_full + b);
    }
    // ---- final epilogue: acc2 (+b2, activation, multiplier) -> swizzled panels in the dead X/ring memory -> TMA store;
    //      this warp takes output columns [chalf*128, chalf*128+128)
    mbar_wait(acc2_full, 0);
    tcgen05_fence_after();
    const int row = m_blk * BM + r_in;
    const bool row_ok = row < M;
    if (ln.res) {
      // ---- LayerNorm epilogue: this thread owns columns [chalf*128, +128) of its row; the other half of the row is with
      //      the warp four above / below, the two partial moments meet in shared memory (the dead H buffer) ----
      if constexpr (sizeof(TC) == 2) {
        const float *resp = ln.res + (long)row * ln.ldres;
        auto chunk = [&](int c0, float (&v)[32]) {            // acc2 + b2 + residual for 32 columns
          uint32_t r[32];
          tmem_ld32(tmem_base + lane_off + (uint32_t)(2 * HC + c0), r);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bb = __ldg(reinterpret_cast<const float4 *>(ep.bias + c0 + j));
            const float4 rr = row_ok ? *reinterpret_cast<const float4 *>(resp + c0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[j] = __uint_as_float(r[j]) + bb.x + rr.x, v[j + 1] = __uint_as_float(r[j + 1]) + bb.y + rr.y;
            v[j + 2] = __uint_as_float(r[j + 2]) + bb.z + rr.z, v[j + 3] = __uint_as_float(r[j + 3]) + bb.w + rr.w;
          }
        };
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
        for (int c0 = chalf * 128; c0 < chalf * 128 + 128; c0 += 32) {
          float v[32];
          chunk(c0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) s1 += v[j], s2 = fmaf(v[j], v[j], s2);
        }
        float2 *stat = reinterpret_cast<float2 *>(hbuf);
        stat[chalf * BM + r_in] = make_float2(s1, s2);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float2 o = stat[(1 - chalf) * BM + r_in];
        const float mean = (s1 + o.x) * (1.f / 256.f);
        const float rstd = rsqrtf(fmaxf((s2 + o.y) * (1.f / 256.f) - mean * mean, 0.f) + ln.eps);
        const __nv_bfloat16 *posp = (const __nv_bfloat16 *)ln.pos + (long)row * ln.ldpos;
#pragma unroll 1
        for (int c0 = chalf * 128; c0 < chalf * 128 + 128; c0 += 32) {
          float v[32];
          chunk(c0, v);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 g = __ldg(reinterpret_cast<const float4 *>(ln.gamma + c0 + j));
            const float4 be = __ldg(reinterpret_cast<const float4 *>(ln.beta + c0 + j));
            v[j] = (v[j] - mean) * rstd * g.x + be.x, v[j + 1] = (v[j + 1] - mean) * rstd * g.y + be.y;
            v[j + 2] = (v[j + 2] - mean) * rstd * g.z + be.z, v[j + 3] = (v[j + 3] - mean) * rstd * g.w + be.w;
          }
          if (ln.y32 && row_ok) {
            float4 *yp = reinterpret_cast<float4 *>(ln.y32 + (long)row * ln.ld32 + c0);
#pragma unroll
            for (int k = 0; k < 8; ++k) yp[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          }
          // y -> panels 0..3, y + pos -> panels 4..7 (64 bf16 columns per panel, 16-byte chunk k at k ^ (row & 7))
          const int panel = c0 / 64, kbase = (c0 % 64) / 8;
          uint8_t *prow = smem + panel * PANEL + r_in * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float t[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) t[i] = v[8 * k + i];
            *reinterpret_cast<uint4 *>(prow + (((kbase + k) ^ (r_in & 7)) << 4)) = f32x8_to_bf16(t);
            if (ln.pos) {
              float pz[8];
              bf16x8_to_f32(row_ok ? __ldg(reinterpret_cast<const uint4 *>(posp + c0 + 8 * k)) : make_uint4(0, 0, 0, 0), pz);
#pragma unroll
              for (int i = 0; i < 8; ++i) t[i] += pz[i];
              *reinterpret_cast<uint4 *>(prow + 4 * PANEL + (((kbase + k) ^ (r_in & 7)) << 4)) = f32x8_to_bf16(t);
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp == 2 && lane == 0) {
#pragma unroll 1
          for (int p = 0; p < 4; ++p) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmC),
                         "r"(smem_u32(smem + p * PANEL)), "r"(p * 64), "r"(m_blk * BM)
                         : "memory");
            if (ln.pos)
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmQ),
                           "r"(smem_u32(smem + (4 + p) * PANEL)), "r"(p * 64), "r"(m_blk * BM)
                           : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
      }
    } else {
    const __nv_bfloat16 *mulp = (const __nv_bfloat16 *)ep.mul + (long)row * ep.ldmul;
    constexpr int PANEL_COLS = 128 / (int)sizeof(TC);
    constexpr int N_PANELS = N2 / PANEL_COLS;
    static_assert(N_PANELS * PANEL <= OFF_H, "staging must fit in the X + ring area");
#pragma unroll 1
    for (int c0 = chalf * 128; c0 < chalf * 128 + 128; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + lane_off + (uint32_t)(2 * HC + c0), r);
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (ep.bias && blockIdx.y == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bb = __ldg(reinterpret_cast<const float4 *>(ep.bias + c0 + j));
          v[j] += bb.x, v[j + 1] += bb.y, v[j + 2] += bb.z, v[j + 3] += bb.w;
        }
      }
      if (ep.act == ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      } else if (ep.act == ACT_SIGMOID) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + __expf(-v[j]));
      }
      if (ep.mul && row_ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float t[8];
          bf16x8_to_f32(__ldg(reinterpret_cast<const uint4 *>(mulp + c0 + j)), t);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[j + i] *= t[i];
        }
      }
      uint8_t *prow = smem + (c0 / PANEL_COLS) * PANEL + r_in * 128;
      if constexpr (sizeof(TC) == 4) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          *reinterpret_cast<float4 *>(prow + ((k ^ (r_in & 7)) << 4)) =
              make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
      } else {
        const int kbase = (c0 % PANEL_COLS) / 8;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float t[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] = v[8 * k + i];
          *reinterpret_cast<uint4 *>(prow + (((kbase + k) ^ (r_in & 7)) << 4)) = f32x8_to_bf16(t);
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (warp == 2 && lane == 0) {
#pragma unroll 1
      for (int p = 0; p < N_PANELS; ++p) {
        if (split)
          asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                           &tmC),
                       "r"(smem_u32(smem + p * PANEL)), "r"(p * PANEL_COLS), "r"(m_blk * BM)
                       : "memory");
        else
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmC),
                       "r"(smem_u32(smem + p * PANEL)), "r"(p * PANEL_COLS), "r"(m_blk * BM)
                       : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

template <typename TC>
static int launch_mlp2(const void *X, int ldx, const void *W1, const float *b1, const void *W2, void *C, int ldc, int M,
                       int Hd, const Epilogue &ep, cudaStream_t st, int nsplit = 1, int tile0 = 0, int ntiles = -1,
                       const LnOut &ln = LnOut{}, void *Q = nullptr, int ldq = 0) {
  using namespace mlp;
  CUtensorMap tmX, tmW1, tmW2, tmC, tmQ;
  if (!make_map(&tmX, X, M, K1, ldx, BM) || !make_map(&tmW1, W1, Hd, K1, K1, HC) || !make_map(&tmW2, W2, N2, Hd, Hd, 256) ||
      !make_map(&tmC, C, M, N2, ldc, BM, sizeof(TC) == 4) || !make_map(&tmQ, Q ? Q : C, M, N2, Q ? ldq : ldc, BM, sizeof(TC) == 4))
    return fail(MEMOTR_ECUDA, "mlp2(tc): cuTensorMapEncodeTiled failed (M=%d Hd=%d)", M, Hd);
  auto kern = mlp2_tc_kernel<TC>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TOTAL);
    if (e != cudaSuccess) return fail(MEMOTR_ECUDA, "mlp2(tc): smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles = ntiles < 0 ? ceil_div(M, BM) - tile0 : ntiles;
  if (nsplit > 1) {   // partial products are reduce-added: start from zero (a memset node in a graph)
    const int r0 = tile0 * BM, nr = (M - r0 < tiles * BM) ? M - r0 : tiles * BM;
    const cudaError_t e = cudaMemset2DAsync(reinterpret_cast<TC *>(C) + (size_t)r0 * ldc, (size_t)ldc * sizeof(TC), 0,
                                            (size_t)N2 * sizeof(TC), nr, st);
    if (e != cudaSuccess) return fail(MEMOTR_ECUDA, "mlp2(tc): memset: %s", cudaGetErrorString(e));
  }
  MEMOTR_LAUNCH((kern), dim3(tiles, nsplit), 320, TOTAL, st, tmX, tmW1, tmW2, tmC, tmQ, b1, M, Hd, ep, tile0, ln);
  return check_launch("mlp2_tc");
}

}  // namespace tc
}  // namespace memotr

using namespace memotr;

static int sm_count() {
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  }
  return n_sm;
}

// hidden-dimension split for `tiles` row tiles: the largest power of two (<= 8) that still fits one CTA per SM
static int pick_split(int tiles, int chunks, int n_sm) {
  int ns = 1;
  while (ns * 2 <= 8 && tiles * ns * 2 <= n_sm && chunks % (ns * 2) == 0) ns *= 2;
  return ns;
}

// fp32-output FFN without activation / multiplier.  One CTA per SM: 175 row tiles on 148 SMs are two rounds of 33 us, the
// second 18 % full -- so the first n_sm tiles run as usual and the remaining ones (and any launch with few tiles) are
// launched with the hidden dimension split over as many CTAs as fit on the GPU (27 tiles x 4: 6.6 + 4 x 1.65 us).
// MEMOTR_MLP_TAIL=0 switches the split off (A/B).
static int mlp2_f32_balanced(const void *X, int ldx, const void *W1, const float *b1, const void *W2, void *C, int ldc, int M,
                             int Hd, const Epilogue &ep, cudaStream_t st) {
  const int n_sm = sm_count(), tiles = ceil_div(M, tc::BM), chunks = Hd / tc::mlp::HC;
  const char *tl = getenv("MEMOTR_MLP_TAIL");
  if (tl && tl[0] == '0') return tc::launch_mlp2<float>(X, ldx, W1, b1, W2, C, ldc, M, Hd, ep, st);
  if (tiles * 2 <= n_sm) return tc::launch_mlp2<float>(X, ldx, W1, b1, W2, C, ldc, M, Hd, ep, st, pick_split(tiles, chunks, n_sm));
  const int tail = tiles - n_sm;
  if (tail > 0 && tail * 2 <= n_sm) {
    const int ns = pick_split(tail, chunks, n_sm);
    if (ns > 1) {
      const int rc = tc::launch_mlp2<float>(X, ldx, W1, b1, W2, C, ldc, M, Hd, ep, st, 1, 0, n_sm);
      if (rc != MEMOTR_OK) return rc;
      return tc::launch_mlp2<float>(X, ldx, W1, b1, W2, C, ldc, M, Hd, ep, st, ns, n_sm, tail);
    }
  }
  return tc::launch_mlp2<float>(X, ldx, W1, b1, W2, C, ldc, M, Hd, ep, st);
}

extern "C" int memotr_mlp2(const void *X, int ldx, const void *W1, const float *b1, const void *W2, const float *b2,
                           const void *mul, int ldmul, void *C, int ldc, int M, int K1, int Hd, int N2, int c_dtype,
                           int act2, void *stream) {
  MEMOTR_REQUIRE(M >= 0 && X && W1 && b1 && W2 && C, "mlp2: bad arguments");
  MEMOTR_REQUIRE(K1 == tc::mlp::K1 && N2 == tc::mlp::N2 && Hd > 0 && Hd % tc::mlp::HC == 0,
                 "mlp2: needs K1 == 256, N2 == 256, hidden %% 128 == 0 (got %d, %d, %d)", K1, N2, Hd);
  MEMOTR_REQUIRE(c_dtype == MEMOTR_F32 || c_dtype == MEMOTR_BF16, "mlp2: output dtype must be f32 or bf16");
  MEMOTR_REQUIRE(act2 >= 0 && act2 <= 2, "mlp2: unknown activation");
  const int cal = c_dtype == MEMOTR_F32 ? 4 : 8;
  MEMOTR_REQUIRE(ldx % 8 == 0 && ldc % cal == 0 && aligned16(X) && aligned16(W1) && aligned16(W2) && aligned16(C) &&
                     aligned16(b1) && (!b2 || aligned16(b2)) && (!mul || (aligned16(mul) && ldmul % 8 == 0)),
                 "mlp2: misaligned buffer");
  MEMOTR_REQUIRE(tc::encode_fn() != nullptr, "mlp2: cuTensorMapEncodeTiled unavailable");
  if (M == 0) return MEMOTR_OK;
  Epilogue ep{b2, mul, nullptr, nullptr, ldmul, 0, act2};
  cudaStream_t st = (cudaStream_t)stream;
  if (c_dtype == MEMOTR_F32 && act2 == 0 && !mul) return mlp2_f32_balanced(X, ldx, W1, b1, W2, C, ldc, M, Hd, ep, st);
  if (c_dtype == MEMOTR_F32) return tc::launch_mlp2<float>(X, ldx, W1, b1, W2, C, ldc, M, Hd, ep, st);
  return tc::launch_mlp2<__nv_bfloat16>(X, ldx, W1, b1, W2, C, ldc, M, Hd, ep, st);
}

// The encoder FFN with its LayerNorm in the epilogue (deformable_encoder.py:103-107 + :128-131):
//   y = LayerNorm(res + relu(X W1^T + b1) W2^T + b2) * gamma + beta  ->  y (bf16), y32 (fp32, may be null), ypos = y + pos (bf16;
//   pos / ypos may be null).  One launch of ceil(M / 128) CTAs (callers with more row tiles than SMs pass the first round here
//   and run the remaining rows through memotr_mlp2 + memotr_layernorm, whose few tiles split the hidden dimension).
extern "C" int memotr_mlp2_lnout(const void *X, int ldx, const void *W1, const float *b1, const void *W2, const float *b2,
                                 const float *res, int ldres, const float *gamma, const float *beta, float eps, void *y, int ldy,
                                 float *y32, int ld32, const void *pos, int ldpos, void *ypos, int ldypos, int M, int Hd,
                                 void *stream) {
  MEMOTR_REQUIRE(M >= 0 && X && W1 && b1 && W2 && b2 && res && gamma && beta && y, "mlp2_lnout: bad arguments");
  MEMOTR_REQUIRE(Hd > 0 && Hd % tc::mlp::HC == 0, "mlp2_lnout: hidden %% 128 != 0 (got %d)", Hd);
  MEMOTR_REQUIRE((pos == nullptr) == (ypos == nullptr), "mlp2_lnout: pos and ypos come together");
  MEMOTR_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && ldres % 4 == 0 && (!y32 || ld32 % 4 == 0) && (!pos || (ldpos % 8 == 0 && ldypos % 8 == 0)) &&
                     aligned16(X) && aligned16(W1) && aligned16(W2) && aligned16(y) && aligned16(b1) && aligned16(b2) && aligned16(res) &&
                     aligned16(gamma) && aligned16(beta) && (!y32 || aligned16(y32)) && (!pos || (aligned16(pos) && aligned16(ypos))),
                 "mlp2_lnout: misaligned buffer");
  MEMOTR_REQUIRE(tc::encode_fn() != nullptr, "mlp2_lnout: cuTensorMapEncodeTiled unavailable");
  if (M == 0) return MEMOTR_OK;
  Epilogue ep{b2, nullptr, nullptr, nullptr, 0, 0, ACT_NONE};
  const tc::LnOut ln{res, gamma, beta, y32, pos, ldres, ld32, ldpos, eps};
  return tc::launch_mlp2<__nv_bfloat16>(X, ldx, W1, b1, W2, y, ldy, M, Hd, ep, (cudaStream_t)stream, 1, 0, -1, ln, ypos, ldypos);
}
