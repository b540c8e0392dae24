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
  void *y;                           // bf16 result
  float *y32;                        // fp32 copy of the result (may be null)
  const void *pos;                   // bf16 rows added to the result for the second bf16 output ypos (both null: none)
  void *ypos;
  int ldres, ldy, ld32, ldpos, ldypos;
  float eps;
};

// Front GEMM + LayerNorm (on == 0: off, X comes from tmX): X = LayerNorm(res0 + A W0^T + bias0) * gamma1 + beta1 computed in
// the kernel -- A (the attention rows, bf16) arrives through tmX, W0 (256 x 256) through the weight ring, the product sits in
// the (not yet used) second accumulator, the epilogue warps normalise it straight into the shared-memory A tile of the FFN.
// That is `src = norm1(src + output_proj(attn))` of deformable_encoder.py:124-126 + ms_deform_attn.py:129 without a GEMM
// launch, a LayerNorm launch and two round trips through HBM.  x32 receives the fp32 result (the residual of norm2).
struct Front {
  const float *bias0, *res0, *gamma1, *beta1;
  float *x32;
  int ldres0, ldx32, on;
  float eps;
  long long *stamps;                 // profiling (tools/micro_dense.py): 8 clock64 stamps per CTA, null in production
  float *zptr;                       // rows of an fp32 buffer this launch clears on the side (zrows x 256, row stride zld): the
  int zrows, zld;                    // output of the split-K launch that follows it, which ADDS its partial products
};

// Row-coalesced LayerNorm over a 128 x 256 fp32 tile staged in shared memory (8 panels of 32 columns, 128-byte rows, 16-byte
// chunk k of row r at k ^ (r & 7): the layout the accumulator dump below writes).  Warp w (0..7) owns rows 16w .. 16w+15, eight
// at a time; lane L owns columns 32 (L/8) + 4 (L%8) .. +3 and the same +128: a quarter-warp reads one 128-byte panel row
// (conflict-free) and touches 128 contiguous bytes of every global row.  v = staged + res; y = LN(v) * gamma + beta;
// `emit(row, r_in, ok, colA, colB, yA, yB, posA, posB)` consumes the two float4 of a lane.
// The phase is bound by bytes in flight per SM (measured: 4 rows per batch = 32 KB in flight = 10 us per 128 rows), so ALL
// global reads of a batch -- the residual and, when POS, the bf16 rows that emit() adds -- are issued before the first use.
template <bool POS, typename Emit>
__device__ __forceinline__ void ln_rows_256(const uint8_t *stage, int w, int lane, int row0, int M, const float *__restrict__ res,
                                            int ldres, const float *__restrict__ gamma, const float *__restrict__ beta, float eps,
                                            const __nv_bfloat16 *__restrict__ pos, int ldpos, Emit emit) {
  constexpr int PANEL = BM * 128, RB = 8;
  const int pA = lane >> 3, ck = lane & 7, colA = pA * 32 + ck * 4, colB = colA + 128;
  const float4 gA = ldg_f4(gamma + colA), gB = ldg_f4(gamma + colB), bA = ldg_f4(beta + colA), bB = ldg_f4(beta + colB);
#pragma unroll 1
  for (int rb = 0; rb < 16; rb += RB) {
    float4 ra[RB], rc[RB];
    uint2 qa[POS ? RB : 1], qb[POS ? RB : 1];
#pragma unroll
    for (int j = 0; j < RB; ++j) {
      const int row = row0 + w * 16 + rb + j;
      ra[j] = rc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (POS) qa[j] = qb[j] = make_uint2(0u, 0u);
      if (row < M) {
        ra[j] = *reinterpret_cast<const float4 *>(res + (long)row * ldres + colA);
        rc[j] = *reinterpret_cast<const float4 *>(res + (long)row * ldres + colB);
        if constexpr (POS) {
          qa[j] = __ldg(reinterpret_cast<const uint2 *>(pos + (long)row * ldpos + colA));
          qb[j] = __ldg(reinterpret_cast<const uint2 *>(pos + (long)row * ldpos + colB));
        }
      }
    }
#pragma unroll
    for (int j = 0; j < RB; ++j) {
      const int r_in = w * 16 + rb + j, row = row0 + r_in;
      const uint8_t *prow = stage + r_in * 128 + ((ck ^ (r_in & 7)) << 4);
      float4 va = *reinterpret_cast<const float4 *>(prow + pA * PANEL);
      float4 vb = *reinterpret_cast<const float4 *>(prow + (pA + 4) * PANEL);
      va.x += ra[j].x, va.y += ra[j].y, va.z += ra[j].z, va.w += ra[j].w;
      vb.x += rc[j].x, vb.y += rc[j].y, vb.z += rc[j].z, vb.w += rc[j].w;
      float sum = va.x + va.y + va.z + va.w + vb.x + vb.y + vb.z + vb.w;
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum * (1.f / 256.f);
      const float d0 = va.x - mean, d1 = va.y - mean, d2 = va.z - mean, d3 = va.w - mean;
      const float d4 = vb.x - mean, d5 = vb.y - mean, d6 = vb.z - mean, d7 = vb.w - mean;
      float q = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3 + d4 * d4 + d5 * d5 + d6 * d6 + d7 * d7;
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rstd = rsqrtf(q * (1.f / 256.f) + eps);
      const bool ok = row < M;
      const float4 yA = ok ? make_float4(d0 * rstd * gA.x + bA.x, d1 * rstd * gA.y + bA.y, d2 * rstd * gA.z + bA.z, d3 * rstd * gA.w + bA.w)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 yB = ok ? make_float4(d4 * rstd * gB.x + bB.x, d5 * rstd * gB.y + bB.y, d6 * rstd * gB.z + bB.z, d7 * rstd * gB.w + bB.w)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      emit(row, r_in, ok, colA, colB, yA, yB, qa[POS ? j : 0], qb[POS ? j : 0]);
    }
  }
}

__device__ __forceinline__ uint2 f32x4_to_bf16(const float4 &v) {
  uint2 u;
  *reinterpret_cast<__nv_bfloat162 *>(&u.x) = __floats2bfloat162_rn(v.x, v.y);
  *reinterpret_cast<__nv_bfloat162 *>(&u.y) = __floats2bfloat162_rn(v.z, v.w);
  return u;
}

template <typename TC>
__global__ void __launch_bounds__(320, 1)
mlp2_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
               const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmC,
               const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmW0,
               const float *__restrict__ b1, int M, int Hd, Epilogue ep, int tile0, LnOut ln, Front fr) {
  // tile0: first row tile of this launch (the tail tiles of a GEMM are launched separately with a hidden-dimension split)
  // gridDim.y > 1: split-K over the hidden dimension -- CTA (x, y) handles hidden chunks [y*NC, (y+1)*NC) of row tile x
  // and ADDS its partial product into the (zero-initialised, fp32) output with a TMA reduce-store; bias from split 0.
  using namespace mlp;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + OFF_BAR);
  uint64_t *x_full = bars, *full = bars + 1, *empty = full + NSLOT, *acc1_full = empty + NSLOT, *acc1_empty = acc1_full + 2,
           *h_full = acc1_empty + 2, *h_empty = h_full + 2, *acc2_full = h_empty + 2, *a_full = acc2_full + 1,
           *acc0_full = a_full + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc0_full + 1);

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
    mbar_init(x_full, fr.on ? 256 : 1);     // front mode: the 8 epilogue warps produce X instead of the TMA
    mbar_init(a_full, 1);
    mbar_init(acc0_full, 1);
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
      uint64_t *xbar = fr.on ? a_full : x_full;
      mbar_expect_tx(xbar, X_BYTES);
      for (int p = 0; p < XP; ++p) tma_load_2d(smem + p * PANEL, &tmX, xbar, p * BK, m_blk * BM);
      int t = 0;
      if (fr.on) {                          // W0: four k-blocks of 256 output rows x 64 columns, one ring slot each
        for (int kb = 0; kb < XP; ++kb, ++t) {
          const int s = t % NSLOT;
          mbar_wait(empty + s, ((t / NSLOT) & 1) ^ 1);
          mbar_expect_tx(full + s, SLOT);
          tma_load_2d(ring + s * SLOT, &tmW0, full + s, kb * BK, 0);
        }
      }
      if (fr.on && NC == 0) return;         // front-only launch (memotr_linear256_layernorm): no FFN behind the LayerNorm
      if (fr.on) mbar_wait(x_full, 0);      // the front LayerNorm stages the accumulator in the ring: hold W1 until it is done
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
      const uint32_t x_addr = smem_u32(smem), ring_addr = smem_u32(ring), h_addr = smem_u32(hbuf);
      int t = 0;
      if (fr.on) {                          // front GEMM: acc2 region = A . W0^T
        mbar_wait(a_full, 0);
        tcgen05_fence_after();
        for (int kb = 0; kb < XP; ++kb, ++t) {
          const int s = t % NSLOT;
          mbar_wait(full + s, (t / NSLOT) & 1);
          tcgen05_fence_after();
          const uint64_t adesc = umma_desc(x_addr + kb * PANEL), bdesc = umma_desc(ring_addr + s * SLOT);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(tmem_base + 2 * HC, adesc + 2 * k, bdesc + 2 * k, idesc2, (kb | k) != 0);
          umma_commit(empty + s);
        }
        umma_commit(acc0_full);
      }
      if (NC > 0) {
      mbar_wait(x_full, 0);
      tcgen05_fence_after();
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
    }
  } else {
    // ---- epilogue warps 2..9: lane quarter = warp % 4 (one accumulator row per thread), column half = (warp - 2) / 4 ----
    const int quarter = warp & 3, chalf = (warp - 2) >> 2;
    const int r_in = quarter * 32 + lane;
    long long *stamps = fr.stamps ? fr.stamps + 8 * (blockIdx.y * gridDim.x + blockIdx.x) : nullptr;
    auto stamp = [&](int i) {
      if (stamps && warp == 2 && lane == 0) stamps[i] = clock64();
    };
    stamp(0);
    if (fr.zptr) {
      const int et = threadIdx.x - 64;
      for (int idx = blockIdx.x * 256 + et; idx < fr.zrows * 64; idx += gridDim.x * 256)
        *reinterpret_cast<float4 *>(fr.zptr + (long)(idx >> 6) * fr.zld + (idx & 63) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    // accumulator (acc2 region) + bias -> fp32 staging tile in shared memory (this thread: row r_in, columns [chalf*128, +128))
    auto dump_acc = [&](uint8_t *stage, const float *bias) {
#pragma unroll 1
      for (int c0 = chalf * 128; c0 < chalf * 128 + 128; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + lane_off + (uint32_t)(2 * HC + c0), r);
        uint8_t *prow = stage + (c0 / 32) * PANEL + r_in * 128;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 bb = __ldg(reinterpret_cast<const float4 *>(bias + c0 + 4 * k));
          *reinterpret_cast<float4 *>(prow + ((k ^ (r_in & 7)) << 4)) =
              make_float4(__uint_as_float(r[4 * k]) + bb.x, __uint_as_float(r[4 * k + 1]) + bb.y,
                          __uint_as_float(r[4 * k + 2]) + bb.z, __uint_as_float(r[4 * k + 3]) + bb.w);
        }
      }
    };
    if (fr.on) {
      // ---- front LayerNorm: acc (A W0^T) + bias0 -> staging (the ring + the head of the H buffer: nothing streams yet) ->
      //      row-coalesced LayerNorm with the fp32 residual -> bf16 X tile in the swizzled K-major panels the TMA load would
      //      have produced (the attention tile there is dead once the front GEMM has committed) + fp32 copy to global ----
      mbar_wait(acc0_full, 0);
      tcgen05_fence_after();
      stamp(1);
      uint8_t *stage = ring;
      dump_acc(stage, fr.bias0);
      tcgen05_fence_before();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      stamp(2);
      ln_rows_256<false>(stage, warp - 2, lane, m_blk * BM, M, fr.res0, fr.ldres0, fr.gamma1, fr.beta1, fr.eps, nullptr, 0,
                  [&](int row, int rr, bool ok, int colA, int colB, const float4 &yA, const float4 &yB, uint2, uint2) {
                    if (ok && fr.x32 && blockIdx.y == 0) {
                      *reinterpret_cast<float4 *>(fr.x32 + (long)row * fr.ldx32 + colA) = yA;
                      *reinterpret_cast<float4 *>(fr.x32 + (long)row * fr.ldx32 + colB) = yB;
                    }
                    // X panel = 64 bf16 columns; 8-byte half of the 16-byte chunk (col % 64) / 8
                    uint8_t *xa = smem + (colA / 64) * PANEL + rr * 128 + ((((colA % 64) / 8) ^ (rr & 7)) << 4) + (colA % 8) * 2;
                    uint8_t *xb = smem + (colB / 64) * PANEL + rr * 128 + ((((colB % 64) / 8) ^ (rr & 7)) << 4) + (colB % 8) * 2;
                    *reinterpret_cast<uint2 *>(xa) = f32x4_to_bf16(yA);
                    *reinterpret_cast<uint2 *>(xb) = f32x4_to_bf16(yB);
                  });
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // st.shared -> visible to the UMMA (async proxy)
      mbar_arrive(x_full);                                           // releases the MMA warp (X ready) and the producer (ring free)
      stamp(3);
    }
    if (fr.on && NC == 0) {
      // front-only launch: the normalised bf16 tile (the four swizzled panels a TMA load would have produced) goes to global
      // memory as it is; x32 was written by the LayerNorm rows
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (warp == 2 && lane == 0) {
#pragma unroll 1
        for (int p = 0; p < XP; ++p)
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmC),
                       "r"(smem_u32(smem + p * PANEL)), "r"(p * BK), "r"(m_blk * BM)
                       : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (fr.stamps) fr.stamps[8 * blockIdx.x + 7] = clock64();
      }
    } else {
    for (int c = 0; c < NC; ++c) {
      const int b = c & 1;
      mbar_wait(acc1_full + b, (c >> 1) & 1);
      tcgen05_fence_after();
      if (c == 0) stamp(4);
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
    stamp(5);
    const int row = m_blk * BM + r_in;
    const bool row_ok = row < M;
    if (ln.res) {
      // ---- LayerNorm epilogue: acc2 + b2 -> staging (the dead X / ring memory) -> row-coalesced LayerNorm with the fp32
      //      residual -> y (bf16), y32 (fp32), y + pos (bf16), every global row written in 128-byte pieces ----
      if constexpr (sizeof(TC) == 2) {
        dump_acc(smem, ep.bias);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        stamp(6);
        __nv_bfloat16 *yp = reinterpret_cast<__nv_bfloat16 *>(ln.y), *qp = reinterpret_cast<__nv_bfloat16 *>(ln.ypos);
        const __nv_bfloat16 *posp = reinterpret_cast<const __nv_bfloat16 *>(ln.pos);
        auto emit = [&](int row, int rr, bool ok, int colA, int colB, const float4 &yA, const float4 &yB, uint2 pa, uint2 pb) {
          if (!ok) return;
          if (ln.y32) {
            *reinterpret_cast<float4 *>(ln.y32 + (long)row * ln.ld32 + colA) = yA;
            *reinterpret_cast<float4 *>(ln.y32 + (long)row * ln.ld32 + colB) = yB;
          }
          *reinterpret_cast<uint2 *>(yp + (long)row * ln.ldy + colA) = f32x4_to_bf16(yA);
          *reinterpret_cast<uint2 *>(yp + (long)row * ln.ldy + colB) = f32x4_to_bf16(yB);
          if (posp) {
            const float2 a0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&pa.x));
            const float2 a1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&pa.y));
            const float2 b0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&pb.x));
            const float2 b1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&pb.y));
            *reinterpret_cast<uint2 *>(qp + (long)row * ln.ldypos + colA) =
                f32x4_to_bf16(make_float4(yA.x + a0.x, yA.y + a0.y, yA.z + a1.x, yA.w + a1.y));
            *reinterpret_cast<uint2 *>(qp + (long)row * ln.ldypos + colB) =
                f32x4_to_bf16(make_float4(yB.x + b0.x, yB.y + b0.y, yB.z + b1.x, yB.w + b1.y));
          }
        };
        if (posp) ln_rows_256<true>(smem, warp - 2, lane, m_blk * BM, M, ln.res, ln.ldres, ln.gamma, ln.beta, ln.eps, posp, ln.ldpos, emit);
        else ln_rows_256<false>(smem, warp - 2, lane, m_blk * BM, M, ln.res, ln.ldres, ln.gamma, ln.beta, ln.eps, nullptr, 0, emit);
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
  }
  if (fr.stamps && NC > 0 && warp == 2 && lane == 0) fr.stamps[8 * (blockIdx.y * gridDim.x + blockIdx.x) + 7] = clock64();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

static long long *g_stamps = nullptr;   // memotr_mlp2_debug_stamps

template <typename TC>
static int launch_mlp2(const void *X, int ldx, const void *W1, const float *b1, const void *W2, void *C, int ldc, int M,
                       int Hd, const Epilogue &ep, cudaStream_t st, int nsplit = 1, int tile0 = 0, int ntiles = -1,
                       const LnOut &ln = LnOut{}, void *Q = nullptr, int ldq = 0, const Front &fr = Front{},
                       const void *W0 = nullptr, bool zeroed = false) {
  using namespace mlp;
  CUtensorMap tmX, tmW1, tmW2, tmC, tmQ, tmW0;
  const int Hm = Hd > 0 ? Hd : 256;      // (front-only launches carry no W1 / W2: the maps only have to be well-formed)
  if (!make_map(&tmX, X, M, K1, ldx, BM) || !make_map(&tmW1, W1, Hm, K1, K1, HC) || !make_map(&tmW2, W2, N2, Hm, Hm, 256) ||
      !make_map(&tmC, C, M, N2, ldc, BM, sizeof(TC) == 4) || !make_map(&tmQ, Q ? Q : C, M, N2, Q ? ldq : ldc, BM, sizeof(TC) == 4) ||
      !make_map(&tmW0, W0 ? W0 : W1, 256, K1, K1, 256))
    return fail(MEMOTR_ECUDA, "mlp2(tc): cuTensorMapEncodeTiled failed (M=%d Hd=%d)", M, Hd);
  auto kern = mlp2_tc_kernel<TC>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TOTAL);
    if (e != cudaSuccess) return fail(MEMOTR_ECUDA, "mlp2(tc): smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles = ntiles < 0 ? ceil_div(M, BM) - tile0 : ntiles;
  if (nsplit > 1 && !zeroed) {   // partial products are reduce-added: start from zero (a memset node in a graph)
    const int r0 = tile0 * BM, nr = (M - r0 < tiles * BM) ? M - r0 : tiles * BM;
    const cudaError_t e = cudaMemset2DAsync(reinterpret_cast<TC *>(C) + (size_t)r0 * ldc, (size_t)ldc * sizeof(TC), 0,
                                            (size_t)N2 * sizeof(TC), nr, st);
    if (e != cudaSuccess) return fail(MEMOTR_ECUDA, "mlp2(tc): memset: %s", cudaGetErrorString(e));
  }
  Front frs = fr;
  frs.stamps = g_stamps;
  if (nsplit > 1) frs.zptr = nullptr;
  MEMOTR_LAUNCH((kern), dim3(tiles, nsplit), 320, TOTAL, st, tmX, tmW1, tmW2, tmC, tmQ, tmW0, b1, M, Hd, ep, tile0, ln, frs);
  return check_launch("mlp2_tc");
}

}  // namespace tc
}  // namespace memotr

using namespace memotr;

// Profiling hook (tools/micro_dense.py): every later mlp2 launch writes 8 clock64 stamps per CTA to `buf` (device memory,
// 8 x CTAs int64; null switches it off): 0 start, 1 front GEMM done, 2 front tile staged, 3 front LayerNorm done, 4 first
// hidden chunk ready, 5 last GEMM2 done, 6 output tile staged, 7 end.
extern "C" int memotr_mlp2_debug_stamps(long long *buf) {
  tc::g_stamps = buf;
  return MEMOTR_OK;
}

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
  const int n_sm = sm_limit(sm_count()), tiles = ceil_div(M, tc::BM), chunks = Hd / tc::mlp::HC;
  const char *tl = getenv("MEMOTR_MLP_TAIL");
  if (tl && tl[0] == '0') return tc::launch_mlp2<float>(X, ldx, W1, b1, W2, C, ldc, M, Hd, ep, st);
  if (tiles * 2 <= n_sm) return tc::launch_mlp2<float>(X, ldx, W1, b1, W2, C, ldc, M, Hd, ep, st, pick_split(tiles, chunks, n_sm));
  const int tail = tiles - n_sm;
  if (tail > 0 && tail <= n_sm) {
    const int ns = pick_split(tail, chunks, n_sm);
    if (ns > 1) {   // the first round clears the rows the split round adds into (no memset node between the two launches)
      tc::Front z{};
      z.zptr = reinterpret_cast<float *>(C) + (size_t)n_sm * tc::BM * ldc, z.zrows = M - n_sm * tc::BM, z.zld = ldc;
      const int rc = tc::launch_mlp2<float>(X, ldx, W1, b1, W2, C, ldc, M, Hd, ep, st, 1, 0, n_sm, tc::LnOut{}, nullptr, 0, z);
      if (rc != MEMOTR_OK) return rc;
      return tc::launch_mlp2<float>(X, ldx, W1, b1, W2, C, ldc, M, Hd, ep, st, ns, n_sm, tail, tc::LnOut{}, nullptr, 0, tc::Front{},
                                    nullptr, true);
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
  const tc::LnOut ln{res, gamma, beta, y, y32, pos, ypos, ldres, ldy, ld32, ldpos, ldypos, eps};
  return tc::launch_mlp2<__nv_bfloat16>(X, ldx, W1, b1, W2, y, ldy, M, Hd, ep, (cudaStream_t)stream, 1, 0, -1, ln, ypos, ldypos);
}

extern "C" int memotr_layernorm(const void *x, int x_dtype, int ldx, const void *x2, int ldx2, const float *gamma,
                                const float *beta, float eps, void *y, int y_dtype, int ldy, const void *pos, int ldpos,
                                void *ypos, int ldypos, float *y32, int ld32, int M, int C, void *stream);

// y = LayerNorm(res + A W^T + b) * gamma + beta for a 256 x 256 projection: y (bf16) and y32 (fp32) -- the encoder's
// `src = norm1(src + output_proj(attn))` (deformable_encoder.py:124-126, ms_deform_attn.py:129) as ONE kernel per 128-row tile:
// the product never leaves the SM (TMEM -> shared-memory staging tile -> row-coalesced LayerNorm), instead of a GEMM writing
// fp32 rows that a LayerNorm kernel reads back (114 -> 69 MB of HBM traffic per encoder layer at the DanceTrack size).
extern "C" int memotr_linear256_layernorm(const void *A, int lda, const void *W, const float *bias, const float *res, int ldres,
                                          const float *gamma, const float *beta, float eps, void *y, int ldy, float *y32, int ld32,
                                          int M, void *stream) {
  MEMOTR_REQUIRE(M >= 0 && A && W && bias && res && gamma && beta && y && y32, "linear256_layernorm: null pointer");
  MEMOTR_REQUIRE(lda % 8 == 0 && ldy % 8 == 0 && ldres % 4 == 0 && ld32 % 4 == 0 && aligned16(A) && aligned16(W) && aligned16(bias) &&
                     aligned16(res) && aligned16(gamma) && aligned16(beta) && aligned16(y) && aligned16(y32),
                 "linear256_layernorm: misaligned buffer");
  MEMOTR_REQUIRE(tc::encode_fn() != nullptr, "linear256_layernorm: cuTensorMapEncodeTiled unavailable");
  if (M == 0) return MEMOTR_OK;
  const tc::Front fr{bias, res, gamma, beta, y32, ldres, ld32, 1, eps, nullptr, nullptr, 0, 0};
  Epilogue ep{nullptr, nullptr, nullptr, nullptr, 0, 0, ACT_NONE};
  return tc::launch_mlp2<__nv_bfloat16>(A, lda, W, bias, W, y, ldy, M, 0, ep, (cudaStream_t)stream, 1, 0, -1, tc::LnOut{}, nullptr, 0, fr, W);
}

// The dense half of an encoder layer in one kernel per 128-row tile (deformable_encoder.py:124-131, ms_deform_attn.py:129):
//   x   = LayerNorm1(src32 + att Wout^T + bout)                          (front GEMM + LayerNorm, X never leaves the SM)
//   y   = LayerNorm2(x + relu(x W1^T + b1) W2^T + b2)                    (fused FFN + LayerNorm epilogue)
// outputs: x32 (fp32 x: the residual of norm2), y (bf16), y32 (fp32), ypos = y + pos (bf16).  Rows beyond the first round of
// SMs run the same front with the hidden dimension split over the idle SMs into `pre` (fp32 scratch, M x 256) and get their
// LayerNorm2 from the stand-alone kernel.
extern "C" int memotr_encoder_dense_block(const void *att, int ldatt, const void *Wout, const float *bout, const float *src32,
                                          int ldsrc, const float *gamma1, const float *beta1, float *x32, int ldx32,
                                          const void *W1, const float *b1, const void *W2, const float *b2, const float *gamma2,
                                          const float *beta2, const void *pos, int ldpos, void *y, int ldy, float *y32, int ld32,
                                          void *ypos, int ldypos, float *pre, int ldpre, int M, int Hd, float eps, void *stream) {
  MEMOTR_REQUIRE(M >= 0 && att && Wout && bout && src32 && gamma1 && beta1 && x32 && W1 && b1 && W2 && b2 && gamma2 && beta2 && pos &&
                     y && y32 && ypos && pre,
                 "encoder_dense_block: null pointer");
  MEMOTR_REQUIRE(Hd > 0 && Hd % tc::mlp::HC == 0, "encoder_dense_block: hidden %% 128 != 0 (got %d)", Hd);
  MEMOTR_REQUIRE(ldatt % 8 == 0 && ldy % 8 == 0 && ldpos % 8 == 0 && ldypos % 8 == 0 && ldsrc % 4 == 0 && ldx32 % 4 == 0 && ld32 % 4 == 0 &&
                     ldpre % 4 == 0 && aligned16(att) && aligned16(Wout) && aligned16(bout) && aligned16(src32) && aligned16(gamma1) &&
                     aligned16(beta1) && aligned16(x32) && aligned16(W1) && aligned16(b1) && aligned16(W2) && aligned16(b2) &&
                     aligned16(gamma2) && aligned16(beta2) && aligned16(pos) && aligned16(y) && aligned16(y32) && aligned16(ypos) && aligned16(pre),
                 "encoder_dense_block: misaligned buffer");
  MEMOTR_REQUIRE(tc::encode_fn() != nullptr, "encoder_dense_block: cuTensorMapEncodeTiled unavailable");
  if (M == 0) return MEMOTR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int n_sm = sm_count(), tiles = ceil_div(M, tc::BM), chunks = Hd / tc::mlp::HC;
  const tc::Front fr{bout, src32, gamma1, beta1, x32, ldsrc, ldx32, 1, eps, nullptr, nullptr, 0, 0};
  const tc::LnOut ln{x32, gamma2, beta2, y, y32, pos, ypos, ldx32, ldy, ld32, ldpos, ldypos, eps};
  Epilogue ep{b2, nullptr, nullptr, nullptr, 0, 0, ACT_NONE};
  const int main_tiles = tiles <= n_sm ? tiles : n_sm;
  const int tail = tiles - main_tiles, r0 = main_tiles * tc::BM;
  const int ns = tail > 0 ? pick_split(tail, chunks, n_sm) : 1;
  tc::Front frm = fr;
  if (ns > 1) frm.zptr = pre + (size_t)r0 * ldpre, frm.zrows = M - r0, frm.zld = ldpre;   // cleared under the main launch
  int rc = tc::launch_mlp2<__nv_bfloat16>(att, ldatt, W1, b1, W2, y, ldy, M, Hd, ep, st, 1, 0, main_tiles, ln, ypos, ldypos, frm, Wout);
  if (rc != MEMOTR_OK || main_tiles == tiles) return rc;
  rc = tc::launch_mlp2<float>(att, ldatt, W1, b1, W2, pre, ldpre, M, Hd, ep, st, ns, main_tiles, tail, tc::LnOut{}, nullptr, 0, fr, Wout,
                              ns > 1);
  if (rc != MEMOTR_OK) return rc;
  return memotr_layernorm(pre + (size_t)r0 * ldpre, MEMOTR_F32, ldpre, x32 + (size_t)r0 * ldx32, ldx32, gamma2, beta2, eps,
                          reinterpret_cast<__nv_bfloat16 *>(y) + (size_t)r0 * ldy, MEMOTR_BF16, ldy,
                          reinterpret_cast<const __nv_bfloat16 *>(pos) + (size_t)r0 * ldpos, ldpos,
                          reinterpret_cast<__nv_bfloat16 *>(ypos) + (size_t)r0 * ldypos, ldypos, y32 + (size_t)r0 * ld32, ld32, M - r0,
                          256, stream);
}
