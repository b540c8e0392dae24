// decoder_fused.cu -- the whole DeformableDecoder (all layers) + box / class heads as ONE persistent kernel (bf16 engine).
//
// Replaces, for the bf16 engine, the ~150 launches per frame that DeformableDecoder.forward / DeformableDecoderLayer.forward
// (models/deformable_decoder.py:56-160, 276-319) and the per-layer heads (models/memotr.py:147-162) became in the
// launch-per-op engine: with <= 400 query rows every one of those kernels is launch-latency-bound (3-5 us each, 733 us
// per frame, profiles/r01_launches_bench_steps2_v2_warm.csv) although the arithmetic is ~0.4 GFLOP per layer.
//
// Decomposition: a CTA owns a block of 16 query rows for the WHOLE decoder.  Everything in a decoder layer is row-local
// except the self-attention, which needs the keys / values of all queries: so a layer is
//     sine embed -> ref_point_head -> query_scale -> q/k/v projections   | K, V to global, ONE grid barrier |
//     attention (warp = head) -> out_proj + LN -> offsets/logits -> softmax + bilinear gather from the value map ->
//     out_proj + LN -> FFN + LN -> box head + refinement + class head
// and the activations of the 16 rows never leave shared memory between layers.  Dense layers run on mma.sync.m16n8k16
// (M = 16 is far below the tcgen05 minimum tile): the A operand is the row block in shared memory, the weights stream
// from L2 through a 3-slot x 33 KB ring, one cp.async.bulk per slot: the host packs every weight matrix as a sequence of
// slot images (64 output rows x 256 k, rows padded to 528 B so that ldmatrix is bank-conflict-free) in exactly the order
// the kernel consumes them (per-row 512-byte bulk copies were tried first: 64 requests per slot made the copy engine the
// bottleneck, 275 us per layer).  The weight
// stream is the roofline of this kernel: 3.8 MB per layer per CTA at the ~140-200 GB/s one SM sustains (tools/tma_stream.cu);
// the producer runs ahead across op and layer boundaries (the program is static), so the stream never waits for the
// epilogues, the attention or the grid barrier.  Attention: q, k in fp16 (11-bit mantissa; bf16 logits were measurably
// too coarse), p, v in fp16, fp32 accumulate, online softmax, K rows / V^T rows fetched from L2 as 16-byte fragments with
// a permuted k index so that no shared-memory staging is needed.
//
// Numerics are those of the bf16 engine (bf16 GEMM operands, fp32 accumulate / residual / LayerNorm / geometry); the
// fp32 engine keeps the launch-per-op path (bit-exactness tests live there).
#include "decoder_common.cuh"

namespace memotr {
namespace dec {

constexpr int OFF_X32 = 0, OFF_XB = OFF_X32 + R * C * 4, OFF_QP = OFF_XB + R * P256, OFF_A = OFF_QP + R * P256,
              OFF_B = OFF_A + R * P512, OFF_H = OFF_B + R * P256, OFF_F0 = OFF_H + R * P1024,
              OFF_RING = OFF_F0 + R * F0P * 4, OFF_MISC = OFF_RING + NSLOT * SLOT_BYTES, OFF_PROG = OFF_MISC + 512,
              MAX_PROG = 15 * MEMOTR_DEC_MAX_LAYERS, SMEM_TOTAL = OFF_PROG + MAX_PROG * 24;
static_assert(OFF_RING % 16 == 0 && SMEM_TOTAL + 128 <= 227 * 1024 && sizeof(memotr_dec_gemm) == 24, "shared memory plan");

__global__ void __launch_bounds__(NTHREADS, 1) decoder_fused_kernel(const __grid_constant__ memotr_dec_params P) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  float *x32 = reinterpret_cast<float *>(smem + OFF_X32);
  uint8_t *xb = smem + OFF_XB, *qp = smem + OFF_QP, *bufA = smem + OFF_A, *bufB = smem + OFF_B, *hbuf = smem + OFF_H;
  float *f0 = reinterpret_cast<float *>(smem + OFF_F0);
  float *refs = reinterpret_cast<float *>(smem + OFF_MISC);             // [16][4] current reference boxes (sigmoid space)
  float *delta = refs + 64;                                             // [16][4]
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + OFF_MISC + 384), *empty = full + NSLOT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int row0 = blockIdx.x * R;
  const int nq = P.nq, nd = P.nd;

  memotr_dec_gemm *sprog = reinterpret_cast<memotr_dec_gemm *>(smem + OFF_PROG);   // the weight program, read many times
  for (int i = tid; i < P.n_prog * 6; i += NTHREADS)
    reinterpret_cast<uint32_t *>(sprog)[i] = reinterpret_cast<const uint32_t *>(P.prog)[i];
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) mbar_init(full + s, 1), mbar_init(empty + s, NCW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == NCW) {
    // ------------------------------------------------------------------ producer: stream the weight program through the ring
    uint32_t t = 0;
    for (int gi = 0; gi < P.n_prog; ++gi) {
      const memotr_dec_gemm d = sprog[gi];
      const uint8_t *W = reinterpret_cast<const uint8_t *>(d.W);   // pre-packed slot images, in consumption order
      const int nslots = (d.N / SLOT_ROWS) * (d.K / SLOT_K);
      for (int i = 0; i < nslots; ++i, ++t) {
        const int s = t % NSLOT;
        if (lane == 0) {
          mbar_wait(empty + s, ((t / NSLOT) & 1) ^ 1);
          mbar_expect_tx(full + s, SLOT_BYTES);
          bulk_row(smem + OFF_RING + s * SLOT_BYTES, W + (long)i * SLOT_BYTES, SLOT_BYTES, full + s);
        }
      }
    }
    return;
  }

  // ---------------------------------------------------------------------- consumers (8 warps, 256 threads)
  Ring rg{smem + OFF_RING, full, empty, 0u, 0};
  const int g = lane >> 2, c = lane & 3;
  // layer input: rows of tgt (fp32) and the reference boxes
  for (int i = tid; i < R * C / 4; i += 256) {
    const int r = i / (C / 4), c4 = (i % (C / 4)) * 4;
    const int row = min(row0 + r, nq - 1);
    const float4 v = *reinterpret_cast<const float4 *>(P.tgt_in + (long)row * C + c4);
    *reinterpret_cast<float4 *>(x32 + r * C + c4) = v;
    *reinterpret_cast<uint2 *>(xb + r * P256 + c4 * 2) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
  if (tid < R * 4) refs[tid] = P.ref_in[(long)min(row0 + tid / 4, nq - 1) * 4 + (tid & 3)];
  csync();

  // optional phase timestamps (tools/prof_decoder.py): P.prof[(block * n_layers + layer) * 16 + k] = clock64 at boundary k
#define STAMP(k)                                                                                        \
  if (P.prof && tid == 0) P.prof[((long)blockIdx.x * P.n_layers + lid) * 16 + (k)] = clock64();
  for (int lid = 0; lid < P.n_layers; ++lid) {
    const memotr_dec_layer &Lp = P.layers[lid];
    STAMP(0)
    const int n = lid >= P.merge ? nq : nd;               // rows taking part in this layer (deformable_decoder.py:292-297)
    const int par = lid & 1;
    __half *Kh = reinterpret_cast<__half *>(P.kbuf) + (long)par * P.np * C;     // [np][256]
    __half *Vt = reinterpret_cast<__half *>(P.vbuf) + (long)par * C * P.np;     // [256][np]
    // init_ref_pts / last_ref_pts of the output dict (memotr.py:183-187): inverse_sigmoid of the references that enter the
    // first / the last layer
    if (tid < R * 4 && row0 + (tid >> 2) < nq) {
      const long o = (long)(row0 + (tid >> 2)) * 4 + (tid & 3);
      if (lid == 0 && P.init_ref_out) P.init_ref_out[o] = inv_sigm(refs[tid]);
      if (lid == P.n_layers - 1 && P.last_ref_out) P.last_ref_out[o] = inv_sigm(refs[tid]);
    }

    // ---- DAB positional query (deformable_decoder.py:82-95): sine embedding of ref * valid_ratio(level 0)
    {
      // e = p * 2pi / dim_t with the correctly rounded reciprocal; |e| <= 2pi, where __sinf / __cosf are good to ~1e-6
      // absolute -- the result is rounded to bf16 (4e-3) anyway; sinf / cosf cost ~100 instructions each on 8 warps
      const float vr0x = __ldg(P.valid_ratios), vr0y = __ldg(P.valid_ratios + 1);   // level-0 ratios (deformable_decoder.py:82-91)
      const float scl[4] = {vr0x, vr0y, vr0x, vr0y};
      for (int i = tid; i < R * 256; i += 256) {
        const int r = i >> 8, cc = (i >> 6) & 3, j = i & 63;
        const float e = refs[r * 4 + cc] * scl[cc] * 6.283185307179586f * __frcp_rn(__ldg(P.dim_t + 2 * j));
        *reinterpret_cast<uint32_t *>(bufA + r * P512 + (cc * 128 + 2 * j) * 2) = pack_bf16(__sinf(e), __cosf(e));
      }
    }
    csync();
    gemm(sprog, rg, bufA, P512, P.rph0_b, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {          // ref_point_head.0 + ReLU
      *reinterpret_cast<uint32_t *>(bufB + g * P256 + col * 2) = pack_bf16(fmaxf(a[0] + b0, 0.f), fmaxf(a[1] + b1, 0.f));
      *reinterpret_cast<uint32_t *>(bufB + (g + 8) * P256 + col * 2) = pack_bf16(fmaxf(a[2] + b0, 0.f), fmaxf(a[3] + b1, 0.f));
    });
    csync();
    if (lid == 0) {
      gemm(sprog, rg, bufB, P256, P.rph1_b, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {       // ref_point_head.1 -> query_pos
        *reinterpret_cast<uint32_t *>(qp + g * P256 + col * 2) = pack_bf16(a[0] + b0, a[1] + b1);
        *reinterpret_cast<uint32_t *>(qp + (g + 8) * P256 + col * 2) = pack_bf16(a[2] + b0, a[3] + b1);
      });
      csync();
    } else {
      gemm(sprog, rg, bufB, P256, P.rph1_b, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {       // raw query pos -> bufA (bf16)
        *reinterpret_cast<uint32_t *>(bufA + g * P512 + col * 2) = pack_bf16(a[0] + b0, a[1] + b1);
        *reinterpret_cast<uint32_t *>(bufA + (g + 8) * P512 + col * 2) = pack_bf16(a[2] + b0, a[3] + b1);
      });
      gemm(sprog, rg, xb, P256, P.qs0_b, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {         // query_scale.0 + ReLU
        *reinterpret_cast<uint32_t *>(bufB + g * P256 + col * 2) = pack_bf16(fmaxf(a[0] + b0, 0.f), fmaxf(a[1] + b1, 0.f));
        *reinterpret_cast<uint32_t *>(bufB + (g + 8) * P256 + col * 2) = pack_bf16(fmaxf(a[2] + b0, 0.f), fmaxf(a[3] + b1, 0.f));
      });
      csync();
      gemm(sprog, rg, bufB, P256, P.qs1_b, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {       // query_scale.1 * raw query pos
        const float2 m0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(bufA + g * P512 + col * 2));
        const float2 m1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(bufA + (g + 8) * P512 + col * 2));
        *reinterpret_cast<uint32_t *>(qp + g * P256 + col * 2) = pack_bf16((a[0] + b0) * m0.x, (a[1] + b1) * m0.y);
        *reinterpret_cast<uint32_t *>(qp + (g + 8) * P256 + col * 2) = pack_bf16((a[2] + b0) * m1.x, (a[3] + b1) * m1.y);
      });
      csync();
    }

    STAMP(1)
    // ---- self-attention projections (deformable_decoder.py:245-247): q = k = tgt + query_pos, v = tgt
    for (int i = tid; i < R * C / 8; i += 256) {
      const int r = i / (C / 8), c8 = (i % (C / 8)) * 8;
      float a[8], b[8];
      bf16x8_to_f32(*reinterpret_cast<const uint4 *>(xb + r * P256 + c8 * 2), a);
      bf16x8_to_f32(*reinterpret_cast<const uint4 *>(qp + r * P256 + c8 * 2), b);
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] += b[k];
      *reinterpret_cast<uint4 *>(bufA + r * P512 + c8 * 2) = f32x8_to_bf16(a);
    }
    csync();
    gemm(sprog, rg, bufA, P512, Lp.qk_b, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {         // [q | k] (fp16)
      if (col < C) {
        const float sc = 0.17677669529663687f;                                            // 1 / sqrt(32), folded into q
        *reinterpret_cast<uint32_t *>(bufB + g * P256 + col * 2) = pack_f16((a[0] + b0) * sc, (a[1] + b1) * sc);
        *reinterpret_cast<uint32_t *>(bufB + (g + 8) * P256 + col * 2) = pack_f16((a[2] + b0) * sc, (a[3] + b1) * sc);
      } else {
        if (row0 + g < nq) *reinterpret_cast<uint32_t *>(Kh + (long)(row0 + g) * C + col - C) = pack_f16(a[0] + b0, a[1] + b1);
        if (row0 + g + 8 < nq)
          *reinterpret_cast<uint32_t *>(Kh + (long)(row0 + g + 8) * C + col - C) = pack_f16(a[2] + b0, a[3] + b1);
      }
    });
    // (V of a padded key is written as zero: its softmax weight is exactly 0, but 0 x a stale non-finite row would be NaN)
    const bool keep0 = !(row0 + g < nq && P.query_pad && P.query_pad[row0 + g]);
    const bool keep1 = !(row0 + g + 8 < nq && P.query_pad && P.query_pad[row0 + g + 8]);
    gemm(sprog, rg, xb, P256, Lp.v_b, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {           // v, stored transposed (fp16)
      if (row0 + g < nq) {
        Vt[(long)col * P.np + row0 + g] = __float2half_rn(keep0 ? a[0] + b0 : 0.f);
        Vt[(long)(col + 1) * P.np + row0 + g] = __float2half_rn(keep0 ? a[1] + b1 : 0.f);
      }
      if (row0 + g + 8 < nq) {
        Vt[(long)col * P.np + row0 + g + 8] = __float2half_rn(keep1 ? a[2] + b0 : 0.f);
        Vt[(long)(col + 1) * P.np + row0 + g + 8] = __float2half_rn(keep1 ? a[3] + b1 : 0.f);
      }
    });
    STAMP(2)
    grid_barrier(P.barrier, (unsigned int)(lid + 1) * gridDim.x);
    STAMP(3)

    // ---- attention: warp = head, 16 queries x n keys, online softmax over blocks of 64 keys
    {
      const int h = warp;
      const uint4 qv0 = *reinterpret_cast<const uint4 *>(bufB + g * P256 + (h * 32 + 8 * c) * 2);
      const uint4 qv1 = *reinterpret_cast<const uint4 *>(bufB + (g + 8) * P256 + (h * 32 + 8 * c) * 2);
      float o[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
      float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
      const int nblk = (n + 63) / 64;
      for (int blk = 0; blk < nblk; ++blk) {
        const int kb = blk * 64;
        uint4 kr[8], vr[4][2];
#pragma unroll
        for (int j = 0; j < 8; ++j) {   // S tile j, column n = g  <->  key kb + 16*(g/2) + 4*(j/2) + 2*(j%2) + g%2
          const int key = kb + 16 * (g >> 1) + 4 * (j >> 1) + 2 * (j & 1) + (g & 1);
          kr[j] = __ldcg(reinterpret_cast<const uint4 *>(Kh + (long)key * C + h * 32 + 8 * c));   // L2 only: written by peers
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {   // O tile i: d = 8i + g; this thread's 16 keys kb + 16c .. +15
          const __half *vp = Vt + (long)(h * 32 + 8 * i + g) * P.np + kb + 16 * c;
          vr[i][0] = __ldcg(reinterpret_cast<const uint4 *>(vp));
          vr[i][1] = __ldcg(reinterpret_cast<const uint4 *>(vp + 8));
        }
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
          mma_f16(s[j], qv0.x, qv1.x, qv0.y, qv1.y, kr[j].x, kr[j].y);
          mma_f16(s[j], qv0.z, qv1.z, qv0.w, qv1.w, kr[j].z, kr[j].w);
        }
        // this thread's columns of tile j: n = 2c + e  <->  key kb + 16c + 4*(j/2) + 2*(j%2) + e
        float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int key = kb + 16 * c + 4 * (j >> 1) + 2 * (j & 1) + e;
            const bool dead = key >= n || (P.query_pad && P.query_pad[key]);
            if (dead) s[j][e] = -INFINITY, s[j][2 + e] = -INFINITY;
            bm0 = fmaxf(bm0, s[j][e]), bm1 = fmaxf(bm1, s[j][2 + e]);
          }
        bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)), bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
        bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)), bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
        const float n0 = fmaxf(m0, bm0), n1 = fmaxf(m1, bm1);
        const float u0 = n0 == -INFINITY ? 0.f : n0, u1 = n1 == -INFINITY ? 0.f : n1;   // all keys so far padded
        const float f0s = __expf(m0 - u0), f1s = __expf(m1 - u1);                       // exp(-inf) = 0 on the first block
        m0 = n0, m1 = n1;
        l0 *= f0s, l1 *= f1s;
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i][0] *= f0s, o[i][1] *= f0s, o[i][2] *= f1s, o[i][3] *= f1s;
        uint32_t pa[8][2];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float p0 = __expf(s[j][0] - u0), p1 = __expf(s[j][1] - u0), p2 = __expf(s[j][2] - u1), p3 = __expf(s[j][3] - u1);
          l0 += p0 + p1, l1 += p2 + p3;
          pa[j][0] = pack_f16(p0, p1), pa[j][1] = pack_f16(p2, p3);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t vv[8] = {vr[i][0].x, vr[i][0].y, vr[i][0].z, vr[i][0].w, vr[i][1].x, vr[i][1].y, vr[i][1].z, vr[i][1].w};
#pragma unroll
          for (int st = 0; st < 4; ++st)   // k16 step st: P tiles 2st, 2st+1; keys 16c + 4st + {0,1} and + {2,3}
            mma_f16(o[i], pa[2 * st][0], pa[2 * st][1], pa[2 * st + 1][0], pa[2 * st + 1][1], vv[2 * st], vv[2 * st + 1]);
        }
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1), l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1), l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int col = h * 32 + 8 * i + 2 * c;
        *reinterpret_cast<uint32_t *>(bufA + g * P512 + col * 2) = pack_bf16(o[i][0] * i0, o[i][1] * i0);
        *reinterpret_cast<uint32_t *>(bufA + (g + 8) * P512 + col * 2) = pack_bf16(o[i][2] * i1, o[i][3] * i1);
      }
    }
    csync();
    STAMP(4)
    gemm(sprog, rg, bufA, P512, Lp.sao_b, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {         // out_proj + residual
      *reinterpret_cast<float2 *>(f0 + g * F0P + col) = make_float2(a[0] + b0 + x32[g * C + col], a[1] + b1 + x32[g * C + col + 1]);
      *reinterpret_cast<float2 *>(f0 + (g + 8) * F0P + col) =
          make_float2(a[2] + b0 + x32[(g + 8) * C + col], a[3] + b1 + x32[(g + 8) * C + col + 1]);
    });
    csync();
    layer_norm(f0, Lp.n2_g, Lp.n2_b, x32, xb, qp, bufA, warp, lane);                      // norm2; bufA = t1 + query_pos
    csync();

    STAMP(5)
    // ---- cross-attention into the encoder memory (ms_deform_attn.py:88-130)
    gemm(sprog, rg, bufA, P512, Lp.ol_b, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {         // [offsets | logits], fp32
      *reinterpret_cast<float2 *>(f0 + g * F0P + col) = make_float2(a[0] + b0, a[1] + b1);
      *reinterpret_cast<float2 *>(f0 + (g + 8) * F0P + col) = make_float2(a[2] + b0, a[3] + b1);
    });
    csync();
    STAMP(6)
    {
      const int Kp = P.n_points, Lv = P.n_levels, LK = Lv * Kp;
      const __half *value = reinterpret_cast<const __half *>(Lp.value);
      const int xs = P.value_stride;
      for (int pass = 0; pass < 2; ++pass) {
        const int pair = pass * 64 + (tid >> 2), sub = tid & 3;
        const int r = pair >> 3, h = pair & 7;
        const float *rowp = f0 + r * F0P;
        float mx = -INFINITY;
        for (int i = 0; i < LK; ++i) mx = fmaxf(mx, rowp[2 * 8 * LK + h * LK + i]);
        float sum = 0.f;
        for (int i = 0; i < LK; ++i) sum += __expf(rowp[2 * 8 * LK + h * LK + i] - mx);
        const float rs = __frcp_rn(sum), rk = __frcp_rn((float)Kp);
        const float rx = refs[r * 4], ry = refs[r * 4 + 1], rw = refs[r * 4 + 2], rh = refs[r * 4 + 3];
        const __half *vb = value + h * 32 + sub * 8;
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        for (int l = 0; l < Lv; ++l) {
          const int Hh = P.shapes[2 * l], Ww = P.shapes[2 * l + 1];
          const float Hf = (float)Hh, Wf = (float)Ww;
          const float vx = __ldg(P.valid_ratios + 2 * l), vy = __ldg(P.valid_ratios + 2 * l + 1);
          const long base = (long)P.lsi[l] * xs;
          const int ys = Ww * xs;
          __half2 a2[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) a2[j] = __float2half2_rn(0.f);
          for (int pb = 0; pb < Kp; pb += 4) {       // 4 points = 16 corner rows in flight per thread
            uint4 rv[4][4];
            __half2 wq[4][4];
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
              const int p = min(pb + pp, Kp - 1), i = l * Kp + p;
              const bool pv = pb + pp < Kp;
              const float2 off = *reinterpret_cast<const float2 *>(rowp + (h * LK + i) * 2);
              const float aw = pv ? __expf(rowp[2 * 8 * LK + h * LK + i] - mx) * rs : 0.f;
              const float lx = rx * vx + off.x * rk * (rw * vx) * 0.5f, ly = ry * vy + off.y * rk * (rh * vy) * 0.5f;
              const float h_im = __fmaf_rn(ly, Hf, -0.5f), w_im = __fmaf_rn(lx, Wf, -0.5f);
              const bool inside = h_im > -1.f && w_im > -1.f && h_im < Hf && w_im < Wf;
              const float hfl = floorf(h_im), wfl = floorf(w_im);
              const int y0 = (int)hfl, x0 = (int)wfl;
              const float lh = h_im - hfl, lw = w_im - wfl, hh = 1.f - lh, hw = 1.f - lw;
              const bool y0ok = inside && y0 >= 0, y1ok = inside && y0 + 1 <= Hh - 1, x0ok = x0 >= 0, x1ok = x0 + 1 <= Ww - 1;
              const int yc0 = min(max(y0, 0), Hh - 1), yc1 = min(max(y0 + 1, 0), Hh - 1);
              const int xc0 = min(max(x0, 0), Ww - 1), xc1 = min(max(x0 + 1, 0), Ww - 1);
              rv[pp][0] = __ldg(reinterpret_cast<const uint4 *>(vb + base + (long)yc0 * ys + xc0 * xs));
              rv[pp][1] = __ldg(reinterpret_cast<const uint4 *>(vb + base + (long)yc0 * ys + xc1 * xs));
              rv[pp][2] = __ldg(reinterpret_cast<const uint4 *>(vb + base + (long)yc1 * ys + xc0 * xs));
              rv[pp][3] = __ldg(reinterpret_cast<const uint4 *>(vb + base + (long)yc1 * ys + xc1 * xs));
              wq[pp][0] = __float2half2_rn((y0ok && x0ok) ? hh * hw * aw : 0.f);
              wq[pp][1] = __float2half2_rn((y0ok && x1ok) ? hh * lw * aw : 0.f);
              wq[pp][2] = __float2half2_rn((y1ok && x0ok) ? lh * hw * aw : 0.f);
              wq[pp][3] = __float2half2_rn((y1ok && x1ok) ? lh * lw * aw : 0.f);
            }
#pragma unroll
            for (int pp = 0; pp < 4; ++pp)
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                const __half2 *v2 = reinterpret_cast<const __half2 *>(&rv[pp][q4]);
#pragma unroll
                for (int j = 0; j < 4; ++j) a2[j] = __hfma2(wq[pp][q4], v2[j], a2[j]);
              }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(a2[j]);
            acc[2 * j] += f.x, acc[2 * j + 1] += f.y;
          }
        }
        *reinterpret_cast<uint4 *>(bufA + r * P512 + (h * 32 + sub * 8) * 2) = f32x8_to_bf16(acc);
      }
    }
    csync();
    STAMP(7)
    gemm(sprog, rg, bufA, P512, Lp.cao_b, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {         // output_proj + residual
      *reinterpret_cast<float2 *>(f0 + g * F0P + col) = make_float2(a[0] + b0 + x32[g * C + col], a[1] + b1 + x32[g * C + col + 1]);
      *reinterpret_cast<float2 *>(f0 + (g + 8) * F0P + col) =
          make_float2(a[2] + b0 + x32[(g + 8) * C + col], a[3] + b1 + x32[(g + 8) * C + col + 1]);
    });
    csync();
    layer_norm(f0, Lp.n1_g, Lp.n1_b, x32, xb, nullptr, nullptr, warp, lane);              // norm1
    csync();

    STAMP(8)
    // ---- FFN (deformable_decoder.py:263-273) in two halves of the hidden dimension
    const int n_half = P.d_ffn > 1024 ? 2 : 1;            // the hidden row block (16 x 1024 bf16) holds half of d_ffn = 2048
    for (int half = 0; half < n_half; ++half) {
      const int hoff = half * (P.d_ffn / n_half);
      gemm(sprog, rg, xb, P256, Lp.f1_b + hoff, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {         // linear1 + ReLU -> hidden (bf16)
        *reinterpret_cast<uint32_t *>(hbuf + g * P1024 + col * 2) = pack_bf16(fmaxf(a[0] + b0, 0.f), fmaxf(a[1] + b1, 0.f));
        *reinterpret_cast<uint32_t *>(hbuf + (g + 8) * P1024 + col * 2) = pack_bf16(fmaxf(a[2] + b0, 0.f), fmaxf(a[3] + b1, 0.f));
      });
      csync();
      gemm(sprog, rg, hbuf, P1024, half == n_half - 1 ? Lp.f2_b : nullptr, warp, lane,
           [&](int col, const float (&a)[4], float b0, float b1) {                          // linear2 (accumulated) + residual
        float2 *d0 = reinterpret_cast<float2 *>(f0 + g * F0P + col), *d1 = reinterpret_cast<float2 *>(f0 + (g + 8) * F0P + col);
        float2 p0 = make_float2(a[0], a[1]), p1 = make_float2(a[2], a[3]);
        if (half > 0) p0.x += d0->x, p0.y += d0->y, p1.x += d1->x, p1.y += d1->y;
        if (half == n_half - 1) {
          p0.x += b0 + x32[g * C + col], p0.y += b1 + x32[g * C + col + 1];
          p1.x += b0 + x32[(g + 8) * C + col], p1.y += b1 + x32[(g + 8) * C + col + 1];
        }
        *d0 = p0, *d1 = p1;
      });
      csync();
    }
    layer_norm(f0, Lp.n3_g, Lp.n3_b, x32, xb, nullptr, nullptr, warp, lane);              // norm3 -> the layer output
    csync();
    STAMP(9)
    // rows that do not take part in this layer pass through unchanged (:316-317); write the layer output
    {
      const float *prev = lid == 0 ? P.tgt_in : P.layers[lid - 1].tgt_out;
      for (int i = tid; i < R * C / 4; i += 256) {
        const int r = i / (C / 4), c4 = (i % (C / 4)) * 4, row = row0 + r;
        if (row >= nq) continue;
        float4 v = *reinterpret_cast<const float4 *>(x32 + r * C + c4);
        if (row >= n) {
          v = *reinterpret_cast<const float4 *>(prev + (long)row * C + c4);
          *reinterpret_cast<float4 *>(x32 + r * C + c4) = v;
          *reinterpret_cast<uint2 *>(xb + r * P256 + c4 * 2) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
        }
        *reinterpret_cast<float4 *>(Lp.tgt_out + (long)row * C + c4) = v;
      }
    }
    csync();

    STAMP(10)
    // ---- box refinement + heads (deformable_decoder.py:139-159, memotr.py:147-162)
    gemm(sprog, rg, xb, P256, Lp.bb0_b, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {           // bbox_embed.0 + ReLU
      *reinterpret_cast<uint32_t *>(bufB + g * P256 + col * 2) = pack_bf16(fmaxf(a[0] + b0, 0.f), fmaxf(a[1] + b1, 0.f));
      *reinterpret_cast<uint32_t *>(bufB + (g + 8) * P256 + col * 2) = pack_bf16(fmaxf(a[2] + b0, 0.f), fmaxf(a[3] + b1, 0.f));
    });
    csync();
    gemm(sprog, rg, bufB, P256, Lp.bb1_b, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {         // bbox_embed.1 + ReLU
      *reinterpret_cast<uint32_t *>(bufA + g * P512 + col * 2) = pack_bf16(fmaxf(a[0] + b0, 0.f), fmaxf(a[1] + b1, 0.f));
      *reinterpret_cast<uint32_t *>(bufA + (g + 8) * P512 + col * 2) = pack_bf16(fmaxf(a[2] + b0, 0.f), fmaxf(a[3] + b1, 0.f));
    });
    csync();
    {
      const int o = tid >> 2, part = tid & 3;                       // 64 outputs: (row, coordinate)
      const float dv = head_dot(bufA, P512, reinterpret_cast<const bf16 *>(Lp.bb2_w), o >> 2, o & 3, part);
      if (part == 0) delta[o] = dv + __ldg(Lp.bb2_b + (o & 3));
      for (int oo = tid >> 2; oo < R * P.ncls; oo += 64) {          // class head on the layer output
        const int r = oo / P.ncls, j = oo % P.ncls;
        const float lv = head_dot(xb, P256, reinterpret_cast<const bf16 *>(Lp.cls_w), r, j, part);
        if (part == 0 && row0 + r < nq) Lp.pred_logit[(long)(row0 + r) * P.ncls + j] = lv + __ldg(Lp.cls_b + j);
      }
    }
    csync();
    if (tid < R * 4) {
      const int r = tid >> 2, row = row0 + r;
      const float rf = refs[tid];
      const float nr = sigm(delta[tid] + inv_sigm(rf));
      const float nxt = row < n ? nr : rf;                           // n_take = n: bypassed rows keep their reference
      if (row < nq) {
        Lp.pred_box[(long)row * 4 + (tid & 3)] = nr;
        Lp.ref_out[(long)row * 4 + (tid & 3)] = nxt;
      }
      refs[tid] = nxt;
    }
    csync();
    STAMP(11)
  }
#undef STAMP
}

}  // namespace dec
}  // namespace memotr

using namespace memotr;

extern "C" int memotr_decoder_forward(const memotr_dec_params *p, void *stream) {
  MEMOTR_REQUIRE(p && p->prog && p->n_prog > 0 && p->tgt_in && p->ref_in && p->kbuf && p->vbuf && p->barrier && p->dim_t &&
                     p->valid_ratios,
                 "decoder_forward: null pointer");
  MEMOTR_REQUIRE(p->n_layers >= 1 && p->n_layers <= MEMOTR_DEC_MAX_LAYERS && p->nq >= 1 && p->nd >= 0 && p->nd <= p->nq &&
                     p->n_prog <= dec::MAX_PROG,
                 "decoder_forward: bad sizes");
  MEMOTR_REQUIRE(p->n_levels >= 1 && p->n_levels <= 8 && p->n_points >= 1 && p->n_levels * p->n_points * 3 * 8 <= dec::F0P,
                 "decoder_forward: levels x points too large");
  MEMOTR_REQUIRE((p->d_ffn > 1024 ? p->d_ffn % 512 == 0 && p->d_ffn <= 2048 : p->d_ffn % 256 == 0) && p->ncls >= 1 &&
                     p->np % 64 == 0 && p->np >= p->nq,
                 "decoder_forward: bad d_ffn / ncls / np");
  const int blocks = ceil_div(p->nq, dec::R);
  int dev = 0, n_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  MEMOTR_REQUIRE(blocks <= n_sm, "decoder_forward: %d row blocks exceed the %d SMs (grid barrier needs co-residency)", blocks, n_sm);
  static bool attr_set = false;
  if (!attr_set) {
    const cudaError_t e = cudaFuncSetAttribute(dec::decoder_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               dec::SMEM_TOTAL + 128);
    if (e != cudaSuccess) return fail(MEMOTR_ECUDA, "decoder_forward: smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(p->barrier, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess) return fail(MEMOTR_ECUDA, "decoder_forward: memset: %s", cudaGetErrorString(e));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(dec::NTHREADS);
  cfg.dynamicSmemBytes = dec::SMEM_TOTAL + 128;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;      // all row blocks resident: the grid barrier cannot dead-lock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, dec::decoder_fused_kernel, *p);
  if (e != cudaSuccess) return fail(MEMOTR_ECUDA, "decoder_forward: launch: %s", cudaGetErrorString(e));
  return check_launch("decoder_fused");
}
