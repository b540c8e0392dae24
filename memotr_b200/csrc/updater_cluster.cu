// updater_cluster.cu -- QueryUpdater.update_tracks_embedding (models/query_updater.py:82-166, DAB branch) as ONE persistent
// kernel on the device-resident track table, same machinery as decoder_cluster.cu: a 4-CTA cluster per block of 16 track
// rows, dense layers split over output columns (FFNs over the hidden dimension) with the weights streamed from L2 in
// program order, the long-term-memory attention split over heads with its keys / values exchanged through global memory
// and ONE grid barrier.  Replaces ~25 launch-latency-bound launches per frame (131 us) of the launch-per-op engine, and
// also writes the fed-back track queries of the next frame (submit_engine.py:64-70).
//
//   is_pos = max_c sigmoid(logits) > update_thresh ; ref = is_pos ? inverse_sigmoid(boxes) : ref_pts          (:84-85,99-102)
//   conf = sigmoid(MLP(out_e)) ; short = MLP([conf * out_e | last_output])                                     (:109-118)
//   query_pos = MLP(sine(sigmoid(ref))) ; q = short + query_pos ; k = long_memory + query_pos ; v = out_e      (:103,120-124)
//   tgt = FFN(LN(out_e + MHA(q, k, v))) ; feat = FFN(LN(long_memory + tgt))                                   (:125-133)
//   where is_pos: long_memory <- (1 - lambda) long_memory + lambda out_e ; last_output <- out_e ; query_embed <- feat  (:135-147)
// Numerics: those of the bf16 engine (bf16 GEMM operands, fp32 accumulate / residual / LayerNorm / state).
#include "decoder_common.cuh"

namespace memotr {
namespace dec {
namespace upd {

using namespace cl;
constexpr int HP = 512 * 2 + 16;
constexpr int OFF_X32 = 0, OFF_XB = OFF_X32 + R * C * 4, OFF_QP = OFF_XB + R * P256, OFF_A = OFF_QP + R * P256,
              OFF_B = OFF_A + R * P512, OFF_H = OFF_B + R * P256, OFF_F0 = OFF_H + R * HP, OFF_RED = OFF_F0 + R * F0P * 4,
              OFF_RING = OFF_RED + CS * R * 64 * 4, OFF_MISC = OFF_RING + NSLOT * SLOT_BYTES, OFF_PROG = OFF_MISC + 512,
              MAX_PROG = 16, SMEM_TOTAL = OFF_PROG + MAX_PROG * 24;
static_assert(OFF_RING % 16 == 0 && SMEM_TOTAL + 128 <= 227 * 1024, "shared memory plan");

__global__ void __launch_bounds__(NTHREADS, 1) updater_cluster_kernel(const __grid_constant__ memotr_upd_params P) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  float *x32 = reinterpret_cast<float *>(smem + OFF_X32);
  uint8_t *xb = smem + OFF_XB, *qp = smem + OFF_QP, *bufA = smem + OFF_A, *bufB = smem + OFF_B, *hbuf = smem + OFF_H;
  float *f0 = reinterpret_cast<float *>(smem + OFF_F0), *red = reinterpret_cast<float *>(smem + OFF_RED);
  float *refs = reinterpret_cast<float *>(smem + OFF_MISC);               // [16][4] updated reference points (logit space)
  int *ispos = reinterpret_cast<int *>(refs + 64);                        // [16]
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + OFF_MISC + 384), *empty = full + NSLOT, *xbar = empty + NSLOT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  uint32_t rk;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rk));
  const int row0 = (blockIdx.x / CS) * R, nt = P.nt;
  const uint32_t sbase = s32(smem);

  memotr_dec_gemm *sprog = reinterpret_cast<memotr_dec_gemm *>(smem + OFF_PROG);
  for (int i = tid; i < P.n_prog * 6; i += NTHREADS)
    reinterpret_cast<uint32_t *>(sprog)[i] = reinterpret_cast<const uint32_t *>(P.prog + (long)rk * P.n_prog)[i];
  for (int i = tid; i < R * HP / 4; i += NTHREADS) reinterpret_cast<uint32_t *>(hbuf)[i] = 0u;
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) mbar_init(full + s, 1), mbar_init(empty + s, NCW);
    mbar_init(xbar, CS * NCW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();

  if (warp == NCW) {
    uint32_t t = 0;
    for (int gi = 0; gi < P.n_prog; ++gi) {
      const memotr_dec_gemm d = sprog[gi];
      const uint8_t *W = reinterpret_cast<const uint8_t *>(d.W);
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
  } else {
    Ring rg{smem + OFF_RING, full, empty, 0u, 0};
    Xchg xc;
    xc.bar = xbar, xc.phase = 0;
    uint32_t peer[CS];
#pragma unroll
    for (int p = 0; p < CS; ++p) peer[p] = mapa(sbase, p), xc.bar_remote[p] = mapa(s32(xbar), p);
    const int g = lane >> 2;
    const int cb = 64 * (int)rk;
    auto bc_bf16 = [&](int off, int pitch, int r, int col, float v0, float v1) {
      const uint32_t o = off + r * pitch + col * 2, v = pack_bf16(v0, v1);
#pragma unroll
      for (int p = 0; p < CS; ++p) st_cl_u32(peer[p] + o, v);
    };
    auto bc_f32 = [&](int off, int pitchf, int r, int col, float v0, float v1) {
      const uint32_t o = off + (r * pitchf + col) * 4;
#pragma unroll
      for (int p = 0; p < CS; ++p) st_cl_f32x2(peer[p] + o, v0, v1);
    };
    auto rowc = [&](int r) { return min(row0 + r, nt - 1); };   // rows beyond the table shadow the last row (never stored)

    // ---- inputs of the block: out_e (fp32 + bf16), [ . | last_output ] half of the fusion input, is_pos, updated reference
    for (int i = tid; i < R * C / 4; i += 256) {
      const int r = i / (C / 4), c4 = (i % (C / 4)) * 4;
      const float4 v = *reinterpret_cast<const float4 *>(P.output_embed + (long)rowc(r) * C + c4);
      *reinterpret_cast<float4 *>(x32 + r * C + c4) = v;
      *reinterpret_cast<uint2 *>(xb + r * P256 + c4 * 2) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
      const float4 lo = *reinterpret_cast<const float4 *>(P.last_output + (long)rowc(r) * C + c4);
      *reinterpret_cast<uint2 *>(bufA + r * P512 + (C + c4) * 2) = make_uint2(pack_bf16(lo.x, lo.y), pack_bf16(lo.z, lo.w));
    }
    if (tid < R) {
      float s = -INFINITY;
      for (int c = 0; c < P.ncls; ++c) s = fmaxf(s, sigm(P.logits[(long)rowc(tid) * P.ncls + c]));
      ispos[tid] = s > P.update_thresh;
    }
    csync();
    if (tid < R * 4) {
      const int r = tid >> 2;
      const float v = ispos[r] ? inv_sigm(P.boxes[(long)rowc(r) * 4 + (tid & 3)]) : P.ref_pts[(long)rowc(r) * 4 + (tid & 3)];
      refs[tid] = v;
    }
    // ---- confidence gate and short-memory fusion
    gemm(sprog, rg, xb, P256, P.conf0_b + cb, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {
      bc_bf16(OFF_B, P256, g, cb + col, fmaxf(a[0] + b0, 0.f), fmaxf(a[1] + b1, 0.f));
      bc_bf16(OFF_B, P256, g + 8, cb + col, fmaxf(a[2] + b0, 0.f), fmaxf(a[3] + b1, 0.f));
    });
    xc.sync(lane);
    gemm(sprog, rg, bufB, P256, P.conf1_b + cb, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {
      const float2 o0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(xb + g * P256 + (cb + col) * 2));
      const float2 o1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(xb + (g + 8) * P256 + (cb + col) * 2));
      bc_bf16(OFF_A, P512, g, cb + col, sigm(a[0] + b0) * o0.x, sigm(a[1] + b1) * o0.y);        // conf * out_e
      bc_bf16(OFF_A, P512, g + 8, cb + col, sigm(a[2] + b0) * o1.x, sigm(a[3] + b1) * o1.y);
    });
    xc.sync(lane);
    gemm(sprog, rg, bufA, P512, P.fus0_b + 2 * cb, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {
      bc_bf16(OFF_H, HP, g, 2 * cb + col, fmaxf(a[0] + b0, 0.f), fmaxf(a[1] + b1, 0.f));       // 128 of 512 columns per rank
      bc_bf16(OFF_H, HP, g + 8, 2 * cb + col, fmaxf(a[2] + b0, 0.f), fmaxf(a[3] + b1, 0.f));
    });
    xc.sync(lane);
    gemm(sprog, rg, hbuf, HP, P.fus1_b + cb, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {
      bc_bf16(OFF_QP, P256, g, cb + col, a[0] + b0, a[1] + b1);                                  // short memory
      bc_bf16(OFF_QP, P256, g + 8, cb + col, a[2] + b0, a[3] + b1);
    });
    // ---- query_pos = MLP(sine(sigmoid(ref))) (pos_to_pos_embed, models/utils.py:78-85)
    for (int i = tid; i < R * 256; i += 256) {
      const int r = i >> 8, cc = (i >> 6) & 3, j = i & 63;
      const float e = sigm(refs[r * 4 + cc]) * 6.283185307179586f * __frcp_rn(__ldg(P.dim_t + 2 * j));
      *reinterpret_cast<uint32_t *>(bufA + r * P512 + (cc * 128 + 2 * j) * 2) = pack_bf16(__sinf(e), __cosf(e));
    }
    xc.sync(lane);                                                                               // short memory complete; anchor local
    gemm(sprog, rg, bufA, P512, P.ph0_b + cb, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {
      bc_bf16(OFF_B, P256, g, cb + col, fmaxf(a[0] + b0, 0.f), fmaxf(a[1] + b1, 0.f));
      bc_bf16(OFF_B, P256, g + 8, cb + col, fmaxf(a[2] + b0, 0.f), fmaxf(a[3] + b1, 0.f));
    });
    xc.sync(lane);
    gemm(sprog, rg, bufB, P256, P.ph1_b + cb, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {
      bc_bf16(OFF_A, P512, g, cb + col, a[0] + b0, a[1] + b1);                                   // query_pos -> bufA[:, 0:256]
      bc_bf16(OFF_A, P512, g + 8, cb + col, a[2] + b0, a[3] + b1);
    });
    xc.sync(lane);
    // q input = short + query_pos -> bufB ; k input = long_memory + query_pos -> bufA[:, 256:512]
    for (int i = tid; i < R * C / 8; i += 256) {
      const int r = i / (C / 8), c8 = (i % (C / 8)) * 8;
      float sh[8], qq[8], lm[8];
      bf16x8_to_f32(*reinterpret_cast<const uint4 *>(qp + r * P256 + c8 * 2), sh);
      bf16x8_to_f32(*reinterpret_cast<const uint4 *>(bufA + r * P512 + c8 * 2), qq);
      const float4 l0 = *reinterpret_cast<const float4 *>(P.long_memory + (long)rowc(r) * C + c8);
      const float4 l1 = *reinterpret_cast<const float4 *>(P.long_memory + (long)rowc(r) * C + c8 + 4);
      const float lf[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        sh[k] += qq[k];
        lm[k] = __bfloat162float(__float2bfloat16_rn(lf[k])) + qq[k];     // the engine rounds long_memory to bf16 first
      }
      *reinterpret_cast<uint4 *>(bufB + r * P256 + c8 * 2) = f32x8_to_bf16(sh);
      *reinterpret_cast<uint4 *>(bufA + r * P512 + (C + c8) * 2) = f32x8_to_bf16(lm);
    }
    csync();
    // ---- long-term-memory attention: projections of this CTA's two heads, k / v to global, grid barrier
    __half *Kh = reinterpret_cast<__half *>(P.kbuf), *Vt = reinterpret_cast<__half *>(P.vbuf);
    gemm(sprog, rg, bufB, P256, P.q_b + cb, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {
      const float sc = 0.17677669529663687f;
      *reinterpret_cast<uint32_t *>(qp + g * P256 + col * 2) = pack_f16((a[0] + b0) * sc, (a[1] + b1) * sc);
      *reinterpret_cast<uint32_t *>(qp + (g + 8) * P256 + col * 2) = pack_f16((a[2] + b0) * sc, (a[3] + b1) * sc);
    });
    gemm(sprog, rg, bufA + C * 2, P512, P.k_b + cb, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {
      if (row0 + g < nt) *reinterpret_cast<uint32_t *>(Kh + (long)(row0 + g) * C + cb + col) = pack_f16(a[0] + b0, a[1] + b1);
      if (row0 + g + 8 < nt)
        *reinterpret_cast<uint32_t *>(Kh + (long)(row0 + g + 8) * C + cb + col) = pack_f16(a[2] + b0, a[3] + b1);
    });
    // (V of a padded key is written as zero: its softmax weight is exactly 0, but 0 x a stale non-finite row would be NaN)
    const float keep0 = (row0 + g < nt && P.track_pad && P.track_pad[row0 + g]) ? 0.f : 1.f;
    const float keep1 = (row0 + g + 8 < nt && P.track_pad && P.track_pad[row0 + g + 8]) ? 0.f : 1.f;
    gemm(sprog, rg, xb, P256, P.v_b + cb, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {
      if (row0 + g < nt) {
        Vt[(long)(cb + col) * P.np + row0 + g] = __float2half_rn(keep0 != 0.f ? a[0] + b0 : 0.f);
        Vt[(long)(cb + col + 1) * P.np + row0 + g] = __float2half_rn(keep0 != 0.f ? a[1] + b1 : 0.f);
      }
      if (row0 + g + 8 < nt) {
        Vt[(long)(cb + col) * P.np + row0 + g + 8] = __float2half_rn(keep1 != 0.f ? a[2] + b0 : 0.f);
        Vt[(long)(cb + col + 1) * P.np + row0 + g + 8] = __float2half_rn(keep1 != 0.f ? a[3] + b1 : 0.f);
      }
    });
    grid_barrier(P.barrier, gridDim.x);
    attention_2heads(qp, Kh, Vt, P.np, nt, P.track_pad, (int)rk, f0, peer, OFF_B, P256, warp, lane, tid);
    xc.sync(lane);
    gemm(sprog, rg, bufB, P256, P.out_b + cb, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {
      bc_f32(OFF_F0, F0P, g, cb + col, a[0] + b0 + x32[g * C + cb + col], a[1] + b1 + x32[g * C + cb + col + 1]);
      bc_f32(OFF_F0, F0P, g + 8, cb + col, a[2] + b0 + x32[(g + 8) * C + cb + col], a[3] + b1 + x32[(g + 8) * C + cb + col + 1]);
    });
    xc.sync(lane);
    layer_norm(f0, P.mn_g, P.mn_b, x32, xb, nullptr, nullptr, warp, lane);                       // memory_norm
    csync();
    // ---- the two FFN blocks (memory_ffn, then query_feat_ffn on long_memory + tgt), hidden dimension split over the cluster
    const int hw_ = P.d_ffn / CS;
    for (int blk = 0; blk < 2; ++blk) {
      const float *b1p = blk ? P.ff1_b : P.mf1_b, *b2p = blk ? P.ff2_b : P.mf2_b;
      if (blk == 1) {                                   // query_feat_norm(long_memory + tgt): row-local, every CTA
        for (int i = tid; i < R * C / 4; i += 256) {
          const int r = i / (C / 4), c4 = (i % (C / 4)) * 4;
          const float4 l = *reinterpret_cast<const float4 *>(P.long_memory + (long)rowc(r) * C + c4);
          const float4 t = *reinterpret_cast<const float4 *>(x32 + r * C + c4);
          *reinterpret_cast<float4 *>(f0 + r * F0P + c4) = make_float4(l.x + t.x, l.y + t.y, l.z + t.z, l.w + t.w);
        }
        csync();
        layer_norm(f0, P.fn_g, P.fn_b, x32, xb, nullptr, nullptr, warp, lane);
        csync();
      }
      gemm(sprog, rg, xb, P256, b1p + hw_ * (int)rk, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {
        *reinterpret_cast<uint32_t *>(hbuf + g * HP + col * 2) = pack_bf16(fmaxf(a[0] + b0, 0.f), fmaxf(a[1] + b1, 0.f));
        *reinterpret_cast<uint32_t *>(hbuf + (g + 8) * HP + col * 2) = pack_bf16(fmaxf(a[2] + b0, 0.f), fmaxf(a[3] + b1, 0.f));
      });
      csync();
      gemm(sprog, rg, hbuf, HP, nullptr, warp, lane, [&](int col, const float (&a)[4], float, float) {
        const uint32_t o = OFF_RED + (((int)rk * R + g) * 64 + (col & 63)) * 4, dst = peer[col >> 6];
        st_cl_f32x2(dst + o, a[0], a[1]);
        st_cl_f32x2(dst + o + 8 * 64 * 4, a[2], a[3]);
      });
      xc.sync(lane);
      for (int i = tid; i < R * 32; i += 256) {
        const int r = i >> 5, c2 = (i & 31) * 2;
        float s0 = __ldg(b2p + cb + c2) + x32[r * C + cb + c2], s1 = __ldg(b2p + cb + c2 + 1) + x32[r * C + cb + c2 + 1];
#pragma unroll
        for (int p = 0; p < CS; ++p) s0 += red[(p * R + r) * 64 + c2], s1 += red[(p * R + r) * 64 + c2 + 1];
        bc_f32(OFF_F0, F0P, r, cb + c2, s0, s1);
      }
      xc.sync(lane);
      layer_norm(f0, blk ? P.ffn_g : P.mfn_g, blk ? P.ffn_b : P.mfn_b, x32, xb, nullptr, nullptr, warp, lane);
      csync();
    }
    // ---- masked state writes + the next frame's track queries; each rank stores four of the sixteen rows
    for (int i = tid; i < R * C / 4; i += 256) {
      const int r = i / (C / 4), c4 = (i % (C / 4)) * 4, row = row0 + r;
      if (row >= nt || (r & 3) != (int)rk) continue;
      const long o = (long)row * C + c4;
      float4 qe = *reinterpret_cast<const float4 *>(P.query_embed + o);
      if (ispos[r]) {
        const float lam = P.long_memory_lambda;
        const float4 oe = *reinterpret_cast<const float4 *>(P.output_embed + o);
        const float4 lm = *reinterpret_cast<const float4 *>(P.long_memory + o);
        *reinterpret_cast<float4 *>(P.long_memory + o) = make_float4((1.f - lam) * lm.x + lam * oe.x, (1.f - lam) * lm.y + lam * oe.y,
                                                                     (1.f - lam) * lm.z + lam * oe.z, (1.f - lam) * lm.w + lam * oe.w);
        *reinterpret_cast<float4 *>(P.last_output + o) = oe;
        qe = *reinterpret_cast<const float4 *>(x32 + r * C + c4);
        *reinterpret_cast<float4 *>(P.query_embed + o) = qe;
      }
      if (P.feedback_embed) *reinterpret_cast<float4 *>(P.feedback_embed + o) = qe;
    }
    if (tid < R * 4 && row0 + (tid >> 2) < nt && ((tid >> 2) & 3) == (int)rk) {
      const long o = (long)(row0 + (tid >> 2)) * 4 + (tid & 3);
      P.ref_pts[o] = refs[tid];
      if (P.feedback_ref) P.feedback_ref[o] = refs[tid];
    }
  }
  cluster_sync_all();
}

}  // namespace upd
}  // namespace dec
}  // namespace memotr

using namespace memotr;

extern "C" int memotr_updater_forward_cluster(const memotr_upd_params *p, void *stream) {
  MEMOTR_REQUIRE(p && p->prog && p->n_prog > 0 && p->n_prog <= dec::upd::MAX_PROG && p->logits && p->boxes && p->output_embed &&
                     p->ref_pts && p->query_embed && p->long_memory && p->last_output && p->kbuf && p->vbuf && p->barrier &&
                     p->dim_t,
                 "updater_forward_cluster: null pointer");
  MEMOTR_REQUIRE(p->nt >= 1 && p->ncls >= 1 && p->d_ffn % 256 == 0 && p->d_ffn <= 2048 && p->np % 64 == 0 && p->np >= p->nt,
                 "updater_forward_cluster: bad sizes");
  const int blocks = ceil_div(p->nt, dec::R) * dec::cl::CS;
  int dev = 0, n_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  MEMOTR_REQUIRE(blocks <= n_sm, "updater_forward_cluster: %d CTAs exceed the %d SMs (grid barrier needs co-residency)", blocks,
                 n_sm);
  static bool attr_set = false;
  if (!attr_set) {
    const cudaError_t e = cudaFuncSetAttribute(dec::upd::updater_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               dec::upd::SMEM_TOTAL + 128);
    if (e != cudaSuccess) return fail(MEMOTR_ECUDA, "updater_forward_cluster: smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(p->barrier, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess) return fail(MEMOTR_ECUDA, "updater_forward_cluster: memset: %s", cudaGetErrorString(e));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(dec::NTHREADS);
  cfg.dynamicSmemBytes = dec::upd::SMEM_TOTAL + 128;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = dec::cl::CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeCooperative;
  attr[1].val.cooperative = 1;
  cfg.attrs = attr;
  // MEMOTR_NONCOOP=1 (profiling only): plain cluster launch -- ncu cannot replay cooperative cluster launches; with the GPU to
  // itself (kernels serialised under the profiler, grid <= number of SMs) all CTAs are resident anyway
  const char *nc = getenv("MEMOTR_NONCOOP");
  cfg.numAttrs = (nc && nc[0] == '1') ? 1 : 2;
  e = cudaLaunchKernelEx(&cfg, dec::upd::updater_cluster_kernel, *p);
  if (e != cudaSuccess) return fail(MEMOTR_ECUDA, "updater_forward_cluster: launch: %s", cudaGetErrorString(e));
  return check_launch("updater_cluster");
}
