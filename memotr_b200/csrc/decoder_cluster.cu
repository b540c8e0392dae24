// decoder_cluster.cu -- the fused decoder of decoder_fused.cu with a 4-CTA thread-block CLUSTER per 16-row block.
//
// decoder_fused.cu showed where one SM per row block ends (tools/prof_decoder.py, per layer): 58 us of dense layers bound by
// what one SM's 99 KB weight ring can pull from L2 (~75 GB/s), 19 us of bilinear gather and 11 us of attention bound by
// the memory-level parallelism of 8 warps -- all per-SM latency limits, with 123 of 148 SMs idle.  Here four CTAs on
// neighbouring SMs share a row block:
//   * every dense layer is split over its OUTPUT columns: rank r streams only the weight rows of its 64-column slice
//     (a quarter of the bytes per SM) and writes its slice of the result into the shared memory of all four CTAs
//     (st.shared::cluster), so each CTA again holds the full 16 x N activation as the next A operand;
//   * attention and the deformable gather are split over HEADS (two per CTA), the q / offset / logit projections of a
//     CTA's own heads are computed locally and never exchanged; the four warps of a head split the key blocks
//     (flash-decoding style) and combine (m, l, O) through shared memory;
//   * the FFN is split over the HIDDEN dimension: hidden slice local, partial outputs reduce-scattered and all-gathered;
//   * cheap row-local work (sine embedding, LayerNorm, box refinement, class head) is simply repeated on every CTA.
// Hand-offs inside the cluster are one mbarrier phase per exchange (remote arrive.release.cluster by one lane per warp,
// local try_wait.acquire.cluster); the only grid-wide dependency remains the K/V barrier per layer.
#include "decoder_common.cuh"

namespace memotr {
namespace dec {
namespace cl {

constexpr int HP = 512 * 2 + 16;                   // hidden slice pitch (<= 512 hidden columns per CTA)
constexpr int OFF_X32 = 0, OFF_XB = OFF_X32 + R * C * 4, OFF_QP = OFF_XB + R * P256, OFF_A = OFF_QP + R * P256,
              OFF_B = OFF_A + R * P512, OFF_H = OFF_B + R * P256, OFF_F0 = OFF_H + R * HP, OFF_RED = OFF_F0 + R * F0P * 4,
              OFF_RING = OFF_RED + CS * R * 64 * 4, OFF_MISC = OFF_RING + NSLOT * SLOT_BYTES, OFF_PROG = OFF_MISC + 512,
              MAX_PROG = 15 * MEMOTR_DEC_MAX_LAYERS, SMEM_TOTAL = OFF_PROG + MAX_PROG * 24;
static_assert(OFF_RING % 16 == 0 && SMEM_TOTAL + 128 <= 227 * 1024, "shared memory plan");

__global__ void __launch_bounds__(NTHREADS, 1) decoder_cluster_kernel(const __grid_constant__ memotr_dec_params P) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  float *x32 = reinterpret_cast<float *>(smem + OFF_X32);
  uint8_t *xb = smem + OFF_XB, *qp = smem + OFF_QP, *bufA = smem + OFF_A, *bufB = smem + OFF_B, *hbuf = smem + OFF_H;
  float *f0 = reinterpret_cast<float *>(smem + OFF_F0), *red = reinterpret_cast<float *>(smem + OFF_RED);
  float *refs = reinterpret_cast<float *>(smem + OFF_MISC), *delta = refs + 64;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + OFF_MISC + 384), *empty = full + NSLOT, *xbar = empty + NSLOT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  uint32_t rk;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rk));
  const int row0 = (blockIdx.x / CS) * R;
  const int nq = P.nq, nd = P.nd;
  const uint32_t sbase = s32(smem);

  memotr_dec_gemm *sprog = reinterpret_cast<memotr_dec_gemm *>(smem + OFF_PROG);   // this rank's weight program
  for (int i = tid; i < P.n_prog * 6; i += NTHREADS)
    reinterpret_cast<uint32_t *>(sprog)[i] = reinterpret_cast<const uint32_t *>(P.prog + (long)rk * P.n_prog)[i];
  for (int i = tid; i < R * HP / 4; i += NTHREADS) reinterpret_cast<uint32_t *>(hbuf)[i] = 0u;   // zero k-padding of the hidden slice
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) mbar_init(full + s, 1), mbar_init(empty + s, NCW);
    mbar_init(xbar, CS * NCW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();            // every CTA's barriers exist before the first remote arrive

  if (warp == NCW) {
    // ------------------------------------------------------------------ producer: this rank's slice of the weight program
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
    // -------------------------------------------------------------------- consumers (8 warps, 256 threads)
    Ring rg{smem + OFF_RING, full, empty, 0u, 0};
    Xchg xc;
    xc.bar = xbar, xc.phase = 0;
    uint32_t peer[CS];           // shared::cluster base address of every CTA's shared memory window
#pragma unroll
    for (int p = 0; p < CS; ++p) peer[p] = mapa(sbase, p), xc.bar_remote[p] = mapa(s32(xbar), p);
    const int g = lane >> 2, c = lane & 3;
    const int cb = 64 * (int)rk;                 // first output column of this rank's 64-column slice
    // write a bf16 pair / fp32 pair at (row r, column col) of buffer `off` (pitch in bytes) in ALL four CTAs
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

    for (int i = tid; i < R * C / 4; i += 256) {
      const int r = i / (C / 4), c4 = (i % (C / 4)) * 4;
      const int row = min(row0 + r, nq - 1);
      const float4 v = *reinterpret_cast<const float4 *>(P.tgt_in + (long)row * C + c4);
      *reinterpret_cast<float4 *>(x32 + r * C + c4) = v;
      *reinterpret_cast<uint2 *>(xb + r * P256 + c4 * 2) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
    if (tid < R * 4) refs[tid] = P.ref_in[(long)min(row0 + tid / 4, nq - 1) * 4 + (tid & 3)];
    csync();

#define STAMP(k) \
  if (P.prof && tid == 0) P.prof[((long)blockIdx.x * P.n_layers + lid) * 16 + (k)] = clock64();
    for (int lid = 0; lid < P.n_layers; ++lid) {
      const memotr_dec_layer &Lp = P.layers[lid];
      const int n = lid >= P.merge ? nq : nd;
      const int par = lid & 1;
      __half *Kh = reinterpret_cast<__half *>(P.kbuf) + (long)par * P.np * C;
      __half *Vt = reinterpret_cast<__half *>(P.vbuf) + (long)par * C * P.np;
      STAMP(0)
      // init_ref_pts / last_ref_pts of the output dict (memotr.py:183-187): inverse_sigmoid of the references that enter
      // the first / the last layer
      if (rk == 0 && tid < R * 4 && row0 + (tid >> 2) < nq) {
        const long o = (long)(row0 + (tid >> 2)) * 4 + (tid & 3);
        if (lid == 0 && P.init_ref_out) P.init_ref_out[o] = inv_sigm(refs[tid]);
        if (lid == P.n_layers - 1 && P.last_ref_out) P.last_ref_out[o] = inv_sigm(refs[tid]);
      }
      // ---- DAB positional query: sine embedding (every CTA, it is the A operand of a split GEMM)
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
      gemm(sprog, rg, bufA, P512, P.rph0_b + cb, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {
        bc_bf16(OFF_B, P256, g, cb + col, fmaxf(a[0] + b0, 0.f), fmaxf(a[1] + b1, 0.f));          // ref_point_head.0 + ReLU
        bc_bf16(OFF_B, P256, g + 8, cb + col, fmaxf(a[2] + b0, 0.f), fmaxf(a[3] + b1, 0.f));
      });
      xc.sync(lane);
      if (lid == 0) {
        gemm(sprog, rg, bufB, P256, P.rph1_b + cb, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {
          bc_bf16(OFF_QP, P256, g, cb + col, a[0] + b0, a[1] + b1);                                 // -> query_pos
          bc_bf16(OFF_QP, P256, g + 8, cb + col, a[2] + b0, a[3] + b1);
        });
        xc.sync(lane);
      } else {
        float raw[4];            // this thread's fragment of the raw query pos: the multiplier of its query_scale fragment
        gemm(sprog, rg, bufB, P256, P.rph1_b + cb, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {
          const __nv_bfloat162 r0 = __floats2bfloat162_rn(a[0] + b0, a[1] + b1), r1 = __floats2bfloat162_rn(a[2] + b0, a[3] + b1);
          raw[0] = __low2float(r0), raw[1] = __high2float(r0), raw[2] = __low2float(r1), raw[3] = __high2float(r1);
        });
        gemm(sprog, rg, xb, P256, P.qs0_b + cb, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {
          bc_bf16(OFF_B, P256, g, cb + col, fmaxf(a[0] + b0, 0.f), fmaxf(a[1] + b1, 0.f));        // query_scale.0 + ReLU
          bc_bf16(OFF_B, P256, g + 8, cb + col, fmaxf(a[2] + b0, 0.f), fmaxf(a[3] + b1, 0.f));
        });
        xc.sync(lane);
        gemm(sprog, rg, bufB, P256, P.qs1_b + cb, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {
          bc_bf16(OFF_QP, P256, g, cb + col, (a[0] + b0) * raw[0], (a[1] + b1) * raw[1]);           // * raw query pos
          bc_bf16(OFF_QP, P256, g + 8, cb + col, (a[2] + b0) * raw[2], (a[3] + b1) * raw[3]);
        });
        xc.sync(lane);
      }
      STAMP(1)
      // ---- self-attention projections of this CTA's two heads: q stays local, k / v go to global for every block
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
      gemm(sprog, rg, bufA, P512, Lp.qk_b + cb, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {   // q (fp16)
        const float sc = 0.17677669529663687f;
        *reinterpret_cast<uint32_t *>(bufB + g * P256 + col * 2) = pack_f16((a[0] + b0) * sc, (a[1] + b1) * sc);
        *reinterpret_cast<uint32_t *>(bufB + (g + 8) * P256 + col * 2) = pack_f16((a[2] + b0) * sc, (a[3] + b1) * sc);
      });
      gemm(sprog, rg, bufA, P512, Lp.qk_b + C + cb, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {   // k
        if (row0 + g < nq) *reinterpret_cast<uint32_t *>(Kh + (long)(row0 + g) * C + cb + col) = pack_f16(a[0] + b0, a[1] + b1);
        if (row0 + g + 8 < nq)
          *reinterpret_cast<uint32_t *>(Kh + (long)(row0 + g + 8) * C + cb + col) = pack_f16(a[2] + b0, a[3] + b1);
      });
      // (V of a padded key is written as zero: its softmax weight is exactly 0, but 0 x a stale non-finite row would be NaN)
      const bool keep0 = !(row0 + g < nq && P.query_pad && P.query_pad[row0 + g]);
      const bool keep1 = !(row0 + g + 8 < nq && P.query_pad && P.query_pad[row0 + g + 8]);
      gemm(sprog, rg, xb, P256, Lp.v_b + cb, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {       // v^T
        if (row0 + g < nq) {
          Vt[(long)(cb + col) * P.np + row0 + g] = __float2half_rn(keep0 ? a[0] + b0 : 0.f);
          Vt[(long)(cb + col + 1) * P.np + row0 + g] = __float2half_rn(keep0 ? a[1] + b1 : 0.f);
        }
        if (row0 + g + 8 < nq) {
          Vt[(long)(cb + col) * P.np + row0 + g + 8] = __float2half_rn(keep1 ? a[2] + b0 : 0.f);
          Vt[(long)(cb + col + 1) * P.np + row0 + g + 8] = __float2half_rn(keep1 ? a[3] + b1 : 0.f);
        }
      });
      STAMP(2)
      grid_barrier(P.barrier, (unsigned int)(lid + 1) * gridDim.x);
      STAMP(3)
      // ---- attention: heads 2*rk, 2*rk+1 (decoder_common.cuh); the result is broadcast into bufA of all four CTAs
      attention_2heads(bufB, Kh, Vt, P.np, n, P.query_pad, (int)rk, f0, peer, OFF_A, P512, warp, lane, tid);
      xc.sync(lane);
      STAMP(4)
      gemm(sprog, rg, bufA, P512, Lp.sao_b + cb, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {   // out_proj
        bc_f32(OFF_F0, F0P, g, cb + col, a[0] + b0 + x32[g * C + cb + col], a[1] + b1 + x32[g * C + cb + col + 1]);
        bc_f32(OFF_F0, F0P, g + 8, cb + col, a[2] + b0 + x32[(g + 8) * C + cb + col], a[3] + b1 + x32[(g + 8) * C + cb + col + 1]);
      });
      xc.sync(lane);
      layer_norm(f0, Lp.n2_g, Lp.n2_b, x32, xb, qp, bufA, warp, lane);                      // norm2; bufA = t1 + query_pos
      csync();
      STAMP(5)
      // ---- cross-attention: offsets / logits of this CTA's heads only (local), gather, then output_proj
      gemm(sprog, rg, bufA, P512, Lp.ol_b + cb, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {   // offsets
        *reinterpret_cast<float2 *>(f0 + g * F0P + col) = make_float2(a[0] + b0, a[1] + b1);
        *reinterpret_cast<float2 *>(f0 + (g + 8) * F0P + col) = make_float2(a[2] + b0, a[3] + b1);
      });
      gemm(sprog, rg, bufA, P512, nullptr, warp, lane, [&](int col, const float (&a)[4], float, float) {              // logits
        if (col < 32) {                                                             // rows 32..63 of the slot are zero padding
          const int LKh = P.n_levels * P.n_points;                                  // logits per head
          const float b0 = __ldg(Lp.ol_b + 16 * LKh + 2 * LKh * (int)rk + col), b1 = __ldg(Lp.ol_b + 16 * LKh + 2 * LKh * (int)rk + col + 1);
          *reinterpret_cast<float2 *>(f0 + g * F0P + 64 + col) = make_float2(a[0] + b0, a[1] + b1);
          *reinterpret_cast<float2 *>(f0 + (g + 8) * F0P + 64 + col) = make_float2(a[2] + b0, a[3] + b1);
        }
      });
      csync();
      STAMP(6)
      {
        // local layout of a row of f0: [offsets of head 2rk (LK x 2) | offsets of head 2rk+1 | logits 2rk (LK) | logits 2rk+1]
        const int Kp = P.n_points, Lv = P.n_levels, LK = Lv * Kp;
        const int xs = P.value_stride;
        const int pair = (tid & 127) >> 2, sub = tid & 3, lhalf = tid >> 7;     // 32 (q, head) pairs x 4 lanes x 2 level halves
        const int r = pair >> 1, hl = pair & 1, h = 2 * (int)rk + hl;
        const float *rowp = f0 + r * F0P;
        const float *lg = rowp + 2 * 2 * LK + hl * LK;       // the two heads' offsets take 2 * LK * 2 floats
        float mx = -INFINITY;
        for (int i = 0; i < LK; ++i) mx = fmaxf(mx, lg[i]);
        float sum = 0.f;
        for (int i = 0; i < LK; ++i) sum += __expf(lg[i] - mx);
        const float rs = __frcp_rn(sum), rkk = __frcp_rn((float)Kp);
        const float rx = refs[r * 4], ry = refs[r * 4 + 1], rw = refs[r * 4 + 2], rh = refs[r * 4 + 3];
        const __half *vb = reinterpret_cast<const __half *>(Lp.value) + h * 32 + sub * 8;
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        for (int l = lhalf; l < Lv; l += 2) {
          const int Hh = P.shapes[2 * l], Ww = P.shapes[2 * l + 1];
          const float Hf = (float)Hh, Wf = (float)Ww;
          const float vx = __ldg(P.valid_ratios + 2 * l), vy = __ldg(P.valid_ratios + 2 * l + 1);
          const long base = (long)P.lsi[l] * xs;
          const int ys = Ww * xs;
          __half2 a2[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) a2[j] = __float2half2_rn(0.f);
          for (int pb = 0; pb < Kp; pb += 4) {
            uint4 rv[4][4];
            __half2 wq[4][4];
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
              const int p = min(pb + pp, Kp - 1), i = l * Kp + p;
              const bool pv = pb + pp < Kp;
              const float2 off = *reinterpret_cast<const float2 *>(rowp + (hl * LK + i) * 2);
              const float aw = pv ? __expf(lg[i] - mx) * rs : 0.f;
              const float lx = rx * vx + off.x * rkk * (rw * vx) * 0.5f, ly = ry * vy + off.y * rkk * (rh * vy) * 0.5f;
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
        float *part = red + (tid & 127) * 8;          // the odd-level half hands its sum to the even-level half
        if (lhalf == 1) {
#pragma unroll
          for (int k = 0; k < 8; ++k) part[k] = acc[k];
        }
        csync();
        if (lhalf == 0) {
#pragma unroll
          for (int k = 0; k < 8; k += 2)
            bc_bf16(OFF_A, P512, r, cb + hl * 32 + sub * 8 + k, acc[k] + part[k], acc[k + 1] + part[k + 1]);
        }
      }
      xc.sync(lane);
      STAMP(7)
      gemm(sprog, rg, bufA, P512, Lp.cao_b + cb, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {   // output_proj
        bc_f32(OFF_F0, F0P, g, cb + col, a[0] + b0 + x32[g * C + cb + col], a[1] + b1 + x32[g * C + cb + col + 1]);
        bc_f32(OFF_F0, F0P, g + 8, cb + col, a[2] + b0 + x32[(g + 8) * C + cb + col], a[3] + b1 + x32[(g + 8) * C + cb + col + 1]);
      });
      xc.sync(lane);
      layer_norm(f0, Lp.n1_g, Lp.n1_b, x32, xb, nullptr, nullptr, warp, lane);              // norm1
      csync();
      STAMP(8)
      // ---- FFN: this CTA's quarter of the hidden dimension stays local; partial outputs are reduce-scattered
      const int hw_ = P.d_ffn / CS;                                                          // hidden columns per CTA
      gemm(sprog, rg, xb, P256, Lp.f1_b + hw_ * (int)rk, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {
        *reinterpret_cast<uint32_t *>(hbuf + g * HP + col * 2) = pack_bf16(fmaxf(a[0] + b0, 0.f), fmaxf(a[1] + b1, 0.f));
        *reinterpret_cast<uint32_t *>(hbuf + (g + 8) * HP + col * 2) = pack_bf16(fmaxf(a[2] + b0, 0.f), fmaxf(a[3] + b1, 0.f));
      });
      csync();
      gemm(sprog, rg, hbuf, HP, nullptr, warp, lane, [&](int col, const float (&a)[4], float, float) {
        // partial sum of output columns col, col+1: they belong to rank col / 64 -> its red[rk] slot ([src][row][64])
        const uint32_t o = OFF_RED + (((int)rk * R + g) * 64 + (col & 63)) * 4, dst = peer[col >> 6];
        st_cl_f32x2(dst + o, a[0], a[1]);
        st_cl_f32x2(dst + o + 8 * 64 * 4, a[2], a[3]);
      });
      xc.sync(lane);
      for (int i = tid; i < R * 32; i += 256) {            // sum the four partials of this rank's 64 columns, + bias + residual
        const int r = i >> 5, c2 = (i & 31) * 2;
        float s0 = __ldg(Lp.f2_b + cb + c2) + x32[r * C + cb + c2], s1 = __ldg(Lp.f2_b + cb + c2 + 1) + x32[r * C + cb + c2 + 1];
#pragma unroll
        for (int p = 0; p < CS; ++p) s0 += red[(p * R + r) * 64 + c2], s1 += red[(p * R + r) * 64 + c2 + 1];
        bc_f32(OFF_F0, F0P, r, cb + c2, s0, s1);
      }
      xc.sync(lane);
      layer_norm(f0, Lp.n3_g, Lp.n3_b, x32, xb, nullptr, nullptr, warp, lane);              // norm3 -> the layer output
      csync();
      STAMP(9)
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
          if ((r & 3) == (int)rk) *reinterpret_cast<float4 *>(Lp.tgt_out + (long)row * C + c4) = v;   // each rank stores 4 rows
        }
      }
      csync();
      STAMP(10)
      // ---- box refinement + heads
      gemm(sprog, rg, xb, P256, Lp.bb0_b + cb, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {
        bc_bf16(OFF_B, P256, g, cb + col, fmaxf(a[0] + b0, 0.f), fmaxf(a[1] + b1, 0.f));
        bc_bf16(OFF_B, P256, g + 8, cb + col, fmaxf(a[2] + b0, 0.f), fmaxf(a[3] + b1, 0.f));
      });
      xc.sync(lane);
      gemm(sprog, rg, bufB, P256, Lp.bb1_b + cb, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {
        bc_bf16(OFF_A, P512, g, cb + col, fmaxf(a[0] + b0, 0.f), fmaxf(a[1] + b1, 0.f));
        bc_bf16(OFF_A, P512, g + 8, cb + col, fmaxf(a[2] + b0, 0.f), fmaxf(a[3] + b1, 0.f));
      });
      xc.sync(lane);
      {
        const int o = tid >> 2, part = tid & 3;
        const float dv = head_dot(bufA, P512, reinterpret_cast<const bf16 *>(Lp.bb2_w), o >> 2, o & 3, part);
        if (part == 0) delta[o] = dv + __ldg(Lp.bb2_b + (o & 3));
        if (rk == 0) {
          for (int oo = tid >> 2; oo < R * P.ncls; oo += 64) {
            const int r = oo / P.ncls, j = oo % P.ncls;
            const float lv = head_dot(xb, P256, reinterpret_cast<const bf16 *>(Lp.cls_w), r, j, part);
            if (part == 0 && row0 + r < nq) Lp.pred_logit[(long)(row0 + r) * P.ncls + j] = lv + __ldg(Lp.cls_b + j);
          }
        }
      }
      csync();
      if (tid < R * 4) {
        const int r = tid >> 2, row = row0 + r;
        const float rf = refs[tid];
        const float nr = sigm(delta[tid] + inv_sigm(rf));
        const float nxt = row < n ? nr : rf;
        if (row < nq && rk == 0) {
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
  cluster_sync_all();            // no CTA leaves while a peer may still store into its shared memory
}

}  // namespace cl
}  // namespace dec
}  // namespace memotr

using namespace memotr;

extern "C" int memotr_decoder_forward_cluster(const memotr_dec_params *p, void *stream) {
  MEMOTR_REQUIRE(p && p->prog && p->n_prog > 0 && p->tgt_in && p->ref_in && p->kbuf && p->vbuf && p->barrier && p->dim_t &&
                     p->valid_ratios,
                 "decoder_forward_cluster: null pointer");
  MEMOTR_REQUIRE(p->n_layers >= 1 && p->n_layers <= MEMOTR_DEC_MAX_LAYERS && p->nq >= 1 && p->nd >= 0 && p->nd <= p->nq &&
                     p->n_prog <= dec::cl::MAX_PROG,
                 "decoder_forward_cluster: bad sizes");
  MEMOTR_REQUIRE(p->n_levels >= 1 && p->n_levels <= 8 && p->n_points >= 1 && p->n_levels * p->n_points == 16,
                 "decoder_forward_cluster: needs levels x points == 16 (one 64-row slot of offsets, 32 logits per CTA)");
  MEMOTR_REQUIRE(p->d_ffn % (dec::cl::CS * 64) == 0 && p->d_ffn <= dec::cl::CS * 512 && p->ncls >= 1 && p->np % 64 == 0 &&
                     p->np >= p->nq,
                 "decoder_forward_cluster: bad d_ffn / ncls / np");
  const int blocks = ceil_div(p->nq, dec::R) * dec::cl::CS;
  int dev = 0, n_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  MEMOTR_REQUIRE(blocks <= n_sm, "decoder_forward_cluster: %d CTAs exceed the %d SMs (grid barrier needs co-residency)", blocks,
                 n_sm);
  static bool attr_set = false;
  if (!attr_set) {
    const cudaError_t e = cudaFuncSetAttribute(dec::cl::decoder_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               dec::cl::SMEM_TOTAL + 128);
    if (e != cudaSuccess) return fail(MEMOTR_ECUDA, "decoder_forward_cluster: smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(p->barrier, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess) return fail(MEMOTR_ECUDA, "decoder_forward_cluster: memset: %s", cudaGetErrorString(e));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(dec::NTHREADS);
  cfg.dynamicSmemBytes = dec::cl::SMEM_TOTAL + 128;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = dec::cl::CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeCooperative;      // all CTAs resident: the grid barrier cannot dead-lock
  attr[1].val.cooperative = 1;
  cfg.attrs = attr;
  // MEMOTR_NONCOOP=1 (profiling only): plain cluster launch -- ncu cannot replay cooperative cluster launches; with the GPU to
  // itself (kernels serialised under the profiler, grid <= number of SMs) all CTAs are resident anyway
  const char *nc = getenv("MEMOTR_NONCOOP");
  cfg.numAttrs = (nc && nc[0] == '1') ? 1 : 2;
  e = cudaLaunchKernelEx(&cfg, dec::cl::decoder_cluster_kernel, *p);
  if (e != cudaSuccess) return fail(MEMOTR_ECUDA, "decoder_forward_cluster: launch: %s", cudaGetErrorString(e));
  return check_launch("decoder_cluster");
}
