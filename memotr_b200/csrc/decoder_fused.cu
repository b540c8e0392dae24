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
#include "common.cuh"

namespace memotr {
namespace dec {

using bf16 = __nv_bfloat16;
constexpr int R = 16, C = 256, NCW = 8, NTHREADS = (NCW + 1) * 32;
constexpr int SLOT_ROWS = 64, SLOT_K = 256, WP = SLOT_K * 2 + 16, SLOT_BYTES = SLOT_ROWS * WP, NSLOT = 3;
constexpr int P256 = 256 * 2 + 16, P512 = 512 * 2 + 16, P1024 = 1024 * 2 + 16;  // bf16 row pitches (ldmatrix conflict-free)
constexpr int F0P = 512;                                                         // fp32 scratch pitch (floats)
constexpr int OFF_X32 = 0, OFF_XB = OFF_X32 + R * C * 4, OFF_QP = OFF_XB + R * P256, OFF_A = OFF_QP + R * P256,
              OFF_B = OFF_A + R * P512, OFF_H = OFF_B + R * P256, OFF_F0 = OFF_H + R * P1024,
              OFF_RING = OFF_F0 + R * F0P * 4, OFF_MISC = OFF_RING + NSLOT * SLOT_BYTES, OFF_PROG = OFF_MISC + 512,
              MAX_PROG = 15 * MEMOTR_DEC_MAX_LAYERS, SMEM_TOTAL = OFF_PROG + MAX_PROG * 24;
static_assert(OFF_RING % 16 == 0 && SMEM_TOTAL + 128 <= 227 * 1024 && sizeof(memotr_dec_gemm) == 24, "shared memory plan");

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tW_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra W_DONE;\n\tbra W_LOOP;\n\tW_DONE:\n\t}" ::"r"(
          s32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_row(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)),
               "l"(src), "r"(bytes), "r"(s32(bar))
               : "memory");
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const void *p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(s32(p)));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                        uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 consumer warps
__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float inv_sigm(float x) {  // utils/utils.py:61-74, eps = 1e-5
  x = fminf(fmaxf(x, 0.f), 1.f);
  return logf(fmaxf(x, 1e-5f) / fmaxf(1.f - x, 1e-5f));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t *>(&v);
}
__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  const __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t *>(&v);
}

struct Ring {
  uint8_t *buf;
  uint64_t *full, *empty;
  uint32_t t;     // slots consumed so far
  int gi;         // next program entry
};

// out(16 x N) = A(16 x K, bf16 in shared memory, pitch pa bytes) . W^T + bias, W streamed through the ring.
// epi(col, acc, b0, b1) gets the thread's fragment: rows lane/4 (acc[0..1]) and lane/4+8 (acc[2..3]), columns col, col+1
// and the two bias values.  The A fragments of a 256-wide k-slice live in REGISTERS (64 per thread) for the whole
// slice: re-reading them from shared memory for every slot (8 warps x 8 KB per 33 KB slot) made the kernel
// shared-memory-bandwidth-bound.  K == 256: slots are the n-blocks in order.  K > 256 (N == 256 only): slots are
// k-slice-major, all four n-blocks' accumulators are live.  Two accumulator chains per tile hide the HMMA latency.
__device__ __forceinline__ void load_a_slice(uint32_t (&af)[16][4], const uint8_t *A, int pa, int ks, int lane) {
  const uint8_t *a = A + (lane & 15) * pa + (ks * SLOT_K + (lane >> 4) * 8) * 2;
#pragma unroll
  for (int i = 0; i < 16; ++i) ldsm4(af[i], a + i * 32);
}
__device__ __forceinline__ void slot_mma(float (&c0)[4], float (&c1)[4], const uint32_t (&af)[16][4], const uint8_t *w) {
#pragma unroll
  for (int kk = 0; kk < SLOT_K / 32; ++kk) {
    uint32_t bq[4];
    ldsm4(bq, w + kk * 64);
    mma_bf16(c0, af[2 * kk], bq[0], bq[1]);
    mma_bf16(c1, af[2 * kk + 1], bq[2], bq[3]);
  }
}
template <class Epi>
__device__ __forceinline__ void gemm(const memotr_dec_gemm *prog, Ring &rg, const uint8_t *A, int pa,
                                     const float *__restrict__ bias, int warp, int lane, Epi epi) {
  const memotr_dec_gemm d = prog[rg.gi++];
  const int nb = d.N / SLOT_ROWS, nk = d.K / SLOT_K;
  const int cw = warp * 8 + 2 * (lane & 3);
  const int woff = (warp * 8 + (lane & 7)) * WP + (lane >> 3) * 16;
  uint32_t af[16][4];
  if (nk == 1) {
    load_a_slice(af, A, pa, 0, lane);
    for (int b = 0; b < nb; ++b, ++rg.t) {
      const int col = b * SLOT_ROWS + cw;
      const float b0 = bias ? __ldg(bias + col) : 0.f, b1 = bias ? __ldg(bias + col + 1) : 0.f;   // in flight during the MMAs
      float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
      const int s = rg.t % NSLOT;
      mbar_wait(rg.full + s, (rg.t / NSLOT) & 1);
      slot_mma(c0, c1, af, rg.buf + s * SLOT_BYTES + woff);
      __syncwarp();
      if (lane == 0) mbar_arrive(rg.empty + s);
      c0[0] += c1[0], c0[1] += c1[1], c0[2] += c1[2], c0[3] += c1[3];
      epi(col, c0, b0, b1);
    }
  } else {   // nb == 4 (checked on the host): k-slice-major slots
    float acc[4][2][4];
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[b][0][j] = acc[b][1][j] = 0.f;
    for (int ks = 0; ks < nk; ++ks) {
      load_a_slice(af, A, pa, ks, lane);
#pragma unroll
      for (int b = 0; b < 4; ++b, ++rg.t) {
        const int s = rg.t % NSLOT;
        mbar_wait(rg.full + s, (rg.t / NSLOT) & 1);
        slot_mma(acc[b][0], acc[b][1], af, rg.buf + s * SLOT_BYTES + woff);
        __syncwarp();
        if (lane == 0) mbar_arrive(rg.empty + s);
      }
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int col = b * SLOT_ROWS + cw;
      const float b0 = bias ? __ldg(bias + col) : 0.f, b1 = bias ? __ldg(bias + col + 1) : 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[b][0][j] += acc[b][1][j];
      epi(col, acc[b][0], b0, b1);
    }
  }
}

// LayerNorm over the 256 columns of 16 fp32 rows in `pre` (pitch F0P); warp w takes rows 2w, 2w+1.
// writes x32 (fp32 master), xb (bf16), and optionally sum = bf16(value + qp) into `sumb` (pitch P512)
__device__ __forceinline__ void layer_norm(const float *pre, const float *__restrict__ gamma, const float *__restrict__ beta,
                                           float *x32, uint8_t *xb, const uint8_t *qp, uint8_t *sumb, int warp, int lane) {
  const float4 g0 = ldg_f4(gamma + lane * 8), g1 = ldg_f4(gamma + lane * 8 + 4), b0 = ldg_f4(beta + lane * 8),
               b1 = ldg_f4(beta + lane * 8 + 4);                  // issued before the reductions that hide their latency
  for (int rr = 0; rr < 2; ++rr) {
    const int r = warp * 2 + rr, c0 = lane * 8;
    float v[8];
    const float4 p0 = *reinterpret_cast<const float4 *>(pre + r * F0P + c0), p1 = *reinterpret_cast<const float4 *>(pre + r * F0P + c0 + 4);
    v[0] = p0.x, v[1] = p0.y, v[2] = p0.z, v[3] = p0.w, v[4] = p1.x, v[5] = p1.y, v[6] = p1.z, v[7] = p1.w;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / 256.f);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float dd = v[i] - mean;
      q += dd * dd;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.f / 256.f) + 1e-5f);
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * g[i] + bb[i];
    *reinterpret_cast<float4 *>(x32 + r * C + c0) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4 *>(x32 + r * C + c0 + 4) = make_float4(v[4], v[5], v[6], v[7]);
    *reinterpret_cast<uint4 *>(xb + r * P256 + c0 * 2) = f32x8_to_bf16(v);
    if (sumb) {
      float p[8];
      bf16x8_to_f32(*reinterpret_cast<const uint4 *>(qp + r * P256 + c0 * 2), p);
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] += v[i];
      *reinterpret_cast<uint4 *>(sumb + r * P512 + c0 * 2) = f32x8_to_bf16(p);
    }
  }
}

// skinny head: out[r][j] = dot(A[r][0..255] (bf16, pitch P), W[j][0..255] (bf16)) + bias[j], j < nout; 4 threads per output
__device__ __forceinline__ float head_dot(const uint8_t *A, int pitch, const bf16 *__restrict__ W, int r, int j, int part) {
  float s = 0.f;
  const uint8_t *a = A + r * pitch + part * 128;
  const bf16 *w = W + j * C + part * 64;
#pragma unroll
  for (int k = 0; k < 64; k += 8) {
    float x[8], y[8];
    bf16x8_to_f32(*reinterpret_cast<const uint4 *>(a + k * 2), x);
    bf16x8_to_f32(__ldg(reinterpret_cast<const uint4 *>(w + k)), y);
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(x[i], y[i], s);
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  return s;
}

__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int target) {
  csync();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
    __threadfence();
  }
  csync();
}

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

  for (int lid = 0; lid < P.n_layers; ++lid) {
    const memotr_dec_layer &Lp = P.layers[lid];
    const int n = lid >= P.merge ? nq : nd;               // rows taking part in this layer (deformable_decoder.py:292-297)
    const int par = lid & 1;
    __half *Kh = reinterpret_cast<__half *>(P.kbuf) + (long)par * P.np * C;     // [np][256]
    __half *Vt = reinterpret_cast<__half *>(P.vbuf) + (long)par * C * P.np;     // [256][np]

    // ---- DAB positional query (deformable_decoder.py:82-95): sine embedding of ref * valid_ratio(level 0)
    {
      const float4 sc = ldg_f4(P.vr_scale4);
      const float scl[4] = {sc.x, sc.y, sc.z, sc.w};
      for (int i = tid; i < R * 256; i += 256) {
        const int r = i >> 8, cc = (i >> 6) & 3, j = i & 63;
        const float e = refs[r * 4 + cc] * scl[cc] * 6.283185307179586f / __ldg(P.dim_t + 2 * j);
        *reinterpret_cast<uint32_t *>(bufA + r * P512 + (cc * 128 + 2 * j) * 2) = pack_bf16(sinf(e), cosf(e));
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
    gemm(sprog, rg, xb, P256, Lp.v_b, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {           // v, stored transposed (fp16)
      if (row0 + g < nq) {
        Vt[(long)col * P.np + row0 + g] = __float2half_rn(a[0] + b0);
        Vt[(long)(col + 1) * P.np + row0 + g] = __float2half_rn(a[1] + b1);
      }
      if (row0 + g + 8 < nq) {
        Vt[(long)col * P.np + row0 + g + 8] = __float2half_rn(a[2] + b0);
        Vt[(long)(col + 1) * P.np + row0 + g + 8] = __float2half_rn(a[3] + b1);
      }
    });
    grid_barrier(P.barrier, (unsigned int)(lid + 1) * gridDim.x);

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
    gemm(sprog, rg, bufA, P512, Lp.sao_b, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {         // out_proj + residual
      *reinterpret_cast<float2 *>(f0 + g * F0P + col) = make_float2(a[0] + b0 + x32[g * C + col], a[1] + b1 + x32[g * C + col + 1]);
      *reinterpret_cast<float2 *>(f0 + (g + 8) * F0P + col) =
          make_float2(a[2] + b0 + x32[(g + 8) * C + col], a[3] + b1 + x32[(g + 8) * C + col + 1]);
    });
    csync();
    layer_norm(f0, Lp.n2_g, Lp.n2_b, x32, xb, qp, bufA, warp, lane);                      // norm2; bufA = t1 + query_pos
    csync();

    // ---- cross-attention into the encoder memory (ms_deform_attn.py:88-130)
    gemm(sprog, rg, bufA, P512, Lp.ol_b, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {         // [offsets | logits], fp32
      *reinterpret_cast<float2 *>(f0 + g * F0P + col) = make_float2(a[0] + b0, a[1] + b1);
      *reinterpret_cast<float2 *>(f0 + (g + 8) * F0P + col) = make_float2(a[2] + b0, a[3] + b1);
    });
    csync();
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
    gemm(sprog, rg, bufA, P512, Lp.cao_b, warp, lane, [&](int col, const float (&a)[4], float b0, float b1) {         // output_proj + residual
      *reinterpret_cast<float2 *>(f0 + g * F0P + col) = make_float2(a[0] + b0 + x32[g * C + col], a[1] + b1 + x32[g * C + col + 1]);
      *reinterpret_cast<float2 *>(f0 + (g + 8) * F0P + col) =
          make_float2(a[2] + b0 + x32[(g + 8) * C + col], a[3] + b1 + x32[(g + 8) * C + col + 1]);
    });
    csync();
    layer_norm(f0, Lp.n1_g, Lp.n1_b, x32, xb, nullptr, nullptr, warp, lane);              // norm1
    csync();

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
  }
}

}  // namespace dec
}  // namespace memotr

using namespace memotr;

extern "C" int memotr_decoder_forward(const memotr_dec_params *p, void *stream) {
  MEMOTR_REQUIRE(p && p->prog && p->n_prog > 0 && p->tgt_in && p->ref_in && p->kbuf && p->vbuf && p->barrier && p->dim_t &&
                     p->vr_scale4 && p->valid_ratios,
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
