// decoder_common.cuh -- device helpers shared by the fused decoder kernels (decoder_fused.cu: one CTA per 16-row block;
// decoder_cluster.cu: a 4-CTA cluster per 16-row block): PTX wrappers (mbarrier, bulk copy, ldmatrix, mma.sync), the
// register-resident-A GEMM over a weight ring, LayerNorm, the skinny heads, and the grid barrier.
#pragma once
#include "common.cuh"

namespace memotr {
namespace dec {

using bf16 = __nv_bfloat16;
constexpr int R = 16, C = 256, NCW = 8, NTHREADS = (NCW + 1) * 32;
constexpr int SLOT_ROWS = 64, SLOT_K = 256, WP = SLOT_K * 2 + 16, SLOT_BYTES = SLOT_ROWS * WP, NSLOT = 3;
constexpr int P256 = 256 * 2 + 16, P512 = 512 * 2 + 16, P1024 = 1024 * 2 + 16;  // bf16 row pitches (ldmatrix conflict-free)
constexpr int F0P = 512;                                                         // fp32 scratch pitch (floats)

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
  } else {   // nb <= 4 (checked on the host): k-slice-major slots
    float acc[4][2][4];
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[b][0][j] = acc[b][1][j] = 0.f;
    for (int ks = 0; ks < nk; ++ks) {
      load_a_slice(af, A, pa, ks, lane);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (b >= nb) break;
        const int s = rg.t % NSLOT;
        mbar_wait(rg.full + s, (rg.t / NSLOT) & 1);
        slot_mma(acc[b][0], acc[b][1], af, rg.buf + s * SLOT_BYTES + woff);
        __syncwarp();
        if (lane == 0) mbar_arrive(rg.empty + s);
        ++rg.t;
      }
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      if (b >= nb) break;
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


}  // namespace dec
}  // namespace memotr
