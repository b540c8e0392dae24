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


// ---- 4-CTA cluster helpers (decoder_cluster.cu, updater_cluster.cu) ---------------------------------------------
namespace cl {
constexpr int CS = 4;                              // CTAs per row block

__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cl_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_cl_f32x2(uint32_t addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

struct Xchg {                 // one hand-off between the four CTAs of a cluster (consumer warps only)
  uint32_t bar_remote[CS];    // shared::cluster addresses of every CTA's exchange barrier
  uint64_t *bar;              // the local one
  uint32_t phase;
  __device__ __forceinline__ void sync(int lane) {
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int p = 0; p < CS; ++p)
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_remote[p]) : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\tX_LOOP:\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t@p bra X_DONE;\n\tbra X_LOOP;\n\tX_DONE:\n\t}" ::"r"(
            s32(bar)),
        "r"(phase & 1)
        : "memory");
    ++phase;
  }
};


// Self-attention of a 16-row block for the two heads (2 rk, 2 rk + 1) of one CTA of a 4-CTA cluster: q (fp16, already scaled,
// local, columns hl * 32 of `qbuf`), keys Kh (np, 256) fp16 and values Vt (256, np) fp16 (transposed) in global memory
// (written by all row blocks before a grid barrier), n keys, optional key-padding mask.  Warp = (head, quarter of the
// 64-key blocks); online softmax per warp, the four partial (m, l, O) of a head are combined through `scratch` (>= 18.5 KB)
// and the normalised output (bf16) is written at (row, 64 rk + hl * 32 + d) of the buffer at byte offset `out_off`
// (row pitch `out_pitch`) in the shared memory of ALL four CTAs.
__device__ __forceinline__ void attention_2heads(const uint8_t *bufB, const __half *Kh, const __half *Vt, int np, int n,
                                                 const unsigned char *pad, int rk, float *f0, const uint32_t (&peer)[CS],
                                                 int out_off, int out_pitch, int warp, int lane, int tid) {
  const int g = lane >> 2, c = lane & 3;

    const int hl = warp >> 2, part = warp & 3, h = 2 * rk + hl;
    const uint4 qv0 = *reinterpret_cast<const uint4 *>(bufB + g * P256 + (hl * 32 + 8 * c) * 2);
    const uint4 qv1 = *reinterpret_cast<const uint4 *>(bufB + (g + 8) * P256 + (hl * 32 + 8 * c) * 2);
    float o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    const int nblk = (n + 63) / 64;
    for (int blk = part; blk < nblk; blk += 4) {
      const int kb = blk * 64;
      uint4 kr[8], vr[4][2];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int key = kb + 16 * (g >> 1) + 4 * (j >> 1) + 2 * (j & 1) + (g & 1);
        kr[j] = __ldcg(reinterpret_cast<const uint4 *>(Kh + (long)key * C + h * 32 + 8 * c));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __half *vp = Vt + (long)(h * 32 + 8 * i + g) * np + kb + 16 * c;
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
      float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = kb + 16 * c + 4 * (j >> 1) + 2 * (j & 1) + e;
          const bool dead = key >= n || (pad && pad[key]);
          if (dead) s[j][e] = -INFINITY, s[j][2 + e] = -INFINITY;
          bm0 = fmaxf(bm0, s[j][e]), bm1 = fmaxf(bm1, s[j][2 + e]);
        }
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)), bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)), bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
      const float n0 = fmaxf(m0, bm0), n1 = fmaxf(m1, bm1);
      const float u0 = n0 == -INFINITY ? 0.f : n0, u1 = n1 == -INFINITY ? 0.f : n1;
      const float f0s = __expf(m0 - u0), f1s = __expf(m1 - u1);
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
        for (int st = 0; st < 4; ++st)
          mma_f16(o[i], pa[2 * st][0], pa[2 * st][1], pa[2 * st + 1][0], pa[2 * st + 1][1], vv[2 * st], vv[2 * st + 1]);
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1), l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1), l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    // partial (m, l, O) of this warp -> f0 scratch: [warp][row][36] floats (32 O columns, m, l)
    float *sc = f0 + warp * (R * 36);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      *reinterpret_cast<float2 *>(sc + g * 36 + 8 * i + 2 * c) = make_float2(o[i][0], o[i][1]);
      *reinterpret_cast<float2 *>(sc + (g + 8) * 36 + 8 * i + 2 * c) = make_float2(o[i][2], o[i][3]);
    }
    if (c == 0) sc[g * 36 + 32] = m0, sc[g * 36 + 33] = l0, sc[(g + 8) * 36 + 32] = m1, sc[(g + 8) * 36 + 33] = l1;
    csync();
    // combine the four key quarters: 2 heads x 16 rows x 32 columns = 1024 outputs, two adjacent columns per thread
    for (int oi = tid; oi < 2 * R * 16; oi += 256) {
      const int hh = oi / (R * 16), r = (oi / 16) % R, d2 = (oi % 16) * 2;
      float M = -INFINITY;
#pragma unroll
      for (int p = 0; p < 4; ++p) M = fmaxf(M, f0[(hh * 4 + p) * (R * 36) + r * 36 + 32]);
      const float U = M == -INFINITY ? 0.f : M;
      float num0 = 0.f, num1 = 0.f, den = 0.f;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float *pp = f0 + (hh * 4 + p) * (R * 36) + r * 36;
        const float w = __expf(pp[32] - U);
        num0 += w * pp[d2], num1 += w * pp[d2 + 1], den += w * pp[33];
      }
      const float inv = den > 0.f ? 1.f / den : 0.f;
      {
      const uint32_t o = out_off + r * out_pitch + (64 * rk + hh * 32 + d2) * 2, v = pack_bf16(num0 * inv, num1 * inv);
#pragma unroll
      for (int p = 0; p < CS; ++p) st_cl_u32(peer[p] + o, v);
    }
    }
  }

}  // namespace cl

}  // namespace dec
}  // namespace memotr
