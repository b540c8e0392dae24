// attention.cu -- multi-head attention over a few hundred queries/keys (head dim 32), fp32 math.
//
// Replaces the math path of nn.MultiheadAttention the reference takes at both call sites (need_weights=True, q/k/v
// distinct tensors => torch/nn/functional.py multi_head_attention_forward: q scaled by 1/sqrt(d) before QK^T, additive
// -inf key-padding mask, softmax over keys, PV):  decoder self-attention over the 300 detect + N track queries
// (models/deformable_decoder.py:245-252) and the QueryUpdater's long-term-memory attention
// (models/query_updater.py:45,125).  The in/out projections are GEMMs (memotr_linear); this kernel is the core.
//
// One CTA per (block of queries, head): K and V of the head are staged once in shared memory as fp32 (K rows padded to
// 33 floats so that lanes reading different keys hit different banks), then each warp walks its queries: scores with
// one key per lane, warp-shuffle max/sum, and the PV product with one output channel per lane.
#include <cstdlib>

#include "common.cuh"

namespace memotr {

__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, s));
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

template <typename T, typename TO>
__global__ void mha32_kernel(const T *__restrict__ Q, int ldq, const T *__restrict__ K, int ldk,
                             const T *__restrict__ V, int ldv, const unsigned char *__restrict__ kpm,
                             TO *__restrict__ O, int ldo, int Nq, int Nk, float scale, int qpb) {
  pdl_grid_sync();
  extern __shared__ float sm[];
  float *Ks = sm;             // [Nk][33]
  float *Vs = Ks + Nk * 33;   // [Nk][32]
  float *Ps = Vs + Nk * 32;   // [warps][Nk]
  const int h = blockIdx.y, nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int idx = threadIdx.x; idx < Nk * 32; idx += blockDim.x) {
    const int j = idx >> 5, d = idx & 31;
    Ks[j * 33 + d] = to_f32<T>(K[(long)j * ldk + h * 32 + d]);
    // padded keys get probability 0; zero their values too so that 0 * (whatever a padding row holds, even NaN) = 0
    Vs[j * 32 + d] = (kpm && kpm[j]) ? 0.f : to_f32<T>(V[(long)j * ldv + h * 32 + d]);
  }
  __syncthreads();
  const int q0 = blockIdx.x * qpb;
  const int q1 = min(q0 + qpb, Nq);
  float *P = Ps + warp * Nk;
  for (int qi = q0 + warp; qi < q1; qi += nw) {
    const float qv = to_f32<T>(Q[(long)qi * ldq + h * 32 + lane]) * scale;
    float qreg[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) qreg[d] = __shfl_sync(0xffffffffu, qv, d);
    float mx = -INFINITY;
    for (int j = lane; j < Nk; j += 32) {
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int d = 0; d < 32; d += 2) {
        s0 = fmaf(qreg[d], Ks[j * 33 + d], s0);
        s1 = fmaf(qreg[d + 1], Ks[j * 33 + d + 1], s1);
      }
      float s = s0 + s1;
      if (kpm && kpm[j]) s = -INFINITY;
      P[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max_f(mx);
    if (mx == -INFINITY) mx = 0.f;  // every key padded (an empty track table): probabilities 0, output 0 (not NaN)
    float sum = 0.f;
    for (int j = lane; j < Nk; j += 32) {
      const float e = expf(P[j] - mx);
      P[j] = e;
      sum += e;
    }
    sum = warp_sum_f(sum);
    __syncwarp();
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;  // four independent chains hide the FMA / LDS latency
    int j = 0;
    for (; j + 4 <= Nk; j += 4) {
      a0 = fmaf(P[j], Vs[j * 32 + lane], a0);
      a1 = fmaf(P[j + 1], Vs[(j + 1) * 32 + lane], a1);
      a2 = fmaf(P[j + 2], Vs[(j + 2) * 32 + lane], a2);
      a3 = fmaf(P[j + 3], Vs[(j + 3) * 32 + lane], a3);
    }
    for (; j < Nk; ++j) a0 = fmaf(P[j], Vs[j * 32 + lane], a0);
    const float acc = (a0 + a1) + (a2 + a3);
    O[(long)qi * ldo + h * 32 + lane] = from_f32<TO>(sum > 0.f ? acc / sum : 0.f);
    __syncwarp();
  }
}


// v2: four queries per warp share every shared-memory operand (one K element feeds 4 FMAs in the score pass, one V
// element 4 FMAs in the PV pass), K is staged transposed ([d][key], lane = key => conflict-free), scores / probabilities
// of the warp's 4 queries sit interleaved as float4 per key, q as float4 per channel (both broadcast LDS.128).
// v1 measured 49 us per decoder self-attention (400x400, 8 heads) -- 2 % of the fp32 FMA peak, most of it re-staging K/V
// for only 16 queries per CTA and one LDS per FMA.
template <typename T, typename TO>
__global__ void __launch_bounds__(256)
mha32_v2_kernel(const T *__restrict__ Q, int ldq, const T *__restrict__ K, int ldk, const T *__restrict__ V, int ldv,
                const unsigned char *__restrict__ kpm, TO *__restrict__ O, int ldo, int Nq, int Nk, int NkP, float scale) {
  pdl_grid_sync();
  extern __shared__ __align__(16) float sm2[];
  float *Kt = sm2;                                   // [32][NkP], NkP odd => conflict-free transposed stores
  float *Vs = Kt + 32 * NkP;                         // [Nk][32]
  float4 *Qs = reinterpret_cast<float4 *>(Vs + (size_t)Nk * 32 + ((32 * NkP + Nk * 32) % 4 ? 4 - (32 * NkP + Nk * 32) % 4 : 0));
  const int NkR = (Nk + 31) & ~31;
  float4 *Ps = Qs + 8 * 32;                          // [8 warps][NkR]
  const int h = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if constexpr (sizeof(T) == 4) {   // fp32 projections (the engine's case): 16-byte loads, 4x fewer staging iterations
    for (int idx = threadIdx.x; idx < Nk * 8; idx += 256) {
      const int j = idx >> 3, d4 = (idx & 7) * 4;
      const float4 kv = __ldg(reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(K) + (long)j * ldk + h * 32 + d4));
      float4 vv = __ldg(reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(V) + (long)j * ldv + h * 32 + d4));
      if (kpm && kpm[j]) vv = make_float4(0.f, 0.f, 0.f, 0.f);  // padded key: probability 0 and value 0 (0 * NaN safe)
      Kt[(d4 + 0) * NkP + j] = kv.x, Kt[(d4 + 1) * NkP + j] = kv.y, Kt[(d4 + 2) * NkP + j] = kv.z, Kt[(d4 + 3) * NkP + j] = kv.w;
      *reinterpret_cast<float4 *>(Vs + j * 32 + d4) = vv;
    }
  } else {
    for (int idx = threadIdx.x; idx < Nk * 32; idx += 256) {
      const int j = idx >> 5, d = idx & 31;
      Kt[d * NkP + j] = to_f32<T>(K[(long)j * ldk + h * 32 + d]);
      Vs[j * 32 + d] = (kpm && kpm[j]) ? 0.f : to_f32<T>(V[(long)j * ldv + h * 32 + d]);
    }
  }
  const int q0 = blockIdx.x * 32 + warp * 4;
  {
    float4 qv;
    qv.x = q0 + 0 < Nq ? to_f32<T>(Q[(long)(q0 + 0) * ldq + h * 32 + lane]) * scale : 0.f;
    qv.y = q0 + 1 < Nq ? to_f32<T>(Q[(long)(q0 + 1) * ldq + h * 32 + lane]) * scale : 0.f;
    qv.z = q0 + 2 < Nq ? to_f32<T>(Q[(long)(q0 + 2) * ldq + h * 32 + lane]) * scale : 0.f;
    qv.w = q0 + 3 < Nq ? to_f32<T>(Q[(long)(q0 + 3) * ldq + h * 32 + lane]) * scale : 0.f;
    Qs[warp * 32 + lane] = qv;
  }
  __syncthreads();
  if (q0 >= Nq) return;
  const float4 *qw = Qs + warp * 32;
  float4 *P = Ps + (size_t)warp * NkR;
  float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  for (int j = lane; j < Nk; j += 32) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int d = 0; d < 32; ++d) {
      const float4 qv = qw[d];
      const float kv = Kt[d * NkP + j];
      s.x = fmaf(qv.x, kv, s.x), s.y = fmaf(qv.y, kv, s.y), s.z = fmaf(qv.z, kv, s.z), s.w = fmaf(qv.w, kv, s.w);
    }
    if (kpm && kpm[j]) s = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    P[j] = s;
    mx.x = fmaxf(mx.x, s.x), mx.y = fmaxf(mx.y, s.y), mx.z = fmaxf(mx.z, s.z), mx.w = fmaxf(mx.w, s.w);
  }
  mx.x = warp_max_f(mx.x), mx.y = warp_max_f(mx.y), mx.z = warp_max_f(mx.z), mx.w = warp_max_f(mx.w);
  if (mx.x == -INFINITY) mx = make_float4(0.f, 0.f, 0.f, 0.f);  // every key padded: output 0 instead of NaN
  float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = lane; j < Nk; j += 32) {
    float4 e = P[j];
    e.x = expf(e.x - mx.x), e.y = expf(e.y - mx.y), e.z = expf(e.z - mx.z), e.w = expf(e.w - mx.w);
    P[j] = e;
    sum.x += e.x, sum.y += e.y, sum.z += e.z, sum.w += e.w;
  }
  sum.x = warp_sum_f(sum.x), sum.y = warp_sum_f(sum.y), sum.z = warp_sum_f(sum.z), sum.w = warp_sum_f(sum.w);
  __syncwarp();
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
  int j = 0;
  for (; j + 2 <= Nk; j += 2) {
    const float4 p0 = P[j], p1 = P[j + 1];
    const float v0 = Vs[j * 32 + lane], v1 = Vs[(j + 1) * 32 + lane];
    a0.x = fmaf(p0.x, v0, a0.x), a0.y = fmaf(p0.y, v0, a0.y), a0.z = fmaf(p0.z, v0, a0.z), a0.w = fmaf(p0.w, v0, a0.w);
    a1.x = fmaf(p1.x, v1, a1.x), a1.y = fmaf(p1.y, v1, a1.y), a1.z = fmaf(p1.z, v1, a1.z), a1.w = fmaf(p1.w, v1, a1.w);
  }
  if (j < Nk) {
    const float4 p0 = P[j];
    const float v0 = Vs[j * 32 + lane];
    a0.x = fmaf(p0.x, v0, a0.x), a0.y = fmaf(p0.y, v0, a0.y), a0.z = fmaf(p0.z, v0, a0.z), a0.w = fmaf(p0.w, v0, a0.w);
  }
  TO *o = O + (long)q0 * ldo + h * 32 + lane;
  const bool live = sum.x > 0.f;  // the mask is per key: all four queries see the same keys
  o[0] = from_f32<TO>(live ? (a0.x + a1.x) / sum.x : 0.f);
  if (q0 + 1 < Nq) o[ldo] = from_f32<TO>(live ? (a0.y + a1.y) / sum.y : 0.f);
  if (q0 + 2 < Nq) o[2 * (long)ldo] = from_f32<TO>(live ? (a0.z + a1.z) / sum.z : 0.f);
  if (q0 + 3 < Nq) o[3 * (long)ldo] = from_f32<TO>(live ? (a0.w + a1.w) / sum.w : 0.f);
}

}  // namespace memotr

using namespace memotr;

extern "C" int memotr_mha(const void *Q, int ldq, const void *K, int ldk, const void *V, int ldv,
                          const unsigned char *key_padding_mask, void *O, int ldo, int Nq, int Nk, int n_heads,
                          int head_dim, int in_dtype, int out_dtype, void *stream) {
  MEMOTR_REQUIRE(Q && K && V && O && Nq >= 0 && Nk > 0 && n_heads > 0, "mha: bad arguments");
  MEMOTR_REQUIRE(head_dim == 32, "mha: only head_dim == 32 is implemented (got %d)", head_dim);
  MEMOTR_REQUIRE((in_dtype == MEMOTR_F32 || in_dtype == MEMOTR_BF16) && (out_dtype == MEMOTR_F32 || out_dtype == MEMOTR_BF16),
                 "mha: dtypes must be f32 or bf16");
  if (Nq == 0) return MEMOTR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  using bf = __nv_bfloat16;
  {  // v2 (4 queries per warp) whenever K^T, V and the score tiles of 8 warps fit in shared memory (Nk <~ 560)
    const int NkP = ((Nk + 31) & ~31) + 1, NkR = (Nk + 31) & ~31;
    size_t words = (size_t)32 * NkP + (size_t)Nk * 32;
    words += (words % 4) ? 4 - words % 4 : 0;
    const size_t smem2 = (words + 8 * 32 * 4 + (size_t)8 * NkR * 4) * sizeof(float);
    const char *force = getenv("MEMOTR_MHA_KERNEL");
    const bool al = in_dtype != MEMOTR_F32 || (ldk % 4 == 0 && ldv % 4 == 0 && aligned16(K) && aligned16(V));
    if (al && smem2 <= 227 * 1024 && !(force && force[0] == 'v' && force[1] == '1')) {
      static bool attr2 = false;
      if (!attr2) {
        cudaError_t e = cudaFuncSetAttribute(mha32_v2_kernel<float, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(mha32_v2_kernel<float, bf>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(mha32_v2_kernel<bf, bf>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(mha32_v2_kernel<bf, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return fail(MEMOTR_ECUDA, "mha: %s", cudaGetErrorString(e));
        attr2 = true;
      }
      dim3 grid2(ceil_div(Nq, 32), n_heads);
      const float scale2 = 1.0f / sqrtf((float)head_dim);
#define MHA2_LAUNCH(TI, TO_)                                                                                          \
  MEMOTR_LAUNCH((mha32_v2_kernel<TI, TO_>), grid2, 256, smem2, st, (const TI *)Q, ldq, (const TI *)K, ldk, (const TI *)V, \
                ldv, key_padding_mask, (TO_ *)O, ldo, Nq, Nk, NkP, scale2)
      if (in_dtype == MEMOTR_F32 && out_dtype == MEMOTR_F32) MHA2_LAUNCH(float, float);
      else if (in_dtype == MEMOTR_F32) MHA2_LAUNCH(float, bf);
      else if (out_dtype == MEMOTR_BF16) MHA2_LAUNCH(bf, bf);
      else MHA2_LAUNCH(bf, float);
#undef MHA2_LAUNCH
      return check_launch("mha_v2");
    }
  }
  int threads = 256;
  size_t smem = (size_t)Nk * (33 + 32 + threads / 32) * sizeof(float);
  if (smem > 220 * 1024) {
    threads = 128;
    smem = (size_t)Nk * (33 + 32 + threads / 32) * sizeof(float);
  }
  if (smem > 227 * 1024) return fail(MEMOTR_ENOSYS, "mha: %d keys do not fit in shared memory", Nk);
  const int qpb = 16;  // 2 queries per warp: (Nq/16) x heads CTAs keep all SMs busy at Nq = 300..800
  dim3 grid(ceil_div(Nq, qpb), n_heads);
  const float scale = 1.0f / sqrtf((float)head_dim);
  static bool attr_set = false;  // one-time, idempotent: allow up to 227 KB of dynamic shared memory
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mha32_kernel<float, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mha32_kernel<float, bf>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mha32_kernel<bf, bf>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mha32_kernel<bf, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail(MEMOTR_ECUDA, "mha: %s", cudaGetErrorString(e));
    attr_set = true;
  }
#define MHA_LAUNCH(TI, TO_)                                                                                          \
  MEMOTR_LAUNCH((mha32_kernel<TI, TO_>), grid, threads, smem, st, (const TI *)Q, ldq, (const TI *)K, ldk, (const TI *)V, ldv,      \
                                                     key_padding_mask, (TO_ *)O, ldo, Nq, Nk, scale, qpb)
  if (in_dtype == MEMOTR_F32 && out_dtype == MEMOTR_F32) MHA_LAUNCH(float, float);
  else if (in_dtype == MEMOTR_F32) MHA_LAUNCH(float, bf);
  else if (out_dtype == MEMOTR_BF16) MHA_LAUNCH(bf, bf);
  else MHA_LAUNCH(bf, float);
#undef MHA_LAUNCH
  return check_launch("mha");
}
