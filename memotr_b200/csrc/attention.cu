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
    Vs[j * 32 + d] = to_f32<T>(V[(long)j * ldv + h * 32 + d]);
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
    O[(long)qi * ldo + h * 32 + lane] = from_f32<TO>(acc / sum);
    __syncwarp();
  }
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
  using bf = __nv_bfloat16;
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
