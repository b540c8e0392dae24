// gemm_simt.cu -- C = epilogue(A . W^T + bias) on the CUDA cores (fp32 accumulate), any M/N/K, fp32 or bf16 I/O.
//
// Stands in for the nn.Linear calls on the hot path: value_proj / sampling_offsets /
// attention_weights / output_proj (models/ops/modules/ms_deform_attn.py:104-129), the FFNs
// (models/deformable_encoder.py:97-107, deformable_decoder.py:263-273, models/ffn.py:15-25), the MLPs (models/mlp.py),
// and the in/out projections of nn.MultiheadAttention (deformable_decoder.py:245-252, query_updater.py:125).
//
// This is the exact-fp32 path (TF32 is off in the reference, main.py:96-97, so an fp32-accurate GEMM is what parity
// at 1e-4 needs) and the any-shape fallback for the bf16 engine (N not a multiple of 64, e.g. the 4-wide box head and
// the 1-wide class head).  The bf16 tensor-core path is gemm_tc.cu; memotr_linear() picks between them.
//
// Tiling: BM x BN output tile per 256-thread CTA, BK = 16, each thread a TM x TN register block, operands staged
// through shared memory transposed to [k][m] so the inner product reads are conflict-free broadcasts/vectors.
#include "common.cuh"

namespace memotr {

template <typename TA, typename TC, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TA *__restrict__ A, int lda, const TA *__restrict__ W, int ldw, TC *__restrict__ C, int ldc,
                 int M, int N, int K, Epilogue ep) {
  pdl_grid_sync();
  constexpr int BK = 16;
  static_assert((BM / TM) * (BN / TN) == 256, "256 threads per CTA");
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // stage A tile (BM x BK) and W tile (BN x BK): element (r, kk) -> smem[kk][r]
    for (int e = tid; e < BM * BK; e += 256) {
      const int r = e / BK, kk = e % BK;
      const int gm = m0 + r, gk = k0 + kk;
      As[kk][r] = (gm < M && gk < K) ? to_f32<TA>(A[(long)gm * lda + gk]) : 0.f;
    }
    for (int e = tid; e < BN * BK; e += 256) {
      const int r = e / BK, kk = e % BK;
      const int gn = n0 + r, gk = k0 + kk;
      Ws[kk][r] = (gn < N && gk < K) ? to_f32<TA>(W[(long)gn * ldw + gk]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], w[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) w[j] = Ws[kk][j * (BN / TN) + tx];  // strided columns: conflict-free
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + ty * TM + i;
    if (gm >= M) continue;
    const bool zero_row = ep.rowzero && ep.rowzero[gm];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + j * (BN / TN) + tx;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (ep.bias) v += ep.bias[gn];
      if (ep.act == ACT_RELU) v = fmaxf(v, 0.f);
      if (ep.act == ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
      if (ep.mul) v *= to_f32<TA>(((const TA *)ep.mul)[(long)gm * ep.ldmul + gn]);
      if (ep.add) v += to_f32<TA>(((const TA *)ep.add)[(long)gm * ep.ldadd + gn]);
      if (zero_row) v = 0.f;
      C[(long)gm * ldc + gn] = from_f32<TC>(v);
    }
  }
}


// Skinny outputs (N <= 8: the 4-wide box-delta layer and the 1..8-wide class heads, memotr.py:153-154): one warp per
// row of A, lanes stride over K, N running sums reduced with xor-shuffles.  The tiled kernels would leave 15/16 of a
// 64-wide tile empty and launch only M/32 CTAs.
template <typename TA>
__device__ __forceinline__ void load8_as_f32(const TA *p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8_as_f32<float>(const float *p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p + 4));
  v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8_as_f32<__nv_bfloat16>(const __nv_bfloat16 *p, float (&v)[8]) {
  bf16x8_to_f32(__ldg(reinterpret_cast<const uint4 *>(p)), v);
}

template <typename TA, typename TC>
__global__ void __launch_bounds__(256)
gemv_rows_kernel(const TA *__restrict__ A, int lda, const TA *__restrict__ W, int ldw, TC *__restrict__ C, int ldc, int M,
                 int N, int K, Epilogue ep, int vec) {
  pdl_grid_sync();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  float acc[8];
#pragma unroll
  for (int n = 0; n < 8; ++n) acc[n] = 0.f;
  if (vec) {  // K % 8 == 0 and 16-byte aligned rows: 8 elements per lane per step, all loads issued before the math
    for (int k = lane * 8; k < K; k += 256) {
      float a[8];
      load8_as_f32<TA>(A + (long)row * lda + k, a);
#pragma unroll
      for (int n = 0; n < 8; ++n)
        if (n < N) {
          float w[8];
          load8_as_f32<TA>(W + (long)n * ldw + k, w);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[n] = fmaf(a[i], w[i], acc[n]);
        }
    }
  } else {
    for (int k = lane; k < K; k += 32) {
      const float a = to_f32<TA>(A[(long)row * lda + k]);
#pragma unroll
      for (int n = 0; n < 8; ++n)
        if (n < N) acc[n] = fmaf(a, to_f32<TA>(W[(long)n * ldw + k]), acc[n]);
    }
  }
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], s);
  if (lane < N) {
    float v = 0.f;
#pragma unroll
    for (int n = 0; n < 8; ++n)
      if (lane == n) v = acc[n];
    if (ep.bias) v += ep.bias[lane];
    if (ep.act == ACT_RELU) v = fmaxf(v, 0.f);
    if (ep.act == ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
    if (ep.mul) v *= to_f32<TA>(((const TA *)ep.mul)[(long)row * ep.ldmul + lane]);
    if (ep.add) v += to_f32<TA>(((const TA *)ep.add)[(long)row * ep.ldadd + lane]);
    if (ep.rowzero && ep.rowzero[row]) v = 0.f;
    C[(long)row * ldc + lane] = from_f32<TC>(v);
  }
}

template <typename TA, typename TC>
static int launch_simt(const void *A, int lda, const void *W, int ldw, void *C, int ldc, int M, int N, int K,
                       const Epilogue &ep, cudaStream_t st) {
  if (N <= 8) {
    const int vec = (K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && aligned16(A) && aligned16(W)) ? 1 : 0;
    MEMOTR_LAUNCH((gemv_rows_kernel<TA, TC>), ceil_div(M, 8), 256, 0, st, (const TA *)A, lda, (const TA *)W, ldw, (TC *)C, ldc, M, N,
                                                             K, ep, vec);
    return check_launch("gemv_rows");
  }
  // big problems: 128x128 tiles (8x8 per thread); small ones (decoder / updater rows): 32x64 tiles for more CTAs
  if ((long)M * N >= 128L * 128 * kNumSMs) {
    dim3 grid(ceil_div(N, 128), ceil_div(M, 128));
    MEMOTR_LAUNCH((gemm_simt_kernel<TA, TC, 128, 128, 8, 8>), grid, 256, 0, st, (const TA *)A, lda, (const TA *)W, ldw, (TC *)C,
                                                                    ldc, M, N, K, ep);
  } else {
    dim3 grid(ceil_div(N, 64), ceil_div(M, 32));
    MEMOTR_LAUNCH((gemm_simt_kernel<TA, TC, 32, 64, 2, 4>), grid, 256, 0, st, (const TA *)A, lda, (const TA *)W, ldw, (TC *)C, ldc,
                                                                  M, N, K, ep);
  }
  return check_launch("gemm_simt");
}

int linear_tc_bf16(const void *A, int lda, const void *W, int ldw, void *C, int ldc, int c_dtype, int M, int N, int K,
                   const Epilogue &ep, cudaStream_t st);  // gemm_tc.cu
bool linear_tc_supported(int lda, int ldw, int ldc, int c_dtype, int M, int N, int K, const void *A, const void *W,
                         const void *C);
int linear_mma_bf16(const void *A, int lda, const void *W, int ldw, void *C, int ldc, int c_dtype, int M, int N, int K,
                    const Epilogue &ep, cudaStream_t st);  // gemm_mma.cu
bool linear_mma_supported(int lda, int ldw, int ldc, int c_dtype, int M, int N, int K, const void *A, const void *W,
                          const void *C);

}  // namespace memotr

using namespace memotr;

extern "C" int memotr_linear(const void *A, int lda, const void *W, int ldw, const float *bias, const void *mul,
                             int ldmul, const void *add, int ldadd, const unsigned char *rowzero, void *C, int ldc,
                             int M, int N, int K, int ab_dtype, int c_dtype, int act, int path, void *stream) {
  MEMOTR_REQUIRE(M >= 0 && N > 0 && K > 0, "linear: bad sizes M=%d N=%d K=%d", M, N, K);
  if (M == 0) return MEMOTR_OK;
  MEMOTR_REQUIRE(A && W && C, "linear: null pointer");
  MEMOTR_REQUIRE(lda >= K && ldw >= K && ldc >= N, "linear: leading dimension too small");
  MEMOTR_REQUIRE(act >= 0 && act <= 2, "linear: unknown activation %d", act);
  MEMOTR_REQUIRE(ab_dtype == MEMOTR_F32 || ab_dtype == MEMOTR_BF16, "linear: A/W dtype must be f32 or bf16");
  MEMOTR_REQUIRE(c_dtype == MEMOTR_F32 || ((c_dtype == MEMOTR_BF16 || c_dtype == MEMOTR_F16) && ab_dtype == MEMOTR_BF16),
                 "linear: output dtype must be f32, or bf16/fp16 with bf16 inputs");
  cudaStream_t st = (cudaStream_t)stream;
  Epilogue ep{bias, mul, add, rowzero, ldmul, ldadd, act};
  // path: 0 = auto, 1 = force CUDA-core kernel, 2 = force tcgen05 kernel, 3 = force the few-rows mma.sync kernel
  if (ab_dtype == MEMOTR_BF16 && (path == 3 || (path == 0 && M <= 1024))) {
    const bool ok = linear_mma_supported(lda, ldw, ldc, c_dtype, M, N, K, A, W, C);
    if (ok) return linear_mma_bf16(A, lda, W, ldw, C, ldc, c_dtype, M, N, K, ep, st);
    if (path == 3) return fail(MEMOTR_ENOSYS, "linear: shape M=%d N=%d K=%d not supported by the mma.sync path", M, N, K);
  }
  if (ab_dtype == MEMOTR_BF16 && path != 1) {
    const bool ok = linear_tc_supported(lda, ldw, ldc, c_dtype, M, N, K, A, W, C);
    if (ok) return linear_tc_bf16(A, lda, W, ldw, C, ldc, c_dtype, M, N, K, ep, st);
    if (path == 2) return fail(MEMOTR_ENOSYS, "linear: shape M=%d N=%d K=%d not supported by the tcgen05 path", M, N, K);
  }
  MEMOTR_REQUIRE(c_dtype != MEMOTR_F16, "linear: fp16 output is only produced by the tensor-core path");
  if (ab_dtype == MEMOTR_F32) return launch_simt<float, float>(A, lda, W, ldw, C, ldc, M, N, K, ep, st);
  if (c_dtype == MEMOTR_F32) return launch_simt<__nv_bfloat16, float>(A, lda, W, ldw, C, ldc, M, N, K, ep, st);
  return launch_simt<__nv_bfloat16, __nv_bfloat16>(A, lda, W, ldw, C, ldc, M, N, K, ep, st);
}
