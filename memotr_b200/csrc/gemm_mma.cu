// gemm_mma.cu -- latency-optimised bf16 GEMM for FEW rows (decoder / query-updater: M <= ~1000 query rows).
//
// Same contract as gemm_tc.cu (C = epilogue(A . W^T + bias), A (M,K), W (N,K) K-major, fp32 accumulate).  The ~150
// dense layers of the decoder and the query updater each see at most 400 rows: a tcgen05/TMA/TMEM kernel spends its
// time in set-up (descriptor prefetch, barrier init, TMEM allocation, TMA round trips, TMA-store drain: 6-10 us per
// launch measured, profiles/r01_launches_bench_steps2_v2_warm.csv) on 8 CTAs.  Here a 32 x 64 output tile per 128-thread
// CTA (13 x N/64 CTAs for 400 rows) streams its operands with cp.async (double-buffered 128-wide K slabs), feeds
// mma.sync.m16n8k16 from ldmatrix fragments and stores straight from registers -- no allocation, no barriers beyond
// __syncthreads, one global round trip of latency.  Throughput is irrelevant at this size (0.05-0.4 GFLOP per launch);
// the large-M GEMMs stay on tcgen05.
#include "common.cuh"

namespace memotr {
namespace mma {

constexpr int BM = 32, BN = 64, BKS = 128;         // tile and K slab
constexpr int LDS_ROW = BKS * 2 + 16;              // bytes per smem row: 256 B of bf16 + 16 B pad (ldmatrix conflict-free)
constexpr int A_BYTES = BM * LDS_ROW, W_BYTES = BN * LDS_ROW, STAGE = A_BYTES + W_BYTES;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool valid) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void *smem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(s));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <typename TC>
__global__ void __launch_bounds__(128)
gemm_mma_kernel(const __nv_bfloat16 *__restrict__ A, int lda, const __nv_bfloat16 *__restrict__ W, int ldw,
                TC *__restrict__ C, int ldc, int M, int N, int K, Epilogue ep) {
  pdl_grid_sync();
  extern __shared__ __align__(16) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int wm = (warp >> 1) * 16, wn = (warp & 1) * 32;   // warp tile: 16 rows x 32 cols
  const int nslab = K / BKS;

  auto issue = [&](int slab, int stage) {
    uint8_t *sa = smem + stage * STAGE, *sw = sa + A_BYTES;
    const int k0 = slab * BKS;
    for (int i = tid; i < BM * 16; i += 128) {          // A: 32 rows x 16 chunks of 16 B
      const int r = i >> 4, c = i & 15;
      const bool ok = m0 + r < M;
      cp_async16(sa + r * LDS_ROW + c * 16, A + (long)(ok ? m0 + r : 0) * lda + k0 + c * 8, ok);
    }
    for (int i = tid; i < BN * 16; i += 128) {          // W: 64 rows x 16 chunks
      const int r = i >> 4, c = i & 15;
      cp_async16(sw + r * LDS_ROW + c * 16, W + (long)(n0 + r) * ldw + k0 + c * 8, true);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  issue(0, 0);
  for (int s = 0; s < nslab; ++s) {
    if (s + 1 < nslab) {
      issue(s + 1, (s + 1) & 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const uint8_t *sa = smem + (s & 1) * STAGE, *sw = sa + A_BYTES;
#pragma unroll
    for (int kk = 0; kk < BKS; kk += 16) {
      uint32_t a[4], b01[4], b23[4];
      // A fragment: lanes 0-15 -> rows 0-15 at k, lanes 16-31 -> rows 0-15 at k+8
      ldmatrix_x4(a, sa + (wm + (lane & 15)) * LDS_ROW + (kk + (lane >> 4) * 8) * 2);
      // B fragments for two pairs of n-tiles: lanes 0-7 n 0-7 @k, 8-15 n 0-7 @k+8, 16-23 n 8-15 @k, 24-31 n 8-15 @k+8
      const int brow = (lane & 7) + ((lane >> 4) << 3), bcol = (kk + ((lane >> 3) & 1) * 8) * 2;
      ldmatrix_x4(b01, sw + (wn + brow) * LDS_ROW + bcol);
      ldmatrix_x4(b23, sw + (wn + 16 + brow) * LDS_ROW + bcol);
      mma_bf16(acc[0], a, b01[0], b01[1]);
      mma_bf16(acc[1], a, b01[2], b01[3]);
      mma_bf16(acc[2], a, b23[0], b23[1]);
      mma_bf16(acc[3], a, b23[2], b23[3]);
    }
    __syncthreads();
  }

  // epilogue straight from the accumulator fragments: thread holds rows (lane/4, lane/4+8), column pairs 2*(lane%4)
  const __nv_bfloat16 *mul = (const __nv_bfloat16 *)ep.mul, *add = (const __nv_bfloat16 *)ep.add;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int row = m0 + wm + (lane >> 2) + half * 8;
    if (row >= M) continue;
    const bool zero_row = ep.rowzero && ep.rowzero[row];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int col = n0 + wn + t * 8 + 2 * (lane & 3);
      float v0 = acc[t][2 * half], v1 = acc[t][2 * half + 1];
      if (ep.bias) v0 += __ldg(ep.bias + col), v1 += __ldg(ep.bias + col + 1);
      if (ep.act == ACT_RELU) v0 = fmaxf(v0, 0.f), v1 = fmaxf(v1, 0.f);
      if (ep.act == ACT_SIGMOID) v0 = 1.f / (1.f + __expf(-v0)), v1 = 1.f / (1.f + __expf(-v1));
      if (mul) {
        const float2 m2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(mul + (long)row * ep.ldmul + col));
        v0 *= m2.x, v1 *= m2.y;
      }
      if (add) {
        const float2 a2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(add + (long)row * ep.ldadd + col));
        v0 += a2.x, v1 += a2.y;
      }
      if (zero_row) v0 = v1 = 0.f;
      TC *dst = C + (long)row * ldc + col;
      if constexpr (sizeof(TC) == 4) *reinterpret_cast<float2 *>(dst) = make_float2(v0, v1);
      else *reinterpret_cast<__nv_bfloat162 *>(dst) = __floats2bfloat162_rn(v0, v1);
    }
  }
}

}  // namespace mma

bool linear_mma_supported(int lda, int ldw, int ldc, int c_dtype, int M, int N, int K, const void *A, const void *W,
                          const void *C) {
  (void)M;
  return (c_dtype == MEMOTR_F32 || c_dtype == MEMOTR_BF16) && N % mma::BN == 0 && K % mma::BKS == 0 && lda % 8 == 0 &&
         ldw % 8 == 0 && ldc % 2 == 0 && aligned16(A) && aligned16(W) && ((reinterpret_cast<uintptr_t>(C) & 7u) == 0);
}

int linear_mma_bf16(const void *A, int lda, const void *W, int ldw, void *C, int ldc, int c_dtype, int M, int N, int K,
                    const Epilogue &ep, cudaStream_t st) {
  if ((ep.mul && ep.ldmul % 2) || (ep.add && ep.ldadd % 2)) return fail(MEMOTR_EINVAL, "linear(mma): odd mul/add stride");
  dim3 grid(N / mma::BN, ceil_div(M, mma::BM));
  const size_t smem = 2 * mma::STAGE;
  using bf = __nv_bfloat16;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mma::gemm_mma_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(mma::gemm_mma_kernel<bf>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(MEMOTR_ECUDA, "linear(mma): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  if (c_dtype == MEMOTR_F32)
    MEMOTR_LAUNCH((mma::gemm_mma_kernel<float>), grid, 128, smem, st, (const bf *)A, lda, (const bf *)W, ldw, (float *)C,
                  ldc, M, N, K, ep);
  else
    MEMOTR_LAUNCH((mma::gemm_mma_kernel<bf>), grid, 128, smem, st, (const bf *)A, lda, (const bf *)W, ldw, (bf *)C, ldc, M,
                  N, K, ep);
  return check_launch("gemm_mma");
}

}  // namespace memotr
