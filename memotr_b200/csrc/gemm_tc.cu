// gemm_tc.cu -- bf16 tensor-core GEMM for sm_100a:  C = epilogue(A . W^T + bias),  A (M,K) and W (N,K) both K-major.
//
// Stands in for the dense layers the reference runs as fp32 cuBLAS SGEMM:
// the MSDeformAttn projections (models/ops/modules/ms_deform_attn.py:104-129), the encoder/decoder FFNs
// (models/deformable_encoder.py:97-107, models/deformable_decoder.py:263-273), the MLPs and the MHA in/out
// projections (models/deformable_decoder.py:245-252, models/query_updater.py:109-132).
//
// Blackwell design (one CTA = one 128 x BN output tile, 6 warps, warp-specialised):
//   warp 0   TMA producer: cp.async.bulk.tensor.2d of a 128x64 A box and a BNx64 W box per k-block into a 3-stage
//            shared-memory ring (SWIZZLE_128B), completion counted on an mbarrier (expect_tx).
//   warp 1   allocates TMEM, then one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16,
//            bf16 x bf16 -> fp32 accumulator in TMEM), 4 per k-block; tcgen05.commit frees the smem stage and, after the
//            last k-block, signals the epilogue.
//   warps 2-5 epilogue: tcgen05.ld 32x32b (one accumulator row per thread, 32 columns at a time) -> bias / ReLU /
//            sigmoid / multiplier / residual / padding-mask -> 16-byte st.shared into the (by then dead) pipeline
//            stages, laid out as 128-row x 128-byte panels with the 128B swizzle -> one TMA store per panel
//            (cp.async.bulk.tensor.2d.global.shared::cta).  v1 wrote one row per thread straight to global memory
//            (32 scattered 16-byte sectors per store instruction); the panels make every global write a full line.
// Rows beyond M are zero-filled by TMA on load and clipped by TMA on store.  96 KB (BN=128) or 72 KB (BN=64) of shared memory
// and BN TMEM columns per CTA, so 2-3 CTAs are resident per SM and one tile's epilogue overlaps another's main loop.
#include <type_traits>

#include "tc_common.cuh"

namespace memotr {

namespace tc {

constexpr int STAGES = 3;
constexpr int A_BYTES = BM * BK * 2;  // 16 KB

template <int BN>
struct Smem {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 128 + 1024;  // + barriers/tmem slot + slack for 1024 B alignment
};

template <int BN, typename TC>
__global__ void __launch_bounds__(192)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmC, int M, int N, int K, Epilogue ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  using SM = Smem<BN>;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + SM::BAR_OFFSET);
  uint64_t *empty = full + STAGES;
  uint64_t *tmem_full = empty + STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blk = blockIdx.y, n_blk = blockIdx.x;
  const int num_k = K / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // whole warp: allocate BN TMEM columns, publish the base address through shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_grid_sync();  // everything above (descriptor prefetch, barrier init, TMEM allocation) overlapped the previous kernel

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(empty + s, ph ^ 1);
        mbar_expect_tx(full + s, SM::STAGE_BYTES);
        uint8_t *sa = smem + s * SM::STAGE_BYTES;
        tma_load_2d(sa, &tmA, full + s, kb * BK, m_blk * BM);
        tma_load_2d(sa + A_BYTES, &tmW, full + s, kb * BK, n_blk * BN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ep.in_f16 ? umma_idesc_f16(BN) : umma_idesc(BN);
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(full + s, ph);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + s * SM::STAGE_BYTES);
        const uint64_t adesc = umma_desc(sa), bdesc = umma_desc(sa + A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)  // advance 16 bf16 = 32 B = 2 x 16-byte units inside the swizzle atom
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
        umma_commit(empty + s);  // implicit fence::before_thread_sync; frees the stage when the MMAs have read it
      }
      umma_commit(tmem_full);
    }
  } else {
    // ---- epilogue: warps 2..5 own TMEM lane quarters (warp % 4) ----
    mbar_wait(tmem_full, 0);   // all MMAs have completed => every smem stage has been consumed and may be reused
    tcgen05_fence_after();
    const int quarter = warp & 3;
    const int r_in = quarter * 32 + lane;  // row inside the tile == TMEM lane
    const int row = m_blk * BM + r_in;
    const bool row_ok = row < M;
    const bool zero_row = row_ok && ep.rowzero && ep.rowzero[row];
    const __nv_bfloat16 *mulp = (const __nv_bfloat16 *)ep.mul + (long)row * ep.ldmul;
    const __nv_bfloat16 *addp = (const __nv_bfloat16 *)ep.add + (long)row * ep.ldadd;
    constexpr int PANEL_COLS = 128 / (int)sizeof(TC);       // columns per 128-byte panel row: 32 (fp32) or 64 (bf16)
    constexpr int N_PANELS = BN / PANEL_COLS;
    static_assert(N_PANELS * BM * 128 <= STAGES * SM::STAGE_BYTES, "staging panels must fit in the dead pipeline stages");
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, r);
      const int col0 = n_blk * BN + c0;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (ep.out_scale != 0.f) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= ep.out_scale;
      }
      if (ep.bias) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b = __ldg(reinterpret_cast<const float4 *>(ep.bias + col0 + j));
          v[j] += b.x, v[j + 1] += b.y, v[j + 2] += b.z, v[j + 3] += b.w;
        }
      }
      if (ep.act == ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      } else if (ep.act == ACT_SIGMOID) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + __expf(-v[j]));
      }
      if (ep.mul && row_ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float t[8];
          bf16x8_to_f32(__ldg(reinterpret_cast<const uint4 *>(mulp + col0 + j)), t);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[j + i] *= t[i];
        }
      }
      if (ep.add && row_ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float t[8];
          bf16x8_to_f32(__ldg(reinterpret_cast<const uint4 *>(addp + col0 + j)), t);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[j + i] += t[i];
        }
      }
      if (zero_row) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      // stage: panel p holds PANEL_COLS columns; row r_in occupies 128 bytes; 16-byte chunk k sits at k ^ (r_in & 7)
      const int panel = c0 / PANEL_COLS;
      uint8_t *prow = smem + panel * (BM * 128) + r_in * 128;
      if constexpr (sizeof(TC) == 4) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          *reinterpret_cast<float4 *>(prow + ((k ^ (r_in & 7)) << 4)) =
              make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
      } else {
        const int kbase = (c0 % PANEL_COLS) / 8;  // 0 or 4: which half of the 64-column panel row
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float t[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] = v[8 * k + i];
          *reinterpret_cast<uint4 *>(prow + (((kbase + k) ^ (r_in & 7)) << 4)) = pack8<TC>(t);
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to TMA
    asm volatile("bar.sync 1, 128;" ::: "memory");                 // the four epilogue warps only
    if (warp == 2 && lane == 0) {
#pragma unroll 1
      for (int p = 0; p < N_PANELS; ++p)
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmC),
                     "r"(smem_u32(smem + p * (BM * 128))), "r"(n_blk * BN + p * PANEL_COLS), "r"(m_blk * BM)
                     : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem must stay intact until TMA has read it
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN) : "memory");
  }
}

template <int BN, typename TC>
static int launch(const void *A, int lda, const void *W, int ldw, void *C, int ldc, int M, int N, int K,
                  const Epilogue &ep, cudaStream_t st) {
  CUtensorMap tmA, tmW, tmC;
  if (!make_map(&tmA, A, M, K, lda, BM) || !make_map(&tmW, W, N, K, ldw, BN) ||
      !make_map(&tmC, C, M, N, ldc, BM, sizeof(TC) == 4, std::is_same<TC, __half>::value))
    return fail(MEMOTR_ECUDA, "linear(tc): cuTensorMapEncodeTiled failed (M=%d N=%d K=%d lda=%d)", M, N, K, lda);
  auto kern = gemm_tc_kernel<BN, TC>;
  static bool attr_set = false;  // idempotent attribute; benign if two threads race to set the same value
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BN>::TOTAL);
    if (e != cudaSuccess) return fail(MEMOTR_ECUDA, "linear(tc): smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid(N / BN, ceil_div(M, BM));
  MEMOTR_LAUNCH((kern), grid, 192, Smem<BN>::TOTAL, st, tmA, tmW, tmC, M, N, K, ep);
  return check_launch("gemm_tc");
}

}  // namespace tc

bool linear_tc_supported(int lda, int ldw, int ldc, int c_dtype, int M, int N, int K, const void *A, const void *W,
                         const void *C) {
  (void)M;
  const int cal = c_dtype == MEMOTR_F32 ? 4 : 8;  // 16-byte row segments on store (bf16 / fp16: 8 elements)
  return N % 64 == 0 && K % tc::BK == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldc % cal == 0 && aligned16(A) &&
         aligned16(W) && aligned16(C) && tc::encode_fn() != nullptr;
}

bool linear_tc_persist_wanted(int M, int N, int *n_sm_out);
int linear_tc_persist_bf16(const void *A, int lda, const void *W, int ldw, void *C, int ldc, int c_dtype, int M, int N,
                           int K, const Epilogue &ep, int n_sm, cudaStream_t st);

int linear_tc_bf16(const void *A, int lda, const void *W, int ldw, void *C, int ldc, int c_dtype, int M, int N, int K,
                   const Epilogue &ep, cudaStream_t st) {
  if (ep.mul && (ep.ldmul % 8 != 0 || !aligned16(ep.mul))) return fail(MEMOTR_EINVAL, "linear(tc): mul misaligned");
  if (ep.add && (ep.ldadd % 8 != 0 || !aligned16(ep.add))) return fail(MEMOTR_EINVAL, "linear(tc): add misaligned");
  if (ep.bias && !aligned16(ep.bias)) return fail(MEMOTR_EINVAL, "linear(tc): bias misaligned");
  const bool wide = N % 128 == 0;
  int n_sm = 0;
  if (linear_tc_persist_wanted(M, N, &n_sm))      // more tiles than SMs: persistent CTAs, double-buffered accumulators
    return linear_tc_persist_bf16(A, lda, W, ldw, C, ldc, c_dtype, M, N, K, ep, n_sm, st);
  if (c_dtype == MEMOTR_F16)
    return wide ? tc::launch<128, __half>(A, lda, W, ldw, C, ldc, M, N, K, ep, st)
                : tc::launch<64, __half>(A, lda, W, ldw, C, ldc, M, N, K, ep, st);
  if (c_dtype == MEMOTR_F32)
    return wide ? tc::launch<128, float>(A, lda, W, ldw, C, ldc, M, N, K, ep, st)
                : tc::launch<64, float>(A, lda, W, ldw, C, ldc, M, N, K, ep, st);
  return wide ? tc::launch<128, __nv_bfloat16>(A, lda, W, ldw, C, ldc, M, N, K, ep, st)
              : tc::launch<64, __nv_bfloat16>(A, lda, W, ldw, C, ldc, M, N, K, ep, st);
}

}  // namespace memotr
