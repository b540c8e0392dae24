// gemm_tc_persist.cu -- persistent variant of gemm_tc.cu for GEMMs with more output tiles than SMs.
//
// The one-tile-per-CTA kernel is a latency chain per tile (TMA load ~1 us -> 16 MMAs 0.5 us -> commit -> TMEM read,
// epilogue math, staging, TMA store ~2 us) that only co-residency of 2-3 CTAs per SM overlaps: the encoder projections
// (350-525 tiles of 128 x 128 x 256) run at 160-190 TFLOP/s, 18-22 us each, with the tensor pipe busy < 15 % of the time.
// Here one CTA per SM walks its tiles (tile = blockIdx.x + i * gridDim.x, n fastest so that neighbours share the A tile
// in L2): the producer warp streams the k-blocks of consecutive tiles through a 4-stage ring without ever draining, the
// MMA warp alternates between TWO TMEM accumulators, and the epilogue warps drain accumulator (i & 1) -- bias / activation
// / multiplier / residual / padding mask, swizzled staging panels in their own shared-memory region, TMA store -- while
// the MMAs of tile i + 1 run.  Same contract and epilogue as gemm_tc.cu (models/ops/modules/ms_deform_attn.py:104-129
// projections and the other large-M nn.Linear call sites).
#include <type_traits>

#include "tc_common.cuh"

namespace memotr {
namespace tc {
namespace persist {

constexpr int BN = 128;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;   // 32 KB
// Shared-memory plan per output type.  The staging tile is DOUBLE-BUFFERED: the in-kernel stamps (tools/gemm_phases.py)
// showed the epilogue of tile i + 1 waiting 1.2 us for the TMA store of tile i to finish reading a single staging buffer
// (64 KB of fp32 per tile) -- as long as the epilogue itself.  fp32 output: 3 ring stages + 2 x 64 KB; 16-bit output: 4 ring
// stages + 2 x 32 KB.
template <typename TC>
struct Plan {
  static constexpr int NST = sizeof(TC) == 4 ? 3 : 4;
  static constexpr int OUT_BYTES = (BN * (int)sizeof(TC) / 128) * BM * 128;                   // 4 (fp32) or 2 panels of 16 KB
  static constexpr int OFF_OUT = NST * STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_OUT + 2 * OUT_BYTES;
  static constexpr int TOTAL = OFF_BAR + 256 + 1024;
  static_assert(TOTAL <= 227 * 1024, "shared memory plan");
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <typename TC>
__global__ void __launch_bounds__(320, 1)
gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                       const __grid_constant__ CUtensorMap tmC, int M, int N, int K, Epilogue ep) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int NST = Plan<TC>::NST, OFF_OUT = Plan<TC>::OFF_OUT, OUT_BYTES = Plan<TC>::OUT_BYTES, OFF_BAR = Plan<TC>::OFF_BAR;
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + OFF_BAR), *empty = full + NST, *acc_full = empty + NST,
           *acc_empty = acc_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = K / BK, n_nblk = N / BN, n_tiles = n_nblk * ceil_div(M, BM);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    for (int s = 0; s < NST; ++s) mbar_init(full + s, 1), mbar_init(empty + s, 1);
    for (int b = 0; b < 2; ++b) mbar_init(acc_full + b, 1), mbar_init(acc_empty + b, 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_grid_sync();

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int m_blk = t / n_nblk, n_blk = t % n_nblk;
        for (int kb = 0; kb < num_k; ++kb, ++it) {
          const int s = it % NST;
          mbar_wait(empty + s, ((it / NST) & 1) ^ 1);
          mbar_expect_tx(full + s, STAGE_BYTES);
          uint8_t *sa = smem + s * STAGE_BYTES;
          tma_load_2d(sa, &tmA, full + s, kb * BK, m_blk * BM);
          tma_load_2d(sa + A_BYTES, &tmW, full + s, kb * BK, n_blk * BN);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ep.in_f16 ? umma_idesc_f16(BN) : umma_idesc(BN);
      uint32_t it = 0;
      int i = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
        const int buf = i & 1;
        mbar_wait(acc_empty + buf, ((i >> 1) & 1) ^ 1);      // the epilogue has drained this accumulator (tile i - 2)
        tcgen05_fence_after();
        for (int kb = 0; kb < num_k; ++kb, ++it) {
          const int s = it % NST;
          mbar_wait(full + s, (it / NST) & 1);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint64_t adesc = umma_desc(sa), bdesc = umma_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(tmem_base + buf * BN, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(empty + s);
        }
        umma_commit(acc_full + buf);
      }
    }
  } else {
    // ---- epilogue warps 2..9: TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 (the epilogue is the pace-maker
    //      of this kernel: with four warps the tile rate was bound by 128 threads x 128 values each) ----
    const int chalf = (warp - 2) >> 2;
    long long *stamps = (ep.stamps && warp == 2 && lane == 0) ? ep.stamps + 20 * blockIdx.x : nullptr;
    if (stamps) stamps[0] = clock64();
    int i = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
      const int m_blk = t / n_nblk, n_blk = t % n_nblk, buf = i & 1;
      // staging buffer (i & 1): the TMA store of tile i - 2 must have finished READING it (the store of tile i - 1 may still run)
      uint8_t *stage_out = smem + OFF_OUT + (i & 1) * OUT_BYTES;
      if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (stamps && i < 6) stamps[1 + 3 * i] = clock64();        // staging free (previous store has read it)
      mbar_wait(acc_full + buf, (i >> 1) & 1);
      tcgen05_fence_after();
      if (stamps && i < 6) stamps[2 + 3 * i] = clock64();        // accumulator ready
        const int quarter = warp & 3;
        const int r_in = quarter * 32 + lane;  // row inside the tile == TMEM lane
        const int row = m_blk * BM + r_in;
        const bool row_ok = row < M;
        float prep_bx = 0.f, prep_by = 0.f;         // encoder reference point of this row (pixel centre / valid extent)
        float lvl_sx[4], lvl_sy[4], lvl_rw[4], lvl_rh[4];   // K = 4 fast path: per value level, reference * valid ratio and 1 / extent
        if (ep.prep) {
          const int rr = row_ok ? row : 0;
          int lq = 0;
          for (int t2 = 1; t2 < ep.prep_L; ++t2)
            if (rr >= ep.prep_lsi[t2]) lq = t2;
          const int Wq = ep.prep_hw[2 * lq + 1], Hq = ep.prep_hw[2 * lq], pp = rr - ep.prep_lsi[lq];
          const int yy = (int)(((float)pp + 0.5f) * __frcp_rn((float)Wq)), xx = pp - yy * Wq;
          prep_bx = ((float)xx + 0.5f) * __frcp_rn(__ldg(ep.prep_vr + 2 * lq) * (float)Wq);
          prep_by = ((float)yy + 0.5f) * __frcp_rn(__ldg(ep.prep_vr + 2 * lq + 1) * (float)Hq);
          if (ep.prep_K == 4) {
#pragma unroll
            for (int l = 0; l < 4; ++l) {
              lvl_sx[l] = prep_bx * __ldg(ep.prep_vr + 2 * l), lvl_sy[l] = prep_by * __ldg(ep.prep_vr + 2 * l + 1);
              lvl_rw[l] = __frcp_rn((float)ep.prep_hw[2 * l + 1]), lvl_rh[l] = __frcp_rn((float)ep.prep_hw[2 * l]);
            }
          }
        }
        const bool zero_row = row_ok && ep.rowzero && ep.rowzero[row];
        const __nv_bfloat16 *mulp = (const __nv_bfloat16 *)ep.mul + (long)row * ep.ldmul;
        const __nv_bfloat16 *addp = (const __nv_bfloat16 *)ep.add + (long)row * ep.ldadd;
        constexpr int PANEL_COLS = 128 / (int)sizeof(TC);       // columns per 128-byte panel row: 32 (fp32) or 64 (bf16)
        constexpr int N_PANELS = BN / PANEL_COLS;
    #pragma unroll 1
        for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN + c0), r);
          if (c0 + 32 == (chalf + 1) * (BN / 2)) {   // this warp's half of the accumulator has been read: hand it back
            tcgen05_fence_before();
            mbar_arrive(acc_empty + buf);
          }
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
          if (ep.prep) {
            if (col0 < ep.prep_nh * 32 && ep.prep_K == 4) {   // 16 (x, y) offsets of one head -> sampling locations
#pragma unroll
              for (int i2 = 0; i2 < 16; ++i2) {              // (same products and sums as the general branch below, the
                v[2 * i2] = lvl_sx[i2 >> 2] + v[2 * i2] * lvl_rw[i2 >> 2];          //  per-level factors hoisted out of the loop)
                v[2 * i2 + 1] = lvl_sy[i2 >> 2] + v[2 * i2 + 1] * lvl_rh[i2 >> 2];
              }
            } else if (col0 < ep.prep_nh * 32) {
#pragma unroll
              for (int i2 = 0; i2 < 16; ++i2) {
                const int l = i2 / ep.prep_K;
                v[2 * i2] = prep_bx * __ldg(ep.prep_vr + 2 * l) + v[2 * i2] * __frcp_rn((float)ep.prep_hw[2 * l + 1]);
                v[2 * i2 + 1] = prep_by * __ldg(ep.prep_vr + 2 * l + 1) + v[2 * i2 + 1] * __frcp_rn((float)ep.prep_hw[2 * l]);
              }
            } else {                                  // the logits of two heads -> softmax over each head's 16 points
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                float mx = v[16 * hh];
#pragma unroll
                for (int i2 = 1; i2 < 16; ++i2) mx = fmaxf(mx, v[16 * hh + i2]);
                float sum = 0.f;
#pragma unroll
                for (int i2 = 0; i2 < 16; ++i2) v[16 * hh + i2] = __expf(v[16 * hh + i2] - mx), sum += v[16 * hh + i2];
                const float rs = __frcp_rn(sum);
#pragma unroll
                for (int i2 = 0; i2 < 16; ++i2) v[16 * hh + i2] *= rs;
              }
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
          uint8_t *prow = stage_out + panel * (BM * 128) + r_in * 128;
          if constexpr (sizeof(TC) == 4) {
            if (ep.split3_n) {
              // split output: panels 0, 1 = fp16 hi of columns [0, 64), [64, 128) of the tile; panels 2, 3 = the lo terms
              uint8_t *ph = stage_out + (c0 / 64) * (BM * 128) + r_in * 128, *pl = ph + 2 * (BM * 128);
              const int kb = (c0 % 64) / 8;
    #pragma unroll
              for (int k = 0; k < 4; ++k) {
                uint32_t hw[4], lw[4];
    #pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const __half2 h = __floats2half2_rn(v[8 * k + 2 * i], v[8 * k + 2 * i + 1]);
                  const float2 f = __half22float2(h);
                  const __half2 l = __floats2half2_rn(v[8 * k + 2 * i] - f.x, v[8 * k + 2 * i + 1] - f.y);
                  hw[i] = *reinterpret_cast<const uint32_t *>(&h), lw[i] = *reinterpret_cast<const uint32_t *>(&l);
                }
                *reinterpret_cast<uint4 *>(ph + (((kb + k) ^ (r_in & 7)) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4 *>(pl + (((kb + k) ^ (r_in & 7)) << 4)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
              }
              continue;
            }
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

      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (stamps && i < 6) stamps[3 + 3 * i] = clock64();        // tile staged
      if (warp == 2 && lane == 0) {
        if (sizeof(TC) == 4 && ep.split3_n) {       // (tmC: the (M, 3N) fp16 tensor; hi to columns n and N + n, lo to 2N + n)
#pragma unroll 1
          for (int p = 0; p < 6; ++p) {
            const int src = p < 4 ? (p & 1) : 2 + (p & 1), col = (p >> 1) * ep.split3_n + n_blk * BN + (p & 1) * 64;
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmC),
                         "r"(smem_u32(stage_out + src * (BM * 128))), "r"(col), "r"(m_blk * BM)
                         : "memory");
          }
        } else {
#pragma unroll 1
        for (int p = 0; p < N_PANELS; ++p)
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmC),
                       "r"(smem_u32(stage_out + p * (BM * 128))), "r"(n_blk * BN + p * PANEL_COLS), "r"(m_blk * BM)
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete
    if (stamps) stamps[19] = clock64();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}

static long long *g_stamps = nullptr;   // memotr_gemm_debug_stamps

template <typename TC>
static int launch(const void *A, int lda, const void *W, int ldw, void *C, int ldc, int M, int N, int K,
                  const Epilogue &ep, int n_sm, cudaStream_t st) {
  CUtensorMap tmA, tmW, tmC;
  const bool split3 = sizeof(TC) == 4 && ep.split3_n != 0;    // the output tensor is then (M, 3N) fp16, 64-column panels
  if (!make_map(&tmA, A, M, K, lda, BM) || !make_map(&tmW, W, N, K, ldw, BN) ||
      !(split3 ? make_map(&tmC, C, M, 3 * N, ldc, BM, false, true)
               : make_map(&tmC, C, M, N, ldc, BM, sizeof(TC) == 4, std::is_same<TC, __half>::value)))
    return fail(MEMOTR_ECUDA, "linear(tc, persistent): cuTensorMapEncodeTiled failed (M=%d N=%d K=%d lda=%d)", M, N, K, lda);
  auto kern = gemm_tc_persist_kernel<TC>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Plan<TC>::TOTAL);
    if (e != cudaSuccess) return fail(MEMOTR_ECUDA, "linear(tc, persistent): smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles = (N / BN) * ceil_div(M, BM);
  Epilogue eps = ep;
  eps.stamps = g_stamps;
  const int ctas = sm_limit(n_sm);      // (memotr_set_sm_budget: leave SMs to a concurrent kernel)
  MEMOTR_LAUNCH((kern), tiles < ctas ? tiles : ctas, 320, Plan<TC>::TOTAL, st, tmA, tmW, tmC, M, N, K, eps);
  return check_launch("gemm_tc_persist");
}

}  // namespace persist
}  // namespace tc

// used by linear_tc_bf16 (gemm_tc.cu) when there are more 128 x 128 tiles than SMs; MEMOTR_GEMM_PERSIST=0 switches it off
bool linear_tc_persist_wanted(int M, int N, int *n_sm_out) {
  static int n_sm = 0, enabled = -1;
  if (enabled < 0) {
    const char *e = getenv("MEMOTR_GEMM_PERSIST");
    enabled = !(e && e[0] == '0');
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  }
  *n_sm_out = n_sm;
  return enabled && N % tc::persist::BN == 0 && (long)(N / tc::persist::BN) * ceil_div(M, tc::BM) > n_sm;
}

int linear_tc_persist_bf16(const void *A, int lda, const void *W, int ldw, void *C, int ldc, int c_dtype, int M, int N,
                           int K, const Epilogue &ep, int n_sm, cudaStream_t st) {
  if (c_dtype == MEMOTR_F16) return tc::persist::launch<__half>(A, lda, W, ldw, C, ldc, M, N, K, ep, n_sm, st);
  if (c_dtype == MEMOTR_F32) return tc::persist::launch<float>(A, lda, W, ldw, C, ldc, M, N, K, ep, n_sm, st);
  return tc::persist::launch<__nv_bfloat16>(A, lda, W, ldw, C, ldc, M, N, K, ep, n_sm, st);
}

}  // namespace memotr

using namespace memotr;

// Profiling hook (tools/micro_gemm.py): every later persistent-GEMM launch writes 20 clock64 stamps per CTA into `buf` (device
// int64[20 x CTAs]; null: off): 0 start; per tile i < 6: 1+3i staging free, 2+3i accumulator ready, 3+3i tile staged; 19 end.
extern "C" int memotr_gemm_debug_stamps(long long *buf) {
  tc::persist::g_stamps = buf;
  return MEMOTR_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// fp32-accurate GEMM on the tensor cores (the reference's precision contract is TF32 OFF, main.py:96-97; the fp32 engine ran
// its encoder GEMMs on CUDA cores: 7.3 of 8.1 ms per frame).  Two-term fp16 splits: x = x_hi + x_lo, 2^s W = w_hi + w_lo
// (11 + 11 mantissa bits each; the power-of-two scale keeps w_lo a normal fp16), and
//     x W^T  =  2^-s (x_hi w_hi + x_hi w_lo + x_lo w_hi)  +  O(2^-22)
// evaluated as ONE fp16 GEMM with three times the K extent: A3 = [x_hi | x_hi | x_lo] (M, 3K), W3 = [w_hi | w_lo | w_hi]
// (N, 3K), fp32 accumulation in TMEM, the scale applied to the accumulator before the bias.  oracle/frame.py models exactly
// this ("fp16x3": 9.4e-5 on the white-noise full-size golden where plain fp32 summation orders already differ by 1e-5).
__global__ void __launch_bounds__(256)
split3_kernel(const float *__restrict__ x, int ldx, __half *__restrict__ a3, int M, int K) {
  pdl_grid_sync();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;            // one thread per 4 consecutive elements of a row
  const int k4 = K >> 2;
  if (idx >= (long)M * k4) return;
  const int r = (int)(idx / k4), c = (int)(idx % k4) * 4;
  const float4 v = *reinterpret_cast<const float4 *>(x + (long)r * ldx + c);
  const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
  const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
  const __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
  __half *row = a3 + (long)r * 3 * K + c;
  const uint2 hi = make_uint2(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1));
  const uint2 lo = make_uint2(*reinterpret_cast<const uint32_t *>(&l0), *reinterpret_cast<const uint32_t *>(&l1));
  *reinterpret_cast<uint2 *>(row) = hi;
  *reinterpret_cast<uint2 *>(row + K) = hi;
  *reinterpret_cast<uint2 *>(row + 2 * K) = lo;
}

namespace memotr {
int linear_tc_bf16(const void *A, int lda, const void *W, int ldw, void *C, int ldc, int c_dtype, int M, int N, int K,
                   const Epilogue &ep, cudaStream_t st);
}

extern "C" int memotr_linear_f32x3(const float *A, int lda, const void *W3, const float *bias, const unsigned char *rowzero, float *C,
                                   int ldc, int M, int N, int K, int act, float w_scale_inv, void *scratch_a3, void *split_out,
                                   void *stream) {
  // A == NULL: scratch_a3 already holds the split operand [hi | hi | lo] (M, 3K) -- the split_out of a previous call.
  // split_out != NULL (C may be NULL): the result is written as the split fp16 operand (M, 3N) of the next call instead of as
  // fp32 (the FFN: linear1 -> relu -> linear2 without the fp32 hidden activation in HBM).
  MEMOTR_REQUIRE(W3 && (C || split_out) && scratch_a3 && M > 0 && N > 0 && K > 0, "linear_f32x3: bad arguments");
  MEMOTR_REQUIRE(K % 64 == 0 && (!A || (lda % 4 == 0 && aligned16(A))) && aligned16(W3) && aligned16(scratch_a3) &&
                     (!C || (ldc % 4 == 0 && aligned16(C))) && (!split_out || aligned16(split_out)) && (!bias || aligned16(bias)) &&
                     act >= 0 && act <= 2 && tc::encode_fn() != nullptr,
                 "linear_f32x3: needs K %% 64 == 0 and 16-byte aligned buffers");
  MEMOTR_REQUIRE(N % 64 == 0, "linear_f32x3: needs N %% 64 == 0 (N = %d)", N);
  cudaStream_t st = (cudaStream_t)stream;
  if (A) {
    const long n4 = (long)M * (K / 4);
    MEMOTR_LAUNCH((split3_kernel), (int)((n4 + 255) / 256), 256, 0, st, A, lda, (__half *)scratch_a3, M, K);
    const int rc = check_launch("split3");
    if (rc != MEMOTR_OK) return rc;
  }
  Epilogue ep{bias, nullptr, nullptr, rowzero, 0, 0, act};
  ep.in_f16 = 1, ep.out_scale = w_scale_inv;
  if (split_out) {
    int n_sm = 0;
    MEMOTR_REQUIRE(N % 128 == 0 && linear_tc_persist_wanted(M, N, &n_sm),
                   "linear_f32x3: split output needs N %% 128 == 0 and more 128 x 128 tiles than SMs (M = %d, N = %d)", M, N);
    ep.split3_n = N;
    return linear_tc_persist_bf16(scratch_a3, 3 * K, W3, 3 * K, split_out, 3 * N, MEMOTR_F32, M, N, 3 * K, ep, n_sm, st);
  }
  return linear_tc_bf16(scratch_a3, 3 * K, W3, 3 * K, C, ldc, MEMOTR_F32, M, N, 3 * K, ep, st);   // persistent when tiles > SMs
}

extern "C" int memotr_linear_msda_prep(const void *A, int lda, const void *W, int ldw, const float *bias, float *out, int ldo,
                                       int M, int K, int n_heads, int n_levels, int n_points, const int *shapes_hw,
                                       const int *level_start, const float *valid_ratios, void *stream) {
  MEMOTR_REQUIRE(A && W && out && shapes_hw && level_start && valid_ratios && M > 0, "linear_msda_prep: bad arguments");
  MEMOTR_REQUIRE(n_levels >= 1 && n_levels <= 8 && n_levels * n_points == 16 && n_heads % 2 == 0,
                 "linear_msda_prep: needs levels x points == 16 and an even head count");
  const int N = n_heads * 48;
  int n_sm = 0;
  MEMOTR_REQUIRE(K % tc::BK == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldo % 4 == 0 && aligned16(A) && aligned16(W) && aligned16(out) &&
                     (!bias || aligned16(bias)) && tc::encode_fn() != nullptr,
                 "linear_msda_prep: misaligned buffer");
  MEMOTR_REQUIRE(linear_tc_persist_wanted(M, N, &n_sm),
                 "linear_msda_prep: needs N %% 128 == 0 and more 128 x 128 tiles than SMs (M = %d, N = %d)", M, N);
  Epilogue ep{bias, nullptr, nullptr, nullptr, 0, 0, ACT_NONE};
  ep.prep = 1, ep.prep_L = n_levels, ep.prep_K = n_points, ep.prep_nh = n_heads, ep.prep_vr = valid_ratios;
  for (int l = 0; l < n_levels; ++l)
    ep.prep_hw[2 * l] = shapes_hw[2 * l], ep.prep_hw[2 * l + 1] = shapes_hw[2 * l + 1], ep.prep_lsi[l] = level_start[l];
  return linear_tc_persist_bf16(A, lda, W, ldw, out, ldo, MEMOTR_F32, M, N, K, ep, n_sm, (cudaStream_t)stream);
}
