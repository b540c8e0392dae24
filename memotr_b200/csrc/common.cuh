// common.cuh -- shared helpers for the sm_100a kernels behind the C ABI (include/memotr_b200.h).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <utility>

#include "../../include/memotr_b200.h"

namespace memotr {

// ---- error plumbing: thread-local message, integer return codes across the C ABI ---------------------------
inline char *err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
inline int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
inline int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MEMOTR_ECUDA, "%s: %s", what, cudaGetErrorString(e));
  return MEMOTR_OK;
}

#define MEMOTR_REQUIRE(cond, ...)                                  \
  do {                                                             \
    if (!(cond)) return ::memotr::fail(MEMOTR_EINVAL, __VA_ARGS__); \
  } while (0)

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------------
// Half of a frame is ~200 small kernels on <= 400 rows whose cost is launch latency, not work.  Every kernel of this
// library starts with pdl_grid_sync() (griddepcontrol.wait: returns once the preceding grid in the stream has completed
// and its writes are visible -- the ordinary stream-order guarantee) followed by griddepcontrol.launch_dependents, and
// is launched with the programmatic-stream-serialization attribute, so the NEXT kernel's launch, block scheduling and
// (for the GEMM) barrier/TMEM set-up overlap the tail of this one instead of following it.  Semantics are unchanged:
// no dependent data is touched before the wait.  MEMOTR_PDL=0 launches plainly (A/B measurement, debugging).
__device__ __forceinline__ void pdl_grid_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

inline bool pdl_enabled() {
  static const bool on = []() {
    const char *e = getenv("MEMOTR_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

template <typename... KArgs, typename... Args>
inline void launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(std::forward<Args>(args))...);  // errors: check_launch()
}
// same, for kernels launched as thread-block clusters of `cluster_x` CTAs (TMA multicast between neighbouring SMs)
template <typename... KArgs, typename... Args>
inline void launch_kernel_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                  unsigned cluster_x, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(std::forward<Args>(args))...);
}
#define MEMOTR_LAUNCH(kern, grid, block, smem, st, ...) \
  ::memotr::launch_kernel(kern, dim3(grid), dim3(block), (size_t)(smem), st, __VA_ARGS__)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- small device helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_f4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// 8 bf16 (16 bytes) -> 8 floats
__device__ __forceinline__ void bf16x8_to_f32(const uint4 &u, float (&f)[8]) {
  const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 f32x8_to_bf16(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

__device__ __forceinline__ uint4 f32x8_to_f16(const float (&f)[8]) {
  uint4 u;
  __half2 *h = reinterpret_cast<__half2 *>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
// 8 floats -> 16 bytes of the 2-byte type T (bf16 or fp16)
template <typename T>
__device__ __forceinline__ uint4 pack8(const float (&f)[8]);
template <>
__device__ __forceinline__ uint4 pack8<__nv_bfloat16>(const float (&f)[8]) { return f32x8_to_bf16(f); }
template <>
__device__ __forceinline__ uint4 pack8<__half>(const float (&f)[8]) { return f32x8_to_f16(f); }

// ---- dtype conversion + the GEMM epilogue description shared by gemm_simt.cu and gemm_tc.cu -------------------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2 };

// SM budget of the persistent kernels (memotr_set_sm_budget, capi.cu): min(n_sm, budget) when a budget is set
int sm_limit(int n_sm);

struct Epilogue {
  const float *bias;             // (N) or null
  const void *mul;               // (M,N) activation dtype or null: result *= mul
  const void *add;               // (M,N) activation dtype or null: result += add  (residual)
  const unsigned char *rowzero;  // (M) or null: rows with rowzero[m] != 0 are written as 0 (padding mask)
  int ldmul, ldadd, act;
  // MSDA "prep" epilogue of the offsets / attention-logits projection (gemm_tc_persist.cu only; prep = 0: off): the raw
  // row [offsets (H, L*K, 2) | logits (H, L*K)] is turned in place into [sampling locations | softmax weights]
  // (ms_deform_attn.py:108-120, encoder reference points of deformable_encoder.py:29-40); needs L*K == 16
  int prep, prep_L, prep_K, prep_nh;
  int prep_hw[16], prep_lsi[8];   // (H_l, W_l) and first row of every level
  const float *prep_vr;           // device (L, 2) valid ratios
  long long *stamps;              // profiling (tools/micro_gemm.py): 20 clock64 stamps per CTA of the persistent GEMM, else null
  int in_f16;                     // persistent tcgen05 GEMM: the 16-bit operands are fp16, not bf16 (memotr_linear_f32x3)
  float out_scale;                // != 0: the accumulator is multiplied by it before the bias (exact power of two of the split weights)
  int split3_n;                   // persistent GEMM, fp32 instantiation: != 0 = N of this GEMM; the result is written as the split fp16
                                  // A operand [hi | hi | lo] (M, 3N) of the NEXT memotr_linear_f32x3 instead of as fp32
};

}  // namespace memotr
