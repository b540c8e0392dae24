/*
 * include/memotr_b200.h -- C ABI of the B200-native MeMOTR hot path (libmemotr_b200.so).
 *
 * Plain pointers and sizes only: no torch / ATen types cross this boundary.  Every entry point
 *   - takes DEVICE pointers to caller-allocated, contiguous buffers (inputs are borrowed, never freed);
 *   - enqueues its kernels on the CUDA stream passed as `stream` (a cudaStream_t cast to void*; NULL = legacy
 *     default stream) and never synchronises, allocates device memory, or keeps global mutable state
 *     (re-entrant; safe to call from the autograd engine thread and under CUDA-graph capture);
 *   - returns 0 on success or a negative MEMOTR_E* code; memotr_last_error() then returns a thread-local
 *     message.  (The reference only printf()s launch errors -- ms_deform_im2col_cuda.cuh:948-952 -- its argument
 *     errors are AT_ASSERTM/AT_ERROR exceptions, ms_deform_attn_cuda.cu:28-52; the Python shim turns a
 *     non-zero return into the same RuntimeError.)
 *
 * Reference interfaces replaced (paths relative to /root/reference/models/ops/):
 *   memotr_msda_forward   <- ms_deform_attn_forward   src/ms_deform_attn.h:20-39,  src/vision.cpp:14
 *                            ms_deform_attn_cuda_forward  src/cuda/ms_deform_attn_cuda.cu:20-80
 *   memotr_msda_backward  <- ms_deform_attn_backward  src/ms_deform_attn.h:41-61,  src/vision.cpp:15
 *                            ms_deform_attn_cuda_backward src/cuda/ms_deform_attn_cuda.cu:83-153
 * The remaining entry points have no native counterpart in the reference (there the work is a chain of ATen
 * calls issued from Python); each cites the Python lines it stands in for.
 */
#ifndef MEMOTR_B200_H_
#define MEMOTR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MEMOTR_ABI_VERSION 1

#if defined(__GNUC__)
#define MEMOTR_API __attribute__((visibility("default")))
#else
#define MEMOTR_API
#endif

/* element types of the floating-point buffers */
#define MEMOTR_F32 0
#define MEMOTR_F64 1
#define MEMOTR_BF16 2

/* return codes */
#define MEMOTR_OK 0
#define MEMOTR_EINVAL (-1)   /* bad argument (shape, dtype, null pointer, misalignment) */
#define MEMOTR_ECUDA (-2)    /* CUDA runtime / launch error                             */
#define MEMOTR_ENOSYS (-3)   /* combination not implemented                             */

MEMOTR_API int memotr_abi_version(void);
MEMOTR_API const char *memotr_last_error(void);

/*
 * Multi-scale deformable attention, forward.
 *   value            (B, S, H, D)        dtype     pixel-major, S = sum_l H_l*W_l
 *   spatial_shapes   (L, 2)              int64     (H_l, W_l), device memory
 *   level_start_idx  (L,)                int64     device memory
 *   sampling_loc     (B, Lq, H, L, K, 2) dtype     last dim (x, y) in [0,1] image coordinates
 *   attn_weight      (B, Lq, H, L, K)    dtype
 *   output           (B, Lq, H*D)        dtype     every element is written (no pre-zeroing needed)
 * dtype F32 with D == 32 reproduces the reference kernel's fp32 rounding sequence bit for bit.
 * The reference's `im2col_step` batching argument has no effect on the result and is not part of this ABI.
 */
MEMOTR_API int memotr_msda_forward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_idx,
                        const void *sampling_loc, const void *attn_weight, void *output, int B, int S, int H,
                        int D, int L, int Lq, int K, int dtype, void *stream);

/*
 * Multi-scale deformable attention, backward.
 *   grad_output      (B, Lq, H*D)  dtype
 *   grad_value       like value         -- MUST be zero-filled by the caller (accumulated with atomics)
 *   grad_sampling_loc like sampling_loc -- fully written
 *   grad_attn_weight like attn_weight   -- fully written
 * dtype F32 or F64.
 */
MEMOTR_API int memotr_msda_backward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_idx,
                         const void *sampling_loc, const void *attn_weight, const void *grad_output,
                         void *grad_value, void *grad_sampling_loc, void *grad_attn_weight, int B, int S, int H,
                         int D, int L, int Lq, int K, int dtype, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MEMOTR_B200_H_ */
