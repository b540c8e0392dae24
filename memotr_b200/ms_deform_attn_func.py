"""MSDeformAttnFunction and the two operator entry points, on the B200 kernels.

Mirrors /root/reference/models/ops/functions/ms_deform_attn_func.py:24-41 (the autograd Function) and the pybind
module `MultiScaleDeformableAttention` (src/vision.cpp:13-16): same names, positional signatures, saved tensors and
return arity.  `im2col_step` is accepted for compatibility; the kernels process the whole batch in one launch and
the result does not depend on it (only the reference's divisibility check is kept).
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib


def _dims(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    B, S, H, D = value.shape
    L = spatial_shapes.shape[0]
    Lq, K = sampling_loc.shape[1], sampling_loc.shape[4]
    if tuple(sampling_loc.shape) != (B, Lq, H, L, K, 2) or tuple(attn_weight.shape) != (B, Lq, H, L, K):
        raise RuntimeError(f"inconsistent shapes: value {tuple(value.shape)}, sampling_loc "
                           f"{tuple(sampling_loc.shape)}, attn_weight {tuple(attn_weight.shape)}")
    if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
        raise RuntimeError("spatial_shapes and level_start_index must be int64")
    if not (sampling_loc.dtype == value.dtype and attn_weight.dtype == value.dtype):
        raise RuntimeError("value, sampling_loc and attn_weight must share one dtype")
    step = min(B, int(im2col_step))
    if B > 0 and B % step != 0:                       # ms_deform_attn_cuda.cu:52
        raise RuntimeError(f"batch({B}) must divide im2col_step({step})")
    return B, S, H, D, L, Lq, K


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step=64):
    """-> (B, Lq, H*D).  Drop-in for MSDA.ms_deform_attn_forward (src/ms_deform_attn.h:20-39)."""
    _lib.require_cuda(value=value, spatial_shapes=spatial_shapes, level_start_index=level_start_index,
                      sampling_loc=sampling_loc, attn_weight=attn_weight)
    B, S, H, D, L, Lq, K = _dims(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step)
    out = torch.empty((B, Lq, H * D), dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device):
        rc = _lib.lib().memotr_msda_forward(_lib.ptr(value), _lib.ptr(spatial_shapes), _lib.ptr(level_start_index),
                                            _lib.ptr(sampling_loc), _lib.ptr(attn_weight), _lib.ptr(out),
                                            B, S, H, D, L, Lq, K, _lib.dtype_code(value), _lib.stream_ptr())
    _lib.check(rc, "ms_deform_attn_forward")
    return out


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step=64):
    """-> [grad_value, grad_sampling_loc, grad_attn_weight].  Drop-in for MSDA.ms_deform_attn_backward."""
    _lib.require_cuda(value=value, spatial_shapes=spatial_shapes, level_start_index=level_start_index,
                      sampling_loc=sampling_loc, attn_weight=attn_weight)
    grad_output = grad_output.contiguous()
    B, S, H, D, L, Lq, K = _dims(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step)
    grad_value = torch.zeros_like(value)
    grad_loc = torch.empty_like(sampling_loc)
    grad_attn = torch.empty_like(attn_weight)
    with torch.cuda.device(value.device):
        rc = _lib.lib().memotr_msda_backward(_lib.ptr(value), _lib.ptr(spatial_shapes), _lib.ptr(level_start_index),
                                             _lib.ptr(sampling_loc), _lib.ptr(attn_weight), _lib.ptr(grad_output),
                                             _lib.ptr(grad_value), _lib.ptr(grad_loc), _lib.ptr(grad_attn),
                                             B, S, H, D, L, Lq, K, _lib.dtype_code(value), _lib.stream_ptr())
    _lib.check(rc, "ms_deform_attn_backward")
    return [grad_value, grad_loc, grad_attn]


class MSDeformAttnFunction(Function):
    # Under torch.autocast the linear layers around the op hand it bf16 / fp16 tensors.  The reference's op has no half
    # instantiation at all (AT_DISPATCH_FLOATING_TYPES, src/cuda/ms_deform_attn_cuda.cu:64) -- its training runs in fp32 -- so
    # a mixed-precision step (BASELINE.json config 3) keeps the sampling core in fp32: inputs are cast up on entry and
    # autocast is off inside, exactly what custom_fwd(cast_inputs=float32) is for.  Outside autocast nothing changes.
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step):
        ctx.im2col_step = im2col_step
        output = ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                                        attention_weights, ctx.im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return output

    @staticmethod
    @once_differentiable
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_output):
        value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights = ctx.saved_tensors
        grad_value, grad_sampling_loc, grad_attn_weight = ms_deform_attn_backward(
            value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights, grad_output,
            ctx.im2col_step)
        return grad_value, None, None, grad_sampling_loc, grad_attn_weight, None
