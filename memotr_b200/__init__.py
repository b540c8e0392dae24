"""memotr_b200 -- B200-native (sm_100a) implementation of MeMOTR's per-frame deformable-transformer hot path.

Host side: Python mirrors of the reference's operator/module surfaces (same names, arguments, state_dict keys).
Device side: hand-written CUDA behind a C ABI (include/memotr_b200.h, memotr_b200/csrc/).  No CPU fallback.
"""
from .ms_deform_attn_func import MSDeformAttnFunction, ms_deform_attn_backward, ms_deform_attn_forward  # noqa: F401

__all__ = ["MSDeformAttnFunction", "ms_deform_attn_forward", "ms_deform_attn_backward"]
