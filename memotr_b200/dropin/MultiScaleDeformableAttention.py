"""Drop-in for the reference's compiled module `MultiScaleDeformableAttention`
(/root/reference/models/ops/src/vision.cpp:13-16, built by models/ops/setup.py:53).

Put this directory on sys.path ahead of (or instead of) the reference's build and
`import MultiScaleDeformableAttention as MSDA` (models/ops/functions/ms_deform_attn_func.py:21) resolves to the
B200 kernels: same two callables, same positional arguments, same return types.
"""
from memotr_b200.ms_deform_attn_func import ms_deform_attn_backward, ms_deform_attn_forward  # noqa: F401
