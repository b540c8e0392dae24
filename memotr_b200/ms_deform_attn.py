"""MSDeformAttn -- the reference's nn.Module surface over the B200 operator.

Mirror of /root/reference/models/ops/modules/ms_deform_attn.py:36-130: constructor arguments, attribute and
parameter names (checkpoint keys `sampling_offsets`, `attention_weights`, `value_proj`, `output_proj`),
`reset_parameters()` initialisation scheme (:72-86) and the forward signature are kept so released MeMOTR checkpoints
load and `models/deformable_transformer.py:116-118` can keep calling `reset_parameters()`.

This module is the autograd-capable, any-shape path (projections through torch.nn.functional.linear, the sampling
core through MSDeformAttnFunction -> C ABI -> CUDA).  The inference engine (memotr_b200/engine.py) runs the same
arithmetic through fused kernels instead.
"""
import math
import warnings

import torch
import torch.nn.functional as F
from torch import nn

from .ms_deform_attn_func import MSDeformAttnFunction


def _power_of_two(n):
    if not isinstance(n, int) or n < 0:
        raise ValueError(f"invalid input for _is_power_of_2: {n} (type: {type(n)})")
    return n != 0 and (n & (n - 1)) == 0


class MSDeformAttn(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4, sigmoid_attn=False, visualize=False):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError(f"d_model must be divisible by n_heads, but got {d_model} and {n_heads}")
        if not _power_of_two(d_model // n_heads):
            warnings.warn("MSDeformAttn: a power-of-two head dimension (32 in MeMOTR) selects the vectorised kernels; "
                          "other sizes run the generic one.")
        self.im2col_step = 64
        self.sigmoid_attn = sigmoid_attn
        self.visualize = visualize
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self.reset_parameters()

    @torch.no_grad()
    def reset_parameters(self):
        """Zero offset/attention weights; offset bias = a ring of directions (one per head, max-norm 1) scaled by
        the point index 1..K; Xavier for the two dense projections (ms_deform_attn.py:72-86)."""
        H, L, K = self.n_heads, self.n_levels, self.n_points
        angle = torch.arange(H, dtype=torch.float32) * (2.0 * math.pi / H)
        direction = torch.stack([angle.cos(), angle.sin()], -1)
        direction = direction / direction.abs().max(-1, keepdim=True)[0]
        ring = direction.view(H, 1, 1, 2) * torch.arange(1, K + 1, dtype=torch.float32).view(1, 1, K, 1)
        self.sampling_offsets.weight.zero_()
        self.sampling_offsets.bias.copy_(ring.expand(H, L, K, 2).reshape(-1))
        self.attention_weights.weight.zero_()
        self.attention_weights.bias.zero_()
        nn.init.xavier_uniform_(self.value_proj.weight)
        self.value_proj.bias.zero_()
        nn.init.xavier_uniform_(self.output_proj.weight)
        self.output_proj.bias.zero_()

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None):
        """query (N,Lq,C); reference_points (N,Lq,L,2|4) in [0,1]; input_flatten (N,S,C); input_spatial_shapes (L,2)
        int64 (H_l,W_l); input_level_start_index (L,); input_padding_mask (N,S) bool, True = padding. -> (N,Lq,C)."""
        N, Lq, _ = query.shape
        _, S, _ = input_flatten.shape
        H, L, K = self.n_heads, self.n_levels, self.n_points
        # The reference asserts sum(H_l*W_l) == S on a CUDA tensor, i.e. one host sync per call (:102).  The
        # kernels index only through level_start_index/spatial_shapes, and the check is kept where it is free.
        if not input_spatial_shapes.is_cuda:
            assert int((input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum()) == S

        value = self.value_proj(input_flatten)
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None], 0.0)
        value = value.view(N, S, H, self.d_model // H)
        offsets = self.sampling_offsets(query).view(N, Lq, H, L, K, 2)
        weights = self.attention_weights(query).view(N, Lq, H, L * K)
        weights = (weights.sigmoid() if self.sigmoid_attn else F.softmax(weights, -1)).view(N, Lq, H, L, K)
        if reference_points.shape[-1] == 2:
            wh = torch.stack([input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1).to(offsets.dtype)
            locations = reference_points[:, :, None, :, None, :] + offsets / wh[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:
            locations = reference_points[:, :, None, :, None, :2] \
                + offsets / K * reference_points[:, :, None, :, None, 2:] * 0.5
        else:
            raise ValueError(f"Last dim of reference_points must be 2 or 4, but get {reference_points.shape[-1]} instead.")
        out = MSDeformAttnFunction.apply(value.contiguous(), input_spatial_shapes, input_level_start_index,
                                         locations.contiguous(), weights.contiguous(), self.im2col_step)
        return self.output_proj(out)
