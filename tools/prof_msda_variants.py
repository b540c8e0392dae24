#!/usr/bin/env python
"""tools/prof_msda_variants.py -- one encoder-shaped launch of each MSDA forward variant (ncu target)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from memotr_b200 import kernels, synthetic as synth  # noqa: E402

shapes = synth.DANCETRACK_SHAPES
S = sum(h * w for h, w in shapes)
value, shp, lsi, loc, attn = (x.cuda() for x in synth.msda_inputs(shapes, Lq=S, K=4, seed=1))
vb = value.reshape(S, 256).bfloat16()
lc, aw = loc[0].contiguous(), attn[0].contiguous()
for rep in range(2):
    os.environ["MEMOTR_MSDA_KERNEL"] = "v1"
    kernels.msda_forward_ex(vb, shp, lsi, lc, aw, 8)
    os.environ["MEMOTR_MSDA_KERNEL"] = "v2"
    os.environ["MEMOTR_MSDA_U"] = "1"
    kernels.msda_forward_ex(vb, shp, lsi, lc, aw, 8)
    pr = kernels.msda_pairs_layout(vb, shp, lsi, 8)
    kernels.msda_forward_pairs(pr, shp, lsi, lc, aw)
torch.cuda.synchronize()
