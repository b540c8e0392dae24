#!/usr/bin/env python
"""tools/debug_sine.py -- hunt for the intermittent 1.5e-4 deviation of tests/test_engine_gpu.py::test_sine_embed_and_box_refine:
repeat the call many times under different orderings and report the deviations seen."""
import collections
import sys

import torch

sys.path.insert(0, ".")
from memotr_b200 import kernels as K          # noqa: E402
from oracle import frame as oframe            # noqa: E402  (tool, not product)

g = torch.Generator().manual_seed(9)
i = torch.arange(128, dtype=torch.float32)
dim_t = 10000 ** (2 * torch.div(i, 2, rounding_mode="trunc") / 128)
pts = torch.rand(50, 4, generator=g)
want = oframe.pos_to_pos_embed(pts, num_pos_feats=128)


def err(got):
    return float((got - want).abs().max() / want.abs().max())


for mode in ("direct", "sync_between", "resident"):
    seen = collections.Counter()
    pd, dd = pts.to("cuda"), dim_t.to("cuda")
    for it in range(1500):
        if mode == "direct":
            got = K.sine_embed(pts.to("cuda"), dim_t.to("cuda")).cpu()
        elif mode == "sync_between":
            a, b = pts.to("cuda"), dim_t.to("cuda")
            torch.cuda.synchronize()
            got = K.sine_embed(a, b).cpu()
        else:
            got = K.sine_embed(pd, dd).cpu()
        if it % 7 == 0:      # churn the allocator a little, as a test suite does
            junk = [torch.randn(n, device="cuda") for n in (128, 200, 50 * 512, 4096)]
            del junk
        seen[f"{err(got):.2e}"] += 1
    print(mode, dict(seen))
# where is the deviation?  compare GPU sinf / division with float64 truth
got = K.sine_embed(pts.to("cuda"), dim_t.to("cuda")).cpu()
truth = oframe.pos_to_pos_embed(pts.double(), num_pos_feats=128)
print("gpu vs fp64", float((got.double() - truth).abs().max()), "cpu-fp32 oracle vs fp64", float((want.double() - truth).abs().max()))
