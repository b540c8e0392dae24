#!/usr/bin/env python
"""tools/prof_step.py -- run a few eager (non-graph) hot-path steps; the target command for `ncu --set full` captures.
Usage: python tools/prof_step.py [bf16|fp32] [n_steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from memotr_b200 import synthetic as synth  # noqa: E402
from memotr_b200.engine import FrameEngine  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = synth.dancetrack_cfg()
sd = synth.hot_path_state_dict(cfg, seed=0)
x = synth.frame_inputs(cfg, synth.DANCETRACK_SHAPES, 100, seed=1)
eng = FrameEngine(sd, cfg, synth.DANCETRACK_SHAPES, 100, "cuda", mode=mode)
eng.load_frame(x["srcs"], x["masks"], x["pos"], x["tracks"]["ref_pts"], x["tracks"]["query_embed"])
eng.load_tracks(x["tracks"])
for _ in range(n):
    eng.step()
torch.cuda.synchronize()
print("launches per step:", eng.launches // n)
