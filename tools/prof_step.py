#!/usr/bin/env python
"""tools/prof_step.py -- one eager (non-graph) hot-path step of the bench configuration between cudaProfilerStart / Stop:
the target of `ncu --profile-from-start off ...` (launch list without any kernel-name filter, or --set full captures).
Usage: python tools/prof_step.py [bf16|fp32] [n_steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from memotr_b200 import synthetic as synth  # noqa: E402
from memotr_b200.engine import FrameEngine  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = synth.dancetrack_cfg()
sd = synth.reference_init_state_dict(cfg, seed=0)
x = synth.frame_inputs(cfg, synth.DANCETRACK_SHAPES, 100, seed=1, padded=True)
tracker = dict(det_score_thresh=2.0, track_score_thresh=0.0, miss_tolerance=30, result_score_thresh=0.5)
eng = FrameEngine(sd, cfg, synth.DANCETRACK_SHAPES, 100, "cuda", mode=mode, tracker=tracker, pos_embed=dict(temperature=20))
eng.load_frame(x["srcs"], x["masks"], None, x["tracks"]["ref_pts"], x["tracks"]["query_embed"])
eng.load_tracks(x["tracks"])
eng.trk.reset(x["tracks"], max_obj_id=100)
for _ in range(2):
    eng.step()
torch.cuda.synchronize()
eng.launches = 0
torch.cuda.profiler.start()
for _ in range(n):
    eng.step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches per step:", eng.launches // n)
