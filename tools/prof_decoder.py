"""tools/prof_decoder.py -- run a few eager bf16 engine forwards at the DanceTrack size (target for `ncu -k regex:decoder_fused`)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memotr_b200 import synthetic as synth
from memotr_b200.engine import FrameEngine

cfg = synth.dancetrack_cfg()
sd = synth.hot_path_state_dict(cfg, seed=0)
x = synth.frame_inputs(cfg, synth.DANCETRACK_SHAPES, 100, seed=1)
eng = FrameEngine(sd, cfg, synth.DANCETRACK_SHAPES, 100, "cuda", mode="bf16")
eng.load_frame(x["srcs"], x["masks"], x["pos"], x["tracks"]["ref_pts"], x["tracks"]["query_embed"])
for _ in range(3):
    eng.forward()
torch.cuda.synchronize()
