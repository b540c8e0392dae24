#!/usr/bin/env python
"""tools/overlap_probe.py -- feasibility probe (timing only, data hazards ignored): how much of the recurrent tail of frame k
(decoder + heads + tracker + updater: a latency chain on ~100 SMs) hides under the encoder of frame k + 1 when the two run on
two streams, as CUDA graphs, on one B200.  Prints sequential vs overlapped time per frame."""
import sys

import torch

sys.path.insert(0, ".")
from memotr_b200 import synthetic as synth          # noqa: E402
from memotr_b200.engine import FrameEngine           # noqa: E402

dev = torch.device("cuda:0")
cfg = synth.dancetrack_cfg()
sd = synth.reference_init_state_dict(cfg, seed=0)
x = synth.frame_inputs(cfg, synth.DANCETRACK_SHAPES, 100, seed=1, padded=True)
tracker = dict(det_score_thresh=2.0, track_score_thresh=0.0, miss_tolerance=30, result_score_thresh=0.5)
eng = FrameEngine(sd, cfg, synth.DANCETRACK_SHAPES, 100, dev, mode="bf16", tracker=tracker, pos_embed=dict(temperature=20))
eng.load_frame(x["srcs"], x["masks"], None, x["tracks"]["ref_pts"], x["tracks"]["query_embed"])
eng.load_tracks(x["tracks"])
eng.trk.reset(x["tracks"], max_obj_id=100)
for _ in range(3):
    eng.step()
torch.cuda.synchronize()


def graph_of(fn, stream):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        fn()
        stream.synchronize()
        with torch.cuda.graph(g, stream=stream):
            fn()
    return g


sa, sb = torch.cuda.Stream(dev), torch.cuda.Stream(dev, priority=-1)
g_enc, g_tail = graph_of(eng.encode, sa), graph_of(eng.step_tail, sb)
g_all = graph_of(eng.step, sa)
N = 200


def timed(body):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    body()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / N


def sequential():
    with torch.cuda.stream(sa):
        for _ in range(N):
            g_all.replay()
    torch.cuda.current_stream().wait_stream(sa)


def enc_only():
    with torch.cuda.stream(sa):
        for _ in range(N):
            g_enc.replay()
    torch.cuda.current_stream().wait_stream(sa)


def tail_only():
    with torch.cuda.stream(sb):
        for _ in range(N):
            g_tail.replay()
    torch.cuda.current_stream().wait_stream(sb)


def overlapped():
    # frame loop: tail(k) on sb and encode(k+1) on sa start together, both must finish before the next pair
    cur = torch.cuda.current_stream()
    sa.wait_stream(cur), sb.wait_stream(cur)
    for _ in range(N):
        with torch.cuda.stream(sb):
            g_tail.replay()
        with torch.cuda.stream(sa):
            g_enc.replay()
        sa.wait_stream(sb), sb.wait_stream(sa)
    cur.wait_stream(sa), cur.wait_stream(sb)


for name, fn in (("sequential step", sequential), ("encode only", enc_only), ("tail only", tail_only), ("overlapped", overlapped),
                 ("sequential step", sequential), ("overlapped", overlapped)):
    fn()
    print(f"{name:18s} {timed(fn):8.1f} us per frame")
