#!/usr/bin/env python
"""tools/micro_msda.py -- MSDA forward micro-benchmark on one B200 (development aid; bench.py is the contract).

Encoder-shaped launches (Lq = S) of the fp16-value-map gathers on encoder-like sampling patterns (ring offsets + noise,
memotr_b200/synthetic.py:encoder_msda_inputs): the windowed kernel (csrc/msda_window.cu) against the global-memory kernel
(msda_fwd_h16), for the DanceTrack pyramid and the BDD100K sweep of BASELINE.json config 5 (K in {4,8,16} x L in {4,5}).
CUDA events on the current stream, L2 flushed between iterations, algorithmic bytes = fp16 value map + fp32 locations /
weights + bf16 output.  Writes gpurun_out/micro_msda.json.
  --ncu : run each kernel of the DanceTrack case exactly once (profiling target), no timing
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from memotr_b200 import kernels, synthetic as synth  # noqa: E402

DEV = "cuda"
PEAK = 6567.4
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]


def timeit(fn, iters=30, warmup=5, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def case(shapes, H, K, noise, radius=None, classes=0):
    value, vr, loc, attn, shift = synth.encoder_msda_inputs(shapes, H=H, K=K, seed=7, noise_px=noise)
    shp = torch.as_tensor(shapes, dtype=torch.long)
    lsi = torch.cat((shp.new_zeros(1), shp.prod(1).cumsum(0)[:-1]))
    d = dict(value=value.half().to(DEV), vr=vr.to(DEV), loc=loc.to(DEV), attn=attn.to(DEV), shift=shift, shp=shp.to(DEV),
             lsi=lsi.to(DEV), shapes=shapes, H=H, K=K, L=len(shapes), S=value.shape[0],
             radius=radius if radius is not None else min((K - 1) / 2 + 2.5 * noise + 0.5, 6.0), classes=classes)
    d["bytes"] = d["S"] * H * 32 * 2 + d["S"] * H * d["L"] * K * 12 + d["S"] * H * 32 * 2
    return d


def run_window(c, stats=None):
    return kernels.msda_forward_window(c["value"], c["shapes"], c["vr"], n_heads=c["H"], n_points=c["K"], loc=c["loc"],
                                       attn=c["attn"], shift=c["shift"], radius=c["radius"], max_classes=c["classes"], stats=stats)


def run_global(c):
    return kernels.msda_forward_strided(c["value"], c["shp"], c["lsi"], n_heads=c["H"], n_levels=c["L"], n_points=c["K"],
                                        loc=c["loc"], attn=c["attn"])


def main():
    if "--ncu" in sys.argv:
        c = case(synth.DANCETRACK_SHAPES, 8, 4, 0.15)
        for _ in range(3):
            run_window(c), run_global(c)
        torch.cuda.synchronize()
        return
    flush = torch.empty(256 * 1024 * 1024 // 4, device=DEV)   # 256 MB > 126 MB L2
    res = []
    cases = [("dancetrack_refinit_like", synth.DANCETRACK_SHAPES, 4, 0.15, None, 0),
             ("dancetrack_refinit_like_c1", synth.DANCETRACK_SHAPES, 4, 0.15, None, 1),
             ("dancetrack_noise0.5", synth.DANCETRACK_SHAPES, 4, 0.5, None, 0),
             ("dancetrack_noise1.0", synth.DANCETRACK_SHAPES, 4, 1.0, None, 0),
             ("dancetrack_noise1.0_r2.5", synth.DANCETRACK_SHAPES, 4, 1.0, 2.5, 0)]
    if "--quick" in sys.argv:
        cases = [cases[0], cases[3]]
    else:
        for L, shapes in ((4, synth.BDD_SHAPES), (5, synth.BDD_SHAPES_L5)):
            for K in (4, 8, 16):
                cases.append((f"bdd_L{L}K{K}", shapes, K, 0.15, None, 0))
    for name, shapes, K, noise, radius, classes in cases:
        c = case(shapes, 8, K, noise, radius, classes)
        stats = torch.zeros(2, dtype=torch.int64, device=DEV)
        a, b = run_window(c, stats), run_global(c)
        torch.cuda.synchronize()
        row = {"case": name, "S": c["S"], "L": c["L"], "K": K, "noise_px": noise, "radius": c["radius"], "bit_equal": bool(torch.equal(a, b)),
               "plan": kernels.window_plan(shapes, 8, K, c["radius"], classes), "bytes": c["bytes"]}
        nw, ng = (int(v) for v in stats.tolist())
        row["staged_points"], row["global_points_in_staged_units"] = nw, ng
        def run_generic():
            os.environ["MEMOTR_WINDOW_GENERIC"] = "1"      # (the library reads it at every launch)
            try:
                return run_window(c)
            finally:
                os.environ.pop("MEMOTR_WINDOW_GENERIC")
        row["generic_bit_equal"] = bool(torch.equal(run_generic(), b))
        for tag, fn in (("window", lambda: run_window(c)), ("window_generic", run_generic), ("global", lambda: run_global(c))):
            t, tmin = timeit(fn, flush=flush)
            row[f"{tag}_us"], row[f"{tag}_min_us"] = t, tmin
            row[f"{tag}_gbs"] = c["bytes"] / t / 1e3
            row[f"{tag}_frac"] = row[f"{tag}_gbs"] / PEAK
            t, _ = timeit(fn, flush=None)
            row[f"{tag}_l2warm_us"] = t
        print(json.dumps({k: v for k, v in row.items() if k != "plan"}), flush=True)
        res.append(row)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "micro_msda.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
