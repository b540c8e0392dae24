#!/usr/bin/env python
"""tools/micro_msda.py -- MSDA micro-benchmark on one B200 (development aid; bench.py is the contract).

Times memotr_msda_forward/backward (fp32, bf16) and, when oracle/_ref travelled, the reference CUDA op on the same
tensors: encoder-shaped (Lq = S = 22323) and decoder-shaped (Lq = 400) calls of the DanceTrack config, with
  * "uniform" sampling locations (models/ops/test.py recipe: rand in [0,1) -- worst-case locality), and
  * "encoder" locations (pixel-centre reference points + Gaussian offsets of a few pixels -- what a trained encoder
    produces, ms_deform_attn.py:115-117).
CUDA events on the current stream, L2 flushed between iterations, algorithmic bytes per BASELINE.md section 3.
Writes gpurun_out/micro_msda.json.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import memotr_b200  # noqa: E402
from oracle import synth  # noqa: E402

DEV = "cuda"
PEAK = 6567.4
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]


def timeit(fn, iters=20, warmup=3, flush=None):
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


def encoder_locs(shapes, H, K, sigma_px, seed):
    g = torch.Generator().manual_seed(seed)
    refs = []
    for (h, w) in shapes:
        ys, xs = torch.meshgrid(torch.linspace(0.5, h - 0.5, h) / h, torch.linspace(0.5, w - 0.5, w) / w, indexing="ij")
        refs.append(torch.stack((xs.reshape(-1), ys.reshape(-1)), -1))
    ref = torch.cat(refs, 0)                                  # (S, 2)
    S, L = ref.shape[0], len(shapes)
    off = torch.randn(1, S, H, L, K, 2, generator=g) * sigma_px
    wh = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32)
    return ref[None, :, None, None, None, :] + off / wh[None, None, None, :, None, :]


def main():
    shapes = synth.DANCETRACK_SHAPES
    S = sum(h * w for h, w in shapes)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=DEV)   # 256 MB > 126 MB L2
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    MSDA = None
    if os.path.exists(os.path.join(ref_dir, "MultiScaleDeformableAttention.so")):
        sys.path.insert(0, ref_dir)
        import MultiScaleDeformableAttention as MSDA
    res = []
    for tag, Lq, K in (("enc", S, 4), ("dec", 400, 4), ("enc_K8", S, 8), ("enc_K16", S, 16)):
        value, shp, lsi, loc, attn = (x.to(DEV) for x in synth.msda_inputs(shapes, Lq=Lq, K=K, seed=1))
        locs = {"uniform": loc}
        if Lq == S:
            locs["encoder"] = encoder_locs(shapes, 8, K, 2.0, 2).to(DEV).contiguous()
        for lname, lc in locs.items():
            nbytes = (S * 256 + Lq * 8 * 4 * K * 3 + Lq * 256) * 4
            row = {"case": tag, "loc": lname, "Lq": Lq, "K": K, "fp32_bytes": nbytes}
            if Lq == S:
                os.environ["MEMOTR_MSDA_MAPPING"] = "linear"
                row["ours_fwd_fp32_linear_us"], _ = timeit(
                    lambda: memotr_b200.ms_deform_attn_forward(value, shp, lsi, lc, attn, 64), flush=flush)
                os.environ["MEMOTR_MSDA_MAPPING"] = "tiled"
            t, tmin = timeit(lambda: memotr_b200.ms_deform_attn_forward(value, shp, lsi, lc, attn, 64), flush=flush)
            os.environ.pop("MEMOTR_MSDA_MAPPING", None)
            row["ours_fwd_fp32_us"], row["ours_fwd_fp32_min_us"] = t, tmin
            row["ours_fwd_fp32_gbs"] = nbytes / t / 1e3
            row["ours_fwd_fp32_frac"] = row["ours_fwd_fp32_gbs"] / PEAK
            vb, lb, ab = value.bfloat16(), lc.bfloat16(), attn.bfloat16()
            t, tmin = timeit(lambda: memotr_b200.ms_deform_attn_forward(vb, shp, lsi, lb, ab, 64), flush=flush)
            row["ours_fwd_bf16_us"] = t
            row["ours_fwd_bf16_gbs"] = nbytes / 2 / t / 1e3
            from memotr_b200 import kernels
            for nm, vv in (("fp32", value.reshape(S, 256)), ("bf16", value.reshape(S, 256).bfloat16())):
                for kern, U in (("v1", "2"), ("v2", "1"), ("v2", "2"), ("v2", "4")):
                    os.environ["MEMOTR_MSDA_KERNEL"] = kern
                    os.environ["MEMOTR_MSDA_U"] = U
                    t, _ = timeit(lambda: kernels.msda_forward_ex(vv, shp, lsi, lc[0], attn[0], 8), flush=flush)
                    row[f"ex_{kern}{'_U' + U if kern == 'v2' else ''}_{nm}_us"] = t
            os.environ.pop("MEMOTR_MSDA_KERNEL", None)
            os.environ.pop("MEMOTR_MSDA_U", None)
            vh16 = value.reshape(S, 256).half()
            row["ex_v4_fp16_us"], _ = timeit(lambda: kernels.msda_forward_ex(vh16, shp, lsi, lc[0], attn[0], 8), flush=flush)
            if K in (4, 8):
                vb16 = value.reshape(S, 256).bfloat16()
                row["pairs_layout_us"], _ = timeit(lambda: kernels.msda_pairs_layout(vb16, shp, lsi, 8), flush=flush)
                pr = kernels.msda_pairs_layout(vb16, shp, lsi, 8)
                row["pairs_gather_us"], _ = timeit(lambda: kernels.msda_forward_pairs(pr, shp, lsi, lc[0], attn[0]), flush=flush)
            if K == 4:
                go = torch.randn(1, Lq, 256, device=DEV)
                t, _ = timeit(lambda: memotr_b200.ms_deform_attn_backward(value, shp, lsi, lc, attn, go, 64), flush=flush)
                row["ours_bwd_fp32_us"] = t
            if MSDA is not None:
                t, tmin = timeit(lambda: MSDA.ms_deform_attn_forward(value, shp, lsi, lc, attn, 64), flush=flush)
                row["ref_fwd_fp32_us"], row["ref_fwd_fp32_gbs"] = t, nbytes / t / 1e3
                if K == 4:
                    t, _ = timeit(lambda: MSDA.ms_deform_attn_backward(value, shp, lsi, lc, attn, go, 64), flush=flush)
                    row["ref_bwd_fp32_us"] = t
            print(json.dumps(row), flush=True)
            res.append(row)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "micro_msda.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
