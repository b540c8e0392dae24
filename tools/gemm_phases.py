#!/usr/bin/env python
"""tools/gemm_phases.py -- where a persistent tcgen05 GEMM launch of the encoder spends its time: CUDA-event time of the
launch + the in-kernel clock64 stamps of every CTA (memotr_gemm_debug_stamps) for value_proj (fp16 out), output_proj (fp32
out) and the offsets+logits projection with the location / softmax epilogue, at the DanceTrack size.  GPU only."""
import json
import sys

import torch

sys.path.insert(0, ".")
from memotr_b200 import _lib, kernels as K, synthetic as synth   # noqa: E402

DEV = torch.device("cuda:0")
S = 22323
GHZ = 1.965e3


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


def phases(fn, n_cta=148):
    st = torch.zeros(20 * n_cta, dtype=torch.int64, device=DEV)
    _lib.check(_lib.lib().memotr_gemm_debug_stamps(_lib.ptr(st)), "stamps")
    fn()
    torch.cuda.synchronize()
    _lib.check(_lib.lib().memotr_gemm_debug_stamps(None), "stamps")
    s = st.view(n_cta, 20).cpu().double()
    out = {"cta_total_us_median": round(float((s[:, 19] - s[:, 0]).median()) / GHZ, 2),
           "cta_total_us_max": round(float((s[:, 19] - s[:, 0]).max()) / GHZ, 2)}
    tiles = []
    for i in range(6):
        ok = s[:, 3 + 3 * i] > 0
        if i >= 4 and "tiles_5plus" not in out:
            out["tiles_5plus"] = "stamps cover the first 6 tiles of a CTA only"
        if ok.sum() == 0:
            break
        a = s[ok]
        tiles.append({"ctas": int(ok.sum()),
                      "wait_staging_us": round(float((a[:, 1 + 3 * i] - (a[:, 3 * i] if i else a[:, 0])).median()) / GHZ, 2),
                      "wait_acc_us": round(float((a[:, 2 + 3 * i] - a[:, 1 + 3 * i]).median()) / GHZ, 2),
                      "epilogue_us": round(float((a[:, 3 + 3 * i] - a[:, 2 + 3 * i]).median()) / GHZ, 2)})
    out["tiles"] = tiles
    last = torch.stack([s[:, 3 + 3 * i] for i in range(6)], 1).max(1).values
    out["final_store_us"] = round(float((s[:, 19] - last).median()) / GHZ, 2)
    return out


def main():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(S, 256, generator=g).bfloat16().to(DEV)
    res = {}
    for name, N, odt in (("value_proj_f16", 256, torch.float16), ("out_proj_f32", 256, torch.float32), ("n384_f32_plain", 384, torch.float32),
                         ("dec_value_all_f16", 1536, torch.float16)):
        w = (torch.randn(N, 256, generator=g) / 16).bfloat16().to(DEV)
        b = torch.randn(N, generator=g).to(DEV)
        out = torch.empty(S, N, dtype=odt, device=DEV)
        fn = lambda: K.linear(x, w, b, out=out, path="tc")   # noqa: E731
        res[name] = {"launch_us": round(timeit(fn), 2), **phases(fn)}
    shapes = synth.DANCETRACK_SHAPES
    w = (torch.randn(384, 256, generator=g) / 16).bfloat16().to(DEV)
    b = torch.randn(384, generator=g).to(DEV)
    vr = torch.ones(4, 2, device=DEV)
    lsi = [0]
    for h, wd in shapes[:-1]:
        lsi.append(lsi[-1] + h * wd)
    fn = lambda: K.linear_msda_prep(x, w, b, shapes, lsi, vr, 8, 4, 4)   # noqa: E731
    res["offsets_logits_prep"] = {"launch_us": round(timeit(fn), 2), **phases(fn)}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
