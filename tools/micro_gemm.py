#!/usr/bin/env python
"""tools/micro_gemm.py -- GEMM micro-benchmark on one B200 (development aid): the hot-path shapes through
memotr_linear (tcgen05 path, bf16) next to torch.matmul (cuBLAS) on the same tensors.  CUDA events, L2 flush."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from memotr_b200 import kernels  # noqa: E402

DEV = "cuda"


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
    return ts[len(ts) // 2]


def main():
    flush = torch.empty(256 * 1024 * 1024 // 4, device=DEV)
    res = []
    S = 22323
    for name, M, N, K, act, odt in [("value_proj", S, 256, 256, None, torch.bfloat16),
                                    ("offs_logits", S, 384, 256, None, torch.float32),
                                    ("ffn1", S, 2048, 256, "relu", torch.bfloat16),
                                    ("ffn2", S, 256, 2048, None, torch.bfloat16),
                                    ("dec_value_all", S, 1536, 256, None, torch.bfloat16),
                                    ("dec_small", 400, 256, 256, None, torch.bfloat16),
                                    ("dec_ffn1", 400, 2048, 256, "relu", torch.bfloat16)]:
        x = torch.randn(M, K, device=DEV).bfloat16()
        w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
        b = torch.randn(N, device=DEV)
        out = torch.empty(M, N, device=DEV, dtype=odt)
        t = timeit(lambda: kernels.linear(x, w, b, act=act, out=out, path="tc"), flush=flush)
        t_ref = timeit(lambda: torch.nn.functional.linear(x, w, b.bfloat16()), flush=flush)
        row = {"gemm": name, "M": M, "N": N, "K": K, "ours_us": t, "ours_tflops": 2 * M * N * K / t / 1e6,
               "cublas_us": t_ref, "cublas_tflops": 2 * M * N * K / t_ref / 1e6}
        if M > 1000:
            xf, wf = x.float(), w.float()
            outf = torch.empty(M, N, device=DEV)
            row["simt_fp32_us"] = timeit(lambda: kernels.linear(xf, wf, b, act=act, out=outf), iters=5, flush=flush)
        print(json.dumps(row), flush=True)
        res.append(row)
    for name, M, Hd in (("ffn_fused_enc", S, 2048), ("ffn_fused_dec", 400, 2048), ("mlp_fused_dec_h256", 400, 256),
                        ("mlp_fused_upd_h256", 100, 256)):
        x = torch.randn(M, 256, device=DEV).bfloat16()
        w1 = (torch.randn(Hd, 256, device=DEV) / 16).bfloat16()
        w2 = (torch.randn(256, Hd, device=DEV) / Hd ** 0.5).bfloat16()
        b1, b2 = torch.randn(Hd, device=DEV), torch.randn(256, device=DEV)
        out = torch.empty(M, 256, device=DEV)
        t = timeit(lambda: kernels.mlp2(x, w1, b1, w2, b2, out=out), flush=flush)
        extra = {}
        for cs in ("1", "2", "4"):
            os.environ["MEMOTR_MLP_CLUSTER"] = cs
            extra[f"fused_cluster{cs}_us"] = timeit(lambda: kernels.mlp2(x, w1, b1, w2, b2, out=out), flush=flush)
        os.environ.pop("MEMOTR_MLP_CLUSTER", None)
        hid = torch.empty(M, Hd, device=DEV, dtype=torch.bfloat16)

        def two():
            kernels.linear(x, w1, b1, act="relu", out=hid, path="tc")
            kernels.linear(hid, w2, b2, out=out, path="tc")
        t2 = timeit(two, flush=flush)
        row = {"gemm": name, "M": M, "fused_us": t, "Hd": Hd, "fused_tflops": 2 * 2 * M * Hd * 256 / t / 1e6, "two_gemms_us": t2, **extra}
        print(json.dumps(row), flush=True)
        res.append(row)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "micro_gemm.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
