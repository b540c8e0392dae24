import os, sys, torch, json
sys.path.insert(0, "/root/repo")
from memotr_b200 import kernels
sys.path.insert(0, "/root/repo/tools")
from micro_gemm import timeit
M, Hd = 22323, 2048
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
x = torch.randn(M, 256, device="cuda").bfloat16()
w1 = (torch.randn(Hd, 256, device="cuda") / 16).bfloat16()
w2 = (torch.randn(256, Hd, device="cuda") / Hd ** 0.5).bfloat16()
b1, b2 = torch.randn(Hd, device="cuda"), torch.randn(256, device="cuda")
out = torch.empty(M, 256, device="cuda")
for sp in ("1", "2", "4", "8", "16", None):
    if sp: os.environ["MEMOTR_MLP_SPLIT"] = sp
    else: os.environ.pop("MEMOTR_MLP_SPLIT", None)
    print("split", sp, "us", round(timeit(lambda: kernels.mlp2(x, w1, b1, w2, b2, out=out), flush=flush), 1), flush=True)
