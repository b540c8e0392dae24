"""tools/prof_mlp2.py -- launch the fused FFN kernel a few times at the encoder shape (target for `ncu -k regex:mlp2`)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memotr_b200 import kernels

M, Hd = 22323, 2048
x = torch.randn(M, 256, device="cuda").bfloat16()
w1 = (torch.randn(Hd, 256, device="cuda") / 16).bfloat16()
w2 = (torch.randn(256, Hd, device="cuda") / Hd ** 0.5).bfloat16()
b1, b2 = torch.randn(Hd, device="cuda"), torch.randn(256, device="cuda")
out = torch.empty(M, 256, device="cuda")
for _ in range(3):
    kernels.mlp2(x, w1, b1, w2, b2, out=out)
torch.cuda.synchronize()
