import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from memotr_b200 import kernels
from oracle import frame as oframe
g = torch.Generator().manual_seed(9)
i = torch.arange(128, dtype=torch.float32)
dim_t = 10000 ** (2 * torch.div(i, 2, rounding_mode="trunc") / 128)
pts = torch.rand(50, 4, generator=g)
want = oframe.pos_to_pos_embed(pts, num_pos_feats=128)
got = kernels.sine_embed(pts.cuda(), dim_t.cuda()).cpu()
d = (got - want).abs()
idx = torch.topk(d.flatten(), 8).indices
for k in idx.tolist():
    n, col = divmod(k, 512); c, r = divmod(col, 128); j, sc = divmod(r, 2)
    p = pts[n, c].item(); dt = dim_t[2 * j].item()
    e32 = np.float32(np.float32(p) * np.float32(6.283185307179586)) / np.float32(dt)
    print(n, c, j, "sin" if sc == 0 else "cos", "p", p, "dim", dt, "e", float(e32), "got", got[n, col].item(), "want", want[n, col].item(),
          "f64", (np.sin if sc == 0 else np.cos)(np.float64(p) * 2 * np.pi / np.float64(dt)))
gw = oframe.pos_to_pos_embed(pts.cuda(), num_pos_feats=128).cpu()
print("torch-cuda vs torch-cpu", (gw - want).abs().max().item(), "ours vs torch-cuda", (got - gw).abs().max().item())
