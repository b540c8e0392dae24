"""Input projections (memotr_b200/input_proj.py, csrc/input_proj.cu) against the reference's own formulation -- nn.Conv2d +
nn.GroupNorm as models/memotr.py:66-78 builds them -- evaluated by stock PyTorch in fp32 (TF32 off, main.py:96-97)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err
from memotr_b200.input_proj import InputProj

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _state(cins, n_extra, seed):
    g = torch.Generator().manual_seed(seed)
    sd, l = {}, 0
    for cin in cins:
        sd[f"feature_projs.{l}.0.weight"] = torch.randn(256, cin, 1, 1, generator=g) / cin ** 0.5
        l += 1
    cin = cins[-1]
    for _ in range(n_extra):
        sd[f"feature_projs.{l}.0.weight"] = torch.randn(256, cin, 3, 3, generator=g) / (9 * cin) ** 0.5
        cin, l = 256, l + 1
    for i in range(l):
        sd[f"feature_projs.{i}.0.bias"] = torch.randn(256, generator=g) * 0.1
        sd[f"feature_projs.{i}.1.weight"] = 1 + 0.1 * torch.randn(256, generator=g)
        sd[f"feature_projs.{i}.1.bias"] = 0.1 * torch.randn(256, generator=g)
    return sd


def _reference(sd, feats):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    d = {k: v.to(DEV) for k, v in sd.items()}
    n = len({k.split(".")[1] for k in sd})
    out = []
    for l in range(n):
        x = feats[l] if l < len(feats) else (feats[-1] if l == len(feats) else out[-1])
        w = d[f"feature_projs.{l}.0.weight"]
        y = F.conv2d(x, w, d[f"feature_projs.{l}.0.bias"], stride=1 if w.shape[-1] == 1 else 2, padding=0 if w.shape[-1] == 1 else 1)
        out.append(F.group_norm(y, 32, d[f"feature_projs.{l}.1.weight"], d[f"feature_projs.{l}.1.bias"], 1e-5))
    return out


@pytest.mark.parametrize("shapes,cins,n_extra", [
    (((100, 168), (50, 84), (25, 42)), (512, 1024, 2048), 1),          # DanceTrack at 1333 x 800: ResNet-50 C3-C5 + one extra level
    (((37, 61), (19, 31)), (96, 160), 2),                               # odd sizes, two chained extra levels
])
def test_input_projections_match_conv2d_groupnorm(shapes, cins, n_extra):
    sd = _state(cins, n_extra, seed=len(shapes))
    g = torch.Generator().manual_seed(5)
    feats = [torch.randn(1, c, h, w, generator=g).to(DEV) for c, (h, w) in zip(cins, shapes)]
    got = InputProj(sd, DEV)(feats)
    want = _reference(sd, feats)
    assert len(got) == len(shapes) + n_extra
    for l, (a, b) in enumerate(zip(got, want)):
        assert a.shape == b.shape, (l, a.shape, b.shape)
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-5, l


def test_input_projections_reject_cpu_tensors_and_wrong_channels():
    sd = _state((64,), 0, seed=1)
    proj = InputProj(sd, DEV)
    with pytest.raises(RuntimeError):
        proj([torch.randn(1, 64, 8, 8)])
    with pytest.raises(RuntimeError):
        proj([torch.randn(1, 32, 8, 8, device=DEV)])
