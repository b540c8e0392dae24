"""memotr_b200/input_proj.py -- the input projections of MeMOTR on our kernels.

Mirrors what /root/reference/models/memotr.py:66-78 builds (`feature_projs`: per backbone level Conv2d(k=1) + GroupNorm(32, C),
per extra level Conv2d(k=3, s=2, p=1) + GroupNorm) and :107-123 applies: level l < n_backbone from the l-th backbone map, the
first extra level from the LAST backbone map, further extra levels from the previous projected map.  Batch 1, fp32; consumes the
reference's `state_dict` keys `feature_projs.<l>.0.{weight,bias}` / `.1.{weight,bias}`; output = the `srcs` list (each (1, C,
H, W)) that the transformer / FrameEngine.load_frame takes.  csrc/input_proj.cu; no CPU fallback."""
import torch

from . import _lib


class InputProj:
    def __init__(self, state_dict, device="cuda", prefix="feature_projs", groups=32, eps=1e-5):
        self.dev, self.groups, self.eps = torch.device(device), groups, eps
        self.levels = []
        l = 0
        while f"{prefix}.{l}.0.weight" in state_dict:
            w = state_dict[f"{prefix}.{l}.0.weight"].to(self.dev, torch.float32).contiguous()
            self.levels.append(dict(k=int(w.shape[-1]), cin=int(w.shape[1]), cout=int(w.shape[0]),
                                    w=w.reshape(w.shape[0], -1).contiguous(),
                                    b=state_dict[f"{prefix}.{l}.0.bias"].to(self.dev, torch.float32).contiguous(),
                                    g=state_dict[f"{prefix}.{l}.1.weight"].to(self.dev, torch.float32).contiguous(),
                                    beta=state_dict[f"{prefix}.{l}.1.bias"].to(self.dev, torch.float32).contiguous()))
            l += 1
        if not self.levels:
            raise KeyError(f"no '{prefix}.<l>.0.weight' in the state dict")

    def _project(self, L, x):
        """x (1, Cin, H, W) fp32 on the device -> (1, Cout, Ho, Wo)."""
        _lib.require_cuda(x=x)
        _, cin, H, W = x.shape
        if cin != L["cin"] or x.shape[0] != 1 or x.dtype != torch.float32:
            raise RuntimeError(f"InputProj: expected (1, {L['cin']}, H, W) fp32, got {tuple(x.shape)} {x.dtype}")
        x = x.contiguous()
        lib, st = _lib.lib(), _lib.stream_ptr(x.device)
        with torch.cuda.device(x.device):
            if L["k"] == 1:
                Ho, Wo, src, K = H, W, x, cin
            elif L["k"] == 3:
                Ho, Wo, K = (H - 1) // 2 + 1, (W - 1) // 2 + 1, 9 * cin
                src = torch.empty((K, Ho * Wo), dtype=torch.float32, device=x.device)
                _lib.check(lib.memotr_im2col_3x3s2(_lib.ptr(x), cin, H, W, _lib.ptr(src), st), "memotr_im2col_3x3s2")
            else:
                raise RuntimeError(f"InputProj: kernel size {L['k']} is not one the reference builds")
            y = torch.empty((1, L["cout"], Ho, Wo), dtype=torch.float32, device=x.device)
            _lib.check(lib.memotr_conv_gemm(_lib.ptr(L["w"]), _lib.ptr(src), _lib.ptr(L["b"]), _lib.ptr(y), L["cout"], Ho * Wo, K, st),
                       "memotr_conv_gemm")
            _lib.check(lib.memotr_groupnorm_cm(_lib.ptr(y), _lib.ptr(L["g"]), _lib.ptr(L["beta"]), self.groups, L["cout"], Ho * Wo,
                                               float(self.eps), st), "memotr_groupnorm_cm")
        return y

    def __call__(self, features):
        """features: the backbone maps, list of (1, C_l, H_l, W_l).  -> srcs, one per projection (memotr.py:107-123)."""
        nb = len(features)
        if nb > len(self.levels):
            raise RuntimeError(f"InputProj: {nb} backbone maps but {len(self.levels)} projections")
        srcs = [self._project(self.levels[l], features[l]) for l in range(nb)]
        for l in range(nb, len(self.levels)):
            srcs.append(self._project(self.levels[l], features[-1] if l == nb else srcs[-1]))
        return srcs
