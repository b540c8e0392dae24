"""GPU parity tests of the windowed encoder gather (csrc/msda_window.cu, memotr_msda_forward_window).

The bar for this kernel is bit-equality with the global-memory gather (memotr_msda_forward_strided: same arithmetic,
csrc/msda_h16.cuh) on the same inputs -- whatever ends up staged -- plus the usual <= 1e-2 (bf16 output of an fp16 value
map) against the C restatement of the reference kernel (oracle/msda_oracle.c,
/root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299).
"""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _sequential_level_order(monkeypatch):
    """Small pyramids: keep the global-memory gather on its sequential (level-pair) order, the one the windowed kernel
    reproduces bit for bit (its level-split variant for launches too small to fill the GPU rounds differently)."""
    monkeypatch.setenv("MEMOTR_MSDA_NO_SPLIT", "1")


def K():
    from memotr_b200 import kernels
    return kernels


def _dev(shapes):
    shp = torch.as_tensor(shapes, dtype=torch.long)
    lsi = torch.cat((shp.new_zeros(1), shp.prod(1).cumsum(0)[:-1]))
    return shp.to(DEV), lsi.to(DEV), shp, lsi


def _case(shapes, H, Kp, seed, noise, valid=(1.0, 1.0)):
    value, vr, loc, attn, shift = synth.encoder_msda_inputs(shapes, H=H, K=Kp, seed=seed, noise_px=noise, valid=valid)
    return value.half().to(DEV), vr.to(DEV), loc.to(DEV), attn.to(DEV), shift, (value, loc, attn)


@pytest.mark.parametrize("shapes,H,Kp,noise,valid", [
    (synth.DANCETRACK_SHAPES, 8, 4, 0.5, (1.0, 1.0)),
    (synth.DANCETRACK_SHAPES, 8, 4, 0.3, (1333 / 1344, 0.875)),      # padded frame: valid ratios < 1
    (synth.BDD_SHAPES, 8, 4, 1.0, (1.0, 1.0)),
    (synth.BDD_SHAPES, 8, 8, 0.5, (1.0, 1.0)),
    (synth.BDD_SHAPES, 8, 16, 0.5, (1.0, 1.0)),
    (synth.BDD_SHAPES_L5, 8, 4, 0.5, (1.0, 1.0)),
    (synth.BDD_SHAPES_L5, 8, 8, 0.5, (0.9, 0.95)),
    (((40, 64), (20, 32), (10, 16)), 4, 2, 0.5, (1.0, 1.0)),
    (((33, 47), (17, 24), (9, 12), (5, 6)), 16, 4, 0.5, (1.0, 1.0)),   # odd extents, 16 heads, edge tiles everywhere
    (synth.SMALL_SHAPES, 8, 4, 0.5, (1.0, 1.0)),                        # windows larger than the coarse levels
])
def test_window_gather_bit_equal_to_global_gather_and_close_to_oracle(shapes, H, Kp, noise, valid):
    from oracle import msda as omsda
    value, vr, loc, attn, shift, cpu = _case(shapes, H, Kp, 11, noise, valid)
    shp_d, lsi_d, shp, lsi = _dev(shapes)
    L = len(shapes)
    want = K().msda_forward_strided(value, shp_d, lsi_d, n_heads=H, n_levels=L, n_points=Kp, loc=loc, attn=attn)
    stats = torch.zeros(2, dtype=torch.int64, device=DEV)
    full = (Kp - 1) / 2 + 3 * noise                 # the ring spreads the K points over +-(K-1)/2 pixels, plus the noise
    for kw in (dict(shift=shift, radius=full), dict(shift=None, radius=4.0), dict(shift=shift, radius=0.0),
               dict(shift=shift, radius=full, max_classes=1)):
        stats.zero_()
        got = K().msda_forward_window(value, shapes, vr, n_heads=H, n_points=Kp, loc=loc, attn=attn, stats=stats, **kw)
        assert torch.equal(got, want), kw
        plan = K().window_plan(shapes, H, Kp, kw["radius"], kw.get("max_classes", 0))
        n_win, n_glob = (int(v) for v in stats.tolist())
        staged_q = sum(min(c["tile"][0] * c["tiles"][0], shapes[c["level"]][1]) * min(c["tile"][1] * c["tiles"][1], shapes[c["level"]][0])
                       for c in plan["cls"])
        assert 0 < n_win + n_glob <= staged_q * H * L * Kp, (kw, plan)    # (points outside every image are counted nowhere)
        if kw.get("shift") is not None and kw["radius"] >= full and plan["classes"] and all(all(c["ww"]) for c in plan["cls"]):
            assert n_win >= 0.9 * (n_win + n_glob), (kw, n_win, n_glob)         # the hint does its job
    v32, l32, a32 = cpu
    S = v32.shape[0]
    ref = omsda.forward(v32.half().float().reshape(1, S, H, 32).numpy(), shp.numpy(), lsi.numpy(), l32[None].numpy(),
                        a32[None].numpy(), fma=True)[0]
    assert rel_err(want.float().cpu().numpy(), ref) < 6e-3


def test_window_gather_arbitrary_sampling_patterns():
    """Uniform random locations (models/ops/test.py recipe, incl. the border variant): nearly every tap leaves its window and
    is read from global memory -- same result."""
    shapes, H, Kp = synth.DANCETRACK_SHAPES, 8, 4
    shp_d, lsi_d, shp, lsi = _dev(shapes)
    S = int(shp.prod(1).sum())
    value, _, _, loc, attn = synth.msda_inputs(shapes, B=1, H=H, D=32, K=Kp, Lq=S, seed=5, border=True)
    value = (value.reshape(S, 256) * 100).half().to(DEV)
    loc, attn = loc[0].contiguous().to(DEV), attn[0].contiguous().to(DEV)
    vr = torch.ones(4, 2, device=DEV)
    want = K().msda_forward_strided(value, shp_d, lsi_d, n_heads=H, n_levels=4, n_points=Kp, loc=loc, attn=attn)
    got = K().msda_forward_window(value, shapes, vr, n_heads=H, n_points=Kp, loc=loc, attn=attn, radius=3.0)
    assert torch.equal(got, want)
    # non-finite / huge coordinates contribute nothing and do not fault
    loc2 = loc.clone()
    loc2[::7, 0, 0, 0, 0] = float("inf")
    loc2[1::7, 1, 1, 1, 1] = float("nan")
    loc2[2::7, 2, 2, 2] = -1e30
    want = K().msda_forward_strided(value, shp_d, lsi_d, n_heads=H, n_levels=4, n_points=Kp, loc=loc2, attn=attn)
    got = K().msda_forward_window(value, shapes, vr, n_heads=H, n_points=Kp, loc=loc2, attn=attn, radius=3.0)
    assert torch.equal(got, want) and torch.isfinite(got.float()).all()


def test_window_gather_strided_rows_and_pixel_stride():
    """Locations / weights read from the [locations | weights] rows of the projection GEMM, value map read through a pixel
    stride (one of several interleaved maps)."""
    shapes, H, Kp = synth.DANCETRACK_SHAPES, 8, 4
    value, vr, loc, attn, shift, _ = _case(shapes, H, Kp, 21, 0.5)
    shp_d, lsi_d, _, _ = _dev(shapes)
    S = value.shape[0]
    rows = torch.cat((loc.reshape(S, -1), attn.reshape(S, -1)), 1).contiguous()
    wide = torch.randn(S, 3 * 256, device=DEV).half()
    wide[:, 256:512] = value
    dense = K().msda_forward_strided(value, shp_d, lsi_d, n_heads=H, n_levels=4, n_points=Kp, loc=loc, attn=attn)
    got = K().msda_forward_window(wide[:, 256:512], shapes, vr, rows=rows, n_heads=H, n_points=Kp, shift=shift, radius=3.0)
    assert torch.equal(got, dense)


def test_window_gather_argument_errors():
    shapes, H = synth.SMALL_SHAPES, 8
    value, vr, loc, attn, shift, _ = _case(shapes, H, 4, 3, 0.5)
    with pytest.raises(RuntimeError, match="even number of points"):
        K().msda_forward_window(value, shapes, vr, n_heads=H, n_points=3, loc=loc, attn=attn)
    with pytest.raises(RuntimeError, match="radius"):
        K().msda_forward_window(value, shapes, vr, n_heads=H, n_points=4, loc=loc, attn=attn, radius=100.0)
    with pytest.raises(RuntimeError, match="sum of the level sizes"):
        K().msda_forward_window(value[:-1], shapes, vr, n_heads=H, n_points=4, loc=loc[:-1], attn=attn[:-1])
