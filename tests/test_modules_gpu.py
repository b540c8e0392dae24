"""GPU tests of the nn.Module mirrors (autograd path): forward against the reference modules' golden outputs and
backward against autograd through the functional CPU oracle (which differentiates through grid_sample, the
reference's own CPU formulation of the sampling core)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch
from torch import nn

from conftest import GOLDEN, rel_err
from memotr_b200 import modules, synthetic as synth
from oracle import frame as oframe

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _load(tag):
    g = np.load(os.path.join(GOLDEN, f"frame_{tag}.npz"))
    nt, seed_w, seed_x, padded = (int(v) for v in g["meta"])
    cfg = synth.small_cfg()
    sd = synth.hot_path_state_dict(cfg, seed=seed_w)
    x = synth.frame_inputs(cfg, synth.SMALL_SHAPES, nt, seed=seed_x, padded=bool(padded))
    rc = oframe.to_reference_config(cfg)
    tr, qu = modules.build_transformer(rc), modules.build_query_updater(rc)
    bbox = nn.ModuleList([modules.MLP(256, 256, 4, 3) for _ in range(cfg["n_dec_layers"])])
    tr.set_refine_bbox_embed(bbox)
    tsd = {k[len("transformer."):]: v for k, v in sd.items() if k.startswith("transformer.")}
    tsd.update({"decoder.bbox_embed." + k[len("bbox_embed."):]: v for k, v in sd.items() if k.startswith("bbox_embed.")})
    tr.load_state_dict(tsd, strict=True)
    qu.load_state_dict({k[len("query_updater."):]: v for k, v in sd.items() if k.startswith("query_updater.")}, strict=True)
    return g, cfg, sd, x, tr.to(DEV).eval(), qu.to(DEV).eval()


def _queries(sd, x):
    ref = torch.cat((sd["det_anchor"], x["tracks"]["ref_pts"]), 0)[None]
    emb = torch.cat((sd["det_query_embed"], x["tracks"]["query_embed"]), 0)[None]
    return emb, ref, torch.zeros(1, ref.shape[1], dtype=torch.bool)


@pytest.mark.parametrize("tag", ["small", "small_padded"])
def test_transformer_and_updater_modules_match_reference_golden(tag):
    g, cfg, sd, x, tr, qu = _load(tag)
    emb, ref, qmask = _queries(sd, x)
    d = lambda ts: [t.to(DEV) for t in ts]                                      # noqa: E731
    with torch.no_grad():
        outs, init_ref, refs, queries = tr(d(x["srcs"]), d(x["masks"]), d(x["pos"]), emb.to(DEV), ref.to(DEV), qmask.to(DEV))
    assert rel_err(outs[-1].cpu().numpy(), g["outputs"]) < 1e-4
    assert rel_err(queries[1:].cpu().numpy(), g["aux_queries"]) < 1e-4
    assert rel_err(oframe.inverse_sigmoid(refs[-2]).cpu().numpy(), g["last_ref_pts"]) < 1e-4
    t = SimpleNamespace(**{k: v.clone().to(DEV) for k, v in x["tracks"].items()})
    with torch.no_grad():
        qu.update_tracks_embedding([t])
    for k in ("ref_pts", "query_embed", "long_memory", "last_output"):
        assert rel_err(getattr(t, k).cpu().numpy(), g["upd_" + k]) < 1e-4, k


def test_training_path_gradients_match_cpu_oracle_autograd():
    """loss = <outputs, R>: gradients w.r.t. the input feature maps and a few parameters through our MSDA backward
    kernels vs torch autograd through the functional oracle on the CPU."""
    g, cfg, sd, x, tr, _ = _load("small_padded")
    tr.train()                                           # DROPOUT is 0.0 in every shipped config
    emb, ref, qmask = _queries(sd, x)
    R = torch.randn(cfg["n_dec_layers"], 1, emb.shape[1], 256, generator=torch.Generator().manual_seed(3))
    srcs_gpu = [s.to(DEV).requires_grad_(True) for s in x["srcs"]]
    outs, _, _, _ = tr(srcs_gpu, [m.to(DEV) for m in x["masks"]], [p.to(DEV) for p in x["pos"]], emb.to(DEV),
                       ref.to(DEV), qmask.to(DEV))
    (outs * R.to(DEV)).sum().backward()
    # oracle
    sd_c = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    srcs_c = [s.clone().requires_grad_(True) for s in x["srcs"]]
    o_outs, _, _, _, _ = oframe.transformer(sd_c, srcs_c, x["masks"], x["pos"], emb, ref, qmask, cfg)
    (o_outs * R).sum().backward()
    for a, b in zip(srcs_gpu, srcs_c):
        assert rel_err(a.grad.cpu().numpy(), b.grad.numpy()) < 2e-4
    named = dict(tr.named_parameters())
    for key in ("encoder.layers.0.self_attn.sampling_offsets.weight", "encoder.layers.1.self_attn.value_proj.weight",
                "decoder.layers.2.cross_attn.attention_weights.bias", "decoder.layers.0.self_attn.in_proj_weight",
                "level_embed", "decoder.ref_point_head.layers.0.weight"):
        assert rel_err(named[key].grad.cpu().numpy(), sd_c["transformer." + key].grad.numpy()) < 2e-4, key


# ------------------------------------------------------------------------------------------------ MeMOTR surface
class _NT:
    """Duck-typed NestedTensor (utils/nested_tensor.py:9-59)."""

    def __init__(self, tensors, masks):
        self.tensors, self.masks = tensors, masks

    def decompose(self):
        return self.tensors, self.masks


class _Tracks:
    """Duck-typed TrackInstances (structures/track_instances.py:7-129): the fields and the two operations the model uses."""
    FIELDS = ("ref_pts", "query_embed", "ids", "boxes", "labels", "logits", "output_embed", "disappear_time", "iou",
              "last_output", "long_memory")

    def __init__(self, frame_height=1.0, frame_width=1.0, hidden_dim=256, num_classes=1, n=0, device="cpu"):
        self.hidden_dim, self.num_classes = hidden_dim, num_classes
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=device)        # noqa: E731
        self.ref_pts, self.boxes = z(n, 4), z(n, 4)
        self.query_embed, self.output_embed, self.last_output, self.long_memory = (z(n, hidden_dim) for _ in range(4))
        self.ids, self.labels, self.disappear_time = (z(n, dt=torch.long) for _ in range(3))
        self.logits, self.iou = z(n, num_classes), z(n)

    def __len__(self):
        return self.query_embed.shape[0]

    def __getitem__(self, item):
        r = _Tracks(hidden_dim=self.hidden_dim, num_classes=self.num_classes)
        for k in self.FIELDS:
            setattr(r, k, getattr(self, k)[item])
        return r

    @staticmethod
    def cat_tracked_instances(a, b):
        r = _Tracks(hidden_dim=a.hidden_dim, num_classes=a.num_classes)
        for k in _Tracks.FIELDS:
            setattr(r, k, torch.cat((getattr(a, k), getattr(b, k))))
        return r


class _FakeBackbone(nn.Module):
    """Stands in for BackboneWithPE (models/backbone.py:101-125): hands MeMOTR.forward a synthetic pyramid."""

    def __init__(self, srcs, masks, pos):
        super().__init__()
        self.feats, self.pos = [_NT(s, m) for s, m in zip(srcs[:3], masks[:3])], pos

    def position_embedding(self, nt):
        return self.pos[3]

    def n_inter_layers(self):
        return 3

    def n_inter_channels(self):
        return [8, 8, 8]

    def forward(self, frame):
        return self.feats, list(self.pos[:3])


class _Const(nn.Module):
    def __init__(self, t):
        super().__init__()
        self.t = t

    def forward(self, _):
        return self.t


def _memotr(tag, mode):
    from memotr_b200 import memotr as mm
    import test_numerics_cpu as tn
    g, cfg, sd, x, shapes, nt = tn.load_case(tag)
    d = lambda ts: [t.to(DEV) for t in ts]                                      # noqa: E731
    srcs, masks, pos = d(x["srcs"]), d(x["masks"]), d(x["pos"])
    rc = dict(oframe.to_reference_config(cfg), DATASET="DanceTrack", ENGINE_MODE=mode)
    model = mm.build(rc, _FakeBackbone(srcs, masks, pos))
    model.feature_projs = nn.ModuleList([nn.Identity(), nn.Identity(), nn.Identity(), _Const(srcs[3])])
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith(("feature_projs", "transformer.decoder.bbox_embed", "backbone")) for k in missing)
    h3, w3 = shapes[3]
    frame = _NT(torch.zeros(1, 3, h3 * 8, w3 * 8, device=DEV), masks[3].repeat_interleave(8, 1).repeat_interleave(8, 2))
    tr = _Tracks(hidden_dim=256, num_classes=cfg["num_classes"], n=nt, device=DEV)
    tr.ref_pts, tr.query_embed = x["tracks"]["ref_pts"].to(DEV), x["tracks"]["query_embed"].to(DEV)
    return g, cfg, x, model.to(DEV).eval(), frame, tr


@pytest.mark.parametrize("tag,mode,tol", [("small", "fp32", 1e-4), ("small_padded", "fp32", 1e-4), ("full_refinit", "fp32", 1e-4),
                                          ("full_refinit", "bf16", 1e-2)])
def test_memotr_module_eval_routes_through_the_engine_and_matches_reference(tag, mode, tol):
    """`MeMOTR.forward(frame, tracks)` -- the reference's model surface (models/memotr.py:97-195) -- in eval mode: the
    fast engine behind it (capacity above the number of tracks: padded rows), the reference's output dict, against the
    outputs of the reference nn.Modules."""
    g, cfg, x, model, frame, tr = _memotr(tag, mode)
    with torch.no_grad():
        res = model(frame=frame, tracks=[tr])
    assert len(model._engines) == 1 and next(iter(model._engines.values())).nt >= len(tr) and len(model._engines) == 1
    n = cfg["n_det_queries"] + len(tr)
    assert res["pred_logits"].shape == (1, n, cfg["num_classes"]) and res["query_mask"].shape == (1, n)
    got = {"pred_logits": res["pred_logits"], "pred_bboxes": res["pred_bboxes"], "last_ref_pts": res["last_ref_pts"],
           "init_ref_pts": res["init_ref_pts"], "outputs": res["outputs"],
           "aux_logits": torch.stack([a["pred_logits"] for a in res["aux_outputs"]]),
           "aux_bboxes": torch.stack([a["pred_bboxes"] for a in res["aux_outputs"]]),
           "aux_queries": torch.stack([a["queries"] for a in res["aux_outputs"]])}
    for k, v in got.items():
        assert rel_err(v.cpu().numpy(), g[k]) < tol, (k, rel_err(v.cpu().numpy(), g[k]))
    with torch.no_grad():                       # a second frame reuses the engine (graph replay) and gives the same answer
        res2 = model(frame=frame, tracks=[tr])
    # (bit-equal in fp32 mode; the bf16 FFN adds its split-K partial sums with reduce-stores whose order is not fixed)
    assert rel_err(res2["pred_bboxes"].cpu().numpy(), res["pred_bboxes"].cpu().numpy()) < 1e-3 and len(model._engines) == 1


def test_memotr_module_autograd_route_and_query_updater_forward():
    """Outside eval/no_grad the same module runs the autograd mirrors (same outputs), and postprocess_single_frame =
    QueryUpdater.forward = select_active_tracks (eval branch, models/query_updater.py:243-254) + update_tracks_embedding."""
    g, cfg, x, model, frame, tr = _memotr("small", "fp32")
    res = model(frame=frame, tracks=[tr])               # grad enabled -> module path
    assert not model._engines and res["outputs"].requires_grad
    assert rel_err(res["outputs"].detach().cpu().numpy(), g["outputs"]) < 1e-4
    assert rel_err(res["pred_bboxes"].detach().cpu().numpy(), g["pred_bboxes"]) < 1e-4
    prev = _Tracks(hidden_dim=256, num_classes=1, n=len(tr), device=DEV)
    for k, v in x["tracks"].items():
        setattr(prev, k, v.clone().to(DEV))
    prev.ids = torch.arange(len(tr), device=DEV)
    prev.ids[1] = -1                                     # a dead track is dropped by select_active_tracks
    new = _Tracks(hidden_dim=256, num_classes=1, n=2, device=DEV)
    gen = torch.Generator().manual_seed(7)
    for k in ("ref_pts", "boxes"):
        setattr(new, k, torch.rand(2, 4, generator=gen).to(DEV))
    new.query_embed, new.output_embed = torch.randn(2, 256, generator=gen).to(DEV), torch.randn(2, 256, generator=gen).to(DEV)
    new.logits = torch.tensor([[3.0], [-3.0]], device=DEV)
    new.ids = torch.tensor([10, 11], device=DEV)
    want_in = {k: torch.cat((getattr(prev, k), getattr(new, k) if k not in ("last_output", "long_memory") else
                             (new.output_embed if k == "last_output" else new.query_embed)))[torch.cat((prev.ids, new.ids)) >= 0]
               for k in ("ref_pts", "query_embed", "output_embed", "last_output", "long_memory", "logits", "boxes")}
    with torch.no_grad():
        out = model.postprocess_single_frame([prev], [new], None)
    assert len(out) == 1 and out[0].ids.tolist() == [0, 2, 3, 4, 10, 11][:len(out[0].ids)] or out[0].ids.tolist() == [i for i in [0] + list(range(2, len(tr))) + [10, 11]]
    want = oframe.update_tracks({k: v.cpu() for k, v in model.state_dict().items()}, {k: v.cpu() for k, v in want_in.items()}, cfg)
    for k in ("ref_pts", "query_embed", "long_memory", "last_output"):
        assert rel_err(getattr(out[0], k).cpu().numpy(), want[k].numpy()) < 1e-4, k


def test_memotr_module_input_projections_run_on_our_kernels_in_eval():
    """MeMOTR._pyramid: with the reference's own feature_projs (Conv2d + GroupNorm) the eval / no-grad / batch-1 route computes
    the projections with memotr_b200.input_proj (csrc/input_proj.cu); same pyramid as the nn.Module route."""
    from memotr_b200 import memotr as mm
    torch.backends.cuda.matmul.allow_tf32 = False          # the reference's precision contract (main.py:96-97); cuDNN's
    torch.backends.cudnn.allow_tf32 = False                # default would run the yardstick convolutions in TF32
    cfg = synth.small_cfg()
    rc = dict(oframe.to_reference_config(cfg), DATASET="DanceTrack", ENGINE_MODE="fp32")
    g = torch.Generator().manual_seed(4)
    shapes = [(24, 40), (12, 20), (6, 10)]
    feats = [torch.randn(1, 8, h, w, generator=g).to(DEV) for h, w in shapes]
    masks = [torch.zeros(1, h, w, dtype=torch.bool, device=DEV) for h, w in shapes]
    pos = [torch.randn(1, 256, h, w, generator=g).to(DEV) for h, w in shapes] + [torch.randn(1, 256, 3, 5, generator=g).to(DEV)]
    model = mm.build(rc, _FakeBackbone(feats, masks, pos)).to(DEV)
    assert isinstance(model.feature_projs[0][0], nn.Conv2d) and len(model.feature_projs) == 4
    frame = _NT(torch.zeros(1, 3, 24 * 8, 40 * 8, device=DEV), torch.zeros(1, 24 * 8, 40 * 8, dtype=torch.bool, device=DEV))
    model.train()
    want, _, _ = model._pyramid(frame)                              # nn.Conv2d + nn.GroupNorm
    assert getattr(model, "_input_proj", None) is None
    model.eval()
    with torch.no_grad():
        got, _, _ = model._pyramid(frame)
    assert model._input_proj is not None and len(got) == 4
    for a, b in zip(got, want):
        assert a.shape == b.shape and rel_err(a.cpu().numpy(), b.detach().cpu().numpy()) < 1e-5
