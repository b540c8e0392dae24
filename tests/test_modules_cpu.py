"""CPU checks of the nn.Module mirrors: parameter names and shapes are exactly the reference's state_dict entries
(the table in memotr_b200/synthetic.py is asserted against the instantiated reference modules by oracle/make_golden.py),
so released checkpoints load; build() contracts accept the reference's flat config dict."""
import torch
from torch import nn

from memotr_b200 import modules, synthetic as synth
from oracle import frame as oframe


def _build(cfg):
    rc = oframe.to_reference_config(cfg)
    tr, qu = modules.build_transformer(rc), modules.build_query_updater(rc)
    bbox = nn.ModuleList([modules.MLP(cfg["d_model"], cfg["d_model"], 4, 3) for _ in range(cfg["n_dec_layers"])])
    tr.set_refine_bbox_embed(bbox)                    # memotr.py:91-92
    return tr, qu, bbox


def test_state_dict_keys_and_shapes_match_reference_table():
    cfg = synth.small_cfg()
    tr, qu, bbox = _build(cfg)
    want = synth.hot_path_param_shapes(cfg)
    have = {"transformer." + k: tuple(v.shape) for k, v in tr.state_dict().items()}
    have.update({"query_updater." + k: tuple(v.shape) for k, v in qu.state_dict().items()})
    alias = {k for k in have if k.startswith("transformer.decoder.bbox_embed.")}      # same tensors as bbox_embed.*
    assert len(alias) == 6 * cfg["n_dec_layers"]
    have = {k: v for k, v in have.items() if k not in alias}
    sub = {k: v for k, v in want.items() if k.startswith(("transformer.", "query_updater."))}
    assert have == sub


def test_reference_checkpoint_layout_loads_strictly():
    cfg = synth.small_cfg()
    tr, qu, bbox = _build(cfg)
    sd = synth.hot_path_state_dict(cfg, seed=0)
    tsd = {k[len("transformer."):]: v for k, v in sd.items() if k.startswith("transformer.")}
    tsd.update({"decoder.bbox_embed." + k[len("bbox_embed."):]: v for k, v in sd.items() if k.startswith("bbox_embed.")})
    tr.load_state_dict(tsd, strict=True)
    qu.load_state_dict({k[len("query_updater."):]: v for k, v in sd.items() if k.startswith("query_updater.")}, strict=True)
    assert torch.equal(bbox[1].layers[2].weight, sd["bbox_embed.1.layers.2.weight"])


def test_modules_refuse_cpu_execution_like_the_reference_op():
    cfg = synth.small_cfg()
    tr, _, _ = _build(cfg)
    x = synth.frame_inputs(cfg, synth.SMALL_SHAPES, 3, seed=1)
    q = torch.randn(1, cfg["n_det_queries"] + 3, 256)
    try:
        tr(x["srcs"], x["masks"], x["pos"], q, torch.randn(1, q.shape[1], 4), torch.zeros(1, q.shape[1], dtype=torch.bool))
    except RuntimeError as e:
        assert "Not implemented on the CPU" in str(e)
    else:
        raise AssertionError("expected the CUDA-only operator to reject CPU tensors")
