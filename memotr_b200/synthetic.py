"""memotr_b200/synthetic.py -- deterministic synthetic weights, inputs and configuration tables for the hot path.

No network, datasets or checkpoints exist in the build/benchmark environment, so bench.py, the tests and the golden
generator (oracle/make_golden.py) all draw their tensors from here.  Pure data generation: no model arithmetic.

Everything is drawn from a seeded *CPU* torch.Generator, so the authoring container (where the reference is
importable and golden outputs are produced, oracle/make_golden.py) and the GPU box (where only this repo
exists) regenerate bit-identical tensors from (cfg, seed) and the golden files only need to hold outputs.

Input recipes follow SURVEY.md section 8d:
  op level    models/ops/test.py:33-36  (value = rand*0.01, loc = rand [or rand*1.2-0.1 for the border
              variant], attention weights = rand+1e-5 normalised over (L,K))
  frame level srcs/pos = randn(1,C,H_l,W_l), masks all-False or right/bottom padded, det_anchor/det_query_embed
              = randn, Nt synthetic tracks with randn embeddings and logits kept >= 0.1 away from the 0.5 score
              threshold.
"""
import math

import torch

DANCETRACK_SHAPES = ((100, 168), (50, 84), (25, 42), (13, 21))          # 1333x800 -> padded 800x1344, S=22323
BDD_SHAPES = ((92, 160), (46, 80), (23, 40), (12, 20))                  # 1280x720 -> padded 736x1280, S=19560
BDD_SHAPES_L5 = BDD_SHAPES + ((6, 10),)


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def msda_inputs(shapes, B=1, H=8, D=32, K=4, Lq=100, seed=3, border=False, dtype=torch.float32):
    """-> value (B,S,H,D), shapes (L,2) i64, level_start_index (L,) i64, loc (B,Lq,H,L,K,2), attn (B,Lq,H,L,K)."""
    g = _gen(seed)
    shapes_t = torch.as_tensor(shapes, dtype=torch.long)
    L = shapes_t.shape[0]
    S = int((shapes_t[:, 0] * shapes_t[:, 1]).sum())
    lsi = torch.cat((shapes_t.new_zeros((1,)), shapes_t.prod(1).cumsum(0)[:-1]))
    value = torch.rand(B, S, H, D, generator=g) * 0.01
    loc = torch.rand(B, Lq, H, L, K, 2, generator=g)
    if border:
        loc = loc * 1.2 - 0.1
    attn = torch.rand(B, Lq, H, L, K, generator=g) + 1e-5
    attn = attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    return value.to(dtype), shapes_t, lsi, loc.to(dtype), attn.to(dtype)


def encoder_msda_inputs(shapes, H=8, K=4, seed=3, noise_px=0.5, valid=(1.0, 1.0), dtype=torch.float32):
    """Inputs of an ENCODER-shaped MSDA call (the queries are the S pixels of the pyramid, deformable_encoder.py:124):
    reference points = pixel centres over the valid extent (deformable_encoder.py:29-40), sampling offsets = the ring
    pattern MSDeformAttn initialises its bias to (head direction x point index, ms_deform_attn.py:72-81) plus Gaussian
    noise of `noise_px` pixels -- the distribution a (lightly) trained encoder produces.  `valid` = (w, h) valid ratio of
    every level.  -> value (S, H*32), valid_ratios (L, 2), loc (S, H, L, K, 2), attn (S, H, L, K), shift (H, L, 2) [mean
    offset per head and level in pixels: the staging hint of the windowed gather]."""
    g = _gen(seed)
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    vr = torch.tensor([[float(valid[0]), float(valid[1])]] * L)
    refs = []
    for l, (h, w) in enumerate(shapes):
        ys = (torch.arange(h, dtype=torch.float32) + 0.5) / (vr[l, 1] * h)
        xs = (torch.arange(w, dtype=torch.float32) + 0.5) / (vr[l, 0] * w)
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        refs.append(torch.stack((xx.reshape(-1), yy.reshape(-1)), -1))
    ref = torch.cat(refs, 0)                                                    # (S, 2), normalised by the valid extent
    th = torch.arange(H, dtype=torch.float32) * (2.0 * math.pi / H)
    ring = torch.stack([th.cos(), th.sin()], -1)
    ring = (ring / ring.abs().max(-1, keepdim=True)[0]).view(1, H, 1, 1, 2)
    off = ring * torch.arange(1, K + 1, dtype=torch.float32).view(1, 1, 1, K, 1)      # (1, H, 1, K, 2) pixels
    off = off + noise_px * torch.randn(S, H, L, K, 2, generator=g)
    wh = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32).view(1, 1, L, 1, 2)
    loc = ref.view(S, 1, 1, 1, 2) * vr.view(1, 1, L, 1, 2) + off / wh
    attn = torch.rand(S, H, L, K, generator=g) + 1e-5
    attn = attn / attn.sum((-1, -2), keepdim=True)
    value = torch.randn(S, H * 32, generator=g)
    shift = (ring * (K + 1) / 2.0).view(H, 1, 2).repeat(1, L, 1)
    return value.to(dtype), vr, loc.to(dtype).contiguous(), attn.to(dtype).contiguous(), shift


def hot_path_param_shapes(cfg):
    """Key -> shape for every hot-path parameter, in the reference's state_dict naming.
    oracle/make_golden.py asserts this table against the instantiated reference modules."""
    C, Fd, H, L = cfg["d_model"], cfg["d_ffn"], cfg["n_heads"], cfg["n_levels"]
    ncls = cfg["num_classes"]
    t = {}

    def lin(key, o, i):
        t[key + ".weight"] = (o, i)
        t[key + ".bias"] = (o,)

    def ln(key):
        t[key + ".weight"] = (C,)
        t[key + ".bias"] = (C,)

    def msda(key, K):
        lin(key + ".sampling_offsets", H * L * K * 2, C)
        lin(key + ".attention_weights", H * L * K, C)
        lin(key + ".value_proj", C, C)
        lin(key + ".output_proj", C, C)

    t["transformer.level_embed"] = (L, C)
    for i in range(cfg["n_enc_layers"]):
        k = f"transformer.encoder.layers.{i}"
        msda(k + ".self_attn", cfg["n_enc_points"])
        ln(k + ".norm1"); lin(k + ".linear1", Fd, C); lin(k + ".linear2", C, Fd); ln(k + ".norm2")
    for i in range(cfg["n_dec_layers"]):
        k = f"transformer.decoder.layers.{i}"
        t[k + ".self_attn.in_proj_weight"] = (3 * C, C)
        t[k + ".self_attn.in_proj_bias"] = (3 * C,)
        lin(k + ".self_attn.out_proj", C, C)
        ln(k + ".norm2")
        msda(k + ".cross_attn", cfg["n_dec_points"])
        ln(k + ".norm1"); lin(k + ".linear1", Fd, C); lin(k + ".linear2", C, Fd); ln(k + ".norm3")
    lin("transformer.decoder.query_scale.layers.0", C, C)
    lin("transformer.decoder.query_scale.layers.1", C, C)
    lin("transformer.decoder.ref_point_head.layers.0", C, 2 * C)
    lin("transformer.decoder.ref_point_head.layers.1", C, C)
    q = "query_updater"
    lin(q + ".confidence_weight_net.0.layers.0", C, C)
    lin(q + ".confidence_weight_net.0.layers.1", C, C)
    lin(q + ".short_memory_fusion.layers.0", 2 * C, 2 * C)
    lin(q + ".short_memory_fusion.layers.1", C, 2 * C)
    t[q + ".memory_attn.in_proj_weight"] = (3 * C, C)
    t[q + ".memory_attn.in_proj_bias"] = (3 * C,)
    lin(q + ".memory_attn.out_proj", C, C)
    ln(q + ".memory_norm")
    lin(q + ".memory_ffn.linear1", Fd, C); lin(q + ".memory_ffn.linear2", C, Fd); ln(q + ".memory_ffn.norm")
    ln(q + ".query_feat_norm")
    lin(q + ".query_feat_ffn.linear1", Fd, C); lin(q + ".query_feat_ffn.linear2", C, Fd); ln(q + ".query_feat_ffn.norm")
    lin(q + ".query_pos_head.layers.0", C, 2 * C)
    lin(q + ".query_pos_head.layers.1", C, C)
    for i in range(cfg["n_dec_layers"]):
        lin(f"class_embed.{i}", ncls, C)
        lin(f"bbox_embed.{i}.layers.0", C, C)
        lin(f"bbox_embed.{i}.layers.1", C, C)
        lin(f"bbox_embed.{i}.layers.2", 4, C)
    t["det_anchor"] = (cfg["n_det_queries"], 4)
    t["det_query_embed"] = (cfg["n_det_queries"], C)
    return t


def hot_path_state_dict(cfg, seed=0):
    """Deterministic fp32 weights with magnitudes that keep the graph well-conditioned: fan-in scaled normal
    weights, small biases, LayerNorm gains near 1, sampling-offset biases on the reference's ring pattern
    (ms_deform_attn.py:72-81) plus noise so the samples spread a few pixels around each reference point."""
    g = _gen(seed)
    sd = {}
    for key, shape in hot_path_param_shapes(cfg).items():
        parts = key.split(".")
        leaf = parts[-1]
        is_norm = len(parts) >= 2 and "norm" in parts[-2]
        if is_norm and leaf == "weight":
            v = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif is_norm and leaf == "bias":
            v = 0.1 * torch.randn(shape, generator=g)
        elif key.endswith("sampling_offsets.bias"):
            H, L = cfg["n_heads"], cfg["n_levels"]
            K = shape[0] // (H * L * 2)
            th = torch.arange(H, dtype=torch.float32) * (2.0 * math.pi / H)
            ring = torch.stack([th.cos(), th.sin()], -1)
            ring = (ring / ring.abs().max(-1, keepdim=True)[0]).view(H, 1, 1, 2).repeat(1, L, K, 1)
            ring = ring * torch.arange(1, K + 1, dtype=torch.float32).view(1, 1, K, 1)
            v = ring.reshape(-1) + 0.5 * torch.randn(shape, generator=g)
        elif key.endswith("sampling_offsets.weight"):
            v = torch.randn(shape, generator=g) * (0.5 / math.sqrt(shape[1]))
        elif key in ("det_anchor", "det_query_embed", "transformer.level_embed"):
            v = torch.randn(shape, generator=g)
        elif leaf in ("weight", "in_proj_weight"):
            v = torch.randn(shape, generator=g) / math.sqrt(shape[1])
        else:
            v = 0.05 * torch.randn(shape, generator=g)
        sd[key] = v.float()
    return sd


def reference_init_state_dict(cfg, seed=0, zero_init_scale=0.1):
    """Weights drawn from the reference's OWN initialisation distributions -- what SURVEY.md 8d names for the
    single-frame configuration ("default init of the reference modules"), regenerated from a seed so that the GPU box
    needs no checkpoint:
      * every matrix of the transformer and the query updater: xavier_uniform (deformable_transformer.py:111-114,
        query_updater.py:67-70); nn.Linear biases U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (torch default);
        MultiheadAttention in_proj_bias / out_proj.bias 0; LayerNorm 1 / 0
      * MSDeformAttn: sampling_offsets.bias = the ring pattern, value_proj / output_proj xavier with zero bias
        (ms_deform_attn.py:72-86); level_embed, det_anchor, det_query_embed ~ N(0, 1) (memotr.py:59-60)
      * class_embed: torch default weight, bias -log(99) (memotr.py:79-81); bbox_embed: torch default, last layer's bias
        (0, 0, -2, -2) on layer 0 (memotr.py:82-91)
    The four matrices the reference initialises to exactly ZERO (sampling_offsets.weight, attention_weights.weight/bias,
    the last bbox_embed layer) get `zero_init_scale` x their xavier range instead, so that their GEMMs carry signal in
    the parity tests (a zero matrix cannot expose a wrong kernel); zero_init_scale=0 reproduces the reference init."""
    g = _gen(seed)
    C, H, L = cfg["d_model"], cfg["n_heads"], cfg["n_levels"]
    sd = {}

    def uniform(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    def xavier(shape, gain=1.0):
        return uniform(shape, gain * math.sqrt(6.0 / (shape[0] + shape[1])))

    shapes = hot_path_param_shapes(cfg)
    for key, shape in shapes.items():
        parts = key.split(".")
        leaf = parts[-1]
        is_norm = len(parts) >= 2 and "norm" in parts[-2]
        in_core = key.startswith("transformer.") or key.startswith("query_updater.")
        if is_norm:
            v = torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
        elif key.endswith("sampling_offsets.bias"):
            K = shape[0] // (H * L * 2)
            th = torch.arange(H, dtype=torch.float32) * (2.0 * math.pi / H)
            ring = torch.stack([th.cos(), th.sin()], -1)
            ring = (ring / ring.abs().max(-1, keepdim=True)[0]).view(H, 1, 1, 2).repeat(1, L, K, 1)
            v = (ring * torch.arange(1, K + 1, dtype=torch.float32).view(1, 1, K, 1)).reshape(-1)
        elif key.endswith("sampling_offsets.weight") or key.endswith("attention_weights.weight"):
            v = xavier(shape, zero_init_scale)
        elif key.endswith("attention_weights.bias"):
            v = uniform(shape, zero_init_scale / math.sqrt(C))
        elif key.endswith("value_proj.bias") or key.endswith("output_proj.bias") or leaf == "in_proj_bias" \
                or key.endswith("out_proj.bias"):
            v = torch.zeros(shape)
        elif key in ("det_anchor", "det_query_embed", "transformer.level_embed"):
            v = torch.randn(shape, generator=g)
        elif key.startswith("class_embed") and leaf == "bias":
            v = torch.full(shape, -math.log(99.0))
        elif key.startswith("bbox_embed") and parts[-2] == "2":          # last layer of the box MLP
            if leaf == "weight":
                v = xavier(shape, zero_init_scale)
            else:
                v = torch.tensor([0.0, 0.0, -2.0, -2.0]) if parts[1] == "0" else torch.zeros(shape)
        elif leaf in ("weight", "in_proj_weight"):
            v = xavier(shape) if in_core else uniform(shape, 1.0 / math.sqrt(shape[1]))
        else:                                                             # nn.Linear bias, torch default
            fan_in = shapes[key[:-len("bias")] + "weight"][1]
            v = uniform(shape, 1.0 / math.sqrt(fan_in))
        sd[key] = v.float()
    return sd


def frame_inputs(cfg, shapes=DANCETRACK_SHAPES, n_tracks=100, seed=1, padded=False):
    """One synthetic frame for batch size 1.  -> dict(srcs, masks, pos: lists per level;  tracks: dict)."""
    g = _gen(seed)
    C = cfg["d_model"]
    srcs = [torch.randn(1, C, h, w, generator=g) for h, w in shapes]
    pos = [torch.randn(1, C, h, w, generator=g) for h, w in shapes]
    masks = []
    for h, w in shapes:
        m = torch.zeros(1, h, w, dtype=torch.bool)
        if padded:                                   # valid region = 1333/1344 wide and 7/8 high
            m[:, :, int(math.ceil(w * 1333 / 1344)):] = True
            m[:, int(math.ceil(h * 0.875)):, :] = True
        masks.append(m)
    nt = n_tracks
    logits = torch.randn(nt, cfg["num_classes"], generator=g)
    logits = logits + torch.sign(logits) * 0.5       # |logit| >= 0.5 => score at least 0.12 away from 0.5
    tracks = {
        "ref_pts": torch.randn(nt, 4, generator=g),
        "query_embed": torch.randn(nt, C, generator=g),
        "output_embed": torch.randn(nt, C, generator=g),
        "last_output": torch.randn(nt, C, generator=g),
        "long_memory": torch.randn(nt, C, generator=g),
        "logits": logits,
        "boxes": torch.rand(nt, 4, generator=g) * 0.8 + 0.1,
    }
    return {"srcs": srcs, "masks": masks, "pos": pos, "tracks": tracks}


def small_cfg():
    """A reduced configuration (same head dim 32 / d_model 256 as DanceTrack, fewer layers, FFN 256,
    12 detect queries) used for the committed golden file."""
    return dict(d_model=256, d_ffn=256, n_levels=4, n_heads=8, n_enc_points=4, n_dec_points=4, n_enc_layers=2,
                n_dec_layers=3, merge_det_track_layer=1, n_det_queries=12, update_thresh=0.5,
                long_memory_lambda=0.01, num_classes=1)


SMALL_SHAPES = ((12, 20), (6, 10), (3, 5), (2, 3))


def dancetrack_cfg():
    """Hot-path hyper-parameters of /root/reference/configs/train_dancetrack.yaml:60-75,89-90."""
    return dict(d_model=256, d_ffn=2048, n_levels=4, n_heads=8, n_enc_points=4, n_dec_points=4, n_enc_layers=6,
                n_dec_layers=6, merge_det_track_layer=1, n_det_queries=300, update_thresh=0.5,
                long_memory_lambda=0.01, num_classes=1)
