"""oracle/frame.py -- functional torch restatement of the reference graph ABOVE the MSDA op.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Plain fp32/fp64 PyTorch ops on whatever device the
tensors live on (CPU for the checker and the cpu_baseline); no nn.Module, no autograd state: every function
takes the reference's own ``state_dict`` (same key names and shapes) plus explicit tensors.

Restated reference code (paths relative to /root/reference/):
  msda_core            models/ops/functions/ms_deform_attn_func.py:44-64   (grid_sample formulation)
  msda_module          models/ops/modules/ms_deform_attn.py:88-130
  mha                  torch.nn.functional.multi_head_attention_forward math path (torch/nn/functional.py),
                       reached from models/deformable_decoder.py:245-252 and models/query_updater.py:125
  encoder              models/deformable_encoder.py:29-60, 97-131
  decoder              models/deformable_decoder.py:56-171, 245-319   (USE_DAB=True branch)
  transformer          models/deformable_transformer.py:175-259
  heads                models/memotr.py:147-195
  update_tracks        models/query_updater.py:82-166                  (USE_DAB=True branch)
  pos_to_pos_embed     models/utils.py:78-85 ; inverse_sigmoid utils/utils.py:61-74

Parity status: the reference ships no test for anything in this file ("parity unpinned" by the reference's
own tests); it is pinned instead against the reference modules themselves, imported and run in the
authoring container by oracle/make_golden.py -> tests/golden/frame_small.npz (tests/test_oracle_cpu.py).
"""
import math

import torch
import torch.nn.functional as F

LN_EPS = 1e-5


# ----------------------------------------------------------------------------------------------- numerics model
# The engine's arithmetic modes restated as ROUNDING POINTS on the fp32 graph (test infrastructure: this is how the
# parity tests separate "the kernels compute what they claim" from "the network amplifies operand rounding"):
#   gemm   None      fp32 operands (the reference's precision contract, main.py:96-97: TF32 off)
#          "bf16"    both operands of every nn.Linear rounded to bf16, fp32 accumulation (engine mode "bf16")
#          "bf16x3"  a = a_hi + a_lo (two bf16 terms each), products hi*hi + hi*lo + lo*hi, fp32 accumulation (~2^-16)
#          "bf16x6"  three bf16 terms per operand, the six products down to 2^-24 (engine mode "fp32tc")
#          "tf32x3"  two tf32 terms per operand, three products (~2^-21)
#          "fp16x3"  two fp16 terms per operand (22 mantissa bits), three products (engine mode "fp32tc")
#   value  None | "fp16" | "bf16"   storage type of the projected value maps the gather reads
#   gather None | "bf16"            storage type of the gather's output rows (a GEMM operand)
NUMERICS = {"gemm": None, "value": None, "gather": None, "fp16_weight_shift": 6}


class numerics:
    """with numerics(gemm="bf16", value="fp16", gather="bf16"): ... -- scoped switch of the rounding model."""

    def __init__(self, **kw):
        assert set(kw) <= set(NUMERICS), kw
        self.kw = kw

    def __enter__(self):
        self.saved = dict(NUMERICS)
        NUMERICS.update(self.kw)
        return self

    def __exit__(self, *exc):
        NUMERICS.clear()
        NUMERICS.update(self.saved)
        return False


def _tf32(x):
    """Round fp32 to the nearest tf32 (10 explicit mantissa bits), ties to even."""
    i = x.contiguous().view(torch.int32)
    r = ((i >> 13) & 1) + 0xFFF
    return ((i + r) & ~0x1FFF).view(torch.float32)


def _split(x, n, rnd):
    parts, rest = [], x
    for _ in range(n):
        p = rnd(rest)
        parts.append(p)
        rest = rest - p
    return parts


def _rounded_matmul(x, w, mode):
    """x (…, K) @ w (N, K)^T under the operand-rounding model `mode`; fp32 accumulation."""
    bf = lambda t: t.to(torch.bfloat16).to(torch.float32)              # noqa: E731
    if mode == "bf16":
        return F.linear(bf(x), bf(w))
    hf = lambda t: t.to(torch.float16).to(torch.float32)               # noqa: E731  (subnormals kept, like the MMA)
    n, rnd = {"bf16x3": (2, bf), "bf16x6": (3, bf), "tf32x3": (2, _tf32), "fp16x3": (2, hf)}[mode]
    sc = 1.0
    if mode == "fp16x3":       # the packed weights carry an exact power-of-two scale so that their low halves stay normal
        sc = 2.0 ** NUMERICS.get("fp16_weight_shift", 6)
    xs, ws = _split(x, n, rnd), _split(w * sc, n, rnd)
    out = None
    for i in reversed(range(n)):                 # smallest terms first
        for j in reversed(range(n)):
            if i + j < n:
                t = F.linear(xs[i], ws[j])
                out = t if out is None else out + t
    return out / sc if sc != 1.0 else out


def _store(x, kind):
    if kind is None:
        return x
    return x.to({"fp16": torch.float16, "bf16": torch.bfloat16}[kind]).to(x.dtype)


# ----------------------------------------------------------------------------------------------- small pieces
def linear(sd, key, x):
    w, b = sd[key + ".weight"], sd[key + ".bias"]
    mode = NUMERICS["gemm"]
    if isinstance(mode, dict):                   # per-region model: longest matching key prefix wins ("" = default)
        hit = max((p for p in mode if key.startswith(p)), key=len, default=None)
        mode = mode[hit] if hit is not None else None
    if mode is None or x.dtype != torch.float32:
        return F.linear(x, w, b)
    return _rounded_matmul(x, w, mode) + b


def layer_norm(sd, key, x):
    return F.layer_norm(x, (x.shape[-1],), sd[key + ".weight"], sd[key + ".bias"], LN_EPS)


def mlp(sd, key, x, n_layers):
    """models/mlp.py:13-25 -- ReLU between layers, none after the last."""
    for i in range(n_layers):
        x = linear(sd, f"{key}.layers.{i}", x)
        if i < n_layers - 1:
            x = torch.relu(x)
    return x


def ffn_block(sd, key, x, norm_key=None):
    """models/ffn.py:15-25: LN(x + W2 relu(W1 x)).  norm_key defaults to '<key>.norm'."""
    y = linear(sd, key + ".linear2", torch.relu(linear(sd, key + ".linear1", x)))
    return layer_norm(sd, norm_key or key + ".norm", x + y)


def inverse_sigmoid(x, eps=1e-5):
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def _r32(x, like):
    """A float64 intermediate rounded the way a float32 evaluation rounds it (identity for float64 inputs)."""
    return x.float().double() if like.dtype != torch.float64 else x


def _dim_table(num_pos_feats, temperature, like):
    i = torch.arange(num_pos_feats, dtype=torch.float64, device=like.device)
    e32 = _r32(2 * torch.div(i, 2, rounding_mode="trunc") / num_pos_feats, like)
    return _r32(torch.as_tensor(float(temperature), dtype=torch.float64, device=like.device) ** e32, like)


def pos_to_pos_embed(pos, num_pos_feats=64, temperature=10000, scale=2 * math.pi):
    """models/utils.py:78-85: interleaved sin/cos, coordinate-major.
    Evaluated as the correctly rounded float32 sequence of the reference (every step in float64, rounded to float32 where the
    reference holds a float32): the result does not depend on which vectorised float32 pow / sin / cos the host's torch build
    dispatches to (a GPU-box host type was seen to deviate by 1.5e-4 here, enough to fail 1e-5 kernel tests)."""
    like = pos
    s32 = float(torch.tensor(scale, dtype=torch.float32)) if pos.dtype != torch.float64 else scale
    p = _r32(pos.double() * s32, like)
    e = _r32(p[..., None] / _dim_table(num_pos_feats, temperature, like), like)
    e = torch.stack((e[..., 0::2].sin(), e[..., 1::2].cos()), dim=-1)
    return torch.flatten(e, start_dim=-3).to(like.dtype)


def mha(sd, key, q, k, v, n_heads, key_padding_mask=None):
    """nn.MultiheadAttention(batch_first=True), need_weights path: separate q/k/v in-projections from the
    chunks of in_proj_weight, q scaled by 1/sqrt(d) before QK^T, additive -inf key padding mask, softmax, PV,
    out_proj.  q,k,v: (B, N, C)."""
    B, Nq, C = q.shape
    Nk = k.shape[1]
    d = C // n_heads
    w, b = sd[key + ".in_proj_weight"], sd[key + ".in_proj_bias"]
    tmp = {"q.weight": w[:C], "q.bias": b[:C], "k.weight": w[C:2 * C], "k.bias": b[C:2 * C],
           "v.weight": w[2 * C:], "v.bias": b[2 * C:]}
    qp, kp, vp = linear(tmp, "q", q), linear(tmp, "k", k), linear(tmp, "v", v)
    qp = qp.view(B, Nq, n_heads, d).transpose(1, 2) * math.sqrt(1.0 / d)
    kp = kp.view(B, Nk, n_heads, d).transpose(1, 2)
    vp = vp.view(B, Nk, n_heads, d).transpose(1, 2)
    logits = qp @ kp.transpose(-1, -2)
    if key_padding_mask is not None:
        logits = logits.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    att = torch.softmax(logits, dim=-1)
    out = (att @ vp).transpose(1, 2).reshape(B, Nq, C)
    return linear(sd, key + ".out_proj", out)


# ----------------------------------------------------------------------------------------------- MSDA
def msda_core(value, shapes, loc, attn):
    """ms_deform_attn_core_pytorch (func.py:44-64): per level grid_sample(bilinear, zeros,
    align_corners=False) on 2*loc-1, weighted by attn and summed over levels*points."""
    B, S, H, D = value.shape
    _, Lq, _, L, K, _ = loc.shape
    sizes = [int(h) * int(w) for h, w in shapes]
    grids = 2 * loc - 1
    sampled = []
    for lvl, (h, w) in enumerate(shapes):
        h, w = int(h), int(w)
        start = sum(sizes[:lvl])
        v = value[:, start:start + h * w].flatten(2).transpose(1, 2).reshape(B * H, D, h, w)
        g = grids[:, :, :, lvl].transpose(1, 2).flatten(0, 1)
        sampled.append(F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    a = attn.transpose(1, 2).reshape(B * H, 1, Lq, L * K)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * a).sum(-1).view(B, H * D, Lq)
    return out.transpose(1, 2).contiguous()


def msda_module(sd, key, query, ref, src, shapes, lsi, padding_mask, n_heads, n_levels, n_points, core=None):
    """MSDeformAttn.forward (modules/ms_deform_attn.py:88-130).  `core(value, shapes, lsi, loc, attn)` may
    replace the grid_sample formulation (e.g. with the C oracle or the CUDA op under test)."""
    B, Lq, C = query.shape
    S = src.shape[1]
    value = linear(sd, key + ".value_proj", src)
    if padding_mask is not None:
        value = value.masked_fill(padding_mask[..., None], 0.0)
    value = _store(value, NUMERICS["value"]).view(B, S, n_heads, C // n_heads)
    off = linear(sd, key + ".sampling_offsets", query).view(B, Lq, n_heads, n_levels, n_points, 2)
    aw = linear(sd, key + ".attention_weights", query).view(B, Lq, n_heads, n_levels * n_points)
    aw = torch.softmax(aw, -1).view(B, Lq, n_heads, n_levels, n_points)
    shapes_t = torch.as_tensor([[int(h), int(w)] for h, w in shapes], dtype=torch.long, device=query.device)
    if ref.shape[-1] == 2:
        norm = torch.stack([shapes_t[:, 1], shapes_t[:, 0]], -1)        # (W_l, H_l), ms_deform_attn.py:117
        loc = ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    else:
        loc = ref[:, :, None, :, None, :2] + off / n_points * ref[:, :, None, :, None, 2:] * 0.5
    if core is None:
        out = msda_core(value, shapes, loc, aw)
    else:
        out = core(value, shapes_t, lsi, loc, aw)
    return linear(sd, key + ".output_proj", _store(out, NUMERICS["gather"]))


# ----------------------------------------------------------------------------------------------- encoder
def encoder_reference_points(shapes, valid_ratios, device):
    """DeformableEncoder.get_reference_points (deformable_encoder.py:29-40)."""
    refs = []
    for lvl, (h, w) in enumerate(shapes):
        h, w = int(h), int(w)
        ry, rx = torch.meshgrid(torch.linspace(0.5, h - 0.5, h, dtype=torch.float32, device=device),
                                torch.linspace(0.5, w - 0.5, w, dtype=torch.float32, device=device), indexing="ij")
        ry = ry.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * h)
        rx = rx.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * w)
        refs.append(torch.stack((rx, ry), -1))
    refs = torch.cat(refs, 1)
    return refs[:, :, None] * valid_ratios[:, None]


def encoder_layer(sd, key, src, pos, ref, shapes, lsi, padding_mask, cfg, core=None):
    a = msda_module(sd, key + ".self_attn", src + pos, ref, src, shapes, lsi, padding_mask,
                    cfg["n_heads"], cfg["n_levels"], cfg["n_enc_points"], core)
    src = layer_norm(sd, key + ".norm1", src + a)
    return ffn_block(sd, key, src, norm_key=key + ".norm2")


def encoder(sd, key, src, pos, shapes, lsi, valid_ratios, padding_mask, cfg, core=None):
    ref = encoder_reference_points(shapes, valid_ratios, src.device)
    for i in range(cfg["n_enc_layers"]):
        src = encoder_layer(sd, f"{key}.layers.{i}", src, pos, ref, shapes, lsi, padding_mask, cfg, core)
    return src


# ----------------------------------------------------------------------------------------------- decoder
def decoder_layer(sd, key, tgt, query_pos, ref_in, memory, shapes, lsi, query_mask, mem_mask, cfg, merge, core=None):
    """DeformableDecoderLayer.forward (deformable_decoder.py:275-319)."""
    nd = cfg["n_det_queries"]
    track_tgt = None
    if not merge:
        track_tgt, tgt = tgt[:, nd:], tgt[:, :nd]
        query_pos, ref_in, query_mask = query_pos[:, :nd], ref_in[:, :nd], query_mask[:, :nd]
    qk = tgt + query_pos
    tgt = layer_norm(sd, key + ".norm2", tgt + mha(sd, key + ".self_attn", qk, qk, tgt, cfg["n_heads"], query_mask))
    a = msda_module(sd, key + ".cross_attn", tgt + query_pos, ref_in, memory, shapes, lsi, mem_mask,
                    cfg["n_heads"], cfg["n_levels"], cfg["n_dec_points"], core)
    tgt = layer_norm(sd, key + ".norm1", tgt + a)
    tgt = ffn_block(sd, key, tgt, norm_key=key + ".norm3")
    if not merge:
        tgt = torch.cat((tgt, track_tgt), dim=1)
    return tgt


def decoder_step(sd, lid, out, ref, memory, shapes, lsi, valid_ratios, query_mask, mem_mask, cfg, core=None,
                 key="transformer.decoder", bbox_key="bbox_embed"):
    """One iteration of DeformableDecoder.forward's layer loop (deformable_decoder.py:80-159, USE_DAB + box refinement):
    (layer input `out` (B,Nq,C), sigmoid-space boxes `ref` (B,Nq,4)) -> (layer output, refined boxes).  Also the unit the
    teacher-forced parity tests drive layer by layer."""
    C = out.shape[-1]
    nd, merge_layer = cfg["n_det_queries"], cfg["merge_det_track_layer"]
    ref_in = ref[:, :, None] * torch.cat([valid_ratios, valid_ratios], -1)[:, None]
    anchor = pos_to_pos_embed(ref_in[:, :, 0, :], num_pos_feats=C // 2)
    raw_pos = mlp(sd, key + ".ref_point_head", anchor, 2)
    scale = mlp(sd, key + ".query_scale", out, 2) if lid != 0 else 1
    query_pos = scale * raw_pos
    out = decoder_layer(sd, f"{key}.layers.{lid}", out, query_pos, ref_in, memory, shapes, lsi, query_mask,
                        mem_mask, cfg, merge=(lid >= merge_layer), core=core)
    new_ref = (mlp(sd, f"{bbox_key}.{lid}", out, 3) + inverse_sigmoid(ref)).sigmoid()
    if lid < merge_layer:                  # refined points are detached (deformable_decoder.py:150-159)
        ref = torch.cat((new_ref[:, :nd].detach(), ref[:, nd:]), dim=1)
    else:
        ref = new_ref.detach()
    return out, ref


def decoder(sd, key, bbox_key, tgt, ref, memory, shapes, lsi, valid_ratios, query_mask, mem_mask, cfg, core=None):
    """DeformableDecoder.forward, USE_DAB branch with iterative box refinement
    (deformable_decoder.py:56-171).  Returns (outputs[n,B,Nq,C], refs[n,B,Nq,4], queries[n,B,Nq,C])."""
    outs, refs, queries = [], [], []
    out = tgt
    for lid in range(cfg["n_dec_layers"]):
        queries.append(out)
        out, ref = decoder_step(sd, lid, out, ref, memory, shapes, lsi, valid_ratios, query_mask, mem_mask, cfg, core,
                                key, bbox_key)
        outs.append(out)
        refs.append(ref)
    return torch.stack(outs), torch.stack(refs), torch.stack(queries)


# ----------------------------------------------------------------------------------------------- transformer
def valid_ratio(mask):
    """DeformableTransformer.get_valid_ratio (deformable_transformer.py:175-190) -> (B, 2) as (w, h)."""
    _, H, W = mask.shape
    vh = torch.sum(~mask[:, :, 0], 1).float() / H
    vw = torch.sum(~mask[:, 0, :], 1).float() / W
    return torch.stack([vw, vh], -1)


def flatten_levels(sd, key, srcs, masks, pos_embeds):
    """deformable_transformer.py:196-220."""
    src_f, mask_f, pos_f, shapes = [], [], [], []
    for lvl, (s, m, p) in enumerate(zip(srcs, masks, pos_embeds)):
        shapes.append((s.shape[2], s.shape[3]))
        src_f.append(s.flatten(2).transpose(1, 2))
        mask_f.append(m.flatten(1))
        pos_f.append(p.flatten(2).transpose(1, 2) + sd[key + ".level_embed"][lvl].view(1, 1, -1))
    sizes = [h * w for h, w in shapes]
    lsi = torch.as_tensor([sum(sizes[:i]) for i in range(len(sizes))], dtype=torch.long, device=srcs[0].device)
    vr = torch.stack([valid_ratio(m) for m in masks], 1)
    return torch.cat(src_f, 1), torch.cat(mask_f, 1), torch.cat(pos_f, 1), shapes, lsi, vr


def transformer(sd, srcs, masks, pos_embeds, query_embed, ref_pts, query_mask, cfg, core=None, key="transformer",
                bbox_key="bbox_embed"):
    """DeformableTransformer.forward (deformable_transformer.py:192-259), DAB branch.
    -> outputs (n,B,Nq,C), init_ref (B,Nq,4), inter_refs (n,B,Nq,4), inter_queries (n,B,Nq,C), memory."""
    src, mask, pos, shapes, lsi, vr = flatten_levels(sd, key, srcs, masks, pos_embeds)
    memory = encoder(sd, key + ".encoder", src, pos, shapes, lsi, vr, mask, cfg, core)
    init_ref = ref_pts.sigmoid()
    outs, refs, queries = decoder(sd, key + ".decoder", bbox_key, query_embed, init_ref, memory, shapes, lsi, vr,
                                  query_mask, mask, cfg, core)
    return outs, init_ref, refs, queries, memory


def heads(sd, outs, init_ref, refs):
    """models/memotr.py:147-162: per-level class / box heads with inverse-sigmoid reference."""
    logits, boxes = [], []
    for lvl in range(outs.shape[0]):
        r = inverse_sigmoid(init_ref if lvl == 0 else refs[lvl - 1])
        logits.append(linear(sd, f"class_embed.{lvl}", outs[lvl]))
        boxes.append((mlp(sd, f"bbox_embed.{lvl}", outs[lvl], 3) + r).sigmoid())
    return torch.stack(logits), torch.stack(boxes)


def frame_forward(sd, srcs, masks, pos_embeds, track_ref_pts, track_query_embed, cfg, core=None):
    """MeMOTR.forward minus the backbone / input projections (models/memotr.py:128-195), batch 1.
    track_ref_pts (Nt,4) and track_query_embed (Nt,C) are TrackInstances.ref_pts / .query_embed."""
    ref_pts = torch.cat((sd["det_anchor"], track_ref_pts), 0)[None]
    query_embed = torch.cat((sd["det_query_embed"], track_query_embed), 0)[None]
    query_mask = torch.zeros((1, ref_pts.shape[1]), dtype=torch.bool, device=ref_pts.device)
    outs, init_ref, refs, queries, memory = transformer(sd, srcs, masks, pos_embeds, query_embed, ref_pts,
                                                        query_mask, cfg, core)
    logits, boxes = heads(sd, outs, init_ref, refs)
    return {
        "pred_logits": logits[-1], "pred_bboxes": boxes[-1],
        "last_ref_pts": inverse_sigmoid(refs[-2]), "init_ref_pts": inverse_sigmoid(init_ref),
        "query_mask": query_mask, "outputs": outs[-1],
        "aux_logits": logits[:-1], "aux_bboxes": boxes[:-1], "aux_queries": queries[1:],
        "memory": memory,
    }


# ----------------------------------------------------------------------------------------------- query updater
def update_tracks(sd, t, cfg, key="query_updater"):
    """QueryUpdater.update_tracks_embedding for one batch item (query_updater.py:82-166, DAB branch).
    `t` is a dict with ref_pts, query_embed, output_embed, last_output, long_memory, logits, boxes.
    Returns a new dict with the five updated fields (+ is_pos)."""
    C = t["output_embed"].shape[-1]
    lam, thr = cfg["long_memory_lambda"], cfg["update_thresh"]
    scores = t["logits"].sigmoid().max(dim=1).values
    is_pos = scores > thr
    ref_pts = t["ref_pts"].clone()
    ref_pts[is_pos] = inverse_sigmoid(t["boxes"][is_pos])
    query_pos = pos_to_pos_embed(ref_pts.sigmoid(), num_pos_feats=C // 2)
    out_e, last_e, long_m = t["output_embed"], t["last_output"], t["long_memory"]
    conf = torch.sigmoid(mlp(sd, key + ".confidence_weight_net.0", out_e, 2))
    short = mlp(sd, key + ".short_memory_fusion", torch.cat((conf * out_e, last_e), -1), 2)
    query_pos = mlp(sd, key + ".query_pos_head", query_pos, 2)
    q, k = short + query_pos, long_m + query_pos
    tgt = out_e + mha(sd, key + ".memory_attn", q[None], k[None], out_e[None], 8)[0]
    tgt = ffn_block(sd, key + ".memory_ffn", layer_norm(sd, key + ".memory_norm", tgt))
    feat = ffn_block(sd, key + ".query_feat_ffn", layer_norm(sd, key + ".query_feat_norm", long_m + tgt))
    m = is_pos[:, None]
    new_long = (1 - lam) * long_m + lam * out_e
    query_embed = t["query_embed"].clone()
    query_embed[is_pos] = feat[is_pos]
    return {
        "ref_pts": ref_pts, "query_embed": query_embed,
        "long_memory": long_m * ~m + new_long * m,
        "last_output": last_e * ~m + out_e * m,
        "is_pos": is_pos,
    }


# ----------------------------------------------------------------------------------------------- configs
def dancetrack_cfg():
    from memotr_b200.synthetic import dancetrack_cfg as _cfg
    return _cfg()


def to_reference_config(cfg):
    """The same numbers as the flat upper-case dict the reference build() functions read."""
    return {
        "HIDDEN_DIM": cfg["d_model"], "FFN_DIM": cfg["d_ffn"], "NUM_FEATURE_LEVELS": cfg["n_levels"],
        "NUM_HEADS": cfg["n_heads"], "NUM_ENC_POINTS": cfg["n_enc_points"], "NUM_DEC_POINTS": cfg["n_dec_points"],
        "NUM_ENC_LAYERS": cfg["n_enc_layers"], "NUM_DEC_LAYERS": cfg["n_dec_layers"],
        "MERGE_DET_TRACK_LAYER": cfg["merge_det_track_layer"], "DROPOUT": 0.0, "ACTIVATION": "ReLU",
        "RETURN_INTER_DEC": True, "NUM_DET_QUERIES": cfg["n_det_queries"], "EXTRA_TRACK_ATTN": False,
        "USE_CHECKPOINT": False, "CHECKPOINT_LEVEL": 2, "USE_DAB": True, "VISUALIZE": False,
        "UPDATE_THRESH": cfg["update_thresh"], "LONG_MEMORY_LAMBDA": cfg["long_memory_lambda"],
        "TP_DROP_RATE": 0.0, "FP_INSERT_RATE": 0.0,
    }


def position_embedding_sine(mask, num_pos_feats=128, temperature=20, scale=2 * math.pi):
    """PositionEmbeddingSine.forward with normalize=True (models/position_embedding.py:23-43; built with
    num_pos_feats = hidden_dim / 2, temperature 20, scale 2 pi at :46-49).  mask (B, H, W) bool -> (B, 2*npf, H, W)."""
    not_mask = ~mask
    y = not_mask.cumsum(dim=1, dtype=torch.float32)
    x = not_mask.cumsum(dim=2, dtype=torch.float32)
    eps = 1e-6
    y = (y - 0.5) / (y[:, -1:, :] + eps) * scale
    x = (x - 0.5) / (x[:, :, -1:] + eps) * scale
    dim_i = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_i = temperature ** (2 * torch.div(dim_i, 2, rounding_mode="trunc") / num_pos_feats)
    pos_x = x[:, :, :, None] / dim_i
    pos_y = y[:, :, :, None] / dim_i
    pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
    pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)
