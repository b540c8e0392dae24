"""memotr_b200/modules.py -- nn.Module mirrors of the reference transformer surfaces (autograd-capable path).

Same class names, constructor arguments, forward signatures, parameter names and shapes as
  models/deformable_encoder.py   (DeformableEncoderLayer :63-131, DeformableEncoder :22-60)
  models/deformable_decoder.py   (DeformableDecoderLayer :174-319, DeformableDecoder :22-171)
  models/deformable_transformer.py (DeformableTransformer :25-274, build :277-298)
  models/query_updater.py        (QueryUpdater :15-255, build :258-270)
  models/mlp.py, models/ffn.py
so a reference checkpoint loads with `load_state_dict` and `models/memotr.py` can import these instead
(INTEGRATION.md).  The sampling core of every MSDeformAttn goes through MSDeformAttnFunction -> C ABI -> our CUDA
forward/backward kernels; the dense layers here use torch autograd (this is the TRAINING / any-shape path -- the
inference engine in engine.py runs the same graph through our own GEMM/attention kernels).

Restrictions (asserted): the DAB configuration of the released MeMOTR models (USE_DAB True, iterative box refinement,
no two-stage, no extra track attention, VISUALIZE off).  Activation checkpointing flags are accepted and honoured with
torch.utils.checkpoint at layer granularity.
"""
import copy
import math

import torch
import torch.nn.functional as F
from torch import nn
from torch.utils.checkpoint import checkpoint as _ckpt

from .ms_deform_attn import MSDeformAttn


def _clones(m, n):
    return nn.ModuleList([copy.deepcopy(m) for _ in range(n)])


def inverse_sigmoid(x, eps=1e-5):
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def pos_to_pos_embed(pos, num_pos_feats=64, temperature=10000, scale=2 * math.pi):
    idx = torch.arange(num_pos_feats, dtype=torch.float32, device=pos.device)
    div = temperature ** (2 * torch.div(idx, 2, rounding_mode="trunc") / num_pos_feats)
    e = (pos * scale)[..., None] / div
    return torch.stack((e[..., 0::2].sin(), e[..., 1::2].cos()), dim=-1).flatten(-3)


class MLP(nn.Module):
    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        dims = [input_dim] + [hidden_dim] * (num_layers - 1) + [output_dim]
        self.layers = nn.ModuleList(nn.Linear(i, o) for i, o in zip(dims[:-1], dims[1:]))

    def forward(self, x):
        for i, layer in enumerate(self.layers):
            x = layer(x)
            if i + 1 < self.num_layers:
                x = F.relu(x)
        return x


class FFN(nn.Module):
    def __init__(self, d_model, d_ffn, dropout: float):
        super().__init__()
        self.linear1, self.linear2 = nn.Linear(d_model, d_ffn), nn.Linear(d_ffn, d_model)
        self.activation = nn.ReLU(inplace=True)
        self.dropout1, self.dropout2 = nn.Dropout(dropout), nn.Dropout(dropout)
        self.norm = nn.LayerNorm(d_model)

    def forward(self, tgt):
        return self.norm(tgt + self.dropout2(self.linear2(self.dropout1(self.activation(self.linear1(tgt))))))


def _act(name):
    if name == "ReLU":
        return nn.ReLU(True)
    if name == "GELU":
        return nn.GELU()
    raise ValueError(f"Do not support activation layer: {name}")


class DeformableEncoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="ReLU", n_levels=4, n_heads=8, n_points=4,
                 sigmoid_attn=False):
        super().__init__()
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points, sigmoid_attn)
        self.dropout1, self.norm1 = nn.Dropout(dropout), nn.LayerNorm(d_model)
        self.linear1, self.activation = nn.Linear(d_model, d_ffn), _act(activation)
        self.dropout2, self.linear2 = nn.Dropout(dropout), nn.Linear(d_ffn, d_model)
        self.dropout3, self.norm2 = nn.Dropout(dropout), nn.LayerNorm(d_model)

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, padding_mask=None):
        q = src if pos is None else src + pos
        src = self.norm1(src + self.dropout1(self.self_attn(q, reference_points, src, spatial_shapes, level_start_index,
                                                            padding_mask)))
        return self.norm2(src + self.dropout3(self.linear2(self.dropout2(self.activation(self.linear1(src))))))


class DeformableEncoder(nn.Module):
    def __init__(self, encoder_layer, num_layers, use_checkpoint: bool):
        super().__init__()
        self.layers, self.num_layers, self.use_checkpoint = _clones(encoder_layer, num_layers), num_layers, use_checkpoint

    @staticmethod
    def get_reference_points(spatial_shapes, valid_ratios, device):
        """Pixel centres of every level in valid-region-normalised coordinates, then re-scaled per target level."""
        per_level = []
        for lvl, (h, w) in enumerate(spatial_shapes):
            h, w = int(h), int(w)
            ys = torch.linspace(0.5, h - 0.5, h, dtype=torch.float32, device=device)
            xs = torch.linspace(0.5, w - 0.5, w, dtype=torch.float32, device=device)
            gy, gx = torch.meshgrid(ys, xs, indexing="ij")
            gy = gy.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * h)
            gx = gx.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * w)
            per_level.append(torch.stack((gx, gy), -1))
        return torch.cat(per_level, 1)[:, :, None] * valid_ratios[:, None]

    def forward(self, src, spatial_shapes, level_start_index, valid_ratios, pos=None, padding_mask=None):
        ref = self.get_reference_points(spatial_shapes, valid_ratios, device=src.device)
        for layer in self.layers:
            if self.use_checkpoint and src.requires_grad:
                src = _ckpt(layer, src, pos, ref, spatial_shapes, level_start_index, padding_mask, use_reentrant=False)
            else:
                src = layer(src, pos, ref, spatial_shapes, level_start_index, padding_mask)
        return src


class DeformableDecoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="ReLU", n_levels=4, n_heads=8, n_points=4,
                 sigmoid_attn=False, extra_track_attn=False, n_det_queries=300, visualize: bool = False):
        super().__init__()
        assert not extra_track_attn and not visualize, "EXTRA_TRACK_ATTN / VISUALIZE are not part of the hot path"
        self.n_det_queries, self.n_heads = n_det_queries, n_heads
        self.self_attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout, batch_first=True)
        self.dropout2, self.norm2 = nn.Dropout(dropout), nn.LayerNorm(d_model)
        self.cross_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points, sigmoid_attn)
        self.dropout1, self.norm1 = nn.Dropout(dropout), nn.LayerNorm(d_model)
        self.linear1, self.activation = nn.Linear(d_model, d_ffn), _act(activation)
        self.dropout3, self.linear2 = nn.Dropout(dropout), nn.Linear(d_ffn, d_model)
        self.dropout4, self.norm3 = nn.Dropout(dropout), nn.LayerNorm(d_model)

    def forward(self, tgt, query_pos, reference_points, src, src_spatial_shapes, level_start_index, query_mask,
                src_padding_mask=None, merge_det_track=False):
        bypass = None
        if not merge_det_track:                      # detect queries only; track queries skip the layer
            n = self.n_det_queries
            bypass, tgt = tgt[:, n:], tgt[:, :n]
            query_pos, reference_points, query_mask = query_pos[:, :n], reference_points[:, :n], query_mask[:, :n]
        qk = tgt + query_pos
        tgt = self.norm2(tgt + self.dropout2(self.self_attn(qk, qk, tgt, key_padding_mask=query_mask)[0]))
        tgt = self.norm1(tgt + self.dropout1(self.cross_attn(tgt + query_pos, reference_points, src, src_spatial_shapes,
                                                             level_start_index, src_padding_mask)))
        tgt = self.norm3(tgt + self.dropout4(self.linear2(self.dropout3(self.activation(self.linear1(tgt))))))
        return tgt if bypass is None else torch.cat((tgt, bypass), dim=1)


class DeformableDecoder(nn.Module):
    def __init__(self, decoder_layer, num_layers, return_intermediate=False, merge_det_track_layer: int = 0,
                 n_det_queries: int = 300, d_model: int = 256, use_checkpoint: bool = False, use_dab: bool = False,
                 visualize: bool = False):
        super().__init__()
        assert use_dab and return_intermediate and not visualize, "only the DAB / RETURN_INTER_DEC configuration is built"
        self.layers, self.num_layers = _clones(decoder_layer, num_layers), num_layers
        self.return_intermediate, self.merge_det_track_layer = return_intermediate, merge_det_track_layer
        self.n_det_queries, self.d_model = n_det_queries, d_model
        self.bbox_embed = self.class_embed = None
        self.use_checkpoint, self.use_dab = use_checkpoint, use_dab
        self.query_scale = MLP(d_model, d_model, d_model, 2)
        self.ref_point_head = MLP(2 * d_model, d_model, d_model, 2)

    def forward(self, tgt, reference_points, src, src_spatial_shapes, src_level_start_index, src_valid_ratios,
                query_pos, query_mask, src_padding_mask):
        out, nd = tgt, self.n_det_queries
        outs, refs, queries = [], [], []
        vr4 = torch.cat([src_valid_ratios, src_valid_ratios], -1)[:, None]
        for lid, layer in enumerate(self.layers):
            ref_in = reference_points[:, :, None] * vr4
            raw_pos = self.ref_point_head(pos_to_pos_embed(ref_in[:, :, 0, :], num_pos_feats=self.d_model // 2))
            qpos = raw_pos if lid == 0 else self.query_scale(out) * raw_pos
            queries.append(out)
            merge = lid >= self.merge_det_track_layer
            args = (out, qpos, ref_in, src, src_spatial_shapes, src_level_start_index, query_mask, src_padding_mask, merge)
            out = _ckpt(layer, *args, use_reentrant=False) if (self.use_checkpoint and out.requires_grad) else layer(*args)
            if self.bbox_embed is not None:
                new_ref = (self.bbox_embed[lid](out) + inverse_sigmoid(reference_points)).sigmoid()
                if merge:
                    reference_points = new_ref.detach()
                else:
                    reference_points = torch.cat((new_ref[:, :nd].detach(), reference_points[:, nd:]), dim=1)
            outs.append(out)
            refs.append(reference_points)
        return torch.stack(outs), torch.stack(refs), torch.stack(queries)


class DeformableTransformer(nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, n_feature_levels=4, n_heads=8, n_enc_points=4, n_dec_points=4,
                 n_enc_layers=6, n_dec_layers=6, merge_det_track_layer=0, dropout=0.1, activation="ReLU",
                 return_intermediate_dec=False, n_det_queries=300, extra_track_attn=False, two_stage=False,
                 two_stage_num_proposals=300, use_checkpoint: bool = False, checkpoint_level: int = 2,
                 use_dab: bool = False, visualize: bool = False):
        super().__init__()
        assert not two_stage and use_dab, "two-stage / non-DAB variants are outside the hot-path scope"
        self.d_model, self.n_heads, self.two_stage = d_model, n_heads, two_stage
        self.n_feature_levels, self.n_enc_points, self.n_dec_points = n_feature_levels, n_enc_points, n_dec_points
        self.two_stage_num_proposals, self.use_checkpoint = two_stage_num_proposals, use_checkpoint
        self.checkpoint_level, self.use_dab, self.visualize = checkpoint_level, use_dab, visualize
        enc = DeformableEncoderLayer(d_model, d_ffn, dropout, activation, n_feature_levels, n_heads, n_enc_points, False)
        dec = DeformableDecoderLayer(d_model, d_ffn, dropout, activation, n_feature_levels, n_heads, n_dec_points, False,
                                     extra_track_attn, n_det_queries, visualize)
        self.encoder = DeformableEncoder(enc, n_enc_layers, use_checkpoint and checkpoint_level == 1)
        self.decoder = DeformableDecoder(dec, n_dec_layers, return_intermediate_dec, merge_det_track_layer, n_det_queries,
                                         d_model, use_checkpoint, use_dab, visualize)
        self.level_embed = nn.Parameter(torch.Tensor(n_feature_levels, d_model))
        self.reset_parameters()

    def reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m.reset_parameters()
        nn.init.normal_(self.level_embed)

    @staticmethod
    def get_valid_ratio(mask):
        _, H, W = mask.shape
        return torch.stack([(~mask[:, 0, :]).sum(1).float() / W, (~mask[:, :, 0]).sum(1).float() / H], -1)

    def forward(self, srcs, masks, pos_embeds, query_embed, ref_pts, query_mask):
        assert query_embed is not None and ref_pts is not None
        shapes = [(s.shape[2], s.shape[3]) for s in srcs]
        src = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)
        mask = torch.cat([m.flatten(1) for m in masks], 1)
        pos = torch.cat([p.flatten(2).transpose(1, 2) + self.level_embed[l].view(1, 1, -1)
                         for l, p in enumerate(pos_embeds)], 1)
        spatial_shapes = torch.as_tensor(shapes, dtype=torch.long, device=src.device)
        level_start_index = torch.cat((spatial_shapes.new_zeros(1), spatial_shapes.prod(1).cumsum(0)[:-1]))
        valid_ratios = torch.stack([self.get_valid_ratio(m) for m in masks], 1)
        memory = self.encoder(src, spatial_shapes, level_start_index, valid_ratios, pos, mask)
        init_ref = ref_pts.sigmoid()
        out, refs, queries = self.decoder(query_embed, init_ref, memory, spatial_shapes, level_start_index, valid_ratios,
                                          None, query_mask, mask)
        return out, init_ref, refs, queries

    def get_d_model(self):
        return self.d_model

    def get_n_dec_layers(self):
        return self.decoder.num_layers

    def set_refine_bbox_embed(self, bbox_embed: nn.Module):
        self.decoder.bbox_embed = bbox_embed


def build_transformer(config: dict):
    """Same contract as models/deformable_transformer.py:277-298 (flat upper-case config dict)."""
    return DeformableTransformer(
        d_model=config["HIDDEN_DIM"], d_ffn=config["FFN_DIM"], n_feature_levels=config["NUM_FEATURE_LEVELS"],
        n_heads=config["NUM_HEADS"], n_enc_points=config["NUM_ENC_POINTS"], n_dec_points=config["NUM_DEC_POINTS"],
        n_enc_layers=config["NUM_ENC_LAYERS"], n_dec_layers=config["NUM_DEC_LAYERS"],
        merge_det_track_layer=config.get("MERGE_DET_TRACK_LAYER", 0), dropout=config["DROPOUT"],
        activation=config["ACTIVATION"], return_intermediate_dec=config["RETURN_INTER_DEC"],
        n_det_queries=config["NUM_DET_QUERIES"], extra_track_attn=config["EXTRA_TRACK_ATTN"], two_stage=False,
        use_checkpoint=config["USE_CHECKPOINT"], checkpoint_level=config["CHECKPOINT_LEVEL"], use_dab=config["USE_DAB"],
        visualize=config["VISUALIZE"])


class QueryUpdater(nn.Module):
    """models/query_updater.py:15-255: forward = select_active_tracks + update_tracks_embedding.  `tracks` items are
    duck-typed: any objects with the TrackInstances attributes (ref_pts, query_embed, output_embed, last_output,
    long_memory, logits, boxes, ids, iou) and its `cat_tracked_instances` / `__getitem__` -- the reference's own class
    plugs in unchanged.  Track augmentation (TP_DROP_RATE / FP_INSERT_RATE > 0; 0 in every shipped config) is not built."""

    def __init__(self, hidden_dim: int, ffn_dim: int, tp_drop_ratio: float, fp_insert_ratio: float, dropout: float,
                 use_checkpoint: bool, use_dab: bool, update_threshold: float, long_memory_lambda: float,
                 visualize: bool = False):
        super().__init__()
        assert use_dab and not visualize
        self.hidden_dim, self.ffn_dim, self.dropout = hidden_dim, ffn_dim, dropout
        self.tp_drop_ratio, self.fp_insert_ratio = tp_drop_ratio, fp_insert_ratio
        self.use_checkpoint, self.use_dab, self.visualize = use_checkpoint, use_dab, visualize
        self.update_threshold, self.long_memory_lambda = update_threshold, long_memory_lambda
        self.confidence_weight_net = nn.Sequential(MLP(hidden_dim, hidden_dim, hidden_dim, 2), nn.Sigmoid())
        self.short_memory_fusion = MLP(2 * hidden_dim, 2 * hidden_dim, hidden_dim, 2)
        self.memory_attn = nn.MultiheadAttention(hidden_dim, 8, batch_first=True)
        self.memory_dropout, self.memory_norm = nn.Dropout(dropout), nn.LayerNorm(hidden_dim)
        self.memory_ffn = FFN(hidden_dim, ffn_dim, dropout)
        self.query_feat_dropout, self.query_feat_norm = nn.Dropout(dropout), nn.LayerNorm(hidden_dim)
        self.query_feat_ffn = FFN(hidden_dim, ffn_dim, dropout)
        self.query_pos_head = MLP(2 * hidden_dim, hidden_dim, hidden_dim, 2)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def forward(self, previous_tracks, new_tracks, unmatched_dets, no_augment: bool = False):
        """models/query_updater.py:72-80."""
        tracks = self.select_active_tracks(previous_tracks, new_tracks, unmatched_dets, no_augment=no_augment)
        return self.update_tracks_embedding(tracks=tracks)

    def select_active_tracks(self, previous_tracks, new_tracks, unmatched_dets, no_augment: bool = False):
        """models/query_updater.py:168-254.  Eval (batch 1): newborn tracks get last_output / long_memory from their own
        output / query embeddings, previous + new are concatenated and the dead ones (ids < 0) dropped.  Training without
        augmentation: previous + new + unmatched detections, kept where the score passes update_threshold or the id is
        alive, ids of low-IoU rows cleared; an empty result becomes one fake track (:227-253) so the updater always runs."""
        out = []
        if not self.training:
            assert len(previous_tracks) == 1 and len(new_tracks) == 1, "eval runs batch 1"
            new = new_tracks[0]
            new.last_output, new.long_memory = new.output_embed, new.query_embed
            act = type(new).cat_tracked_instances(previous_tracks[0], new)
            out.append(act[act.ids >= 0])
            return out
        if self.tp_drop_ratio != 0.0 or self.fp_insert_ratio != 0.0:
            raise NotImplementedError("track augmentation (TP_DROP_RATE / FP_INSERT_RATE > 0) is not part of this build")
        for b in range(len(new_tracks)):
            new, um = new_tracks[b], unmatched_dets[b]
            new.last_output, new.long_memory = new.output_embed, new.query_embed
            um.last_output, um.long_memory = um.output_embed, um.query_embed
            cat = type(new).cat_tracked_instances
            act = cat(cat(previous_tracks[b], new), um)
            scores = act.logits.sigmoid().max(dim=1).values
            act = act[(scores > self.update_threshold) | (act.ids >= 0)]
            act.ids[act.iou < 0.5] = -1
            if len(act) == 0:
                dev, C = next(self.query_feat_ffn.parameters()).device, self.hidden_dim
                fake = type(new)(frame_height=1.0, frame_width=1.0, hidden_dim=C).to(dev)
                r = lambda *s: torch.randn(s, dtype=torch.float, device=dev)           # noqa: E731
                fake.query_embed, fake.output_embed, fake.ref_pts, fake.boxes = r(1, C), r(1, C), r(1, 4), r(1, 4)
                fake.ids = torch.as_tensor([-2], dtype=torch.long, device=dev)
                fake.matched_idx = torch.as_tensor([-2], dtype=torch.long, device=dev)
                fake.logits, fake.iou = r(1, act.logits.shape[1]), torch.zeros((1,), dtype=torch.float, device=dev)
                fake.last_output, fake.long_memory = r(1, C), r(1, C)
                act = fake
            out.append(act)
        return out

    def update_tracks_embedding(self, tracks):
        lam = self.long_memory_lambda
        for t in tracks:
            is_pos = t.logits.sigmoid().max(dim=1).values > self.update_threshold
            t.ref_pts[is_pos] = inverse_sigmoid(t.boxes[is_pos].detach().clone())
            qpos = self.query_pos_head(pos_to_pos_embed(t.ref_pts.sigmoid(), num_pos_feats=self.hidden_dim // 2))
            out_e, long_m = t.output_embed, t.long_memory.detach()
            short = self.short_memory_fusion(torch.cat((self.confidence_weight_net(out_e) * out_e, t.last_output), -1))
            att = self.memory_attn((short + qpos)[None], (long_m + qpos)[None], out_e[None])[0][0]
            tgt = self.memory_ffn(self.memory_norm(out_e + self.memory_dropout(att)))
            feat = self.query_feat_ffn(self.query_feat_norm(long_m + self.query_feat_dropout(tgt)))
            m = is_pos[:, None]
            t.long_memory = t.long_memory * ~m + ((1 - lam) * long_m + lam * out_e) * m
            t.last_output = t.last_output * ~m + out_e * m
            t.query_embed[is_pos] = feat[is_pos]
        return tracks


def build_query_updater(config: dict):
    """Same contract as models/query_updater.py:258-270."""
    return QueryUpdater(hidden_dim=config["HIDDEN_DIM"], ffn_dim=config["FFN_DIM"], dropout=config["DROPOUT"],
                        tp_drop_ratio=config.get("TP_DROP_RATE", 0.0), fp_insert_ratio=config.get("FP_INSERT_RATE", 0.0),
                        use_checkpoint=config["USE_CHECKPOINT"], use_dab=config["USE_DAB"],
                        update_threshold=config["UPDATE_THRESH"], long_memory_lambda=config["LONG_MEMORY_LAMBDA"],
                        visualize=config["VISUALIZE"])
