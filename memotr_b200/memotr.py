"""memotr_b200/memotr.py -- the reference's model surface, `MeMOTR` (models/memotr.py:22-321), on this package's kernels.

Same constructor arguments, parameter names / shapes (a reference checkpoint loads with `load_state_dict`), `forward(frame:
NestedTensor, tracks: list[TrackInstances]) -> dict` with the reference's keys, and `postprocess_single_frame(...)`.  The
backbone (ResNet-50 + position embedding, models/backbone.py) is handed in by the caller -- it is upstream of the hot path
 -- and everything downstream runs here:

  eval mode, batch 1, CUDA   ->  memotr_b200.engine.FrameEngine: one CUDA-graph replay per frame (every kernel our own,
                                 include/memotr_b200.h); a variable number of track queries is served by a fixed-capacity
                                 engine whose unused rows are masked as padded keys
  training / batches / CPU-less autograd  ->  the nn.Module mirrors of memotr_b200/modules.py (MSDeformAttnFunction ->
                                 C ABI -> our forward / backward kernels inside torch autograd)

`frame` / `tracks` are duck-typed: anything with the reference's NestedTensor (`tensors`, `masks`, `decompose()`) and
TrackInstances attributes works, so the reference's own classes plug in unchanged.
"""
import copy
import math

import torch
import torch.nn.functional as F
from torch import nn

from .modules import MLP, build_query_updater, build_transformer, inverse_sigmoid


class _Nested:
    """Minimal stand-in for utils/nested_tensor.py:9-59 used when an extra pyramid level has to be handed to the
    position embedding (models/memotr.py:121)."""

    def __init__(self, tensors, masks):
        self.tensors, self.masks = tensors, masks

    def decompose(self):
        return self.tensors, self.masks


class MeMOTR(nn.Module):
    def __init__(self, backbone, transformer, query_updater, num_classes, n_det_queries, n_feature_levels, hidden_dim,
                 ffn_dim, dropout, aux_loss, with_box_refine, use_checkpoint, checkpoint_level=2, use_dab=True,
                 visualize=False, engine_mode="bf16", engine_tracks=100):
        super().__init__()
        assert use_dab and with_box_refine and not visualize, "the DAB + box-refinement configuration of the released models"
        self.num_classes, self.n_det_queries, self.n_feature_levels = num_classes, n_det_queries, n_feature_levels
        self.hidden_dim, self.ffn_dim, self.dropout, self.aux_loss = hidden_dim, ffn_dim, dropout, aux_loss
        self.with_box_refine, self.use_checkpoint, self.checkpoint_level = with_box_refine, use_checkpoint, checkpoint_level
        self.use_dab, self.visualize = use_dab, visualize
        self.backbone, self.transformer, self.query_updater = backbone, transformer, query_updater
        n_layers = transformer.get_n_dec_layers()
        # heads and queries (models/memotr.py:56-95): same parameter names, shapes and initial values
        class_embed = nn.Linear(hidden_dim, num_classes)
        bbox_embed = MLP(hidden_dim, hidden_dim, 4, 3)
        class_embed.bias.data = torch.ones(num_classes) * (-math.log((1 - 0.01) / 0.01))
        nn.init.constant_(bbox_embed.layers[-1].weight.data, 0)
        nn.init.constant_(bbox_embed.layers[-1].bias.data, 0)
        self.det_anchor = nn.Parameter(torch.randn(n_det_queries, 4))
        self.det_query_embed = nn.Parameter(torch.randn(n_det_queries, hidden_dim))
        n_inter, chans = backbone.n_inter_layers(), backbone.n_inter_channels()
        projs = [nn.Sequential(nn.Conv2d(chans[i], hidden_dim, kernel_size=1), nn.GroupNorm(32, hidden_dim)) for i in range(n_inter)]
        projs += [nn.Sequential(nn.Conv2d(chans[-1], hidden_dim, kernel_size=3, stride=2, padding=1), nn.GroupNorm(32, hidden_dim))
                  for _ in range(n_feature_levels - n_inter)]
        self.feature_projs = nn.ModuleList(projs)
        for proj in self.feature_projs:
            nn.init.xavier_uniform_(proj[0].weight, gain=1)
            nn.init.constant_(proj[0].bias, 0)
        self.class_embed = nn.ModuleList([copy.deepcopy(class_embed) for _ in range(n_layers)])
        self.bbox_embed = nn.ModuleList([copy.deepcopy(bbox_embed) for _ in range(n_layers)])
        nn.init.constant_(self.bbox_embed[0].layers[-1].bias.data[2:], -2.0)
        self.transformer.set_refine_bbox_embed(self.bbox_embed)
        # fast path
        self.engine_mode, self.engine_tracks = engine_mode, int(engine_tracks)
        self._engines = {}

    # ------------------------------------------------------------------------------------------------ pyramid
    def _pyramid(self, frame):
        """models/memotr.py:101-123: backbone features -> input projections -> (srcs, masks, pos) of every level."""
        features, pos = self.backbone(frame)
        pos = list(pos)
        srcs, masks = [], []
        maps = [feat.decompose() for feat in features]
        # inference on one CUDA frame: the projections on our kernels (memotr_b200/input_proj.py); otherwise the nn.Modules
        proj = None
        standard = all(isinstance(p, nn.Sequential) and len(p) == 2 and isinstance(p[0], nn.Conv2d) and isinstance(p[1], nn.GroupNorm)
                       for p in self.feature_projs)
        if (standard and not self.training and not torch.is_grad_enabled()
                and all(m[0].is_cuda and m[0].shape[0] == 1 and m[0].dtype == torch.float32 for m in maps)):
            if getattr(self, "_input_proj", None) is None:
                from .input_proj import InputProj
                self._input_proj = InputProj({k: v for k, v in self.state_dict().items() if k.startswith("feature_projs.")},
                                             maps[0][0].device)
            proj = self._input_proj([m[0] for m in maps])
        for layer, (src, mask) in enumerate(maps):
            srcs.append(proj[layer] if proj is not None else self.feature_projs[layer](src))
            masks.append(mask)
        for layer in range(len(srcs), self.n_feature_levels):
            src = proj[layer] if proj is not None else self.feature_projs[layer](
                features[-1].tensors if layer == len(features) else srcs[-1])
            mask = F.interpolate(frame.masks[None].float(), size=src.shape[-2:])[0].to(torch.bool)
            pos.append(self.backbone.position_embedding(_Nested(src, mask)).to(src.device))
            srcs.append(src)
            masks.append(mask)
        return srcs, masks, pos

    # ------------------------------------------------------------------------------------------------ queries
    def _queries(self, tracks, device):
        """Reference points, query embeddings and the key-padding mask of detect + track queries (models/memotr.py:209-278)."""
        B, n = len(tracks), max(len(t.ref_pts) for t in tracks)
        ref = torch.zeros((B, n, 4), device=device)
        emb = torch.zeros((B, n, self.hidden_dim), device=device)
        pad = torch.zeros((B, n), dtype=torch.bool, device=device)
        for i, t in enumerate(tracks):
            k = len(t.ref_pts)
            ref[i, :k], emb[i, :k] = t.ref_pts.to(device), t.query_embed.to(device)
            if k > 0:
                pad[i, k:] = True
        ref = torch.cat((self.det_anchor[None].expand(B, -1, -1), ref), 1)
        emb = torch.cat((self.det_query_embed[None].expand(B, -1, -1), emb), 1)
        pad = torch.cat((torch.zeros((B, self.n_det_queries), dtype=torch.bool, device=device), pad), 1)
        return ref, emb, pad

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self, frame, tracks):
        srcs, masks, pos = self._pyramid(frame)
        if self._engine_eligible(srcs, tracks):
            return self._forward_engine(srcs, masks, pos, tracks[0])
        ref, emb, pad = self._queries(tracks, srcs[0].device)
        outputs, init_ref, inter_refs, inter_queries = self.transformer(srcs=srcs, masks=masks, pos_embeds=pos,
                                                                       query_embed=emb, ref_pts=ref, query_mask=pad)
        logits, boxes = [], []
        for lvl in range(outputs.shape[0]):             # models/memotr.py:147-162
            r = inverse_sigmoid(init_ref if lvl == 0 else inter_refs[lvl - 1])
            logits.append(self.class_embed[lvl](outputs[lvl]))
            boxes.append((self.bbox_embed[lvl](outputs[lvl]) + r).sigmoid())
        logits, boxes = torch.stack(logits), torch.stack(boxes)
        res = {"pred_logits": logits[-1], "pred_bboxes": boxes[-1], "last_ref_pts": inverse_sigmoid(inter_refs[-2]),
               "query_mask": pad, "det_query_embed": emb[0][:self.n_det_queries], "init_ref_pts": inverse_sigmoid(init_ref)}
        if self.aux_loss:
            res["aux_outputs"] = [{"pred_logits": a, "pred_bboxes": b, "query_mask": pad, "queries": c}
                                  for a, b, c in zip(logits[:-1], boxes[:-1], inter_queries[1:])]
        res["outputs"] = outputs[-1]
        return res

    def postprocess_single_frame(self, previous_tracks, new_tracks, unmatched_dets, no_augment=False):
        """Query updating (models/memotr.py:280-287)."""
        return self.query_updater(previous_tracks, new_tracks, unmatched_dets, no_augment)

    # ------------------------------------------------------------------------------------------------ engine route
    def _engine_eligible(self, srcs, tracks):
        t = self.transformer
        return (not self.training and not torch.is_grad_enabled() and len(tracks) == 1 and srcs[0].is_cuda
                and srcs[0].shape[0] == 1 and self.engine_mode in ("fp32", "bf16", "fp32tc") and self.hidden_dim == 256
                and t.n_heads == 8 and len(srcs) == t.n_feature_levels)

    def hot_path_config(self):
        t = self.transformer
        return dict(d_model=self.hidden_dim, d_ffn=self.ffn_dim, n_levels=t.n_feature_levels, n_heads=t.n_heads,
                    n_enc_points=t.n_enc_points, n_dec_points=t.n_dec_points, n_enc_layers=t.encoder.num_layers,
                    n_dec_layers=t.decoder.num_layers, merge_det_track_layer=t.decoder.merge_det_track_layer,
                    n_det_queries=self.n_det_queries, update_thresh=self.query_updater.update_threshold,
                    long_memory_lambda=self.query_updater.long_memory_lambda, num_classes=self.num_classes)

    def reset_engines(self):
        """Drop the cached engines (call after changing the weights: an engine packs them once)."""
        self._engines = {}
        self._input_proj = None

    def _engine(self, shapes, n_tracks, device):
        from .engine import FrameEngine
        cap = self.engine_tracks
        while cap < n_tracks:
            cap *= 2
        key = (tuple(shapes), cap, str(device), self.engine_mode)
        if key not in self._engines:
            sd = {k: v.detach() for k, v in self.state_dict().items()
                  if not k.startswith(("backbone.", "feature_projs.", "transformer.decoder.bbox_embed."))}
            eng = FrameEngine(sd, self.hot_path_config(), shapes, cap, device, mode=self.engine_mode, pad_tracks=True)
            eng.capture(eng.forward)
            self._engines[key] = eng
        return self._engines[key]

    def _forward_engine(self, srcs, masks, pos, tracks):
        nd, n, dev = self.n_det_queries, len(tracks.ref_pts), srcs[0].device
        shapes = [tuple(s.shape[-2:]) for s in srcs]
        eng = self._engine(shapes, n, dev)
        ref = torch.zeros((eng.nt, 4), device=dev)
        emb = torch.zeros((eng.nt, self.hidden_dim), device=dev)
        ref[:n], emb[:n] = tracks.ref_pts.to(dev), tracks.query_embed.to(dev)
        eng.load_frame(srcs, masks, pos, ref, emb)
        eng.query_pad[nd:nd + n] = 0
        eng.query_pad[nd + n:] = 1
        eng.replay()
        L, nq = eng.n_dec, nd + n
        c = lambda t: t[:nq].float().clone()[None]                                     # noqa: E731
        pad = torch.zeros((1, nq), dtype=torch.bool, device=dev)
        res = {"pred_logits": c(eng.pred_logit[L - 1]), "pred_bboxes": c(eng.pred_box[L - 1]), "last_ref_pts": c(eng.last_ref_pts),
               "query_mask": pad, "det_query_embed": self.det_query_embed, "init_ref_pts": c(eng.init_ref_pts)}
        if self.aux_loss:
            res["aux_outputs"] = [{"pred_logits": c(eng.pred_logit[l]), "pred_bboxes": c(eng.pred_box[l]), "query_mask": pad,
                                   "queries": c(eng.tgt32[l + 1])} for l in range(L - 1)]
        res["outputs"] = c(eng.tgt32[L])
        return res


def build(config: dict, backbone):
    """models/memotr.py:290-321 with the backbone handed in (build it with the reference's models.backbone.build)."""
    num_classes = {"DanceTrack": 1, "SportsMOT": 1, "MOT17": 1, "MOT17_SPLIT": 1, "BDD100K": 8}[config["DATASET"]]
    return MeMOTR(backbone=backbone, transformer=build_transformer(config), query_updater=build_query_updater(config),
                  num_classes=num_classes, n_det_queries=config["NUM_DET_QUERIES"],
                  n_feature_levels=config["NUM_FEATURE_LEVELS"], hidden_dim=config["HIDDEN_DIM"], ffn_dim=config["FFN_DIM"],
                  dropout=config["DROPOUT"], aux_loss=True, with_box_refine=True, use_checkpoint=config["USE_CHECKPOINT"],
                  checkpoint_level=config["CHECKPOINT_LEVEL"], use_dab=config["USE_DAB"], visualize=config["VISUALIZE"],
                  engine_mode=config.get("ENGINE_MODE", "bf16"))
