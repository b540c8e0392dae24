"""oracle/tracker.py -- CPU restatement of the reference's per-frame tracker glue.  TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py): imported by tests/ and never by memotr_b200/.

Track instances are plain dicts of tensors with the fields of structures/track_instances.py:11-38 that the eval path
uses: ids, labels, disappear_time (int64), boxes, logits, ref_pts, query_embed, output_embed, last_output, long_memory.

Pinned against the reference's own classes (RuntimeTracker, TrackInstances, QueryUpdater.select_active_tracks) by
oracle/make_golden.py -> tests/golden/tracker.npz, checked in tests/test_oracle_cpu.py.
"""
import torch

FLOAT_FIELDS = ("boxes", "logits", "ref_pts", "query_embed", "output_embed", "last_output", "long_memory")
INT_FIELDS = ("ids", "labels", "disappear_time")


def empty_tracks(C=256, ncls=1):
    t = {k: torch.zeros((0, C)) for k in ("query_embed", "output_embed", "last_output", "long_memory")}
    t["boxes"], t["ref_pts"], t["logits"] = torch.zeros((0, 4)), torch.zeros((0, 4)), torch.zeros((0, ncls))
    for k in INT_FIELDS:
        t[k] = torch.zeros((0,), dtype=torch.long)
    return t


def runtime_tracker_update(out, tracks, max_obj_id, det_thresh, track_thresh, miss_tolerance):
    """RuntimeTracker.update with use_motion=False, use_dab=True (models/runtime_tracker.py:29-101).

    out: dict with pred_logits (Nq, ncls), pred_bboxes (Nq, 4), outputs (Nq, C), last_ref_pts (Nq, 4),
    aux_queries (Nq, C) = aux_outputs[-1]["queries"]; the first n_det rows are the detect queries, then one row per
    track.  Returns (previous_tracks, new_tracks, max_obj_id)."""
    n = len(tracks["ids"])
    n_det = out["pred_logits"].shape[0] - n
    scores = out["pred_logits"].sigmoid()                                     # logits_to_scores, models/utils.py:171
    prev = {k: v.clone() for k, v in tracks.items()}
    prev["boxes"] = out["pred_bboxes"][n_det:].clone()                        # :43-45
    prev["logits"] = out["pred_logits"][n_det:].clone()
    prev["output_embed"] = out["outputs"][n_det:].clone()
    tscores = prev["logits"].sigmoid()
    for i in range(n):                                                        # :47-57
        if tscores[i][prev["labels"][i]] < track_thresh:
            prev["disappear_time"][i] += 1
        else:
            prev["disappear_time"][i] = 0
        if prev["disappear_time"][i] >= miss_tolerance:
            prev["ids"][i] = -1
    idx = torch.max(scores[:n_det], dim=-1).values >= det_thresh             # :60-62
    new = {
        "logits": out["pred_logits"][:n_det][idx], "boxes": out["pred_bboxes"][:n_det][idx],
        "ref_pts": out["last_ref_pts"][:n_det][idx], "output_embed": out["outputs"][:n_det][idx],
        "query_embed": out["aux_queries"][:n_det][idx],                        # :68-69 (use_dab)
    }
    k = int(idx.sum())
    new["disappear_time"] = torch.zeros((k,), dtype=torch.long)               # :76
    new["labels"] = torch.max(scores[:n_det][idx], dim=-1).indices if k else torch.zeros((0,), dtype=torch.long)
    new["ids"] = torch.arange(max_obj_id, max_obj_id + k, dtype=torch.long)  # :85-89
    return prev, new, max_obj_id + k


def select_active_tracks(prev, new):
    """QueryUpdater.select_active_tracks, eval branch (models/query_updater.py:243-254)."""
    new = dict(new)
    new["last_output"] = new["output_embed"]
    new["long_memory"] = new["query_embed"]
    cat = {k: torch.cat((prev[k], new[k]), dim=0) for k in FLOAT_FIELDS + INT_FIELDS}
    keep = cat["ids"] >= 0
    return {k: v[keep] for k, v in cat.items()}


def frame_results(tracks, score_thresh, ori_w, ori_h, area_thresh=100):
    """submit_engine.py:89-102: score filter, area filter, cxcywh -> xyxy in pixels.  Returns (ids, boxes_xyxy, scores)."""
    b = tracks["boxes"]
    area = b[:, 2] * ori_w * b[:, 3] * ori_h
    scores = tracks["logits"].sigmoid()
    keep = (torch.max(scores, dim=-1).values > score_thresh) if len(b) else torch.zeros((0,), dtype=torch.bool)
    keep = keep & (area > area_thresh)
    cx, cy, w, h = b.unbind(-1)
    xyxy = torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)
    xyxy = xyxy * torch.as_tensor([ori_w, ori_h, ori_w, ori_h], dtype=torch.float)
    return tracks["ids"][keep], xyxy[keep], torch.max(scores, dim=-1).values[keep] if len(b) else scores.reshape(0)


def clip_step(sd, cfg, frame, tracks, max_obj_id, det_thresh, track_thresh, miss_tolerance):
    """One iteration of the submit loop (submit_engine.py:64-70) on the functional restatement: model forward with the
    current tracks, RuntimeTracker.update, select_active_tracks, update_tracks_embedding.
    Returns (active tracks, max_obj_id, model outputs, decision margins)."""
    from oracle import frame as oframe
    with torch.no_grad():
        out = oframe.frame_forward(sd, frame["srcs"], frame["masks"], frame["pos"], tracks["ref_pts"],
                                   tracks["query_embed"], cfg)
        o = {"pred_logits": out["pred_logits"][0], "pred_bboxes": out["pred_bboxes"][0], "outputs": out["outputs"][0],
             "last_ref_pts": out["last_ref_pts"][0], "aux_queries": out["aux_queries"][-1, 0]}
        n_det = cfg["n_det_queries"]
        sc = o["pred_logits"].sigmoid()
        margins = [float((sc[:n_det].max(-1).values - det_thresh).abs().min())]
        if len(tracks["ids"]):
            own = sc[n_det:].gather(1, tracks["labels"][:, None])[:, 0]
            margins.append(float((own - track_thresh).abs().min()))
        prev, new, max_obj_id = runtime_tracker_update(o, tracks, max_obj_id, det_thresh, track_thresh, miss_tolerance)
        act = select_active_tracks(prev, new)
        if len(act["ids"]):
            margins.append(float((act["logits"].sigmoid().max(-1).values - cfg["update_thresh"]).abs().min()))
            upd = oframe.update_tracks(sd, act, cfg)
            for k in ("ref_pts", "query_embed", "long_memory", "last_output"):
                act[k] = upd[k]
    return act, max_obj_id, o, min(margins)
