"""memotr_b200/tracker.py -- the reference's per-frame tracker glue, resident on the device.

Host-side mirror of
  * ``TrackInstances``                         /root/reference/structures/track_instances.py:7-38     -> ``TrackTable``
  * ``RuntimeTracker.update``                  /root/reference/models/runtime_tracker.py:29-101
  * ``QueryUpdater.select_active_tracks`` eval /root/reference/models/query_updater.py:243-254        -> ``DeviceTracker.update``
  * result filters of the submit loop          /root/reference/submit_engine.py:89-102                 -> ``DeviceTracker.results``

The reference keeps growing/shrinking tensors and walks them with python loops (one device synchronisation per track
and per field).  Here a track table has a fixed capacity; rows ``[0, n_active)`` are live and ordered as the reference
orders them, the remaining rows are padding (masked as padded keys in the attention kernels), and ``n_active`` is a
device scalar -- so the frame loop never leaves the GPU and can be replayed as one CUDA graph.  Work is done by
``memotr_tracker_update`` / ``memotr_tracker_results`` (include/memotr_b200.h); there is no CPU implementation here.
"""
import ctypes

import torch

from . import _lib

FLOAT_FIELDS = ("query_embed", "output_embed", "last_output", "long_memory", "ref_pts", "boxes", "logits")
INT_FIELDS = ("ids", "labels", "disappear_time")


class TrackTable:
    """Fixed-capacity structure-of-arrays with the TrackInstances fields the eval path uses."""

    def __init__(self, capacity, hidden_dim=256, num_classes=1, device="cuda"):
        if torch.device(device).type != "cuda":
            raise RuntimeError("Not implemented on the CPU")          # as the reference op, src/ms_deform_attn.h:38
        self.capacity, self.C, self.ncls, self.dev = int(capacity), int(hidden_dim), int(num_classes), torch.device(device)
        f = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.dev)        # noqa: E731
        self.fields = {k: f(self.capacity, self.C) for k in FLOAT_FIELDS[:4]}
        self.fields["ref_pts"], self.fields["boxes"] = f(self.capacity, 4), f(self.capacity, 4)
        self.fields["logits"] = f(self.capacity, self.ncls)
        self.fields["ids"] = torch.full((self.capacity,), -1, dtype=torch.long, device=self.dev)
        self.fields["labels"] = torch.zeros(self.capacity, dtype=torch.long, device=self.dev)
        self.fields["disappear_time"] = torch.zeros(self.capacity, dtype=torch.long, device=self.dev)
        self.n_active = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._struct = _lib.TrackTable(**{k: self.fields[k].data_ptr() for k in FLOAT_FIELDS + INT_FIELDS},
                                       n_active=self.n_active.data_ptr())

    def struct(self):
        return ctypes.byref(self._struct)

    def __getitem__(self, k):
        return self.fields[k]

    def clear(self):
        self.n_active.zero_()
        for k in FLOAT_FIELDS:
            self.fields[k].zero_()
        self.fields["ids"].fill_(-1)
        self.fields["labels"].zero_()
        self.fields["disappear_time"].zero_()

    def load(self, tracks):
        """Fill rows [0, n) from a dict of tensors (host or device) with n rows each; missing int fields default to
        ids = arange(n), labels = 0, disappear_time = 0."""
        n = len(tracks["query_embed"])
        if n > self.capacity:
            raise RuntimeError(f"TrackTable.load: {n} tracks exceed the capacity {self.capacity}")
        self.clear()
        for k in FLOAT_FIELDS:
            if k in tracks:
                self.fields[k][:n].copy_(tracks[k])
        self.fields["ids"][:n].copy_(tracks["ids"] if "ids" in tracks else torch.arange(n, device=self.fields["ids"].device))
        for k in ("labels", "disappear_time"):
            if k in tracks:
                self.fields[k][:n].copy_(tracks[k])
        self.n_active.fill_(n)

    def active(self):
        """The live rows as a dict of tensors (synchronises: reads n_active)."""
        n = int(self.n_active.item())
        return {k: v[:n].clone() for k, v in self.fields.items()}


class DeviceTracker:
    """RuntimeTracker + select_active_tracks + result filter on a TrackTable (constructor arguments as
    RuntimeTracker.__init__, runtime_tracker.py:14-17, and Submitter.__init__, submit_engine.py:26)."""

    def __init__(self, table: TrackTable, n_det_queries, det_score_thresh=0.7, track_score_thresh=0.6,
                 miss_tolerance=5, result_score_thresh=0.7, area_thresh=100.0):
        self.table, self.nd = table, int(n_det_queries)
        self.det_score_thresh, self.track_score_thresh = float(det_score_thresh), float(track_score_thresh)
        self.miss_tolerance, self.result_score_thresh = int(miss_tolerance), float(result_score_thresh)
        self.area_thresh = float(area_thresh)
        cap, dev = table.capacity, table.dev
        self.scratch = TrackTable(cap, table.C, table.ncls, dev)
        self.max_obj_id = torch.zeros(1, dtype=torch.long, device=dev)
        self.src_index = torch.full((cap,), -1, dtype=torch.int32, device=dev)
        self.track_pad = torch.ones(cap, dtype=torch.uint8, device=dev)              # empty table: every row is padding
        self.overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        # result rows in ONE buffer [ids int64 | boxes 4 x fp32 | scores fp32 | keep u8] so that they leave with one copy
        self.res_flat = torch.zeros(cap * (8 + 16 + 4 + 1), dtype=torch.uint8, device=dev)
        self.res_ids, self.res_boxes, self.res_scores, self.res_keep = self.split_results(self.res_flat)
        self.res_ids.fill_(-1)
        self.lib = _lib.lib()

    def split_results(self, flat):
        """Views of a flat result buffer (device or host copy): ids (cap) int64, boxes (cap, 4), scores (cap), keep (cap)."""
        cap = self.table.capacity
        return (flat[:cap * 8].view(torch.long), flat[cap * 8:cap * 24].view(torch.float32).view(cap, 4),
                flat[cap * 24:cap * 28].view(torch.float32), flat[cap * 28:cap * 29])

    def reset(self, tracks=None, max_obj_id=0):
        """Start of a clip (submit_engine.py:60-62): empty table, identities from 0 -- or a given state."""
        if tracks is None:
            self.table.clear()
        else:
            self.table.load(tracks)
        n = int(self.table.n_active.item())
        self.track_pad.fill_(1)
        self.track_pad[:n] = 0
        self.max_obj_id.fill_(int(max_obj_id))
        self.overflow.zero_()

    def reset_async(self, tracks, max_obj_id=0):
        """reset(tracks, max_obj_id) without reading anything back from the device (the row count is len(tracks)): usable
        between the frames of a timed loop."""
        self.table.load(tracks)
        n = len(tracks["query_embed"])
        self.track_pad.fill_(1)
        self.track_pad[:n] = 0
        self.max_obj_id.fill_(int(max_obj_id))
        self.overflow.zero_()

    def state_tensors(self):
        """Every tensor of the recurrent tracker state (for snapshot / restore around a warm-up step)."""
        return list(self.table.fields.values()) + [self.table.n_active, self.max_obj_id, self.track_pad, self.overflow,
                                                   self.src_index]

    def update(self, pred_logits, pred_bboxes, outputs, last_ref_pts, aux_queries):
        """One frame.  Every argument is a contiguous fp32 device tensor with n_det + capacity rows (the engine's
        output buffers): detect queries first, then one row per table row."""
        t, nq = self.table, self.nd + self.table.capacity
        for name, x, w in (("pred_logits", pred_logits, t.ncls), ("pred_bboxes", pred_bboxes, 4), ("outputs", outputs, t.C),
                           ("last_ref_pts", last_ref_pts, 4), ("aux_queries", aux_queries, t.C)):
            _lib.require_cuda(**{name: x})
            if x.dtype != torch.float32 or not x.is_contiguous() or tuple(x.shape[-2:]) != (nq, w):
                raise RuntimeError(f"DeviceTracker.update: {name} must be contiguous fp32 ({nq}, {w}), got "
                                   f"{x.dtype} {tuple(x.shape)}")
        fo = _lib.FrameOutputs(pred_logits.data_ptr(), pred_bboxes.data_ptr(), outputs.data_ptr(),
                               last_ref_pts.data_ptr(), aux_queries.data_ptr())
        _lib.check(self.lib.memotr_tracker_update(
            ctypes.byref(fo), self.nd, t.ncls, t.C, t.struct(), self.scratch.struct(), t.capacity,
            self.det_score_thresh, self.track_score_thresh, self.miss_tolerance, self.max_obj_id.data_ptr(),
            self.src_index.data_ptr(), self.track_pad.data_ptr(), self.overflow.data_ptr(),
            _lib.stream_ptr(t.dev)), "tracker_update")

    def results(self, ori_w, ori_h):
        """Enqueue the per-frame result rows into res_ids / res_boxes (xyxy, pixels) / res_scores / res_keep."""
        t = self.table
        _lib.check(self.lib.memotr_tracker_results(
            t.struct(), t.capacity, t.ncls, self.result_score_thresh, self.area_thresh, float(ori_w), float(ori_h),
            self.res_ids.data_ptr(), self.res_boxes.data_ptr(), self.res_scores.data_ptr(), self.res_keep.data_ptr(),
            _lib.stream_ptr(t.dev)), "tracker_results")

    def check_overflow(self):
        """Raise if newborn tracks were dropped because the table was full (synchronises)."""
        n = int(self.overflow.item())
        if n:
            raise RuntimeError(f"DeviceTracker: {n} newborn tracks did not fit into the table (capacity "
                               f"{self.table.capacity}); enlarge n_tracks")
