"""memotr_b200/engine.py -- the per-frame hot path of MeMOTR on one B200, every kernel our own.

`FrameEngine` runs what `MeMOTR.forward` does after the backbone / input projections
(/root/reference/models/memotr.py:128-195: reference-point/query assembly, DeformableTransformer.forward
deformable_transformer.py:192-259, the per-layer heads) and `QueryUpdater.update_tracks_embedding`
(models/query_updater.py:82-166), for batch size 1 and the DAB configuration the released MeMOTR checkpoints use
(USE_DAB True, iterative box refinement, MERGE_DET_TRACK_LAYER 1).  It consumes the reference's own state_dict (same
keys/shapes), packs the weights once into the layouts the kernels want, owns a fixed workspace in HBM, and issues
nothing but C-ABI calls (include/memotr_b200.h) on the current CUDA stream -- so a whole frame can be captured into a
CUDA graph (`capture()`), which removes the per-launch host cost and the host syncs listed in SURVEY.md 2.2.

Two arithmetic modes:
  "fp32"  fp32 storage, fp32 CUDA-core GEMMs -> matches the reference's fp32 (TF32-off) path to <= 1e-4
  "bf16"  bf16 GEMM operands (activations fed to GEMMs, weights, value maps), tcgen05 tensor-core GEMMs with fp32
          accumulation; the residual stream, the pre-LayerNorm sums, LayerNorm itself and all geometry (reference points,
          sampling locations, boxes, logits) stay fp32 -> <= 1e-2
"""
import math

import torch

from . import _lib

F32, BF16, F16 = _lib.F32, _lib.BF16, _lib.F16


def _p(t):
    return _lib.ptr(t)


class _Lin:
    """One nn.Linear packed for the kernels: weight (N,K) in the activation dtype, bias fp32."""

    def __init__(self, w, b, dtype, device):
        self.N, self.K = w.shape
        self.w = w.to(device=device, dtype=dtype).contiguous()
        self.b = b.to(device=device, dtype=torch.float32).contiguous() if b is not None else None


class FrameEngine:
    def __init__(self, state_dict, cfg, shapes, n_tracks, device="cuda", mode="fp32", tracker=None,
                 ori_size=(1920, 1080), pos_embed=None, pad_tracks=False):
        """`n_tracks` is the number of track-query rows (= the capacity of the track table).  `tracker=None`: the
        caller owns the track bookkeeping and every row is live (the reference's model + query-updater path only).
        `tracker=dict(det_score_thresh=, track_score_thresh=, miss_tolerance=, result_score_thresh=)`: the
        RuntimeTracker glue runs on the device inside step() (memotr_b200/tracker.py) with a variable number of live
        rows; `ori_size` = (width, height) of the original image for the result boxes (submit_engine.py:89-98)."""
        # pos_embed=dict(temperature=20, scale=2*pi): the position maps are rebuilt on the device from the padding masks
        # (PositionEmbeddingSine, models/position_embedding.py:23-49) instead of being an input (None)
        self.pos_cfg = dict(pos_embed) if pos_embed is not None else None
        # pad_tracks: the caller owns the tracks but uses fewer than n_tracks rows: query_pad (device, uint8 per query, 1 =
        # unused row) masks them as padded keys, exactly like the device tracker's padding
        self.use_pad = tracker is not None or bool(pad_tracks)
        assert mode in ("fp32", "bf16", "fp32tc")
        # "fp32tc": the fp32 engine with its large-M nn.Linear call sites (the encoder's projections and FFN, the stacked decoder
        # value projection) on the tensor cores at fp32 accuracy -- two-term fp16 operand splits, three products, fp32
        # accumulation (memotr_linear_f32x3); everything else is the fp32 engine
        self.tc3 = mode == "fp32tc"
        mode = "fp32" if self.tc3 else mode
        self.tracker_cfg, self.ori_size = (dict(tracker) if tracker is not None else None), tuple(ori_size)
        self.cfg, self.mode = dict(cfg), mode
        self.dev = torch.device(device)
        self.ta = torch.float32 if mode == "fp32" else torch.bfloat16
        self.dt = F32 if mode == "fp32" else BF16
        self.shapes = [(int(h), int(w)) for h, w in shapes]
        self.L = len(self.shapes)
        self.C, self.H, self.Fd = cfg["d_model"], cfg["n_heads"], cfg["d_ffn"]
        assert self.C == 256 and self.C // self.H == 32, "kernels are specialised for d_model 256 / head dim 32"
        assert self.L == cfg["n_levels"]
        self.S = sum(h * w for h, w in self.shapes)
        self.nd, self.nt = cfg["n_det_queries"], int(n_tracks)
        self.nq = self.nd + self.nt
        self.ncls = cfg["num_classes"]
        self.n_enc, self.n_dec = cfg["n_enc_layers"], cfg["n_dec_layers"]
        self.merge = cfg["merge_det_track_layer"]
        self.lib = _lib.lib()
        self.launches = 0
        import os
        self.fused_mlp = os.environ.get("MEMOTR_FUSED_MLP", "1") != "0"   # A/B switch for the on-chip FFN/MLP kernel
        self.tokens_levels = os.environ.get("MEMOTR_TOKENS_LEVELS", "1") != "0"   # all levels' tokens in one launch
        # value maps in fp16 (bf16 mode): more mantissa than bf16 for this read-only intermediate, and the gather can
        # blend corners with packed HFMA2 instead of widening every bf16 element on the (binding) ALU pipe; A/B switch
        self.value_f16 = mode == "bf16" and os.environ.get("MEMOTR_VALUE_F16", "1") != "0"
        self.vdt = F16 if self.value_f16 else self.dt
        self.tv = torch.float16 if self.value_f16 else self.ta
        # the whole decoder + heads as ONE persistent kernel, a 4-CTA cluster per block of 16 query rows
        # (csrc/decoder_cluster.cu); MEMOTR_DEC_FUSED=0: one launch per op (A/B, and the path of the fp32 modes)
        self.dec_fused = (mode == "bf16" and self.value_f16 and os.environ.get("MEMOTR_DEC_FUSED", "1") != "0"
                          and (self.nq + 15) // 16 * 4 <= 132 and cfg["d_ffn"] % 256 == 0 and cfg["d_ffn"] <= 2048
                          and cfg["n_levels"] * cfg["n_dec_points"] == 16)
        self.dec_cluster = self.dec_fused
        # encoder: sampling locations / softmax weights computed in the epilogue of the offsets+logits GEMM (A/B switch
        # MEMOTR_FUSE_PREP=0).  With four epilogue warps in the persistent GEMM this was 7.5 us per layer SLOWER than the
        # separate prep kernel (the kernel is epilogue-bound); with eight it is 9.5 us per layer faster
        n_sm = torch.cuda.get_device_properties(self.dev).multi_processor_count if self.dev.type == "cuda" else 0
        self.fuse_prep = (mode == "bf16" and self.value_f16 and os.environ.get("MEMOTR_FUSE_PREP", "1") != "0"
                          and cfg["n_levels"] * cfg["n_enc_points"] == 16 and self.H % 2 == 0
                          and 3 * ((self.S + 127) // 128) > n_sm > 0 and self.H * 48 % 128 == 0)
        # encoder: output_proj + norm1 as one kernel (memotr_linear256_layernorm; MEMOTR_FUSE_OUTLN=0: GEMM + LayerNorm launches)
        self.fuse_outln = (mode == "bf16" and os.environ.get("MEMOTR_FUSE_OUTLN", "0") != "0" and self.fused_mlp
                           and self.C == 256 and self.S >= 2048 and n_sm > 0)
        # encoder: norm2 (+ the fp32 master, + the next layer's query = y + pos) in the EPILOGUE of the fused FFN kernel
        # (memotr_mlp2_lnout): no LayerNorm launch, no fp32 round trip of the pre-norm sum.  Opt-in (MEMOTR_FUSE_LN2=1):
        # measured 9.5 us per layer SLOWER -- the LayerNorm phase of a tile runs at HBM speed on every SM at once with the
        # tensor pipes idle, and a one-tile-per-CTA kernel has nothing to overlap it with (tools/micro_dense.py)
        self.fuse_ln2 = (mode == "bf16" and os.environ.get("MEMOTR_FUSE_LN2", "0") != "0" and self.fused_mlp
                         and self.Fd % 128 == 0 and self.S >= 2048 and n_sm > 0)
        self.n_sm = n_sm
        # encoder: output_proj + norm1 + FFN + norm2 as ONE kernel per 128-row tile (memotr_encoder_dense_block: the front GEMM
        # and norm1 run inside the FFN kernel, its input tile never exists in HBM).  MEMOTR_FUSE_BLOCK=0: A/B
        self.fuse_block = self.fuse_ln2 and os.environ.get("MEMOTR_FUSE_BLOCK", "0") != "0"
        # encoder gather from TMA-staged value-map windows in shared memory (csrc/msda_window.cu); MEMOTR_MSDA_WINDOW=0: the
        # global-memory gather (bit-identical results)
        self.msda_window = (self.fuse_prep and self.L <= 5 and self.H <= 16 and cfg["n_enc_points"] % 2 == 0
                            and os.environ.get("MEMOTR_MSDA_WINDOW", "1") != "0")
        self.window_classes = int(os.environ.get("MEMOTR_WINDOW_CLASSES", "0"))
        self.window_stats = None             # device int64[2] when profiling: taps served from windows / from global memory
        self._pack(state_dict)
        self._alloc()
        if self.dec_fused:
            self._build_decoder_program()
            self._build_decoder_program_cluster()
        # the one-CTA-per-row-block decoder (csrc/decoder_fused.cu) for frame-pipelined clips; MEMOTR_DEC_SINGLE=0: never
        F_ = cfg["d_ffn"]
        self.dec_single_ok = (self.dec_fused and os.environ.get("MEMOTR_DEC_SINGLE", "1") != "0"
                              and (F_ % 512 == 0 if F_ > 1024 else F_ % 256 == 0))
        self.dec_use_single = False
        if self.dec_single_ok:
            self._build_decoder_program_single()
        # the query updater as one persistent cluster kernel (csrc/updater_cluster.cu); A/B switch MEMOTR_UPD_FUSED=0
        self.upd_fused = (self.dec_cluster and os.environ.get("MEMOTR_UPD_FUSED", "1") != "0"
                          and 1 <= self.nt and (self.nt + 15) // 16 * 4 <= 132)
        if self.upd_fused:
            self._build_updater_program()
        self.graph = None
        self.timer = None
        self.debug_enc = None        # tests: a list that forward() fills with the fp32 encoder state before / after every layer

    # ------------------------------------------------------------------------------------------------ measurement
    def enable_msda_timer(self):
        """Bracket every encoder-layer MSDA forward launch with a pair of (graph-capturable) CUDA events, and mark the
        section boundaries of a step (prep | encoder | decoder | updater) with five more."""
        self.timer = self.lib.memotr_timer_create(4 * self.n_enc + 5)     # [msda pairs | 5 section marks | ffn pairs]
        self._timer_slot = 0

    def _mark(self, i):
        if self.timer is not None:
            _lib.check(self.lib.memotr_timer_record(self.timer, 2 * self.n_enc + i, self._st()), "timer_record")

    def section_times_us(self):
        """(prep, encoder, decoder+heads, updater) durations of the most recent step."""
        import ctypes
        ms, out, b = ctypes.c_float(), {}, 2 * self.n_enc
        for i, name in enumerate(("prep", "encoder", "decoder_heads", "updater")):
            _lib.check(self.lib.memotr_timer_elapsed_ms(self.timer, b + i, b + i + 1, ctypes.byref(ms)), "timer_elapsed")
            out[name] = ms.value * 1e3
        return out

    def ffn_times_us(self):
        """Durations of the encoder FFN launches (fused mlp2, incl. its tail-split launch) of the most recent step."""
        import ctypes
        out, ms, b = [], ctypes.c_float(), 2 * self.n_enc + 5
        for i in range(self.n_enc):
            _lib.check(self.lib.memotr_timer_elapsed_ms(self.timer, b + 2 * i, b + 2 * i + 1, ctypes.byref(ms)), "timer_elapsed")
            out.append(ms.value * 1e3)
        return out

    def msda_times_us(self):
        """Durations of the encoder MSDA forward launches of the most recent step (call after a synchronize)."""
        import ctypes
        out = []
        ms = ctypes.c_float()
        for i in range(self.n_enc):
            _lib.check(self.lib.memotr_timer_elapsed_ms(self.timer, 2 * i, 2 * i + 1, ctypes.byref(ms)), "timer_elapsed")
            out.append(ms.value * 1e3)
        return out

    # ------------------------------------------------------------------------------------------------ weights
    def _window_hint(self, bias, weight):
        """Where the windowed gather should stage value-map windows for one MSDeformAttn: per (head, level) the mean
        sampling offset over the K points (pixels of that level: ms_deform_attn.py:115-117 divides the offsets by the
        level's extent), and a radius around it = spread of the bias over the points (90th percentile) + two standard
        deviations of the query-dependent part (row norms of sampling_offsets.weight x an r.m.s. query of 1.5) + 0.5.
        Only a hint: taps outside a window are read from global memory with the same arithmetic."""
        import ctypes
        import os
        H, L = self.H, self.L
        b = bias.detach().float().cpu().view(H, L, -1, 2)
        shift = b.mean(2)                                           # (H, L, 2)
        dev = (b - shift[:, :, None]).abs().flatten()
        spread = float(torch.quantile(dev, 0.9)) if dev.numel() > 1 else 0.0
        wn = float(weight.detach().float().norm(dim=1).max())
        radius = min(max(spread + 2.0 * wn * 1.5 + 0.5, 1.5), 6.0)
        if os.environ.get("MEMOTR_WINDOW_RADIUS"):
            radius = float(os.environ["MEMOTR_WINDOW_RADIUS"])
        return (ctypes.c_float * (H * L * 2))(*[float(v) for v in shift.reshape(-1).tolist()]), radius

    def _pack(self, sd):
        dev, ta = self.dev, self.ta
        lin = lambda k: _Lin(sd[k + ".weight"], sd[k + ".bias"], ta, dev)          # noqa: E731
        f32 = lambda k: sd[k].to(device=dev, dtype=torch.float32).contiguous()     # noqa: E731

        def msda(k):
            ol_w = torch.cat([sd[k + ".sampling_offsets.weight"], sd[k + ".attention_weights.weight"]], 0)
            ol_b = torch.cat([sd[k + ".sampling_offsets.bias"], sd[k + ".attention_weights.bias"]], 0)
            d = {"ol": _Lin(ol_w, ol_b, ta, dev), "value": lin(k + ".value_proj"), "out": lin(k + ".output_proj")}
            d["win_shift"], d["win_radius"] = self._window_hint(sd[k + ".sampling_offsets.bias"],
                                                                sd[k + ".sampling_offsets.weight"])
            return d

        def ln(k):
            return f32(k + ".weight"), f32(k + ".bias")

        def mha(k):
            w, b = sd[k + ".in_proj_weight"], sd[k + ".in_proj_bias"]
            C = self.C
            return {"q": _Lin(w[:C], b[:C], ta, dev), "k": _Lin(w[C:2 * C], b[C:2 * C], ta, dev),
                    "qk": _Lin(w[:2 * C], b[:2 * C], ta, dev), "v": _Lin(w[2 * C:], b[2 * C:], ta, dev),
                    "out": lin(k + ".out_proj")}

        self.level_embed = f32("transformer.level_embed")
        self.enc = []
        for i in range(self.n_enc):
            k = f"transformer.encoder.layers.{i}"
            self.enc.append({"attn": msda(k + ".self_attn"), "norm1": ln(k + ".norm1"), "lin1": lin(k + ".linear1"),
                             "lin2": lin(k + ".linear2"), "norm2": ln(k + ".norm2")})
        self.dec = []
        for i in range(self.n_dec):
            k = f"transformer.decoder.layers.{i}"
            self.dec.append({"self": mha(k + ".self_attn"), "norm2": ln(k + ".norm2"), "attn": msda(k + ".cross_attn"),
                             "norm1": ln(k + ".norm1"), "lin1": lin(k + ".linear1"), "lin2": lin(k + ".linear2"),
                             "norm3": ln(k + ".norm3"),
                             "bbox": [lin(f"bbox_embed.{i}.layers.{j}") for j in range(3)],
                             "cls": lin(f"class_embed.{i}")})
        # value projections of all decoder layers share the input (the encoder memory): one stacked GEMM
        self.dec_value = _Lin(torch.cat([sd[f"transformer.decoder.layers.{i}.cross_attn.value_proj.weight"]
                                         for i in range(self.n_dec)], 0),
                              torch.cat([sd[f"transformer.decoder.layers.{i}.cross_attn.value_proj.bias"]
                                         for i in range(self.n_dec)], 0), ta, dev)
        self.query_scale = [lin(f"transformer.decoder.query_scale.layers.{j}") for j in range(2)]
        self.ref_point_head = [lin(f"transformer.decoder.ref_point_head.layers.{j}") for j in range(2)]
        self.det_anchor = f32("det_anchor")
        self.det_query_embed = f32("det_query_embed")
        q = "query_updater"
        self.upd = {
            "conf": [lin(f"{q}.confidence_weight_net.0.layers.{j}") for j in range(2)],
            "fusion": [lin(f"{q}.short_memory_fusion.layers.{j}") for j in range(2)],
            "pos_head": [lin(f"{q}.query_pos_head.layers.{j}") for j in range(2)],
            "attn": mha(f"{q}.memory_attn"), "memory_norm": ln(f"{q}.memory_norm"),
            "mffn": (lin(f"{q}.memory_ffn.linear1"), lin(f"{q}.memory_ffn.linear2"), ln(f"{q}.memory_ffn.norm")),
            "feat_norm": ln(f"{q}.query_feat_norm"),
            "fffn": (lin(f"{q}.query_feat_ffn.linear1"), lin(f"{q}.query_feat_ffn.linear2"),
                     ln(f"{q}.query_feat_ffn.norm")),
        }
        i = torch.arange(128, dtype=torch.float32)                                   # models/utils.py:80-81
        self.dim_t = (10000 ** (2 * torch.div(i, 2, rounding_mode="trunc") / 128)).to(dev)

    # ------------------------------------------------------------------------------------------------ workspace
    def _alloc(self):
        dev, ta, S, C, nq, nt = self.dev, self.ta, self.S, self.C, self.nq, self.nt
        e = lambda *s, dtype=ta: torch.empty(*s, dtype=dtype, device=dev)            # noqa: E731
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)              # noqa: E731
        LK = self.L * max(self.cfg["n_enc_points"], self.cfg["n_dec_points"])
        self.shapes_t = torch.as_tensor(self.shapes, dtype=torch.long, device=dev)
        # fp32tc: scratch for the split A operand [hi | hi | lo] of the widest large-M GEMM (linear2: K = d_ffn)
        tc3 = getattr(self, "tc3", False)
        self.a3 = torch.empty(S * 3 * max(self.Fd, C) if tc3 else 0, dtype=torch.float16, device=dev)
        self.a3h = torch.empty(S * 3 * self.Fd if tc3 else 0, dtype=torch.float16, device=dev)     # linear1's split output
        sizes = [h * w for h, w in self.shapes]
        self.lsi_host = [sum(sizes[:i]) for i in range(self.L)]
        self.lsi_t = torch.as_tensor(self.lsi_host, dtype=torch.long, device=dev)
        # static inputs (filled by load_frame; fixed addresses so a captured graph can be replayed)
        # one flat buffer [src (fp32) per level | pos (fp32) per level, if it is an input | masks (u8) per level] so that a
        # staged frame arrives with ONE device-to-device copy (ClipRunner); in_src / in_pos / in_mask are views
        self.in_layout = self.input_layout(self.shapes, C, with_pos=self.pos_cfg is None)
        self.in_flat = torch.zeros(self.in_layout["bytes"], dtype=torch.uint8, device=dev)
        self.in_src, self.in_pos, self.in_mask = self.input_views(self.in_flat, self.in_layout, self.shapes, C)
        if self.pos_cfg is not None:
            self.pos_scratch = f(2 * self.S)
            i = torch.arange(C // 2, dtype=torch.float32)                              # models/position_embedding.py:33-34
            self.pos_dim_i = (float(self.pos_cfg.get("temperature", 20)) **
                              (2 * torch.div(i, 2, rounding_mode="trunc") / (C // 2))).to(dev)
        self.in_track_ref = torch.zeros(nt, 4, dtype=torch.float32, device=dev)
        self.in_track_embed = torch.zeros(nt, C, dtype=torch.float32, device=dev)
        # encoder
        m0 = self.in_layout["mask"][0]
        self.mask_flat = self.in_flat[m0:m0 + S]        # the levels' masks are consecutive in the input buffer: a view
        self.vr = f(self.L, 2)
        self.src_tok, self.pos_tok, self.q_tok = e(S, C), e(S, C), e(S, C)
        self.value = e(S, C, dtype=self.tv)
        self.ol = f(S, 3 * self.H * LK)
        self.loc = f(S, self.H, LK, 2)
        self.attw = f(S, self.H, LK)
        self.att = e(S, C)
        self.pre = f(S, C)                                           # pre-LayerNorm GEMM output, always fp32
        self.src1 = e(S, C)
        # fp32 residual stream: in fp32 mode the activation buffers are their own masters
        fp32 = self.mode == "fp32"
        self.src32 = self.src_tok if fp32 else f(S, C)
        self.src1_32 = self.src1 if fp32 else f(S, C)
        self.hid = e(S, self.Fd)
        self.value_all = e(S, self.n_dec * C, dtype=self.tv)
        # decoder (Nq rows)
        self.ref = [f(nq, 4) for _ in range(self.n_dec + 1)]         # ref[0] = init (sigmoid), ref[l+1] after layer l
        self.pred_box = [f(nq, 4) for _ in range(self.n_dec)]
        self.pred_logit = [f(nq, self.ncls) for _ in range(self.n_dec)]
        self.tgt = [e(nq, C) for _ in range(self.n_dec + 1)]         # tgt[0] = queries, tgt[l+1] = output of layer l
        self.tgt32 = self.tgt if fp32 else [f(nq, C) for _ in range(self.n_dec + 1)]
        self.ref_raw = f(nq, 4)
        self.vr_scale4 = f(4)
        self.anchor = e(nq, 2 * C)
        self.d_a, self.d_b, self.d_c = e(nq, C), e(nq, C), e(nq, C)
        self.query_pos = e(nq, C)
        self.qk_in = e(nq, C)
        self.qk = f(nq, 2 * C)                                        # projected q|k and v stay fp32 (see mha())
        self.v = f(nq, C)
        self.t1, self.t1q, self.t2 = e(nq, C), e(nq, C), e(nq, C)
        self.t1_32 = self.t1 if fp32 else f(nq, C)
        self.t2_32 = self.t2 if fp32 else f(nq, C)
        self.d_pre = f(nq, C)
        self.d_hid = e(nq, self.Fd)
        self.delta = f(nq, 4)
        self.last_ref_pts, self.init_ref_pts = f(nq, 4), f(nq, 4)
        # query updater (Nt rows); the track state itself is fp32 (TrackInstances fields)
        from .tracker import TrackTable, DeviceTracker, FLOAT_FIELDS
        self.table = TrackTable(nt, C, self.ncls, dev)
        self.st = {k: self.table[k] for k in FLOAT_FIELDS}
        self.query_pad = torch.zeros(nq, dtype=torch.uint8, device=dev)     # key-padding mask of the decoder queries
        self.trk = None
        if self.tracker_cfg is not None:
            self.trk = DeviceTracker(self.table, self.nd, **self.tracker_cfg)
            self.trk.track_pad = self.query_pad[self.nd:]                    # written by the tracker, read by the MHAs
            self.trk.reset()
        self.is_pos = torch.zeros(nt, dtype=torch.uint8, device=dev)
        self.u_ref = f(nt, 4)
        self.u_sine = e(nt, 2 * C)
        self.u_oe, self.u_last, self.u_long = e(nt, C), e(nt, C), e(nt, C)
        self.u_cat = e(nt, 2 * C)
        self.u_a, self.u_b, self.u_c, self.u_d = e(nt, C), e(nt, C), e(nt, C), e(nt, C)
        self.u_big = e(nt, 2 * C)
        self.u_q, self.u_k, self.u_v = f(nt, C), f(nt, C), f(nt, C)
        self.u_hid = e(nt, self.Fd)
        self.u_pre = f(nt, C)
        self.u_a32 = self.u_a if fp32 else f(nt, C)
        self.u_c32 = self.u_c if fp32 else f(nt, C)

    # ------------------------------------------------------------------------------------------------ fused decoder
    def _build_decoder_program(self):
        """Parameter block shared by the fused decoder kernels (memotr_dec_params, include/memotr_b200.h)."""
        import ctypes
        C, nq, dev, nl = self.C, self.nq, self.dev, self.n_dec
        F = self.Fd
        self.dec_np = (nq + 63) // 64 * 64
        self.dec_kbuf = torch.zeros(2, self.dec_np, C, dtype=torch.float16, device=dev)
        self.dec_vbuf = torch.zeros(2, C, self.dec_np, dtype=torch.float16, device=dev)
        self.dec_barrier = torch.zeros(1, dtype=torch.int32, device=dev)
        P = _lib.DecParams()
        P.prog, P.n_prog, P.n_layers, P.nq, P.nd, P.merge = None, 0, nl, nq, self.nd, self.merge
        P.ncls, P.n_levels, P.n_points, P.d_ffn = self.ncls, self.L, self.cfg["n_dec_points"], F
        P.value_stride, P.np = nl * C, self.dec_np
        P.rph0_b, P.rph1_b = self.ref_point_head[0].b.data_ptr(), self.ref_point_head[1].b.data_ptr()
        P.qs0_b, P.qs1_b = self.query_scale[0].b.data_ptr(), self.query_scale[1].b.data_ptr()
        P.tgt_in, P.ref_in = self.tgt32[0].data_ptr(), self.ref[0].data_ptr()
        P.vr_scale4, P.valid_ratios, P.dim_t = self.vr_scale4.data_ptr(), self.vr.data_ptr(), self.dim_t.data_ptr()
        P.query_pad = self.query_pad.data_ptr() if self.use_pad else None
        P.kbuf, P.vbuf, P.barrier = self.dec_kbuf.data_ptr(), self.dec_vbuf.data_ptr(), self.dec_barrier.data_ptr()
        for l, (h, w) in enumerate(self.shapes):
            P.shapes[2 * l], P.shapes[2 * l + 1], P.lsi[l] = h, w, self.lsi_host[l]
        for lid, ly in enumerate(self.dec):
            D = P.layers[lid]
            D.qk_b, D.v_b, D.sao_b = ly["self"]["qk"].b.data_ptr(), ly["self"]["v"].b.data_ptr(), ly["self"]["out"].b.data_ptr()
            D.ol_b, D.cao_b = ly["attn"]["ol"].b.data_ptr(), ly["attn"]["out"].b.data_ptr()
            D.f1_b, D.f2_b = ly["lin1"].b.data_ptr(), ly["lin2"].b.data_ptr()
            D.bb0_b, D.bb1_b, D.bb2_b = (ly["bbox"][j].b.data_ptr() for j in range(3))
            D.cls_b, D.bb2_w, D.cls_w = ly["cls"].b.data_ptr(), ly["bbox"][2].w.data_ptr(), ly["cls"].w.data_ptr()
            (D.n1_g, D.n1_b), (D.n2_g, D.n2_b), (D.n3_g, D.n3_b) = (
                tuple(t.data_ptr() for t in ly[k]) for k in ("norm1", "norm2", "norm3"))
            D.value = self.value_all.data_ptr() + lid * C * 2
            D.tgt_out, D.ref_out = self.tgt32[lid + 1].data_ptr(), self.ref[lid + 1].data_ptr()
            D.pred_box, D.pred_logit = self.pred_box[lid].data_ptr(), self.pred_logit[lid].data_ptr()
        self.dec_params = P

    def _build_decoder_program_cluster(self):
        """Per-rank weight programs of memotr_decoder_forward_cluster (include/memotr_b200.h)."""
        import copy
        import ctypes
        dev, F, CS = self.dev, self.Fd, 4
        hw = F // CS
        self.dec_packed_cl = []

        def pack(w):
            """(rows, K) bf16 -> slot images; rows padded to 64, K to 256 with zeros."""
            rows, K = w.shape
            rp, Kp = (rows + 63) // 64 * 64, (K + 255) // 256 * 256
            z = torch.zeros(rp, Kp, dtype=torch.bfloat16, device=dev)
            z[:rows, :K] = w
            v = z.reshape(rp // 64, 64, Kp // 256, 256)
            if Kp == 256:
                img = torch.zeros(rp // 64, 1, 64, 264, dtype=torch.bfloat16, device=dev)
                img[..., :256] = v.permute(0, 2, 1, 3)
            else:
                assert rp // 64 <= 4
                img = torch.zeros(Kp // 256, rp // 64, 64, 264, dtype=torch.bfloat16, device=dev)
                img[..., :256] = v.permute(2, 0, 1, 3)
            self.dec_packed_cl.append(img)
            return (img.data_ptr(), 264, rp, Kp)

        progs = []
        for r in range(CS):
            pr, rows = [], slice(64 * r, 64 * r + 64)
            for lid, ly in enumerate(self.dec):
                pr.append(pack(self.ref_point_head[0].w[rows])), pr.append(pack(self.ref_point_head[1].w[rows]))
                if lid > 0:
                    pr.append(pack(self.query_scale[0].w[rows])), pr.append(pack(self.query_scale[1].w[rows]))
                qk, ol = ly["self"]["qk"].w, ly["attn"]["ol"].w
                pr.append(pack(qk[rows])), pr.append(pack(qk[256 + 64 * r:256 + 64 * r + 64]))
                pr.append(pack(ly["self"]["v"].w[rows])), pr.append(pack(ly["self"]["out"].w[rows]))
                pr.append(pack(ol[rows])), pr.append(pack(ol[256 + 32 * r:256 + 32 * r + 32]))
                pr.append(pack(ly["attn"]["out"].w[rows]))
                pr.append(pack(ly["lin1"].w[hw * r:hw * r + hw])), pr.append(pack(ly["lin2"].w[:, hw * r:hw * r + hw]))
                pr.append(pack(ly["bbox"][0].w[rows])), pr.append(pack(ly["bbox"][1].w[rows]))
            progs.append(pr)
        n_prog = len(progs[0])
        flat = [e for pr in progs for e in pr]
        arr = (_lib.DecGemm * len(flat))(*[_lib.DecGemm(w, ldw, n, k, 0) for (w, ldw, n, k) in flat])
        self.dec_prog_cl = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        P = _lib.DecParams()
        ctypes.memmove(ctypes.byref(P), ctypes.byref(self.dec_params), ctypes.sizeof(P))
        P.prog, P.n_prog = self.dec_prog_cl.data_ptr(), n_prog
        P.init_ref_out, P.last_ref_out = self.init_ref_pts.data_ptr(), self.last_ref_pts.data_ptr()
        self.dec_params_cl = P

    def _build_decoder_program_single(self):
        """Weight program of memotr_decoder_forward (one CTA per row block, csrc/decoder_fused.cu): the variant the pipelined clip
        runs next to the encoder of the following frame."""
        import ctypes
        dev, F = self.dev, self.Fd
        prog, self.dec_packed = [], []

        def g(L, row0=0, rows=None, col0=0, K=None):
            """Pack W[row0:row0+rows, col0:col0+K] as slot images (see memotr_dec_gemm in include/memotr_b200.h)."""
            rows, K = (L.N if rows is None else rows), (L.K if K is None else K)
            assert rows % 64 == 0 and K % 256 == 0 and L.w.dtype == torch.bfloat16
            w = L.w[row0:row0 + rows, col0:col0 + K].reshape(rows // 64, 64, K // 256, 256)
            if K == 256:
                img = torch.zeros(rows // 64, 1, 64, 264, dtype=torch.bfloat16, device=dev)
                img[..., :256] = w.permute(0, 2, 1, 3)
            else:                                   # k-slice-major; the kernel keeps all four n-blocks' accumulators live
                assert rows == 256
                img = torch.zeros(K // 256, rows // 64, 64, 264, dtype=torch.bfloat16, device=dev)
                img[..., :256] = w.permute(2, 0, 1, 3)
            self.dec_packed.append(img)
            prog.append((img.data_ptr(), 264, rows, K))

        for lid, ly in enumerate(self.dec):
            g(self.ref_point_head[0]), g(self.ref_point_head[1])
            if lid > 0:
                g(self.query_scale[0]), g(self.query_scale[1])
            g(ly["self"]["qk"]), g(ly["self"]["v"]), g(ly["self"]["out"]), g(ly["attn"]["ol"]), g(ly["attn"]["out"])
            nh = 2 if F > 1024 else 1
            for half in range(nh):
                g(ly["lin1"], row0=half * F // nh, rows=F // nh)
                g(ly["lin2"], col0=half * F // nh, K=F // nh)
            g(ly["bbox"][0]), g(ly["bbox"][1])
        arr = (_lib.DecGemm * len(prog))(*[_lib.DecGemm(w, ldw, n, k, 0) for (w, ldw, n, k) in prog])
        self.dec_prog = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        self.dec_kbuf1 = torch.zeros_like(self.dec_kbuf)
        self.dec_vbuf1 = torch.zeros_like(self.dec_vbuf)
        self.dec_barrier1 = torch.zeros_like(self.dec_barrier)
        P = _lib.DecParams()
        ctypes.memmove(ctypes.byref(P), ctypes.byref(self.dec_params_cl), ctypes.sizeof(P))
        P.prog, P.n_prog = self.dec_prog.data_ptr(), len(prog)
        P.kbuf, P.vbuf, P.barrier = self.dec_kbuf1.data_ptr(), self.dec_vbuf1.data_ptr(), self.dec_barrier1.data_ptr()
        self.dec_params_single = P

    def _build_updater_program(self):
        """Per-rank weight programs + parameter block of memotr_updater_forward_cluster (include/memotr_b200.h)."""
        dev, F, C, nt, u, CS = self.dev, self.Fd, self.C, self.nt, self.upd, 4
        hw = F // CS
        self.upd_packed = []

        def pack(w):
            rows, K = w.shape
            rp, Kp = (rows + 63) // 64 * 64, (K + 255) // 256 * 256
            z = torch.zeros(rp, Kp, dtype=torch.bfloat16, device=dev)
            z[:rows, :K] = w
            v = z.reshape(rp // 64, 64, Kp // 256, 256)
            if Kp == 256:
                img = torch.zeros(rp // 64, 1, 64, 264, dtype=torch.bfloat16, device=dev)
                img[..., :256] = v.permute(0, 2, 1, 3)
            else:
                assert rp // 64 <= 4
                img = torch.zeros(Kp // 256, rp // 64, 64, 264, dtype=torch.bfloat16, device=dev)
                img[..., :256] = v.permute(2, 0, 1, 3)
            self.upd_packed.append(img)
            return (img.data_ptr(), 264, rp, Kp)

        progs = []
        for r in range(CS):
            rows = slice(64 * r, 64 * r + 64)
            ma = u["attn"]
            pr = [pack(u["conf"][0].w[rows]), pack(u["conf"][1].w[rows]),
                  pack(u["fusion"][0].w[128 * r:128 * r + 128]), pack(u["fusion"][1].w[rows]),
                  pack(u["pos_head"][0].w[rows]), pack(u["pos_head"][1].w[rows]),
                  pack(ma["q"].w[rows]), pack(ma["k"].w[rows]), pack(ma["v"].w[rows]), pack(ma["out"].w[rows])]
            for key in ("mffn", "fffn"):
                l1, l2, _ = u[key]
                pr.append(pack(l1.w[hw * r:hw * r + hw])), pr.append(pack(l2.w[:, hw * r:hw * r + hw]))
            progs.append(pr)
        flat = [e for pr in progs for e in pr]
        arr = (_lib.DecGemm * len(flat))(*[_lib.DecGemm(w, ldw, n, k, 0) for (w, ldw, n, k) in flat])
        self.upd_prog = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        self.upd_np = (nt + 63) // 64 * 64
        self.upd_kbuf = torch.zeros(self.upd_np, C, dtype=torch.float16, device=dev)
        self.upd_vbuf = torch.zeros(C, self.upd_np, dtype=torch.float16, device=dev)
        self.upd_barrier = torch.zeros(1, dtype=torch.int32, device=dev)
        P = _lib.UpdParams()
        P.prog, P.n_prog, P.nt, P.ncls, P.d_ffn, P.np = self.upd_prog.data_ptr(), len(progs[0]), nt, self.ncls, F, self.upd_np
        P.update_thresh, P.long_memory_lambda = float(self.cfg["update_thresh"]), float(self.cfg["long_memory_lambda"])
        ma = u["attn"]
        for name, L in (("conf0_b", u["conf"][0]), ("conf1_b", u["conf"][1]), ("fus0_b", u["fusion"][0]),
                        ("fus1_b", u["fusion"][1]), ("ph0_b", u["pos_head"][0]), ("ph1_b", u["pos_head"][1]),
                        ("q_b", ma["q"]), ("k_b", ma["k"]), ("v_b", ma["v"]), ("out_b", ma["out"]),
                        ("mf1_b", u["mffn"][0]), ("mf2_b", u["mffn"][1]), ("ff1_b", u["fffn"][0]), ("ff2_b", u["fffn"][1])):
            setattr(P, name, L.b.data_ptr())
        for pre, (gam, bet) in (("mn", u["memory_norm"]), ("mfn", u["mffn"][2]), ("fn", u["feat_norm"]), ("ffn", u["fffn"][2])):
            setattr(P, pre + "_g", gam.data_ptr()), setattr(P, pre + "_b", bet.data_ptr())
        P.dim_t = self.dim_t.data_ptr()
        P.track_pad = self.trk.track_pad.data_ptr() if self.trk is not None else None
        st = self.st
        P.logits, P.boxes, P.output_embed = st["logits"].data_ptr(), st["boxes"].data_ptr(), st["output_embed"].data_ptr()
        P.ref_pts, P.query_embed = st["ref_pts"].data_ptr(), st["query_embed"].data_ptr()
        P.long_memory, P.last_output = st["long_memory"].data_ptr(), st["last_output"].data_ptr()
        P.feedback_ref, P.feedback_embed = self.in_track_ref.data_ptr(), self.in_track_embed.data_ptr()
        P.kbuf, P.vbuf, P.barrier = self.upd_kbuf.data_ptr(), self.upd_vbuf.data_ptr(), self.upd_barrier.data_ptr()
        self.upd_params = P

    def _decoder_fused(self):
        import ctypes
        if self.dec_use_single:      # one CTA per row block: slower, but a third of the SM-time (pipelined clips)
            self._ck(self.lib.memotr_decoder_forward(ctypes.byref(self.dec_params_single), self._st()), "decoder_forward")
        else:
            self._ck(self.lib.memotr_decoder_forward_cluster(ctypes.byref(self.dec_params_cl), self._st()),
                     "decoder_forward_cluster")
        self.launches += 1

    # ------------------------------------------------------------------------------------------------ launch helpers
    def _st(self):
        return _lib.stream_ptr(self.dev)

    def _ck(self, rc, what):
        self.launches += 1
        _lib.check(rc, what)

    def lin(self, x, ldx, L, out, ldo, M, act=0, mul=None, ldmul=0, add=None, ldadd=0, rowzero=None, c_dtype=None,
            path=0):
        cd = self.dt if c_dtype is None else c_dtype
        if (self.tc3 and mul is None and add is None and cd == F32 and L.N % 64 == 0 and L.K % 64 == 0 and ldx % 4 == 0
                and ldo % 4 == 0 and M >= 64 and M * 3 * L.K <= self.a3.numel() and L.w.dtype == torch.float32):
            if getattr(L, "w3", None) is None:
                from .kernels import pack_w3
                L.w3 = pack_w3(L.w)
            self._ck(self.lib.memotr_linear_f32x3(_p(x), ldx, _p(L.w3), _p(L.b), _p(rowzero), _p(out), ldo, M, L.N, L.K, act,
                                                  2.0 ** -6, _p(self.a3), None, self._st()), "linear_f32x3")
            self.launches += 1
            return
        self._ck(self.lib.memotr_linear(_p(x), ldx, _p(L.w), L.K, _p(L.b), _p(mul), ldmul, _p(add), ldadd, _p(rowzero),
                                        _p(out), ldo, M, L.N, L.K, self.dt, cd, act, path, self._st()), "linear")

    def mlp2(self, x, ldx, L1, L2, out, ldo, M, hid, act2=0, mul=None, ldmul=0, c_dtype=None, force=False):
        """out = act2(relu(x L1^T + b1) L2^T + b2) [* mul].  bf16 mode with a 256-wide input/output: ONE tensor-core kernel
        with the hidden activation kept on chip (memotr_mlp2); otherwise two GEMMs through the scratch buffer `hid`."""
        cd = self.dt if c_dtype is None else c_dtype
        # one CTA per 128 rows walks the hidden chunks serially: worth it when there are enough row tiles to fill the
        # GPU (encoder, 175 tiles); on <= 400 decoder rows two GEMMs are as fast or faster (measured 51 vs 30 us for the
        # 2048-wide FFN, 22.5 vs 21.4 us for the 256-wide MLPs; profiles/r01_micro_gemm_v3_fused.json)
        worth = M >= 2048 or force
        if self.fused_mlp and worth and self.mode == "bf16" and L1.K == 256 and L2.N == 256 and L1.N % 128 == 0 \
                and L2.K == L1.N:
            self._ck(self.lib.memotr_mlp2(_p(x), ldx, _p(L1.w), _p(L1.b), _p(L2.w), _p(L2.b), _p(mul), ldmul, _p(out), ldo,
                                          M, L1.K, L1.N, L2.N, cd, act2, self._st()), "mlp2")
            return
        if (self.tc3 and mul is None and cd == F32 and act2 == 0 and L1.N % 128 == 0 and L1.K % 64 == 0 and L2.N % 64 == 0
                and (L1.N // 128) * ((M + 127) // 128) > self.n_sm > 0 and M * 3 * L1.N <= self.a3h.numel()
                and M * 3 * L1.K <= self.a3.numel() and L1.w.dtype == torch.float32 and ldx % 4 == 0 and ldo % 4 == 0):
            # fp32tc FFN: linear1 writes relu(.) straight as the split fp16 operand of linear2 (no fp32 hidden tensor in HBM)
            from .kernels import pack_w3
            for L in (L1, L2):
                if getattr(L, "w3", None) is None:
                    L.w3 = pack_w3(L.w)
            self._ck(self.lib.memotr_linear_f32x3(_p(x), ldx, _p(L1.w3), _p(L1.b), None, None, 0, M, L1.N, L1.K, 1, 2.0 ** -6,
                                                  _p(self.a3), _p(self.a3h), self._st()), "linear_f32x3(split out)")
            self._ck(self.lib.memotr_linear_f32x3(None, 0, _p(L2.w3), _p(L2.b), None, _p(out), ldo, M, L2.N, L2.K, 0, 2.0 ** -6,
                                                  _p(self.a3h), None, self._st()), "linear_f32x3(split in)")
            self.launches += 1
            return
        self.lin(x, ldx, L1, hid, L1.N, M, act=1)
        self.lin(hid, L1.N, L2, out, ldo, M, act=act2, mul=mul, ldmul=ldmul, c_dtype=cd)

    def ln(self, x, wb, y, M, x2=None, y32=None, pos=None, ypos=None):
        """y = LN(x + x2) in the activation dtype (+ fp32 master y32, + ypos = y + pos).  x, x2, y32: fp32, ld = C."""
        C = self.C
        if y32 is not None and y32.data_ptr() == y.data_ptr():
            y32 = None                                           # fp32 mode: y is its own master
        self._ck(self.lib.memotr_layernorm(_p(x), F32, C, _p(x2), C, _p(wb[0]), _p(wb[1]), 1e-5, _p(y), self.dt, C,
                                           _p(pos), C, _p(ypos), C, _p(y32), C, M, C, self._st()), "layernorm")

    def add(self, a, lda, b, ldb, out, ldo, M, N):
        self._ck(self.lib.memotr_add(_p(a), lda, _p(b), ldb, _p(out), ldo, M, N, self.dt, self._st()), "add")

    def convert(self, src, sdt, lds, dst, ddt, ldd, M, N):
        self._ck(self.lib.memotr_convert(_p(src), sdt, lds, _p(dst), ddt, ldd, M, N, self._st()), "convert")

    def msda(self, value, stride, ol, ldol, mode, ref4, out, Lq, K):
        self._ck(self.lib.memotr_msda_prep(_p(ol), ldol, _p(self.shapes_t), _p(self.lsi_t), _p(self.vr), _p(ref4), mode,
                                           _p(self.loc), _p(self.attw), Lq, self.H, self.L, K, self._st()), "msda_prep")
        timed = self.timer is not None and mode == 0
        if timed:
            slot = self._timer_slot % self.n_enc
            self._timer_slot += 1
            _lib.check(self.lib.memotr_timer_record(self.timer, 2 * slot, self._st()), "timer_record")
        self._ck(self.lib.memotr_msda_forward_ex(_p(value), stride, _p(self.shapes_t), _p(self.lsi_t), _p(self.loc),
                                                 _p(self.attw), _p(out), 1, self.S, self.H, self.L, Lq, K, self.vdt,
                                                 self._st()), "msda_forward_ex")
        if timed:
            _lib.check(self.lib.memotr_timer_record(self.timer, 2 * slot + 1, self._st()), "timer_record")

    def _encoder_ol_and_gather(self, a, K):
        """offsets/logits GEMM with the location + softmax epilogue, then the gather straight from its output rows."""
        import ctypes
        L, S, H, N = self.L, self.S, self.H, a["ol"].N
        if not hasattr(self, "_prep_hw"):
            self._prep_hw = (ctypes.c_int * (2 * L))(*[v for hw in self.shapes for v in hw])
            self._prep_lsi = (ctypes.c_int * L)(*self.lsi_host)
        self._ck(self.lib.memotr_linear_msda_prep(_p(self.q_tok), self.C, _p(a["ol"].w), a["ol"].K, _p(a["ol"].b), _p(self.ol),
                                                  N, S, self.C, H, L, K, self._prep_hw, self._prep_lsi, _p(self.vr),
                                                  self._st()), "linear_msda_prep")
        timed = self.timer is not None
        if timed:
            slot = self._timer_slot % self.n_enc
            self._timer_slot += 1
            _lib.check(self.lib.memotr_timer_record(self.timer, 2 * slot, self._st()), "timer_record")
        attw = self.ol.view(-1)[H * L * K * 2:]                    # the weights start after the locations in every row
        if self.msda_window:
            self._ck(self.lib.memotr_msda_forward_window(_p(self.value), self.C, self._prep_hw, self._prep_lsi, _p(self.ol), N,
                                                         _p(attw), N, _p(self.vr), a["win_shift"], a["win_radius"],
                                                         self.window_classes, _p(self.window_stats), _p(self.att),
                                                         S, H, L, K, self._st()), "msda_forward_window")
        else:
            self._ck(self.lib.memotr_msda_forward_strided(_p(self.value), self.C, _p(self.shapes_t), _p(self.lsi_t), _p(self.ol),
                                                          N, _p(attw), N, _p(self.att), 1, S, H, L, S, K, self._st()),
                     "msda_forward_strided")
        if timed:
            _lib.check(self.lib.memotr_timer_record(self.timer, 2 * slot + 1, self._st()), "timer_record")

    def mha(self, q, ldq, k, ldk, v, ldv, out, ldo, Nq, Nk, kpm=None):
        """q/k/v are the fp32 projections (kept in fp32 in both modes: rounding them to bf16 perturbs the attention
        logits by ~1e-2); the output is an activation (GEMM operand) in the engine dtype."""
        self._ck(self.lib.memotr_mha(_p(q), ldq, _p(k), ldk, _p(v), ldv, _p(kpm), _p(out), ldo, Nq, Nk, self.H, 32,
                                     F32, self.dt, self._st()), "mha")

    def sine(self, pts, scale4, sigmoid, out, N):
        self._ck(self.lib.memotr_sine_embed(_p(pts), 4, _p(scale4), int(sigmoid), _p(self.dim_t), _p(out), 2 * self.C, N,
                                            self.dt, self._st()), "sine_embed")

    # ------------------------------------------------------------------------------------------------ inputs
    @staticmethod
    def input_layout(shapes, C, with_pos=True):
        off, lay = 0, {"src": [], "pos": [], "mask": []}
        for key in (("src", "pos") if with_pos else ("src",)):
            for h, w in shapes:
                lay[key].append(off)
                off += C * h * w * 4
        for h, w in shapes:
            lay["mask"].append(off)
            off += h * w
        lay["bytes"] = (off + 15) // 16 * 16
        return lay

    @staticmethod
    def input_views(flat, lay, shapes, C):
        f32 = lambda o, h, w: flat[o:o + C * h * w * 4].view(torch.float32).view(C, h * w)   # noqa: E731
        src = [f32(o, h, w) for o, (h, w) in zip(lay["src"], shapes)]
        pos = [f32(o, h, w) for o, (h, w) in zip(lay["pos"], shapes)] if lay["pos"] else None
        mask = [flat[o:o + h * w] for o, (h, w) in zip(lay["mask"], shapes)]
        return src, pos, mask

    def load_frame(self, srcs, masks, pos, track_ref_pts, track_query_embed, non_blocking=True):
        """Copy one frame's inputs (host or device tensors, fp32 NCHW with batch 1) into the static input buffers."""
        for l in range(self.L):
            self.in_src[l].copy_(srcs[l].reshape(self.C, -1), non_blocking=non_blocking)
            if self.pos_cfg is None:
                self.in_pos[l].copy_(pos[l].reshape(self.C, -1), non_blocking=non_blocking)
            self.in_mask[l].copy_(masks[l].reshape(-1).to(torch.uint8), non_blocking=non_blocking)
        self.in_track_ref.copy_(track_ref_pts, non_blocking=non_blocking)
        self.in_track_embed.copy_(track_query_embed, non_blocking=non_blocking)

    def load_tracks(self, tracks, non_blocking=True):
        for k in ("query_embed", "output_embed", "last_output", "long_memory", "ref_pts", "boxes", "logits"):
            self.st[k].copy_(tracks[k], non_blocking=non_blocking)

    # ------------------------------------------------------------------------------------------------ the frame
    def forward(self):
        """Transformer + heads on the loaded frame.  Results stay in the workspace; see results()."""
        self.encode()
        self.decode()

    def encode(self, sm_plan=None):
        """Everything that depends on the frame alone (phase 1 of an exact sharded clip, memotr_b200/clip.py): level
        flattening, the encoder, and the stacked value projection of all decoder layers.
        sm_plan = (budget, n_layers): the persistent kernels of the first n_layers encoder layers size their grids for `budget`
        SMs (a concurrent kernel holds the others: capture_pipeline)."""
        C, S, H, L, dt = self.C, self.S, self.H, self.L, self.dt
        st = self._st
        if sm_plan is not None:
            _lib.check(self.lib.memotr_set_sm_budget(int(sm_plan[0])), "set_sm_budget")
        try:
            self._encode(sm_plan)
        finally:
            if sm_plan is not None:
                _lib.check(self.lib.memotr_set_sm_budget(0), "set_sm_budget")

    def _encode(self, sm_plan):
        C, S, H, L, dt = self.C, self.S, self.H, self.L, self.dt
        st = self._st
        self._mark(0)
        # -- level flattening, level embedding, valid ratios (deformable_transformer.py:196-220)
        tokens_levels = self.tokens_levels and self.pos_cfg is not None and C % 64 == 0 and self.L <= 8
        for l, (h, w) in enumerate(self.shapes):
            if self.pos_cfg is not None:      # position map of this level evaluated inside the token kernel
                if l == 0:                    # normalised cumulative counts of all levels: one launch
                    import ctypes
                    if not hasattr(self, "_pos_hw"):
                        self._pos_hw = (ctypes.c_int * (2 * self.L))(*[v for hw in self.shapes for v in hw])
                        self._pos_lsi = (ctypes.c_int * self.L)(*self.lsi_host)
                    self._ck(self.lib.memotr_pos_cumsum_levels(_p(self.in_mask[0]), self._pos_hw, self._pos_lsi, self.L,
                                                               float(self.pos_cfg.get("scale", 2 * math.pi)),
                                                               _p(self.pos_scratch), _p(self.vr), st()),
                             "pos_cumsum_levels")                      # (also writes the valid ratios of all levels)
                    if tokens_levels:         # every level's tokens in one launch (MEMOTR_TOKENS_LEVELS=0: one launch per level)
                        if not hasattr(self, "_src_ptrs"):
                            self._src_ptrs = (ctypes.c_void_p * self.L)(*[t.data_ptr() for t in self.in_src])
                        self._ck(self.lib.memotr_tokens_from_nchw_levels(
                            self._src_ptrs, _p(self.pos_scratch), _p(self.pos_dim_i), _p(self.level_embed), _p(self.src_tok),
                            _p(self.pos_tok), _p(self.q_tok), _p(None if self.mode == "fp32" else self.src32), C, self._pos_hw,
                            self._pos_lsi, self.L, C, dt, st()), "tokens_levels")
                if tokens_levels:
                    continue
                self._ck(self.lib.memotr_tokens_from_nchw_emb(
                    _p(self.in_src[l]), _p(self.pos_scratch[2 * self.lsi_host[l]:]), _p(self.pos_dim_i),
                    _p(self.level_embed[l]), _p(self.src_tok), _p(self.pos_tok), _p(self.q_tok),
                    _p(None if self.mode == "fp32" else self.src32), C, h * w, self.lsi_host[l], C, dt, st()), "tokens_emb")
            else:
                self._ck(self.lib.memotr_tokens_from_nchw(_p(self.in_src[l]), _p(self.in_pos[l]), _p(self.level_embed[l]),
                                                          _p(self.src_tok), _p(self.pos_tok), _p(self.q_tok),
                                                          _p(None if self.mode == "fp32" else self.src32), C, h * w,
                                                          self.lsi_host[l], C, dt, st()), "tokens")
            if self.pos_cfg is None:
                self._ck(self.lib.memotr_valid_ratio(_p(self.in_mask[l]), h, w, _p(self.vr[l]), st()), "valid_ratio")
        # -- encoder (deformable_encoder.py:109-131)
        self._mark(1)
        Ke = self.cfg["n_enc_points"]
        for i, ly in enumerate(self.enc):
            a = ly["attn"]
            if sm_plan is not None and i == sm_plan[1]:            # (sm_plan[1] = k: full budget from layer k on;
                _lib.check(self.lib.memotr_set_sm_budget(0), "set_sm_budget")
            if self.debug_enc is not None:
                self.debug_enc.append(self.src32.float().clone())
            self.lin(self.src_tok, C, a["value"], self.value, C, S, rowzero=self.mask_flat, c_dtype=self.vdt)
            if self.fuse_prep:
                self._encoder_ol_and_gather(a, Ke)
            else:
                self.lin(self.q_tok, C, a["ol"], self.ol, a["ol"].N, S, c_dtype=F32)
                self.msda(self.value, C, self.ol, a["ol"].N, 0, None, self.att, S, Ke)
            if sm_plan is not None and i + 0.5 == sm_plan[1]:      #  k + 0.5: from the dense half of layer k on)
                _lib.check(self.lib.memotr_set_sm_budget(0), "set_sm_budget")
            if self.fuse_block:
                (g1, b1n), (g2, b2n), l1, l2 = ly["norm1"], ly["norm2"], ly["lin1"], ly["lin2"]
                self._ck(self.lib.memotr_encoder_dense_block(
                    _p(self.att), C, _p(a["out"].w), _p(a["out"].b), _p(self.src32), C, _p(g1), _p(b1n), _p(self.src1_32), C,
                    _p(l1.w), _p(l1.b), _p(l2.w), _p(l2.b), _p(g2), _p(b2n), _p(self.pos_tok), C, _p(self.src_tok), C,
                    _p(self.src32), C, _p(self.q_tok), C, _p(self.pre), C, S, self.Fd, 1e-5, st()), "encoder_dense_block")
                continue
            if self.fuse_outln:
                g1, b1n = ly["norm1"]
                self._ck(self.lib.memotr_linear256_layernorm(_p(self.att), C, _p(a["out"].w), _p(a["out"].b), _p(self.src32), C,
                                                             _p(g1), _p(b1n), 1e-5, _p(self.src1), C, _p(self.src1_32), C, S, st()),
                         "linear256_layernorm")
            else:
                self.lin(self.att, C, a["out"], self.pre, C, S, c_dtype=F32)
                self.ln(self.pre, ly["norm1"], self.src1, S, x2=self.src32, y32=self.src1_32)
            tf = self.timer is not None and getattr(self, "time_ffn", False)   # (event nodes cost PDL overlap: opt-in)
            if tf:
                _lib.check(self.lib.memotr_timer_record(self.timer, 2 * self.n_enc + 5 + 2 * i, st()), "timer_record")
            if self.fuse_ln2:
                # first round of row tiles (one CTA per SM): FFN + norm2 in one kernel; the remaining rows: FFN with the hidden
                # dimension split over the idle SMs, then the stand-alone LayerNorm on those rows only
                g2, b2 = ly["norm2"]
                tiles = (S + 127) // 128
                main = S if tiles <= self.n_sm else self.n_sm * 128
                l1, l2 = ly["lin1"], ly["lin2"]
                self._ck(self.lib.memotr_mlp2_lnout(_p(self.src1), C, _p(l1.w), _p(l1.b), _p(l2.w), _p(l2.b), _p(self.src1_32), C,
                                                    _p(g2), _p(b2), 1e-5, _p(self.src_tok), C, _p(self.src32), C, _p(self.pos_tok), C,
                                                    _p(self.q_tok), C, main, self.Fd, st()), "mlp2_lnout")
                if main < S:
                    self.mlp2(self.src1[main:], C, l1, l2, self.pre[main:], C, S - main, self.hid, c_dtype=F32, force=True)
                    self.ln(self.pre[main:], ly["norm2"], self.src_tok[main:], S - main, x2=self.src1_32[main:],
                            y32=self.src32[main:], pos=self.pos_tok[main:], ypos=self.q_tok[main:])
            else:
                self.mlp2(self.src1, C, ly["lin1"], ly["lin2"], self.pre, C, S, self.hid, c_dtype=F32)
                self.ln(self.pre, ly["norm2"], self.src_tok, S, x2=self.src1_32, y32=self.src32, pos=self.pos_tok,
                        ypos=self.q_tok)
            if tf:
                _lib.check(self.lib.memotr_timer_record(self.timer, 2 * self.n_enc + 6 + 2 * i, st()), "timer_record")
        memory = self.src_tok
        if self.debug_enc is not None:
            self.debug_enc.append(self.src32.float().clone())
        self._mark(2)
        # value maps of all decoder layers in one GEMM over the memory (ms_deform_attn.py:104-106, x6)
        self.lin(memory, C, self.dec_value, self.value_all, self.n_dec * C, S, rowzero=self.mask_flat, c_dtype=self.vdt)

    def decode(self):
        """The recurrent part of a frame: the decoder and heads on the CURRENT track queries (in_track_ref / in_track_embed)
        and the value maps / valid ratios encode() left in the workspace."""
        C, S, H, L, dt = self.C, self.S, self.H, self.L, self.dt
        st = self._st
        # -- decoder inputs (memotr.py:209-278, deformable_transformer.py:239-242)
        nd, nt, nq = self.nd, self.nt, self.nq
        if not self.dec_cluster or not getattr(self, "_det_rows_ready", False):
            # the detect-query rows are constants: with the fused decoder (which never touches these buffers' detect rows)
            # they are written once -- capture() always runs one eager warm-up step first
            self.convert(self.det_anchor, F32, 4, self.ref_raw, F32, 4, nd, 4)
            self.convert(self.det_query_embed, F32, C, self.tgt32[0], F32, C, nd, C)
            self._det_rows_ready = True
        self.convert(self.in_track_ref, F32, 4, self.ref_raw[nd:], F32, 4, nt, 4)
        self._ck(self.lib.memotr_unary(_p(self.ref_raw), _p(self.ref[0]), nq * 4, 0, st()), "sigmoid")
        self.convert(self.in_track_embed, F32, C, self.tgt32[0][nd:], F32, C, nt, C)
        if not self.dec_fused:
            if self.mode != "fp32":
                self.convert(self.tgt32[0], F32, C, self.tgt[0], dt, C, nq, C)
            self.convert(self.vr, F32, 2, self.vr_scale4, F32, 2, 1, 2)          # (vr0.w, vr0.h, vr0.w, vr0.h)
            self.convert(self.vr, F32, 2, self.vr_scale4[2:], F32, 2, 1, 2)
        Kd = self.cfg["n_dec_points"]
        if self.dec_fused:
            self._decoder_fused()
        for lid, ly in enumerate(() if self.dec_fused else self.dec):
            out, ref = self.tgt[lid], self.ref[lid]
            n = nq if lid >= self.merge else nd              # det/track split before the merge layer (:292-297)
            # DAB positional query (deformable_decoder.py:88-95)
            self.sine(ref, self.vr_scale4, False, self.anchor, nq)
            self.lin(self.anchor, 2 * C, self.ref_point_head[0], self.d_a, C, nq, act=1)
            if lid == 0:
                self.lin(self.d_a, C, self.ref_point_head[1], self.query_pos, C, nq)
            else:
                self.lin(self.d_a, C, self.ref_point_head[1], self.d_b, C, nq)
                self.mlp2(out, C, self.query_scale[0], self.query_scale[1], self.query_pos, C, nq, self.d_c,
                          mul=self.d_b, ldmul=C)
            # self-attention (deformable_decoder.py:245-252)
            sa = ly["self"]
            self.add(out, C, self.query_pos, C, self.qk_in, C, n, C)
            self.lin(self.qk_in, C, sa["qk"], self.qk, 2 * C, n, c_dtype=F32)
            self.lin(out, C, sa["v"], self.v, C, n, c_dtype=F32)
            self.mha(self.qk, 2 * C, self.qk[:, C:], 2 * C, self.v, C, self.d_a, C, n, n,
                     kpm=self.query_pad if (self.use_pad and n == nq) else None)
            self.lin(self.d_a, C, sa["out"], self.d_pre, C, n, c_dtype=F32)
            self.ln(self.d_pre, ly["norm2"], self.t1, n, x2=self.tgt32[lid], y32=self.t1_32, pos=self.query_pos,
                    ypos=self.t1q)
            # cross-attention into the encoder memory (deformable_decoder.py:303-313)
            a = ly["attn"]
            self.lin(self.t1q, C, a["ol"], self.ol, a["ol"].N, n, c_dtype=F32)
            self.msda(self.value_all[:, lid * C:], self.n_dec * C, self.ol, a["ol"].N, 1, ref, self.d_a, n, Kd)
            self.lin(self.d_a, C, a["out"], self.d_pre, C, n, c_dtype=F32)
            self.ln(self.d_pre, ly["norm1"], self.t2, n, x2=self.t1_32, y32=self.t2_32)
            # FFN (deformable_decoder.py:263-273)
            self.mlp2(self.t2, C, ly["lin1"], ly["lin2"], self.d_pre, C, n, self.d_hid, c_dtype=F32)
            new, new32 = self.tgt[lid + 1], self.tgt32[lid + 1]
            self.ln(self.d_pre, ly["norm3"], new, n, x2=self.t2_32, y32=new32)
            if n < nq:                                       # track queries bypass the layer (:316-317)
                self.convert(self.tgt32[lid][n:], F32, C, new32[n:], F32, C, nq - n, C)
                if self.mode != "fp32":
                    self.convert(out[n:], dt, C, new[n:], dt, C, nq - n, C)
            # box refinement + heads (deformable_decoder.py:139-159, memotr.py:147-162)
            bb = ly["bbox"]
            self.mlp2(new, C, bb[0], bb[1], self.d_c, C, nq, self.d_a, act2=1)
            self.lin(self.d_c, C, bb[2], self.delta, 4, nq, c_dtype=F32)
            self._ck(self.lib.memotr_box_refine(_p(self.delta), _p(ref), _p(self.pred_box[lid]), _p(self.ref[lid + 1]),
                                                nq, nq if lid >= self.merge else nd, st()), "box_refine")
            self.lin(new, C, ly["cls"], self.pred_logit[lid], self.ncls, nq, c_dtype=F32)
        if not self.dec_cluster:        # (the cluster decoder kernel writes both itself)
            self._ck(self.lib.memotr_unary(_p(self.ref[self.n_dec - 1]), _p(self.last_ref_pts), nq * 4, 1, st()), "inv_sig")
            self._ck(self.lib.memotr_unary(_p(self.ref[0]), _p(self.init_ref_pts), nq * 4, 1, st()), "inv_sig")
        self._mark(3)

    def convert_u8(self, src, dst, n):
        dst[:n].copy_(src)       # device-to-device byte copy on the current stream (cudaMemcpyAsync; graph-capturable)

    def results(self):
        """The reference's output dict (memotr.py:180-195), fp32, batch dimension restored."""
        n = self.n_dec
        f = lambda t: t.float()[None]                                                  # noqa: E731
        return {
            "pred_logits": self.pred_logit[n - 1][None], "pred_bboxes": self.pred_box[n - 1][None],
            "last_ref_pts": self.last_ref_pts[None], "init_ref_pts": self.init_ref_pts[None],
            "outputs": self.tgt32[n][None],
            "aux_logits": torch.stack(self.pred_logit[:-1])[:, None], "aux_bboxes": torch.stack(self.pred_box[:-1])[:, None],
            "aux_queries": torch.stack(self.tgt32[1:n])[:, None],
            "memory": f(self.src_tok),
        }

    # ------------------------------------------------------------------------------------------------ query updater
    def tracks_from_frame(self):
        """What RuntimeTracker.update writes into the track instances before the updater runs
        (runtime_tracker.py:43-45): boxes, logits and output_embed of the track rows of this frame."""
        nd, nt, C, n = self.nd, self.nt, self.C, self.n_dec
        self.convert(self.pred_box[n - 1][nd:], F32, 4, self.st["boxes"], F32, 4, nt, 4)
        self.convert(self.pred_logit[n - 1][nd:], F32, self.ncls, self.st["logits"], F32, self.ncls, nt, self.ncls)
        self.convert(self.tgt32[n][nd:], F32, C, self.st["output_embed"], F32, C, nt, C)

    def update_tracks(self):
        """QueryUpdater.update_tracks_embedding on the fp32 track state in self.st (query_updater.py:82-166)."""
        C, nt, dt, st, u = self.C, self.nt, self.dt, self.st, self.upd
        if nt == 0:
            return
        if self.upd_fused:      # one persistent kernel; it also writes the fed-back track queries (in_track_ref / _embed)
            import ctypes
            self._ck(self.lib.memotr_updater_forward_cluster(ctypes.byref(self.upd_params), self._st()), "updater_forward")
            self.launches += 1
            return
        self._ck(self.lib.memotr_upd_prepare(_p(st["logits"]), self.ncls, _p(st["boxes"]), _p(st["ref_pts"]),
                                             float(self.cfg["update_thresh"]), _p(self.is_pos), _p(self.u_ref), nt,
                                             self._st()), "upd_prepare")
        self.sine(self.u_ref, None, True, self.u_sine, nt)
        self.convert(st["output_embed"], F32, C, self.u_oe, dt, C, nt, C)
        self.convert(st["last_output"], F32, C, self.u_cat[:, C:], dt, 2 * C, nt, C)
        self.convert(st["long_memory"], F32, C, self.u_long, dt, C, nt, C)
        # confidence gate and short-memory fusion (:109-118)
        self.mlp2(self.u_oe, C, u["conf"][0], u["conf"][1], self.u_cat, 2 * C, nt, self.u_a, act2=2, mul=self.u_oe,
                  ldmul=C)
        self.lin(self.u_cat, 2 * C, u["fusion"][0], self.u_big, 2 * C, nt, act=1)
        self.lin(self.u_big, 2 * C, u["fusion"][1], self.u_a, C, nt)                 # short memory
        self.lin(self.u_sine, 2 * C, u["pos_head"][0], self.u_b, C, nt, act=1)
        self.lin(self.u_b, C, u["pos_head"][1], self.u_c, C, nt)                     # query_pos
        self.add(self.u_a, C, self.u_c, C, self.u_b, C, nt, C)                        # q = short + pos
        self.add(self.u_long, C, self.u_c, C, self.u_d, C, nt, C)                     # k = long + pos
        # long-term-memory attention (:125-128)
        ma = u["attn"]
        self.lin(self.u_b, C, ma["q"], self.u_q, C, nt, c_dtype=F32)
        self.lin(self.u_d, C, ma["k"], self.u_k, C, nt, c_dtype=F32)
        self.lin(self.u_oe, C, ma["v"], self.u_v, C, nt, c_dtype=F32)
        self.mha(self.u_q, C, self.u_k, C, self.u_v, C, self.u_a, C, nt, nt,
                 kpm=self.trk.track_pad if self.trk is not None else None)
        self.lin(self.u_a, C, ma["out"], self.u_pre, C, nt, c_dtype=F32)
        self.ln(self.u_pre, u["memory_norm"], self.u_a, nt, x2=st["output_embed"], y32=self.u_a32)
        l1, l2, nrm = u["mffn"]
        self.mlp2(self.u_a, C, l1, l2, self.u_pre, C, nt, self.u_hid, c_dtype=F32)
        self.ln(self.u_pre, nrm, self.u_c, nt, x2=self.u_a32, y32=self.u_c32)
        # long-memory residual branch (:130-133)
        self.ln(self.u_c32, u["feat_norm"], self.u_a, nt, x2=st["long_memory"], y32=self.u_a32)
        l1, l2, nrm = u["fffn"]
        self.mlp2(self.u_a, C, l1, l2, self.u_pre, C, nt, self.u_hid, c_dtype=F32)
        self.ln(self.u_pre, nrm, self.u_c, nt, x2=self.u_a32, y32=self.u_c32)        # query_feat
        # masked state writes (:135-147); ref_pts was already replaced where is_pos (:99-102)
        self._ck(self.lib.memotr_upd_finalize(_p(self.is_pos), _p(self.u_c32), F32, C, _p(st["output_embed"]),
                                              _p(st["query_embed"]), _p(st["long_memory"]), _p(st["last_output"]),
                                              float(self.cfg["long_memory_lambda"]), nt, C, self._st()), "upd_finalize")
        self.convert(self.u_ref, F32, 4, st["ref_pts"], F32, 4, nt, 4)

    def track_state(self):
        return {k: v.clone() for k, v in self.st.items()}

    # ------------------------------------------------------------------------------------------------ whole step
    def step(self):
        """One hot-path step: frame forward, hand the track rows to the updater, update the track embeddings, and
        feed the updated (ref_pts, query_embed) back as the next frame's track queries (submit_engine.py:64-72 with
        the host-side RuntimeTracker glue reduced to the field hand-off of runtime_tracker.py:43-45)."""
        self.encode()
        self.step_tail()

    # -- exact sharded clips (memotr_b200/clip.py:run_clip_two_phase): phase-1 token and the track memory hand-off ------
    def frame_token(self):
        """What decode() needs from encode(), detached from the workspace: the decoder value maps and the valid ratios."""
        return self.value_all.clone(), self.vr.clone()

    def load_token(self, token):
        self.value_all.copy_(token[0], non_blocking=True)
        self.vr.copy_(token[1], non_blocking=True)

    def get_track_memory(self):
        """The complete recurrent state as one packed uint8 tensor (clip.pack_track_state): every TrackInstances field of the
        table, the live-row count and the tracker's id counter."""
        from . import clip
        if self.trk is None:
            return clip.pack_track_state(self.st)
        return clip.pack_track_state({k: self.table[k] for k in clip.FLOAT_FIELDS + clip.INT_FIELDS},
                                     self.table.n_active, self.trk.max_obj_id)

    def set_track_memory(self, packed):
        """Adopt a packed track memory (from another rank): table, padding mask, id counter and the fed-back track queries."""
        from . import clip
        m = clip.unpack_track_state(packed, self.nt, self.C, self.ncls)
        if self.trk is None:
            for k in clip.FLOAT_FIELDS:
                self.st[k].copy_(m[k])
        else:
            n = int(m["n_active"].item())
            self.trk.reset({k: m[k][:n] for k in clip.FLOAT_FIELDS + clip.INT_FIELDS}, max_obj_id=int(m["max_obj_id"].item()))
        self.in_track_ref.copy_(self.st["ref_pts"])
        self.in_track_embed.copy_(self.st["query_embed"])

    def run_clip_two_phase(self, n_frames, feed, group=None):
        """The exact sharded clip (clip.run_clip_two_phase) on this engine: `feed(i)` puts frame i into the input buffers;
        phase 1 = encode() per frame of this rank, phase 2 = the recurrent tail on the track memory handed along the
        ranks.  Eager launches (no graph).  Returns the rank's frame indices."""
        from . import clip

        def encode(i):
            feed(i)
            self.encode()
            return self.frame_token()

        def decode(i, token):
            self.load_token(token)
            self.step_tail()

        out = clip.run_clip_two_phase(n_frames, encode, decode, self.get_track_memory, self.set_track_memory, group=group)
        return [i for i, _ in out]

    def step_tail(self):
        """The recurrent tail of a step: decoder + heads, tracker glue, query updater, feedback."""
        self.decode()
        if self.trk is None:
            self.tracks_from_frame()
        else:       # RuntimeTracker.update + select_active_tracks on the device (submit_engine.py:66-70)
            n = self.n_dec
            self.trk.update(self.pred_logit[n - 1], self.pred_box[n - 1], self.tgt32[n], self.last_ref_pts,
                            self.tgt32[n - 1])
            self.launches += 3
        self.update_tracks()
        if self.trk is not None:
            self.trk.results(*self.ori_size)
            self.launches += 1
        if not self.upd_fused:      # (the fused updater kernel writes the fed-back track queries itself)
            self.convert(self.st["ref_pts"], F32, 4, self.in_track_ref, F32, 4, self.nt, 4)
            self.convert(self.st["query_embed"], F32, self.C, self.in_track_embed, F32, self.C, self.nt, self.C)
        self._mark(4)

    def _recurrent_state(self):
        """The tensors a step mutates and the next step reads: track table / tracker bookkeeping and the fed-back queries."""
        ts = [self.in_track_ref, self.in_track_embed, self.query_pad]
        ts += self.trk.state_tensors() if self.trk is not None else list(self.st.values())
        return ts

    def capture(self, fn=None):
        """Record `fn` (default: step) into a CUDA graph; replay() then re-issues the whole frame with one launch.
        The eager warm-up run that precedes the capture leaves the recurrent state (track table, identities, padding
        mask, fed-back track queries) exactly as it found it."""
        fn = fn or self.step
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            state = self._recurrent_state()
            saved = [t.clone() for t in state]
            fn()                                            # warm-up outside capture (one-time attribute calls etc.)
            for t, v in zip(state, saved):
                t.copy_(v)
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self.launches = 0
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        self.graph, self.graph_launches = g, self.launches
        return g

    def replay(self):
        self.graph.replay()

    # ------------------------------------------------------------------------------------------------ frame pipelining
    # The recurrent tail of a frame (decoder + heads, tracker glue, query updater) is a latency chain on ~100 of the 148 SMs
    # (profiles/r02_decoder_cluster_ncu.md: issue slots 15 %), and the encoder of the NEXT frame depends on nothing it
    # produces.  For a clip whose frames are at hand (the submit loop reads a video) the two run concurrently: tail(k) on
    # the capture stream, encode(k + 1) on a second stream, one CUDA graph with a fork and a join per frame.  encode() hands
    # decode() two things -- the stacked decoder value maps and the valid ratios -- so those exist twice (`use_set`), and
    # there is one graph per parity.  Same kernels, same per-frame order, same results as step() (tests/test_engine_gpu.py).
    def enable_pipeline(self):
        import ctypes
        if getattr(self, "_sets", None) is not None:
            return
        if not self.dec_cluster:
            raise RuntimeError("FrameEngine.enable_pipeline: needs the fused decoder (bf16 mode)")
        C = self.C
        single = self.dec_params_single if self.dec_single_ok else None
        sets = [dict(value_all=self.value_all, vr=self.vr, dec=self.dec_params_cl, dec1=single)]
        v2, vr2 = torch.empty_like(self.value_all), torch.empty_like(self.vr)

        def second(src):
            if src is None:
                return None
            P = _lib.DecParams()
            ctypes.memmove(ctypes.byref(P), ctypes.byref(src), ctypes.sizeof(P))
            P.valid_ratios = vr2.data_ptr()
            for lid in range(self.n_dec):
                P.layers[lid].value = v2.data_ptr() + lid * C * 2
            return P
        sets.append(dict(value_all=v2, vr=vr2, dec=second(self.dec_params_cl), dec1=second(single)))
        self._sets = sets

    def use_set(self, i):
        """Which copy of (value_all, vr) the following encode() writes / decode() reads (enqueue-time pointers)."""
        d = self._sets[i]
        self.value_all, self.vr, self.dec_params_cl, self.dec_params_single = d["value_all"], d["vr"], d["dec"], d["dec1"]

    def capture_pipeline(self, single=None, split=None):
        """Five graphs: encode-only into set 0 (clip prologue), [tail(set c) || encode(set 1 - c)] for c = 0, 1, tail-only for
        c = 0, 1 (clip epilogue).  Call after capture() (which did the one-time warm-up of every kernel).
        In the forked graphs the decoder is the one-CTA-per-row-block kernel (`single`, default on when available): ~600 us on 25
        SMs instead of ~380 us on 100 -- its latency hides behind the encoder, its SM-time is what the encoder loses -- and the
        first `split` encoder layers (default: two thirds) size their persistent grids for the SMs that leaves free
        (memotr_set_sm_budget); by the time the later layers launch, the tail is done.  Measured at the DanceTrack size
        (frames/s): cluster decoder 666; single-CTA decoder with split 0 / 2 / 3 / 4 / 5 / 6: 652 / 681 / 714 / 726 / 717 / 702
        (25 SMs reserved); k + 0.5 = budget until after the gather of layer k; 4.5 with 28 reserved: 733 (the plateau is flat).
        MEMOTR_PIPE_SINGLE / MEMOTR_PIPE_SPLIT / MEMOTR_PIPE_RESERVE override."""
        import os
        if self.graph is None:
            raise RuntimeError("FrameEngine.capture_pipeline: call capture() first")
        self.enable_pipeline()
        if single is None:
            single = os.environ.get("MEMOTR_PIPE_SINGLE", "1") != "0"
        single = bool(single) and self.dec_single_ok
        if split is None:
            split = float(os.environ.get("MEMOTR_PIPE_SPLIT", str((2 * self.n_enc + 2) // 3 + 0.5)))
        # SMs left to the tail: the decoder's CTAs (one per 16 query rows) or the updater's (a 4-CTA cluster per 16 track rows)
        reserve = int(os.environ.get("MEMOTR_PIPE_RESERVE", str(max((self.nq + 15) // 16, 4 * ((self.nt + 15) // 16)))))
        budget = self.n_sm - reserve if single else 0
        if single:           # one-time attribute calls of the kernel outside capture (a warm-up launch; state restored)
            state = self._recurrent_state()
            saved = [t.clone() for t in state]
            self.dec_use_single = True
            self.step_tail()
            self.dec_use_single = False
            for t, v in zip(state, saved):
                t.copy_(v)
            torch.cuda.synchronize(self.dev)
        timer, self.timer = self.timer, None                   # (event nodes only in the sequential instrumented graph)
        main = torch.cuda.Stream(self.dev, priority=-1)        # the tail: fewer, longer CTAs -- scheduled first
        side = torch.cuda.Stream(self.dev)
        torch.cuda.synchronize(self.dev)

        def graph_of(body):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=main):
                body()
            return g

        def pair(c):
            def body():
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    self.use_set(1 - c)
                    self.encode(sm_plan=(budget, split) if budget > 0 else None)
                self.use_set(c)
                self.dec_use_single = single
                try:
                    self.step_tail()
                finally:
                    self.dec_use_single = False
                main.wait_stream(side)
            return body

        def only(fn, c):
            def body():
                self.use_set(c)
                fn()
            return body
        try:
            self.g_first = graph_of(only(self.encode, 0))
            self.g_pipe = [graph_of(pair(0)), graph_of(pair(1))]
            self.g_last = [graph_of(only(self.step_tail, 0)), graph_of(only(self.step_tail, 1))]
        finally:
            self.use_set(0)
            self.timer = timer

    def run_clip_pipelined(self, n_frames, feed):
        """Frames 0 .. n_frames - 1 of a clip with frame pipelining; feed(j) loads frame j's inputs into the static input buffers
        (stream-ordered, e.g. device-to-device copies).  After it returns the recurrent state is that of the sequential clip."""
        if n_frames <= 0:
            return
        feed(0)
        self.g_first.replay()
        for j in range(n_frames - 1):
            feed(j + 1)                                        # (encode(j) has finished with the input buffers: graph order)
            self.g_pipe[j & 1].replay()
        self.g_last[(n_frames - 1) & 1].replay()

    def run_clips_pipelined(self, frames_per_clip, feed, between=None):
        """A STREAM of clips as one pipeline: like run_clip_pipelined per clip, except that the tail of a clip's last frame runs
        concurrently with the encoder of the NEXT clip's first frame (which depends on no track state), so only the stream's very
        first encoder and very last tail run alone -- what a rank with 8-frame sub-clips pays per clip otherwise is 4 % of it.
        frames_per_clip: list of frame counts; feed(c, j) loads frame j of clip c; between(c) is called when every frame of clip c
        has been enqueued (its track state is final in stream order): exchange it, then reset the tracks for clip c + 1 there."""
        clips = [(c, n) for c, n in enumerate(frames_per_clip) if n > 0]
        if not clips:
            return
        p = 0                                                   # the set that holds the encoded, not yet decoded frame
        feed(clips[0][0], 0)
        self.g_first.replay()
        for i, (c, n) in enumerate(clips):
            for j in range(n):
                if j + 1 < n:
                    nxt = (c, j + 1)
                else:
                    nxt = (clips[i + 1][0], 0) if i + 1 < len(clips) else None
                if nxt is None:
                    self.g_last[p].replay()
                else:
                    feed(*nxt)
                    self.g_pipe[p].replay()                     # tail(c, j) from set p || encode(nxt) into set 1 - p
                    p ^= 1
            if between is not None:
                between(c)


class ClipRunner:
    """Public host-buffer API for a clip: frames arrive as pinned HOST tensors, results go back to pinned host tensors.

    The host->device copy of frame i+1 runs on a dedicated copy stream into one of two device staging slots while the
    captured step of frame i runs on the compute stream (events order slot reuse), so the PCIe transfer (45.7 MB per
    1333x800 frame) overlaps the compute instead of preceding it.  One replayed CUDA graph per frame; results of frame i
    are read back asynchronously into pinned host buffers.
    """

    def __init__(self, eng: "FrameEngine"):
        self.eng = eng
        dev = eng.dev
        if eng.graph is None:
            eng.capture()
        self.copy_stream = torch.cuda.Stream(dev)
        lay = eng.in_layout
        self.stage = []
        for _ in range(2):          # two staging slots with the engine's own input layout: one D2D copy hands a frame over
            flat = torch.zeros(lay["bytes"], dtype=torch.uint8, device=dev)
            src, pos, mask = eng.input_views(flat, lay, eng.shapes, eng.C)
            self.stage.append({"flat": flat, "src": src, "pos": pos, "mask": mask})
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [torch.cuda.Event(), torch.cuda.Event()]
        for e in self.free:
            e.record(torch.cuda.current_stream(dev))
        self.raw_outputs = eng.trk is None
        self.host_out = {}
        if self.raw_outputs:     # the caller keeps the tracks: raw frame outputs + the updated track queries
            self.host_out = {"pred_logits": torch.empty(eng.nq, eng.ncls).pin_memory(),
                             "pred_bboxes": torch.empty(eng.nq, 4).pin_memory(),
                             "track_query_embed": torch.empty(eng.nt, eng.C).pin_memory(),
                             "track_ref_pts": torch.empty(eng.nt, 4).pin_memory()}
        else:                    # tracks live on the device: only the frame's result rows travel (submit_engine.py:99-102)
            self.host_out = {"results": torch.empty(eng.trk.res_flat.numel(), dtype=torch.uint8).pin_memory(),
                             "n_active": torch.empty(1, dtype=torch.int32).pin_memory()}
        n_in = ("src", "pos", "mask") if eng.pos_cfg is None else ("src", "mask")
        self.h2d_bytes = sum(t.numel() * t.element_size() for k in n_in for t in self.stage[0][k])
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in self.host_out.values())

    def prefetch(self, slot, srcs, pos, masks):
        """Enqueue the H2D copy of one frame (pinned host tensors: per level (1,C,H,W) fp32 and (1,H,W) uint8/bool; `pos`
        is None when the engine rebuilds the position maps on the device)."""
        eng, st = self.eng, self.stage[slot]
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[slot])
            for l in range(eng.L):
                st["src"][l].copy_(srcs[l].reshape(eng.C, -1), non_blocking=True)
                if eng.pos_cfg is None:
                    st["pos"][l].copy_(pos[l].reshape(eng.C, -1), non_blocking=True)
                st["mask"][l].copy_(masks[l].reshape(-1), non_blocking=True)
            self.ready[slot].record(self.copy_stream)

    def run(self, slot):
        """Run the hot-path step on the frame staged in `slot`; enqueue the read-back of its results."""
        eng, st = self.eng, self.stage[slot]
        cur = torch.cuda.current_stream(eng.dev)
        cur.wait_event(self.ready[slot])
        eng.in_flat.copy_(st["flat"], non_blocking=True)
        self.free[slot].record(cur)
        eng.replay()
        if self.raw_outputs:
            n = eng.n_dec
            self.host_out["pred_logits"].copy_(eng.pred_logit[n - 1], non_blocking=True)
            self.host_out["pred_bboxes"].copy_(eng.pred_box[n - 1], non_blocking=True)
            self.host_out["track_query_embed"].copy_(eng.st["query_embed"], non_blocking=True)
            self.host_out["track_ref_pts"].copy_(eng.st["ref_pts"], non_blocking=True)
        else:
            self.host_out["results"].copy_(eng.trk.res_flat, non_blocking=True)
            self.host_out["n_active"].copy_(eng.table.n_active, non_blocking=True)
        return self.host_out

    def check(self):
        """Raise if the device track table overflowed at any point since the last reset (newborn tracks dropped:
        identities would diverge from the reference).  Synchronises; call it at the end of a clip."""
        if self.eng.trk is not None:
            self.eng.trk.check_overflow()

    def results(self):
        """Tracker mode, after a synchronize: the frame's result rows (ids, xyxy boxes in pixels, scores) of the kept tracks."""
        ids, boxes, scores, keep = self.eng.trk.split_results(self.host_out["results"])
        k = keep.bool()
        return ids[k], boxes[k], scores[k]

    def run_clip_pipelined(self, frames, sync=True):
        """run_clip with frame pipelining (FrameEngine.run_clip_pipelined): while the recurrent tail of frame k runs, the
        encoder of frame k + 1 runs on a second stream and frame k + 2 crosses PCIe.  frames: list of (srcs, pos, masks) pinned
        host tensors.  Tracker mode returns the per-frame result buffers, one row per frame (pinned; `frame_results(i)`
        decodes row i); raw mode returns the last frame's outputs.  Same results as run_clip."""
        eng = self.eng
        frames = list(frames)
        n = len(frames)
        if n == 0:
            return None
        if getattr(eng, "g_pipe", None) is None:
            eng.capture_pipeline()
        if not self.raw_outputs and (getattr(self, "clip_results", None) is None or self.clip_results.shape[0] < n):
            self.clip_results = torch.empty(n, eng.trk.res_flat.numel(), dtype=torch.uint8).pin_memory()
            self.clip_n_active = torch.empty(n, dtype=torch.int32).pin_memory()
        cur = torch.cuda.current_stream(eng.dev)

        def hand_over(j):                      # staged frame j -> the static input buffers (one device-to-device copy)
            slot = j % 2
            cur.wait_event(self.ready[slot])
            eng.in_flat.copy_(self.stage[slot]["flat"], non_blocking=True)
            self.free[slot].record(cur)

        def read_back(j):
            if self.raw_outputs:
                k = eng.n_dec
                self.host_out["pred_logits"].copy_(eng.pred_logit[k - 1], non_blocking=True)
                self.host_out["pred_bboxes"].copy_(eng.pred_box[k - 1], non_blocking=True)
                self.host_out["track_query_embed"].copy_(eng.st["query_embed"], non_blocking=True)
                self.host_out["track_ref_pts"].copy_(eng.st["ref_pts"], non_blocking=True)
            else:
                self.clip_results[j].copy_(eng.trk.res_flat, non_blocking=True)
                self.clip_n_active[j:j + 1].copy_(eng.table.n_active, non_blocking=True)

        self.prefetch(0, *frames[0])
        if n > 1:
            self.prefetch(1, *frames[1])
        hand_over(0)
        eng.g_first.replay()
        for j in range(n - 1):
            hand_over(j + 1)                   # (encode(j) is done with the input buffers: graph order on this stream)
            if j + 2 < n:
                self.prefetch(j % 2, *frames[j + 2])
            eng.g_pipe[j & 1].replay()
            read_back(j)
        eng.g_last[(n - 1) & 1].replay()
        read_back(n - 1)
        if sync:
            cur.synchronize()
            self.check()
        return self.host_out if self.raw_outputs else self.clip_results[:n]

    def frame_results(self, i):
        """After run_clip_pipelined (tracker mode, synchronised): frame i's result rows (ids, xyxy boxes in pixels, scores)."""
        ids, boxes, scores, keep = self.eng.trk.split_results(self.clip_results[i])
        k = keep.bool()
        return ids[k], boxes[k], scores[k]

    def run_clip(self, frames):
        """frames: iterable of (srcs, pos, masks) pinned host tensors.  Returns the last frame's host outputs."""
        frames = list(frames)
        if not frames:
            return None
        self.prefetch(0, *frames[0])
        out = None
        for i in range(len(frames)):
            if i + 1 < len(frames):
                self.prefetch((i + 1) % 2, *frames[i + 1])
            out = self.run(i % 2)
        torch.cuda.current_stream(self.eng.dev).synchronize()
        self.check()
        return out
