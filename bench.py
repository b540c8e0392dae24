#!/usr/bin/env python
"""bench.py -- frames/sec of the MeMOTR per-frame hot path on B200 (contract: see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--mode bf16|fp32] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one frame through the hot path: level flattening -> 6-layer deformable encoder -> 6-layer decoder (300
detect + 100 track queries) -> class/box heads -> QueryUpdater.update_tracks_embedding, on a synthetic 1333x800
4-scale pyramid (S = 22323 tokens), DanceTrack hyper-parameters (BASELINE.json configs[1] minus the ResNet-50
backbone, which SURVEY.md section 8 marks out of scope).  The track queries a step produces feed the next step, so K
steps are a K-frame clip.

  value   frames/s with the frame inputs already resident in HBM (CUDA-graph replay of the whole step; 4 distinct
          frames rotate through the input buffers, device-to-device, inside the timed region).
  e2e     the same metric through the public API (memotr_b200.engine.ClipRunner) with HOST buffers: every step copies
          the frame (4 feature maps, 4 position maps, 4 masks) from pinned host memory -- on a copy stream, double
          buffered, so the transfer of frame i+1 overlaps the compute of frame i -- runs the step and reads pred_logits /
          pred_bboxes / the updated track queries back to pinned host memory.
  roofline   MSDA forward (encoder-shaped launch, the dominant kernel): algorithmic bytes / duration, duration from CUDA
          events recorded inside the captured graph around that launch.
  cpu_baseline / --impl reference   the reference's CPU path (oracle/frame.py, the torch restatement pinned against the
          reference modules) on the host cores.
N > 1: every rank runs its own sub-clip of K frames (weak scaling) and the ranks exchange their packed track-query
memory with ONE NCCL all-gather at the end of the clip; time = max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

def _metric_name():
    """BASELINE.json's metric string, verbatim (the file is part of the repo snapshot)."""
    try:
        return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "BASELINE.json")))["metric"]
    except Exception:                                                    # noqa: BLE001
        return "frames/sec at 1333x800, 300 det+100 track queries"


METRIC = _metric_name()
N_TRACKS = 100
N_ROT = 6            # resident frames rotating through the input buffers: 6 x 22.9 MB (45.8 MB with position maps) > L2


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
            "clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def _pick_cpu_threads(sd, x, cfg):
    """torch intra-op thread count that runs the reference CPU path fastest on this host.  "All the host threads" is
    not automatically the fastest setting: on the 128-core GPU hosts the default (one thread per core) ran a frame in
    52 s against a few seconds with fewer threads (oversubscription on many small ops).  Probe = one encoder layer."""
    from oracle import frame as oframe
    src, mask, pos, shapes, lsi, vr = oframe.flatten_levels(sd, "transformer", x["srcs"], x["masks"], x["pos"])
    ref = oframe.encoder_reference_points(shapes, vr, "cpu")
    n_cpu = os.cpu_count() or 1
    best, best_t = None, float("inf")
    for t in sorted({min(c, n_cpu) for c in (8, 16, 32, 64, n_cpu)}):
        torch.set_num_threads(t)
        with torch.no_grad():
            for rep in range(2):
                t0 = time.perf_counter()
                oframe.encoder_layer(sd, "transformer.encoder.layers.0", src, pos, ref, shapes, lsi, mask, cfg)
                dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = t, dt
        if dt > 4 * best_t:
            break
    torch.set_num_threads(best)
    return best


def cpu_reference_fps(steps, warmup):
    """The reference's CPU path (torch fp32 on the host cores) through the functional oracle."""
    from oracle import frame as oframe
    from oracle import synth
    cfg = oframe.dancetrack_cfg()
    sd = synth.hot_path_state_dict(cfg, seed=0)
    x = synth.frame_inputs(cfg, synth.DANCETRACK_SHAPES, N_TRACKS, seed=1)
    threads = _pick_cpu_threads(sd, x, cfg)
    tracks = dict(x["tracks"])
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            res = oframe.frame_forward(sd, x["srcs"], x["masks"], x["pos"], tracks["ref_pts"], tracks["query_embed"], cfg)
            nd = cfg["n_det_queries"]
            tracks.update(boxes=res["pred_bboxes"][0, nd:], logits=res["pred_logits"][0, nd:],
                          output_embed=res["outputs"][0, nd:])
            tracks.update({k: v for k, v in oframe.update_tracks(sd, tracks, cfg).items() if k != "is_pos"})
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return len(times) / sum(times), threads


def gpu_reference_fps(dev, steps, warmup):
    """The reference GPU path restated: stock PyTorch fp32 ops (TF32 off, main.py:96-97) + the reference's own CUDA op
    compiled into oracle/_ref.  Reported beside our numbers; None when the .so did not travel."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "MultiScaleDeformableAttention.so")):
        return None
    sys.path.insert(0, ref_dir)
    import MultiScaleDeformableAttention as MSDA
    from oracle import frame as oframe
    from oracle import synth
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = oframe.dancetrack_cfg()
    sd = {k: v.to(dev) for k, v in synth.hot_path_state_dict(cfg, seed=0).items()}
    x = synth.frame_inputs(cfg, synth.DANCETRACK_SHAPES, N_TRACKS, seed=1)
    srcs, masks, pos = ([t.to(dev) for t in x[k]] for k in ("srcs", "masks", "pos"))
    tracks = {k: v.to(dev) for k, v in x["tracks"].items()}

    def core(value, shapes_t, lsi, loc, attn):
        return MSDA.ms_deform_attn_forward(value.contiguous(), shapes_t, lsi, loc.contiguous(), attn.contiguous(), 64)

    def step():
        res = oframe.frame_forward(sd, srcs, masks, pos, tracks["ref_pts"], tracks["query_embed"], cfg, core=core)
        nd = cfg["n_det_queries"]
        tracks.update(boxes=res["pred_bboxes"][0, nd:], logits=res["pred_logits"][0, nd:],
                      output_embed=res["outputs"][0, nd:])
        tracks.update({k: v for k, v in oframe.update_tracks(sd, tracks, cfg).items() if k != "is_pos"})

    with torch.no_grad():
        for _ in range(warmup):
            step()
        torch.cuda.synchronize(dev)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            step()
        e.record()
        torch.cuda.synchronize(dev)
    return steps / (s.elapsed_time(e) * 1e-3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)   # ~0.25 s timed region: a few nvidia-smi clock samples
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-baselines", action="store_true", help="skip the cpu_baseline / gpu_reference legs")
    ap.add_argument("--upload-pos", action="store_true",
                    help="upload the position maps with every frame (A/B; default: rebuilt on the device from the masks)")
    ap.add_argument("--no-tracker", action="store_true",
                    help="leave the RuntimeTracker glue out of the step (A/B; default: on the device, inside the graph)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3)
    workload = "DanceTrack hot path: transformer(6 enc + 6 dec, d256, ffn2048, 4 levels S=22323) + heads + " \
               "QueryUpdater, 300 det + 100 track queries, batch 1, synthetic 1333x800 pyramid, backbone excluded"

    if args.impl == "reference":
        # the reference's own CPU implementation of the path, all host threads, rank 0 only
        if rank != 0:
            return
        steps = max(1, min(args.steps, 8))              # bounded sample: one step is one full frame (seconds of CPU)
        fps, cores = cpu_reference_fps(steps, min(W, 2))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": min(W, 2), "ms_per_step": 1e3 / fps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "note": "reference CPU path = PyTorch fp32 ops on the host cores with "
                       "ms_deform_attn_core_pytorch as the sampling core (oracle/frame.py, pinned to the reference modules)"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{steps} full frames after {min(W, 2)} warm-up; thread count = fastest of "
                                       f"8/16/32/64/{os.cpu_count()} on a one-encoder-layer probe"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback for the product path)"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)

    from memotr_b200 import synthetic as synth
    from memotr_b200.engine import FrameEngine
    cfg = synth.dancetrack_cfg()
    sd = synth.hot_path_state_dict(cfg, seed=0)
    frames = [synth.frame_inputs(cfg, synth.DANCETRACK_SHAPES, N_TRACKS, seed=1 + i + 17 * rank) for i in range(N_ROT)]
    # Tracker glue on the device (memotr_b200/tracker.py).  The weights are random, so the thresholds are pinned such that
    # the 100 loaded tracks stay live and nothing is born: the step keeps BASELINE.json's 300 det + 100 track queries.
    tracker = None if args.no_tracker else dict(det_score_thresh=2.0, track_score_thresh=0.0, miss_tolerance=30,
                                                result_score_thresh=0.5)
    # Position maps: a function of the padding masks alone (PositionEmbeddingSine), rebuilt on the device every frame
    # instead of crossing PCIe (the box's pinned H2D rate, ~26 GB/s, would cap e2e at 564 frames/s with them).
    pos_embed = None if args.upload_pos else dict(temperature=20)
    eng = FrameEngine(sd, cfg, synth.DANCETRACK_SHAPES, N_TRACKS, dev, mode=args.mode, tracker=tracker,
                      ori_size=(1920, 1080), pos_embed=pos_embed)
    eng.enable_msda_timer()
    L, C, K = eng.L, eng.C, args.steps

    # resident copies of the rotating frames + pinned host copies for the e2e leg
    res_src = [[f["srcs"][l].reshape(C, -1).to(dev) for l in range(L)] for f in frames]
    res_pos = [[f["pos"][l].reshape(C, -1).to(dev) for l in range(L)] for f in frames] if args.upload_pos else None
    pin = lambda t: t.contiguous().pin_memory()                                      # noqa: E731
    host = [{"srcs": [pin(t) for t in f["srcs"]], "pos": [pin(t) for t in f["pos"]],
             "masks": [pin(t.to(torch.uint8)) for t in f["masks"]]} for f in frames]
    x0 = frames[0]
    eng.load_frame(x0["srcs"], x0["masks"], x0["pos"] if args.upload_pos else None, x0["tracks"]["ref_pts"],
                   x0["tracks"]["query_embed"])
    eng.load_tracks(x0["tracks"])
    if eng.trk is not None:
        eng.trk.reset(x0["tracks"])
    eng.capture()                                       # records step(): forward + hand-off + updater + feedback

    def feed_resident(i):
        for l in range(L):
            eng.in_src[l].copy_(res_src[i % N_ROT][l], non_blocking=True)
            if res_pos is not None:
                eng.in_pos[l].copy_(res_pos[i % N_ROT][l], non_blocking=True)

    def reset_clip():
        eng.in_track_ref.copy_(x0["tracks"]["ref_pts"])
        eng.in_track_embed.copy_(x0["tracks"]["query_embed"])
        eng.load_tracks(x0["tracks"], non_blocking=False)
        if eng.trk is not None:
            eng.trk.reset(x0["tracks"], max_obj_id=N_TRACKS)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    from memotr_b200 import clip as clip_mod

    def clip_exchange():
        """One NCCL all-gather of the packed track-query memory per clip (memotr_b200/clip.py)."""
        return clip_mod.gather_track_memory(eng.st) if world > 1 else None

    # ---- resident-input throughput ("value") -----------------------------------------------------------------
    reset_clip()
    for i in range(W):
        feed_resident(i)
        eng.replay()
    clip_exchange()
    reset_clip()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    host_t0 = time.perf_counter()
    for i in range(K):
        feed_resident(i)
        eng.replay()
    host_enqueue_ms = (time.perf_counter() - host_t0) * 1e3 / K     # CPU time to enqueue one step (must stay < GPU time)
    clip_exchange()
    t1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    tracks_live = int(eng.table.n_active.item()) if eng.trk is not None else N_TRACKS
    if eng.trk is not None:
        eng.trk.check_overflow()
        assert tracks_live == N_TRACKS, f"the bench workload drifted: {tracks_live} live tracks instead of {N_TRACKS}"
    # CPU cost of launching one step with an empty queue (the in-loop figure above includes back-pressure from the GPU)
    torch.cuda.synchronize(dev)
    h0 = time.perf_counter()
    for _ in range(4):
        eng.replay()
    host_launch_ms = (time.perf_counter() - h0) * 1e3 / 4
    torch.cuda.synchronize(dev)
    msda_us = eng.msda_times_us()                       # the 6 encoder MSDA launches of the last timed step
    sections = eng.section_times_us()
    # the fused encoder FFN (largest kernel by time, tensor-bound) timed on its own after the run: event nodes around it
    # inside the graph would cost the programmatic-launch overlap with its neighbours (measured: -4 % step throughput)
    ffn_us = None
    if eng.fused_mlp and rank == 0:
        ly = eng.enc[0]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(12)]
        for i in range(8):
            if i >= 2:
                ev[2 * (i - 2)].record()
            eng.mlp2(eng.src1, eng.C, ly["lin1"], ly["lin2"], eng.pre, eng.C, eng.S, eng.hid, c_dtype=0)
            if i >= 2:
                ev[2 * (i - 2) + 1].record()
        torch.cuda.synchronize(dev)
        ffn_us = [ev[2 * i].elapsed_time(ev[2 * i + 1]) * 1e3 for i in range(6)]

    # ---- end to end through the public API with host buffers ("e2e") ----------------------------------------------
    from memotr_b200.engine import ClipRunner
    runner = ClipRunner(eng)
    h2d, d2h = runner.h2d_bytes, runner.d2h_bytes
    hf = [(h["srcs"], h["pos"] if args.upload_pos else None, h["masks"]) for h in host]

    def e2e_clip(n):
        runner.prefetch(0, *hf[0])
        for i in range(n):
            if i + 1 < n:
                runner.prefetch((i + 1) % 2, *hf[(i + 1) % N_ROT])
            runner.run(i % 2)

    reset_clip()
    e2e_clip(W)
    reset_clip()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_clip(K)
    clip_exchange()
    e1.record()
    barrier()
    ems = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_fps = world * K / (float(ems.item()) * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (MSDA forward, encoder-shaped launch) ----------------------------------------
    S, H, LK = eng.S, eng.H, eng.L * cfg["n_enc_points"]
    esz = 2 if args.mode == "bf16" else 4
    # value read once + output written once (activation dtype) + sampling locations and weights (fp32): BASELINE.md sec. 3
    alg_bytes = S * H * 32 * esz + S * H * 32 * esz + S * H * LK * 3 * 4
    peak, peak_src = peaks()
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_msda_fwd_h16_traffic.json")
    if eng.value_f16 and os.path.exists(tpath):          # dram bytes of this kernel from the committed ncu --set full capture
        tj = json.load(open(tpath))
        traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    dur = sum(msda_us) / len(msda_us)
    achieved = alg_bytes / dur / 1e3
    fps = world * K / (ms_total * 1e-3)
    out = {
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.mode == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": workload, "clip": f"{K} chained frames per GPU; N>1: one NCCL all-gather of the packed "
                   "track-query memory per clip", "l2": f"inputs larger than L2: {N_ROT} resident frames x {h2d / 1e6:.1f} MB rotate "
                   "through the input buffers and a step touches ~0.5 GB of workspace (L2 = 126 MB)",
                   "position_maps": "uploaded with every frame (--upload-pos)" if args.upload_pos else
                   "PositionEmbeddingSine rebuilt on the device from the padding masks inside the captured step",
                   "tracker": ("RuntimeTracker.update + select_active_tracks + result filter on the device inside the "
                               f"captured step; thresholds pinned so that {tracks_live} tracks stay live and none is born")
                   if eng.trk is not None else "off (--no-tracker)",
                   "arithmetic": "bf16 GEMM operands + fp32 accumulate/residual/LayerNorm/geometry" if args.mode == "bf16"
                   else "fp32 everywhere (TF32 off, as the reference)"},
        "clocks": clocks,
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": eng.graph_launches * K,
        "sections_us": {k: round(v, 1) for k, v in sections.items()},
        "host_enqueue_ms_per_step": round(host_enqueue_ms, 3),
        "host_graph_launch_ms": round(host_launch_ms, 3),
        "roofline": {"kernel": ("msda_fwd_h16 (fp16 value map)" if eng.value_f16 else "msda_fwd_vec") +
                               " -- encoder-shaped launch, Lq = S = 22323", "bound": "hbm",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": "profiles/r01_msda_fwd_h16_traffic.json (ncu --set full)" if traffic else None,
                     "peak_source": peak_src, "algorithmic_bytes": alg_bytes, "duration_us": dur,
                     "samples": f"{len(msda_us)} launches (the encoder layers of the last timed step), CUDA events "
                                "recorded inside the captured graph",
                     "ceiling_note": "on-chip gather traffic (S*H*L*K*4 corners*32 ch) is ~18x the algorithmic bytes; see DESIGN.md"},
    }
    if ffn_us:      # second roofline object: the largest kernel by time is tensor-bound (fused encoder FFN)
        fdur = sum(ffn_us) / len(ffn_us)
        flops = 2.0 * 2.0 * eng.S * eng.C * eng.Fd
        tpeak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", None) \
            if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else None
        tpeak = float(tpeak) if tpeak else 1430.2
        out["roofline_tensor"] = {"kernel": "mlp2_tc_kernel -- fused encoder FFN (22323 x 256 -> 2048 -> 256), main launch + tail-split launch",
                                  "bound": "tensor", "achieved": flops / fdur / 1e6, "peak": tpeak, "unit": "TFLOP/s",
                                  "frac": flops / fdur / 1e6 / tpeak, "flop": flops, "duration_us": fdur,
                                  "peak_source": "MEASURED_PEAKS.json bf16 dense, sustained",
                                  "samples": "6 back-to-back launches after the timed region (CUDA events; operands L2-warm)",
                                  "note": "shared-memory-bandwidth-bound in its cta_group::1 form, see profiles/r01_mlp2_ncu.md"}
    if not args.no_baselines and world == 1:      # the CPU / reference-GPU legs are timed at N = 1 only
        cpu_fps, cores = cpu_reference_fps(3, 1)
        out["cpu_baseline"] = {"value": cpu_fps, "unit": "frames/s", "cores": cores, "kind": "port",
                               "sample": "3 full frames after 1 warm-up (oracle/frame.py on the host cores, torch fp32; "
                                         f"thread count = fastest of 8/16/32/64/{os.cpu_count()} on a one-layer probe)"}
        g = gpu_reference_fps(dev, 10, 3)
        out["gpu_reference"] = {"value": g, "unit": "frames/s",
                                "what": "reference models/ops CUDA op (oracle/_ref, compiled from /root/reference) + stock "
                                        "PyTorch fp32 eager modules (TF32 off) on the same GPU and inputs"} if g else None
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
