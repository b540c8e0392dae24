"""memotr_b200/clip.py -- sharding a clip over the GPUs of one box (BASELINE.json configs[3]).

The reference only shards whole sequences over ranks (submit_engine.py:225-231).  Here the frames of ONE clip are split
into contiguous sub-clips, one per rank (`shard_frames`), in one of two ways:

  independent sub-clips (`gather_track_memory`)   every rank runs its sub-clip with its own track state and the ranks
      exchange their complete track memory -- every TrackInstances field (structures/track_instances.py:18-37) --
      exactly once per clip with a single all-gather.  No collective on the per-frame path.  Sub-clips that do not start
      from the previous rank's tracks are NOT the sequential result (SURVEY.md 8e "parity caveat").

  exact two-phase clip (`run_clip_two_phase`)     phase 1, frame-parallel: every rank runs everything that depends on the
      frame alone (level flattening, the six encoder layers, the stacked decoder value projection: >= 95 % of the FLOPs)
      for its frames and keeps the per-frame result; phase 2, the recurrent tail (decoder, heads, tracker glue, query
      updater: consumes the previous frame's tracks, submit_engine.py:64-72) runs as a hand-off chain: rank r receives the
      track memory from rank r-1, advances it over its own frames, and sends it on.  Bit-identical to one rank running
      the whole clip; the speed-up is bounded by the serial tail.

The same code runs on NCCL / CUDA tensors and on gloo / CPU tensors (tests/test_clip_cpu.py, world size 2).

Packed track memory of one rank (bytes, contiguous; `pack_track_state` / `unpack_track_state`):
    ids | labels | disappear_time   int64 (Nt each),  n_active, max_obj_id int64 (1 each)
    query_embed | long_memory | last_output | output_embed   fp32 (Nt x C each)
    ref_pts | boxes fp32 (Nt x 4 each) | logits fp32 (Nt x n_cls)
"""
import torch
import torch.distributed as dist

FLOAT_FIELDS = ("query_embed", "long_memory", "last_output", "output_embed", "ref_pts", "boxes", "logits")
INT_FIELDS = ("ids", "labels", "disappear_time")
FIELDS = FLOAT_FIELDS                      # (historical name)


def shard_frames(n_frames: int, world: int, rank: int) -> range:
    """Contiguous block of frame indices owned by `rank`; the blocks partition range(n_frames) and differ by <= 1."""
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def _widths(d_model: int, n_cls: int) -> dict:
    return {"query_embed": d_model, "long_memory": d_model, "last_output": d_model, "output_embed": d_model,
            "ref_pts": 4, "boxes": 4, "logits": n_cls}


def packed_nbytes(n_tracks: int, d_model: int, n_cls: int) -> int:
    return n_tracks * (4 * d_model + 8 + n_cls) * 4 + (3 * n_tracks + 2) * 8


def pack_track_state(state: dict, n_active=None, max_obj_id=0) -> torch.Tensor:
    """state: the float fields (Nt rows each) and, optionally, the int64 fields ids / labels / disappear_time (defaults:
    arange, 0, 0); n_active: int, 1-element tensor or None (= Nt); max_obj_id: int or 1-element tensor (the tracker's id
    counter, runtime_tracker.py:85-89).  -> uint8 (packed_nbytes,) on the fields' device."""
    ref = state["query_embed"]
    nt, dev = ref.shape[0], ref.device
    floats = torch.cat([state[k].reshape(-1).to(torch.float32) for k in FLOAT_FIELDS])
    ints = [state.get("ids", None), state.get("labels", None), state.get("disappear_time", None)]
    ints[0] = torch.arange(nt, dtype=torch.long, device=dev) if ints[0] is None else ints[0]
    ints = [torch.zeros(nt, dtype=torch.long, device=dev) if t is None else t.reshape(-1).to(torch.long) for t in ints]
    if n_active is None:
        n_active = torch.full((1,), nt, dtype=torch.long, device=dev)
    elif not torch.is_tensor(n_active):
        n_active = torch.full((1,), int(n_active), dtype=torch.long, device=dev)
    ints.append(n_active.reshape(-1).to(device=dev, dtype=torch.long))
    if not torch.is_tensor(max_obj_id):
        max_obj_id = torch.full((1,), int(max_obj_id), dtype=torch.long, device=dev)
    ints.append(max_obj_id.reshape(-1).to(device=dev, dtype=torch.long))
    return torch.cat([torch.cat(ints).contiguous().view(torch.uint8), floats.contiguous().view(torch.uint8)])


def unpack_track_state(flat: torch.Tensor, n_tracks: int, d_model: int, n_cls: int) -> dict:
    """Inverse of pack_track_state -> dict of views with every field plus "n_active" / "max_obj_id" (1-element int64)."""
    assert flat.dtype == torch.uint8 and flat.numel() == packed_nbytes(n_tracks, d_model, n_cls), (flat.dtype, flat.numel())
    ni = (3 * n_tracks + 2) * 8                    # the int64 block comes first so that both views are aligned
    ints, floats = flat[:ni].view(torch.long), flat[ni:].view(torch.float32)
    out, o = {}, 0
    for k, w in _widths(d_model, n_cls).items():
        out[k] = floats[o:o + n_tracks * w].view(n_tracks, w)
        o += n_tracks * w
    for i, k in enumerate(INT_FIELDS):
        out[k] = ints[i * n_tracks:(i + 1) * n_tracks]
    out["n_active"], out["max_obj_id"] = ints[3 * n_tracks:3 * n_tracks + 1], ints[3 * n_tracks + 1:]
    return out


def _world(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def gather_track_memory(state: dict, n_active=None, max_obj_id=0, group=None) -> torch.Tensor:
    """THE collective of a clip sharded into independent sub-clips: all-gather every rank's packed track memory ->
    uint8 (world, packed_nbytes).  One call per clip."""
    packed = pack_track_state(state, n_active, max_obj_id)
    world = _world(group)
    if world == 1:
        return packed[None]
    out = torch.empty(world * packed.numel(), dtype=torch.uint8, device=packed.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    return out.view(world, -1)


def run_clip_two_phase(n_frames, encode, decode, get_state, set_state, group=None):
    """The exact clip over the ranks of `group` (see the module docstring).

      encode(i) -> token      phase 1 for frame i: whatever depends on the frame alone; `token` is handed back to decode
      decode(i, token) -> r   phase 2 for frame i: the recurrent tail on the CURRENT track state; r = the frame's results
      get_state() -> uint8 tensor / set_state(uint8 tensor)   the packed track memory (pack_track_state) of this rank

    Rank r owns frames shard_frames(n_frames, world, r).  Returns the list of this rank's (frame index, result) pairs."""
    world = _world(group)
    rank = dist.get_rank(group) if world > 1 else 0
    mine = shard_frames(n_frames, world, rank)
    tokens = [encode(i) for i in mine]                                   # phase 1: frame-parallel, no communication
    if world > 1 and rank > 0 and len(mine) > 0:                          # phase 2: wait for the previous rank's tracks
        buf = torch.empty_like(get_state())
        dist.recv(buf, src=_prev_nonempty(n_frames, world, rank), group=group)
        set_state(buf)
    out = [(i, decode(i, tok)) for i, tok in zip(mine, tokens)]
    nxt = _next_nonempty(n_frames, world, rank)
    if world > 1 and nxt is not None and len(mine) > 0:
        dist.send(get_state().contiguous(), dst=nxt, group=group)
    return out


def _prev_nonempty(n_frames, world, rank):
    for r in range(rank - 1, -1, -1):
        if len(shard_frames(n_frames, world, r)):
            return r
    return None


def _next_nonempty(n_frames, world, rank):
    for r in range(rank + 1, world):
        if len(shard_frames(n_frames, world, r)):
            return r
    return None
