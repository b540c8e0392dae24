"""memotr_b200/clip.py -- sharding a clip over the GPUs of one box (BASELINE.json configs[3]).

The reference only shards whole sequences over ranks (submit_engine.py:225-231).  Here the frames of ONE clip are split
into contiguous sub-clips, one per rank; every rank runs its sub-clip with its own FrameEngine and track state, and the
ranks exchange their track-query memory exactly once per clip with a single all-gather (NCCL over NVLink on GPUs; the
same code runs on gloo/CPU tensors in the tests).  There is no collective on the per-frame path.

Packed layout of one rank's track memory (fp32, contiguous):
    query_embed | long_memory | last_output | output_embed   (Nt x C each)   -- TrackInstances fields, track_instances.py:18-37
    ref_pts | boxes                                          (Nt x 4 each)
    logits                                                   (Nt x n_cls)
"""
import torch
import torch.distributed as dist

FIELDS = ("query_embed", "long_memory", "last_output", "output_embed", "ref_pts", "boxes", "logits")


def shard_frames(n_frames: int, world: int, rank: int) -> range:
    """Contiguous block of frame indices owned by `rank`; the blocks partition range(n_frames) and differ by <= 1."""
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def packed_numel(n_tracks: int, d_model: int, n_cls: int) -> int:
    return n_tracks * (4 * d_model + 8 + n_cls)


def pack_track_state(state: dict) -> torch.Tensor:
    return torch.cat([state[k].reshape(-1).float() for k in FIELDS])


def unpack_track_state(flat: torch.Tensor, n_tracks: int, d_model: int, n_cls: int) -> dict:
    assert flat.numel() == packed_numel(n_tracks, d_model, n_cls)
    widths = {"query_embed": d_model, "long_memory": d_model, "last_output": d_model, "output_embed": d_model,
              "ref_pts": 4, "boxes": 4, "logits": n_cls}
    out, o = {}, 0
    for k in FIELDS:
        n = n_tracks * widths[k]
        out[k] = flat[o:o + n].view(n_tracks, widths[k])
        o += n
    return out


def gather_track_memory(state: dict, group=None) -> torch.Tensor:
    """THE collective of a sharded clip: all-gather every rank's packed track memory -> (world, packed_numel).
    One call per clip; with an initialised process group only (single-process runs return a (1, n) view)."""
    packed = pack_track_state(state)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return packed[None]
    world = dist.get_world_size(group)
    out = torch.empty(world * packed.numel(), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed.contiguous(), group=group)
    return out.view(world, -1)
