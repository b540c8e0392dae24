"""Host-side logic of clip sharding on CPU: frame partition, pack/unpack, and the single all-gather of the
track-query memory exercised with the gloo backend at world_size 2 (the N>1 path of bench.py uses the same function
over NCCL)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from memotr_b200 import clip


def test_shard_frames_partitions_the_clip():
    for n, w in [(64, 8), (64, 1), (7, 3), (5, 8), (0, 4)]:
        blocks = [clip.shard_frames(n, w, r) for r in range(w)]
        assert [i for b in blocks for i in b] == list(range(n))                # contiguous, ordered, complete
        assert max(len(b) for b in blocks) - min(len(b) for b in blocks) <= 1   # balanced
    assert clip.shard_frames(64, 8, 3) == range(24, 32)                         # BASELINE configs[3]: 8 frames per GPU


def _state(nt, C, ncls, seed):
    g = torch.Generator().manual_seed(seed)
    return {"query_embed": torch.randn(nt, C, generator=g), "long_memory": torch.randn(nt, C, generator=g),
            "last_output": torch.randn(nt, C, generator=g), "output_embed": torch.randn(nt, C, generator=g),
            "ref_pts": torch.randn(nt, 4, generator=g), "boxes": torch.rand(nt, 4, generator=g),
            "logits": torch.randn(nt, ncls, generator=g)}


def test_pack_unpack_roundtrip_bit_exact():
    for nt, C, ncls in [(100, 256, 1), (500, 256, 8), (0, 256, 1)]:
        st = _state(nt, C, ncls, 1)
        flat = clip.pack_track_state(st)
        assert flat.numel() == clip.packed_numel(nt, C, ncls)
        back = clip.unpack_track_state(flat, nt, C, ncls)
        for k in clip.FIELDS:
            assert torch.equal(back[k], st[k])
    assert clip.gather_track_memory(_state(3, 8, 1, 2)).shape == (1, clip.packed_numel(3, 8, 1))   # no process group


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nt, C, ncls = 7, 16, 2
        gathered = clip.gather_track_memory(_state(nt, C, ncls, 100 + rank))
        ok = gathered.shape == (world, clip.packed_numel(nt, C, ncls))
        for r in range(world):                                   # every rank sees every rank's memory, in rank order
            want = _state(nt, C, ncls, 100 + r)
            got = clip.unpack_track_state(gathered[r], nt, C, ncls)
            ok = ok and all(torch.equal(got[k], want[k]) for k in clip.FIELDS)
        frames = list(clip.shard_frames(10, world, rank))
        q.put((rank, bool(ok), frames))
    finally:
        dist.destroy_process_group()


def test_single_all_gather_of_track_memory_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert res[0][2] + res[1][2] == list(range(10))
