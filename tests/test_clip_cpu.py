"""Host-side logic of clip sharding on CPU: frame partition, pack/unpack of the complete track memory (every
TrackInstances field, structures/track_instances.py:18-37), the single all-gather of the independent-sub-clip mode and the
hand-off chain of the exact two-phase mode, exercised with the gloo backend at world_size 2 and 3 (the N>1 paths of bench.py
use the same functions over NCCL)."""
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


def _state(nt, C, ncls, seed, ints=True):
    g = torch.Generator().manual_seed(seed)
    st = {"query_embed": torch.randn(nt, C, generator=g), "long_memory": torch.randn(nt, C, generator=g),
          "last_output": torch.randn(nt, C, generator=g), "output_embed": torch.randn(nt, C, generator=g),
          "ref_pts": torch.randn(nt, 4, generator=g), "boxes": torch.rand(nt, 4, generator=g),
          "logits": torch.randn(nt, ncls, generator=g)}
    if ints:
        st["ids"] = torch.randint(-1, 10 ** 12, (nt,), generator=g)            # int64 range, -1 = dead
        st["labels"] = torch.randint(0, max(ncls, 1), (nt,), generator=g)
        st["disappear_time"] = torch.randint(0, 30, (nt,), generator=g)
    return st


def test_pack_unpack_roundtrip_bit_exact():
    for nt, C, ncls in [(100, 256, 1), (500, 256, 8), (0, 256, 1)]:
        st = _state(nt, C, ncls, 1)
        flat = clip.pack_track_state(st, n_active=max(nt - 3, 0), max_obj_id=nt + 11)
        assert flat.dtype == torch.uint8 and flat.numel() == clip.packed_nbytes(nt, C, ncls)
        back = clip.unpack_track_state(flat, nt, C, ncls)
        for k in clip.FLOAT_FIELDS + clip.INT_FIELDS:
            assert torch.equal(back[k], st[k]) and back[k].dtype == st[k].dtype, k
        assert back["n_active"].tolist() == [max(nt - 3, 0)] and back["max_obj_id"].tolist() == [nt + 11]
    # defaults when the caller keeps no bookkeeping: ids = arange, labels = disappear_time = 0, n_active = Nt
    back = clip.unpack_track_state(clip.pack_track_state(_state(5, 8, 1, 2, ints=False)), 5, 8, 1)
    assert back["ids"].tolist() == [0, 1, 2, 3, 4] and back["labels"].sum() == 0 and back["n_active"].item() == 5
    assert clip.gather_track_memory(_state(3, 8, 1, 2)).shape == (1, clip.packed_nbytes(3, 8, 1))   # no process group
    assert clip.packed_nbytes(100, 256, 1) == 100 * (4 * 256 + 8 + 1) * 4 + 302 * 8                # 0.42 MB per rank


# ---- a toy recurrent "engine" with the two-phase interface: the state after frame i depends on every earlier frame ------
NT, C, NCLS = 6, 8, 2


def _toy(seed=0):
    g = torch.Generator().manual_seed(seed)
    M = torch.randn(C, C, generator=g) / 3
    st = {"s": _state(NT, C, NCLS, 7)}

    def encode(i):                                    # frame-only work
        return torch.sin(torch.arange(C, dtype=torch.float32) * (i + 1))

    def decode(i, tok):                               # recurrent tail: consumes the current tracks
        s = st["s"]
        s["query_embed"] = torch.tanh(s["query_embed"] @ M + tok)
        s["long_memory"] = 0.9 * s["long_memory"] + 0.1 * s["query_embed"]
        s["ids"] = s["ids"] + (i % 3 == 0)
        s["disappear_time"] = (s["disappear_time"] + i) % 7
        return float(s["query_embed"].sum())

    return encode, decode, (lambda: clip.pack_track_state(st["s"], 4)), \
        (lambda b: st.__setitem__("s", {k: v.clone() for k, v in clip.unpack_track_state(b, NT, C, NCLS).items()
                                        if k not in ("n_active", "max_obj_id")}))


def _sequential(n_frames):
    encode, decode, get_state, _ = _toy()
    res = [decode(i, encode(i)) for i in range(n_frames)]
    return res, get_state()


def test_two_phase_clip_single_process_is_the_sequential_clip():
    encode, decode, get_state, set_state = _toy()
    out = clip.run_clip_two_phase(9, encode, decode, get_state, set_state)
    want, want_state = _sequential(9)
    assert [i for i, _ in out] == list(range(9)) and [r for _, r in out] == want
    assert torch.equal(get_state(), want_state)


def _worker(rank, world, port, q, n_frames):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nt, Cc, ncls = 7, 16, 2
        gathered = clip.gather_track_memory(_state(nt, Cc, ncls, 100 + rank), n_active=rank + 1)
        ok = gathered.shape == (world, clip.packed_nbytes(nt, Cc, ncls))
        for r in range(world):                                   # every rank sees every rank's memory, in rank order
            want = _state(nt, Cc, ncls, 100 + r)
            got = clip.unpack_track_state(gathered[r], nt, Cc, ncls)
            ok = ok and all(torch.equal(got[k], want[k]) for k in clip.FLOAT_FIELDS + clip.INT_FIELDS)
            ok = ok and got["n_active"].item() == r + 1
        frames = list(clip.shard_frames(10, world, rank))
        # exact two-phase clip: hand-off chain of the packed track memory
        encode, decode, get_state, set_state = _toy()
        out = clip.run_clip_two_phase(n_frames, encode, decode, get_state, set_state)
        q.put((rank, bool(ok), frames, out, get_state()))
    finally:
        dist.destroy_process_group()


def _spawn(world, n_frames):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import socket
    with socket.socket() as sk:                    # a free port chosen by the kernel (a pid-derived one collided once)
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, n_frames)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_all_gather_and_two_phase_chain_gloo_world2():
    res = _spawn(2, 9)
    assert [r[1] for r in res] == [True, True]
    assert res[0][2] + res[1][2] == list(range(10))
    want, want_state = _sequential(9)
    got = [r for rank in res for _, r in rank[3]]
    assert [i for rank in res for i, _ in rank[3]] == list(range(9)) and got == want     # every frame once, same results
    assert torch.equal(res[-1][4], want_state)                                             # last rank ends with the clip's tracks


def test_two_phase_chain_gloo_world3_with_an_idle_rank():
    res = _spawn(3, 2)            # 2 frames over 3 ranks: the last rank owns nothing and must not dead-lock the chain
    want, want_state = _sequential(2)
    assert [r for rank in res for _, r in rank[3]] == want and res[2][3] == []
    assert torch.equal(res[1][4], want_state)
