"""GPU tests of the training path (BASELINE.json config 3, at test size): a mixed-precision (bf16 autocast) step of the
mirrored transformer on our MSDA forward / backward kernels, and a DistributedDataParallel step over NCCL on two GPUs
(skipped on a one-GPU box; run with `gpurun --gpus 2`).  The DDP test also exercises the C library on cuda:1."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import rel_err
from test_modules_gpu import _load, _queries

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _loss(tr, x, emb, ref, qmask, R, dev):
    d = lambda ts: [t.to(dev) for t in ts]                                      # noqa: E731
    outs, _, refs, _ = tr(d(x["srcs"]), d(x["masks"]), d(x["pos"]), emb.to(dev), ref.to(dev), qmask.to(dev))
    return (outs * R.to(dev)).sum() + refs[-1].sum()


def test_bf16_autocast_training_step_matches_fp32_gradients():
    """forward + backward under torch.autocast(bf16): every nn.Linear / MHA GEMM in bf16, the sampling core in fp32 (the
    Function casts its inputs up: the reference op has no half instantiation).  Gradients stay within bf16 noise of the
    fp32 step and one SGD step moves the loss the same way."""
    _, cfg, sd, x, tr, _ = _load("small_padded")
    tr.train()
    emb, ref, qmask = _queries(sd, x)
    R = torch.randn(cfg["n_dec_layers"], 1, emb.shape[1], 256, generator=torch.Generator().manual_seed(3))
    l32 = _loss(tr, x, emb, ref, qmask, R, DEV)
    l32.backward()
    g32 = {k: p.grad.clone() for k, p in tr.named_parameters() if p.grad is not None}
    tr.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        l16 = _loss(tr, x, emb, ref, qmask, R, DEV)
    l16.backward()
    assert abs(float(l16.detach()) - float(l32.detach())) <= 2e-2 * max(1.0, abs(float(l32.detach())))
    # (white-noise synthetic weights amplify rounding noise -- DESIGN.md section 6 -- so single parameters with small gradients can
    #  deviate a lot relative to their own maximum; the step as a whole must point the same way)
    dots, n16, n32, per = 0.0, 0.0, 0.0, []
    for k, p in tr.named_parameters():
        if k not in g32:
            continue
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
        assert p.grad.dtype == torch.float32
        a, b = p.grad.double().flatten(), g32[k].double().flatten()
        dots, n16, n32 = dots + float(a @ b), n16 + float(a @ a), n32 + float(b @ b)
        per.append((float(torch.nn.functional.cosine_similarity(a, b, dim=0)), float(b.norm()), k))
    cos = dots / (n16 * n32) ** 0.5
    rel = (n16 + n32 - 2 * dots) ** 0.5 / n32 ** 0.5
    per.sort()
    print(f"bf16-autocast vs fp32 step: gradient cosine {cos:.5f}, relative L2 deviation {rel:.3e} over {len(per)} parameters; "
          f"lowest per-parameter cosines: {[(round(c, 3), f'{n:.1e}', k) for c, n, k in per[:4]]}")
    # the yardstick: stock PyTorch under the same autocast (functional oracle on the GPU, grid_sample core) against ITS fp32 step
    from oracle import frame as oframe

    def oracle_grads(autocast):
        sd_g = {k: v.to(DEV).clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
        d = lambda ts: [t.to(DEV) for t in ts]                                  # noqa: E731
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            o_outs, _, o_refs, _, _ = oframe.transformer(sd_g, d(x["srcs"]), d(x["masks"]), d(x["pos"]), emb.to(DEV), ref.to(DEV),
                                                         qmask.to(DEV), cfg)
            loss = (o_outs * R.to(DEV)).sum() + o_refs[-1].sum()
        loss.backward()
        return torch.cat([v.grad.double().flatten() for k, v in sorted(sd_g.items())
                          if k.startswith("transformer.") and v.grad is not None])
    o32, o16 = oracle_grads(False), oracle_grads(True)
    o_rel = float((o16 - o32).norm() / o32.norm())
    print(f"stock PyTorch under the same autocast: relative L2 deviation {o_rel:.3e}")
    assert cos > 0.98 and rel < 1.5 * o_rel + 0.02


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _ddp_worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world)
        dev = torch.device("cuda", rank)
        from memotr_b200 import synthetic as synth
        _, cfg, sd, _, tr, _ = _load("small_padded")
        tr = tr.to(dev).train()
        xs = [synth.frame_inputs(cfg, synth.SMALL_SHAPES, 5, seed=40 + r, padded=True) for r in range(world)]
        R = torch.randn(cfg["n_dec_layers"], 1, cfg["n_det_queries"] + 5, 256, generator=torch.Generator().manual_seed(3))
        # reference: the mean over ranks of the local gradients, computed without DDP on this rank
        want = None
        for r in range(world):
            tr.zero_grad()
            emb, ref, qmask = _queries(sd, xs[r])
            _loss(tr, xs[r], emb, ref, qmask, R, dev).backward()
            g = {k: p.grad.clone() for k, p in tr.named_parameters() if p.grad is not None}
            want = g if want is None else {k: want[k] + g[k] for k in g}
        want = {k: v / world for k, v in want.items()}
        tr.zero_grad()
        ddp = torch.nn.parallel.DistributedDataParallel(tr, device_ids=[rank], find_unused_parameters=True)
        emb, ref, qmask = _queries(sd, xs[rank])
        d = lambda ts: [t.to(dev) for t in ts]                                  # noqa: E731
        outs, _, refs, _ = ddp(d(xs[rank]["srcs"]), d(xs[rank]["masks"]), d(xs[rank]["pos"]), emb.to(dev), ref.to(dev), qmask.to(dev))
        ((outs * R.to(dev)).sum() + refs[-1].sum()).backward()
        worst = 0.0
        for k, p in tr.named_parameters():
            if k in want:
                worst = max(worst, rel_err(p.grad.cpu().numpy(), want[k].cpu().numpy()))
        # one optimizer step keeps the replicas identical
        opt = torch.optim.SGD(ddp.parameters(), lr=1e-3)
        opt.step()
        flat = torch.cat([p.detach().flatten() for p in tr.parameters()])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        same = all(torch.equal(gathered[0], t) for t in gathered)
        q.put((rank, worst, bool(same), None))
        dist.destroy_process_group()
    except Exception as e:      # surface the failure in the parent
        import traceback
        q.put((rank, None, False, traceback.format_exc()[-1500:]))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_ddp_training_step_on_two_gpus():
    """DistributedDataParallel over NCCL, one rank per GPU, different frames per rank: the all-reduced gradients equal the
    mean of the per-frame gradients (our autograd Function under DDP's hooks, our kernels on cuda:1), and the replicas
    hold identical parameters after an optimizer step."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, worst, same, err in res:
        assert err is None, f"rank {rank}: {err}"
        print(f"rank {rank}: DDP gradient vs mean of local gradients {worst:.2e}, replicas identical after the step: {same}")
        assert worst < 1e-5 and same
