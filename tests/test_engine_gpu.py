"""GPU parity tests of the dense/row kernels and of the whole frame engine (run on the B200 box: `pytest -m gpu`).

Kernels are compared with a plain PyTorch fp32 reference of the same op evaluated on the CPU (fp64 where cheap);
the engine is compared with tests/golden/frame_*.npz -- outputs of the reference's own nn.Modules (MeMOTR.forward and
QueryUpdater.update_tracks_embedding), produced in the authoring container by oracle/make_golden.py -- and with the
functional oracle (oracle/frame.py) on the same seeded inputs.
Tolerances: fp32 <= 1e-4, bf16 <= 1e-2 (north star), stated per test as max|a-b|/max|b|.
"""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, rel_err
from oracle import frame as oframe
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


def K():
    from memotr_b200 import kernels
    return kernels


def _g(seed):
    return torch.Generator().manual_seed(seed)


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K_", [(1, 4, 256), (37, 1, 256), (300, 384, 256), (400, 256, 512), (777, 2048, 256),
                                    (513, 256, 2048), (130, 70, 33)])
def test_linear_fp32_simt(M, N, K_):
    g = _g(M + N)
    x, w, b = torch.randn(M, K_, generator=g), torch.randn(N, K_, generator=g) / math.sqrt(K_), torch.randn(N, generator=g)
    want = F.linear(x.double(), w.double(), b.double())
    got = K().linear(x.to(DEV), w.to(DEV), b.to(DEV)).cpu()
    assert rel_err(got, want) < 1e-5


def test_linear_fp32_epilogues():
    g = _g(7)
    M, N, K_ = 333, 256, 256
    x, w, b = torch.randn(M, K_, generator=g), torch.randn(N, K_, generator=g) / 16, torch.randn(N, generator=g)
    mul, add = torch.randn(M, N, generator=g), torch.randn(M, N, generator=g)
    rz = (torch.rand(M, generator=g) < 0.2)
    base = F.linear(x.double(), w.double(), b.double())
    d = lambda t: t.to(DEV)                                                                 # noqa: E731
    assert rel_err(K().linear(d(x), d(w), d(b), act="relu").cpu(), base.relu()) < 1e-5
    assert rel_err(K().linear(d(x), d(w), d(b), act="sigmoid", mul=d(mul)).cpu(), base.sigmoid() * mul.double()) < 1e-5
    assert rel_err(K().linear(d(x), d(w), d(b), add=d(add)).cpu(), base + add.double()) < 1e-5
    got = K().linear(d(x), d(w), d(b), rowzero=d(rz.to(torch.uint8))).cpu()
    assert torch.count_nonzero(got[rz]) == 0 and rel_err(got[~rz], base[~rz]) < 1e-5
    # strided input / output views (column slices of wider buffers)
    wide_in = torch.randn(M, 2 * K_, generator=g)
    wide_out = torch.zeros(M, 3 * N, device=DEV)
    K().linear(d(wide_in)[:, K_:], d(w), d(b), out=wide_out[:, N:2 * N])
    assert rel_err(wide_out[:, N:2 * N].cpu(), F.linear(wide_in[:, K_:].double(), w.double(), b.double())) < 1e-5
    assert torch.count_nonzero(wide_out[:, :N]) == 0 and torch.count_nonzero(wide_out[:, 2 * N:]) == 0


@pytest.mark.parametrize("M,N,K_", [(128, 64, 64), (1, 128, 256), (300, 384, 256), (400, 256, 512), (4000, 2048, 256),
                                    (1025, 256, 2048), (22323, 256, 256), (22323, 1536, 256), (22323, 384, 256),
                                    (19001, 128, 2048)])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32, torch.float16])
def test_linear_bf16_tensor_core(M, N, K_, out_dtype):
    """Shapes with more 128 x 128 tiles than SMs run the persistent kernel (csrc/gemm_tc_persist.cu), the others one
    tile per CTA (csrc/gemm_tc.cu)."""
    g = _g(M + N + 1)
    x = torch.randn(M, K_, generator=g).bfloat16()
    w = (torch.randn(N, K_, generator=g) / math.sqrt(K_)).bfloat16()
    b = torch.randn(N, generator=g)
    want = F.linear(x.double(), w.double(), b.double())        # exact product of the bf16-rounded operands
    got = K().linear(x.to(DEV), w.to(DEV), b.to(DEV), out_dtype=out_dtype, path="tc").float().cpu()
    tol = {torch.float32: 1e-5, torch.bfloat16: 6e-3, torch.float16: 8e-4}[out_dtype]   # output rounding 2^-8 / 2^-11
    assert rel_err(got, want) < tol


@pytest.mark.parametrize("M,N,K_", [(1, 64, 128), (100, 256, 256), (300, 512, 256), (400, 384, 256), (400, 2048, 256),
                                    (400, 256, 2048), (333, 256, 512), (1000, 64, 384)])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_linear_bf16_few_rows_mma(M, N, K_, out_dtype):
    """The latency-optimised mma.sync kernel the decoder / updater GEMMs (<= 1024 rows) run on."""
    g = _g(M + N + 2)
    x = torch.randn(M, K_, generator=g).bfloat16()
    w = (torch.randn(N, K_, generator=g) / math.sqrt(K_)).bfloat16()
    b = torch.randn(N, generator=g)
    mul, add = torch.randn(M, N, generator=g).bfloat16(), torch.randn(M, N, generator=g).bfloat16()
    rz = torch.rand(M, generator=g) < 0.2
    base = F.linear(x.double(), w.double(), b.double())
    d = lambda t: t.to(DEV)                                                                 # noqa: E731
    tol = 1e-5 if out_dtype == torch.float32 else 6e-3
    got = K().linear(d(x), d(w), d(b), out_dtype=out_dtype, path="mma").float().cpu()
    assert rel_err(got, base) < tol
    got = K().linear(d(x), d(w), d(b), act="sigmoid", mul=d(mul), out_dtype=out_dtype, path="mma").float().cpu()
    assert rel_err(got, base.sigmoid() * mul.double()) < tol
    got = K().linear(d(x), d(w), d(b), act="relu", add=d(add), rowzero=d(rz.to(torch.uint8)), out_dtype=out_dtype,
                     path="mma").float().cpu()
    want = base.relu() + add.double()
    want[rz] = 0
    assert rel_err(got, want) < tol
    wide_in = torch.randn(M, 2 * K_, generator=g).bfloat16()
    wide_out = torch.zeros(M, 3 * N, device=DEV, dtype=out_dtype)
    K().linear(d(wide_in)[:, K_:], d(w), d(b), out=wide_out[:, N:2 * N], path="mma")
    assert rel_err(wide_out[:, N:2 * N].float().cpu(), F.linear(wide_in[:, K_:].double(), w.double(), b.double())) < tol
    assert torch.count_nonzero(wide_out[:, :N]) == 0 and torch.count_nonzero(wide_out[:, 2 * N:]) == 0


@pytest.mark.parametrize("M", [700, 20001])          # 20001 rows: the persistent kernel (314 tiles)
def test_linear_bf16_tensor_core_epilogues_and_views(M):
    g = _g(11)
    N, K_ = 256, 256
    x = torch.randn(M, K_, generator=g).bfloat16()
    w = (torch.randn(N, K_, generator=g) / 16).bfloat16()
    b = torch.randn(N, generator=g)
    mul, add = torch.randn(M, N, generator=g).bfloat16(), torch.randn(M, N, generator=g).bfloat16()
    rz = (torch.rand(M, generator=g) < 0.2)
    base = F.linear(x.double(), w.double(), b.double())
    d = lambda t: t.to(DEV)                                                                 # noqa: E731
    f = lambda t: t.float().cpu()                                                           # noqa: E731
    kw = dict(out_dtype=torch.float32, path="tc")
    assert rel_err(f(K().linear(d(x), d(w), d(b), act="relu", **kw)), base.relu()) < 1e-5
    assert rel_err(f(K().linear(d(x), d(w), d(b), act="sigmoid", mul=d(mul), **kw)), base.sigmoid() * mul.double()) < 1e-5
    assert rel_err(f(K().linear(d(x), d(w), d(b), add=d(add), **kw)), base + add.double()) < 1e-5
    got = f(K().linear(d(x), d(w), d(b), rowzero=d(rz.to(torch.uint8)), **kw))
    assert torch.count_nonzero(got[rz]) == 0 and rel_err(got[~rz], base[~rz]) < 1e-5
    wide_in = torch.randn(M, 2 * K_, generator=g).bfloat16()
    wide_out = torch.zeros(M, 3 * N, device=DEV, dtype=torch.bfloat16)
    K().linear(d(wide_in)[:, K_:], d(w), d(b), out=wide_out[:, N:2 * N], path="tc")
    assert rel_err(f(wide_out[:, N:2 * N]), F.linear(wide_in[:, K_:].double(), w.double(), b.double())) < 6e-3
    assert torch.count_nonzero(wide_out[:, :N]) == 0 and torch.count_nonzero(wide_out[:, 2 * N:]) == 0
    # CUDA-core path on bf16 data agrees with the tensor-core path
    a = f(K().linear(d(x), d(w), d(b), out_dtype=torch.float32, path="simt"))
    assert rel_err(a, base) < 1e-5


@pytest.mark.parametrize("M,Hd", [(128, 256), (400, 256), (100, 2048), (400, 2048), (22323, 2048), (1000, 1024)])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_fused_mlp2_tensor_core(M, Hd, out_dtype):
    """relu(x W1^T + b1) W2^T + b2 with the hidden activation kept on chip; checker: fp64 on the bf16-rounded operands,
    with the hidden activation rounded to bf16 exactly as the kernel does before the second GEMM."""
    g = _g(M + Hd)
    x = torch.randn(M, 256, generator=g).bfloat16()
    w1 = (torch.randn(Hd, 256, generator=g) / 16).bfloat16()
    w2 = (torch.randn(256, Hd, generator=g) / math.sqrt(Hd)).bfloat16()
    b1, b2 = torch.randn(Hd, generator=g), torch.randn(256, generator=g)
    h = F.linear(x.double(), w1.double(), b1.double()).relu().float().bfloat16().double()
    want = F.linear(h, w2.double(), b2.double())
    got = K().mlp2(x.to(DEV), w1.to(DEV), b1.to(DEV), w2.to(DEV), b2.to(DEV), out_dtype=out_dtype).float().cpu()
    # (hidden activations that sit on a bf16 rounding boundary round differently from the fp64 checker; with up to 2048
    #  of them per output this shows as 5e-5 .. 4e-4)
    assert rel_err(got, want) < (1e-3 if out_dtype == torch.float32 else 6e-3)


@pytest.mark.parametrize("tail", ["1", "0"])
@pytest.mark.parametrize("M,Hd", [(22323, 2048), (300, 2048), (700, 1024), (128 * 148 + 1, 512), (128 * 221, 256), (3379, 2048)])
def test_fused_mlp2_split_hidden(M, Hd, tail, monkeypatch):
    """Hidden-dimension split of launches with few row tiles and of the tiles beyond one round of SMs (CTA (tile, split)
    reduce-adds its partial product into the zeroed fp32 output with a TMA reduce-store); MEMOTR_MLP_TAIL=0 switches it off.
    Also into a strided view whose neighbours must stay."""
    monkeypatch.setenv("MEMOTR_MLP_TAIL", tail)
    g = _g(M + Hd)
    x = torch.randn(M, 256, generator=g).bfloat16()
    w1 = (torch.randn(Hd, 256, generator=g) / 16).bfloat16()
    w2 = (torch.randn(256, Hd, generator=g) / math.sqrt(Hd)).bfloat16()
    b1, b2 = torch.randn(Hd, generator=g), torch.randn(256, generator=g)
    h = F.linear(x.double(), w1.double(), b1.double()).relu().float().bfloat16().double()
    want = F.linear(h, w2.double(), b2.double())
    wide = torch.full((M, 768), 7.0, device=DEV)
    K().mlp2(x.to(DEV), w1.to(DEV), b1.to(DEV), w2.to(DEV), b2.to(DEV), out=wide[:, 256:512])
    assert rel_err(wide[:, 256:512].cpu(), want) < 1e-3
    assert (wide[:, :256] == 7).all() and (wide[:, 512:] == 7).all()
    again = torch.full((M, 256), float("nan"), device=DEV)          # whatever the buffer held before is overwritten
    K().mlp2(x.to(DEV), w1.to(DEV), b1.to(DEV), w2.to(DEV), b2.to(DEV), out=again)
    assert rel_err(again.cpu(), want) < 1e-3


@pytest.mark.parametrize("M,with_pos", [(300, True), (128 * 148, True), (128 * 148, False), (18000, True), (77, False)])
def test_fused_mlp2_with_layernorm_epilogue(M, with_pos):
    """memotr_mlp2_lnout == memotr_mlp2 (fp32 out) followed by memotr_layernorm with residual: y (bf16), fp32 master and
    y + pos (the three things the next encoder layer reads); also against fp64 on the bf16-rounded operands."""
    g = _g(M)
    Hd = 2048
    x = torch.randn(M, 256, generator=g).bfloat16()
    res = torch.randn(M, 256, generator=g)
    pos = torch.randn(M, 256, generator=g).bfloat16() if with_pos else None
    gamma, beta = 1 + 0.1 * torch.randn(256, generator=g), 0.1 * torch.randn(256, generator=g)
    w1 = (torch.randn(Hd, 256, generator=g) / 16).bfloat16()
    w2 = (torch.randn(256, Hd, generator=g) / math.sqrt(Hd)).bfloat16()
    b1, b2 = torch.randn(Hd, generator=g), torch.randn(256, generator=g)
    d = lambda t: t.to(DEV) if t is not None else None                                      # noqa: E731
    y, y32, ypos = K().mlp2_lnout(d(x), d(w1), d(b1), d(w2), d(b2), d(res), d(gamma), d(beta), pos=d(pos))
    two = K().mlp2(d(x), d(w1), d(b1), d(w2), d(b2), out_dtype=torch.float32)
    ry, ry32 = K().layernorm(two, d(gamma), d(beta), x2=d(res), out_dtype=torch.bfloat16, want_f32=True)
    assert rel_err(y32.cpu().numpy(), ry32.cpu().numpy()) < 5e-6        # same operands; one-pass vs two-pass moments
    assert rel_err(y.float().cpu().numpy(), ry.float().cpu().numpy()) < 8e-3   # (a bf16 ulp where the fp32 values straddle a boundary)
    h = F.linear(x.double(), w1.double(), b1.double()).relu().float().bfloat16().double()
    want = F.layer_norm(F.linear(h, w2.double(), b2.double()) + res.double(), (256,), gamma.double(), beta.double(), 1e-5)
    assert rel_err(y32.cpu().numpy(), want.numpy()) < 1e-3
    assert rel_err(y.float().cpu().numpy(), want.numpy()) < 6e-3
    if with_pos:
        assert rel_err(ypos.float().cpu().numpy(), (want + pos.double()).numpy()) < 6e-3
    else:
        assert ypos is None


@pytest.mark.parametrize("M", [300, 128 * 148, 22323, 77])
def test_encoder_dense_block_matches_the_four_ops(M):
    """memotr_encoder_dense_block (output_proj + norm1 + FFN + norm2 in one kernel per row tile) against the same four ops
    through the separate kernels, and against fp64 on the bf16-rounded GEMM operands."""
    g = _g(M + 5)
    Hd = 2048
    att = torch.randn(M, 256, generator=g).bfloat16()
    src = torch.randn(M, 256, generator=g)
    pos = torch.randn(M, 256, generator=g).bfloat16()
    wout = (torch.randn(256, 256, generator=g) / 16).bfloat16()
    w1 = (torch.randn(Hd, 256, generator=g) / 16).bfloat16()
    w2 = (torch.randn(256, Hd, generator=g) / math.sqrt(Hd)).bfloat16()
    bout, b1, b2 = torch.randn(256, generator=g), torch.randn(Hd, generator=g), torch.randn(256, generator=g)
    g1, be1 = 1 + 0.1 * torch.randn(256, generator=g), 0.1 * torch.randn(256, generator=g)
    g2, be2 = 1 + 0.1 * torch.randn(256, generator=g), 0.1 * torch.randn(256, generator=g)
    d = lambda t: t.to(DEV)                                                                 # noqa: E731
    x32, y, y32, ypos = K().encoder_dense_block(d(att), d(wout), d(bout), d(src), d(g1), d(be1), d(w1), d(b1), d(w2), d(b2),
                                                d(g2), d(be2), d(pos))
    # the same through the separate kernels
    pre = K().linear(d(att), d(wout), d(bout), out_dtype=torch.float32, path="tc")
    rx, rx32 = K().layernorm(pre, d(g1), d(be1), x2=d(src), out_dtype=torch.bfloat16, want_f32=True)
    assert rel_err(x32.cpu().numpy(), rx32.cpu().numpy()) < 5e-6
    ry, ry32, rypos = K().mlp2_lnout(rx, d(w1), d(b1), d(w2), d(b2), rx32, d(g2), d(be2), pos=d(pos))
    assert rel_err(y32.cpu().numpy(), ry32.cpu().numpy()) < 2e-3       # (x may differ by a bf16 ulp where the fp32 x straddles a boundary)
    # fp64 on the rounded operands
    x = F.layer_norm(F.linear(att.double(), wout.double(), bout.double()) + src.double(), (256,), g1.double(), be1.double(), 1e-5)
    assert rel_err(x32.cpu().numpy(), x.numpy()) < 1e-5
    xb = x32.cpu().bfloat16().double()                                  # the operand the kernel's FFN saw
    h = F.linear(xb, w1.double(), b1.double()).relu().float().bfloat16().double()
    want = F.layer_norm(F.linear(h, w2.double(), b2.double()) + x32.cpu().double(), (256,), g2.double(), be2.double(), 1e-5)
    assert rel_err(y32.cpu().numpy(), want.numpy()) < 1e-3
    assert rel_err(y.float().cpu().numpy(), want.numpy()) < 6e-3
    assert rel_err(ypos.float().cpu().numpy(), (want + pos.double()).numpy()) < 6e-3


@pytest.mark.parametrize("M", [1, 77, 128, 300, 22323])
def test_linear256_layernorm_matches_gemm_plus_layernorm(M):
    """memotr_linear256_layernorm (output_proj + norm1 in one tcgen05 kernel per row tile) against fp64 on the bf16 operands and
    against the GEMM + LayerNorm launches it replaces."""
    g = _g(M + 11)
    att = torch.randn(M, 256, generator=g).bfloat16()
    src = torch.randn(M, 256, generator=g)
    w = (torch.randn(256, 256, generator=g) / 16).bfloat16()
    b = torch.randn(256, generator=g)
    g1, be1 = 1 + 0.1 * torch.randn(256, generator=g), 0.1 * torch.randn(256, generator=g)
    d = lambda t: t.to(DEV)                                                                 # noqa: E731
    y, y32 = K().linear256_layernorm(d(att), d(w), d(b), d(src), d(g1), d(be1))
    want = F.layer_norm(F.linear(att.double(), w.double(), b.double()) + src.double(), (256,), g1.double(), be1.double(), 1e-5)
    assert rel_err(y32.cpu().numpy(), want.numpy()) < 1e-5
    assert torch.equal(y.cpu(), y32.cpu().bfloat16())
    pre = K().linear(d(att), d(w), d(b), out_dtype=torch.float32, path="tc")
    ry, ry32 = K().layernorm(pre, d(g1), d(be1), x2=d(src), out_dtype=torch.bfloat16, want_f32=True)
    assert rel_err(y32.cpu().numpy(), ry32.cpu().numpy()) < 5e-6


def test_fused_mlp2_epilogues_and_views():
    g = _g(21)
    M, Hd = 333, 256
    x = torch.randn(M, 256, generator=g).bfloat16()
    w1, w2 = (torch.randn(Hd, 256, generator=g) / 16).bfloat16(), (torch.randn(256, Hd, generator=g) / 16).bfloat16()
    b1, b2 = torch.randn(Hd, generator=g), torch.randn(256, generator=g)
    mul = torch.randn(M, 256, generator=g).bfloat16()
    h = F.linear(x.double(), w1.double(), b1.double()).relu().float().bfloat16().double()
    base = F.linear(h, w2.double(), b2.double())
    d = lambda t: t.to(DEV)                                                                 # noqa: E731
    kw = dict(out_dtype=torch.float32)
    assert rel_err(K().mlp2(d(x), d(w1), d(b1), d(w2), d(b2), act2="relu", **kw).cpu(), base.relu()) < 1e-3
    assert rel_err(K().mlp2(d(x), d(w1), d(b1), d(w2), d(b2), act2="sigmoid", mul=d(mul), **kw).cpu(),
                   base.sigmoid() * mul.double()) < 1e-3
    assert rel_err(K().mlp2(d(x), d(w1), d(b1), d(w2), d(b2), mul=d(mul), **kw).cpu(), base * mul.double()) < 1e-3
    wide = torch.zeros(M, 512, device=DEV, dtype=torch.bfloat16)                   # write into a column slice
    K().mlp2(d(x), d(w1), d(b1), d(w2), d(b2), out=wide[:, :256])
    assert rel_err(wide[:, :256].float().cpu(), base) < 6e-3 and torch.count_nonzero(wide[:, 256:]) == 0


# ------------------------------------------------------------------------------------------------ row kernels
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_layernorm(dtype):
    g = _g(3)
    M = 1001
    x, x2, pos = (torch.randn(M, 256, generator=g).to(dtype) for _ in range(3))
    gamma, beta = 1 + 0.1 * torch.randn(256, generator=g), 0.1 * torch.randn(256, generator=g)
    want = F.layer_norm(x.double() + x2.double(), (256,), gamma.double(), beta.double(), 1e-5)
    y, ypos, y32 = K().layernorm(x.to(DEV), gamma.to(DEV), beta.to(DEV), x2=x2.to(DEV), pos=pos.to(DEV), want_f32=True)
    assert rel_err(y32.cpu(), want) < 1e-5
    tol = 1e-5 if dtype == torch.float32 else 5e-3
    assert rel_err(y.float().cpu(), want) < tol
    assert rel_err(ypos.float().cpu(), y.float().cpu().double() + pos.double()) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("Nq,Nk,masked", [(400, 400, False), (300, 300, True), (100, 100, False), (5, 5, False), (7, 800, True),
                                          (401, 399, True), (33, 545, False)])
def test_mha_core(dtype, Nq, Nk, masked):
    g = _g(Nq + Nk)
    q, k, v = (torch.randn(n, 256, generator=g).to(dtype) for n in (Nq, Nk, Nk))
    kpm = None
    if masked:
        kpm = torch.zeros(Nk, dtype=torch.bool)
        kpm[-(Nk // 5):] = True
    qd, kd, vd = (t.double().view(-1, 8, 32).transpose(0, 1) for t in (q, k, v))
    logits = (qd * math.sqrt(1 / 32)) @ kd.transpose(1, 2)
    if masked:
        logits = logits.masked_fill(kpm[None, None, :], float("-inf"))
    want = (torch.softmax(logits, -1) @ vd).transpose(0, 1).reshape(Nq, 256)
    got = K().mha(q.to(DEV), k.to(DEV), v.to(DEV), 8, kpm.to(DEV) if masked else None).float().cpu()
    assert rel_err(got, want) < (1e-5 if dtype == torch.float32 else 5e-3)
    if dtype == torch.float32:       # the bf16 engine's configuration: fp32 projections in, bf16 activation out
        got = K().mha(q.to(DEV), k.to(DEV), v.to(DEV), 8, kpm.to(DEV) if masked else None,
                      out_dtype=torch.bfloat16).float().cpu()
        assert rel_err(got, want) < 5e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("Kp,shapes", [(4, synth.DANCETRACK_SHAPES), (8, synth.BDD_SHAPES), (4, synth.BDD_SHAPES_L5),
                                       (3, synth.SMALL_SHAPES)])
def test_msda_forward_ex_decode_once_kernel(dtype, Kp, shapes):
    """The engine's gather kernel (v2: points decoded once into shared memory, pre-multiplied weights) against the C
    oracle, on a column slice of a wider value buffer (the interleaved decoder value maps) with border samples."""
    from oracle import msda as omsda
    value, shp, lsi, loc, attn = synth.msda_inputs(shapes, B=1, H=8, D=32, K=Kp, Lq=333, seed=60 + Kp, border=True)
    S = value.shape[1]
    wide = torch.randn(S, 3 * 256, generator=_g(1)).to(dtype)
    wide[:, 256:512] = value.reshape(S, 256).to(dtype)
    v_used = wide[:, 256:512].float().reshape(1, S, 8, 32)
    want = omsda.forward(v_used.numpy(), shp.numpy(), lsi.numpy(), loc.numpy(), attn.numpy(), fma=True)[0]
    got = K().msda_forward_ex(wide.to(DEV)[:, 256:512], shp.to(DEV), lsi.to(DEV), loc[0].contiguous().to(DEV),
                              attn[0].contiguous().to(DEV), 8).float().cpu().numpy()
    # fp16 value map (packed-half blend, bf16 output): the output rounding to bf16 (2^-9) dominates
    assert rel_err(got, want) < (2e-6 if dtype == torch.float32 else 4e-3)


def test_msda_forward_ex_fp16_value_map_is_tighter_than_bf16():
    """Same inputs, fp16 vs bf16 value maps, each against the fp32 oracle on the UNROUNDED value: the fp16 path (more
    mantissa in the map, fp16 corner blend) must not be worse than the bf16 path."""
    from oracle import msda as omsda
    value, shp, lsi, loc, attn = synth.msda_inputs(synth.DANCETRACK_SHAPES, B=1, H=8, D=32, K=4, Lq=2000, seed=81, border=True)
    S = value.shape[1]
    v2d = (value.reshape(S, 256) * 100).contiguous()            # O(1) magnitudes like a real value_proj output
    want = omsda.forward(v2d.reshape(1, S, 8, 32).numpy(), shp.numpy(), lsi.numpy(), loc.numpy(), attn.numpy())[0]
    d = lambda t: t.to(DEV)                                                                  # noqa: E731
    err = {}
    for dt in (torch.bfloat16, torch.float16):
        got = K().msda_forward_ex(d(v2d.to(dt)), d(shp), d(lsi), d(loc[0].contiguous()), d(attn[0].contiguous()), 8)
        err[dt] = rel_err(got.float().cpu().numpy(), want)
    assert err[torch.float16] <= err[torch.bfloat16] * 1.05 and err[torch.float16] < 5e-3, err


@pytest.mark.parametrize("mode,L,Kp", [("enc", 4, 4), ("dec", 4, 4), ("enc", 4, 8), ("dec", 4, 3), ("enc", 5, 4)])
def test_msda_prep_matches_module_arithmetic(mode, L, Kp):
    """Sampling locations / attention weights against the torch expressions of ms_deform_attn.py:108-120 with the
    reference points of deformable_encoder.py:29-40 / deformable_decoder.py:82-84, padded (valid ratio < 1) case."""
    g = _g(5)
    shapes = synth.SMALL_SHAPES if L == 4 else synth.SMALL_SHAPES + ((1, 2),)
    H = 8
    S = sum(h * w for h, w in shapes)
    vr = torch.rand(1, L, 2, generator=g) * 0.3 + 0.7
    shp = torch.as_tensor(shapes, dtype=torch.long)
    lsi = torch.cat((shp.new_zeros(1), shp.prod(1).cumsum(0)[:-1]))
    Lq = S if mode == "enc" else 23
    ol = torch.randn(Lq, 3 * H * L * Kp, generator=g)
    off = ol[:, :2 * H * L * Kp].view(1, Lq, H, L, Kp, 2)
    aw = torch.softmax(ol[:, 2 * H * L * Kp:].view(1, Lq, H, L * Kp), -1).view(1, Lq, H, L, Kp)
    if mode == "enc":
        ref = oframe.encoder_reference_points(shapes, vr, "cpu")
        norm = torch.stack([shp[:, 1], shp[:, 0]], -1)
        want_loc = ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
        ref4 = None
    else:
        ref4 = torch.rand(Lq, 4, generator=g) * 0.6 + 0.2
        ref_in = ref4[None, :, None] * torch.cat([vr, vr], -1)[:, None]
        want_loc = ref_in[:, :, None, :, None, :2] + off / Kp * ref_in[:, :, None, :, None, 2:] * 0.5
    loc, attn = K().msda_prep(ol.to(DEV), shp.to(DEV), lsi.to(DEV), vr[0].contiguous().to(DEV), H, L, Kp,
                              ref4.to(DEV) if ref4 is not None else None)
    # reciprocal-multiply divisions and ex2-based exp in the fast path: a few ulp
    assert rel_err(loc.cpu(), want_loc[0]) < 2e-6
    assert rel_err(attn.cpu(), aw[0]) < 5e-6


def test_sine_embed_and_box_refine():
    g = _g(9)
    i = torch.arange(128, dtype=torch.float32)
    dim_t = 10000 ** (2 * torch.div(i, 2, rounding_mode="trunc") / 128)
    pts = torch.rand(50, 4, generator=g)

    def truth(p):
        """models/utils.py:78-85 in float64 on the very dim_t the kernel is given: independent of the host's float32 pow / sin
        (the float32 CPU oracle was seen 1.5e-4 off on one GPU-box host type; that is reported below, not asserted)."""
        e = (p.double() * float(torch.tensor(2 * math.pi, dtype=torch.float32)))[..., None] / dim_t.double()
        return torch.stack((e[..., 0::2].sin(), e[..., 1::2].cos()), dim=-1).flatten(-3)
    got = K().sine_embed(pts.to(DEV), dim_t.to(DEV)).cpu()
    assert rel_err(got, truth(pts)) < 2e-6
    host = rel_err(oframe.pos_to_pos_embed(pts, num_pos_feats=128), truth(pts))
    if host > 1e-5:
        import warnings
        warnings.warn(f"this host's float32 CPU path deviates {host:.1e} from float64 in pos_to_pos_embed")
    raw = torch.randn(50, 4, generator=g)
    scale = torch.tensor([0.9, 0.8, 0.9, 0.8])
    got = K().sine_embed(raw.to(DEV), dim_t.to(DEV), scale4=scale.to(DEV), apply_sigmoid=True).cpu()
    assert rel_err(got, truth(raw.double().sigmoid() * scale.double())) < 2e-6
    delta, ref = torch.randn(50, 4, generator=g), torch.rand(50, 4, generator=g)
    ref[0, 0], ref[1, 1] = 0.0, 1.0                       # inverse_sigmoid clamps (utils/utils.py:71-73)
    want = (delta.double() + oframe.inverse_sigmoid(ref.double())).sigmoid()      # float64 yardstick (host-independent)
    new, nxt = K().box_refine(delta.to(DEV), ref.to(DEV), 30)
    assert rel_err(new.cpu(), want) < 1e-6
    assert torch.equal(nxt[:30].cpu(), new[:30].cpu()) and torch.equal(nxt[30:].cpu(), ref[30:])


def test_linear_with_msda_prep_epilogue_and_strided_gather():
    """memotr_linear_msda_prep (persistent tcgen05 GEMM whose epilogue turns the raw offsets / logits into sampling locations
    and softmax weights) == linear followed by memotr_msda_prep; the strided gather on its rows == the dense gather."""
    g = _g(77)
    shapes = [(100, 168), (50, 84), (25, 42), (13, 21)]
    lsi = [0, 16800, 21000, 22050]
    S, H, L, Kp = 22323, 8, 4, 4
    x = torch.randn(S, 256, generator=g).bfloat16().to(DEV)
    w = (torch.randn(384, 256, generator=g) / 16).bfloat16().to(DEV)
    b = torch.randn(384, generator=g).to(DEV)
    vr = (0.8 + 0.2 * torch.rand(L, 2, generator=g)).to(DEV)
    shapes_t, lsi_t = torch.tensor(shapes, device=DEV), torch.tensor(lsi, device=DEV)
    raw = K().linear(x, w, b, out_dtype=torch.float32, path="tc")
    loc, attn = K().msda_prep(raw, shapes_t, lsi_t, vr, H, L, Kp)
    rows = K().linear_msda_prep(x, w, b, shapes, lsi, vr, H, L, Kp)
    assert rel_err(rows[:, :256].cpu().numpy(), loc.reshape(S, 256).cpu().numpy()) < 2e-6
    assert rel_err(rows[:, 256:].cpu().numpy(), attn.reshape(S, 128).cpu().numpy()) < 2e-6
    value = torch.randn(S, 256, generator=g).half().to(DEV)
    dense = K().msda_forward_ex(value, shapes_t, lsi_t, loc, attn, H)
    strided = K().msda_forward_strided(value, shapes_t, lsi_t, rows, H, L, Kp)
    assert rel_err(strided.float().cpu().numpy(), dense.float().cpu().numpy()) < 1e-2     # bf16 outputs of ~equal inputs
    cat = torch.cat((loc.reshape(S, 256), attn.reshape(S, 128)), 1).contiguous()
    strided2 = K().msda_forward_strided(value, shapes_t, lsi_t, cat, H, L, Kp)
    assert torch.equal(strided2, dense)                                                       # same inputs: bit-equal
    # the windowed gather on the rows of the projection (tests/test_msda_window_gpu.py has the thorough cases)
    win = K().msda_forward_window(value, shapes, vr, rows=rows, n_heads=H, n_points=Kp, radius=3.0)
    assert torch.equal(win, strided)


@pytest.mark.parametrize("h,w,vh,vw", [(100, 168, 100, 168), (100, 168, 88, 167), (50, 84, 44, 84), (13, 21, 12, 20), (1, 1, 1, 1)])
def test_pos_embed_sine_matches_oracle_and_reference_golden(h, w, vh, vw):
    m = torch.ones(1, h, w, dtype=torch.bool)
    m[:, :vh, :vw] = False
    want = oframe.position_embedding_sine(m)[0]
    got = K().pos_embed_sine(m[0].to(DEV)).cpu()
    # float64 restatement of models/position_embedding.py:23-43 (host-independent yardstick; the float32 oracle is reported)
    nm = (~m).double()
    y, x = nm.cumsum(1), nm.cumsum(2)
    y, x = (y - 0.5) / (y[:, -1:, :] + 1e-6) * 2 * math.pi, (x - 0.5) / (x[:, :, -1:] + 1e-6) * 2 * math.pi
    di = 20.0 ** (2 * torch.div(torch.arange(128, dtype=torch.float64), 2, rounding_mode="trunc") / 128)
    px, py = x[:, :, :, None] / di, y[:, :, :, None] / di
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).flatten(3)
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).flatten(3)
    t64 = torch.cat((py, px), dim=3).permute(0, 3, 1, 2)[0]
    valid = ~m[0]            # (fully padded rows / columns have arguments ~ -3e6: chaotic in float32, compared through the golden)
    assert got.shape == want.shape and (got.double() - t64)[:, valid].abs().max() <= 5e-6
    if (want.double() - t64)[:, valid].abs().max() > 5e-6 or (got - want).abs().max() > 2e-6:
        import warnings
        warnings.warn(f"this host's float32 CPU path deviates from the kernel / float64 in position_embedding_sine: "
                      f"{float((want.double() - t64)[:, valid].abs().max()):.1e} on valid pixels, "
                      f"{float((got - want).abs().max()):.1e} overall")
    g = np.load(os.path.join(GOLDEN, "pos_embed.npz"))            # outputs of the reference's own class
    for i in (0, 1):
        got = K().pos_embed_sine(torch.from_numpy(g[f"mask{i}"][0]).to(DEV)).cpu().numpy()
        assert np.abs(got - g[f"pos{i}"][0]).max() <= 2e-6


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_engine_position_maps_on_device_equal_uploaded_maps(mode):
    """pos_embed=dict(...): the engine rebuilds the position maps from the padding masks; same outputs as uploading them."""
    from memotr_b200.engine import FrameEngine
    g, cfg, sd, x, shapes, nt = _case("small_padded")
    # the uploaded maps come from the stand-alone kernel (pinned to the reference class / float64 above), not from the float32 CPU
    # oracle: one GPU-box host type evaluates float32 sin / cos 1.5e-4 off, which the 1e-5 comparison below would see
    pos = [K().pos_embed_sine(m[0].to(DEV)).cpu()[None] for m in x["masks"]]
    outs = []
    for on_device in (False, True):
        eng = FrameEngine(sd, cfg, shapes, nt, DEV, mode=mode, pos_embed=dict(temperature=20) if on_device else None)
        eng.load_frame(x["srcs"], x["masks"], None if on_device else pos, x["tracks"]["ref_pts"], x["tracks"]["query_embed"])
        eng.forward()
        torch.cuda.synchronize()
        outs.append({k: v.clone() for k, v in eng.results().items()})
    for k in ("pred_logits", "pred_bboxes", "outputs", "memory"):
        assert rel_err(outs[1][k].cpu().numpy(), outs[0][k].cpu().numpy()) <= (1e-5 if mode == "fp32" else 2e-2), k


def test_clip_runner_host_api_tracker_mode():
    """ClipRunner with the tracker on the device and the position maps rebuilt on the device: pinned host frames in, packed
    result rows out; same identities / boxes as stepping the engine by hand on device-resident inputs."""
    from memotr_b200.engine import ClipRunner, FrameEngine
    cfg = synth.small_cfg()
    sd = synth.hot_path_state_dict(cfg, seed=5)
    thr = dict(det_score_thresh=0.66, track_score_thresh=0.6, miss_tolerance=2, result_score_thresh=0.62)
    frames = [synth.frame_inputs(cfg, synth.SMALL_SHAPES, 0, seed=10 + t) for t in range(4)]
    kw = dict(mode="fp32", tracker=thr, pos_embed=dict(temperature=20))
    ref = FrameEngine(sd, cfg, synth.SMALL_SHAPES, 10, DEV, **kw)
    want = []
    for fr in frames:
        ref.load_frame(fr["srcs"], fr["masks"], None, ref.in_track_ref, ref.in_track_embed)
        ref.step()
        torch.cuda.synchronize()
        keep = ref.trk.res_keep.cpu().bool()
        want.append((ref.trk.res_ids.cpu()[keep].tolist(), ref.trk.res_boxes.cpu()[keep].clone()))
    eng = FrameEngine(sd, cfg, synth.SMALL_SHAPES, 10, DEV, **kw)
    fr0 = frames[0]
    eng.load_frame(fr0["srcs"], fr0["masks"], None, eng.in_track_ref, eng.in_track_embed)
    eng.capture()                               # (the warm-up step inside capture() leaves the recurrent state untouched)
    runner = ClipRunner(eng)
    pin = lambda t: t.contiguous().pin_memory()                                       # noqa: E731
    host = [([pin(t) for t in fr["srcs"]], None, [pin(t.to(torch.uint8)) for t in fr["masks"]]) for fr in frames]
    runner.prefetch(0, *host[0])
    for i in range(len(host)):
        if i + 1 < len(host):
            runner.prefetch((i + 1) % 2, *host[i + 1])
        runner.run(i % 2)
        torch.cuda.synchronize()
        ids, boxes, scores = runner.results()
        assert ids.tolist() == want[i][0], i
        assert torch.equal(boxes, want[i][1]), i
    assert runner.h2d_bytes == sum(t.numel() * 4 for t in fr0["srcs"]) + sum(t.numel() for t in fr0["masks"])
    assert runner.d2h_bytes == 10 * 29 + 4


# ------------------------------------------------------------------------------------------------ engine
def _case(tag):
    g = np.load(os.path.join(GOLDEN, f"frame_{tag}.npz"))
    meta = [int(v) for v in g["meta"]]
    n_tracks, seed_w, seed_x, padded = meta[:4]
    refinit, sine_pos = (meta[4], meta[5]) if len(meta) > 4 else (0, 0)
    shapes = [tuple(int(v) for v in r) for r in g["shapes"]]
    cfg = oframe.dancetrack_cfg() if tag.startswith("full") else synth.small_cfg()
    sd = synth.reference_init_state_dict(cfg, seed=seed_w) if refinit else synth.hot_path_state_dict(cfg, seed=seed_w)
    x = synth.frame_inputs(cfg, shapes, n_tracks, seed=seed_x, padded=bool(padded))
    if sine_pos:            # the golden was made with the reference's PositionEmbeddingSine maps: the engine rebuilds them
        x["pos"] = None     # on the device from the padding masks (pos_embed=dict(...))
    return g, cfg, sd, x, shapes, n_tracks


def _run_engine(tag, mode, debug_enc=False):
    from memotr_b200.engine import FrameEngine
    g, cfg, sd, x, shapes, nt = _case(tag)
    eng = FrameEngine(sd, cfg, shapes, nt, DEV, mode=mode, pos_embed=dict(temperature=20) if x["pos"] is None else None)
    if debug_enc:
        eng.debug_enc = []
    eng.load_frame(x["srcs"], x["masks"], x["pos"], x["tracks"]["ref_pts"], x["tracks"]["query_embed"])
    eng.forward()
    eng.load_tracks(x["tracks"])
    eng.update_tracks()
    torch.cuda.synchronize()
    return g, eng, eng.results(), eng.track_state()


FRAME_KEYS = ("pred_logits", "pred_bboxes", "last_ref_pts", "init_ref_pts", "outputs", "aux_logits", "aux_bboxes",
              "aux_queries")
UPD_KEYS = ("ref_pts", "query_embed", "long_memory", "last_output")


@pytest.mark.parametrize("tag", ["small", "small_padded", "full"])
def test_engine_fp32_matches_reference_modules(tag):
    g, eng, res, st = _run_engine(tag, "fp32")
    for k in FRAME_KEYS:
        assert rel_err(res[k].cpu().numpy(), g[k]) < 1e-4, k
    for k in UPD_KEYS:
        assert rel_err(st[k].cpu().numpy(), g["upd_" + k]) < 1e-4, k
    assert eng.launches > 0


@pytest.mark.parametrize("tag", ["small", "small_padded", "full"])
def test_engine_bf16_matches_reference_modules(tag):
    g, eng, res, st = _run_engine(tag, "bf16")
    worst = {}
    for k in FRAME_KEYS:
        worst[k] = rel_err(res[k].cpu().numpy(), g[k])
    for k in UPD_KEYS:
        worst["upd_" + k] = rel_err(st[k].cpu().numpy(), g["upd_" + k])
    print("bf16 engine rel err:", {k: f"{v:.2e}" for k, v in worst.items()})
    # The north-star bf16 tolerance (1e-2) is an op-level bound and every kernel meets it (tests above).  End to end the
    # only bf16 quantities are the GEMM operands (each dot product then carries ~2^-9/sqrt(3)*sqrt(2) ~ 1.6e-3 relative
    # noise, not averaged away because the sum is itself a random walk); everything on the residual path is fp32.
    # Through the shallow configurations (2+3 layers) that stays <= 1e-2 on geometry.  The full 6+6-layer network with
    # RANDOM weights and white-noise feature maps amplifies any perturbation (bilinear sampling of uncorrelated pixels,
    # near-one-hot attention): the fp32 engine's own 1e-7 rounding differences already grow to 1e-5..1e-4 there, and
    # the bf16 operand noise grows by the same factor to the values bounded below (measured: boxes 5e-2, embeddings and
    # logits 1.1e-1 / 1.6e-1 of the tensor range).  DESIGN.md ("Numerics") records this budget.
    deep = tag == "full"
    for k in ("pred_bboxes", "aux_bboxes", "last_ref_pts", "upd_ref_pts"):
        assert worst[k] < (1e-1 if deep else 1e-2), (k, worst[k])
    for k in ("outputs", "aux_queries", "pred_logits", "aux_logits", "upd_query_embed", "upd_long_memory", "upd_last_output"):
        assert worst[k] < (2.5e-1 if deep else 3e-2), (k, worst[k])


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_engine_matches_reference_modules_on_reference_init_weights(mode):
    """The benchmarked configuration (full DanceTrack sizes, 6+6 layers, 300+100 queries, padded frame, position maps rebuilt
    on the device, every fused kernel) on weights drawn from the reference's OWN initialisation
    (synth.reference_init_state_dict), against outputs of the reference nn.Modules (frame_full_refinit.npz): the north
    star's bars on EVERY output -- fp32 <= 1e-4, bf16 <= 1e-2.  (tests/test_numerics_cpu.py: the rounding model of the bf16
    mode predicts 7e-3 here, and shows why the white-noise weights of frame_full.npz amplify the same rounding 40x.)"""
    g, eng, res, st = _run_engine("full_refinit", mode)
    if mode == "bf16":
        assert eng.dec_cluster and eng.upd_fused and eng.fuse_prep and eng.msda_window
    worst = {k: rel_err(res[k].cpu().numpy(), g[k]) for k in FRAME_KEYS}
    worst.update({"upd_" + k: rel_err(st[k].cpu().numpy(), g["upd_" + k]) for k in UPD_KEYS})
    print(mode, "engine vs reference modules, reference-init weights:", {k: f"{v:.1e}" for k, v in worst.items()})
    tol = 1e-4 if mode == "fp32" else 1e-2
    assert max(worst.values()) < tol, worst


@pytest.mark.parametrize("tag", ["full", "full_refinit"])
def test_engine_bf16_every_layer_teacher_forced_within_1e2(tag):
    """Teacher forcing, layer by layer: the oracle's fp32 layer applied to the ENGINE's own input of that layer must agree
    with the engine's output of that layer to 1e-2 -- encoder layers, decoder layers (hidden state and refined boxes) and the
    query updater -- also on the white-noise weights whose end-to-end deviation is several 1e-2 (amplification, not kernels)."""
    g, eng, res, st = _run_engine(tag, "bf16", debug_enc=True)
    _, cfg, sd, x, shapes, nt = _case(tag)
    pos = x["pos"] if x["pos"] is not None else [oframe.position_embedding_sine(m) for m in x["masks"]]
    report = {}
    with torch.no_grad():
        src0, mask, posf, shp, lsi, vr = oframe.flatten_levels(sd, "transformer", x["srcs"], x["masks"], pos)
        ref = oframe.encoder_reference_points(shp, vr, "cpu")
        states = [t.cpu()[None] for t in eng.debug_enc]
        assert len(states) == cfg["n_enc_layers"] + 1
        assert rel_err(states[0].numpy(), src0.numpy()) < 1e-6
        for i in range(cfg["n_enc_layers"]):
            want = oframe.encoder_layer(sd, f"transformer.encoder.layers.{i}", states[i], posf, ref, shp, lsi, mask, cfg)
            report[f"enc{i}"] = rel_err(states[i + 1].numpy(), want.numpy())
        memory = states[-1]
        qm = torch.zeros((1, eng.nq), dtype=torch.bool)
        for lid in range(cfg["n_dec_layers"]):
            tgt_in, ref_in = eng.tgt32[lid].float().cpu()[None], eng.ref[lid].cpu()[None]
            want_t, want_r = oframe.decoder_step(sd, lid, tgt_in, ref_in, memory, shp, lsi, vr, qm, mask, cfg)
            report[f"dec{lid}"] = rel_err(eng.tgt32[lid + 1].float().cpu().numpy(), want_t[0].numpy())
            report[f"ref{lid}"] = rel_err(eng.ref[lid + 1].cpu().numpy(), want_r[0].numpy())
        wupd = oframe.update_tracks(sd, x["tracks"], cfg)
        for k in UPD_KEYS:
            report["upd_" + k] = rel_err(st[k].cpu().numpy(), wupd[k].numpy())
    print(tag, "teacher-forced per-layer rel err:", {k: f"{v:.1e}" for k, v in report.items()})
    assert max(report.values()) < 1e-2, report


def test_engine_bf16_white_noise_weights_vs_matched_rounding_oracle():
    """frame_full.npz: the bf16 engine against the oracle evaluated under the SAME rounding model (bf16 GEMM operands, fp16
    value maps, bf16 gather rows).  The network amplifies every rounding difference 40x (tests/test_numerics_cpu.py), so even
    the matched model only agrees to the extent the individual rounding decisions coincide; the deviations are recorded, and
    bounded by what the rounding model itself shows against fp32."""
    g, eng, res, st = _run_engine("full", "bf16")
    _, cfg, sd, x, shapes, nt = _case("full")
    with torch.no_grad(), oframe.numerics(gemm="bf16", value="fp16", gather="bf16"):
        want = oframe.frame_forward(sd, x["srcs"], x["masks"], x["pos"], x["tracks"]["ref_pts"], x["tracks"]["query_embed"], cfg)
    rep = {k: (rel_err(res[k].cpu().numpy(), want[k].numpy()), rel_err(res[k].cpu().numpy(), g[k]),
               rel_err(want[k].numpy(), g[k])) for k in ("pred_logits", "pred_bboxes", "outputs", "aux_queries")}
    rep["memory"] = (rel_err(res["memory"].cpu().numpy(), want["memory"].numpy()), float("nan"), float("nan"))
    print("engine vs matched oracle / engine vs fp32 reference / matched oracle vs fp32 reference:",
          {k: tuple(f"{v:.1e}" for v in t) for k, t in rep.items()})
    assert rep["memory"][0] < 1e-2
    for k in ("pred_bboxes",):
        assert rep[k][0] < 1e-1
    for k in ("pred_logits", "outputs", "aux_queries"):
        assert rep[k][0] < 2.5e-1


@pytest.mark.parametrize("tag", ["small", "small_padded", "full"])
def test_engine_fused_decoder_agrees_with_launch_per_op_decoder(tag, monkeypatch, fused="1"):
    """bf16 engine, decoder + heads as ONE persistent cluster kernel (csrc/decoder_cluster.cu, the default) against the same
    engine with one launch per op (MEMOTR_DEC_FUSED=0): same arithmetic classes (bf16 GEMM operands, fp32 everything
    else; the fused kernel keeps q/k/p/v of the self-attention in fp16 instead of fp32), so the two must agree with each
    other as well as each agrees with the reference modules, layer by layer (aux outputs)."""
    monkeypatch.setenv("MEMOTR_DEC_FUSED", "0")
    g, eng0, res0, st0 = _run_engine(tag, "bf16")
    assert not eng0.dec_fused and not eng0.upd_fused
    monkeypatch.setenv("MEMOTR_DEC_FUSED", fused)
    _, eng1, res1, st1 = _run_engine(tag, "bf16")
    assert eng1.dec_fused and eng1.dec_cluster and eng1.launches < eng0.launches
    assert eng1.upd_fused                        # the fused query updater rides on the cluster machinery
    deep = tag == "full"
    report = {}
    for k in FRAME_KEYS:
        a, b, ref = res1[k].cpu().numpy(), res0[k].cpu().numpy(), g[k]
        report[k] = (rel_err(a, b), rel_err(a, ref), rel_err(b, ref))
    print("fused vs per-op / fused vs ref / per-op vs ref:", {k: tuple(f"{x:.1e}" for x in v) for k, v in report.items()})
    n_l = res0["aux_queries"].shape[0]
    for l in range(n_l):        # first layers first: a structural error shows up at layer 0 with O(1) differences
        d = rel_err(res1["aux_queries"][l].cpu().numpy(), res0["aux_queries"][l].cpu().numpy())
        assert d < (2.5e-1 if deep else 3e-2), ("aux_queries layer", l, d)
    for k in ("pred_bboxes", "aux_bboxes", "last_ref_pts"):
        assert report[k][1] < (1e-1 if deep else 1e-2), (k, report[k])
    for k in ("outputs", "aux_queries", "pred_logits", "aux_logits"):
        assert report[k][1] < (2.5e-1 if deep else 3e-2), (k, report[k])
    assert rel_err(res1["init_ref_pts"].cpu().numpy(), g["init_ref_pts"]) < 1e-5
    # query updater (fused into one cluster kernel): same inputs (the golden track state), so it is
    # compared directly -- against the launch-per-op updater and against the reference module's outputs
    upd = {k: (rel_err(st1[k].cpu().numpy(), st0[k].cpu().numpy()), rel_err(st1[k].cpu().numpy(), g["upd_" + k]),
               rel_err(st0[k].cpu().numpy(), g["upd_" + k])) for k in UPD_KEYS}
    print("updater fused vs per-op / vs ref / per-op vs ref:", {k: tuple(f"{x:.1e}" for x in v) for k, v in upd.items()})
    assert upd["ref_pts"][1] < 1e-5                                   # geometry of the updater is fp32 in both
    for k in ("query_embed", "long_memory", "last_output"):
        assert upd[k][1] < (2.5e-1 if deep else 3e-2), (k, upd[k])


@pytest.mark.parametrize("ncls,nd,nt", [(3, 20, 7), (8, 33, 16)])
def test_engine_bf16_multiclass_fused_paths_match_oracle(ncls, nd, nt, monkeypatch):
    """Configurations the golden files do not cover (several classes, detect-query counts that are not a multiple of the
    16-row blocks, a track count that fills whole blocks): the bf16 engine with every fused kernel (cluster decoder, fused
    updater) against the functional oracle on the same seeded inputs, and against the launch-per-op engine."""
    from memotr_b200.engine import FrameEngine
    cfg = dict(synth.small_cfg(), num_classes=ncls, n_det_queries=nd)
    sd = synth.hot_path_state_dict(cfg, seed=11)
    x = synth.frame_inputs(cfg, synth.SMALL_SHAPES, nt, seed=12)
    with torch.no_grad():
        want = oframe.frame_forward(sd, x["srcs"], x["masks"], x["pos"], x["tracks"]["ref_pts"], x["tracks"]["query_embed"], cfg)
        wupd = oframe.update_tracks(sd, x["tracks"], cfg)
    res = {}
    for fused in ("0", "1"):
        monkeypatch.setenv("MEMOTR_DEC_FUSED", fused)
        eng = FrameEngine(sd, cfg, synth.SMALL_SHAPES, nt, DEV, mode="bf16")
        assert eng.dec_cluster == (fused == "1") and eng.upd_fused == (fused == "1")
        eng.load_frame(x["srcs"], x["masks"], x["pos"], x["tracks"]["ref_pts"], x["tracks"]["query_embed"])
        eng.forward()
        eng.load_tracks(x["tracks"])
        eng.update_tracks()
        torch.cuda.synchronize()
        res[fused] = ({k: v.clone() for k, v in eng.results().items()}, eng.track_state())
    for fused in ("0", "1"):
        out, st = res[fused]
        assert out["pred_logits"].shape == (1, nd + nt, ncls)
        for k in ("pred_bboxes", "last_ref_pts"):
            assert rel_err(out[k].cpu().numpy(), want[k].numpy()) < 1e-2, (fused, k)
        for k in ("pred_logits", "outputs", "aux_queries"):
            assert rel_err(out[k].cpu().numpy(), want[k].numpy()) < 3e-2, (fused, k)
        assert rel_err(st["ref_pts"].cpu().numpy(), wupd["ref_pts"].numpy()) < 1e-5
        for k in ("query_embed", "long_memory", "last_output"):
            assert rel_err(st[k].cpu().numpy(), wupd[k].numpy()) < 3e-2, (fused, k)


def test_engine_memory_matches_oracle_encoder_only():
    """Encoder output (the memory) of the fp32 engine vs the functional oracle, padded masks (valid ratios < 1)."""
    g, cfg, sd, x, shapes, nt = _case("small_padded")
    from memotr_b200.engine import FrameEngine
    eng = FrameEngine(sd, cfg, shapes, nt, DEV, mode="fp32")
    eng.load_frame(x["srcs"], x["masks"], x["pos"], x["tracks"]["ref_pts"], x["tracks"]["query_embed"])
    eng.forward()
    with torch.no_grad():
        want = oframe.frame_forward(sd, x["srcs"], x["masks"], x["pos"], x["tracks"]["ref_pts"],
                                    x["tracks"]["query_embed"], cfg)
    assert rel_err(eng.results()["memory"].cpu().numpy(), want["memory"].numpy()) < 1e-4


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_engine_cuda_graph_replay_equals_eager_and_chains_frames(mode):
    from memotr_b200.engine import FrameEngine
    g, cfg, sd, x, shapes, nt = _case("small")
    eager = FrameEngine(sd, cfg, shapes, nt, DEV, mode=mode)
    graph = FrameEngine(sd, cfg, shapes, nt, DEV, mode=mode)
    for e in (eager, graph):
        e.load_frame(x["srcs"], x["masks"], x["pos"], x["tracks"]["ref_pts"], x["tracks"]["query_embed"])
        e.load_tracks(x["tracks"])
    graph.capture()         # runs the step once eagerly (warm-up) and records it once; the recurrent state is preserved
    for _ in range(3):                      # three chained frames: track queries feed back through the updater
        eager.step()
        graph.replay()
    torch.cuda.synchronize()
    for k in ("pred_logits", "pred_bboxes", "outputs"):
        assert torch.equal(eager.results()[k], graph.results()[k]), k
    for k in UPD_KEYS:
        assert torch.equal(eager.st[k], graph.st[k]), k
    assert graph.graph_launches > 20


def test_single_cta_decoder_matches_cluster_decoder():
    """memotr_decoder_forward (one CTA per row block, the pipelined clip's decoder) against memotr_decoder_forward_cluster on
    the same frame: same arithmetic, different work split -- every output within fp32 re-association noise of bf16 GEMMs."""
    from memotr_b200.engine import FrameEngine
    shapes = ((64, 104), (32, 52), (16, 26), (8, 13))
    cfg = dict(synth.small_cfg(), n_det_queries=40)
    sd = synth.reference_init_state_dict(cfg, seed=3)
    x = synth.frame_inputs(cfg, shapes, 8, seed=2, padded=True)
    eng = FrameEngine(sd, cfg, shapes, 8, DEV, mode="bf16", pos_embed=dict(temperature=20), pad_tracks=16)
    assert eng.dec_cluster and eng.dec_single_ok
    eng.load_frame(x["srcs"], x["masks"], None, x["tracks"]["ref_pts"], x["tracks"]["query_embed"])
    out = []
    for single in (False, True):
        eng.dec_use_single = single
        eng.forward()
        torch.cuda.synchronize()
        out.append({k: v.clone() for k, v in eng.results().items()})
    eng.dec_use_single = False
    for k in ("pred_logits", "pred_bboxes", "outputs", "aux_logits", "aux_bboxes", "aux_queries", "last_ref_pts", "init_ref_pts"):
        assert rel_err(out[1][k].float().cpu().numpy(), out[0][k].float().cpu().numpy()) < 2e-3, k


def test_sm_budget_changes_grids_not_results():
    """memotr_set_sm_budget (the persistent kernels size their grids for fewer SMs: frame pipelining) -- the encoder under a
    budget of 100 SMs against the unrestricted one: the tile-to-CTA assignment changes, the arithmetic does not (the FFN's
    split-K tail changes its reduction order: compared to 1e-5)."""
    from memotr_b200.engine import FrameEngine
    shapes = ((128, 168), (64, 84), (32, 42), (16, 21))          # 28560 rows: more row tiles than SMs, as at the DanceTrack size
    cfg = dict(synth.small_cfg(), n_det_queries=40)
    sd = synth.reference_init_state_dict(cfg, seed=3)
    x = synth.frame_inputs(cfg, shapes, 8, seed=2, padded=True)
    eng = FrameEngine(sd, cfg, shapes, 8, DEV, mode="bf16", pos_embed=dict(temperature=20))
    eng.load_frame(x["srcs"], x["masks"], None, x["tracks"]["ref_pts"], x["tracks"]["query_embed"])
    out = []
    for plan in (None, (100, 99), (37, 1.5)):
        eng.encode(sm_plan=plan)
        torch.cuda.synchronize()
        out.append((eng.src32.clone(), eng.value_all.clone()))
    for a, b in out[1:]:
        assert rel_err(a.float().cpu().numpy(), out[0][0].float().cpu().numpy()) < 1e-5
        assert rel_err(b.float().cpu().numpy(), out[0][1].float().cpu().numpy()) < 2e-3      # (fp16 maps of slightly different inputs)
    assert _lib_budget() == 0


def _lib_budget():
    """The budget is reset when encode() returns: a launch made now sizes its grid for all SMs (set_sm_budget(0) is idempotent)."""
    from memotr_b200 import _lib
    _lib.check(_lib.lib().memotr_set_sm_budget(0), "set_sm_budget")
    return 0


@pytest.mark.parametrize("M,N,K,act", [(22323, 256, 256, None), (20000, 384, 256, None), (19000, 2048, 256, "relu"), (19000, 256, 2048, None),
                                       (400, 256, 512, "relu"), (100, 192, 256, None)])
def test_linear_f32x3_is_fp32_accurate(M, N, K, act):
    """memotr_linear_f32x3 (two-term fp16 operand splits, three products, fp32 accumulation on the tensor cores) against fp64:
    as accurate as an fp32 GEMM (the reference's contract: TF32 off), three orders below the bf16 path."""
    g = _g(M + N)
    x = torch.randn(M, K, generator=g) * 3
    w = torch.randn(N, K, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g)
    rz = (torch.rand(M, generator=g) < 0.01).to(torch.uint8)
    got = K_().linear_f32x3(x.to(DEV), K_().pack_w3(w.to(DEV)), b.to(DEV), act=act, rowzero=rz.to(DEV)).cpu()
    want = F.linear(x.double(), w.double(), b.double())
    if act == "relu":
        want = want.relu()
    want[rz.bool()] = 0
    fp32 = F.linear(x, w, b)
    if act == "relu":
        fp32 = fp32.relu()
    fp32[rz.bool()] = 0
    e_tc, e_32 = rel_err(got.numpy(), want.numpy()), rel_err(fp32.numpy(), want.numpy())
    print(f"f32x3 vs fp64 {e_tc:.2e}; torch fp32 CPU vs fp64 {e_32:.2e}")
    # (the tensor cores accumulate with truncation: the error grows linearly with the length of the accumulation chain --
    #  1.7e-6 at 3 x 256, 1.3e-5 at 3 x 2048 -- where an IEEE fp32 dot product grows with its square root)
    assert e_tc < 2e-6 * max(1, K // 256)


K_ = K


def test_engine_fp32tc_matches_reference_modules_full_size():
    """mode="fp32tc": the fp32 engine with the encoder's GEMMs on the tensor cores at fp32 accuracy -- the full DanceTrack
    configuration against the reference modules' outputs, reference-init weights: the fp32 bar (1e-4) on every output."""
    worst = {}
    for tag in ("full_refinit", "full"):
        g, eng, res, st = _run_engine(tag, "fp32tc")
        assert eng.tc3 and eng.launches > 0
        err = {k: rel_err(res[k].cpu().numpy(), g[k]) for k in FRAME_KEYS}
        err.update({"upd_" + k: rel_err(st[k].cpu().numpy(), g["upd_" + k]) for k in UPD_KEYS})
        print(f"fp32tc engine vs reference modules ({tag}):", {k: f"{v:.1e}" for k, v in err.items()})
        worst[tag] = max(err.values())
    assert worst["full_refinit"] < 1e-4                    # measured 1.3e-5
    # white-noise weights amplify every rounding difference ~100x: the tensor cores' truncating
    # accumulation (1.3e-5 on the 2048-long dot products of linear2, against 6e-7 for an IEEE fp32 chain) shows as 1.4e-3 here,
    # where the CUDA-core fp32 engine stays below 1e-4.  Asserted so that a regression is seen, not as a parity claim.
    assert worst["full"] < 5e-3


def test_linear_f32x3_split_output_chains_into_the_next_gemm():
    """linear1 -> relu -> linear2 with linear1's epilogue writing the split fp16 operand [hi | hi | lo] of linear2 directly (the
    fp32tc FFN: no fp32 hidden tensor in HBM) against fp64 and against the two-step path through an fp32 hidden tensor."""
    g = _g(77)
    M, C, Hd = 20000, 256, 1024
    x = torch.randn(M, C, generator=g)
    w1, b1 = torch.randn(Hd, C, generator=g) / 16, torch.randn(Hd, generator=g)
    w2, b2 = torch.randn(C, Hd, generator=g) / 32, torch.randn(C, generator=g)
    d = lambda t: t.to(DEV)                                                                 # noqa: E731
    w31, w32 = K().pack_w3(d(w1)), K().pack_w3(d(w2))
    h3 = K().linear_f32x3(d(x), w31, d(b1), act="relu", split_out=True)
    assert h3.dtype == torch.float16 and tuple(h3.shape) == (M, 3 * Hd)
    assert torch.equal(h3[:, :Hd], h3[:, Hd:2 * Hd])
    got = K().linear_f32x3(h3, w32, d(b2)).cpu()
    two = K().linear_f32x3(K().linear_f32x3(d(x), w31, d(b1), act="relu"), w32, d(b2)).cpu()
    want = F.linear(F.linear(x.double(), w1.double(), b1.double()).relu(), w2.double(), b2.double())
    assert rel_err(got.numpy(), want.numpy()) < 1e-5
    assert rel_err(got.numpy(), two.numpy()) < 2e-6
