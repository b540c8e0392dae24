"""Host logic of the frame pipeline and of the fp32-accurate GEMM's operand packing (CPU, no GPU, no library)."""
import types

import pytest
import torch


class _Graph:
    def __init__(self, log, name):
        self.log, self.name = log, name

    def replay(self):
        self.log.append(self.name)


def _simulate(log):
    """Replay the logged graph sequence on a model of the double-buffered hand-off: ('feed', c, j) loads inputs, 'first' encodes
    into set 0, 'pipe<p>' = tail(set p) || encode -> set 1 - p, 'last<p>' = tail(set p).  Returns the order in which frames
    reach their tail and asserts that every tail finds its own frame in the set it reads."""
    pending_input, sets, tails = None, [None, None], []
    for ev in log:
        if isinstance(ev, tuple):
            if ev[0] == "feed":
                pending_input = ev[1:]
            else:
                tails.append(ev)
        elif ev == "first":
            sets[0], pending_input = pending_input, None
        elif ev.startswith("pipe"):
            p = int(ev[4:])
            assert sets[p] is not None, "tail reads an empty set"
            tails.append(sets[p])
            assert pending_input is not None, "encoder runs without a freshly fed frame"
            sets[p], sets[1 - p], pending_input = None, pending_input, None
        elif ev.startswith("last"):
            p = int(ev[4:])
            assert sets[p] is not None
            tails.append(sets[p])
            sets[p] = None
    assert sets == [None, None]
    return tails


def _stub(log):
    return types.SimpleNamespace(g_first=_Graph(log, "first"), g_pipe=[_Graph(log, "pipe0"), _Graph(log, "pipe1")],
                                 g_last=[_Graph(log, "last0"), _Graph(log, "last1")])


@pytest.mark.parametrize("n", [1, 2, 3, 8, 64])
def test_run_clip_pipelined_visits_every_frame_once_in_order(n):
    from memotr_b200.engine import FrameEngine
    log = []
    FrameEngine.run_clip_pipelined(_stub(log), n, lambda j: log.append(("feed", 0, j)))
    assert _simulate(log) == [(0, j) for j in range(n)]
    assert sum(1 for e in log if e == "first") == 1 and sum(1 for e in log if isinstance(e, str) and e.startswith("last")) == 1


@pytest.mark.parametrize("lens", [[5, 1, 4], [1, 1, 1], [8, 8, 7], [3], [2, 0, 2]])
def test_run_clips_pipelined_is_one_stream_with_the_clip_boundaries_in_place(lens):
    """Every frame of every clip reaches its tail exactly once and in order; between(c) comes after the last tail of clip c has
    been enqueued and before the first tail of clip c + 1 (where the exchange and the track reset belong); only one encode-only
    and one tail-only graph for the whole stream."""
    from memotr_b200.engine import FrameEngine
    log = []
    FrameEngine.run_clips_pipelined(_stub(log), lens, lambda c, j: log.append(("feed", c, j)), lambda c: log.append(("between", c)))
    want = []
    for c, n in enumerate(lens):
        if n > 0:
            want += [(c, j) for j in range(n)] + [("between", c)]
    assert _simulate(log) == want
    assert sum(1 for e in log if e == "first") == 1 and sum(1 for e in log if isinstance(e, str) and e.startswith("last")) == 1


def test_pack_w3_reconstructs_the_scaled_weight_to_22_bits():
    from memotr_b200.kernels import W3_SHIFT, pack_w3
    g = torch.Generator().manual_seed(0)
    w = torch.randn(192, 256, generator=g) / 16
    w3 = pack_w3(w)
    assert w3.dtype == torch.float16 and tuple(w3.shape) == (192, 768)
    hi, lo, hi2 = w3[:, :256], w3[:, 256:512], w3[:, 512:]
    assert torch.equal(hi, hi2)
    rec = (hi.double() + lo.double()) / 2 ** W3_SHIFT
    rel = ((rec - w.double()).abs() / w.double().abs().clamp_min(1e-3)).max()
    assert rel < 2.0 ** -20
    assert (lo.float().abs() > 0).float().mean() > 0.9 and lo.float().abs().max() < 2.0 ** -9 * hi.float().abs().max() * 4


def test_split_operand_layout_reproduces_the_oracle_fp16x3_model():
    """The data layout memotr_linear_f32x3 multiplies -- A3 = [x_hi | x_hi | x_lo], W3 = pack_w3(w) = [w_hi | w_lo | w_hi] of 2^6 w,
    one GEMM over 3K, the scale taken out of the result -- is the "fp16x3" rounding model of oracle/frame.py (the model the
    full-size parity estimate of DESIGN.md section 6 was made with), and both are fp32-accurate."""
    from memotr_b200.kernels import W3_SHIFT, pack_w3
    from oracle import frame as oframe
    g = torch.Generator().manual_seed(1)
    x = torch.randn(64, 256, generator=g) * 2
    w = torch.randn(128, 256, generator=g) / 16
    hi = x.half()
    lo = (x - hi.float()).half()
    a3 = torch.cat((hi, hi, lo), dim=1).double()
    got = (a3 @ pack_w3(w).double().T) / 2 ** W3_SHIFT
    model = oframe._rounded_matmul(x, w, "fp16x3").double()
    exact = x.double() @ w.double().T
    scale = exact.abs().max()
    assert (got - model).abs().max() / scale < 2e-6          # the same three products; the model sums them in fp32
    assert (got - exact).abs().max() / scale < 5e-7          # the dropped lo x lo term and the fp16 rounding of the lo halves
