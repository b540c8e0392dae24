"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports exactly the symbols
include/memotr_b200.h declares; CPU tensors are rejected the way the reference rejects them.  No kernel runs here."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "memotr_b200.h")).read()
    return sorted(set(re.findall(r"MEMOTR_API\s+[\w\s\*]+?\b(memotr_\w+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from memotr_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    declared = _declared()
    assert declared, "header parse failed"
    lib = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/memotr_b200.h but not exported"
    assert sorted(_lib.exported_symbols()) == declared, "Python binding table out of sync with the header"
    assert _lib.lib().memotr_abi_version() == _lib.ABI_VERSION


def test_cpu_tensors_are_rejected_like_the_reference():
    import memotr_b200
    args = (torch.zeros(1, 4, 1, 4), torch.tensor([[2, 2]]), torch.tensor([0]), torch.zeros(1, 1, 1, 1, 1, 2),
            torch.zeros(1, 1, 1, 1, 1))
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):      # src/ms_deform_attn.h:38
        memotr_b200.ms_deform_attn_forward(*args, 64)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        memotr_b200.ms_deform_attn_backward(*args, torch.zeros(1, 1, 4), 64)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        memotr_b200.MSDeformAttnFunction.apply(*args, 64)


def test_dropin_module_name_and_callables():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "memotr_b200", "dropin"))
    try:
        sys.modules.pop("MultiScaleDeformableAttention", None)
        import MultiScaleDeformableAttention as MSDA
        assert callable(MSDA.ms_deform_attn_forward) and callable(MSDA.ms_deform_attn_backward)
    finally:
        sys.path.pop(0)
        sys.modules.pop("MultiScaleDeformableAttention", None)


def test_module_state_dict_keys_match_reference_table():
    """Parameter names/shapes of MSDeformAttn as listed in SURVEY.md 8a / oracle.synth.hot_path_param_shapes
    (that table is asserted against the instantiated reference modules by oracle/make_golden.py)."""
    from memotr_b200.ms_deform_attn import MSDeformAttn
    from oracle import synth
    cfg = synth.small_cfg()
    want = {k.split("self_attn.")[1]: v for k, v in synth.hot_path_param_shapes(cfg).items()
            if k.startswith("transformer.encoder.layers.0.self_attn.")}
    have = {k: tuple(v.shape) for k, v in MSDeformAttn(256, 4, 8, 4).state_dict().items()}
    assert have == want


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """The parameter blocks passed by value / by pointer across the C ABI: sizes and a few field offsets as a C compiler
    sees include/memotr_b200.h must equal the ctypes mirrors in memotr_b200/_lib.py (a silent drift would corrupt pointers)."""
    import subprocess
    from memotr_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "memotr_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(memotr_track_table), sizeof(memotr_frame_outputs), sizeof(memotr_dec_gemm),
         sizeof(memotr_dec_layer), sizeof(memotr_dec_params), sizeof(memotr_upd_params));
  printf("%zu %zu %zu %zu %zu\n", offsetof(memotr_dec_params, rph0_b), offsetof(memotr_dec_params, shapes),
         offsetof(memotr_dec_params, layers), offsetof(memotr_upd_params, update_thresh), offsetof(memotr_upd_params, kbuf));
  return 0;
}
''')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes, offs = (tuple(int(v) for v in line.split()) for line in subprocess.check_output([str(exe)], text=True).splitlines())
    mirrors = (_lib.TrackTable, _lib.FrameOutputs, _lib.DecGemm, _lib.DecLayer, _lib.DecParams, _lib.UpdParams)
    assert sizes == tuple(ctypes.sizeof(m) for m in mirrors)
    assert offs == (_lib.DecParams.rph0_b.offset, _lib.DecParams.shapes.offset, _lib.DecParams.layers.offset,
                    _lib.UpdParams.update_thresh.offset, _lib.UpdParams.kbuf.offset)


def test_engine_input_layout_is_one_contiguous_buffer():
    """FrameEngine.input_layout / input_views (the flat staging buffer of ClipRunner): views tile the buffer without overlap,
    masks are consecutive (the engine reads them as one (S,) padding mask), with and without uploaded position maps."""
    from memotr_b200.engine import FrameEngine
    shapes, C = [(5, 7), (3, 4), (2, 2)], 8
    for with_pos in (True, False):
        lay = FrameEngine.input_layout(shapes, C, with_pos=with_pos)
        flat = torch.zeros(lay["bytes"], dtype=torch.uint8)
        src, pos, mask = FrameEngine.input_views(flat, lay, shapes, C)
        assert (pos is None) == (not with_pos) and len(src) == len(mask) == 3
        for i, t in enumerate(src + (pos or []) + mask):
            t.reshape(-1).view(torch.uint8).fill_(i + 1)        # tag every BYTE of the view
        used = sum(t.numel() * t.element_size() for t in src + (pos or []) + mask)
        assert used <= lay["bytes"] < used + 16 and int((flat != 0).sum()) == used      # no overlap, no gaps but the tail pad
        S = sum(h * w for h, w in shapes)
        m0 = lay["mask"][0]
        assert [int(v) for v in flat[m0:m0 + S].unique()] == sorted({int(m[0]) for m in mask})
        assert all(int(flat[lay["mask"][l]]) == int(mask[l][0]) for l in range(3))


def test_window_plan_host_logic():
    """The staging plan of the windowed encoder gather (host-only code of csrc/msda_window.cu): units cover every query of
    the staged levels exactly once, the rest goes to the global-memory CTAs, shared memory stays within two CTAs per SM."""
    from memotr_b200 import kernels
    from memotr_b200 import synthetic as synth
    for shapes, K, radius in ((synth.DANCETRACK_SHAPES, 4, 2.5), (synth.DANCETRACK_SHAPES, 4, 5.75), (synth.BDD_SHAPES_L5, 8, 3.0),
                              (synth.BDD_SHAPES, 16, 2.0), (synth.SMALL_SHAPES, 4, 2.0)):
        plan = kernels.window_plan(shapes, 8, K, radius)
        sizes = [h * w for h, w in shapes]
        assert 1 <= plan["classes"] <= 4 and plan["smem"] <= 224 * 1024
        assert plan["global_q0"] == sum(sizes[:plan["classes"]])
        assert plan["global_ctas"] == -(-(sum(sizes) - plan["global_q0"]) * 8 * 4 // 768)
        units = 0
        for c in plan["cls"]:
            h, w = shapes[c["level"]]
            assert c["tiles"] == (-(-w // c["tile"][0]), -(-h // c["tile"][1])) and c["units"] == c["tiles"][0] * c["tiles"][1] * 8
            assert c["tile"][0] * c["tile"][1] % 32 == 0 and c["rec_stride"] % 128 == 32
            assert c["tma_bytes"] == sum(a * b * 64 for a, b in zip(c["ww"], c["wh"])) > 0
            for l, (a, b) in enumerate(zip(c["ww"], c["wh"])):
                assert (a == 0) == (b == 0) and a <= shapes[l][1] + 2 and b <= shapes[l][0] + 2
            units += c["units"]
        assert units == plan["units"]
    # a radius that does not fit drops the finest window first instead of failing
    big = kernels.window_plan(synth.DANCETRACK_SHAPES, 8, 4, 6.0)
    small = kernels.window_plan(synth.DANCETRACK_SHAPES, 8, 4, 2.5)
    assert all(small["cls"][0]["ww"]) and big["cls"][0]["ww"][0] == 0 and all(big["cls"][0]["ww"][1:])
