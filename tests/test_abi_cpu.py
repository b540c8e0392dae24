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
