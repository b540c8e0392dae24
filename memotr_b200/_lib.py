"""memotr_b200/_lib.py -- ctypes binding of the C ABI (include/memotr_b200.h -> csrc/libmemotr_b200.so).

There is no CPU fallback anywhere in this package: if the shared library is missing or a call fails, a
RuntimeError is raised (the reference raises the same exception type from AT_ASSERTM/AT_ERROR,
/root/reference/models/ops/src/cuda/ms_deform_attn_cuda.cu:28-52, src/ms_deform_attn.h:35-38).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmemotr_b200.so")
ABI_VERSION = 2

F32, F64, BF16, F16 = 0, 1, 2, 3
_DTYPES = {torch.float32: F32, torch.float64: F64, torch.bfloat16: BF16, torch.float16: F16}

_lib = None

_vp, _i, _f, _l = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_long


class TrackTable(ctypes.Structure):
    """memotr_track_table (include/memotr_b200.h): device pointers of one fixed-capacity track table."""
    _fields_ = [(k, ctypes.c_void_p) for k in ("ids", "labels", "disappear_time", "query_embed", "output_embed",
                                               "last_output", "long_memory", "ref_pts", "boxes", "logits", "n_active")]


class FrameOutputs(ctypes.Structure):
    """memotr_frame_outputs (include/memotr_b200.h)."""
    _fields_ = [(k, ctypes.c_void_p) for k in ("pred_logits", "pred_boxes", "outputs", "last_ref_pts", "aux_queries")]


MEMOTR_DEC_MAX_LAYERS = 8


class DecGemm(ctypes.Structure):
    """memotr_dec_gemm (include/memotr_b200.h)."""
    _fields_ = [("W", ctypes.c_void_p), ("ldw", ctypes.c_int), ("N", ctypes.c_int), ("K", ctypes.c_int),
                ("pad_", ctypes.c_int)]


class DecLayer(ctypes.Structure):
    """memotr_dec_layer (include/memotr_b200.h)."""
    _fields_ = [(k, ctypes.c_void_p) for k in (
        "qk_b", "v_b", "sao_b", "ol_b", "cao_b", "f1_b", "f2_b", "bb0_b", "bb1_b", "bb2_b", "cls_b",
        "n1_g", "n1_b", "n2_g", "n2_b", "n3_g", "n3_b", "bb2_w", "cls_w", "value",
        "tgt_out", "ref_out", "pred_box", "pred_logit")]


class DecParams(ctypes.Structure):
    """memotr_dec_params (include/memotr_b200.h)."""
    _fields_ = ([("prog", ctypes.c_void_p)] +
                [(k, ctypes.c_int) for k in ("n_prog", "n_layers", "nq", "nd", "merge", "ncls", "n_levels", "n_points",
                                             "d_ffn", "value_stride", "np", "pad_")] +
                [(k, ctypes.c_void_p) for k in ("rph0_b", "rph1_b", "qs0_b", "qs1_b", "tgt_in", "ref_in", "vr_scale4",
                                                "valid_ratios", "dim_t", "query_pad", "kbuf", "vbuf", "barrier", "prof",
                                                "init_ref_out", "last_ref_out")] +
                [("shapes", ctypes.c_int * 16), ("lsi", ctypes.c_int * 8),
                 ("layers", DecLayer * MEMOTR_DEC_MAX_LAYERS)])


class UpdParams(ctypes.Structure):
    """memotr_upd_params (include/memotr_b200.h)."""
    _fields_ = ([("prog", ctypes.c_void_p)] +
                [(k, ctypes.c_int) for k in ("n_prog", "nt", "ncls", "d_ffn", "np", "pad_")] +
                [("update_thresh", ctypes.c_float), ("long_memory_lambda", ctypes.c_float)] +
                [(k, ctypes.c_void_p) for k in (
                    "conf0_b", "conf1_b", "fus0_b", "fus1_b", "ph0_b", "ph1_b", "q_b", "k_b", "v_b", "out_b", "mf1_b", "mf2_b",
                    "ff1_b", "ff2_b", "mn_g", "mn_b", "mfn_g", "mfn_b", "fn_g", "fn_b", "ffn_g", "ffn_b", "dim_t", "track_pad",
                    "logits", "boxes", "output_embed", "ref_pts", "query_embed", "long_memory", "last_output",
                    "feedback_ref", "feedback_embed", "kbuf", "vbuf", "barrier")])


_SIGNATURES = {
    "memotr_abi_version": ([], _i),
    "memotr_last_error": ([], ctypes.c_char_p),
    "memotr_msda_forward": ([_vp] * 6 + [_i] * 8 + [_vp], _i),
    "memotr_msda_backward": ([_vp] * 9 + [_i] * 8 + [_vp], _i),
    "memotr_msda_forward_ex": ([_vp, _i] + [_vp] * 5 + [_i] * 7 + [_vp], _i),
    "memotr_msda_forward_strided": ([_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _vp] + [_i] * 6 + [_vp], _i),
    "memotr_linear_msda_prep": ([_vp, _i, _vp, _i, _vp, _vp, _i] + [_i] * 5 + [_vp, _vp, _vp, _vp], _i),
    "memotr_msda_forward_window": ([_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _f, _i, _vp, _vp, _i, _i, _i, _i, _vp], _i),
    "memotr_msda_window_plan": ([_vp, _vp, _i, _i, _i, _i, _f, _i, _vp], _i),
    "memotr_msda_prep": ([_vp, _i] + [_vp] * 4 + [_i] + [_vp] * 2 + [_i] * 4 + [_vp], _i),
    "memotr_linear": ([_vp, _i, _vp, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _i] + [_i] * 7 + [_vp], _i),
    "memotr_mlp2": ([_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i] + [_i] * 6 + [_vp], _i),
    "memotr_mlp2_lnout": ([_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _f, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _vp], _i),
    "memotr_encoder_dense_block": ([_vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _i,
                                    _vp, _i, _vp, _i, _i, _i, _f, _vp], _i),
    "memotr_mlp2_debug_stamps": ([_vp], _i),
    "memotr_gemm_debug_stamps": ([_vp], _i),
    "memotr_linear256_layernorm": ([_vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _f, _vp, _i, _vp, _i, _i, _vp], _i),
    "memotr_layernorm": ([_vp, _i, _i, _vp, _i, _vp, _vp, _f, _vp, _i, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _vp], _i),
    "memotr_mha": ([_vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i] + [_i] * 6 + [_vp], _i),
    "memotr_tokens_from_nchw": ([_vp] * 7 + [_i] * 5 + [_vp], _i),
    "memotr_tokens_from_nchw_pe": ([_vp, _vp, _i, _i, _vp, _f, _vp] + [_vp] * 5 + [_i] * 4 + [_vp], _i),
    "memotr_pos_cumsum_levels": ([_vp, _vp, _vp, _i, _f, _vp, _vp, _vp], _i),
    "memotr_tokens_from_nchw_emb": ([_vp] * 8 + [_i] * 5 + [_vp], _i),
    "memotr_tokens_from_nchw_levels": ([_vp] * 8 + [_i, _vp, _vp, _i, _i, _i, _vp], _i),
    "memotr_valid_ratio": ([_vp, _i, _i, _vp, _vp], _i),
    "memotr_pos_embed_sine": ([_vp, _i, _i, _vp, _i, _f, _vp, _vp, _vp], _i),
    "memotr_sine_embed": ([_vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _vp], _i),
    "memotr_add": ([_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _vp], _i),
    "memotr_convert": ([_vp, _i, _i, _vp, _i, _i, _i, _i, _vp], _i),
    "memotr_box_refine": ([_vp] * 4 + [_i, _i, _vp], _i),
    "memotr_unary": ([_vp, _vp, _l, _i, _vp], _i),
    "memotr_upd_prepare": ([_vp, _i, _vp, _vp, _f, _vp, _vp, _i, _vp], _i),
    "memotr_upd_finalize": ([_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _f, _i, _i, _vp], _i),
    "memotr_tracker_update": ([_vp, _i, _i, _i, _vp, _vp, _i, _f, _f, _i, _vp, _vp, _vp, _vp, _vp], _i),
    "memotr_tracker_results": ([_vp, _i, _i, _f, _f, _f, _f, _vp, _vp, _vp, _vp, _vp], _i),
    "memotr_decoder_forward_cluster": ([_vp, _vp], _i),
    "memotr_decoder_forward": ([_vp, _vp], _i),
    "memotr_set_sm_budget": ([_i], _i),
    "memotr_conv_gemm": ([_vp, _vp, _vp, _vp, _i, _i, _i, _vp], _i),
    "memotr_im2col_3x3s2": ([_vp, _i, _i, _i, _vp, _vp], _i),
    "memotr_groupnorm_cm": ([_vp, _vp, _vp, _i, _i, _i, _f, _vp], _i),
    "memotr_linear_f32x3": ([_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp], _i),
    "memotr_updater_forward_cluster": ([_vp, _vp], _i),
    "memotr_timer_create": ([_i], _vp),
    "memotr_timer_destroy": ([_vp], None),
    "memotr_timer_record": ([_vp, _i, _vp], _i),
    "memotr_timer_elapsed_ms": ([_vp, _i, _i, _vp], _i),
}


def exported_symbols():
    """Every symbol include/memotr_b200.h declares (checked by tests/test_abi_cpu.py against the header)."""
    return sorted(_SIGNATURES)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"memotr_b200: CUDA library {LIB_PATH} is missing -- run `python -c 'import __graft_entry__ as g; "
                f"g.build()'` (or `python memotr_b200/build.py`).  There is no CPU fallback.")
        l = ctypes.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(l, name)           # AttributeError here = header/library mismatch: fail loudly
            fn.argtypes, fn.restype = argtypes, restype
        if l.memotr_abi_version() != ABI_VERSION:
            raise RuntimeError(f"memotr_b200: ABI version {l.memotr_abi_version()} != expected {ABI_VERSION}")
        _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().memotr_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise RuntimeError(f"memotr_b200: unsupported dtype {t.dtype}") from None


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(**tensors):
    """The reference's argument checks (ms_deform_attn_cuda.cu:28-38, ms_deform_attn.h:35-38)."""
    for name, t in tensors.items():
        if not t.is_cuda:
            raise RuntimeError("Not implemented on the CPU" if name == "value" else f"{name} must be a CUDA tensor")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")
