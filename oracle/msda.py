"""ctypes/numpy front-end of the plain-C MSDA oracle (oracle/msda_oracle.c).  TEST INFRASTRUCTURE ONLY.

forward(...)  restates ms_deformable_im2col_gpu_kernel   (/root/reference/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299)
backward(...) restates the col2im kernels                 (same file, :301-403 and :87-159)
Argument order and layouts are those of MSDA.ms_deform_attn_forward/backward
(/root/reference/models/ops/src/ms_deform_attn.h:20-61).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmsda_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, -ffp-contract=off).  Called by __graft_entry__.build()."""
    src = os.path.join(_HERE, "msda_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libmsda_oracle.so"])
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _prep(value, shapes, lsi, loc, attn):
    dt = np.float64 if value.dtype == np.float64 else np.float32
    value = np.ascontiguousarray(value, dtype=dt)
    loc = np.ascontiguousarray(loc, dtype=dt)
    attn = np.ascontiguousarray(attn, dtype=dt)
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    lsi = np.ascontiguousarray(lsi, dtype=np.int64)
    B, S, H, D = value.shape
    _, Lq, _, L, K, _ = loc.shape
    assert attn.shape == (B, Lq, H, L, K) and shapes.shape == (L, 2) and lsi.shape == (L,)
    assert int((shapes[:, 0] * shapes[:, 1]).sum()) == S
    return dt, value, shapes, lsi, loc, attn, (B, S, H, D, L, Lq, K)


def forward(value, shapes, lsi, loc, attn, fma: bool = True):
    """-> (B, Lq, H*D) array.  fma=True reproduces the reference kernel's contracted rounding bit for bit."""
    dt, value, shapes, lsi, loc, attn, (B, S, H, D, L, Lq, K) = _prep(value, shapes, lsi, loc, attn)
    out = np.empty((B, Lq, H * D), dtype=dt)
    fn = getattr(_load(), "msda_oracle_forward_f64" if dt == np.float64 else "msda_oracle_forward_f32")
    fn(_ptr(value), _ptr(shapes), _ptr(lsi), _ptr(loc), _ptr(attn),
       *(ctypes.c_int(x) for x in (B, S, H, D, L, Lq, K)), ctypes.c_int(int(fma)), _ptr(out))
    return out


def backward(value, shapes, lsi, loc, attn, grad_out):
    """-> (grad_value, grad_sampling_loc, grad_attn_weight), shaped like value / loc / attn."""
    dt, value, shapes, lsi, loc, attn, (B, S, H, D, L, Lq, K) = _prep(value, shapes, lsi, loc, attn)
    grad_out = np.ascontiguousarray(grad_out, dtype=dt).reshape(B, Lq, H * D)
    gv, gl, ga = np.empty_like(value), np.empty_like(loc), np.empty_like(attn)
    fn = getattr(_load(), "msda_oracle_backward_f64" if dt == np.float64 else "msda_oracle_backward_f32")
    fn(_ptr(value), _ptr(shapes), _ptr(lsi), _ptr(loc), _ptr(attn), _ptr(grad_out),
       *(ctypes.c_int(x) for x in (B, S, H, D, L, Lq, K)), _ptr(gv), _ptr(gl), _ptr(ga))
    return gv, gl, ga


def level_start_index(shapes):
    shapes = np.asarray(shapes, dtype=np.int64)
    return np.concatenate(([0], np.cumsum(shapes[:, 0] * shapes[:, 1])[:-1])).astype(np.int64)
