"""memotr_b200/kernels.py -- tensor-level wrappers of the individual C-ABI kernels (allocation + argument plumbing).

The frame engine calls the library directly on its pre-allocated workspace; these functional forms exist for the
module mirrors, the unit tests and ad-hoc use.  All tensors must be CUDA tensors; nothing here falls back to torch math.
"""
import torch

from . import _lib

_ACT = {None: 0, "none": 0, "relu": 1, "sigmoid": 2}
_PATH = {"auto": 0, "simt": 1, "tc": 2, "mma": 3}


def _ld(t):
    assert t.dim() == 2 and t.stride(1) == 1, "expected a 2-d tensor with unit column stride"
    return t.stride(0)


def linear(x, weight, bias=None, act=None, mul=None, add=None, rowzero=None, out_dtype=None, path="auto", out=None):
    """act(x @ weight.T + bias) [* mul] [+ add], rows with rowzero != 0 zeroed.  x (M,K), weight (N,K)."""
    M, K = x.shape
    N = weight.shape[0]
    assert weight.shape[1] == K and x.dtype == weight.dtype
    out_dtype = out_dtype or x.dtype
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=x.device)
    if bias is not None:
        bias = bias.float().contiguous()
    with torch.cuda.device(x.device):
        rc = _lib.lib().memotr_linear(
            _lib.ptr(x), _ld(x), _lib.ptr(weight), _ld(weight), _lib.ptr(bias), _lib.ptr(mul), _ld(mul) if mul is not None else 0,
            _lib.ptr(add), _ld(add) if add is not None else 0, _lib.ptr(rowzero), _lib.ptr(out), _ld(out), M, N, K,
            _lib.dtype_code(x), _lib.dtype_code(out), _ACT[act], _PATH[path], _lib.stream_ptr())
    _lib.check(rc, "memotr_linear")
    return out


def mlp2(x, w1, b1, w2, b2=None, act2=None, mul=None, out_dtype=None, out=None):
    """act2(relu(x @ w1.T + b1) @ w2.T + b2) [* mul] in one tensor-core kernel; x (M,256), w1 (Hd,256), w2 (256,Hd) bf16."""
    M, K1 = x.shape
    Hd, N2 = w1.shape[0], w2.shape[0]
    if out is None:
        out = torch.empty((M, N2), dtype=out_dtype or x.dtype, device=x.device)
    b1 = b1.float().contiguous()
    b2 = b2.float().contiguous() if b2 is not None else None
    with torch.cuda.device(x.device):
        rc = _lib.lib().memotr_mlp2(_lib.ptr(x), _ld(x), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2),
                                    _lib.ptr(mul), _ld(mul) if mul is not None else 0, _lib.ptr(out), _ld(out), M, K1, Hd,
                                    N2, _lib.dtype_code(out), _ACT[act2], _lib.stream_ptr())
    _lib.check(rc, "memotr_mlp2")
    return out


def layernorm(x, gamma, beta, x2=None, pos=None, eps=1e-5, out_dtype=None, want_f32=False):
    """LayerNorm(x [+ x2]) over the last (256-wide) dimension; returns y, or (y, y + pos), plus an fp32 copy on request."""
    M, C = x.shape
    out_dtype = out_dtype or x.dtype
    y = torch.empty((M, C), dtype=out_dtype, device=x.device)
    ypos = torch.empty_like(y) if pos is not None else None
    y32 = torch.empty((M, C), dtype=torch.float32, device=x.device) if want_f32 else None
    with torch.cuda.device(x.device):
        rc = _lib.lib().memotr_layernorm(
            _lib.ptr(x), _lib.dtype_code(x), _ld(x), _lib.ptr(x2), _ld(x2) if x2 is not None else 0,
            _lib.ptr(gamma), _lib.ptr(beta), eps, _lib.ptr(y), _lib.dtype_code(y), C, _lib.ptr(pos),
            _ld(pos) if pos is not None else 0, _lib.ptr(ypos), C, _lib.ptr(y32), C, M, C, _lib.stream_ptr())
    _lib.check(rc, "memotr_layernorm")
    res = (y,) + ((ypos,) if pos is not None else ()) + ((y32,) if want_f32 else ())
    return res[0] if len(res) == 1 else res


def mha(q, k, v, n_heads, key_padding_mask=None, out_dtype=None):
    """softmax(q k^T / sqrt(32) + mask) v per head; q (Nq, n_heads*32), k/v (Nk, n_heads*32)."""
    Nq, C = q.shape
    Nk = k.shape[0]
    out = torch.empty((Nq, C), dtype=out_dtype or q.dtype, device=q.device)
    kpm = key_padding_mask.to(torch.uint8).contiguous() if key_padding_mask is not None else None
    with torch.cuda.device(q.device):
        rc = _lib.lib().memotr_mha(_lib.ptr(q), _ld(q), _lib.ptr(k), _ld(k), _lib.ptr(v), _ld(v), _lib.ptr(kpm),
                                   _lib.ptr(out), C, Nq, Nk, n_heads, C // n_heads, _lib.dtype_code(q),
                                   _lib.dtype_code(out), _lib.stream_ptr())
    _lib.check(rc, "memotr_mha")
    return out


def msda_prep(ol, spatial_shapes, level_start_index, valid_ratios, n_heads, n_levels, n_points, ref4=None):
    """-> sampling_loc (Lq,H,L,K,2), attn_weight (Lq,H,L,K).  ref4=None: encoder mode (queries are the pixels)."""
    Lq = ol.shape[0]
    loc = torch.empty((Lq, n_heads, n_levels, n_points, 2), dtype=torch.float32, device=ol.device)
    attn = torch.empty((Lq, n_heads, n_levels, n_points), dtype=torch.float32, device=ol.device)
    with torch.cuda.device(ol.device):
        rc = _lib.lib().memotr_msda_prep(_lib.ptr(ol), _ld(ol), _lib.ptr(spatial_shapes), _lib.ptr(level_start_index),
                                         _lib.ptr(valid_ratios), _lib.ptr(ref4), 0 if ref4 is None else 1, _lib.ptr(loc),
                                         _lib.ptr(attn), Lq, n_heads, n_levels, n_points, _lib.stream_ptr())
    _lib.check(rc, "memotr_msda_prep")
    return loc, attn


def msda_forward_ex(value, spatial_shapes, level_start_index, loc, attn, n_heads):
    """value (S, >=H*32 columns; may be a column slice of a wider buffer) f32/bf16; loc/attn fp32 -> (Lq, H*32)."""
    S = value.shape[0]
    Lq, H, L, K = attn.shape
    out_dtype = torch.bfloat16 if value.dtype == torch.float16 else value.dtype     # fp16 value maps feed bf16 GEMMs
    out = torch.empty((Lq, H * 32), dtype=out_dtype, device=value.device)
    with torch.cuda.device(value.device):
        rc = _lib.lib().memotr_msda_forward_ex(_lib.ptr(value), _ld(value), _lib.ptr(spatial_shapes),
                                               _lib.ptr(level_start_index), _lib.ptr(loc), _lib.ptr(attn), _lib.ptr(out),
                                               1, S, H, L, Lq, K, _lib.dtype_code(value), _lib.stream_ptr())
    _lib.check(rc, "memotr_msda_forward_ex")
    return out


def linear_msda_prep(x, w, b, shapes, level_start, valid_ratios, n_heads, n_levels, n_points):
    """Offsets / logits projection with the location + softmax epilogue (encoder mode): x (M,K), w (3*H*L*K, K) bf16 ->
    (M, 3*H*L*K) fp32 rows [sampling locations (H, L*K, 2) | attention weights (H, L*K)].  shapes / level_start: python ints."""
    import ctypes
    M, K = x.shape
    N = n_heads * n_levels * n_points * 3
    out = torch.empty((M, N), dtype=torch.float32, device=x.device)
    hw = (ctypes.c_int * (2 * n_levels))(*[int(v) for s_ in shapes for v in s_])
    lsi = (ctypes.c_int * n_levels)(*[int(v) for v in level_start])
    with torch.cuda.device(x.device):
        rc = _lib.lib().memotr_linear_msda_prep(_lib.ptr(x), _ld(x), _lib.ptr(w), _ld(w), _lib.ptr(b), _lib.ptr(out), N, M, K,
                                                n_heads, n_levels, n_points, hw, lsi, _lib.ptr(valid_ratios),
                                                _lib.stream_ptr())
    _lib.check(rc, "memotr_linear_msda_prep")
    return out


def msda_forward_strided(value, spatial_shapes, level_start_index, rows=None, n_heads=8, n_levels=4, n_points=4, loc=None,
                         attn=None):
    """Global-memory gather from an fp16 value map (S, >= H*32): locations / weights either from the (Lq, 3*H*L*K) rows of
    linear_msda_prep (`rows`) or from dense `loc` (Lq,H,L,K,2) / `attn` (Lq,H,L,K) fp32."""
    S = value.shape[0]
    if rows is not None:
        Lq, N = rows.shape[0], rows.shape[1]
        loc, attn, ld_loc, ld_attn = rows, rows.view(-1)[n_heads * n_levels * n_points * 2:], N, N
    else:
        Lq, ld_loc, ld_attn = loc.shape[0], n_heads * n_levels * n_points * 2, n_heads * n_levels * n_points
    out = torch.empty((Lq, n_heads * 32), dtype=torch.bfloat16, device=value.device)
    with torch.cuda.device(value.device):
        rc = _lib.lib().memotr_msda_forward_strided(_lib.ptr(value), _ld(value), _lib.ptr(spatial_shapes),
                                                    _lib.ptr(level_start_index), _lib.ptr(loc), ld_loc, _lib.ptr(attn), ld_attn,
                                                    _lib.ptr(out), 1, S, n_heads, n_levels, Lq, n_points, _lib.stream_ptr())
    _lib.check(rc, "memotr_msda_forward_strided")
    return out


def window_plan(shapes, n_heads=8, n_points=4, radius=2.5, max_classes=0):
    """The staging plan of msda_forward_window (host only): dict(classes, units, global_ctas, global_q0, smem, cls=[...])."""
    import ctypes
    L = len(shapes)
    hw = (ctypes.c_int * (2 * L))(*[int(v) for s_ in shapes for v in s_])
    sizes = [int(h) * int(w) for h, w in shapes]
    lsi = (ctypes.c_int * L)(*[sum(sizes[:i]) for i in range(L)])
    info = (ctypes.c_int * (8 + 4 * 18))()
    _lib.check(_lib.lib().memotr_msda_window_plan(hw, lsi, sum(sizes), n_heads, L, n_points, float(radius), max_classes, info),
               "memotr_msda_window_plan")
    out = dict(classes=info[0], units=info[1], global_ctas=info[2], global_q0=info[3], smem=info[4], cls=[])
    for c in range(info[0]):
        o = info[8 + 18 * c: 8 + 18 * (c + 1)]
        out["cls"].append(dict(level=o[0], tile=(o[1], o[2]), tiles=(o[3], o[4]), units=o[5], tma_bytes=o[6], rec_stride=o[7],
                               ww=list(o[8:8 + L]), wh=list(o[13:13 + L])))
    return out


def msda_forward_window(value, shapes, valid_ratios, rows=None, n_heads=8, n_points=4, shift=None, radius=2.5, max_classes=0,
                        loc=None, attn=None, stats=None):
    """Encoder-shaped gather from TMA-staged windows (csrc/msda_window.cu): value (S, >= H*32) fp16 pixel-major; locations /
    weights as for msda_forward_strided; `shift` (H, L, 2) host floats or None; `stats` device int64[2] or None."""
    import ctypes
    L = len(shapes)
    S = value.shape[0]
    if rows is not None:
        N = rows.shape[1]
        loc, attn, ld_loc, ld_attn = rows, rows.view(-1)[n_heads * L * n_points * 2:], N, N
    else:
        ld_loc, ld_attn = n_heads * L * n_points * 2, n_heads * L * n_points
    out = torch.empty((S, n_heads * 32), dtype=torch.bfloat16, device=value.device)
    hw = (ctypes.c_int * (2 * L))(*[int(v) for s_ in shapes for v in s_])
    sizes = [int(h) * int(w) for h, w in shapes]
    lsi = (ctypes.c_int * L)(*[sum(sizes[:i]) for i in range(L)])
    sh = None
    if shift is not None:
        flat = [float(v) for v in torch.as_tensor(shift, dtype=torch.float32).reshape(-1).tolist()]
        assert len(flat) == n_heads * L * 2
        sh = (ctypes.c_float * len(flat))(*flat)
    with torch.cuda.device(value.device):
        rc = _lib.lib().memotr_msda_forward_window(_lib.ptr(value), _ld(value), hw, lsi, _lib.ptr(loc), ld_loc, _lib.ptr(attn),
                                                   ld_attn, _lib.ptr(valid_ratios), sh, float(radius), int(max_classes),
                                                   _lib.ptr(stats), _lib.ptr(out), S, n_heads, L, n_points, _lib.stream_ptr())
    _lib.check(rc, "memotr_msda_forward_window")
    return out


def sine_embed(pts, dim_t, scale4=None, apply_sigmoid=False, out_dtype=torch.float32):
    N = pts.shape[0]
    out = torch.empty((N, 512), dtype=out_dtype, device=pts.device)
    with torch.cuda.device(pts.device):
        rc = _lib.lib().memotr_sine_embed(_lib.ptr(pts), _ld(pts), _lib.ptr(scale4), int(apply_sigmoid), _lib.ptr(dim_t),
                                          _lib.ptr(out), 512, N, _lib.dtype_code(out), _lib.stream_ptr())
    _lib.check(rc, "memotr_sine_embed")
    return out


def mlp2_lnout(x, w1, b1, w2, b2, res, gamma, beta, pos=None, eps=1e-5, want_f32=True):
    """LayerNorm(res + relu(x W1^T + b1) W2^T + b2) in the FFN kernel's epilogue.  x (M,256) bf16, res (M,256) fp32, pos bf16
    or None.  -> (y bf16, y32 fp32 or None, y + pos bf16 or None)."""
    M, Hd = x.shape[0], w1.shape[0]
    y = torch.empty((M, 256), dtype=torch.bfloat16, device=x.device)
    y32 = torch.empty((M, 256), dtype=torch.float32, device=x.device) if want_f32 else None
    ypos = torch.empty((M, 256), dtype=torch.bfloat16, device=x.device) if pos is not None else None
    with torch.cuda.device(x.device):
        rc = _lib.lib().memotr_mlp2_lnout(_lib.ptr(x), _ld(x), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2),
                                          _lib.ptr(res), _ld(res), _lib.ptr(gamma), _lib.ptr(beta), float(eps), _lib.ptr(y), 256,
                                          _lib.ptr(y32), 256, _lib.ptr(pos), _ld(pos) if pos is not None else 0, _lib.ptr(ypos),
                                          256, M, Hd, _lib.stream_ptr())
    _lib.check(rc, "memotr_mlp2_lnout")
    return y, y32, ypos


def encoder_dense_block(att, wout, bout, src32, g1, be1, w1, b1, w2, b2, g2, be2, pos, eps=1e-5):
    """output_proj + norm1 + FFN + norm2 of an encoder layer in one kernel per row tile (memotr_encoder_dense_block).
    att (M,256) bf16, src32 (M,256) fp32, pos (M,256) bf16 -> (x32 fp32, y bf16, y32 fp32, y + pos bf16)."""
    M, Hd, dev = att.shape[0], w1.shape[0], att.device
    x32, y32, pre = (torch.empty((M, 256), dtype=torch.float32, device=dev) for _ in range(3))
    y, ypos = (torch.empty((M, 256), dtype=torch.bfloat16, device=dev) for _ in range(2))
    with torch.cuda.device(dev):
        rc = _lib.lib().memotr_encoder_dense_block(_lib.ptr(att), _ld(att), _lib.ptr(wout), _lib.ptr(bout), _lib.ptr(src32),
                                                   _ld(src32), _lib.ptr(g1), _lib.ptr(be1), _lib.ptr(x32), 256, _lib.ptr(w1),
                                                   _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), _lib.ptr(g2), _lib.ptr(be2),
                                                   _lib.ptr(pos), _ld(pos), _lib.ptr(y), 256, _lib.ptr(y32), 256, _lib.ptr(ypos), 256,
                                                   _lib.ptr(pre), 256, M, Hd, float(eps), _lib.stream_ptr())
    _lib.check(rc, "memotr_encoder_dense_block")
    return x32, y, y32, ypos


W3_SHIFT = 6          # the split weights carry 2^6: their low halves stay normal fp16 numbers (oracle/frame.py models the same)


def pack_w3(w):
    """(N, K) fp32 weight -> (N, 3K) fp16 [hi | lo | hi] of 2^W3_SHIFT * w: the B operand of memotr_linear_f32x3."""
    ws = w.float() * float(2 ** W3_SHIFT)
    hi = ws.half()
    lo = (ws - hi.float()).half()
    return torch.cat((hi, lo, hi), dim=1).contiguous()


def linear_f32x3(x, w3, bias=None, act=None, rowzero=None, out=None, scratch=None, split_out=False):
    """act(x @ w.T + bias) for fp32 x with fp32-accurate products on the tensor cores (memotr_linear_f32x3); w3 = pack_w3(w).
    x may already be a split operand (fp16, (M, 3K): the split_out=True result of a previous call); split_out=True returns the
    result as such an operand (M, 3N) instead of fp32."""
    pre = x.dtype == torch.float16
    M, K = x.shape[0], (x.shape[1] // 3 if pre else x.shape[1])
    N = w3.shape[0]
    assert w3.shape[1] == 3 * K and w3.dtype == torch.float16 and (pre or x.dtype == torch.float32)
    if split_out:
        res = torch.empty((M, 3 * N), dtype=torch.float16, device=x.device)
    else:
        res = out if out is not None else torch.empty((M, N), dtype=torch.float32, device=x.device)
    if scratch is None:
        scratch = x if pre else torch.empty((M, 3 * K), dtype=torch.float16, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib().memotr_linear_f32x3(None if pre else _lib.ptr(x), 0 if pre else _ld(x), _lib.ptr(w3), _lib.ptr(bias),
                                            _lib.ptr(rowzero), None if split_out else _lib.ptr(res), 0 if split_out else _ld(res),
                                            M, N, K, _ACT[act], float(2.0 ** -W3_SHIFT), _lib.ptr(scratch),
                                            _lib.ptr(res) if split_out else None, _lib.stream_ptr())
    _lib.check(rc, "memotr_linear_f32x3")
    return res


def linear256_layernorm(a, w, b, res, gamma, beta, eps=1e-5):
    """LayerNorm(res + a @ w.T + b) for a 256 x 256 projection in one kernel (memotr_linear256_layernorm).
    a (M,256) bf16, w (256,256) bf16, res (M,256) fp32 -> (y bf16, y32 fp32)."""
    M = a.shape[0]
    y = torch.empty((M, 256), dtype=torch.bfloat16, device=a.device)
    y32 = torch.empty((M, 256), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        rc = _lib.lib().memotr_linear256_layernorm(_lib.ptr(a), _ld(a), _lib.ptr(w), _lib.ptr(b), _lib.ptr(res), _ld(res),
                                                   _lib.ptr(gamma), _lib.ptr(beta), float(eps), _lib.ptr(y), 256, _lib.ptr(y32), 256,
                                                   M, _lib.stream_ptr())
    _lib.check(rc, "memotr_linear256_layernorm")
    return y, y32


def pos_embed_sine(mask, num_pos_feats=128, temperature=20, scale=6.283185307179586):
    """PositionEmbeddingSine(normalize=True) of one (H, W) padding mask (bool / uint8, device) -> (2*num_pos_feats, H, W) fp32
    (models/position_embedding.py:23-49)."""
    _lib.require_cuda(mask=mask)
    H, W = mask.shape
    m = mask.to(torch.uint8).contiguous()
    i = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_i = (float(temperature) ** (2 * torch.div(i, 2, rounding_mode="trunc") / num_pos_feats)).to(mask.device)
    scratch = torch.empty(2 * H * W, dtype=torch.float32, device=mask.device)
    out = torch.empty(2 * num_pos_feats, H, W, dtype=torch.float32, device=mask.device)
    with torch.cuda.device(mask.device):
        rc = _lib.lib().memotr_pos_embed_sine(_lib.ptr(m), H, W, _lib.ptr(dim_i), num_pos_feats, float(scale),
                                              _lib.ptr(scratch), _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "memotr_pos_embed_sine")
    return out


def box_refine(delta, ref, n_take):
    new_ref, ref_next = torch.empty_like(ref), torch.empty_like(ref)
    with torch.cuda.device(ref.device):
        rc = _lib.lib().memotr_box_refine(_lib.ptr(delta), _lib.ptr(ref), _lib.ptr(new_ref), _lib.ptr(ref_next),
                                          ref.shape[0], n_take, _lib.stream_ptr())
    _lib.check(rc, "memotr_box_refine")
    return new_ref, ref_next
