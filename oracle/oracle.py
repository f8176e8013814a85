"""ctypes loader for the CPU oracle — TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libmnv1_oracle.so")
LITERAL_SO = os.path.join(HERE, "_ref", "libmnv1_literal.so")

ACT_NONE, ACT_RELU, ACT_RELU6 = 0, 1, 2
STORE_F32, STORE_BF16, STORE_U8_SAT, STORE_U8_WRAP = 0, 1, 2, 3   # the `rbf16` argument below takes these too
PAD_REF, PAD_TFSAME = 0, 1


def build(force: bool = False) -> None:
    """Compile oracle/ (and oracle/_ref when /root/reference is mounted)."""
    src_new = any(
        os.path.getmtime(os.path.join(HERE, f)) > os.path.getmtime(ORACLE_SO)
        for f in ("mnv1_oracle.c", "mnv1_oracle.h")
    ) if os.path.exists(ORACLE_SO) else True
    if force or src_new or not os.path.exists(LITERAL_SO):
        subprocess.run(["make", "-C", HERE], check=True, capture_output=True)


class _Ep(C.Structure):
    _fields_ = [("scale", C.c_void_p), ("shift", C.c_void_p), ("act", C.c_int), ("round_bf16", C.c_int)]


_lib = None
_lit = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        _lib = C.CDLL(ORACLE_SO)
        _lib.mnv1o_round_bf16.restype = C.c_float
        _lib.mnv1o_round_bf16.argtypes = [C.c_float]
    return _lib


def literal():
    """kernel.cl compiled as C (oracle/_ref); None when it was never built."""
    global _lit
    if _lit is None:
        if not os.path.exists(LITERAL_SO):
            try:
                build()
            except Exception:
                pass
        if not os.path.exists(LITERAL_SO):
            return None
        _lit = C.CDLL(LITERAL_SO)
    return _lit


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _ep(scale, shift, act, round_bf16, keep):
    s, t = _f32(scale), _f32(shift)
    keep += [s, t]
    return _Ep(_p(s), _p(t), act, int(round_bf16))


def round_bf16(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(x.shape)


def convolute(r, g, b, filt, n, rows, cols, stride, op_size, pad_mode=PAD_REF, in_scale=1.0, in_bias=0.0,
              scale=None, shift=None, act=ACT_RELU, rbf16=False, pix_stride=1, img_stride=None):
    keep = []
    filt = _f32(filt)
    out = np.empty((n, op_size, rows // stride, cols // stride), dtype=np.float32)
    ep = _ep(scale, shift, act, rbf16, keep)
    if img_stride is None:
        img_stride = rows * cols * pix_stride
    lib().mnv1o_convolute(_p(out), _p(r), _p(g), _p(b), C.c_int(pix_stride), C.c_long(img_stride), _p(filt),
                          n, rows, cols, 3, stride, op_size, pad_mode, C.c_float(in_scale), C.c_float(in_bias),
                          C.byref(ep))
    return out


def depthwise(x, filt, stride, pad_mode=PAD_REF, scale=None, shift=None, act=ACT_RELU, rbf16=False):
    keep = []
    x, filt = _f32(x), _f32(filt)
    n, c, rows, cols = x.shape
    out = np.empty((n, c, rows // stride, cols // stride), dtype=np.float32)
    ep = _ep(scale, shift, act, rbf16, keep)
    lib().mnv1o_depthwise(_p(out), _p(x), _p(filt), n, rows, cols, 3, stride, c, pad_mode, C.byref(ep))
    return out


def pointwise(x, filt, op_size, scale=None, shift=None, act=ACT_RELU, rbf16=False):
    keep = []
    x, filt = _f32(x), _f32(filt)
    n, cin, rows, cols = x.shape
    out = np.empty((n, op_size, rows, cols), dtype=np.float32)
    ep = _ep(scale, shift, act, rbf16, keep)
    lib().mnv1o_pointwise(_p(out), _p(x), _p(filt), n, rows, cols, cin, op_size, C.byref(ep))
    return out


def pool(x, truncate=False, rbf16=False):
    x = _f32(x)
    n, c, rows, cols = x.shape
    out = np.empty((n, c), dtype=np.float32)
    lib().mnv1o_pool(_p(out), _p(x), n, rows, cols, rows, c, int(truncate), int(rbf16))
    return out


def softmax_argmax(logits):
    logits = _f32(logits)
    n, k = logits.shape
    prob = np.empty((n, k), dtype=np.float64)
    top1 = np.empty(n, dtype=np.int32)
    p1 = np.empty(n, dtype=np.float64)
    lib().mnv1o_softmax_argmax(_p(logits), n, k, _p(prob), _p(top1), _p(p1))
    return prob, top1, p1


_SHAPES = None


def layer_shapes():
    """(cout, hout) of the 29 layers, from the oracle's own table."""
    global _SHAPES
    if _SHAPES is None:
        class L(C.Structure):
            _fields_ = [("kind", C.c_int), ("cin", C.c_int), ("cout", C.c_int), ("hin", C.c_int),
                        ("hout", C.c_int), ("stride", C.c_int), ("w_off", C.c_long), ("w_cnt", C.c_long),
                        ("c_off", C.c_long)]
        lib().mnv1o_layers.restype = C.POINTER(L)
        p = lib().mnv1o_layers()
        _SHAPES = [(p[i].kind, p[i].cin, p[i].cout, p[i].hin, p[i].hout, p[i].stride, p[i].w_off, p[i].w_cnt,
                    p[i].c_off) for i in range(29)]
    return _SHAPES


def forward(images_u8, weights, scale=None, shift=None, pad_mode=PAD_TFSAME, act=ACT_RELU6, rbf16=0,
            in_scale=1.0 / 127.5, in_bias=-1.0, last_layer=29, taps=()):
    """Run layers 1..last_layer.  Returns (final_out, {layer: planar output}) ."""
    images_u8 = np.ascontiguousarray(images_u8, dtype=np.uint8)
    n = images_u8.shape[0]
    weights, scale, shift = _f32(weights), _f32(scale), _f32(shift)
    sh = layer_shapes()
    _, _, cout, _, hout, *_ = sh[last_layer - 1]
    final = np.empty((n, cout, hout, hout), dtype=np.float32)
    tap_arr = (C.c_void_p * 29)()
    tap_out = {}
    for k in taps:
        _, _, c, _, h, *_ = sh[k - 1]
        tap_out[k] = np.empty((n, c, h, h), dtype=np.float32)
        tap_arr[k - 1] = tap_out[k].ctypes.data
    lib().mnv1o_forward(_p(images_u8), n, _p(weights), _p(scale), _p(shift), pad_mode, act, int(rbf16),
                        C.c_float(in_scale), C.c_float(in_bias), last_layer, _p(final), tap_arr)
    if last_layer == 29:
        final = final.reshape(n, 1000)
    return final, tap_out


# ---- literal kernel.cl (oracle/_ref) -------------------------------------------------
def lit_depthwise_per_channel(x_u8, filt_i32, stride=1):
    """kernel.cl `depthwise` launched once per channel (op_size=1, offset pointers): the KAT
    trick of SURVEY §8(c).  x [C][H][W] u8, filt [C][3][3] int32 -> [C][H][W] u8."""
    L = literal()
    c, h, w = x_u8.shape
    guard = np.zeros((c, h + 2, w), dtype=np.uint8)  # zeroed guard rows after each plane (D-08)
    guard[:, :h] = x_u8
    filt = np.ascontiguousarray(filt_i32, dtype=np.int32)
    out = np.zeros((c, h, w), dtype=np.uint8)
    for ch in range(c):
        L.lit_depthwise(C.c_void_p(out[ch].ctypes.data), C.c_void_p(guard[ch].ctypes.data),
                        C.c_void_p(filt[ch].ctypes.data), h, w, 3, stride, 1, w, h)
    return out


def lit_pointwise_per_channel(x_u8, filt_i32):
    """kernel.cl `pointwise`, one launch per output channel with filtersize = Cin."""
    L = literal()
    cin, h, w = x_u8.shape
    x = np.ascontiguousarray(x_u8, dtype=np.uint8)
    filt = np.ascontiguousarray(filt_i32, dtype=np.int32)
    cout = filt.shape[0]
    out = np.zeros((cout, h, w), dtype=np.uint8)
    for co in range(cout):
        L.lit_pointwise(C.c_void_p(out[co].ctypes.data), _p(x), C.c_void_p(filt[co].ctypes.data), h, w, cin, 1, w, h)
    return out


def lit_pool_per_channel(x_u8):
    """kernel.cl `pool` on a 1x1 NDRange, one launch per channel."""
    L = literal()
    c = x_u8.shape[0]
    x = np.ascontiguousarray(x_u8, dtype=np.uint8)
    out = np.zeros(c, dtype=np.uint8)
    for ch in range(c):
        L.lit_pool(C.c_void_p(out[ch:].ctypes.data), C.c_void_p(x[ch].ctypes.data), 7, 7, 7, 1, 1, 1)
    return out


def lit_forward(images_u8, weights_i32):
    L = literal()
    images_u8 = np.ascontiguousarray(images_u8, dtype=np.uint8)
    w = np.ascontiguousarray(weights_i32, dtype=np.int32)
    n = images_u8.shape[0]
    out = np.zeros((n, 1000), dtype=np.uint8)
    L.lit_forward(_p(images_u8), n, _p(w), _p(out))
    return out
