"""Pins oracle/mnv1_oracle.c against the reference's own kernel.cl compiled unchanged as C
(oracle/_ref), on the domain where that code is well defined (SURVEY §8c "KAT trick"):
per-output-channel launches with op_size=1, offset pointers, filtersize=Cin, guard rows, small
integers so the u8 store does not wrap.  Bit-exact."""
import numpy as np
import pytest

import mnv1_b200  # noqa: F401
from mnv1_b200 import synth


def _ints(seed, shape, lo, hi):
    return synth.kat_ints(seed, int(np.prod(shape)), lo, hi).reshape(shape)


@pytest.fixture(scope="module")
def lit(oracle_mod):
    if oracle_mod.literal() is None:
        pytest.skip("oracle/_ref/libmnv1_literal.so was never built (reference not mounted)")
    return oracle_mod


@pytest.mark.parametrize("c,h,w", [(8, 14, 14), (3, 9, 17), (16, 28, 28), (1, 5, 5)])
def test_depthwise_stride1(lit, c, h, w):
    x = _ints(11, (c, h, w), 0, 3).astype(np.uint8)
    f = _ints(12, (c, 3, 3), -2, 2).astype(np.int32)
    want = lit.depthwise(x[None].astype(np.float32), f.astype(np.float32), 1, act=lit.ACT_RELU)[0]
    got = lit.lit_depthwise_per_channel(x, f, 1)
    # the literal kernel has no upper-bound check: its right column reads the next row (D-08)
    assert np.array_equal(got[:, :, :-1].astype(np.float32), want[:, :, :-1])
    # ... and called the way MobileNet.c calls it (one launch, op_size=C) it is NOT the layer (D-01/D-03)
    L = lit.literal()
    import ctypes as C
    guard = np.zeros((c, h + 2, w), np.uint8); guard[:, :h] = x
    out = np.zeros((c, h, w), np.uint8)
    flat = np.ascontiguousarray(guard[:, :h].reshape(-1))
    buf = np.zeros(flat.size + 4 * w, np.uint8); buf[:flat.size] = flat
    L.lit_depthwise(C.c_void_p(out.ctypes.data), C.c_void_p(buf.ctypes.data), C.c_void_p(f.ctypes.data), h, w, 3, 1, c, w, h)
    if c > 1:
        assert not np.array_equal(out[:, :, :-1].astype(np.float32), want[:, :, :-1])


@pytest.mark.parametrize("cin,cout,h", [(8, 16, 14), (32, 64, 7), (3, 5, 9), (64, 8, 4)])
def test_pointwise_and_fc(lit, cin, cout, h):
    hi = 3 if cin <= 32 else 1
    x = _ints(13, (cin, h, h), 0, hi).astype(np.uint8)
    f = _ints(14, (cout, cin), -1, 1).astype(np.int32)
    want = lit.pointwise(x[None].astype(np.float32), f.astype(np.float32), cout, act=lit.ACT_RELU)[0]
    assert want.max() < 256
    assert np.array_equal(lit.lit_pointwise_per_channel(x, f).astype(np.float32), want)


def test_fc_is_pointwise_at_1x1(lit):
    x = _ints(15, (64, 1, 1), 0, 3).astype(np.uint8)
    f = _ints(16, (10, 64), -1, 1).astype(np.int32)
    want = lit.pointwise(x[None].astype(np.float32), f.astype(np.float32), 10, act=lit.ACT_RELU)[0]
    assert np.array_equal(lit.lit_pointwise_per_channel(x, f).astype(np.float32), want)


def test_pool(lit):
    x = _ints(17, (32, 7, 7), 0, 5).astype(np.uint8)
    want = lit.pool(x[None].astype(np.float32), truncate=True)[0]
    assert np.array_equal(lit.lit_pool_per_channel(x).astype(np.float32), want)


def test_literal_forward_runs(lit):
    """the CPU-baseline path (whole schedule through kernel.cl) is deterministic and complete"""
    img = synth.images(2)
    w = synth.kat_ints(7, 4209088, -2, 2).astype(np.int32)
    a = lit.lit_forward(img, w)
    b = lit.lit_forward(img[1:], w)
    assert a.shape == (2, 1000) and np.array_equal(a[1], b[0])


# ---- the oracle's INTEGER store modes (round 2): full-range u8 data, sums far beyond 255 ----------------------
@pytest.mark.parametrize("cin,cout,h", [(64, 32, 9), (512, 24, 14), (1024, 16, 7), (3, 7, 5)])
def test_integer_wrap_mode_pointwise_full_range(lit, cin, cout, h):
    """kernel.cl `pointwise` stores its int sum into an unsigned char (kernel.cl:112): with u8 inputs in 0..255
    and filters in -128..127 the sums reach millions and the store wraps modulo 256.  The oracle's
    STORE_U8_WRAP mode must reproduce that bit for bit — this is what lets the GPU's MNV1_U8 contexts be
    pinned on real-size layers (tests/test_gpu_int8.py)."""
    rng = np.random.default_rng(cin + cout)
    x = rng.integers(0, 256, (cin, h, h), dtype=np.uint8)
    f = rng.integers(-128, 128, (cout, cin)).astype(np.int32)
    want = lit.pointwise(x[None].astype(np.float32), f.astype(np.float32), cout, act=lit.ACT_RELU, rbf16=lit.STORE_U8_WRAP)[0]
    got = lit.lit_pointwise_per_channel(x, f)
    assert np.array_equal(got.astype(np.float32), want)
    sat = lit.pointwise(x[None].astype(np.float32), f.astype(np.float32), cout, act=lit.ACT_RELU, rbf16=lit.STORE_U8_SAT)[0]
    assert sat.max() == 255 and not np.array_equal(sat, want)      # saturation is the other, non-literal store


@pytest.mark.parametrize("c,h,w", [(16, 14, 14), (4, 9, 17), (32, 7, 7)])
def test_integer_wrap_mode_depthwise_full_range(lit, c, h, w):
    rng = np.random.default_rng(c + h)
    x = rng.integers(0, 256, (c, h, w), dtype=np.uint8)
    f = rng.integers(-128, 128, (c, 3, 3)).astype(np.int32)
    want = lit.depthwise(x[None].astype(np.float32), f.astype(np.float32), 1, act=lit.ACT_RELU, rbf16=lit.STORE_U8_WRAP)[0]
    got = lit.lit_depthwise_per_channel(x, f, 1)
    assert np.array_equal(got[:, :, :-1].astype(np.float32), want[:, :, :-1])


def test_integer_mode_bias_shift_and_forward(oracle_mod):
    """out = store(relu(sum + bias) >> s): checked against plain numpy integer arithmetic, and the integer forward
    chain stays in [0, 255] with integer values at every tap"""
    rng = np.random.default_rng(3)
    cin, cout, h, s = 48, 20, 6, 7
    x = rng.integers(0, 256, (2, cin, h, h), dtype=np.uint8)
    f = rng.integers(-128, 128, (cout, cin)).astype(np.int64)
    bias = rng.integers(-3000, 3000, cout).astype(np.int64)
    acc = np.einsum("oc,nchw->nohw", f, x.astype(np.int64)) + bias[None, :, None, None]
    ref_sat = np.clip(np.maximum(acc, 0) >> s, 0, 255)
    ref_wrap = (np.maximum(acc, 0) >> s) & 255
    sc = np.full(cout, 2.0 ** -s, np.float32)
    for mode, ref in ((oracle_mod.STORE_U8_SAT, ref_sat), (oracle_mod.STORE_U8_WRAP, ref_wrap)):
        got = oracle_mod.pointwise(x.astype(np.float32), f.astype(np.float32), cout, scale=sc, shift=bias.astype(np.float32),
                                   act=oracle_mod.ACT_RELU, rbf16=mode)
        assert np.array_equal(got, ref.astype(np.float32))
    neg = oracle_mod.pointwise(x.astype(np.float32), f.astype(np.float32), cout, scale=sc, shift=bias.astype(np.float32),
                               act=oracle_mod.ACT_NONE, rbf16=oracle_mod.STORE_U8_WRAP)
    assert np.array_equal(neg, ((acc >> s) & 255).astype(np.float32))      # no ReLU: floor shift, then the C conversion
