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
