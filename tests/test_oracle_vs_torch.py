"""Pins the oracle where nothing in the reference does (stride 2, borders, BN fold, ReLU6,
bias, TF-SAME padding): an independent implementation — torch.nn.functional on the
CPU, fp64 — must agree to fp32 rounding."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import mnv1_b200  # noqa: F401
from mnv1_b200 import synth


def _pad(x, stride, pad_mode):
    if stride == 2 and pad_mode == 1:  # TF-SAME on even sizes: bottom/right only
        return F.pad(x, (0, 1, 0, 1))
    return F.pad(x, (1, 1, 1, 1))


def _ep(y, sc, sh, relu6=True):
    y = y * torch.from_numpy(sc).double().view(1, -1, 1, 1) + torch.from_numpy(sh).double().view(1, -1, 1, 1)
    return y.clamp(0, 6) if relu6 else y


@pytest.mark.parametrize("stride,pad_mode", [(1, 0), (2, 0), (2, 1)])
def test_depthwise(oracle_mod, stride, pad_mode):
    rng = np.random.default_rng(1)
    c, h = 24, 20
    x = rng.standard_normal((2, c, h, h)).astype(np.float32)
    w = rng.standard_normal((c, 3, 3)).astype(np.float32)
    sc, sh = (0.5 + rng.random(c)).astype(np.float32), rng.standard_normal(c).astype(np.float32)
    got = oracle_mod.depthwise(x, w, stride, pad_mode=pad_mode, scale=sc, shift=sh, act=oracle_mod.ACT_RELU6)
    ref = F.conv2d(_pad(torch.from_numpy(x).double(), stride, pad_mode), torch.from_numpy(w).double().view(c, 1, 3, 3),
                   stride=stride, groups=c)
    ref = _ep(ref, sc, sh)
    assert got.shape == tuple(ref.shape)
    np.testing.assert_allclose(got, ref.numpy(), rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("pad_mode", [0, 1])
def test_stem(oracle_mod, pad_mode):
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (2, 32, 32, 3), dtype=np.uint8)
    w = rng.standard_normal((8, 3, 3, 3)).astype(np.float32)
    sc, sh = (0.5 + rng.random(8)).astype(np.float32), rng.standard_normal(8).astype(np.float32)
    flat = img.reshape(-1)
    got = oracle_mod.convolute(flat, flat[1:], flat[2:], w, 2, 32, 32, 2, 8, pad_mode=pad_mode, in_scale=1 / 127.5,
                               in_bias=-1.0, scale=sc, shift=sh, act=oracle_mod.ACT_RELU6, pix_stride=3,
                               img_stride=32 * 32 * 3)
    x = torch.from_numpy(img).permute(0, 3, 1, 2).double() * float(np.float32(1 / 127.5)) - 1.0
    ref = _ep(F.conv2d(_pad(x, 2, pad_mode), torch.from_numpy(w).double(), stride=2), sc, sh)
    np.testing.assert_allclose(got, ref.numpy(), rtol=2e-5, atol=2e-5)
    # planar call (three planes, MobileNet.c:218-246) gives the same numbers
    planes = [np.ascontiguousarray(img[..., k]) for k in range(3)]
    got2 = oracle_mod.convolute(planes[0], planes[1], planes[2], w, 2, 32, 32, 2, 8, pad_mode=pad_mode,
                                in_scale=1 / 127.5, in_bias=-1.0, scale=sc, shift=sh, act=oracle_mod.ACT_RELU6)
    assert np.array_equal(got, got2)


def test_pointwise_pool_fc_softmax(oracle_mod):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((3, 40, 7, 7)).astype(np.float32)
    w = rng.standard_normal((24, 40)).astype(np.float32)
    sc, sh = (0.5 + rng.random(24)).astype(np.float32), rng.standard_normal(24).astype(np.float32)
    got = oracle_mod.pointwise(x, w, 24, scale=sc, shift=sh, act=oracle_mod.ACT_RELU6)
    ref = _ep(F.conv2d(torch.from_numpy(x).double(), torch.from_numpy(w).double().view(24, 40, 1, 1)), sc, sh)
    np.testing.assert_allclose(got, ref.numpy(), rtol=2e-6, atol=2e-6)
    pooled = oracle_mod.pool(got)
    np.testing.assert_allclose(pooled, F.avg_pool2d(ref, 7).numpy().reshape(3, 24), rtol=2e-6, atol=2e-6)
    logits = rng.standard_normal((3, 1000)).astype(np.float32) * 3
    prob, top1, p1 = oracle_mod.softmax_argmax(logits)
    rp = F.softmax(torch.from_numpy(logits).double(), dim=1).numpy()
    np.testing.assert_allclose(prob, rp, rtol=1e-12)
    assert np.array_equal(top1, rp.argmax(1)) and np.allclose(p1, rp.max(1))


def test_first_five_layers_of_the_net(oracle_mod, synth_net):
    """BASELINE config 1 (MobileNet_L5.c): the oracle's layer chain == torch's, seeded weights."""
    from mnv1_b200.layers import LAYERS
    w, sc, sh = synth_net
    img = synth.images(1)
    out, _ = oracle_mod.forward(img, w, sc, sh, last_layer=5)
    x = torch.from_numpy(img).permute(0, 3, 1, 2).double() * float(np.float32(1 / 127.5)) - 1.0
    for L in LAYERS[:5]:
        wl = torch.from_numpy(w[L.w_off:L.w_off + L.w_cnt]).double()
        s, t = sc[L.c_off:L.c_off + L.cout], sh[L.c_off:L.c_off + L.cout]
        if L.kind == 0:
            x = F.conv2d(_pad(x, 2, 1), wl.view(32, 3, 3, 3), stride=2)
        elif L.kind == 1:
            x = F.conv2d(_pad(x, L.stride, 1), wl.view(L.cout, 1, 3, 3), stride=L.stride, groups=L.cout)
        else:
            x = F.conv2d(x, wl.view(L.cout, L.cin, 1, 1))
        x = _ep(x, s, t)
    np.testing.assert_allclose(out, x.numpy(), rtol=1e-4, atol=1e-5)


def test_bf16_rounding_matches_torch(oracle_mod):
    x = np.random.default_rng(4).standard_normal(10000).astype(np.float32) * 100
    assert np.array_equal(oracle_mod.round_bf16(x), torch.from_numpy(x).bfloat16().float().numpy())
    assert np.array_equal(synth._round_bf16(x), oracle_mod.round_bf16(x))
