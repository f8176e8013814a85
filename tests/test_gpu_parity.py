"""GPU parity tests: CUDA kernels (through the C-ABI, via ctypes) vs the CPU oracle.

Bars (BASELINE.json north_star): integer KATs bit-exact; fp32 within 1e-4 relative per layer
(|a-b| <= 1e-4 * max(1,|b|)); bf16 within one bf16 ulp per layer when both sides get the same
input (|a-b| <= 2^-7 * max(1,|b|)), looser and stated for the end-to-end network.
"""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_TOL = 2.0 ** -7


@pytest.fixture(scope="module")
def mn():
    import mnv1_b200  # noqa: F401
    from mnv1_b200 import binding
    return binding


@pytest.fixture(scope="module", params=["f32", "bf16"])
def ctx(request, mn):
    c = mn.Context(0, mn.F32 if request.param == "f32" else mn.BF16)
    yield c
    c.close()


def _ints(seed, shape, lo, hi):
    from mnv1_b200 import synth
    return synth.kat_ints(seed, int(np.prod(shape)), lo, hi).reshape(shape)


# --------------------------------------------------------------------------- KATs (exact)
@pytest.mark.parametrize("c,h", [(8, 14), (32, 28), (64, 7)])
def test_kat_depthwise_three_way(ctx, mn, oracle_mod, c, h):
    """kernel.cl `depthwise` (per-channel launches) == oracle == CUDA, bit for bit."""
    x = _ints(1, (c, h, h), 0, 3).astype(np.uint8)
    w = _ints(2, (c, 3, 3), -2, 2).astype(np.int32)
    want = oracle_mod.depthwise(x[None].astype(np.float32), w.astype(np.float32), 1, act=oracle_mod.ACT_RELU)
    lit = oracle_mod.lit_depthwise_per_channel(x, w, 1) if oracle_mod.literal() else None
    if lit is not None:  # literal is valid for x < cols-1 (right column wraps, App. C D-08)
        assert np.array_equal(lit[:, :, :-1].astype(np.float32), want[0, :, :, :-1])
    f = ctx.filter(mn.DEPTHWISE, w.astype(np.float32), c, c, act=mn.ACT_RELU)
    xin = ctx.upload_planar(x[None].astype(np.float32))
    out = ctx.malloc(1, c, h, h)
    ctx.depthwise(out, xin, f, h, h, 3, 1, c)
    got = ctx.download_planar(out)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("cin,cout,h,hi", [(32, 64, 14, 3), (64, 128, 7, 1), (8, 16, 5, 3), (128, 256, 12, 1)])
def test_kat_pointwise_three_way(ctx, mn, oracle_mod, cin, cout, h, hi):
    """kernel.cl `pointwise` (filtersize = Cin) == oracle == CUDA (tcgen05 on bf16 contexts)."""
    x = _ints(3, (cin, h, h), 0, hi).astype(np.uint8)
    w = _ints(4, (cout, cin), -1, 1).astype(np.int32)
    want = oracle_mod.pointwise(x[None].astype(np.float32), w.astype(np.float32), cout, act=oracle_mod.ACT_RELU)
    assert want.max() < 256
    if oracle_mod.literal():
        lit = oracle_mod.lit_pointwise_per_channel(x, w)
        assert np.array_equal(lit.astype(np.float32), want[0])
    f = ctx.filter(mn.POINTWISE, w.astype(np.float32), cin, cout, act=mn.ACT_RELU)
    xin = ctx.upload_planar(x[None].astype(np.float32))
    out = ctx.malloc(1, cout, h, h)
    ctx.pointwise(out, xin, f, h, h, cin, cout)
    got = ctx.download_planar(out)
    if ctx.dtype == mn.BF16 and cout % 64 == 0 and cin % 8 == 0:
        assert ctx.last_kernel_name in ("pointwise_tc_kernel", "pointwise_pair_kernel")   # tcgen05 path, never the SIMT GEMM
    assert np.array_equal(got, want)


def test_kat_pool_three_way(ctx, mn, oracle_mod):
    c = 64
    x = (_ints(5, (c, 7, 7), 0, 4) * 49 // 4).astype(np.uint8)  # any u8; compare the means
    x = _ints(5, (c, 7, 7), 0, 5).astype(np.uint8)
    x[:, :, :] = x[:, :1, :1]  # constant planes: the mean is an exact integer in every dtype
    want = oracle_mod.pool(x[None].astype(np.float32), truncate=True)
    if oracle_mod.literal():
        assert np.array_equal(oracle_mod.lit_pool_per_channel(x).astype(np.float32), want[0])
    xin = ctx.upload_planar(x[None].astype(np.float32))
    out = ctx.malloc(1, c, 1, 1)
    ctx.pool(out, xin, 7, 7, 7, c)
    got = ctx.download_planar(out).reshape(1, c)
    assert np.array_equal(got, want)


# --------------------------------------------------------------------------- per-layer parity
def _tol(ctx, mn):
    return FP32_TOL if ctx.dtype == mn.F32 else BF16_TOL


def _prep(ctx, mn, oracle_mod, x):
    """What the device will actually hold: bf16 contexts round the activations on upload."""
    return oracle_mod.round_bf16(x) if ctx.dtype == mn.BF16 else x


@pytest.mark.parametrize("c,h,stride,pad", [(32, 112, 1, 0), (64, 112, 2, 0), (64, 112, 2, 1), (128, 56, 1, 1),
                                            (256, 28, 2, 1), (512, 14, 1, 0), (512, 14, 2, 1), (1024, 7, 1, 0),
                                            (1024, 7, 1, 1), (16, 10, 2, 0), (8, 9, 1, 0)])
def test_depthwise_layer(ctx, mn, oracle_mod, c, h, stride, pad):
    if h % stride:
        pytest.skip("odd size with stride 2")
    rng = np.random.default_rng(10 + c + h)
    n = 2 + pad          # 3 images in the TF-SAME cases: odd counts leave ragged last tiles / blocks
    x = _prep(ctx, mn, oracle_mod, (rng.random((n, c, h, h), dtype=np.float32) * 6))
    w = rng.standard_normal((c, 3, 3)).astype(np.float32) * 0.5
    sc = (0.5 + rng.random(c)).astype(np.float32)
    sh = (rng.standard_normal(c) * 0.1).astype(np.float32)
    want = oracle_mod.depthwise(x, w, stride, pad_mode=pad, scale=sc, shift=sh, act=oracle_mod.ACT_RELU6,
                                rbf16=ctx.dtype == mn.BF16)
    ctx.set_pad_mode(pad)
    f = ctx.filter(mn.DEPTHWISE, w, c, c, sc, sh, mn.ACT_RELU6)
    xin = ctx.upload_planar(x)
    out = ctx.malloc(n, c, h // stride, h // stride)
    ctx.depthwise(out, xin, f, h, h, 3, stride, c)
    got = ctx.download_planar(out)
    ctx.set_pad_mode(mn.PAD_REF)
    assert rel_err(got, want) <= _tol(ctx, mn)
    if ctx.dtype == mn.BF16 and (c, h, stride) == (1024, 7, 1):
        assert ctx.last_kernel_name == "depthwise_cw_kernel"     # thread-per-channel-pair kernel, layer 26
    if ctx.dtype == mn.BF16 and (c, h) == (512, 14):
        assert ctx.last_kernel_name == "depthwise_ring_kernel"


@pytest.mark.parametrize("cin,cout,h,n", [(32, 64, 112, 1), (64, 128, 56, 2), (128, 128, 56, 1), (128, 256, 28, 3),
                                          (256, 256, 28, 1), (256, 512, 14, 2), (512, 512, 14, 5), (512, 1024, 7, 3),
                                          (1024, 1024, 7, 2), (24, 40, 5, 1)])
def test_pointwise_layer(ctx, mn, oracle_mod, cin, cout, h, n):
    rng = np.random.default_rng(20 + cin + cout)
    x = _prep(ctx, mn, oracle_mod, rng.random((n, cin, h, h), dtype=np.float32) * 6)
    w = (rng.standard_normal((cout, cin)) * np.sqrt(2.0 / cin)).astype(np.float32)
    if ctx.dtype == mn.BF16:
        w = oracle_mod.round_bf16(w)
    sc = (0.5 + rng.random(cout)).astype(np.float32)
    sh = (rng.standard_normal(cout) * 0.1).astype(np.float32)
    want = oracle_mod.pointwise(x, w, cout, scale=sc, shift=sh, act=oracle_mod.ACT_RELU6, rbf16=ctx.dtype == mn.BF16)
    f = ctx.filter(mn.POINTWISE, w, cin, cout, sc, sh, mn.ACT_RELU6)
    xin = ctx.upload_planar(x)
    out = ctx.malloc(n, cout, h, h)
    ctx.pointwise(out, xin, f, h, h, cin, cout)
    got = ctx.download_planar(out)
    assert rel_err(got, want) <= _tol(ctx, mn)
    if ctx.dtype == mn.BF16 and cout % 64 == 0 and cin % 8 == 0:
        assert ctx.last_kernel_name in ("pointwise_tc_kernel", "pointwise_pair_kernel")   # tcgen05 path, never the SIMT GEMM
        # the CUDA-core GEMM must agree with the tensor-core one to the same tolerance
        out2 = ctx.malloc(n, cout, h, h)
        ctx.pointwise(out2, xin, f, h, h, cin, cout, simt=True)
        assert rel_err(ctx.download_planar(out2), want) <= _tol(ctx, mn)


@pytest.mark.parametrize("pad", [0, 1])
def test_stem_layer(ctx, mn, oracle_mod, pad):
    from mnv1_b200 import synth
    rng = np.random.default_rng(30)
    n = 2
    img = synth.images(n)
    w = (rng.standard_normal((32, 3, 3, 3)) * np.sqrt(2.0 / 27)).astype(np.float32)
    sc = (0.5 + rng.random(32)).astype(np.float32)
    sh = (rng.standard_normal(32) * 0.1).astype(np.float32)
    wo = w
    if ctx.dtype == mn.BF16:  # the tensor-core stem stores fp16(w * in_scale)
        wo = ((w * np.float32(1 / 127.5)).astype(np.float16).astype(np.float32) / np.float32(1 / 127.5))
    want = oracle_mod.convolute(img, img.reshape(-1)[1:], img.reshape(-1)[2:], wo, n, 224, 224, 2, 32, pad_mode=pad,
                                in_scale=1 / 127.5, in_bias=-1.0, scale=sc, shift=sh, act=oracle_mod.ACT_RELU6,
                                rbf16=ctx.dtype == mn.BF16, pix_stride=3, img_stride=224 * 224 * 3)
    ctx.set_pad_mode(pad)
    ctx.set_input_transform(1 / 127.5, -1.0)
    f = ctx.filter(mn.CONVOLUTE, w, 3, 32, sc, sh, mn.ACT_RELU6)
    out = ctx.malloc(n, 32, 112, 112)
    rgb = ctx.upload_u8(img)
    ctx.convolute_rgb(out, rgb, f, 224, 224, 3, 2, 32)
    got = ctx.download_planar(out)
    assert rel_err(got, want) <= _tol(ctx, mn)
    if ctx.dtype == mn.BF16:
        assert ctx.last_kernel_name == "stem_rows_kernel"   # row-staged tcgen05 stem, not the gather fallback
    # the reference's own calling convention: three separate planes (MobileNet.c:218-246)
    planes = [ctx.upload_u8(np.ascontiguousarray(img[..., k])) for k in range(3)]
    out2 = ctx.malloc(n, 32, 112, 112)
    ctx.convolute(out2, planes[0], planes[1], planes[2], f, 224, 224, 3, 2, 32)
    got2 = ctx.download_planar(out2)
    # planar and interleaved inputs run different kernels with different K orders inside the MMA:
    # same oracle tolerance, and at most a bf16 ulp apart from each other
    assert rel_err(got2, want) <= _tol(ctx, mn)
    assert rel_err(got2, got) <= _tol(ctx, mn)
    if ctx.dtype == mn.BF16:
        assert ctx.last_kernel_name == "stem_tc_kernel"
    ctx.set_pad_mode(mn.PAD_REF)
    ctx.set_input_transform(1.0, 0.0)


@pytest.mark.parametrize("rows,cols,n,kernel", [(32, 32, 3, "stem_rows_kernel"), (18, 96, 2, "stem_rows_kernel"),
                                                (64, 256, 5, "stem_rows_kernel"), (20, 40, 2, "stem_tc_kernel")])
@pytest.mark.parametrize("pad", [0, 1])
def test_stem_other_geometries(ctx, mn, oracle_mod, pad, rows, cols, n, kernel):
    """Row-staged stem on other image sizes (one CTA tile = one output row, ragged tile counts), and
    the gather kernel when a row is not a multiple of 16 bytes."""
    rng = np.random.default_rng(31 + rows + cols)
    img = rng.integers(0, 256, (n, rows, cols, 3), dtype=np.uint8)
    w = (rng.standard_normal((32, 3, 3, 3)) * np.sqrt(2.0 / 27)).astype(np.float32)
    sc = (0.5 + rng.random(32)).astype(np.float32)
    sh = (rng.standard_normal(32) * 0.1).astype(np.float32)
    wo = w
    if ctx.dtype == mn.BF16:
        wo = ((w * np.float32(1 / 127.5)).astype(np.float16).astype(np.float32) / np.float32(1 / 127.5))
    want = oracle_mod.convolute(img, img.reshape(-1)[1:], img.reshape(-1)[2:], wo, n, rows, cols, 2, 32, pad_mode=pad,
                                in_scale=1 / 127.5, in_bias=-1.0, scale=sc, shift=sh, act=oracle_mod.ACT_RELU6,
                                rbf16=ctx.dtype == mn.BF16, pix_stride=3, img_stride=rows * cols * 3)
    ctx.set_pad_mode(pad)
    ctx.set_input_transform(1 / 127.5, -1.0)
    try:
        f = ctx.filter(mn.CONVOLUTE, w, 3, 32, sc, sh, mn.ACT_RELU6)
        out = ctx.malloc(n, 32, rows // 2, cols // 2)
        ctx.convolute_rgb(out, ctx.upload_u8(img), f, rows, cols, 3, 2, 32)
        assert rel_err(ctx.download_planar(out), want) <= _tol(ctx, mn)
        if ctx.dtype == mn.BF16:
            assert ctx.last_kernel_name == kernel
    finally:
        ctx.set_pad_mode(mn.PAD_REF)
        ctx.set_input_transform(1.0, 0.0)


def test_pool_fc_softmax(ctx, mn, oracle_mod):
    rng = np.random.default_rng(40)
    n = 3
    x = _prep(ctx, mn, oracle_mod, rng.random((n, 1024, 7, 7), dtype=np.float32) * 6)
    want_pool = oracle_mod.pool(x, rbf16=ctx.dtype == mn.BF16)
    xin = ctx.upload_planar(x)
    pooled = ctx.malloc(n, 1024, 1, 1)
    ctx.pool(pooled, xin, 7, 7, 7, 1024)
    got_pool = ctx.download_planar(pooled).reshape(n, 1024)
    assert rel_err(got_pool, want_pool) <= _tol(ctx, mn)
    w = (rng.standard_normal((1000, 1024)) / 32).astype(np.float32)
    if ctx.dtype == mn.BF16:
        w = oracle_mod.round_bf16(w)
    bias = (rng.standard_normal(1000) * 0.01).astype(np.float32)
    want_fc = oracle_mod.pointwise(got_pool.reshape(n, 1024, 1, 1), w, 1000, shift=bias, act=oracle_mod.ACT_NONE,
                                   rbf16=ctx.dtype == mn.BF16).reshape(n, 1000)
    f = ctx.filter(mn.FC, w, 1024, 1000, None, bias, mn.ACT_NONE)
    logits = ctx.malloc(n, 1000, 1, 1)
    ctx.pointwise(logits, pooled, f, 1, 1, 1024, 1000)
    got_fc = ctx.download_planar(logits).reshape(n, 1000)
    assert rel_err(got_fc, want_fc) <= _tol(ctx, mn)
    prob, top1, p1 = ctx.softmax(logits, 1000)
    oprob, otop1, op1 = oracle_mod.softmax_argmax(got_fc)
    assert np.array_equal(top1, otop1)
    assert np.allclose(prob, oprob, rtol=1e-4, atol=1e-7)
    assert np.allclose(p1, op1, rtol=1e-4)


# --------------------------------------------------------------------------- whole network
def _net_ctx(mn, dtype, synth_net):
    w, sc, sh = synth_net
    c = mn.Context(0, dtype)
    c.set_pad_mode(mn.PAD_TFSAME)
    c.set_input_transform(1 / 127.5, -1.0)
    c.set_weights(w, sc, sh, mn.ACT_RELU6)
    return c


def test_fp32_network_per_layer(mn, oracle_mod, synth_net):
    """BASELINE configs 1-3: layers 1-5, 1-13, full net at N=1, fp32, <= 1e-4 per layer, same top-1."""
    from mnv1_b200 import synth
    w, sc, sh = synth_net
    img = synth.images(1)
    logits, taps = oracle_mod.forward(img, w, sc, sh, taps=range(1, 29))
    c = _net_ctx(mn, mn.F32, synth_net)
    worst = {}
    for k in list(range(1, 29)):
        got = c.forward_upto(img, k)
        worst[k] = rel_err(got.reshape(taps[k].shape), taps[k])
        assert worst[k] <= FP32_TOL, f"layer {k}: {worst[k]}"
    got_logits, top1, p1 = c.forward(img)
    assert rel_err(got_logits, logits) <= FP32_TOL
    _, otop1, op1 = oracle_mod.softmax_argmax(logits)
    assert np.array_equal(top1, otop1)
    assert abs(p1[0] - op1[0]) <= 1e-4 * op1[0] + 1e-7
    c.close()


def test_bf16_network(mn, oracle_mod, synth_net):
    """bf16 activations / pointwise+FC weights, fp32 accumulate.  Oracle run in the same
    storage precision (bf16-rounded pointwise/FC weights, outputs rounded per layer).
    Rounding flips propagate down the chain, so the stated tolerance grows with depth: per
    layer, relative L2 error <= 2% and 99.9% of elements within 16 bf16 ulp (2^-4 of
    max(1,|ref|)); logits within 0.05 absolute (logit spread is ~0.8); identical top-1
    wherever the oracle's top-1 margin exceeds 0.1."""
    from mnv1_b200 import synth
    w, sc, sh = synth_net
    wq = synth.bf16_storage_weights(w)  # bf16 pointwise/FC filters, fp16 scale-folded stem taps
    n = 4
    img = synth.images(n)
    logits, taps = oracle_mod.forward(img, wq, sc, sh, rbf16=1, taps=(1, 2, 3, 5, 13, 27))
    c = _net_ctx(mn, mn.BF16, synth_net)
    for k in (1, 2, 3, 5, 13, 27):
        got = c.forward_upto(img, k)
        err = np.abs(got - taps[k]) / np.maximum(1.0, np.abs(taps[k]))
        l2 = np.linalg.norm(got - taps[k]) / np.linalg.norm(taps[k])
        print(f"layer {k}: q99.9 rel err {np.quantile(err, 0.999):.4f}, rel L2 {l2:.5f}")
        assert np.quantile(err, 0.999) <= 16 * BF16_TOL / 2, f"layer {k}: q99.9 {np.quantile(err, 0.999)}"
        assert l2 <= 0.02, f"layer {k}: rel L2 {l2}"
    got_logits, top1, _ = c.forward(img)
    assert np.max(np.abs(got_logits - logits)) <= 0.05
    _, otop1, _ = oracle_mod.softmax_argmax(logits)
    srt = np.sort(logits, axis=1)
    margin = srt[:, -1] - srt[:, -2]
    assert np.array_equal(top1[margin > 0.1], otop1[margin > 0.1])
    c.close()


def test_batch_independence_and_graph(mn, synth_net):
    """Size-independent properties at batch 256 (BASELINE config 4): every image's logits are
    bit-identical whether it runs alone, eagerly, or inside a graph-replayed batch of 256."""
    from mnv1_b200 import synth
    c = _net_ctx(mn, mn.BF16, synth_net)
    img = synth.images(256)
    lg, top1, _ = c.forward(img)
    lg2, top1b, _ = c.forward(img)  # graph replay
    assert np.array_equal(lg, lg2) and np.array_equal(top1, top1b)
    for i in (0, 100, 255):
        l1, t1, _ = c.forward(img[i:i + 1])
        assert np.array_equal(l1[0], lg[i])
        assert t1[0] == top1[i]
    c.use_graph(False)
    lg3, _, _ = c.forward(img[:7])
    assert np.array_equal(lg3, lg[:7])
    assert np.isfinite(lg).all()
    c.close()


def test_synth_images_device_matches_numpy(mn):
    import torch
    from mnv1_b200 import synth
    c = mn.Context(0, mn.BF16)
    buf = torch.empty(3 * 224 * 224 * 3, dtype=torch.uint8, device="cuda")
    c.synth_images_device(buf.data_ptr(), 3, 5, synth.IMAGE_SEED)
    c.sync()
    assert np.array_equal(buf.cpu().numpy().reshape(3, 224, 224, 3), synth.images(3, first=5))
    c.close()


# --------------------------------------------------------------------------- edges / errors
def test_errors_and_empty(mn):
    c = mn.Context(0, mn.BF16)
    f = c.filter(mn.DEPTHWISE, np.zeros((8, 3, 3), np.float32), 8, 8)
    a = c.malloc(0, 8, 4, 4)
    b = c.malloc(0, 8, 4, 4)
    c.depthwise(b, a, f, 4, 4, 3, 1, 8)  # empty batch is a no-op
    x = c.malloc(1, 8, 4, 4)
    y = c.malloc(1, 8, 3, 3)
    with pytest.raises(mn.Mnv1Error) as e:
        c.depthwise(y, x, f, 4, 4, 3, 1, 8)  # wrong output shape
    assert e.value.code == -1
    with pytest.raises(mn.Mnv1Error) as e:
        c.depthwise(x, x, f, 4, 4, 5, 1, 8)  # 5x5 not implemented
    assert e.value.code == -6
    with pytest.raises(mn.Mnv1Error) as e:
        c.forward(np.zeros((1, 224, 224, 3), np.uint8))  # no weights
    assert e.value.code == -5
    c.close()


@pytest.mark.parametrize("depth", [2, 3, 5])
def test_pipelined_forward_matches_blocking(mn, synth_net, depth):
    """mnv1_forward_submit/_wait with `depth` batches outstanding (the library keeps three in flight;
    submitting more retires the oldest) returns the same bits as mnv1_forward, for pinned (torch) and
    pageable (numpy) caller buffers."""
    import torch
    from mnv1_b200 import synth
    c = _net_ctx(mn, mn.BF16, synth_net)
    n = 32
    batches = [synth.images(n, first=k * n) for k in range(7)]
    want = [c.forward(b) for b in batches]
    outs = []
    pending = []
    for k, b in enumerate(batches):
        pinned = k % 2 == 0
        if pinned:
            hi = torch.from_numpy(b.copy()).pin_memory()
            lg = torch.empty(n, 1000).pin_memory(); t1 = torch.empty(n, dtype=torch.int32).pin_memory()
            p1 = torch.empty(n).pin_memory()
            ptrs = (hi.data_ptr(), lg.data_ptr(), t1.data_ptr(), p1.data_ptr())
            outs.append((hi, lg, t1, p1))
        else:
            hi = np.ascontiguousarray(b); lg = np.empty((n, 1000), np.float32); t1 = np.empty(n, np.int32)
            p1 = np.empty(n, np.float32)
            ptrs = (hi.ctypes.data, lg.ctypes.data, t1.ctypes.data, p1.ctypes.data)
            outs.append((hi, lg, t1, p1))
        pending.append(c.forward_submit(ptrs[0], n, ptrs[1], ptrs[2], ptrs[3]))
        if len(pending) == depth:
            c.forward_wait(pending.pop(0))
    for t in pending:
        c.forward_wait(t)
    for (hi, lg, t1, p1), (wl, wt, wp) in zip(outs, want):
        lg = lg.numpy() if hasattr(lg, "numpy") and not isinstance(lg, np.ndarray) else lg
        t1 = t1.numpy() if not isinstance(t1, np.ndarray) else t1
        assert np.array_equal(lg, wl) and np.array_equal(t1, wt)
    c.close()


@pytest.mark.parametrize("c,cout,h,stride,pad,n", [
    # resident-filter kernel (fused_rb.cu): the five blocks of layers 2-11, both padding conventions
    (32, 64, 112, 1, 0, 2), (64, 128, 112, 2, 1, 2), (64, 128, 112, 2, 0, 1), (128, 128, 56, 1, 0, 3),
    (128, 256, 56, 2, 1, 3), (128, 256, 56, 2, 0, 2), (256, 256, 28, 1, 0, 5), (256, 256, 28, 1, 0, 40),
    # CTA-pair kernel with a streamed filter (fused_pair.cu): the 512-channel 14x14 blocks, layers 14-23; one image, fewer
    # images than clusters, several rounds per cluster with a ragged last round, the benched batch
    (512, 512, 14, 1, 0, 1), (512, 512, 14, 1, 1, 9), (512, 512, 14, 1, 0, 100), (512, 512, 14, 1, 0, 256)])
def test_fused_dw_pw_block(mn, oracle_mod, c, cout, h, stride, pad, n):
    """depthwise->pointwise fused kernels vs the oracle chain at bf16 storage precision, and vs the
    two separate CUDA kernels (same arithmetic in the same order: bit-identical)."""
    ctx = mn.Context(0, mn.BF16)
    ctx.set_pad_mode(pad)
    rng = np.random.default_rng(50 + c + cout + h)
    ho = h // stride
    x = oracle_mod.round_bf16(rng.random((n, c, h, h), dtype=np.float32) * 6)
    wd = rng.standard_normal((c, 3, 3)).astype(np.float32) * 0.5
    sd, td = (0.5 + rng.random(c)).astype(np.float32), (rng.standard_normal(c) * 0.1).astype(np.float32)
    wp = oracle_mod.round_bf16((rng.standard_normal((cout, c)) * np.sqrt(2.0 / c)).astype(np.float32))
    sp, tp = (0.5 + rng.random(cout)).astype(np.float32), (rng.standard_normal(cout) * 0.1).astype(np.float32)
    fd = ctx.filter(mn.DEPTHWISE, wd, c, c, sd, td, mn.ACT_RELU6)
    fp = ctx.filter(mn.POINTWISE, wp, c, cout, sp, tp, mn.ACT_RELU6)
    xin = ctx.upload_planar(x)
    out = ctx.malloc(n, cout, ho, ho)
    ctx.dw_pw_block(out, xin, fd, fp, h, h, stride)
    assert ctx.last_kernel_name == ("fused_pair_kernel" if c == 512 else "fused_dw_pw_kernel")
    got = ctx.download_planar(out)
    if n * h * h * c <= 8 << 20 or c == 512:   # the oracle chain on the small cases (and every CTA-pair case)
        mid = oracle_mod.depthwise(x, wd, stride, pad_mode=pad, scale=sd, shift=td, act=oracle_mod.ACT_RELU6, rbf16=True)
        want = oracle_mod.pointwise(mid, wp, cout, scale=sp, shift=tp, act=oracle_mod.ACT_RELU6, rbf16=True)
        err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
        # a 1-ulp flip of a depthwise value (fp32 vs double accumulation) moves a few outputs by an ulp
        assert err.max() <= 4 * BF16_TOL and np.quantile(err, 0.999) <= BF16_TOL
    # against the unfused CUDA kernels: same arithmetic, so (nearly) the same bits
    m = ctx.malloc(n, c, ho, ho)
    ctx.depthwise(m, xin, fd, h, h, 3, stride, c)
    out2 = ctx.malloc(n, cout, ho, ho)
    ctx.pointwise(out2, m, fp, ho, ho, c, cout)
    got2 = ctx.download_planar(out2)
    assert np.mean(got2 == got) > 0.999 and rel_err(got, got2) <= 2 * BF16_TOL
    ctx.close()


def test_fused_block_unsupported_shape(mn):
    """blocks whose filter does not fit in shared memory report MNV1_EUNSUPPORTED and launch nothing"""
    ctx = mn.Context(0, mn.BF16)
    c = 1024
    fd = ctx.filter(mn.DEPTHWISE, np.ones((c, 3, 3), np.float32), c, c, None, None, mn.ACT_RELU6)
    fp = ctx.filter(mn.POINTWISE, np.ones((c, c), np.float32), c, c, None, None, mn.ACT_RELU6)
    x = ctx.malloc(1, c, 7, 7)
    out = ctx.malloc(1, c, 7, 7)
    with pytest.raises(mn.Mnv1Error) as e:
        ctx.dw_pw_block(out, x, fd, fp, 7, 7, 1)
    assert e.value.code == -6   # MNV1_EUNSUPPORTED
    ctx.close()


def test_fused_layers_report(mn, synth_net):
    """mnv1_fused_layers: a bf16 context fuses layers 2-11 pairwise inside mnv1_forward (14-23 too with MNV1_FUSED_PAIR=1:
    the CTA-pair kernel is correct but slower than its two kernels), an fp32 context nothing"""
    import os
    c = _net_ctx(mn, mn.BF16, synth_net)
    f = c.fused_layers()
    want = [2, 4, 6, 8, 10] + ([14, 16, 18, 20, 22] if os.environ.get("MNV1_FUSED_PAIR") else [])
    assert [i + 1 for i in range(29) if f[i]] == want
    c.use_fused_blocks(False)
    assert not c.fused_layers().any()
    c.close()
    c32 = _net_ctx(mn, mn.F32, synth_net)
    assert not c32.fused_layers().any()
    c32.close()


def test_fused_and_unfused_network_agree(mn, synth_net):
    from mnv1_b200 import synth
    c = _net_ctx(mn, mn.BF16, synth_net)
    img = synth.images(8)
    lf, tf, _ = c.forward(img)
    c.use_fused_blocks(False)
    lu, tu, _ = c.forward(img)
    assert np.max(np.abs(lf - lu)) <= 0.02
    srt = np.sort(lu, axis=1)
    ok = (srt[:, -1] - srt[:, -2]) > 0.05
    assert np.array_equal(tf[ok], tu[ok])
    c.close()
