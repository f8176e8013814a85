"""Integer "reference-faithful" contexts (MNV1_U8) on the GPU.

u8 activations x s8 filters -> s32 (tcgen05.mma.kind::i8 for pointwise / FC, DP4A stencils for depthwise and the
stem) -> ReLU -> u8.  This is the one mode where the reference ITSELF pins real-size layers: with the wrapping
store, no bias and no shift, every layer must equal the unchanged kernel.cl (oracle/_ref, one launch per output
channel) bit for bit on full-range u8 data — depthwise stride 1 (all but the wrapping right column, App. C
D-08), pointwise, FC, pool — and the oracle's integer mode everywhere (stride 2, borders, stem, saturating
store, bias / shift, the 29-layer chain).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mn():
    import mnv1_b200  # noqa: F401
    from mnv1_b200 import binding
    return binding


@pytest.fixture(scope="module", params=["wrap", "saturate"])
def ictx(request, mn):
    c = mn.Context(0, mn.U8)
    c.set_u8_store(request.param == "wrap")
    c.store = request.param
    yield c
    c.close()


def _store(oracle_mod, ctx):
    return oracle_mod.STORE_U8_WRAP if ctx.store == "wrap" else oracle_mod.STORE_U8_SAT


@pytest.mark.parametrize("cin,cout,h,n", [(32, 64, 112, 1), (64, 128, 56, 2), (128, 128, 56, 1), (256, 512, 14, 3),
                                          (512, 512, 14, 7), (1024, 1024, 7, 5), (16, 16, 5, 2), (48, 40, 6, 1)])
def test_pointwise_u8_full_range(ictx, mn, oracle_mod, cin, cout, h, n):
    """kernel.cl `pointwise` (literal, filtersize = Cin, one launch per output channel) == oracle == tcgen05 kind::i8
    on full-range data: sums reach +-4 M, the u8 store wraps (or saturates)."""
    rng = np.random.default_rng(cin * 7 + cout)
    x = rng.integers(0, 256, (n, cin, h, h), dtype=np.uint8)
    w = rng.integers(-128, 128, (cout, cin)).astype(np.int32)
    want = oracle_mod.pointwise(x.astype(np.float32), w.astype(np.float32), cout, act=oracle_mod.ACT_RELU,
                                rbf16=_store(oracle_mod, ictx)).astype(np.uint8)
    if ictx.store == "wrap" and oracle_mod.literal():
        assert np.array_equal(oracle_mod.lit_pointwise_per_channel(x[0], w), want[0])
    f = ictx.filter(mn.POINTWISE, w.astype(np.float32), cin, cout, act=mn.ACT_RELU)
    xin = ictx.upload_planar_u8(x)
    out = ictx.malloc(n, cout, h, h)
    ictx.pointwise(out, xin, f, h, h, cin, cout)
    assert ictx.last_kernel_name == "pointwise_i8_kernel"
    assert np.array_equal(ictx.download_planar_u8(out), want)


@pytest.mark.parametrize("c,h,stride,pad,n", [(32, 112, 1, 0, 1), (64, 112, 2, 0, 1), (64, 112, 2, 1, 2), (512, 14, 1, 0, 5),
                                              (512, 14, 2, 1, 3), (1024, 7, 1, 0, 4), (8, 9, 1, 0, 2),
                                              # every tile shape of the shared-memory kernel (>= 128 channels), both paddings
                                              (128, 56, 1, 0, 1), (128, 56, 1, 1, 2), (128, 56, 2, 0, 1), (256, 28, 1, 1, 1),
                                              (256, 28, 2, 0, 2), (256, 28, 2, 1, 1), (512, 14, 1, 1, 2), (1024, 7, 1, 1, 3),
                                              (384, 20, 1, 0, 1), (128, 10, 2, 1, 3)])
def test_depthwise_u8_full_range(ictx, mn, oracle_mod, c, h, stride, pad, n):
    if h % stride:
        pytest.skip("odd size with stride 2")
    rng = np.random.default_rng(c + h + stride)
    x = rng.integers(0, 256, (n, c, h, h), dtype=np.uint8)
    w = rng.integers(-128, 128, (c, 3, 3)).astype(np.int32)
    want = oracle_mod.depthwise(x.astype(np.float32), w.astype(np.float32), stride, pad_mode=pad, act=oracle_mod.ACT_RELU,
                                rbf16=_store(oracle_mod, ictx)).astype(np.uint8)
    if ictx.store == "wrap" and stride == 1 and oracle_mod.literal():
        lit = oracle_mod.lit_depthwise_per_channel(x[0], w, 1)
        assert np.array_equal(lit[:, :, :-1], want[0, :, :, :-1])
    ictx.set_pad_mode(pad)
    f = ictx.filter(mn.DEPTHWISE, w.astype(np.float32), c, c, act=mn.ACT_RELU)
    xin = ictx.upload_planar_u8(x)
    out = ictx.malloc(n, c, h // stride, h // stride)
    ictx.depthwise(out, xin, f, h, h, 3, stride, c)
    ictx.set_pad_mode(mn.PAD_REF)
    assert ictx.last_kernel_name == "depthwise_u8_kernel"
    assert np.array_equal(ictx.download_planar_u8(out), want)


@pytest.mark.parametrize("pad", [0, 1])
def test_stem_u8(ictx, mn, oracle_mod, pad):
    """`convolute` on the raw u8 pixels: three planes (the reference's calling convention) and the interleaved payload"""
    rng = np.random.default_rng(5)
    n = 2
    img = rng.integers(0, 256, (n, 224, 224, 3), dtype=np.uint8)
    w = rng.integers(-128, 128, (32, 3, 3, 3)).astype(np.int32)
    want = oracle_mod.convolute(img, img.reshape(-1)[1:], img.reshape(-1)[2:], w.astype(np.float32), n, 224, 224, 2, 32,
                                pad_mode=pad, act=oracle_mod.ACT_RELU, rbf16=_store(oracle_mod, ictx), pix_stride=3,
                                img_stride=224 * 224 * 3).astype(np.uint8)
    ictx.set_pad_mode(pad)
    f = ictx.filter(mn.CONVOLUTE, w.astype(np.float32), 3, 32, act=mn.ACT_RELU)
    out = ictx.malloc(n, 32, 112, 112)
    ictx.convolute_rgb(out, ictx.upload_u8(img), f, 224, 224, 3, 2, 32)
    assert ictx.last_kernel_name == "stem_i8_rows_kernel"      # interleaved payload: tcgen05 kind::i8
    assert np.array_equal(ictx.download_planar_u8(out), want)
    planes = [ictx.upload_u8(np.ascontiguousarray(img[..., k])) for k in range(3)]
    out2 = ictx.malloc(n, 32, 112, 112)
    ictx.convolute(out2, planes[0], planes[1], planes[2], f, 224, 224, 3, 2, 32)
    assert ictx.last_kernel_name == "stem_u8_kernel"           # three planes: DP4A
    assert np.array_equal(ictx.download_planar_u8(out2), want)
    ictx.set_pad_mode(mn.PAD_REF)


def test_pool_and_fc_u8(ictx, mn, oracle_mod):
    rng = np.random.default_rng(6)
    n = 3
    x = rng.integers(0, 256, (n, 1024, 7, 7), dtype=np.uint8)
    want_pool = oracle_mod.pool(x.astype(np.float32), truncate=True).astype(np.uint8)
    if oracle_mod.literal():
        assert np.array_equal(oracle_mod.lit_pool_per_channel(x[0]), want_pool[0])
    xin = ictx.upload_planar_u8(x)
    pooled = ictx.malloc(n, 1024, 1, 1)
    ictx.pool(pooled, xin, 7, 7, 7, 1024)
    got_pool = ictx.download_planar_u8(pooled).reshape(n, 1024)
    assert np.array_equal(got_pool, want_pool)
    w = rng.integers(-128, 128, (1000, 1024)).astype(np.int32)
    want_fc = oracle_mod.pointwise(got_pool.reshape(n, 1024, 1, 1).astype(np.float32), w.astype(np.float32), 1000,
                                   act=oracle_mod.ACT_RELU, rbf16=_store(oracle_mod, ictx)).reshape(n, 1000).astype(np.uint8)
    if ictx.store == "wrap" and oracle_mod.literal():
        lit = oracle_mod.lit_pointwise_per_channel(got_pool[0].reshape(1024, 1, 1), w)
        assert np.array_equal(lit.reshape(1000), want_fc[0])
    f = ictx.filter(mn.FC, w.astype(np.float32), 1024, 1000, act=mn.ACT_RELU)
    logits = ictx.malloc(n, 1000, 1, 1)
    ictx.pointwise(logits, pooled, f, 1, 1, 1024, 1000)     # Cout = 1000: ragged n-tile, byte-wise store path
    got_fc = ictx.download_planar_u8(logits).reshape(n, 1000)
    assert np.array_equal(got_fc, want_fc)
    prob, top1, p1 = ictx.softmax(logits, 1000)             # MobileNet.c:2769-2792 over the u8 logits
    oprob, otop1, op1 = oracle_mod.softmax_argmax(got_fc.astype(np.float32))
    assert np.array_equal(top1, otop1) and np.allclose(p1, op1, rtol=1e-4)


def test_bias_and_shift_requantise(mn, oracle_mod):
    """out = store(relu(acc + bias) >> s): the saturating requantisation that keeps a deep integer chain in range"""
    ctx = mn.Context(0, mn.U8)
    rng = np.random.default_rng(8)
    cin, cout, h, n, s = 256, 128, 14, 2, 9
    x = rng.integers(0, 256, (n, cin, h, h), dtype=np.uint8)
    w = rng.integers(-128, 128, (cout, cin)).astype(np.int32)
    bias = rng.integers(-20000, 20000, cout).astype(np.float32)
    sc = np.full(cout, 2.0 ** -s, np.float32)
    want = oracle_mod.pointwise(x.astype(np.float32), w.astype(np.float32), cout, scale=sc, shift=bias, act=oracle_mod.ACT_RELU,
                                rbf16=oracle_mod.STORE_U8_SAT).astype(np.uint8)
    assert 0 < (want == 255).mean() < 0.5 and (want > 0).mean() > 0.2     # the test exercises both clamps and the middle
    f = ctx.filter(mn.POINTWISE, w.astype(np.float32), cin, cout, sc, bias, mn.ACT_RELU)
    out = ctx.malloc(n, cout, h, h)
    ctx.pointwise(out, ctx.upload_planar_u8(x), f, h, h, cin, cout)
    assert np.array_equal(ctx.download_planar_u8(out), want)
    with pytest.raises(mn.Mnv1Error):
        ctx.filter(mn.POINTWISE, w.astype(np.float32) + 0.5, cin, cout)            # not integers
    with pytest.raises(mn.Mnv1Error):
        ctx.filter(mn.POINTWISE, w.astype(np.float32), cin, cout, np.full(cout, 0.3, np.float32))   # not a power of two
    ctx.close()


@pytest.mark.parametrize("n", [1, 5])
def test_integer_network_29_layers(mn, oracle_mod, n):
    """The whole MobileNet.c schedule in the reference's integers: seeded s8 filters, per-layer shift so that the
    maps stay in range, saturating store.  Every tapped layer and the u8 logits equal the oracle bit for bit;
    top-1 identical.  (BASELINE configs 1-3 in integer arithmetic.)"""
    from mnv1_b200 import synth
    from mnv1_b200.layers import LAYERS, TOTAL_WEIGHTS, TOTAL_CHANNELS, DEPTHWISE, STEM, FC, POOL
    w = synth.kat_ints(99, TOTAL_WEIGHTS, -127, 127).astype(np.float32)
    sc = np.ones(TOTAL_CHANNELS, np.float32)
    sh = np.zeros(TOTAL_CHANNELS, np.float32)
    for L in LAYERS:
        if L.kind == POOL:
            continue
        fan = 27 if L.kind == STEM else 9 if L.kind == DEPTHWISE else L.cin
        s = int(np.ceil(np.log2(np.sqrt(fan) * 74 * 1.2)))                # keeps ~the input's spread after the shift
        if L.kind != FC:
            sc[L.c_off:L.c_off + L.cout] = 2.0 ** -s
            sh[L.c_off:L.c_off + L.cout] = synth.kat_ints(100 + L.index, L.cout, 0, 1 << (s + 5))
        else:
            sc[L.c_off:L.c_off + L.cout] = 2.0 ** -s
    img = synth.images(n)
    taps_at = (1, 2, 3, 5, 12, 13, 24, 27, 28)
    logits, taps = oracle_mod.forward(img, w, sc, sh, act=oracle_mod.ACT_RELU, rbf16=oracle_mod.STORE_U8_SAT,
                                      in_scale=1.0, in_bias=0.0, taps=taps_at)
    assert len(np.unique(taps[27])) > 30 and len(np.unique(logits)) > 20   # the chain did not collapse to 0 / 255
    c = mn.Context(0, mn.U8)
    c.set_pad_mode(mn.PAD_TFSAME)
    # the FC layer's scale slot: set_weights passes no scale to the FC, so fold nothing there
    c.set_weights(w, sc, sh, mn.ACT_RELU)
    for k in taps_at:
        got = c.forward_upto(img, k)
        assert np.array_equal(got.reshape(taps[k].shape), taps[k]), f"layer {k}"
    lg, top1, p1 = c.forward(img)
    assert np.array_equal(lg, logits)
    _, otop1, op1 = oracle_mod.softmax_argmax(logits)
    assert np.array_equal(top1, otop1) and np.allclose(p1, op1, rtol=1e-4)
    lg2, _, _ = c.forward(img)          # graph replay
    assert np.array_equal(lg2, lg)
    c.close()
