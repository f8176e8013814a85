"""GPU parity at the BENCHED configuration (BASELINE config 4: batch 256, bf16) and on the persistent
multi-round paths of the tensor-core kernels — the holes VERDICT r01 / ADVICE r01 name:

* all 256 images of a batch-256 forward vs the oracle (per-layer taps + logits), not 3 of them;
* `pointwise_pair_kernel` with M large enough that every CTA pair runs >= 6 tiles (ring phase wrap,
  two-stage TMEM accumulator flip, ragged last round), against the oracle;
* integer known-answer tests that reach `pointwise_pair_kernel`, `depthwise_ring_kernel`,
  `depthwise_cw_kernel` and the fused block — three-way bit-exact (kernel.cl literal = oracle = CUDA);
* identical top-1 asserted UNCONDITIONALLY on 256 images, with FC weights constructed so that every
  oracle margin is far from a tie.
"""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

BF16_TOL = 2.0 ** -7


@pytest.fixture(scope="module")
def mn():
    import mnv1_b200  # noqa: F401
    from mnv1_b200 import binding
    return binding


def _ints(seed, shape, lo, hi):
    from mnv1_b200 import synth
    return synth.kat_ints(seed, int(np.prod(shape)), lo, hi).reshape(shape)


def _net_ctx(mn, dtype, w, sc, sh):
    c = mn.Context(0, dtype)
    c.set_pad_mode(mn.PAD_TFSAME)
    c.set_input_transform(1 / 127.5, -1.0)
    c.set_weights(w, sc, sh, mn.ACT_RELU6)
    return c


# ------------------------------------------------------------------ pair kernel, many rounds per cluster
@pytest.mark.parametrize("cin,cout,h,n", [
    (512, 512, 14, 256),    # layer 15 at the benched batch: M = 50176 -> 392 units on 74 pairs, 6 rounds, ragged last
    (1024, 1024, 7, 256),   # layer 27: M = 12544 -> 49 m-pairs x 4 n-tiles = 196 units, 16 k-blocks per tile
    (512, 1024, 7, 256),    # layer 25
    (256, 512, 14, 256),    # layer 13 (K = 256: the single-CTA tcgen05 kernel, 4 k-blocks, ring wraps every tile)
    (512, 512, 14, 37),     # ragged M: 7252 rows -> 57 m-tiles (odd), last tile 84 rows
    (512, 512, 14, 101),    # 19796 rows -> 155 m-tiles: odd tile count, units % clusters != 0, >= 2 rounds
])
def test_pointwise_pair_many_rounds(mn, oracle_mod, cin, cout, h, n):
    ctx = mn.Context(0, mn.BF16)
    rng = np.random.default_rng(70 + cin + cout + n)
    x = oracle_mod.round_bf16(rng.random((n, cin, h, h), dtype=np.float32) * 6)
    w = oracle_mod.round_bf16((rng.standard_normal((cout, cin)) * np.sqrt(2.0 / cin)).astype(np.float32))
    sc = (0.5 + rng.random(cout)).astype(np.float32)
    sh = (rng.standard_normal(cout) * 0.1).astype(np.float32)
    want = oracle_mod.pointwise(x, w, cout, scale=sc, shift=sh, act=oracle_mod.ACT_RELU6, rbf16=True)
    f = ctx.filter(mn.POINTWISE, w, cin, cout, sc, sh, mn.ACT_RELU6)
    xin = ctx.upload_planar(x)
    out = ctx.malloc(n, cout, h, h)
    ctx.pointwise(out, xin, f, h, h, cin, cout)
    assert ctx.last_kernel_name == ("pointwise_pair_kernel" if cin >= 512 else "pointwise_tc_kernel")   # K = 256: single CTA
    got = ctx.download_planar(out)
    assert rel_err(got, want) <= BF16_TOL
    # and twice more on the same context: a second launch must not depend on left-over barrier state
    ctx.pointwise(out, xin, f, h, h, cin, cout)
    assert np.array_equal(ctx.download_planar(out), got)
    ctx.close()


# ------------------------------------------------------------------ integer KATs on the big-shape kernels
@pytest.mark.parametrize("cin,cout,h,n", [(256, 256, 14, 3), (256, 512, 14, 2), (512, 512, 14, 9), (512, 1024, 7, 5),
                                          (1024, 1024, 7, 4)])
def test_kat_pointwise_pair_three_way(mn, oracle_mod, cin, cout, h, n):
    """kernel.cl `pointwise` (one launch per output channel, filtersize = Cin) == oracle == the CTA-pair
    tcgen05 kernel, bit for bit, on integer data whose sums stay below 256 (kernel.cl:94-114)."""
    ctx = mn.Context(0, mn.BF16)
    x = _ints(11 + cin, (n, cin, h, h), 0, 1).astype(np.uint8)
    w = _ints(12 + cout, (cout, cin), -1, 1).astype(np.int32)
    want = oracle_mod.pointwise(x.astype(np.float32), w.astype(np.float32), cout, act=oracle_mod.ACT_RELU)
    assert 0 < want.max() < 256
    if oracle_mod.literal():
        for i in range(min(n, 2)):
            lit = oracle_mod.lit_pointwise_per_channel(x[i], w)
            assert np.array_equal(lit.astype(np.float32), want[i])
    f = ctx.filter(mn.POINTWISE, w.astype(np.float32), cin, cout, act=mn.ACT_RELU)
    xin = ctx.upload_planar(x.astype(np.float32))
    out = ctx.malloc(n, cout, h, h)
    ctx.pointwise(out, xin, f, h, h, cin, cout)
    assert ctx.last_kernel_name == ("pointwise_pair_kernel" if cin >= 512 else "pointwise_tc_kernel")
    assert np.array_equal(ctx.download_planar(out), want)
    ctx.close()


@pytest.mark.parametrize("c,h,n,kernel", [(512, 14, 9, "depthwise_ring_kernel"), (1024, 7, 5, "depthwise_cw_kernel"),
                                          (512, 14, 300, "depthwise_ring_kernel")])
def test_kat_depthwise_big_shapes_three_way(mn, oracle_mod, c, h, n, kernel):
    """kernel.cl `depthwise` (per-channel launches) == oracle == the group-ring / channel-walk CUDA
    stencils on the real layer shapes (L14-L22, L26), bit for bit; n = 300 gives every persistent CTA
    several units."""
    ctx = mn.Context(0, mn.BF16)
    x = _ints(21 + c, (n, c, h, h), 0, 3).astype(np.uint8)
    w = _ints(22 + c, (c, 3, 3), -2, 2).astype(np.int32)
    want = oracle_mod.depthwise(x.astype(np.float32), w.astype(np.float32), 1, act=oracle_mod.ACT_RELU)
    assert want.max() < 256
    if oracle_mod.literal():
        lit = oracle_mod.lit_depthwise_per_channel(x[0], w, 1)   # valid for x < cols-1 (App. C D-08)
        assert np.array_equal(lit[:, :, :-1].astype(np.float32), want[0, :, :, :-1])
    f = ctx.filter(mn.DEPTHWISE, w.astype(np.float32), c, c, act=mn.ACT_RELU)
    xin = ctx.upload_planar(x.astype(np.float32))
    out = ctx.malloc(n, c, h, h)
    ctx.depthwise(out, xin, f, h, h, 3, 1, c)
    assert ctx.last_kernel_name == kernel
    assert np.array_equal(ctx.download_planar(out), want)
    ctx.close()


@pytest.mark.parametrize("c,cout,h,stride,n", [(128, 128, 56, 1, 3), (256, 256, 28, 1, 7), (512, 512, 14, 1, 9),
                                               (512, 1024, 14, 2, 5), (1024, 1024, 7, 1, 6)])
def test_kat_dw_pw_block_three_way(mn, oracle_mod, c, cout, h, stride, n):
    """A depthwise->pointwise block on integers through the FUSED kernels where one exists: literal
    depthwise (stride 1) -> literal pointwise == oracle chain == CUDA, bit for bit."""
    ctx = mn.Context(0, mn.BF16)
    x = _ints(31 + c, (n, c, h, h), 0, 1).astype(np.uint8)
    wd = _ints(32 + c, (c, 3, 3), 0, 1).astype(np.int32)
    wd[:, 1, 1] = 1
    wd[::2] = 0
    wd[::2, 1, 1] = 1                  # half of the channels are identity taps: keeps the sums small
    wp = _ints(33 + cout, (cout, c), -1, 1).astype(np.int32)
    wp[:, 1::2] = 0                    # contract over the identity channels only (values 0..1)
    ho = h // stride
    mid = oracle_mod.depthwise(x.astype(np.float32), wd.astype(np.float32), stride, act=oracle_mod.ACT_RELU)
    want = oracle_mod.pointwise(mid, wp.astype(np.float32), cout, act=oracle_mod.ACT_RELU)
    assert mid.max() <= 9 and 0 < want.max() < 256
    if oracle_mod.literal() and stride == 1:
        lmid = oracle_mod.lit_depthwise_per_channel(x[0], wd, 1)
        assert np.array_equal(lmid[:, :, :-1].astype(np.float32), mid[0, :, :, :-1])
        lit = oracle_mod.lit_pointwise_per_channel(mid[0].astype(np.uint8), wp)
        assert np.array_equal(lit.astype(np.float32), want[0])
    fd = ctx.filter(mn.DEPTHWISE, wd.astype(np.float32), c, c, act=mn.ACT_RELU)
    fp = ctx.filter(mn.POINTWISE, wp.astype(np.float32), c, cout, act=mn.ACT_RELU)
    xin = ctx.upload_planar(x.astype(np.float32))
    out = ctx.malloc(n, cout, ho, ho)
    try:
        ctx.dw_pw_block(out, xin, fd, fp, h, h, stride)
        assert ctx.last_kernel_name.startswith("fused_")
    except mn.Mnv1Error as e:          # no fused variant for this shape: the two kernels
        assert e.code == -6
        m = ctx.malloc(n, c, ho, ho)
        ctx.depthwise(m, xin, fd, h, h, 3, stride, c)
        ctx.pointwise(out, m, fp, ho, ho, c, cout)
    assert np.array_equal(ctx.download_planar(out), want)
    ctx.close()


# ------------------------------------------------------------------ batch 256, every image, vs the oracle
def _structured_images(n, seed=0xB10C):
    """u8 [n,224,224,3]: a 4x4 grid of coloured blocks plus mild noise per image.  Unlike uniform noise
    (synth.images), these give every image its own feature vector, so a classifier can tell them apart."""
    rng = np.random.default_rng(seed)
    blocks = rng.integers(0, 256, (n, 4, 4, 3)).astype(np.float32)
    img = np.repeat(np.repeat(blocks, 56, axis=1), 56, axis=2)
    img += rng.normal(0, 12, img.shape).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


def test_batch256_every_image_vs_oracle(mn, oracle_mod, synth_net):
    """BASELINE config 4 itself: batch 256, bf16.  Layers 11, 13, 15, 23, 27 and the logits of ALL 256
    images against the oracle at the same storage precision (tolerances of test_bf16_network), and every
    image's logits bit-identical to its own batch-1 run."""
    from mnv1_b200 import synth
    w, sc, sh = synth_net
    wq = synth.bf16_storage_weights(w)
    n = 256
    img = synth.images(n)
    taps_at = (11, 13, 15, 23, 27)
    logits, taps = oracle_mod.forward(img, wq, sc, sh, rbf16=1, taps=taps_at)
    c = _net_ctx(mn, mn.BF16, w, sc, sh)
    for k in taps_at:
        got = c.forward_upto(img, k)
        err = np.abs(got - taps[k]) / np.maximum(1.0, np.abs(taps[k]))
        # per image: the worst image must meet the bar, not the batch average
        q = np.quantile(err.reshape(n, -1), 0.999, axis=1)
        l2 = np.linalg.norm((got - taps[k]).reshape(n, -1), axis=1) / np.linalg.norm(taps[k].reshape(n, -1), axis=1)
        assert q.max() <= 16 * BF16_TOL / 2, f"layer {k}: worst image q99.9 {q.max()} (image {q.argmax()})"
        assert l2.max() <= 0.02, f"layer {k}: worst image rel L2 {l2.max()} (image {l2.argmax()})"
        del got
    lg, top1, p1 = c.forward(img)
    assert np.abs(lg - logits).max() <= 0.05
    assert np.array_equal(top1, lg.argmax(axis=1))
    for i in range(n):
        l1, t1, _ = c.forward(img[i:i + 1])
        assert np.array_equal(l1[0], lg[i]), f"image {i}: batch-256 logits differ from its batch-1 run"
        assert t1[0] == top1[i]
    c.close()


def test_top1_identity_unconditional_256(mn, oracle_mod, synth_net):
    """Identical top-1 on all 256 images, no margin mask.  The conv stack is the seeded synthetic net; the
    FC layer is CONSTRUCTED ("weights so top-1 margins are not ties"): class t(i) = 3*i + 1
    gets the (ridge-regularised) interpolating filter of image i's centred oracle feature vector, every other class a small random
    row, so the oracle's top-1 margin is large for every image and the 256 predictions are 256 different
    classes (MobileNet.c:2783-2792 argmax, 0-based)."""
    from mnv1_b200 import synth
    from mnv1_b200.layers import LAYERS
    w, sc, sh = synth_net
    n = 256
    img = _structured_images(n)
    wq = synth.bf16_storage_weights(w)
    pooled, _ = oracle_mod.forward(img, wq, sc, sh, rbf16=1, last_layer=28)
    pooled = pooled.reshape(n, 1024).astype(np.float64)
    mean = pooled.mean(axis=0)
    dev = pooled - mean
    rng = np.random.default_rng(77)
    wfc = rng.standard_normal((1000, 1024)) * 1e-3
    bias = np.zeros(1000)
    target = 3 * np.arange(n) + 1
    # ridge-regularised interpolation: logit(image i, class t(j)) = 8 [K (K + lam I)^-1]_ij ~ 8 delta_ij
    K = dev @ dev.T
    lam = 1e-2 * np.trace(K) / n
    wfc[target] = 8.0 * np.linalg.solve(K + lam * np.eye(n), dev)
    wfc = oracle_mod.round_bf16(wfc.astype(np.float32))
    bias[target] = -(wfc[target].astype(np.float64) @ mean)
    fc = LAYERS[28]
    w2, sh2 = w.copy(), sh.copy()
    w2[fc.w_off:fc.w_off + fc.w_cnt] = wfc.reshape(-1)
    sh2[fc.c_off:fc.c_off + 1000] = bias.astype(np.float32)
    logits, _ = oracle_mod.forward(img, synth.bf16_storage_weights(w2), sc, sh2, rbf16=1)
    _, otop1, _ = oracle_mod.softmax_argmax(logits)
    srt = np.sort(logits, axis=1)
    margin = srt[:, -1] - srt[:, -2]
    assert np.array_equal(otop1, target)            # the construction works: 256 distinct classes
    assert margin.min() > 0.5, f"weakest oracle margin {margin.min()}"
    c = _net_ctx(mn, mn.BF16, w2, sc, sh2)
    lg, top1, p1 = c.forward(img)
    assert np.array_equal(top1, otop1)              # unconditional: every one of the 256 images
    assert np.abs(lg - logits).max() <= 0.5         # winning logit ~8, |W_fc| ~ 8 x a unit filter: 6 % of it
    c.close()
    c32 = _net_ctx(mn, mn.F32, w2, sc, sh2)
    logits32, _ = oracle_mod.forward(img[:8], w2, sc, sh2)
    lg32, top32, _ = c32.forward(img[:8])
    assert np.array_equal(top32, target[:8])
    assert rel_err(lg32, logits32) <= 1e-3          # |W_fc| is ~100x a trained layer's: the 1e-4 per-layer bar scales with it
    c32.close()


def test_replan_with_tickets_in_flight(mn, synth_net):
    """a submit with n > plan_batch while earlier tickets are outstanding re-plans the arena;
    the earlier batches' results must still arrive in the caller's (pageable) arrays."""
    from mnv1_b200 import synth
    w, sc, sh = synth_net
    c = _net_ctx(mn, mn.BF16, w, sc, sh)
    small, big = synth.images(4), synth.images(24, first=4)
    want_s, want_st, _ = c.forward(small)
    c2 = _net_ctx(mn, mn.BF16, w, sc, sh)          # fresh context: plan_batch grows 4 -> 24 mid-flight
    ls, ts, ps = np.full((4, 1000), np.nan, np.float32), np.full(4, -1, np.int32), np.empty(4, np.float32)
    lb, tb, pb = np.empty((24, 1000), np.float32), np.empty(24, np.int32), np.empty(24, np.float32)
    t0 = c2.forward_submit(small.ctypes.data, 4, ls.ctypes.data, ts.ctypes.data, ps.ctypes.data)
    t1 = c2.forward_submit(big.ctypes.data, 24, lb.ctypes.data, tb.ctypes.data, pb.ctypes.data)
    c2.forward_wait(t0)
    c2.forward_wait(t1)
    assert np.array_equal(ls, want_s) and np.array_equal(ts, want_st)
    want_b, want_bt, _ = c.forward(big)
    assert np.array_equal(lb, want_b) and np.array_equal(tb, want_bt)
    c.close(); c2.close()


def test_two_contexts_two_devices_one_process(mn, synth_net):
    """ADVICE r01 / SURVEY §8b "one ctx per GPU, distinct ctxs independent": two contexts on two GPUs driven
    from ONE process and thread, interleaved, with the thread's current device left on GPU 0."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from mnv1_b200 import synth
    w, sc, sh = synth_net
    c0 = _net_ctx(mn, mn.BF16, w, sc, sh)
    c1 = mn.Context(1, mn.BF16)
    c1.set_pad_mode(mn.PAD_TFSAME); c1.set_input_transform(1 / 127.5, -1.0); c1.set_weights(w, sc, sh, mn.ACT_RELU6)
    img = synth.images(16)
    a0, t0, _ = c0.forward(img)
    a1, t1, _ = c1.forward(img)
    b0, _, _ = c0.forward(img[:5])
    assert torch.cuda.current_device() == 0
    assert np.array_equal(a0, a1) and np.array_equal(t0, t1) and np.array_equal(b0, a0[:5])
    c0.close(); c1.close()
