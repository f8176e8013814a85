"""Data-parallel mode of the C-ABI (mnv1_dp_*, mnv1_gather_*; SURVEY §8e, BASELINE config 5) on the GPU box.

One process, one context + worker thread per "rank".  With a single GPU the ranks share the device (the peer
stores then land in the same GPU's memory); with >= 2 GPUs they sit on different devices and the stores
cross NVLink.  Either way the check is the one the north_star asks for: the gathered logits of the sharded
run are bit-identical to a single-context run over the same global image indices.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mn():
    import mnv1_b200  # noqa: F401
    from mnv1_b200 import binding
    return binding


def _single(mn, synth_net, img):
    w, sc, sh = synth_net
    c = mn.Context(0, mn.BF16)
    c.set_pad_mode(mn.PAD_TFSAME); c.set_input_transform(1 / 127.5, -1.0); c.set_weights(w, sc, sh, mn.ACT_RELU6)
    out = c.forward(img)
    c.close()
    return out


def _devices(world):
    import torch
    g = torch.cuda.device_count()
    return [r % g for r in range(world)]


@pytest.mark.parametrize("world,n", [(2, 64), (3, 50), (8, 37), (4, 3)])
def test_dp_forward_host_matches_single_context(mn, synth_net, world, n):
    """host in / host out: contiguous shards (remainder on the low ranks; empty shards when n < world)."""
    from mnv1_b200 import synth
    w, sc, sh = synth_net
    img = synth.images(n)
    want_l, want_t, want_p = _single(mn, synth_net, img)
    dp = mn.DataParallel(_devices(world), mn.BF16, max_batch_per_gpu=32)
    assert dp.size == world
    dp.set_pad_mode(mn.PAD_TFSAME); dp.set_input_transform(1 / 127.5, -1.0); dp.set_weights(w, sc, sh, mn.ACT_RELU6)
    lg, t1, p1 = dp.forward(img)
    assert np.array_equal(lg, want_l) and np.array_equal(t1, want_t) and np.array_equal(p1, want_p)
    lg2, t2, _ = dp.forward(img)          # graph replay on every rank
    assert np.array_equal(lg2, want_l) and np.array_equal(t2, want_t)
    with pytest.raises(mn.Mnv1Error):
        dp.forward(synth.images(world * 32 + 1))   # exceeds n_devices * max_batch_per_gpu
    dp.close()


@pytest.mark.parametrize("world,per", [(2, 16), (4, 8)])
def test_dp_forward_device_gathers_by_peer_stores(mn, synth_net, world, per):
    """device in / device out: after mnv1_dp_forward_device EVERY rank's gather block holds all world*per rows
    (logits, top-1, top-1 probability), written by the head kernels of all ranks — no collective kernel."""
    import torch
    from mnv1_b200 import synth
    w, sc, sh = synth_net
    devs = _devices(world)
    img = synth.images(world * per)
    want_l, want_t, want_p = _single(mn, synth_net, img)
    dp = mn.DataParallel(devs, mn.BF16, max_batch_per_gpu=per)
    dp.set_pad_mode(mn.PAD_TFSAME); dp.set_input_transform(1 / 127.5, -1.0); dp.set_weights(w, sc, sh, mn.ACT_RELU6)
    d_imgs = [torch.from_numpy(img[r * per:(r + 1) * per].copy()).to(f"cuda:{devs[r]}") for r in range(world)]
    for rep in range(2):
        dp.forward_device([t.data_ptr() for t in d_imgs], per)
        for r in range(world):
            lp, tp, pp = dp.gather_ptrs(r)
            rows = world * per
            with torch.cuda.device(devs[r]):
                lg = _from_ptr(torch, lp, (rows, 1000), torch.float32, devs[r]).cpu().numpy()
                t1 = _from_ptr(torch, tp, (rows,), torch.int32, devs[r]).cpu().numpy()
                p1 = _from_ptr(torch, pp, (rows,), torch.float32, devs[r]).cpu().numpy()
            assert np.array_equal(lg, want_l), f"rank {r}: gathered logits differ from the single-context run"
            assert np.array_equal(t1, want_t) and np.array_equal(p1, want_p)
    dp.close()


def _from_ptr(torch, ptr, shape, dtype, dev):
    """a torch view of raw device memory owned by the library (via __cuda_array_interface__)"""
    n = int(np.prod(shape))
    typestr = {torch.float32: "<f4", torch.int32: "<i4"}[dtype]

    class _Holder:
        __cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(_Holder(), device=f"cuda:{dev}").view(*shape).clone()


def test_gather_between_two_plain_contexts(mn, synth_net):
    """the building blocks the dp group uses, driven by hand: two contexts, attach, forward_device"""
    import torch
    from mnv1_b200 import synth
    w, sc, sh = synth_net
    devs = _devices(2)
    ctxs = []
    for r in range(2):
        c = mn.Context(devs[r], mn.BF16)
        c.set_pad_mode(mn.PAD_TFSAME); c.set_input_transform(1 / 127.5, -1.0); c.set_weights(w, sc, sh, mn.ACT_RELU6)
        c.gather_create(2, r, 8)
        ctxs.append(c)
    ctxs[0].gather_attach(ctxs[1]); ctxs[1].gather_attach(ctxs[0])
    img = synth.images(16)
    want_l, want_t, _ = _single(mn, synth_net, img)
    outs = []
    for r in range(2):
        with torch.cuda.device(devs[r]):
            d = torch.from_numpy(img[8 * r:8 * r + 8].copy()).to(f"cuda:{devs[r]}")
            lg = torch.empty(8, 1000, device=f"cuda:{devs[r]}"); t1 = torch.empty(8, dtype=torch.int32, device=f"cuda:{devs[r]}")
            ctxs[r].forward_device(d.data_ptr(), 8, lg.data_ptr(), t1.data_ptr())
            outs.append((d, lg, t1))
    for c in ctxs:
        c.sync()
    for r in range(2):
        lp, tp, _ = ctxs[r].gather_ptrs()
        got = _from_ptr(torch, lp, (16, 1000), torch.float32, devs[r]).cpu().numpy()
        assert np.array_equal(got, want_l)
        assert np.array_equal(outs[r][1].cpu().numpy(), want_l[8 * r:8 * r + 8])   # the local outputs are still written
    with pytest.raises(mn.Mnv1Error):
        ctxs[0].forward(synth.images(9))     # more rows than the gather block holds
    for c in ctxs:
        c.close()


def test_fused_head_matches_three_kernel_head(mn, synth_net, monkeypatch):
    """head_fused_kernel (cluster of 8 CTAs: pool -> FC -> softmax; opt-in with MNV1_FUSED_HEAD=1 until it beats the
    three launches) against the pool / fc_mma / softmax launches: same arithmetic in the same order, so the logits
    and top-1 are bit-identical."""
    import subprocess, sys, os, json
    from mnv1_b200 import synth
    img = synth.images(37)
    got_l, got_t, got_p = _single(mn, synth_net, img)
    # the switch is read once per process: run the cluster head in a child
    code = ("import sys, json, numpy as np; sys.path.insert(0, %r); import mnv1_b200; from mnv1_b200 import binding as mn, synth;"
            "w = synth.weights(); sc, sh = synth.batchnorm(); c = mn.Context(0, mn.BF16); c.set_pad_mode(1);"
            "c.set_input_transform(1/127.5, -1.0); c.set_weights(w, sc, sh, mn.ACT_RELU6);"
            "l, t, p = c.forward(synth.images(37)); np.save(sys.argv[1], l); print(json.dumps(t.tolist()))") % \
        os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = "/tmp/mnv1_head3.npy"
    env = dict(os.environ, MNV1_FUSED_HEAD="1")
    r = subprocess.run([sys.executable, "-c", code, path], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    want_l = np.load(path)
    assert np.array_equal(got_l, want_l)
    assert got_t.tolist() == json.loads(r.stdout.strip().splitlines()[-1])


def test_dp_weighted_shards_and_calibration_keep_the_results(mn, synth_net):
    """uneven shards (mnv1_dp_set_shard_weights, or mnv1_dp_calibrate = weights from each GPU's measured pinned H2D rate
    with all GPUs copying): the batch is cut differently, the outputs are bit-identical to the single-context run."""
    from mnv1_b200 import synth
    w, sc, sh = synth_net
    world, n = 4, 41
    img = synth.images(n)
    want_l, want_t, want_p = _single(mn, synth_net, img)
    dp = mn.DataParallel(_devices(world), mn.BF16, max_batch_per_gpu=24)
    dp.set_pad_mode(mn.PAD_TFSAME); dp.set_input_transform(1 / 127.5, -1.0); dp.set_weights(w, sc, sh, mn.ACT_RELU6)
    weights = [1.0, 2.5, 0.5, 2.0]
    counts = [mn.dp_shard_weighted(n, r, world, weights)[1] for r in range(world)]
    assert sum(counts) == n and counts[1] > counts[0] > counts[2]
    dp.set_shard_weights(weights)
    lg, t1, p1 = dp.forward(img)
    assert np.array_equal(lg, want_l) and np.array_equal(t1, want_t) and np.array_equal(p1, want_p)
    with pytest.raises(mn.Mnv1Error):
        dp.forward(synth.images(64))         # rank 1's share of 64 images exceeds max_batch_per_gpu = 24
    with pytest.raises(mn.Mnv1Error):
        dp.set_shard_weights([1.0, 0.0, 1.0, 1.0])
    rates = dp.calibrate()
    assert len(rates) == world and all(r > 0.5 for r in rates)      # GB/s
    lg, t1, _ = dp.forward(img)
    assert np.array_equal(lg, want_l) and np.array_equal(t1, want_t)
    dp.set_shard_weights(None)
    lg, _, _ = dp.forward(img)
    assert np.array_equal(lg, want_l)
    dp.close()


def test_gather_rows_window_for_uneven_shards(mn, synth_net):
    """mnv1_gather_set_rows: two ranks with 11 + 5 images of a 16-row block (2 x 8) — rank 0's rows land at 0..10,
    rank 1's at 11..15, and the block equals the single-context run."""
    import torch
    from mnv1_b200 import synth
    w, sc, sh = synth_net
    devs = _devices(2)
    ctxs = []
    for r in range(2):
        c = mn.Context(devs[r], mn.BF16)
        c.set_pad_mode(mn.PAD_TFSAME); c.set_input_transform(1 / 127.5, -1.0); c.set_weights(w, sc, sh, mn.ACT_RELU6)
        c.gather_create(2, r, 8)
        ctxs.append(c)
    ctxs[0].gather_attach(ctxs[1]); ctxs[1].gather_attach(ctxs[0])
    spans = [mn.dp_shard_weighted(16, r, 2, [11.0, 5.0]) for r in range(2)]
    assert spans == [(0, 11), (11, 5)]
    with pytest.raises(mn.Mnv1Error):
        ctxs[1].gather_set_rows(11, 6)       # leaves the 16-row block
    for r in range(2):
        ctxs[r].gather_set_rows(*spans[r])
    img = synth.images(16)
    want_l, want_t, _ = _single(mn, synth_net, img)
    keep = []
    for r in range(2):
        first, cnt = spans[r]
        with torch.cuda.device(devs[r]):
            d = torch.from_numpy(img[first:first + cnt].copy()).to(f"cuda:{devs[r]}")
            lg = torch.empty(cnt, 1000, device=f"cuda:{devs[r]}"); t1 = torch.empty(cnt, dtype=torch.int32, device=f"cuda:{devs[r]}")
            ctxs[r].forward_device(d.data_ptr(), cnt, lg.data_ptr(), t1.data_ptr())
            keep.append((d, lg, t1))
    for c in ctxs:
        c.sync()
    for r in range(2):
        lp, tp, _ = ctxs[r].gather_ptrs()
        assert np.array_equal(_from_ptr(torch, lp, (16, 1000), torch.float32, devs[r]).cpu().numpy(), want_l)
        assert np.array_equal(_from_ptr(torch, tp, (16,), torch.int32, devs[r]).cpu().numpy(), want_t)
    with pytest.raises(mn.Mnv1Error):
        ctxs[1].forward(synth.images(6))     # more rows than rank 1's window
    for c in ctxs:
        c.close()
