"""Keras-h5 exporter: the HDF5 reader against a libhdf5-written file, the exporter
against a synthetic Keras-layout file, and the exported filter orders against an independent NHWC/HWIO
convolution through the oracle.  CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

import h5_writer

HERE = os.path.dirname(os.path.abspath(__file__))
PLAN = [(32, 64), (64, 128), (128, 128), (128, 256), (256, 256), (256, 512)] + [(512, 512)] * 5 + [(512, 1024), (1024, 1024)]


@pytest.fixture(scope="module")
def kh5():
    import mnv1_b200  # noqa: F401
    from mnv1_b200 import keras_h5
    return keras_h5


def test_reader_on_a_file_written_by_libhdf5(kh5):
    """tests/golden/libhdf5_sample.mat = scipy's testhdf5_7.4_GLNX86.mat: a MATLAB v7.3 file, i.e. HDF5 written
    by libhdf5 behind a 512-byte user block, holding `testdouble` = 0, pi/4, ..., 2 pi (scipy's own expectation)."""
    ds = kh5.H5File(os.path.join(HERE, "golden", "libhdf5_sample.mat")).datasets()
    assert list(ds) == ["testdouble"]
    assert ds["testdouble"].dtype == np.float64 and ds["testdouble"].shape == (9, 1)
    assert np.array_equal(ds["testdouble"].ravel(), np.arange(0, 2 * np.pi + 0.01, np.pi / 4))


def _keras_tree(rng):
    """A model.save_weights()-shaped tree: <layer>/<layer>/<var>:0, plus weightless layers as empty groups."""
    def bn(c):
        return {"gamma:0": (0.5 + rng.random(c)).astype(np.float32), "beta:0": rng.standard_normal(c).astype(np.float32) * 0.1,
                "moving_mean:0": rng.standard_normal(c).astype(np.float32) * 0.1,
                "moving_variance:0": (0.5 + rng.random(c)).astype(np.float32)}
    tree = {"input_1": {}, "conv1_pad": {}, "conv1_relu": {}, "dropout": {}, "reshape_2": {}}
    tree["conv1"] = {"conv1": {"kernel:0": rng.standard_normal((3, 3, 3, 32)).astype(np.float32)}}
    tree["conv1_bn"] = {"conv1_bn": bn(32)}
    for i, (cin, cout) in enumerate(PLAN, start=1):
        tree[f"conv_dw_{i}"] = {f"conv_dw_{i}": {"depthwise_kernel:0": rng.standard_normal((3, 3, cin, 1)).astype(np.float32)}}
        tree[f"conv_dw_{i}_bn"] = {f"conv_dw_{i}_bn": bn(cin)}
        tree[f"conv_dw_{i}_relu"] = {}
        tree[f"conv_pw_{i}"] = {f"conv_pw_{i}": {"kernel:0": (rng.standard_normal((1, 1, cin, cout)) * 0.05).astype(np.float32)}}
        tree[f"conv_pw_{i}_bn"] = {f"conv_pw_{i}_bn": bn(cout)}
        tree[f"conv_pw_{i}_relu"] = {}
    tree["conv_preds"] = {"conv_preds": {"kernel:0": (rng.standard_normal((1, 1, 1024, 1000)) * 0.03).astype(np.float32),
                                         "bias:0": rng.standard_normal(1000).astype(np.float32) * 0.01}}
    return tree


@pytest.fixture(scope="module")
def keras_file(tmp_path_factory):
    tree = _keras_tree(np.random.default_rng(7))
    path = str(tmp_path_factory.mktemp("h5") / "mobilenet_1_0_224_tf.h5")
    h5_writer.write_h5(path, tree)
    return path, tree


def _flatten(tree, prefix=""):
    for k, v in tree.items():
        p = f"{prefix}/{k}" if prefix else k
        if isinstance(v, dict):
            yield from _flatten(v, p)
        else:
            yield p, v


@pytest.mark.parametrize("user_block", [0, 512])
def test_reader_round_trip_of_a_keras_shaped_file(kh5, tmp_path, user_block):
    """84 groups at the root = 11 symbol-table nodes under the B-tree, nested groups, empty groups,
    object headers with continuation blocks, 4.25 M floats."""
    tree = _keras_tree(np.random.default_rng(8))
    path = str(tmp_path / "w.h5")
    h5_writer.write_h5(path, tree, user_block=user_block)
    ds = kh5.H5File(path).datasets()
    want = dict(_flatten(tree))
    assert sorted(ds) == sorted(want)
    assert sum(v.size for v in ds.values()) == 4209088 + 1000 + 4 * 10944
    for k, v in want.items():
        assert ds[k].dtype == np.float32 and np.array_equal(ds[k], v), k


def test_export_matches_a_direct_fold_and_loads_through_the_c_abi(kh5, keras_file, tmp_path):
    import mnv1_b200  # noqa: F401
    from mnv1_b200 import binding as mn
    path, tree = keras_file
    out = str(tmp_path / "weights.bin")
    kh5.export(path, out)
    NW, NC = 4209088, 10944 + 1000
    w, sc, sh = np.empty(NW, np.float32), np.empty(NC, np.float32), np.empty(NC, np.float32)
    assert mn.lib().mnv1_parse_weights(out.encode(), w.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p),
                                       sh.ctypes.data_as(C.c_void_p)) == 0
    # the reference's per-layer token counts (MobileNet.c App. A) and orders, rebuilt here independently
    from mnv1_b200 import layers
    table = layers.LAYERS           # mirrors mnv1_layer_table (tests/test_host_logic.py checks they agree)
    names = ["conv1"] + [f"conv_{k}_{i}" for i in range(1, 14) for k in ("dw", "pw")]
    for info, name in zip(table, names):
        leaf = tree[name][name]
        kern = leaf.get("kernel:0", leaf.get("depthwise_kernel:0"))
        if name == "conv1":
            flat = kern.transpose(3, 2, 0, 1)                    # [O][I][H][W]
        elif "_dw_" in name:
            flat = kern[:, :, :, 0].transpose(2, 0, 1)           # [C][3][3]
        else:
            flat = kern[0, 0].T                                   # [O][I]
        assert np.array_equal(w[info.w_off:info.w_off + info.w_cnt], flat.reshape(-1)), name
        b = tree[name + "_bn"][name + "_bn"]
        s = b["gamma:0"].astype(np.float64) / np.sqrt(b["moving_variance:0"].astype(np.float64) + 1e-3)
        t = b["beta:0"].astype(np.float64) - b["moving_mean:0"].astype(np.float64) * s
        c = s.size
        assert np.array_equal(sc[info.c_off:info.c_off + c], s.astype(np.float32)), name
        assert np.array_equal(sh[info.c_off:info.c_off + c], t.astype(np.float32)), name
    fc = tree["conv_preds"]["conv_preds"]
    assert np.array_equal(w[-1024000:], fc["kernel:0"][0, 0].T.reshape(-1))
    assert np.array_equal(sh[-1000:], fc["bias:0"]) and np.all(sc[-1000:] == 1.0)


def test_exported_orders_mean_what_keras_means(kh5, keras_file, oracle_mod):
    """Layers 1-3 computed by the oracle from the EXPORTED arrays equal a direct NHWC / HWIO evaluation of
    the Keras definition (ZeroPadding bottom/right + valid = TF-SAME, BN with eps 1e-3, ReLU6) on a small map."""
    path, tree = keras_file
    w, sc, sh = kh5.mobilenet_from_datasets(kh5.H5File(path).datasets())
    rng = np.random.default_rng(9)
    n, hw = 2, 12
    img = rng.integers(0, 256, (n, hw, hw, 3), dtype=np.uint8)
    x = img.astype(np.float64) / 127.5 - 1.0

    def bn_relu6(y, name):
        b = tree[name][name]
        y = (y - b["moving_mean:0"]) / np.sqrt(b["moving_variance:0"].astype(np.float64) + 1e-3) * b["gamma:0"] + b["beta:0"]
        return np.clip(y, 0.0, 6.0)

    def conv(xin, k_hwio, stride, depthwise=False):
        nn, h, ww, cin = xin.shape
        ho = h // stride
        pad = np.zeros((nn, h + 2, ww + 2, cin))
        lo = 0 if stride == 2 else 1                              # stride 2: pad bottom/right only
        pad[:, lo:lo + h, lo:lo + ww] = xin
        cout = cin if depthwise else k_hwio.shape[3]
        y = np.zeros((nn, ho, ho, cout))
        for i in range(3):
            for j in range(3):
                patch = pad[:, i:i + stride * ho:stride, j:j + stride * ho:stride]
                y += patch * k_hwio[i, j, :, 0] if depthwise else patch @ k_hwio[i, j]
        return y

    k1 = tree["conv1"]["conv1"]["kernel:0"].astype(np.float64)
    a1 = bn_relu6(conv(x, k1, 2), "conv1_bn")
    a2 = bn_relu6(conv(a1, tree["conv_dw_1"]["conv_dw_1"]["depthwise_kernel:0"].astype(np.float64), 1, depthwise=True), "conv_dw_1_bn")
    a3 = bn_relu6(a2 @ tree["conv_pw_1"]["conv_pw_1"]["kernel:0"][0, 0].astype(np.float64), "conv_pw_1_bn")

    o = oracle_mod
    flat = img.reshape(-1)
    g1 = o.convolute(img, flat[1:], flat[2:], w[:864].reshape(32, 3, 3, 3), n, hw, hw, 2, 32, pad_mode=o.PAD_TFSAME,
                     in_scale=1 / 127.5, in_bias=-1.0, scale=sc[:32], shift=sh[:32], act=o.ACT_RELU6, pix_stride=3,
                     img_stride=hw * hw * 3)
    g2 = o.depthwise(g1, w[864:864 + 288].reshape(32, 3, 3), 1, pad_mode=o.PAD_TFSAME, scale=sc[32:64], shift=sh[32:64],
                     act=o.ACT_RELU6)
    g3 = o.pointwise(g2, w[1152:1152 + 2048].reshape(64, 32), 64, scale=sc[64:128], shift=sh[64:128], act=o.ACT_RELU6)
    for got, want in ((g1, a1), (g2, a2), (g3, a3)):
        assert np.max(np.abs(got.transpose(0, 2, 3, 1) - want)) < 2e-4


def test_errors_are_loud(kh5, tmp_path, keras_file):
    p = tmp_path / "junk.h5"
    p.write_bytes(b"not an hdf5 file" * 100)
    with pytest.raises(kh5.H5Error):
        kh5.H5File(str(p))
    path, _tree = keras_file
    ds = kh5.H5File(path).datasets()
    del ds["conv_pw_7_bn/conv_pw_7_bn/beta:0"]
    with pytest.raises(kh5.H5Error, match="conv_pw_7_bn/beta:0"):
        kh5.mobilenet_from_datasets(ds)
    ds = kh5.H5File(path).datasets()
    ds["conv1/conv1/kernel:0"] = ds["conv1/conv1/kernel:0"].transpose(3, 2, 0, 1).copy()
    with pytest.raises(kh5.H5Error, match="kernel of shape"):
        kh5.mobilenet_from_datasets(ds)


@pytest.mark.gpu
def test_exported_file_runs_the_network_like_the_arrays_it_came_from(kh5, keras_file, oracle_mod, tmp_path):
    """h5 -> export -> mnv1_load_weights -> forward: the same logits (bit for bit) as mnv1_set_weights with the
    folded arrays, the oracle's logits within the fp32 bar, and the oracle's top-1."""
    import mnv1_b200  # noqa: F401
    from mnv1_b200 import binding as mn, synth
    path, _tree = keras_file
    out = str(tmp_path / "weights.bin")
    kh5.export(path, out)
    w, sc, sh = kh5.mobilenet_from_datasets(kh5.H5File(path).datasets())
    imgs = synth.images(2)
    res = []
    for load in (True, False):
        c = mn.Context(0, mn.F32)
        c.set_pad_mode(mn.PAD_TFSAME)
        c.set_input_transform(1 / 127.5, -1.0)
        if load:
            c.load_weights(out, mn.ACT_RELU6)
        else:
            c.set_weights(w, sc, sh, mn.ACT_RELU6)
        res.append(c.forward(imgs))
        c.close()
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    want = oracle_mod.forward(imgs, w, sc, sh, pad_mode=oracle_mod.PAD_TFSAME, act=oracle_mod.ACT_RELU6)
    want_logits = want[0] if isinstance(want, tuple) else want
    assert np.max(np.abs(res[0][0] - want_logits) / np.maximum(1.0, np.abs(want_logits))) <= 1e-3
    srt = np.sort(want_logits, axis=1)
    clear = (srt[:, -1] - srt[:, -2]) > 1e-3 * np.maximum(1.0, np.abs(srt[:, -1]))      # top-1 is only defined up to the bar
    assert np.array_equal(res[0][1][clear], np.argmax(want_logits, axis=1)[clear])
