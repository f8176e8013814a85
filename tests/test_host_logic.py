"""Host-side logic that needs no GPU: the layer schedule, the C-ABI library's exports, weight
and image file readers, error behaviour without a device, the batch sharding."""
import ctypes as C
import os

import numpy as np
import pytest

import mnv1_b200  # noqa: F401
from mnv1_b200 import binding as mn, layers, shard, synth


def test_layer_schedule_matches_the_reference_counts():
    L = layers.LAYERS
    assert len(L) == 29 and layers.TOTAL_WEIGHTS == 4209088 and layers.BN_CHANNELS == 10944
    # readSquezeNetKernel counts, SURVEY App. A column nW (MobileNet.c:241,337,419,...,2696)
    nw = [864, 288, 2048, 576, 8192, 1152, 16384, 1152, 32768, 2304, 65536, 2304, 131072] + \
         [4608, 262144] * 5 + [4608, 524288, 9216, 1048576, 0, 1024000]
    assert [l.w_cnt for l in L] == nw
    assert sum(l.w_cnt for l in L[:5]) == 11968 and sum(l.w_cnt for l in L[:13]) == 264640
    assert abs(sum(2 * l.macs for l in L) / 1e6 - 1137.53) < 0.1  # MFLOP per image
    assert sum(l.in_elems + l.out_elems for l in L) == 10238952
    assert [l.hout for l in L[:5]] == [112, 112, 112, 56, 56] and L[25].stride == 1


def test_library_exports_every_declared_symbol():
    lib = mn.lib()
    syms = mn.declared_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert b"sm_100a" in lib.mnv1_version()


def test_layer_table_from_the_library_agrees_with_python():
    class Info(C.Structure):
        _fields_ = [(n, C.c_int) for n in ("index", "kind", "cin", "cout", "hin", "hout", "stride")] + \
                   [(n, C.c_long) for n in ("w_off", "w_cnt", "c_off")]
    arr = (Info * 29)()
    assert mn.lib().mnv1_layer_table(arr) == 0
    for a, b in zip(arr, layers.LAYERS):
        assert (a.index, a.kind, a.cin, a.cout, a.hin, a.hout, a.stride, a.w_off, a.w_cnt, a.c_off) == \
               (b.index, b.kind, b.cin, b.cout, b.hin, b.hout, b.stride, b.w_off, b.w_cnt, b.c_off)


def test_no_gpu_means_a_loud_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mn.Mnv1Error) as e:
        mn.Context(0, mn.BF16)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def _parse(path):
    w = np.empty(layers.TOTAL_WEIGHTS, np.float32)
    sc = np.empty(layers.TOTAL_CHANNELS, np.float32)
    sh = np.empty(layers.TOTAL_CHANNELS, np.float32)
    rc = mn.lib().mnv1_parse_weights(path.encode(), w.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p),
                                     sh.ctypes.data_as(C.c_void_p))
    return rc, w, sc, sh


def test_weight_file_round_trip_binary_and_text(tmp_path, synth_net):
    w, sc, sh = synth_net
    p = str(tmp_path / "w.bin")
    assert mn.lib().mnv1_save_weights_bin(p.encode(), w.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p),
                                          sh.ctypes.data_as(C.c_void_p)) == 0
    rc, w2, sc2, sh2 = _parse(p)
    assert rc == 0 and np.array_equal(w, w2) and np.array_equal(sc, sc2) and np.array_equal(sh, sh2)
    # the reference's text format: whitespace separated decimals (MobileNet.c:37-44), filters only
    t = str(tmp_path / "weights_c.txt")
    wi = synth.kat_ints(3, layers.TOTAL_WEIGHTS, -5, 5).astype(np.float32)
    with open(t, "w") as f:
        f.write(" ".join("%d" % v for v in wi[:1000]) + "\n")
        np.savetxt(f, wi[1000:], fmt="%.1f", newline=" ")
    rc, w3, sc3, sh3 = _parse(t)
    assert rc == 0 and np.array_equal(w3, wi) and np.all(sc3 == 1) and np.all(sh3 == 0)
    # wrong token count -> MNV1_EIO with a message, not garbage weights
    with open(t, "w") as f:
        f.write("1 2 3")
    rc, *_ = _parse(t)
    assert rc == -4 and b"tokens" in mn.lib().mnv1_last_error(None)
    rc, *_ = _parse(str(tmp_path / "missing.txt"))
    assert rc == -4


def test_ppm_reader_skips_the_header(tmp_path):
    img = synth.images(1)[0]
    p = str(tmp_path / "Cat_Image0.ppm")
    with open(p, "wb") as f:
        f.write(b"P6\n# made by the test\n224 224\n255\n" + img.tobytes())
    out = np.zeros((224, 224, 3), np.uint8)
    assert mn.lib().mnv1_read_ppm(p.encode(), out.ctypes.data_as(C.c_void_p), 224, 224) == 0
    assert np.array_equal(out, img)  # decode_image (MobileNet.c:49-57) would have read the header as pixels
    with open(p, "wb") as f:
        f.write(b"P6\n100 100\n255\n" + bytes(30000))
    assert mn.lib().mnv1_read_ppm(p.encode(), out.ctypes.data_as(C.c_void_p), 224, 224) == -4


def test_shard_ranges_cover_the_batch():
    for n, g in [(2048, 8), (2048, 4), (256, 1), (10, 4), (3, 8), (0, 2)]:
        spans = [shard.shard_range(n, r, g) for r in range(g)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    assert shard.shard_range(2048, 3, 8) == (768, 1024)
    with pytest.raises(ValueError):
        shard.shard_range(8, 8, 8)


def test_bf16_storage_weights():
    w = synth.weights()
    q = synth.bf16_storage_weights(w)
    L = layers.LAYERS
    assert np.array_equal(q[L[1].w_off:L[1].w_off + L[1].w_cnt], w[L[1].w_off:L[1].w_off + L[1].w_cnt])  # depthwise fp32
    pw = q[L[2].w_off:L[2].w_off + L[2].w_cnt]
    assert np.all((pw.view(np.uint32) & 0xFFFF) == 0)  # bf16-representable
    assert np.max(np.abs(pw - w[L[2].w_off:L[2].w_off + L[2].w_cnt]) / np.abs(w[L[2].w_off:L[2].w_off + L[2].w_cnt])) <= 2 ** -8
    st = slice(L[0].w_off, L[0].w_off + L[0].w_cnt)  # stem: fp16 (11 significant bits) after folding 1/127.5
    assert np.max(np.abs(q[st] - w[st])) <= 2.0 ** -11 * np.max(np.abs(w[st]))


# --------------------------------------------------------------------------- tools/compare_logs.py, bench rows
def _load_tool(name):
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location(name, os.path.join(root, "tools" if name != "bench" else "", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_compare_logs_parses_reference_format():
    """the reference's printf formats (MobileNet.c:315, :2792) and the per-layer / class comparison"""
    cl = _load_tool("compare_logs")
    ref = ("Kernel Execution time for Layer 1: 0.002000\nKernel Execution time for Layer 2: 0.004000\n"
           "Highest Probability of the element is present at location 283 and it's value is 0.913000.\n")
    ours = ("Kernel Execution time for Layer 1: 0.000100\nLayer 2 op: 1\t\nKernel Execution time for Layer 2: 0.000100\n"
            "Total kernel time for 2 layer(s), batch 1: 0.000200\n"
            "Highest Probability of the element is present at location 283 and it's value is 0.912700.\n")
    res = cl.compare(ref, ours)
    assert [r[0] for r in res["rows"]] == [1, 2]
    assert abs(res["rows"][0][3] - 20.0) < 1e-9 and abs(res["rows"][1][3] - 40.0) < 1e-9
    assert res["ref_final"] == (283, 0.913) and res["our_final"][0] == 283 and res["same_class"] is True
    assert cl.compare(ref, ours.replace("location 283", "location 7"))["same_class"] is False
    assert cl.compare("", "garbage")["rows"] == [] and cl.parse("nothing")[1] is None


def test_bench_launch_rows_from_graph_prefixes():
    """bench.launch_rows: one row per launch from the cumulative prefix times (-1 = no launch ends at that layer);
    a fused depthwise->pointwise launch is ONE row (dw input + pw output + both filters), the rows add up to the
    last prefix time, and kernel_rooflines groups them per distinct kernel."""
    bench = _load_tool("bench")
    from mnv1_b200.layers import LAYERS
    peaks = {"hbm_gbs": 6543.1, "bf16_tflops": 1395.8}
    cum = [0.1 * (k + 1) for k in range(29)]
    names = ["k%d" % (k % 3) for k in range(29)]
    plain = bench.launch_rows(LAYERS, cum, names, 256, peaks)
    assert len(plain) == 29 and abs(sum(r["us"] for r in plain) - cum[28] * 1e3) < 1e-6
    fused = list(cum)
    fused[1] = -1.0          # layer 2 (dw) runs fused with layer 3 (pw)
    fused[27] = -1.0         # the pool runs inside the head launch
    merged = bench.launch_rows(LAYERS, fused, names, 256, peaks)
    assert len(merged) == 27 and abs(sum(r["us"] for r in merged) - cum[28] * 1e3) < 1e-6
    row = merged[1]
    assert row["kind"] == "dw+pw" and row["layer"] == "2+3" and abs(row["us"] - 200.0) < 1e-6 and row["kernel"] == "k2"
    want = (LAYERS[1].in_elems * 2 + LAYERS[2].out_elems * 2) * 256 + LAYERS[1].w_cnt * 4 + LAYERS[2].w_cnt * 2
    assert row["bytes"] == want and row["bytes"] < plain[1]["bytes"] + plain[2]["bytes"]
    assert abs(row["flops"] - (plain[1]["flops"] + plain[2]["flops"])) < 1.0
    assert merged[-1]["layer"] == "28+29" and merged[-1]["kind"] == "pool+fc"
    ks = bench.kernel_rooflines(merged, cum[28] * 1e3, peaks)
    assert sorted(k["kernel"] for k in ks) == ["k0", "k1", "k2"]
    assert sum(k["launches_per_step"] for k in ks) == 27 and abs(sum(k["share_of_step"] for k in ks) - 1.0) < 0.01
    for k in ks:
        assert k["unit"] == ("TFLOP/s" if k["bound"] == "tensor" else "GB/s") and 0 < k["frac"]


def test_dp_shard_matches_python_sharding():
    """mnv1_dp_shard (C-ABI, no GPU needed) == shard.shard_range: contiguous, remainder on the low ranks."""
    from mnv1_b200 import shard
    for n in (0, 1, 7, 256, 2048, 2049):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                first, count = mn.dp_shard(n, r, world)
                a, b = shard.shard_range(n, r, world)
                assert (first, first + count) == (a, b)
                cover += list(range(first, first + count))
            assert cover == list(range(n))
    with pytest.raises(mn.Mnv1Error):
        mn.dp_shard(4, 2, 2)


def test_dp_shard_weighted_apportions_the_batch():
    """mnv1_dp_shard_weighted (C-ABI, no GPU needed): contiguous cover of the batch, counts within one image of the exact
    share, equal weights == mnv1_dp_shard, bad weights rejected."""
    rng = np.random.default_rng(5)
    rates = [27.2, 25.8, 25.0, 24.4, 38.2, 37.8, 37.7, 38.7]          # GB/s per GPU measured on an 8 x B200 box
    for n in (0, 1, 7, 256, 2048, 2049):
        for world in (1, 2, 3, 8):
            for w in ([1.0] * world, rates[:world], list(rng.uniform(0.1, 5.0, world))):
                cover, counts = [], []
                for r in range(world):
                    first, count = mn.dp_shard_weighted(n, r, world, w)
                    cover += list(range(first, first + count))
                    counts.append(count)
                assert cover == list(range(n))
                for r in range(world):
                    assert abs(counts[r] - n * w[r] / sum(w)) < 1.0 + 1e-6
                for r in range(world):                      # the Python mirror agrees with the C-ABI
                    a, b = shard.weighted_range(n, r, world, w)
                    assert (a, b - a) == mn.dp_shard_weighted(n, r, world, w)
            for r in range(world):
                assert mn.dp_shard_weighted(n, r, world, [3.0] * world) == mn.dp_shard(n, r, world)
                assert mn.dp_shard_weighted(n, r, world, None) == mn.dp_shard(n, r, world)
    first, count = mn.dp_shard_weighted(2048, 0, 8, rates)
    assert count < 256 < mn.dp_shard_weighted(2048, 7, 8, rates)[1]
    for bad in ([1.0, 0.0], [1.0, -2.0], [float("nan"), 1.0]):
        with pytest.raises(mn.Mnv1Error):
            mn.dp_shard_weighted(16, 0, 2, bad)
