"""The oracle and the synthetic-data generators against the committed golden fixtures
(tests/golden/mnv1_golden.npz, produced by tests/golden/make_golden.py)."""
import os
import zlib

import numpy as np
import pytest

import mnv1_b200  # noqa: F401
from mnv1_b200 import synth

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mnv1_golden.npz"))


def test_synthetic_data_is_pinned(synth_net):
    w, sc, sh = synth_net
    img = synth.images(4)
    assert [zlib.crc32(img[i].tobytes()) for i in range(4)] == G["image_crc"].tolist()
    assert np.array_equal(synth.images(2, first=2), img[2:])  # image n does not depend on the batch
    assert np.array_equal(w[::65768][:64], G["weight_probe"])
    assert np.array_equal(sc[::187][:64], G["scale_probe"]) and np.array_equal(sh[::187][:64], G["shift_probe"])


def test_oracle_full_network_fp32(oracle_mod, synth_net):
    w, sc, sh = synth_net
    logits, taps = oracle_mod.forward(synth.images(4), w, sc, sh, taps=(2, 3, 5, 13))
    np.testing.assert_allclose(logits, G["logits_f32"], rtol=1e-5, atol=1e-6)
    _, top1, _ = oracle_mod.softmax_argmax(logits)
    assert np.array_equal(top1, G["top1_f32"])
    for k in (2, 3, 5, 13):
        np.testing.assert_allclose(taps[k][:, ::7, ::5, ::3], G[f"l{k:02d}_sample"], rtol=1e-5, atol=1e-6)


def test_oracle_bf16_storage_emulation(oracle_mod, synth_net):
    w, sc, sh = synth_net
    logits, _ = oracle_mod.forward(synth.images(2), synth.bf16_storage_weights(w), sc, sh, rbf16=1)
    np.testing.assert_allclose(logits, G["logits_bf16"][:2], rtol=1e-5, atol=1e-6)
    # and it stays close to the fp32 network (the stated end-to-end bf16 tolerance is 0.05)
    assert np.max(np.abs(logits - G["logits_f32"][:2])) < 0.05


@pytest.mark.gpu
def test_cuda_against_golden(synth_net):
    """The CUDA path against the committed vectors (no oracle involved at run time)."""
    from mnv1_b200 import binding as mn
    w, sc, sh = synth_net
    img = synth.images(4)
    for dtype, key, tol in ((mn.F32, "logits_f32", 1e-4), (mn.BF16, "logits_bf16", 0.05)):
        c = mn.Context(0, dtype)
        c.set_pad_mode(mn.PAD_TFSAME)
        c.set_input_transform(1 / 127.5, -1.0)
        c.set_weights(w, sc, sh, mn.ACT_RELU6)
        logits, top1, _ = c.forward(img)
        ref = G[key]
        assert np.max(np.abs(logits - ref) / np.maximum(1, np.abs(ref))) <= tol
        if dtype == mn.F32:
            assert np.array_equal(top1, G["top1_f32"])
            for k in (2, 3, 5, 13):
                got = c.forward_upto(img, k)[:, ::7, ::5, ::3]
                assert np.max(np.abs(got - G[f"l{k:02d}_sample"]) / np.maximum(1, np.abs(G[f"l{k:02d}_sample"]))) <= 1e-4
        c.close()
