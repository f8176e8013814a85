"""The C host programs (host/MobileNet*.c on the C-ABI) end to end on a GPU: same printed lines as
the reference (MobileNet.c:315, :2792), same top-1 as the Python path."""
import os
import re
import subprocess

import numpy as np
import pytest

import mnv1_b200  # noqa: F401
from mnv1_b200 import binding as mn, synth

HOST = os.path.join(mn.PKG_DIR, "host")


def _inputs(tmp_path):
    import ctypes as C
    img = synth.images(1)[0]
    (tmp_path / "Cat_Image0.ppm").write_bytes(b"P6\n224 224\n255\n" + img.tobytes())
    w = synth.weights()
    sc, sh = synth.batchnorm()
    assert mn.lib().mnv1_save_weights_bin(str(tmp_path / "weights_c.txt").encode(), w.ctypes.data_as(C.c_void_p),
                                          sc.ctypes.data_as(C.c_void_p), sh.ctypes.data_as(C.c_void_p)) == 0


def test_host_programs_build():
    subprocess.run(["make", "-C", HOST], check=True, capture_output=True)
    for exe in ("out", "out_13layers", "out_l5"):
        assert os.access(os.path.join(HOST, exe), os.X_OK)


@pytest.mark.gpu
@pytest.mark.parametrize("exe,layers,flags", [("out", 29, []), ("out", 29, ["-bf16"]), ("out_13layers", 13, []),
                                              ("out_l5", 5, ["-dump"])])
def test_host_program_runs(tmp_path, exe, layers, flags, synth_net):
    subprocess.run(["make", "-C", HOST], check=True, capture_output=True)
    _inputs(tmp_path)
    r = subprocess.run([os.path.join(HOST, exe)] + flags, cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    times = re.findall(r"Kernel Execution time for Layer (\d+): ([0-9.]+)", r.stdout)
    assert [int(k) for k, _ in times] == list(range(1, layers + 1))
    if layers == 29:
        m = re.search(r"Highest Probability of the element is present at location (\d+) and it's value is ([0-9.]+)\.", r.stdout)
        assert m, r.stdout
        w, sc, sh = synth_net
        c = mn.Context(0, mn.BF16 if "-bf16" in flags else mn.F32)
        c.set_pad_mode(mn.PAD_TFSAME)
        c.set_input_transform(1 / 127.5, -1.0)
        c.set_weights(w, sc, sh, mn.ACT_RELU6)
        _, top1, p1 = c.forward(synth.images(1))
        c.close()
        assert int(m.group(1)) == int(top1[0]) + 1           # the reference prints a 1-based location
        assert abs(float(m.group(2)) - float(p1[0])) < 1e-3
    if "-dump" in flags:
        assert "Layer 5 op:" in r.stdout


def test_host_program_fails_loudly_without_inputs(tmp_path):
    subprocess.run(["make", "-C", HOST], check=True, capture_output=True)
    r = subprocess.run([os.path.join(HOST, "out_l5")], cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "Error:" in r.stdout


@pytest.mark.gpu
def test_host_program_integer_mode_matches_kernel_cl_arithmetic(tmp_path, oracle_mod):
    """`out -u8 -wrap -ref-pad` with a weight file in the reference's own TEXT format (whitespace-separated decimal
    integers, readSquezeNetKernel MobileNet.c:31-47): the 29-layer schedule in kernel.cl's arithmetic — u8 maps, int
    filters, `if (sum <= 0) sum = 0`, the store to unsigned char wrapping modulo 256.  The printed top-1 must be
    the oracle's (integer mode, wrapping store) for the same image and weights."""
    subprocess.run(["make", "-C", HOST], check=True, capture_output=True)
    img = synth.images(1)
    (tmp_path / "Cat_Image0.ppm").write_bytes(b"P6\n224 224\n255\n" + img[0].tobytes())
    w = synth.kat_ints(5, 4209088, -3, 3).astype(np.int32)
    with open(tmp_path / "weights_c.txt", "w") as f:           # the reference's format: one token per value
        f.write("\n".join(str(int(v)) for v in w))
    r = subprocess.run([os.path.join(HOST, "out"), "-u8", "-wrap", "-ref-pad"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"Highest Probability of the element is present at location (\d+) and it's value is ([0-9.]+)\.", r.stdout)
    assert m, r.stdout
    logits, _ = oracle_mod.forward(img, w.astype(np.float32), None, None, pad_mode=oracle_mod.PAD_REF, act=oracle_mod.ACT_RELU,
                                   rbf16=oracle_mod.STORE_U8_WRAP, in_scale=1.0, in_bias=0.0)
    _, otop1, op1 = oracle_mod.softmax_argmax(logits)
    assert int(m.group(1)) == int(otop1[0]) + 1
    assert abs(float(m.group(2)) - float(op1[0])) < 1e-3
