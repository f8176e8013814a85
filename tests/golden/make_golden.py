#!/usr/bin/env python
"""Generates tests/golden/mnv1_golden.npz from the CPU oracle (the reference ships no golden
vectors — SURVEY §4/§8c — so the build commits its own, with this script).

Contents (all fp32 unless noted), seeded synthetic network and images:
  logits_f32   [4][1000]   full network, fp32 storage, TF-SAME padding, ReLU6
  top1_f32     [4] int32
  logits_bf16  [4][1000]   same with bf16 storage emulation (bf16 pw/FC weights, fp16 stem taps,
                           outputs rounded to bf16 per layer)
  l02_sample / l03_sample / l05_sample / l13_sample   strided samples of layer outputs (fp32 run)
  image_crc    [4] uint32  CRC32 of each synthetic image (pins synth.images)
  weight_probe [64]        weights[::65768][:64] (pins synth.weights), scale_probe/shift_probe
Run:  python tests/golden/make_golden.py
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mnv1_b200  # noqa: E402
from mnv1_b200 import synth  # noqa: E402
import oracle  # noqa: E402


def main():
    w = synth.weights()
    sc, sh = synth.batchnorm()
    img = synth.images(4)
    logits, taps = oracle.forward(img, w, sc, sh, taps=(2, 3, 5, 13))
    _, top1, _ = oracle.softmax_argmax(logits)
    logits_b, _ = oracle.forward(img, synth.bf16_storage_weights(w), sc, sh, rbf16=1)
    out = {
        "logits_f32": logits, "top1_f32": top1.astype(np.int32), "logits_bf16": logits_b,
        "image_crc": np.array([zlib.crc32(img[i].tobytes()) for i in range(4)], dtype=np.uint32),
        "weight_probe": w[::65768][:64].copy(), "scale_probe": sc[::187][:64].copy(), "shift_probe": sh[::187][:64].copy(),
    }
    for k in (2, 3, 5, 13):
        out[f"l{k:02d}_sample"] = taps[k][:, ::7, ::5, ::3].copy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mnv1_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
