#!/usr/bin/env python
"""A few device-resident forward passes of the full network (batch 256, bf16) and nothing else: the short command
ncu wraps.  One image-synthesis launch, then `--iters` forwards of 25 launches each (stem, 5 fused dw+pw, L12, L13,
5 x (dw, pw), L24..L27, pool, fc, softmax):
    ncu --set full --clock-control none -s 26 -c 25 -o fwd python tools/run_forward.py --iters 3"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import mnv1_b200  # noqa: E402,F401
from mnv1_b200 import binding as mn, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--dtype", default="bf16")
    a = ap.parse_args()
    ctx = mn.Context(0, {"bf16": mn.BF16, "f32": mn.F32, "u8": mn.U8}[a.dtype])
    ctx.set_pad_mode(mn.PAD_TFSAME)
    if a.dtype == "u8":      # the integer contexts: seeded s8 filters, a per-layer requantisation shift (bench.py's configs.u8)
        import numpy as np
        from mnv1_b200.layers import LAYERS, TOTAL_WEIGHTS, TOTAL_CHANNELS, DEPTHWISE, STEM, POOL
        w = synth.kat_ints(99, TOTAL_WEIGHTS, -127, 127).astype(np.float32)
        sc = np.ones(TOTAL_CHANNELS, np.float32)
        for L in LAYERS:
            if L.kind != POOL:
                fan = 27 if L.kind == STEM else 9 if L.kind == DEPTHWISE else L.cin
                sc[L.c_off:L.c_off + L.cout] = 2.0 ** -int(np.ceil(np.log2(np.sqrt(fan) * 74 * 1.2)))
        ctx.set_weights(w, sc, np.zeros(TOTAL_CHANNELS, np.float32), mn.ACT_RELU)
    else:
        ctx.set_input_transform(1 / 127.5, -1.0)
        ctx.set_weights(synth.weights(), *synth.batchnorm(), mn.ACT_RELU6)
    ctx.plan(a.n)
    img = torch.empty(a.n * 224 * 224 * 3, dtype=torch.uint8, device="cuda")
    lg = torch.empty(a.n, 1000, device="cuda"); t1 = torch.empty(a.n, dtype=torch.int32, device="cuda"); p1 = torch.empty(a.n, device="cuda")
    ctx.synth_images_device(img.data_ptr(), a.n, 0, synth.IMAGE_SEED)
    for _ in range(a.iters):
        ctx.forward_device(img.data_ptr(), a.n, lg.data_ptr(), t1.data_ptr(), p1.data_ptr())
    ctx.sync()
    print("top1[:8]", t1[:8].tolist(), "launches", ctx.launch_count)
    ctx.close()


if __name__ == "__main__":
    main()
