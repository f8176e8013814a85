#!/usr/bin/env python
"""Run one layer of the schedule through the C-ABI a few times and print its device time.
Used for kernel iteration and as the short command ncu wraps:
    python tools/run_layer.py --layer 2 --n 256 --iters 5
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mnv1_b200  # noqa: E402
from mnv1_b200 import binding as mn  # noqa: E402
from mnv1_b200.layers import LAYERS, STEM, DEPTHWISE, POINTWISE  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layer", type=int, required=True)
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--dtype", default="bf16")
    a = ap.parse_args()
    L = LAYERS[a.layer - 1]
    ctx = mn.Context(0, mn.BF16 if a.dtype == "bf16" else mn.F32)
    ctx.set_pad_mode(mn.PAD_TFSAME)
    rng = np.random.default_rng(0)
    sc = (0.5 + rng.random(L.cout)).astype(np.float32)
    sh = (rng.standard_normal(L.cout) * 0.1).astype(np.float32)
    esz = 2 if a.dtype == "bf16" else 4
    if L.kind == STEM:
        img = ctx.upload_u8(rng.integers(0, 256, (a.n, 224, 224, 3), dtype=np.uint8))
        f = ctx.filter(mn.CONVOLUTE, rng.standard_normal((32, 27)).astype(np.float32), 3, 32, sc, sh, mn.ACT_RELU6)
        out = ctx.malloc(a.n, 32, 112, 112)
        run = lambda: ctx.convolute_rgb(out, img, f, 224, 224, 3, 2, 32)
        nbytes = a.n * (224 * 224 * 3 + L.out_elems * esz)
    elif L.kind == DEPTHWISE:
        x = ctx.malloc(a.n, L.cin, L.hin, L.hin)
        f = ctx.filter(mn.DEPTHWISE, rng.standard_normal((L.cin, 9)).astype(np.float32), L.cin, L.cin, sc, sh, mn.ACT_RELU6)
        out = ctx.malloc(a.n, L.cout, L.hout, L.hout)
        run = lambda: ctx.depthwise(out, x, f, L.hin, L.hin, 3, L.stride, L.cin)
        nbytes = a.n * (L.in_elems + L.out_elems) * esz
    elif L.kind == POINTWISE:
        x = ctx.malloc(a.n, L.cin, L.hin, L.hin)
        f = ctx.filter(mn.POINTWISE, (rng.standard_normal((L.cout, L.cin)) * 0.05).astype(np.float32), L.cin, L.cout, sc, sh, mn.ACT_RELU6)
        out = ctx.malloc(a.n, L.cout, L.hout, L.hout)
        run = lambda: ctx.pointwise(out, x, f, L.hin, L.hin, L.cin, L.cout)
        nbytes = a.n * (L.in_elems + L.out_elems) * esz + L.w_cnt * esz
    else:
        raise SystemExit("only stem / depthwise / pointwise layers")
    ts = []
    for _ in range(a.iters + 2):
        run()
        ts.append(ctx.last_kernel_ms())
    t = float(np.median(ts[2:]))
    print(f"layer {a.layer} {ctx.last_kernel_name} n={a.n}: {t * 1e3:.1f} us, {nbytes / t / 1e6:.0f} GB/s, "
          f"{2 * L.macs * a.n / t / 1e9:.1f} TFLOP/s")
    ctx.close()


if __name__ == "__main__":
    main()
