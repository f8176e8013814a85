#!/usr/bin/env python
"""Time one depthwise->pointwise block through mnv1_dw_pw_block (the fused kernels) and through the two separate
kernels:  python tools/run_block.py --c 512 --cout 512 --h 14 --stride 1 --n 256   (MNV1_LIB selects the library)"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mnv1_b200  # noqa: E402,F401
from mnv1_b200 import binding as mn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c", type=int, default=512); ap.add_argument("--cout", type=int, default=512)
    ap.add_argument("--h", type=int, default=14); ap.add_argument("--stride", type=int, default=1)
    ap.add_argument("--n", type=int, default=256); ap.add_argument("--iters", type=int, default=7)
    a = ap.parse_args()
    ctx = mn.Context(0, mn.BF16)
    ctx.set_pad_mode(mn.PAD_TFSAME)
    rng = np.random.default_rng(0)
    fd = ctx.filter(mn.DEPTHWISE, rng.standard_normal((a.c, 9)).astype(np.float32), a.c, a.c,
                    (0.5 + rng.random(a.c)).astype(np.float32), (rng.standard_normal(a.c) * 0.1).astype(np.float32), mn.ACT_RELU6)
    fp = ctx.filter(mn.POINTWISE, (rng.standard_normal((a.cout, a.c)) * 0.05).astype(np.float32), a.c, a.cout,
                    (0.5 + rng.random(a.cout)).astype(np.float32), (rng.standard_normal(a.cout) * 0.1).astype(np.float32), mn.ACT_RELU6)
    ho = a.h // a.stride
    x = ctx.malloc(a.n, a.c, a.h, a.h); mid = ctx.malloc(a.n, a.c, ho, ho); out = ctx.malloc(a.n, a.cout, ho, ho)

    def med(fn):
        ts = []
        for _ in range(a.iters + 2):
            fn(); ts.append(ctx.last_kernel_ms())
        return float(np.median(ts[2:])) * 1e3
    try:
        t = med(lambda: ctx.dw_pw_block(out, x, fd, fp, a.h, a.h, a.stride))
        print(f"{os.path.basename(mn.LIB_PATH)} fused {ctx.last_kernel_name}: {t:.1f} us")
    except mn.Mnv1Error as e:
        print("no fused variant:", e)
    t1 = med(lambda: ctx.depthwise(mid, x, fd, a.h, a.h, 3, a.stride, a.c)); k1 = ctx.last_kernel_name
    t2 = med(lambda: ctx.pointwise(out, mid, fp, ho, ho, a.c, a.cout)); k2 = ctx.last_kernel_name
    print(f"separate: {k1} {t1:.1f} us + {k2} {t2:.1f} us = {t1 + t2:.1f} us")
    ctx.close()


if __name__ == "__main__":
    main()
