#!/usr/bin/env python
"""Run one fused depthwise->pointwise block at a given batch and compare it with the two separate
kernels (bit-identical arithmetic).  usage: fused_stress.py C COUT H STRIDE N [pad]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mnv1_b200  # noqa
from mnv1_b200 import binding as mn

c, cout, h, stride, n = [int(a) for a in sys.argv[1:6]]
pad = int(sys.argv[6]) if len(sys.argv) > 6 else 1
ctx = mn.Context(0, mn.BF16)
ctx.set_pad_mode(pad)
rng = np.random.default_rng(1)
ho = h // stride
x = (rng.random((n, c, h, h), dtype=np.float32) * 6)
wd = rng.standard_normal((c, 3, 3)).astype(np.float32) * 0.5
sd, td = (0.5 + rng.random(c)).astype(np.float32), (rng.standard_normal(c) * 0.1).astype(np.float32)
wp = (rng.standard_normal((cout, c)) * np.sqrt(2.0 / c)).astype(np.float32)
sp, tp = (0.5 + rng.random(cout)).astype(np.float32), (rng.standard_normal(cout) * 0.1).astype(np.float32)
fd = ctx.filter(mn.DEPTHWISE, wd, c, c, sd, td, mn.ACT_RELU6)
fp = ctx.filter(mn.POINTWISE, wp, c, cout, sp, tp, mn.ACT_RELU6)
xin = ctx.upload_planar(x)
out = ctx.malloc(n, cout, ho, ho)
t0 = time.time()
ctx.dw_pw_block(out, xin, fd, fp, h, h, stride)
ctx.sync()
print(f"fused {ctx.last_kernel_name} ok in {time.time() - t0:.3f}s, kernel {ctx.last_kernel_ms():.3f} ms", flush=True)
got = ctx.download_planar(out)
m = ctx.malloc(n, c, ho, ho)
ctx.depthwise(m, xin, fd, h, h, 3, stride, c)
t_dw = ctx.last_kernel_ms()
out2 = ctx.malloc(n, cout, ho, ho)
ctx.pointwise(out2, m, fp, ho, ho, c, cout)
t_pw = ctx.last_kernel_ms()
got2 = ctx.download_planar(out2)
print(f"unfused dw {t_dw:.3f} + pw {t_pw:.3f} ms; identical fraction {np.mean(got == got2):.6f}", flush=True)
reps = int(os.environ.get("REPS", "3"))
for _ in range(reps):
    ctx.dw_pw_block(out, xin, fd, fp, h, h, stride)
ctx.sync()
print(f"fused warm {ctx.last_kernel_ms():.3f} ms after {reps} reps")
got3 = ctx.download_planar(out)
print(f"still identical {np.mean(got3 == got2):.6f}")
ctx.close()
