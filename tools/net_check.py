#!/usr/bin/env python
"""Full-network forward at batch N with fused blocks on and off; prints per-layer times.
usage: net_check.py N [fused=1]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mnv1_b200  # noqa
from mnv1_b200 import binding as mn, synth

n = int(sys.argv[1]); fused = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = mn.Context(0, mn.BF16)
ctx.set_pad_mode(mn.PAD_TFSAME)
ctx.set_input_transform(1 / 127.5, -1.0)
w = synth.weights(); sc, sh = synth.batchnorm()
ctx.set_weights(w, sc, sh, mn.ACT_RELU6)
ctx.use_fused_blocks(bool(fused))
img = synth.images(min(n, 8))
img = np.ascontiguousarray(np.concatenate([img] * ((n + len(img) - 1) // len(img)))[:n])
print("forward...", flush=True)
t0 = time.time()
lg, t1, p1 = ctx.forward(img)
print(f"forward ok {time.time() - t0:.3f}s top1[:8]={t1[:8].tolist()}", flush=True)
d_img = ctx.upload_u8(img)
mn.lib().mnv1_buf_device_ptr.restype = __import__("ctypes").c_void_p
lt = ctx.profile_layers(mn.lib().mnv1_buf_device_ptr(d_img.h), n, iters=5)
print("layer us:", [round(float(x) * 1e3, 1) for x in lt], "sum", round(float(sum(lt)) * 1e3, 1))
ctx.close()
