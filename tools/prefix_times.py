#!/usr/bin/env python
"""In-graph per-launch times of one batch-256 bf16 forward (mnv1_profile_prefixes), the numbers bench.py's `layers` rows are
made of; quick A/B of a kernel inside the graph: MNV1_LIB=<variant .so> python tools/prefix_times.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import mnv1_b200  # noqa: E402,F401
from mnv1_b200 import binding as mn, synth  # noqa: E402

n = 256
ctx = mn.Context(0, mn.BF16)
ctx.set_pad_mode(mn.PAD_TFSAME); ctx.set_input_transform(1 / 127.5, -1.0)
ctx.set_weights(synth.weights(), *synth.batchnorm(), mn.ACT_RELU6)
ctx.plan(n)
img = torch.empty(n * 224 * 224 * 3, dtype=torch.uint8, device="cuda")
ctx.synth_images_device(img.data_ptr(), n, 0, synth.IMAGE_SEED)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    cum = ctx.profile_prefixes(img.data_ptr(), n, iters=21)
    prev, rows = 0.0, []
    for k in range(1, 30):
        if cum[k - 1] >= 0:
            rows.append((k, round((cum[k - 1] - prev) * 1e3, 1)))
            prev = cum[k - 1]
    print("step_us", round(cum[28] * 1e3, 1), "rows", rows)
ctx.close()
