#!/usr/bin/env python
"""The integer mode alone: batch-256 throughput and the in-graph per-launch table (what bench.py prints as configs.u8)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mnv1_b200  # noqa: E402,F401
from mnv1_b200 import binding as mn, synth  # noqa: E402

r = bench.integer_mode_throughput(mn, synth, 0)
print(r["images_per_s"], r["ms_per_step"])
agg = {}
for row in r["per_launch"]:
    print(row["layer"], row["kernel"], row["us"], row["frac_of_hbm_roofline"])
    agg[row["kernel"]] = agg.get(row["kernel"], 0.0) + row["us"]
print(json.dumps({k: round(v, 1) for k, v in agg.items()}))
