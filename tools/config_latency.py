#!/usr/bin/env python
"""BASELINE configs 1-3 alone: batch-1 latency of the 5-, 13- and 29-layer cuts (fp32 and bf16), what bench.py prints as
`configs` (without the integer-mode block)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mnv1_b200  # noqa: E402,F401
from mnv1_b200 import binding as mn, synth  # noqa: E402

bench.integer_mode_throughput = lambda *a, **k: None
r = bench.config_latencies(mn, synth, 0)
for name in ("fp32", "bf16"):
    print(name, r[name]["cuts"])
    print("  ", [(x["upto_layer"], x["ms"]) for x in r[name]["per_launch_ms"]])
