#!/usr/bin/env python
"""External reference point for the tensor-bound 1x1 layers — NEVER on the product path.

Times torch.matmul (cuBLAS / cuBLASLt under the hood) on the pointwise GEMMs of the MobileNet.c schedule at batch 256,
bf16 in, fp32 accumulate, bf16 out, L2-cold operands rotated over 4 buffers, CUDA events, median of 30:
    python experiments/gemm_yardstick.py
and prints it beside the tensor / HBM roofline of the same shape, so that "pointwise_pair_kernel reaches 0.68 of the
sustained tensor peak on layer 15" can be read against what the vendor library does on the very same box.  The library GEMM
has no BatchNorm / ReLU6 epilogue and no NHWC constraints, so it is an upper bound for a fused kernel, not a competitor."""
import json
import os

import torch

SHAPES = [("L13", 50176, 256, 512), ("L15-23", 50176, 512, 512), ("L25", 12544, 512, 1024), ("L27", 12544, 1024, 1024),
          ("L11", 200704, 256, 256)]


def main():
    peaks = {"hbm_gbs": 6543.1, "bf16_tflops": 1395.8}
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        peaks = {"hbm_gbs": j["hbm_gbs"], "bf16_tflops": j.get("bf16_tflops_sustained", j["bf16_tflops"])}
    dev = "cuda"
    for name, m, k, n in SHAPES:
        a = [torch.randn(m, k, device=dev, dtype=torch.bfloat16) for _ in range(4)]
        w = torch.randn(n, k, device=dev, dtype=torch.bfloat16)
        out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
        for i in range(5):
            torch.matmul(a[i % 4], w.t(), out=out)
        ts = []
        for i in range(30):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a[i % 4], w.t(), out=out); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        us = ts[len(ts) // 2]
        flops = 2.0 * m * k * n
        nbytes = 2.0 * (m * k + m * n + n * k)
        roof = max(flops / (peaks["bf16_tflops"] * 1e12), nbytes / (peaks["hbm_gbs"] * 1e9)) * 1e6
        print(f"{name:7s} M={m:6d} K={k:4d} N={n:4d}: cuBLAS {us:6.1f} us = {flops / us / 1e6:7.1f} TFLOP/s, "
              f"roofline {roof:5.1f} us, library / roofline {roof / us:.2f}")


if __name__ == "__main__":
    main()
