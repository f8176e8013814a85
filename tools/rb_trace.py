#!/usr/bin/env python
"""Print the event stamps fused_rb.cu records for CTA 0 under MNV1_RB_TRACE=<file> (debug aid)."""
import sys
import numpy as np
t = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(8, 128, 4).astype(np.int64)
t0 = t[t > 0].min()
rel = np.where(t > 0, t - t0, -1)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
print("producer (unit: first chunk issued, last chunk issued)")
print([tuple(rel[0, i, :2]) for i in range(n)])
print("mma (unit: wait a_full begin, a_full seen, committed)")
print([tuple(rel[1, i, :3]) for i in range(n)])
print("epilogue (tile: wait tm_full begin, seen, done)")
print([tuple(rel[2, i, :3]) for i in range(n)])
print("epilogue detail (tile: before ld wait, after ld wait, after math+sts, after fence)")
print([tuple(rel[6, i, :4]) for i in range(n)])
for g in range(3):
    print(f"stencil group {g} (k-th unit: begin, in_full0 seen, a_empty seen, a_full arrived)")
    print([tuple(rel[3 + g, i, :4]) for i in range(n // 3 + 1)])
