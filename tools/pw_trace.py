#!/usr/bin/env python
"""Print the event stamps pointwise_tc.cu records for CTA 0 under MNV1_PW_TRACE=<file> (debug aid)."""
import sys
import numpy as np
t = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(4, 128, 4).astype(np.int64)
t0 = t[t > 0].min()
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
for i in range(n):
    p, m, e = t[0, i] - t0, t[1, i] - t0, t[2, i] - t0
    print(i, "producer: start", p[0], "last k-block issued +", p[2] - p[0], "| mma: start", m[0], "tmem_empty wait", m[1] - m[0],
          "first full +", m[2] - m[1], "last commit +", m[3] - m[2], "| epilogue: start", e[0], "tmem_full wait", e[1] - e[0], "done +", e[2] - e[1])
