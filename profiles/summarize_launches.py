#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("mnv1::", "").replace("<unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {len(rows) - 1} launches, {tot / 1e3:.1f} us total (ncu: cold-cache, serialised)")
    print(f"{'kernel':58s} {'launches':>8s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:58s} {n:8d} {t / 1e3:10.1f} {t / 1e3 / n:9.2f} {t / tot:6.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
