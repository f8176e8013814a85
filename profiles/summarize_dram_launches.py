#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list per kernel.
usage: summarize_dram_launches.py <launches.csv> > summary.txt"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
    per = collections.OrderedDict()
    to_us = {"ns": 1e-3, "us": 1.0, "ms": 1e3}
    to_mb = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    for r in rows:
        d = per.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0].replace("unnamed>::", "").replace("void ", "")})
        v, u, m = float(r["Metric Value"].replace(",", "")), r["Metric Unit"], r["Metric Name"]
        if m.startswith("gpu__time"):
            d["us"] = v * to_us.get(u, 1e-3)
        elif "read" in m:
            d["rd"] = v * to_mb[u]
        elif "write" in m:
            d["wr"] = v * to_mb[u]
    agg = collections.OrderedDict()
    for d in per.values():
        a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += d.get("us", 0); a[2] += d.get("rd", 0); a[3] += d.get("wr", 0)
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {len(per)} launches, {tot:.1f} us total (ncu: cold-cache, serialised)")
    print("%-52s %8s %9s %9s %12s %12s %6s" % ("kernel", "launches", "total_us", "avg_us", "rd_MB/launch", "wr_MB/launch", "share"))
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-52s %8d %9.1f %9.2f %12.2f %12.2f %6.3f" % (n[-52:], a[0], a[1], a[1] / a[0], a[2] / a[0], a[3] / a[0], a[1] / tot))


if __name__ == "__main__":
    main(sys.argv[1])
