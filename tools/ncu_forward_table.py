#!/usr/bin/env python
"""One forward pass captured with `ncu --set full` -> the per-launch table and profiles/dram_traffic.json.

usage: ncu_forward_table.py <report.ncu-rep> <table.txt> [dram_traffic.json]
The report is expected to hold exactly one forward pass (25 launches with layers 2-11 fused), e.g.
  ncu --set full --clock-control none -s 104 -c 25 -o fwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline
(4 image-synthesis launches + 4 forward passes skipped).  CPU only."""
import csv
import json
import re
import subprocess
import sys


def base_name(name: str) -> str:
    """the kernel's bare name, as mnv1_last_kernel_name / bench.py's rows call it"""
    for k in ("stem_rows_kernel", "stem_tc_kernel", "fused_pair_kernel", "depthwise_tma_kernel", "depthwise_ring_kernel",
              "depthwise_cw_kernel", "pointwise_pair_kernel", "pointwise_tc_kernel", "head_fused_kernel"):
        if name.startswith(k): return "head" if k == "head_fused_kernel" else k
    if name.startswith("fused_rb_kernel"): return "fused_dw_pw_kernel"
    if name.startswith(("pool_kernel", "fc_mma_kernel", "softmax_kernel", "fc_kernel")): return "head"
    return name.split("<")[0].strip()


def family(name: str) -> str:
    if "fused_rb" in name: return "dw+pw"
    if "stem" in name: return "stem"
    if "depthwise" in name: return "dw"
    if "pointwise" in name: return "pw"
    if "pool" in name: return "pool"
    return "fc"                     # fc_mma_kernel + softmax_kernel: the bench's "fc" row is both


def main(rep, table, traffic=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}

    def num(r, key):
        v = r[col[key]].replace(",", "")
        return float(v) if v else 0.0

    def to_bytes(r, key):
        unit = rows[1][col[key]].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
        return num(r, key) * scale

    out = ["# one forward pass (batch 256, bf16) under `ncu --set full --clock-control none`",
           "# per launch: duration (cold-cache, serialised), DRAM bytes read / written, tensor-pipe active %, issue-slot %",
           f"{'kernel':78s} {'us':>6s} {'rd_MB':>8s} {'wr_MB':>8s} {'tensor%':>8s} {'issue%':>7s} {'regs':>5s}"]
    fam, kern = {}, {}
    for r in rows[2:]:
        name = re.sub(r"^void |mnv1::|<unnamed>::|unnamed>::|\(anonymous namespace\)::", "", r[col["Kernel Name"]]).split("(")[0]
        us = num(r, "gpu__time_duration.sum")
        if rows[1][col["gpu__time_duration.sum"]].startswith("ns"): us /= 1e3
        if rows[1][col["gpu__time_duration.sum"]].startswith("ms"): us *= 1e3
        rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
        tens = num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in col else 0.0
        issue = num(r, "sm__inst_executed.sum.pct_of_peak_sustained_elapsed") if "sm__inst_executed.sum.pct_of_peak_sustained_elapsed" in col else 0.0
        regs = int(num(r, "launch__registers_per_thread"))
        out.append(f"{name[:78]:78s} {us:6.1f} {rd / 1e6:8.1f} {wr / 1e6:8.1f} {tens:8.1f} {issue:7.1f} {regs:5d}")
        f = fam.setdefault(family(name), {"launches": 0, "bytes": 0.0, "us": 0.0})
        f["launches"] += 1; f["bytes"] += rd + wr; f["us"] += us
        k = kern.setdefault(base_name(name), {"launches": 0, "bytes": 0.0, "us": 0.0})
        k["launches"] += 1; k["bytes"] += rd + wr; k["us"] += us
    open(table, "w").write("\n".join(out) + "\n")
    if traffic:
        # the bench's rows: "fc" = fc + softmax launches of one pass, every other family per launch
        js = {"source": f"{table} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, averaged over the family's launches of one forward pass)"}
        detail = {}
        for k, f in fam.items():
            per = f["bytes"] / (1 if k == "fc" else f["launches"])
            js[k] = int(per)
            detail[k] = {"launches": f["launches"], "dram_bytes_per_launch": int(per), "ncu_us_per_launch": round(f["us"] / (1 if k == "fc" else f["launches"]), 1)}
        js["detail"] = detail
        # per distinct kernel (bench.py's roofline.kernels[].traffic): DRAM bytes per launch; "head" = its launches together
        js["kernels"] = {k: int(v["bytes"] / (1 if k == "head" else v["launches"])) for k, v in kern.items()}
        json.dump(js, open(traffic, "w"), indent=1)
    print("\n".join(out[:6]))


if __name__ == "__main__":
    main(*sys.argv[1:4])
