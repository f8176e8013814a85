#!/usr/bin/env python
"""Print the handful of ncu metrics the roofline discussion uses from a .ncu-rep (CPU only)."""
import csv
import subprocess
import sys
import collections

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.sum", "sm__inst_executed.sum.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subunit_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print("==", d.get("Kernel Name", "")[:110])
        for k in KEYS:
            if k in d:
                print(f"  {k:75s} {d[k]:>14s} {units[hdr.index(k)]}")
        stalls = {h.split("issue_stalled_")[1].split("_per")[0]: float(d[h]) for h in hdr
                  if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and d[h]}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:6]
        print("  stalls/issue:", ", ".join(f"{k}={v:.2f}" for k, v in top))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    if len(rows) > 2:
        hdr = rows[1]
        S, N, I = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
        h, hs = collections.Counter(), collections.Counter()
        for r in rows[2:]:
            if len(r) <= I or not r[I].isdigit():
                continue
            toks = r[S].split()
            op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
            h[op] += int(r[I]); hs[op] += int(r[N])
        tot, tots = sum(h.values()), max(1, sum(hs.values()))
        print(f"  warp instructions {tot}; opcode mix (share of executed / share of stall samples):")
        print("   ", ", ".join(f"{op} {c / tot:.3f}/{hs[op] / tots:.3f}" for op, c in h.most_common(14)))


if __name__ == "__main__":
    for r in sys.argv[1:]:
        main(r)
