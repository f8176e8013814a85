#!/usr/bin/env python3
"""Selected `ncu --set full` metrics per launch of a forward pass, as a text table.

usage: ncu -i forward.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_metrics_table.py raw.csv > profiles/rNN_forward_metrics.txt
Columns: us | DRAM MB read | DRAM MB written | issue slots % | tensor pipe % | FMA pipe % | shared-memory pipe % |
L2 hit % | registers | warps active % | the four largest stall reasons (warps stalled per issued instruction)."""
import csv
import sys

COLS = [("gpu__time_duration.sum", 1e-3), ("dram__bytes_read.sum", 1e-6), ("dram__bytes_write.sum", 1e-6),
        ("sm__inst_executed.avg.pct_of_peak_sustained_elapsed", 1), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", 1),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", 1), ("lts__t_sector_hit_rate.pct", 1),
        ("launch__registers_per_thread", 1), ("sm__warps_active.avg.pct_of_peak_sustained_active", 1)]
UNIT = {"ns": 1.0, "us": 1e3, "ms": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    print("# ncu --set full of one forward pass (batch 256, bf16): selected metrics per launch (tools/ncu_metrics_table.py)")
    print("# us | DRAM MB read | DRAM MB written | issue % | tensor pipe % | FMA pipe % | shared-memory pipe % | L2 hit % | regs | warps active % | stalls per issue")
    for r in data:
        name = r[ix["Kernel Name"]]
        name = name.split("(")[0].replace("void ", "").replace("mnv1::<unnamed>::", "").replace("<unnamed>::", "")
        vals = []
        for c, scale in COLS:
            v = float(r[ix[c]].replace(",", "")) * UNIT.get(units[ix[c]], 1.0) * scale
            vals.append(v)
        st = sorted(((float(r[ix[s]].replace(",", "") or 0), s[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")])
                     for s in stalls), reverse=True)[:4]
        print(f"{name[-58:]:58s} " + " ".join(f"{v:9.2f}" for v in vals) + "  " + ", ".join(f"{n}={v:.2f}" for v, n in st))


if __name__ == "__main__":
    main(sys.argv[1])
