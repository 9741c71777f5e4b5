#!/usr/bin/env python
"""Condenses `ncu -i X.ncu-rep --page raw --csv` output into the small JSON summaries kept under
profiles/ (the .ncu-rep files themselves stay in gpurun_out/).  Usage:
    ncu -i gpurun_out/walk.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_summary.py raw.csv [--traffic workload sites_per_gpu algorithmic_bytes] > summary.json"""
import csv
import json
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "smsp__warps_eligible.avg.per_cycle_active",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        rec = dict(zip(hdr, vals))
        s = {"kernel": rec.get("Kernel Name")}
        for h, u in zip(hdr, units):
            if h in KEEP or ("issue_stalled" in h and h.endswith("_per_issue_active.ratio")):
                try:
                    v = float(rec[h])
                except ValueError:
                    continue
                if "issue_stalled" in h:
                    if v >= 0.05:
                        s.setdefault("stall_warps_per_issue", {})[h.split("issue_stalled_")[1].split("_per_issue")[0]] = round(v, 3)
                else:
                    s[h] = [v, u]
        out.append(s)
    json.dump(out if len(out) > 1 else out[0], sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
