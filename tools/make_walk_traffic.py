#!/usr/bin/env python
"""profiles/walk_traffic.json from the raw ncu pages of the walk launch at several shard sizes:

    python tools/make_walk_traffic.py cfg4 "<how the captures were made>" 1000000:gpurun_out/x_1000000_raw.csv 500000:... > profiles/walk_traffic.json

bench.py looks its `roofline.traffic` up here by (workload, sites per GPU)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def metric(rec, name):
    return float(rec[name])


def main():
    workload, how = sys.argv[1], sys.argv[2]
    entries = []
    for arg in sys.argv[3:]:
        sites, path = arg.split(":")
        sites = int(sites)
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        rec = dict(zip(hdr, rows[2]))
        unit = dict(zip(hdr, units))

        def byts(name):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "B": 1.0, "KB": 1e3, "MB": 1e6, "GB": 1e9, "TB": 1e12}[unit[name]]
            return metric(rec, name) * scale

        w = bench.make_workload(workload, sites)
        stalls = {}
        for h in hdr:
            if "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                try:
                    v = float(rec[h])
                except ValueError:
                    continue
                if v >= 0.05:
                    stalls[h.split("issue_stalled_")[1].split("_per_issue")[0]] = round(v, 3)
        rd, wr = byts("dram__bytes_read.sum"), byts("dram__bytes_write.sum")
        entries.append({
            "workload": workload, "sites_per_gpu": sites, "kernel": rec.get("Kernel Name"),
            "dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
            "algorithmic_bytes_per_launch": bench.algorithmic_bytes(w["n_taxa"], sites, w["K"], w["R"], True),
            "walk_bytes_per_launch": bench.walk_bytes(w["tree"], sites, w["K"], w["R"], True),
            "gpu_time_duration_ms_under_ncu": metric(rec, "gpu__time_duration.sum") * {"msecond": 1.0, "usecond": 1e-3, "second": 1e3, "nsecond": 1e-6, "ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[unit["gpu__time_duration.sum"]],
            "issue_active_pct": metric(rec, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "fp64_pipe_active_pct": metric(rec, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "dram_throughput_pct": metric(rec, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "warps_active_pct": metric(rec, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "warp_instructions": metric(rec, "smsp__inst_executed.sum"),
            "registers_per_thread": metric(rec, "launch__registers_per_thread"),
            "stall_warps_per_issue": stalls,
            "source": how + f" (--sites {sites}); raw page {os.path.basename(path)}",
        })
    json.dump({"entries": entries}, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
