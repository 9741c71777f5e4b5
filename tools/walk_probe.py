#!/usr/bin/env python
"""Experiment helper (not part of the product or the tests): times the walk kernel of one
workload for several launch shapes through the C ABI and prints CUDA-event kernel times."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import mcphylo_jl_b200 as mcp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg4")
ap.add_argument("--sites", type=int, default=200000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--shapes", default="0:0")
ap.add_argument("--nograd", action="store_true")
a = ap.parse_args()
w = bench.make_workload(a.workload, a.sites)
codes, leaf_nums = bench.make_codes(w, 0, w["S"])
aln = mcp.DeviceAlignment(codes, leaf_nums, w["K"])
pd = mcp.PhyloDist(w["tree"], w["pi"], w["srates"], w["rates"], w["model"])
ctx = mcp.get_context(0)
balg = bench.algorithmic_bytes(w["n_taxa"], w["S"], w["K"], w["R"], not a.nograd)
for shape in a.shapes.split(","):
    parts = [int(v) for v in shape.split(":")]
    blk, cps = parts[0], parts[1]
    ctx.set_launch(blk, cps)
    ctx.set_columns_per_thread(parts[2] if len(parts) > 2 else 0)
    ts = []
    for r in range(a.reps + 1):
        res = mcp.logpdf(pd, aln) if a.nograd else mcp.gradlogpdf(pd, aln)[0]
        st = ctx.stats()
        if r:
            ts.append(st["walk_ms"])
    t = float(np.median(ts))
    print(f"shape={shape:10s} block={st['block']:4d} grid={st['grid']:5d} tiles={st['tiles']:6d} walk_ms={t:9.3f} "
          f"alg_GB/s={balg / t / 1e6:8.1f} device_ms={st['device_ms']:.3f} ll={res:.6f}", flush=True)
