#!/usr/bin/env python
"""Where does the end-to-end time go?  (measurement harness, not product code)
Times the pipelined host-buffer path of bench.py piece by piece on the cfg4 workload:
host-side tree flattening, pure upload, block-split evaluation without upload, and the full thing
for several block counts."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import mcphylo_jl_b200 as mcp  # noqa: E402
from mcphylo_jl_b200.phylodist import _tree_args  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg4")
ap.add_argument("--sites", type=int, default=0)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--blocks", default="1,3,4,5,6")
a = ap.parse_args()

w = bench.make_workload(a.workload, a.sites)
codes, leaf_nums = bench.make_codes(w, 0, w["S"])
d = mcp.PhyloDist(w["tree"], w["pi"], w["srates"], w["rates"], w["model"])
out = {"sites": w["S"]}

t0 = time.perf_counter()
for _ in range(20):
    _tree_args(d)
out["host_tree_args_ms"] = (time.perf_counter() - t0) / 20 * 1e3


def timed(fn, reps):
    fn()
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for nb in [int(x) for x in a.blocks.split(",")]:
    pipe = mcp.PipelinedEvaluator(codes, leaf_nums, w["K"], 0, n_blocks=nb)
    pipe.evaluate(d, True)

    def upload_only():
        for aln, host in pipe.blocks:
            aln.update_codes(host.data_ptr())
        pipe.ctx.synchronize()

    out[f"b{nb}"] = {
        "bounds": [hi - lo for lo, hi in pipe.bounds],
        "no_upload_ms": timed(lambda: pipe.evaluate(d, True, upload=False), a.reps),
        "upload_only_ms": timed(upload_only, a.reps),
        "e2e_ms": timed(lambda: pipe.evaluate(d, True, upload=True), a.reps),
    }
    pipe.close()
    del pipe
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
