#!/usr/bin/env python
"""Device time per gradient call of the small-tree (level-parallel) kernel on MCMC-sized inputs, for A/B runs of two
library builds on one box:  MCPHYLO_B200_LIB=build_exp/<tag>/libmcphylo_b200.so python tools/ab_small_tree.py

    python tools/ab_small_tree.py --cases cfg2,cfg1,t100k4r4:2000 --steps 200
"""
import argparse
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="cfg2,cfg1")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--no-timing", action="store_true", help="mcp_set_timing(0): no per-evaluation timing events (kernel_us reads 0)")
    ap.add_argument("--tag", default=os.environ.get("MCPHYLO_B200_LIB", "lib"))
    args = ap.parse_args()
    import mcphylo_jl_b200 as mcp
    from mcphylo_jl_b200 import capi
    from mcphylo_jl_b200.phylodist import _tree_args

    ctx = capi.Context(0)
    if args.no_timing:
        ctx.set_timing(False)
    flush = bench.L2Flusher(0)
    rows = []
    for case in args.cases.split(","):
        name, _, S = case.partition(":")
        m = re.match(r"^t(\d+)k(\d+)r(\d+)$", name)
        if m and name not in bench.WORKLOADS:
            bench.WORKLOADS[name] = (int(m.group(1)), int(S), int(m.group(2)), int(m.group(3)), 31000 + int(m.group(1)), 32000 + int(m.group(1)))
        w = bench.make_workload(name, int(S) if S else None)
        codes, leaf_nums = bench.make_codes(w, 0, w["S"])
        aln = ctx.alignment_from_codes(codes, w["K"], leaf_nums)
        ft, targs = _tree_args(mcp.PhyloDist(w["tree"], w["pi"], w["srates"], w["rates"], w["model"]))
        prep = capi.PreparedBatch(ctx, [aln], [targs], want_grad=True)
        blv0 = ft.blv.copy()

        def step(i):
            prep.set_blv(0, blv0 * (1.0 + 1e-3 * (i % 7)))
            return prep.eval()

        cold, walk = bench.timed_calls(ctx, step, args.steps, 20, flush)
        warm, walk_w = bench.timed_calls(ctx, step, args.steps, 0, None)
        st = ctx.stats()
        res = step(0)
        row = {"tag": args.tag, "case": case, "grid": st["grid"], "block": st["block"], "launches": st["kernel_launches"],
               "us_per_call_l2_cold": 1e3 * float(np.median(cold)), "us_per_call_l2_warm": 1e3 * float(np.median(warm)),
               "kernel_us_cold": 1e3 * float(np.median(walk)), "kernel_us_warm": 1e3 * float(np.median(walk_w)),
               "ll": float(res[0][0]), "grad_sum": float(np.sum(res[1][0]))}
        rows.append(row)
        print(row, file=sys.stderr, flush=True)
        aln.close()
    print(json.dumps(rows))


if __name__ == "__main__":
    main()
