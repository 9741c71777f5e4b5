#!/usr/bin/env python
"""A/B of walk-kernel variants on cfg4-like inputs (walk_ms from the library's CUDA events):
cherries stored vs recomputed in the gradient pass, static tile ranges vs atomic tickets.

    python tools/ab_walk.py --sites 1000000,125000 > profiles/r2_ab_cherry_tickets.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg4")
    ap.add_argument("--sites", default="1000000,125000")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import mcphylo_jl_b200 as mcp
    from mcphylo_jl_b200 import capi
    from mcphylo_jl_b200.phylodist import _tree_args

    out = {"workload": args.workload, "rows": []}
    ctx = capi.Context(0)
    for S in [int(s) for s in args.sites.split(",")]:
        w = bench.make_workload(args.workload, S)
        codes, leaf_nums = bench.make_codes(w, 0, S)
        aln = ctx.alignment_from_codes(codes, w["K"], leaf_nums)
        ft, targs = _tree_args(mcp.PhyloDist(w["tree"], w["pi"], w["srates"], w["rates"], w["model"]))
        ref = None
        for rnd in range(2):                      # two interleaved rounds: the board drifts with temperature / power
            for cherry, tickets in ((0, 0), (1, 0), (0, 1), (1, 1)):
                ctx.set_cherry_mode(cherry)
                ctx.set_tile_order(tickets)
                ms, res = [], None
                for _ in range(args.reps + 1):
                    res = ctx.eval(aln, *targs, want_grad=True)
                    ms.append(ctx.stats()["walk_ms"])
                if ref is None:
                    ref = res
                row = {"sites": S, "round": rnd, "cherries_recomputed": cherry, "atomic_tickets": tickets,
                       "walk_ms_median": float(np.median(ms[1:])), "walk_ms_min": float(np.min(ms[1:])),
                       "ll_equal_to_first": bool(res[0] == ref[0]),
                       "grad_max_rel_diff_to_first": float(np.max(np.abs(res[1] - ref[1]) / np.maximum(np.abs(ref[1]), 1e-300)))}
                out["rows"].append(row)
                print(row, file=sys.stderr)
        aln.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
