#!/usr/bin/env python
"""Walk-kernel time of cfg4-like shards for every launch shape (tile width x columns per thread),
to calibrate the planner's cost model (csrc/planner.hpp: choose_walk_shape).

    python tools/shape_sweep.py --sites 15625,31250,62500,125000,250000 > profiles/r2_shape_sweep.json
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
    ap.add_argument("--sites", default="31250,62500,125000,250000")
    ap.add_argument("--reps", type=int, default=3)
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
        ctx.set_level_mode(0)
        shapes = [(0, 0)] + [(b, c) for c in (2, 1) for b in (256, 224, 192, 160, 128, 96, 64)]
        for block, cpt in shapes:
            ctx.set_launch(block, 0)
            ctx.set_columns_per_thread(cpt)
            ms = []
            for _ in range(args.reps + 1):
                ctx.eval(aln, *targs, want_grad=True)
                ms.append(ctx.stats()["walk_ms"])
            st = ctx.stats()
            out["rows"].append({"sites": S, "block_req": block, "cpt_req": cpt, "block": st["block"],
                                "cpt": st["columns_per_thread"], "grid": st["grid"], "tiles": st["tiles"],
                                "walk_ms": float(np.min(ms[1:]))})
            print(out["rows"][-1], file=sys.stderr)
        aln.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
