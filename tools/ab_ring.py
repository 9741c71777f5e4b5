#!/usr/bin/env python
"""A/B of the gradient pass's operand ring (mcp_set_ring_mode) on BASELINE-shaped inputs: walk_ms from the
library's CUDA events, two interleaved rounds per input (the boards drift with temperature and power), results
compared bit for bit with the ring off.

    python tools/ab_ring.py --cases cfg4:1000000,cfg4:125000,cfg3:100000 > profiles/r2_ab_ring.json
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
    ap.add_argument("--cases", default="cfg4:1000000,cfg4:125000")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpt", type=int, default=0, help="force columns per thread (0 = automatic)")
    ap.add_argument("--block", type=int, default=0, help="force the tile width in threads (0 = automatic)")
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    import mcphylo_jl_b200 as mcp
    from mcphylo_jl_b200 import capi
    from mcphylo_jl_b200.phylodist import _tree_args

    out = {"tag": args.tag, "rows": []}
    ctx = capi.Context(0)
    if args.cpt:
        ctx.set_columns_per_thread(args.cpt)
    if args.block:
        ctx.set_launch(args.block, 0)
    for case in args.cases.split(","):
        name, S = case.split(":")
        S = int(S)
        m = re.match(r"^t(\d+)k(\d+)r(\d+)$", name)     # ad-hoc shape: t<taxa>k<states>r<rate categories>
        if m and name not in bench.WORKLOADS:
            bench.WORKLOADS[name] = (int(m.group(1)), S, int(m.group(2)), int(m.group(3)), 31000 + int(m.group(1)), 32000 + int(m.group(1)))
        w = bench.make_workload(name, S)
        codes, leaf_nums = bench.make_codes(w, 0, S)
        aln = ctx.alignment_from_codes(codes, w["K"], leaf_nums)
        ft, targs = _tree_args(mcp.PhyloDist(w["tree"], w["pi"], w["srates"], w["rates"], w["model"]))
        ref = None
        for rnd in range(2):
            for ring in (0, 1):
                ctx.set_ring_mode(ring)
                ms, res, st = [], None, None
                for _ in range(args.reps + 1):
                    res = ctx.eval(aln, *targs, want_grad=True)
                    st = ctx.stats()
                    ms.append(st["walk_ms"])
                if ref is None:
                    ref = res
                row = {"case": case, "round": rnd, "ring_mode": ring, "operand_ring": st["operand_ring"],
                       "columns_per_thread": st["columns_per_thread"], "grid": st["grid"], "block": st["block"],
                       "walk_ms_median": float(np.median(ms[1:])), "walk_ms_min": float(np.min(ms[1:])),
                       "ll": res[0], "ll_equal_to_first": bool(res[0] == ref[0]),
                       "grad_equal_to_first": bool(np.array_equal(res[1], ref[1])),
                       "grad_max_rel_diff_to_first": float(np.max(np.abs(res[1] - ref[1]) / np.maximum(np.abs(ref[1]), 1e-300)))}
                out["rows"].append(row)
                print(row, file=sys.stderr, flush=True)
        aln.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
