#!/usr/bin/env python
"""Is the per-site cost of a 1 M-site evaluation (4-5 % above that of a 125 k-site shard) a matter of the code-row
stride (1000 rows x 1 MB apart: one page per row and tile) or of sustained power?  Same total work both ways:
one alignment of S sites vs a batch of 8 trees (the same tree) on S / 8 sites each, one launch each."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import mcphylo_jl_b200 as mcp
    from mcphylo_jl_b200 import capi
    from mcphylo_jl_b200.phylodist import _tree_args

    S, parts = 1000000, 8
    w = bench.make_workload("cfg4", S)
    codes, leaf_nums = bench.make_codes(w, 0, S)
    ctx = capi.Context(0)
    ft, targs = _tree_args(mcp.PhyloDist(w["tree"], w["pi"], w["srates"], w["rates"], w["model"]))
    whole = ctx.alignment_from_codes(codes, w["K"], leaf_nums)
    step = S // parts
    pieces = [ctx.alignment_from_codes(np.ascontiguousarray(codes[:, i * step:(i + 1) * step]), w["K"], leaf_nums) for i in range(parts)]
    out = []
    for rnd in range(2):
        ms = []
        for _ in range(5):
            ll, g = ctx.eval(whole, *targs, want_grad=True)
            ms.append(ctx.stats()["walk_ms"])
        out.append({"round": rnd, "layout": "one alignment, 1 M sites", "walk_ms_median": float(np.median(ms[1:])), "ll": ll})
        ms = []
        for _ in range(5):
            lls, gs = ctx.eval_batch(pieces, [targs] * parts, want_grad=True)
            ms.append(ctx.stats()["walk_ms"])
        out.append({"round": rnd, "layout": "8 alignments of 125 k sites, one batched launch", "walk_ms_median": float(np.median(ms[1:])),
                    "ll": float(np.sum(lls)), "grid": ctx.stats()["grid"], "cpt": ctx.stats()["columns_per_thread"]})
        print(out[-2], out[-1], file=sys.stderr, flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
