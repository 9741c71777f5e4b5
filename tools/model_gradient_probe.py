#!/usr/bin/env python
"""Cost of a model-gradient evaluation (mcp_eval_model_gradient, runtime-K kernel + moment matrices) next to the plain
gradient evaluation (tuned kernels) on BASELINE-shaped inputs: device time of the walk kernel from the library's CUDA
events, wall time of the whole call, and the agreement of logL / branch gradient between the two paths.

    python tools/model_gradient_probe.py --cases cfg3:100000,cfg2:10000,cfg4:20000 > profiles/r2e_model_gradient.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="cfg3:100000,cfg2:10000")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import mcphylo_jl_b200 as mcp
    from mcphylo_jl_b200 import capi
    from mcphylo_jl_b200.phylodist import _tree_args
    from mcphylo_jl_b200.substitution_models import model_derivatives

    out = {"rows": []}
    ctx = capi.Context(0)
    for case in args.cases.split(","):
        name, S = case.split(":")
        S = int(S)
        w = bench.make_workload(name, S)
        codes, leaf_nums = bench.make_codes(w, 0, S)
        aln = ctx.alignment_from_codes(codes, w["K"], leaf_nums)
        ft, targs = _tree_args(mcp.PhyloDist(w["tree"], w["pi"], w["srates"], w["rates"], w["model"]))
        _, dA, dpi = model_derivatives(w["model"], w["pi"], w["srates"])
        plain_ms, plain_wall, mg_ms, mg_wall = [], [], [], []
        for _ in range(args.reps + 1):
            t0 = time.perf_counter()
            ll0, g0 = ctx.eval(aln, *targs, want_grad=True)
            plain_wall.append((time.perf_counter() - t0) * 1e3)
            plain_ms.append(ctx.stats()["walk_ms"])
            t0 = time.perf_counter()
            ll1, g1, pg = ctx.eval_model_gradient(aln, *targs, dA=dA, dpi=dpi)
            mg_wall.append((time.perf_counter() - t0) * 1e3)
            st = ctx.stats()
            mg_ms.append(st["walk_ms"])
        row = {"case": case, "K": w["K"], "R": len(w["rates"]), "n_par": int(dA.shape[2]),
               "plain_walk_ms": float(np.median(plain_ms[1:])), "plain_call_ms": float(np.median(plain_wall[1:])),
               "model_gradient_walk_ms": float(np.median(mg_ms[1:])), "model_gradient_call_ms": float(np.median(mg_wall[1:])),
               "grid": st["grid"], "block": st["block"],
               "ll_rel_diff": abs(ll1 - ll0) / abs(ll0),
               "grad_max_rel_diff": float(np.max(np.abs(g1 - g0)) / np.max(np.abs(g0))),
               "par_grad": [float(v) for v in pg]}
        out["rows"].append(row)
        print(row, file=sys.stderr, flush=True)
        aln.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
