#!/usr/bin/env python
"""K = 20 (protein-sized alphabet): the FP64 tensor-core walk against the runtime-K fallback kernel on
100 taxa x 100 000 sites (walk_ms from the library's CUDA events), plus parity of the two against each other.

    python tools/ab_large_alphabet.py > profiles/r2_ab_large_alphabet.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--K", type=int, default=20)
    ap.add_argument("--taxa", type=int, default=100)
    ap.add_argument("--sites", type=int, default=100000)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import mcphylo_jl_b200 as mcp
    from mcphylo_jl_b200 import capi
    from mcphylo_jl_b200.phylodist import _tree_args

    K = args.K
    rng = np.random.default_rng(2020)
    tree = mcp.random_tree(args.taxa, rng)
    pi = rng.dirichlet(np.ones(K) * 5)
    srates = rng.uniform(0.2, 3.0, size=K * (K - 1) // 2)
    pool, leaf_nums = mcp.simulate_codes(tree, mcp.GTR(pi, srates), pi, np.ones(1), 4096, rng, gap_frac=0.01)
    codes = np.take(pool, rng.integers(0, 4096, size=args.sites), axis=1)
    ctx = capi.Context(0)
    aln = ctx.alignment_from_codes(codes, K, leaf_nums)
    ft, targs = _tree_args(mcp.PhyloDist(tree, pi, srates, [1.0], mcp.GTR))
    out = {"K": K, "taxa": args.taxa, "sites": args.sites, "rows": []}
    res = {}
    for mode, name in ((1, "tensor_core_walk"), (0, "runtime_K_fallback")):
        ctx.set_large_alphabet_mode(mode)
        for want_grad in (True, False):
            ms = []
            for _ in range(args.reps + 1):
                r = ctx.eval(aln, *targs, want_grad=want_grad)
                ms.append(ctx.stats()["walk_ms"])
            st = ctx.stats()
            res[(mode, want_grad)] = r
            out["rows"].append({"kernel": name, "want_grad": want_grad, "walk_ms": float(np.min(ms[1:])),
                                "grid": st["grid"], "block": st["block"], "tiles": st["tiles"]})
            print(out["rows"][-1], file=sys.stderr)
    a, b = res[(1, True)], res[(0, True)]
    out["parity_between_kernels"] = {"ll_rel": float(abs(a[0] - b[0]) / abs(b[0])),
                                     "grad_max_rel": float(np.max(np.abs(a[1] - b[1]) / np.maximum(np.abs(b[1]), 1e-3 * np.max(np.abs(b[1])))))}
    t = {(r["kernel"], r["want_grad"]): r["walk_ms"] for r in out["rows"]}
    out["speedup_grad"] = t[("runtime_K_fallback", True)] / t[("tensor_core_walk", True)]
    out["speedup_logl"] = t[("runtime_K_fallback", False)] / t[("tensor_core_walk", False)]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
