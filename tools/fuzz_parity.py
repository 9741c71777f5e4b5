#!/usr/bin/env python
"""Randomised parity sweep on the GPU (experiment harness): random trees (multifurcations, unary
nodes), state counts, rate categories, site counts, gap fractions and kernel variants (tile width,
columns per thread, scratch placement, level-parallel kernel), each compared with the CPU oracle at
the acceptance tolerances (logL 1e-10, gradients 1e-8 relative).

    python tools/fuzz_parity.py --cases 300 --seed 1
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcphylo_jl_b200 as mcp  # noqa: E402
import oracle  # noqa: E402


def one_case(rng, ctx, idx, stress=False):
    K = int(rng.choice([2, 2, 3, 4, 4, 4, 5, 6, 7, 12, 20]))
    R = int(rng.choice([1, 1, 2, 4]))
    n_taxa = int(rng.choice([2, 3, 5, 8, 13, 21, 40, 77, 150]))
    if K > 6:
        n_taxa = min(n_taxa, 40)
    S = int(rng.choice([1, 7, 31, 32, 33, 100, 257, 1000, 2500]))
    if K > 6:
        S = min(S, 257)
    tree = mcp.random_tree(n_taxa, rng, multifurcate=bool(rng.integers(0, 2)), unary=bool(rng.integers(0, 4) == 0),
                           mean_bl=float(rng.choice([0.001, 0.01, 0.1] if stress else [0.05, 0.1, 0.5])))
    pi = rng.dirichlet(np.ones(K) * 3)
    if K == 2:
        model, srates = mcp.Restriction, np.zeros(1)
    elif rng.integers(0, 2):
        model, srates = mcp.JC, np.zeros(1)
    else:
        model, srates = mcp.GTR, rng.uniform(0.2, 3.0, size=K * (K - 1) // 2)
    # Baseline regime: mean-one Gamma categories with shape >= 0.5 (slowest rate >= 0.03) and branch
    # lengths like BASELINE's Exponential(0.1) clipped at 1e-4.  Stress regime: shapes down to 0.2
    # with free scale and very short branches; there the off-diagonal entries of
    # P = U diag(e) Uinv are cancellation noise in fp64 (in the reference as much as here) and
    # implementations can only agree to ~1e-6 (Conditioning).
    if stress:
        rates = mcp.discrete_gamma_rates(float(rng.uniform(0.2, 2.0)), float(rng.uniform(0.2, 2.0)), R) if R > 1 else np.ones(1)
    else:
        shape = float(rng.uniform(0.5, 2.0))
        rates = mcp.discrete_gamma_rates(shape, shape, R) if R > 1 else np.ones(1)
    codes, leaf_nums = mcp.simulate_codes(tree, model(pi, srates), pi, rates, S, rng, gap_frac=float(rng.choice([0.0, 0.05, 0.5])))
    block = int(rng.choice([0, 0, 32, 64, 128, 256]))
    cpt = int(rng.choice([0, 1, 2]))
    scratch = int(rng.choice([-1, 0, 1]))
    levels = int(rng.choice([-1, 0, 1]))
    ctx.set_launch(block, 0)
    ctx.set_columns_per_thread(cpt)
    ctx.set_scratch_mode(scratch)
    ctx.set_level_mode(levels)
    pd = mcp.PhyloDist(tree, pi, srates, rates, model)
    aln = mcp.DeviceAlignment(codes, leaf_nums, K)
    ll, g = mcp.gradlogpdf(pd, aln)
    ll2 = mcp.logpdf(pd, aln)
    ft = mcp.flatten(tree)
    x = oracle.codes_to_dense(codes, leaf_nums, K, ft.NN)
    U, D, Uinv, mu = model(pi, srates)
    ll_o, g_o = oracle.felsenstein(x, ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, rates, pi, True, 0)
    desc = f"#{idx} K={K} R={R} taxa={n_taxa} S={S} block={block} cpt={cpt} scratch={scratch} levels={levels} model={model.__name__}"
    # relative errors with an absolute floor: an all-gap alignment has logL = 0 and zero gradients,
    # where both sides are rounding noise of size 1e-16
    # (gradient components that cancel to ~0 are rounding noise of size 1e-16 per site on both sides:
    # the floor is 1e-3 of the largest component, and never below 1e-4 per site)
    scale = max(np.max(np.abs(g_o)), 1e-4 * S)
    e_ll = abs(ll - ll_o) / max(abs(ll_o), 1e-3)
    e_ll2 = abs(ll2 - ll_o) / max(abs(ll_o), 1e-3)
    e_g = float(np.max(np.abs(g - g_o) / np.maximum(np.abs(g_o), 1e-3 * scale)))
    ok = (e_ll <= 1e-7 and e_ll2 <= 1e-7 and e_g <= 1e-5) if stress else (e_ll <= 1e-10 and e_ll2 <= 1e-10 and e_g <= 1e-8)
    if not ok:
        j = int(np.argmax(np.abs(g - g_o) / np.maximum(np.abs(g_o), 1e-3 * scale)))
        desc += f" | worst comp {j}: gpu {g[j]:.15e} oracle {g_o[j]:.15e} scale {scale:.3e} blv {ft.blv[j]:.3e} ll {ll:.15e} vs {ll_o:.15e}"
    return ok, desc, e_ll, e_g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--regime", default="baseline", choices=["baseline", "stress"])
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    ctx = mcp.get_context(0)
    worst_ll = worst_g = 0.0
    bad = 0
    for i in range(a.cases):
        ok, desc, e_ll, e_g = one_case(rng, ctx, i, a.regime == "stress")
        worst_ll, worst_g = max(worst_ll, e_ll), max(worst_g, e_g)
        if not ok:
            bad += 1
            print("FAIL", desc, f"logL rel {e_ll:.2e} grad rel {e_g:.2e}", flush=True)
    print(f"regime {a.regime}, seed {a.seed}: {a.cases} cases, {bad} failures, worst logL rel err {worst_ll:.2e}, worst gradient rel err {worst_g:.2e}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
