#!/usr/bin/env python
"""Randomised parity sweep on the GPU (experiment harness): random trees (multifurcations, unary
nodes), state counts, rate categories, site counts, gap fractions and kernel variants (tile width,
columns per thread, scratch placement, level-parallel kernel), each compared with the CPU oracle at
the acceptance tolerances (logL 1e-10, gradients 1e-8 relative).  Where the two disagree, an
extended-precision evaluation (oracle/extended.py) decides which side is off.

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


def make_case(rng, stress=False):
    """Draws one random problem and kernel variant (no GPU involved; consumes the rng identically for
    every run, so (seed, index) names a case)."""
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
    # Baseline regime: mean-one Gamma categories with shape >= 0.5 and branch lengths like BASELINE's
    # Exponential(0.1) clipped at 1e-4.  Stress regime: shapes down to 0.2 with free scale and very
    # short branches.  The slowest category of a shape-0.5 Gamma already has rate ~1e-4: on short
    # branches the off-diagonal entries of the REFERENCE's P = U diag(e) Uinv are then cancellation
    # noise in fp64 (Conditioning), which is what the arbiter below sorts out.
    if stress:
        rates = mcp.discrete_gamma_rates(float(rng.uniform(0.2, 2.0)), float(rng.uniform(0.2, 2.0)), R) if R > 1 else np.ones(1)
    else:
        shape = float(rng.uniform(0.5, 2.0))
        rates = mcp.discrete_gamma_rates(shape, shape, R) if R > 1 else np.ones(1)
    codes, leaf_nums = mcp.simulate_codes(tree, model(pi, srates), pi, rates, S, rng, gap_frac=float(rng.choice([0.0, 0.05, 0.5])))
    return dict(K=K, R=R, n_taxa=n_taxa, S=S, tree=tree, pi=pi, model=model, srates=srates, rates=rates,
                codes=codes, leaf_nums=leaf_nums,
                block=int(rng.choice([0, 0, 32, 64, 128, 256])), cpt=int(rng.choice([0, 1, 2])),
                scratch=int(rng.choice([-1, 0, 1])), levels=int(rng.choice([-1, 0, 1])))


def _errors(ll, g, ll_ref, g_ref, S):
    # relative errors with an absolute floor: an all-gap alignment has logL = 0 and zero gradients,
    # where both sides are rounding noise of size 1e-16; gradient components that cancel to ~0 are
    # rounding noise of size 1e-16 per site on both sides: the floor is 1e-3 of the largest
    # component, and never below 1e-4 per site
    scale = max(np.max(np.abs(g_ref)), 1e-4 * S)
    rel = np.abs(g - g_ref) / np.maximum(np.abs(g_ref), 1e-3 * scale)
    return abs(ll - ll_ref) / max(abs(ll_ref), 1e-3), float(np.max(rel)), int(np.argmax(rel))


def one_case(rng, ctx, idx, stress=False):
    """Returns (ok, description, logL rel err, gradient rel err, verdict).  The CUDA result is compared
    with the fp64 oracle at the acceptance tolerances (logL 1e-10, gradients 1e-8); when they disagree,
    the extended-precision arbiter (oracle/extended.py) decides: verdict "oracle-limited" (still ok)
    if the CUDA result agrees with the arbiter at the acceptance tolerances -- the deviation is then
    the reference formula's own rounding -- else "FAIL"."""
    c = make_case(rng, stress)
    K, S, tree, pi, model, srates, rates = c["K"], c["S"], c["tree"], c["pi"], c["model"], c["srates"], c["rates"]
    ctx.set_launch(c["block"], 0)
    ctx.set_columns_per_thread(c["cpt"])
    ctx.set_scratch_mode(c["scratch"])
    ctx.set_level_mode(c["levels"])
    pd = mcp.PhyloDist(tree, pi, srates, rates, model)
    aln = mcp.DeviceAlignment(c["codes"], c["leaf_nums"], K)
    ll, g = mcp.gradlogpdf(pd, aln)
    ll2 = mcp.logpdf(pd, aln)
    ft = mcp.flatten(tree)
    x = oracle.codes_to_dense(c["codes"], c["leaf_nums"], K, ft.NN)
    U, D, Uinv, mu = model(pi, srates)
    ll_o, g_o = oracle.felsenstein(x, ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, rates, pi, True, 0)
    desc = (f"#{idx} K={K} R={c['R']} taxa={c['n_taxa']} S={S} block={c['block']} cpt={c['cpt']} scratch={c['scratch']} "
            f"levels={c['levels']} model={model.__name__}")
    e_ll, e_g, j = _errors(ll, g, ll_o, g_o, S)
    e_ll2 = abs(ll2 - ll_o) / max(abs(ll_o), 1e-3)
    if e_ll <= 1e-10 and e_ll2 <= 1e-10 and e_g <= 1e-8:
        return True, desc, e_ll, e_g, "ok"
    ll_x, g_x = oracle.felsenstein_extended(c["codes"], c["leaf_nums"], K, ft.postorder_num, ft.parent_num, ft.blv,
                                            U, D, Uinv, mu, rates, pi)
    ll_x, g_x = float(ll_x), g_x.astype(np.float64)
    x_ll, x_g, _ = _errors(ll, g, ll_x, g_x, S)
    x_ll2 = abs(ll2 - ll_x) / max(abs(ll_x), 1e-3)
    o_ll, o_g, _ = _errors(ll_o, g_o, ll_x, g_x, S)
    desc += (f" | vs oracle: logL {e_ll:.2e} grad {e_g:.2e} (comp {j}, blv {ft.blv[j]:.3e});"
             f" vs extended precision: CUDA logL {x_ll:.2e} grad {x_g:.2e}, oracle logL {o_ll:.2e} grad {o_g:.2e}")
    ok = x_ll <= 1e-10 and x_ll2 <= 1e-10 and x_g <= 1e-8
    return ok, desc, e_ll, e_g, ("oracle-limited" if ok else "FAIL")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--regime", default="baseline", choices=["baseline", "stress"])
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    ctx = mcp.get_context(0)
    worst_ll = worst_g = 0.0
    bad = limited = 0
    for i in range(a.cases):
        ok, desc, e_ll, e_g, verdict = one_case(rng, ctx, i, a.regime == "stress")
        if verdict == "ok":
            worst_ll, worst_g = max(worst_ll, e_ll), max(worst_g, e_g)
        else:
            bad += verdict == "FAIL"
            limited += verdict == "oracle-limited"
            print(verdict, desc, flush=True)
    print(f"regime {a.regime}, seed {a.seed}: {a.cases} cases, {bad} failures, {limited} decided by the extended-precision "
          f"arbiter in favour of the CUDA path; among the others worst logL rel err {worst_ll:.2e}, worst gradient rel err {worst_g:.2e}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
