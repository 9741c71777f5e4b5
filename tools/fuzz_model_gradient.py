#!/usr/bin/env python
"""Randomised sweep of mcp_eval_model_gradient on the GPU (experiment harness): random trees (multifurcations, unary
nodes), state counts 2 .. 20 (every compile-time-K instantiation of the column-per-thread kernel and the runtime-K one),
rate categories, site counts, gap fractions, models, multi-device contexts.  Per case: logL / branch gradient against
the CPU oracle (1e-10 / 1e-8; the extended-precision arbiter decides where the fp64 oracle itself is the limit), the branch
gradient re-derived from the device moments against the device's own (1e-8), the returned gradients against the
contraction of the returned moments (bit for bit), and moments / parameter / rate gradients against a numpy restatement
of the reference's two passes (1e-6: structural check).

    python tools/fuzz_model_gradient.py --cases 300 --seed 1
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mcphylo_jl_b200 as mcp  # noqa: E402
import oracle  # noqa: E402
from mcphylo_jl_b200 import capi  # noqa: E402
from mcphylo_jl_b200 import substitution_models as sm  # noqa: E402
from test_model_gradient_cpu import numpy_moments  # noqa: E402


def make_case(rng):
    K = int(rng.choice([2, 2, 3, 4, 4, 5, 6, 7, 12, 20]))
    R = int(rng.choice([1, 2, 4]))
    n_taxa = int(rng.choice([2, 3, 5, 8, 13, 21, 40, 77]))
    if K > 6:
        n_taxa = min(n_taxa, 21)
    S = int(rng.choice([1, 7, 31, 32, 33, 100, 257, 1000]))
    pi = rng.dirichlet(np.ones(K) * 5)
    kind = str(rng.choice(["GTR", "JC", "freeK"] if K in (3, 4) else ["Restriction"] if K == 2 else ["GTR", "JC"]))
    if kind == "GTR":
        model, sr = sm.GTR, rng.uniform(0.3, 3.0, size=K * (K - 1) // 2)
    elif kind == "JC":
        model, sr, pi = sm.JC, np.zeros(0), np.full(K, 1.0 / K)
    elif kind == "freeK":
        model, sr = sm.freeK, rng.uniform(0.3, 3.0, size=K * (K - 1))
    else:
        model, sr = sm.Restriction, np.zeros(0)
    # Gamma shapes of the BASELINE regime (slowest category >= 0.03): with slower categories the exp-form arithmetic of
    # the oracle and of the restatement is itself off by 1e-8 .. 1e-6 on short branches (tools/fuzz_parity.py
    # --regime stress), which is not what this sweep is after
    shape = float(rng.uniform(0.5, 2.0))
    rates = mcp.discrete_gamma_rates(shape, shape, R) if R > 1 else np.ones(1)
    return dict(K=K, R=R, n_taxa=n_taxa, S=S, pi=pi, kind=kind, model=model, sr=sr, rates=np.asarray(rates, float),
                multi=bool(rng.integers(0, 2)), unary=bool(rng.integers(0, 4) == 0), gap=float(rng.choice([0.0, 0.02, 0.3])),
                devices=int(rng.choice([1, 1, 1, 2, 3])))


def errors(ll, g, ll_ref, g_ref, S):
    """Relative errors with the absolute floors of tools/fuzz_parity.py (all-gap columns: logL = 0, vanishing gradients)."""
    scale = max(np.max(np.abs(g_ref)), 1e-4 * S)
    return abs(ll - ll_ref) / max(abs(ll_ref), 1e-3), float(np.max(np.abs(g - g_ref) / np.maximum(np.abs(g_ref), 1e-3 * scale)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    single = capi.Context(0)
    groups = {}
    fails = limited = 0
    worst = {}
    for ci in range(args.cases):
        c = make_case(rng)
        K, R = c["K"], c["R"]
        tree = mcp.random_tree(c["n_taxa"], rng, multifurcate=c["multi"], unary=c["unary"])
        try:
            model_out = c["model"](c["pi"], c["sr"])
        except ValueError:
            continue                        # freeK draw with complex eigenvalues: not on this path
        codes, leaf_nums = mcp.simulate_codes(tree, model_out, c["pi"], c["rates"], c["S"], rng, gap_frac=c["gap"])
        ft = mcp.flatten(tree)
        U, D, Uinv, mu = model_out
        _, dA, dpi = sm.model_derivatives(c["model"], c["pi"], c["sr"])
        G = c["devices"]
        ctx = single if G == 1 else groups.setdefault(G, capi.Context(devices=[0] * G, reduce=capi.REDUCE_HOST))
        aln = ctx.alignment_from_codes(codes, K, leaf_nums)
        ll, g, pg, rg, M, W = ctx.eval_model_gradient(aln, ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, c["rates"], c["pi"],
                                                      dA=dA, dpi=dpi, want_moments=True, want_rate_grad=True)
        aln.close()
        x = oracle.codes_to_dense(codes, leaf_nums, K, ft.NN)
        ll_o, g_o = oracle.felsenstein(x, ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, c["rates"], c["pi"], True, 0)
        e_ll, e_g = errors(ll, g, ll_o, g_o, c["S"])
        verdict = "ok"
        if e_ll > 1e-10 or e_g > 1e-8:      # the fp64 oracle's own rounding?  the extended-precision arbiter decides
            ll_x, g_x = oracle.felsenstein_extended(codes, leaf_nums, K, ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu,
                                                    c["rates"], c["pi"])
            e_ll, e_g = errors(ll, g, float(ll_x), g_x.astype(np.float64), c["S"])
            verdict = "oracle-limited"
        # the moments: (1) the branch gradient re-derived from the DEVICE moments is the device's own branch gradient (two
        # independent reductions of the same pass); (2) the library's results are the contraction of the moments it returns,
        # bit for bit; (3) moments and derived gradients against the numpy restatement, loosely (1e-6: the restatement is
        # exp-form fp64 arithmetic without rescaling -- this catches a wrong index, not the last digits)
        pg_d, gc_d, rg_d = capi.model_gradient_contract(ft.blv, U, D, Uinv, mu, c["rates"], M, W, dA, dpi, want_grad_check=True,
                                                        want_rate_grad=True)
        _, e_gc = errors(ll, gc_d, ll, g, c["S"])
        exact = bool(np.array_equal(pg, pg_d) and np.array_equal(rg, rg_d))
        _, _, M_n, W_n = numpy_moments(ft, codes, leaf_nums, K, model_out, c["rates"], c["pi"])
        pg_n, rg_n = capi.model_gradient_contract(ft.blv, U, D, Uinv, mu, c["rates"], M_n, W_n, dA, dpi, want_rate_grad=True)
        floor = 1e-6 * max(abs(ll_o), 1e-3)
        e = dict(ll=e_ll, grad=e_g, self=e_gc,
                 M=float(np.max(np.abs(M - M_n)) / max(np.max(np.abs(M_n)), floor)),
                 W=float(np.max(np.abs(W - W_n)) / max(np.max(np.abs(W_n)), floor)),
                 par=float(np.max(np.abs(pg - pg_n)) / max(np.max(np.abs(pg_n)), 1.0)) if pg.size else 0.0,
                 rate=float(np.max(np.abs(rg - rg_n)) / max(np.max(np.abs(rg_n)), 1.0)))
        bad = (e["ll"] > 1e-10 or e["grad"] > 1e-8 or e["self"] > 1e-8 or not exact or e["M"] > 1e-6 or e["W"] > 1e-6 or
               e["par"] > 1e-6 or e["rate"] > 1e-6)
        limited += verdict == "oracle-limited" and not bad
        for k in e:
            worst[k] = max(worst.get(k, 0.0), e[k])
        if bad:
            fails += 1
            print(f"FAIL #{ci} K={K} R={R} taxa={c['n_taxa']} S={c['S']} model={c['kind']} devices={G} multi={c['multi']} unary={c['unary']} gap={c['gap']} "
                  f"exact={exact} {verdict} | " + " ".join(f"{k} {v:.2e}" for k, v in e.items()), flush=True)
    print(f"seed {args.seed}: {args.cases} cases, {fails} failures, {limited} logL / branch-gradient comparisons decided by the "
          f"extended-precision arbiter; worst " + " ".join(f"{k} {v:.2e}" for k, v in worst.items()))
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
