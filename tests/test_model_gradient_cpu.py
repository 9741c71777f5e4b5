"""Substitution-model parameter gradient, host half, without a GPU.

The CUDA path accumulates per-(branch, rate) moment matrices M and a root vector W (kernel_generic.cuh) and
the library contracts them with d P / d theta (mcp_model_gradient_contract).  Here the moments come from a
numpy restatement of the reference's two passes (/root/reference/src/Likelihood/LikelihoodCalculator_Node.jl:
3-114, no rescaling -- small trees), the contraction is the library's, and the result is checked against
central differences of the ORACLE's logL under perturbed model parameters.  Tolerances: 2e-6 relative to the
largest component (difference quotients of a ~1e3 logL with h = 1e-5)."""
import numpy as np
import pytest

import mcphylo_jl_b200 as mcp
from mcphylo_jl_b200 import capi
from mcphylo_jl_b200 import substitution_models as sm
from synth import random_tree, simulate_codes


def numpy_moments(ft, codes, leaf_nums, K, model_out, rates, pi):
    """(logL, grad, M[b, r, s, k], W[s]) by dense numpy pruning; node numbers index everything."""
    NN = ft.NN
    S = codes.shape[1]
    R = len(rates)
    children = [[] for _ in range(NN + 1)]
    for num in ft.postorder_num:           # stored child order = order of appearance in post-order
        p = ft.parent_num[num - 1]
        if p > 0:
            children[p].append(int(num))
    row_of = {int(n): i for i, n in enumerate(leaf_nums)}
    U, D, Uinv, mu = model_out
    ll = 0.0
    grad = np.zeros(NN - 1)
    M = np.zeros((NN - 1, R, K, K))
    W = np.zeros(K)
    for r, rate in enumerate(rates):
        P = {b: (U * np.exp(mu * ft.blv[b - 1] * D * rate)[None, :]) @ Uinv for b in range(1, NN)}
        dP = {b: (U * (D * rate * mu * np.exp(mu * ft.blv[b - 1] * D * rate))[None, :]) @ Uinv for b in range(1, NN)}
        L = {}
        for num in ft.postorder_num:
            num = int(num)
            if not children[num]:
                c = codes[row_of[num]]
                v = np.ones((K, S))
                oh = c < K
                v[:, oh] = 0.0
                v[c[oh], np.nonzero(oh)[0]] = 1.0
                L[num] = v
            else:
                v = np.ones((K, S))
                for ch in children[num]:
                    v = v * (P[ch] @ L[ch])
                L[num] = v
        root = NN
        site = pi @ L[root]
        ll += np.log(site).sum()
        W += (L[root] / site[None, :]).sum(axis=1)
        pre = {root: np.repeat(np.asarray(pi)[:, None], S, axis=1)}
        for num in reversed(list(ft.postorder_num)):
            num = int(num)
            if num == root:
                continue
            m = int(ft.parent_num[num - 1])
            q = pre[m].copy()
            for sib in children[m]:
                if sib != num:
                    q = q * (P[sib] @ L[sib])
            grad[num - 1] += (np.einsum("sc,sk,kc->c", q, dP[num], L[num]) / site).sum()
            M[num - 1, r] = np.einsum("sc,kc->sk", q / site[None, :], L[num])
            pre[num] = P[num].T @ q
    return ll, grad, M, W


def oracle_ll(oracle, ft, x, model, pi, sr, rates, root_pi=None):
    U, D, Uinv, mu = model(np.asarray(pi, float), np.asarray(sr, float))
    return oracle.felsenstein(x, ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, np.asarray(rates, float),
                              np.asarray(pi if root_pi is None else root_pi, float), False, 0)[0]


def fd_param_gradient(oracle, ft, x, model, pi, sr, rates, h=1e-5):
    theta = np.concatenate([pi, sr])
    K = len(pi)
    out = np.zeros(theta.size)
    for p in range(theta.size):
        tp, tm = theta.copy(), theta.copy()
        tp[p] += h
        tm[p] -= h
        out[p] = (oracle_ll(oracle, ft, x, model, tp[:K], tp[K:], rates) -
                  oracle_ll(oracle, ft, x, model, tm[:K], tm[K:], rates)) / (2 * h)
    return out


CASES = [
    ("GTR", sm.GTR, np.array([0.1, 0.2, 0.3, 0.4]), np.array([1.0, 2.0, 1.5, 0.8, 2.5, 1.2]), [0.3, 0.9, 1.8]),
    ("Restriction", sm.Restriction, np.array([0.3, 0.7]), np.zeros(0), [1.0]),
    ("JC", sm.JC, np.full(4, 0.25), np.zeros(0), [0.5, 1.5]),
    ("freeK", sm.freeK, np.array([0.2, 0.3, 0.5]), np.array([1.0, 2.0, 0.5, 0.7, 1.3, 0.9]), [1.0, 2.0]),
]


@pytest.mark.parametrize("name,model,pi,sr,rates", CASES, ids=[c[0] for c in CASES])
def test_contract_matches_finite_differences_of_the_oracle(oracle, name, model, pi, sr, rates):
    rng = np.random.default_rng(77)
    K = len(pi)
    tree = random_tree(9, rng, multifurcate=(name == "GTR"))
    model_out = model(pi, sr)
    codes, leaf_nums = simulate_codes(tree, model_out, pi, rates, 120, rng, gap_frac=0.03)
    ft = mcp.flatten(tree)
    x = oracle.codes_to_dense(codes, leaf_nums, K, ft.NN)
    rates = np.asarray(rates, float)
    ll, grad, M, W = numpy_moments(ft, codes, leaf_nums, K, model_out, rates, pi)
    U, D, Uinv, mu = model_out
    ll_o, grad_o = oracle.felsenstein(x, ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, rates, pi, True, 0)
    assert abs(ll - ll_o) <= 1e-10 * abs(ll_o)
    assert np.max(np.abs(grad - grad_o)) <= 1e-8 * np.max(np.abs(grad_o))

    names, dA, dpi = sm.model_derivatives(model, pi, sr)
    pg, gc, rg = capi.model_gradient_contract(ft.blv, U, D, Uinv, mu, rates, M, W, dA, dpi, want_grad_check=True,
                                              want_rate_grad=True)
    # the branch gradient re-derived from the moments is the oracle's
    assert np.max(np.abs(gc - grad_o)) <= 1e-9 * np.max(np.abs(grad_o))
    # d logL / d rates[r] from the same moments against central differences of the oracle's logL
    fd_r = np.zeros(rates.size)
    for r in range(rates.size):
        rp, rm = rates.copy(), rates.copy()
        rp[r] += 1e-6
        rm[r] -= 1e-6
        fd_r[r] = (oracle_ll(oracle, ft, x, model, pi, sr, rp) - oracle_ll(oracle, ft, x, model, pi, sr, rm)) / 2e-6
    assert np.max(np.abs(rg - fd_r)) <= 2e-6 * max(np.max(np.abs(fd_r)), 1.0), (rg, fd_r)
    fd = fd_param_gradient(oracle, ft, x, model, pi, sr, rates)
    assert pg.shape == fd.shape == (len(names),)
    assert np.max(np.abs(pg - fd)) <= 2e-6 * max(np.max(np.abs(fd)), 1.0), (pg, fd)


def test_user_supplied_model_function_uses_difference_quotients(oracle):
    def hky_like(base_freq, rates_):            # a model function the library has never seen
        kappa = rates_[0]
        return sm.GTR(base_freq, np.array([1.0, kappa, 1.0, 1.0, kappa, 1.0]))

    rng = np.random.default_rng(5)
    pi, sr, rates = np.array([0.15, 0.35, 0.2, 0.3]), np.array([2.7]), np.array([0.4, 1.6])
    tree = random_tree(7, rng)
    model_out = hky_like(pi, sr)
    codes, leaf_nums = simulate_codes(tree, model_out, pi, rates, 90, rng)
    ft = mcp.flatten(tree)
    x = oracle.codes_to_dense(codes, leaf_nums, 4, ft.NN)
    _, _, M, W = numpy_moments(ft, codes, leaf_nums, 4, model_out, rates, pi)
    names, dA, dpi = sm.model_derivatives(hky_like, pi, sr)
    U, D, Uinv, mu = model_out
    pg = capi.model_gradient_contract(ft.blv, U, D, Uinv, mu, rates, M, W, dA, dpi)
    fd = fd_param_gradient(oracle, ft, x, hky_like, pi, sr, rates)
    assert np.max(np.abs(pg - fd)) <= 2e-6 * max(np.max(np.abs(fd)), 1.0), (pg, fd)


def test_contract_rejects_bad_arguments():
    with pytest.raises(capi.McpError):
        capi.load()
        rc = capi.load().mcp_model_gradient_contract(0, 1, 0, None, None, None, None, 1.0, None, None, None, 0, None, None, None, None, None)
        if rc:
            raise capi.McpError(rc, capi.load().mcp_last_error(None).decode())


@pytest.mark.parametrize("name,model,pi,sr,rates", CASES, ids=[c[0] for c in CASES])
def test_closed_form_rate_matrix_derivatives_match_difference_quotients(name, model, pi, sr, rates):
    """model_derivatives: the closed forms for the four built-in models against the Richardson-extrapolated difference
    quotient the same function applies to user-supplied models (and that julia/MCPhyloB200.jl applies to every model)."""
    names, dA, dpi = sm.model_derivatives(model, pi, sr)
    _, dA_num, dpi_num = sm.model_derivatives(lambda a, b: model(a, b), pi, sr)     # a lambda is "a model it has never seen"
    K = len(pi)
    assert dA.shape == dA_num.shape == (K, K, K + len(sr)) and len(names) == K + len(sr)
    assert np.max(np.abs(dA - dA_num)) <= 1e-9 * max(np.max(np.abs(dA_num)), 1.0)
    assert np.array_equal(dpi, dpi_num) and np.array_equal(dpi[:, :K], np.eye(K)) and not dpi[:, K:].any()
    # rows of a rate matrix sum to zero, and so do the rows of its derivatives
    assert np.max(np.abs(dA.sum(axis=1))) <= 1e-12 * max(np.max(np.abs(dA)), 1.0)
    A = sm.normalised_rate_matrix(model(pi, sr))
    assert np.max(np.abs(A.sum(axis=1))) <= 1e-12
