"""GPU tests of the substitution-model parameter gradient (mcp_eval_model_gradient), through
the C ABI: the moment matrices the CUDA gradient pass accumulates against a numpy restatement of the reference's
two passes, logL / branch gradient against the oracle (1e-10 / 1e-8), and the parameter gradient against central
differences of the ORACLE's logL."""
import numpy as np
import pytest

import mcphylo_jl_b200 as mcp
from mcphylo_jl_b200 import capi
from mcphylo_jl_b200 import substitution_models as sm
from synth import random_tree, simulate_codes
from test_gpu_parity import _check
from test_model_gradient_cpu import fd_param_gradient, numpy_moments

pytestmark = pytest.mark.gpu


def _protein_like(base_freq, rates_):
    """20-state reversible model: exchangeabilities from a fixed seed scaled by rates_[0] on the first pair."""
    K = len(base_freq)
    ex = np.random.default_rng(123).uniform(0.2, 3.0, size=K * (K - 1) // 2)
    ex[0] *= rates_[0]
    return sm.GTR(np.asarray(base_freq, float), ex)


CASES = [
    # name, model, pi, substitution rates, rate categories, taxa, sites, multifurcate, unary
    ("restriction", sm.Restriction, np.array([0.3, 0.7]), np.zeros(0), np.ones(1), 17, 777, False, False),
    ("gtr_gamma4", sm.GTR, np.array([0.1, 0.2, 0.3, 0.4]), np.array([1.0, 2.0, 1.5, 0.8, 2.5, 1.2]),
     mcp.discrete_gamma_rates(0.5, 0.5, 4), 23, 1031, True, False),
    ("jc_unary", sm.JC, np.full(4, 0.25), np.zeros(0), np.array([0.6, 1.4]), 12, 300, False, True),
    ("freek3", sm.freeK, np.array([0.2, 0.3, 0.5]), np.array([1.0, 2.0, 0.5, 0.7, 1.3, 0.9]), np.ones(1), 10, 129, False, False),
    ("gtr5_gamma2", sm.GTR, np.array([0.1, 0.15, 0.2, 0.25, 0.3]), np.linspace(0.5, 2.3, 10), mcp.discrete_gamma_rates(0.8, 0.8, 2),
     14, 333, True, False),
    ("jc6", sm.JC, np.full(6, 1.0 / 6.0), np.zeros(0), np.ones(1), 9, 65, False, True),
    ("protein20", _protein_like, np.random.default_rng(9).dirichlet(np.ones(20) * 8), np.array([1.3]), np.array([0.5, 1.5]),
     8, 96, False, False),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_moments_and_parameter_gradient(oracle, case):
    name, model, pi, sr, rates, n_taxa, S, multi, unary = case
    rng = np.random.default_rng(len(name) * 101 + S)
    K = len(pi)
    tree = random_tree(n_taxa, rng, multifurcate=multi, unary=unary)
    model_out = model(pi, sr)
    codes, leaf_nums = simulate_codes(tree, model_out, pi, rates, S, rng, gap_frac=0.03)
    ft = mcp.flatten(tree)
    U, D, Uinv, mu = model_out
    names, dA, dpi = sm.model_derivatives(model, pi, sr)
    ctx = mcp.get_context()
    aln = ctx.alignment_from_codes(codes, K, leaf_nums)
    try:
        ll, grad, pg, M, W = ctx.eval_model_gradient(aln, ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, rates, pi,
                                                     dA=dA, dpi=dpi, want_moments=True)
        # the plain evaluation (tuned kernels) gives the same logL and branch gradient
        ll_p, grad_p = ctx.eval(aln, ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, rates, pi, want_grad=True)
    finally:
        aln.close()
    x = oracle.codes_to_dense(codes, leaf_nums, K, ft.NN)
    ll_o, grad_o = oracle.felsenstein(x, ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, rates, pi, True, 0)
    _check(ll, grad, ll_o, grad_o)
    _check(ll_p, grad_p, ll_o, grad_o)
    ll_n, grad_n, M_n, W_n = numpy_moments(ft, codes, leaf_nums, K, model_out, rates, pi)
    assert np.max(np.abs(M - M_n)) <= 1e-9 * np.max(np.abs(M_n))
    assert np.max(np.abs(W - W_n)) <= 1e-10 * np.max(np.abs(W_n))
    # contraction of the device moments = contraction of the restated ones = difference quotients of the oracle
    pg_n, gc = capi.model_gradient_contract(ft.blv, U, D, Uinv, mu, rates, M, W, dA, dpi, want_grad_check=True)
    assert np.array_equal(pg, pg_n)
    assert np.max(np.abs(gc - grad_o)) <= 1e-8 * np.max(np.abs(grad_o))
    if K <= 6:
        fd = fd_param_gradient(oracle, ft, x, model, pi, sr, rates)
        assert np.max(np.abs(pg - fd)) <= 2e-6 * max(np.max(np.abs(fd)), 1.0), (pg, fd)
    else:
        pg_ref = capi.model_gradient_contract(ft.blv, U, D, Uinv, mu, rates, M_n, W_n, dA, dpi)
        assert np.max(np.abs(pg - pg_ref)) <= 1e-8 * max(np.max(np.abs(pg_ref)), 1.0)


def test_phylodist_api_and_plan_cache_keep_apart(oracle):
    """gradlogpdf_model on the drop-in object; a model-gradient evaluation and a plain one of the same tree and
    alignment alternate without disturbing each other's cached plan."""
    rng = np.random.default_rng(4)
    pi, sr = np.array([0.22, 0.28, 0.24, 0.26]), np.array([1.1, 2.2, 0.9, 1.4, 2.0, 0.7])
    rates = mcp.discrete_gamma_rates(0.7, 0.7, 4)
    tree = random_tree(30, rng)
    codes, leaf_nums = simulate_codes(tree, sm.GTR(pi, sr), pi, rates, 2000, rng)
    pd = mcp.PhyloDist(tree, pi, sr, rates, mcp.GTR)
    x = mcp.DeviceAlignment(codes, leaf_nums, 4)
    ll0, g0 = mcp.gradlogpdf(pd, x)
    ll1, g1, g_pi, g_sr = mcp.gradlogpdf_model(pd, x)
    ll2, g2 = mcp.gradlogpdf(pd, x)
    # the rate-category gradient from the moments of ONE evaluation = the one from R single-category evaluations
    ctx = mcp.get_context()
    from mcphylo_jl_b200.phylodist import _device_alignment, _tree_args
    ft_, targs_ = _tree_args(pd)
    aln_ = _device_alignment(x, ft_.leaf_nums, 4, ctx)
    rg_moments = ctx.eval_model_gradient(aln_, *targs_, dA=np.zeros((4, 4, 0)), want_rate_grad=True)[3]
    _, _, rg_ref = mcp.gradlogpdf_rates(pd, x)
    assert np.max(np.abs(rg_moments - rg_ref)) <= 1e-9 * np.max(np.abs(rg_ref)), (rg_moments, rg_ref)
    assert ll2 == ll0 and np.array_equal(g0, g2)
    assert abs(ll1 - ll0) <= 1e-10 * abs(ll0)
    assert np.max(np.abs(g1 - g0)) <= 1e-8 * np.max(np.abs(g0))
    assert g_pi.shape == (4,) and g_sr.shape == (6,)
    ft = mcp.flatten(tree)
    xd = oracle.codes_to_dense(codes, leaf_nums, 4, ft.NN)
    fd = fd_param_gradient(oracle, ft, xd, sm.GTR, pi, sr, rates)
    pg = np.concatenate([g_pi, g_sr])
    assert np.max(np.abs(pg - fd)) <= 2e-6 * max(np.max(np.abs(fd)), 1.0), (pg, fd)
    # scaling every exchangeability leaves the normalised rate matrix unchanged: the gradient is orthogonal to sr
    assert abs(np.dot(g_sr, sr)) <= 1e-8 * np.max(np.abs(g_sr)) * np.max(sr)


@pytest.mark.parametrize("G", [2, 3])
def test_multi_device_context_sums_the_moments(oracle, G):
    rng = np.random.default_rng(40 + G)
    pi, sr = np.array([0.1, 0.2, 0.3, 0.4]), np.array([1.0, 2.0, 1.5, 0.8, 2.5, 1.2])
    rates = mcp.discrete_gamma_rates(0.5, 0.5, 4)
    tree = random_tree(20, rng, multifurcate=True)
    model_out = sm.GTR(pi, sr)
    codes, leaf_nums = simulate_codes(tree, model_out, pi, rates, 1500, rng, gap_frac=0.02)
    ft = mcp.flatten(tree)
    U, D, Uinv, mu = model_out
    _, dA, dpi = sm.model_derivatives(sm.GTR, pi, sr)
    single = mcp.get_context()
    aln1 = single.alignment_from_codes(codes, 4, leaf_nums)
    multi = capi.Context(devices=[0] * G, reduce=capi.REDUCE_HOST)
    try:
        ll1, g1, pg1 = single.eval_model_gradient(aln1, ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, rates, pi, dA=dA, dpi=dpi)
        alnG = multi.alignment_from_codes(codes, 4, leaf_nums)
        llG, gG, pgG = multi.eval_model_gradient(alnG, ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, rates, pi, dA=dA, dpi=dpi)
        alnG.close()
    finally:
        aln1.close()
        multi.close()
    assert abs(llG - ll1) <= 1e-12 * abs(ll1)
    assert np.max(np.abs(gG - g1)) <= 1e-10 * np.max(np.abs(g1))
    assert np.max(np.abs(pgG - pg1)) <= 1e-10 * np.max(np.abs(pg1))


def test_model_gradient_argument_errors():
    ctx = mcp.get_context()
    rng = np.random.default_rng(1)
    tree = random_tree(5, rng)
    pi = np.array([0.3, 0.7])
    codes, leaf_nums = simulate_codes(tree, sm.Restriction(pi), pi, np.ones(1), 50, rng)
    aln = ctx.alignment_from_codes(codes, 2, leaf_nums)
    ft = mcp.flatten(tree)
    U, D, Uinv, mu = sm.Restriction(pi)
    try:
        rc = ctx.lib.mcp_eval_model_gradient(ctx.handle, aln.handle, ft.NN, None, None, None, None, None, None, 1.0, None, 1, None,
                                             0, None, None, None, None, None, None, None)
        assert rc == capi.ERR_ARG if hasattr(capi, "ERR_ARG") else rc < 0
        # the context is still usable
        ll, g = ctx.eval(aln, ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, np.ones(1), pi, want_grad=True)
        assert np.isfinite(ll) and np.all(np.isfinite(g))
    finally:
        aln.close()


def _rank_worker(rank, n_ranks, uid, q, payload):
    """One process per GPU (mcp_create_rank): evaluates its site shard; mcp_eval_model_gradient all-reduces logL, the
    branch gradient and the moment matrices over the ranks, so every rank returns the full-alignment result."""
    import numpy as np

    from mcphylo_jl_b200 import capi as _capi

    codes, leaf_nums, targs, dA, dpi, K = payload
    try:
        ctx = _capi.Context(rank, rank=(n_ranks, rank, uid))
        lo, hi = _capi.shard_bounds(codes.shape[1], n_ranks, rank)
        aln = ctx.alignment_from_codes(np.ascontiguousarray(codes[:, lo:hi]), K, leaf_nums)
        ll, g, pg = ctx.eval_model_gradient(aln, *targs, dA=dA, dpi=dpi)
        ll2, g2 = ctx.eval(aln, *targs, want_grad=True)          # the plain all-reduced evaluation still works afterwards
        aln.close()
        ctx.close()
        q.put((rank, ll, g, pg, ll2, g2, None))
    except Exception as ex:      # noqa: BLE001 - reported to the parent
        q.put((rank, None, None, None, None, None, repr(ex)))


def test_one_process_per_gpu_allreduces_the_moments(oracle):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import multiprocessing as mp_

    rng = np.random.default_rng(314)
    pi, sr = np.array([0.1, 0.2, 0.3, 0.4]), np.array([1.0, 2.0, 1.5, 0.8, 2.5, 1.2])
    rates = mcp.discrete_gamma_rates(0.5, 0.5, 4)
    tree = random_tree(25, rng)
    model_out = sm.GTR(pi, sr)
    codes, leaf_nums = simulate_codes(tree, model_out, pi, rates, 3001, rng, gap_frac=0.02)
    ft = mcp.flatten(tree)
    U, D, Uinv, mu = model_out
    targs = (ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, rates, pi)
    _, dA, dpi = sm.model_derivatives(sm.GTR, pi, sr)
    single = mcp.get_context()
    aln = single.alignment_from_codes(codes, 4, leaf_nums)
    try:
        ll1, g1, pg1 = single.eval_model_gradient(aln, *targs, dA=dA, dpi=dpi)
    finally:
        aln.close()
    mpc = mp_.get_context("spawn")
    q = mpc.Queue()
    uid = capi.nccl_unique_id()
    procs = [mpc.Process(target=_rank_worker, args=(r, 2, uid, q, (codes, leaf_nums, targs, dA, dpi, 4))) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ll, g, pg, ll2, g2, err in results:
        assert err is None, err
        assert abs(ll - ll1) <= 1e-12 * abs(ll1)
        assert np.max(np.abs(g - g1)) <= 1e-10 * np.max(np.abs(g1))
        assert np.max(np.abs(pg - pg1)) <= 1e-10 * np.max(np.abs(pg1))
        assert abs(ll2 - ll1) <= 1e-12 * abs(ll1)
        assert np.max(np.abs(g2 - g1)) <= 1e-10 * np.max(np.abs(g1))
