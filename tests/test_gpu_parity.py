"""GPU parity tests: the CUDA path, called through the C-ABI (capi / PhyloDist API), against the
CPU oracle on identical trees and alignments, against the reference's goldens, and — at sizes the
oracle cannot hold — through size-independent properties.

Tolerances (BASELINE.json north_star): logL relative error <= 1e-10, every gradient component
relative error <= 1e-8 (absolute 1e-8 * max|grad| floor for components that cancel to ~0)."""
import numpy as np
import pytest

import mcphylo_jl_b200 as mcp
from conftest import golden_case
from synth import random_tree, simulate_codes

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-10
GRAD_RTOL = 1e-8


def _model(K, pi, rng):
    if K == 2:
        return mcp.Restriction, pi, np.zeros(1)
    if K == 4:
        return mcp.GTR, pi, rng.uniform(0.5, 2.5, size=6)
    return mcp.JC, pi, np.zeros(1)


def _oracle_eval(oracle, tree, codes, leaf_nums, K, model, pi, srates, rates, want_grad=True):
    ft = mcp.flatten(tree)
    x = oracle.codes_to_dense(codes, leaf_nums, K, ft.NN)
    U, D, Uinv, mu = model(np.asarray(pi, float), np.asarray(srates, float))
    return oracle.felsenstein(x, ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu,
                              np.asarray(rates, float), np.asarray(pi, float), want_grad, 0)


def _check(ll, g, ll_o, g_o):
    assert abs(ll - ll_o) <= LL_RTOL * abs(ll_o), (ll, ll_o)
    if g_o is not None:
        scale = np.max(np.abs(g_o))
        assert np.all(np.abs(g - g_o) <= GRAD_RTOL * np.maximum(np.abs(g_o), 1e-3 * scale)), \
            np.max(np.abs(g - g_o) / np.maximum(np.abs(g_o), 1e-3 * scale))


def test_primates_golden_dense_input():
    tree, x, _, _, fx = golden_case("primates")
    pd = mcp.PhyloDist(tree, fx["base_freq"], [1.0], [1.0], mcp.JC)
    assert pd.size() == tuple(fx["size"])
    ll = mcp.logpdf(pd, x)
    assert abs(ll - fx["logpdf"]) <= LL_RTOL * abs(ll)
    ll2, grad = mcp.gradlogpdf(pd, x)
    assert abs(ll2 - fx["logpdf"]) <= LL_RTOL * abs(ll)
    assert grad.shape == (21,) and np.all(np.isfinite(grad))


def test_simudata_golden_gradient(oracle):
    tree, x, codes, leaf_nums, fx = golden_case("simudata")
    pd = mcp.PhyloDist(tree, fx["base_freq"], [1.0], [1.0], mcp.JC)
    ll, grad = mcp.gradlogpdf(pd, x)
    assert np.max(np.abs(grad - fx["grad"]) / np.abs(fx["grad"])) <= GRAD_RTOL
    assert abs(ll - fx["logpdf_loose"]) <= fx["logpdf_rtol"] * abs(ll)
    ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, 4, mcp.JC, fx["base_freq"], [1.0], [1.0])
    _check(ll, grad, ll_o, g_o)
    # compact codes give the same answer as the dense array
    ll_c, grad_c = mcp.gradlogpdf(pd, mcp.DeviceAlignment(codes, leaf_nums, 4))
    assert ll_c == ll and np.array_equal(grad_c, grad)


def test_multiple_phylodist_golden():
    tree, x, _, _, fx = golden_case("primates")
    mpd = mcp.MultiplePhyloDist([tree, tree], np.full((4, 2), 0.25), np.ones((1, 2)), np.ones((1, 2)), mcp.JC)
    assert mpd.size() == (4, 1, 22, 2)
    mdf = np.asfortranarray(np.stack([x, x], axis=3))
    assert abs(mcp.logpdf(mpd, mdf) - 2 * fx["logpdf"]) <= LL_RTOL * abs(2 * fx["logpdf"])
    res = mcp.multi_gradlogpdf(mpd, mdf)
    assert len(res) == 2
    for ll, g in res:
        assert abs(ll - fx["logpdf"]) <= LL_RTOL * abs(ll) and g.shape == (21,)
    assert np.array_equal(res[0][1], res[1][1])
    with pytest.raises(mcp.DimensionMismatch):
        mcp.MultiplePhyloDist([tree, tree], np.full((4, 3), 0.25), [1.0], [1.0], mcp.JC)
    with pytest.raises(mcp.DimensionMismatch):
        mcp.MultiplePhyloDist([tree, tree], np.full((4, 2), 0.25), np.ones((1, 3)), [1.0], mcp.JC)


@pytest.mark.parametrize("n_taxa,K,R,S,seed,multi,unary", [
    (10, 2, 1, 1000, 20241, False, False),     # BASELINE config 1 shape
    (50, 2, 1, 4000, 20242, False, False),     # config 2 shape (sites subsampled)
    (200, 4, 4, 600, 20243, False, False),     # config 3 shape (sites subsampled)
    (37, 4, 2, 257, 7, True, False),           # trifurcations, ragged tile
    (23, 3, 3, 130, 8, True, True),            # K=3, unary node
    (2, 4, 1, 33, 9, False, False),            # smallest tree
    (64, 5, 1, 100, 10, False, False),
    (16, 6, 2, 64, 12, False, False),
    (300, 2, 1, 31, 11, False, False),         # fewer sites than one warp
])
def test_random_cases_vs_oracle(oracle, n_taxa, K, R, S, seed, multi, unary):
    rng = np.random.default_rng(seed)
    tree = random_tree(n_taxa, rng, multifurcate=multi, unary=unary)
    pi = rng.dirichlet(np.ones(K) * 5)
    model, pi, srates = _model(K, pi, rng)
    rates = mcp.discrete_gamma_rates(0.5, 0.5, R) if R > 1 else np.ones(1)
    codes, leaf_nums = simulate_codes(tree, model(pi, srates), pi, rates, S, rng, gap_frac=0.05)
    pd = mcp.PhyloDist(tree, pi, srates, rates, model)
    aln = mcp.DeviceAlignment(codes, leaf_nums, K)
    ll, grad = mcp.gradlogpdf(pd, aln)
    ll_only = mcp.logpdf(pd, aln)
    ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, K, model, pi, srates, rates)
    _check(ll, grad, ll_o, g_o)
    _check(ll_only, None, ll_o, None)


@pytest.mark.parametrize("K,R,n_taxa,S", [(7, 1, 12, 150), (20, 2, 25, 200), (32, 1, 9, 70)])
def test_large_alphabets_generic_kernel(oracle, K, R, n_taxa, S):
    """K > 6 (e.g. 20-state protein alphabets) runs on the runtime-K kernel."""
    rng = np.random.default_rng(200 + K)
    tree = random_tree(n_taxa, rng, multifurcate=True, unary=(K == 20))
    pi = rng.dirichlet(np.ones(K) * 5)
    # a reversible K-state model: symmetric exchangeabilities times pi (GTR with K states)
    srates = rng.uniform(0.2, 3.0, size=K * (K - 1) // 2)
    rates = mcp.discrete_gamma_rates(0.5, 0.5, R) if R > 1 else np.ones(1)
    codes, leaf_nums = simulate_codes(tree, mcp.GTR(pi, srates), pi, rates, S, rng, gap_frac=0.05)
    pd = mcp.PhyloDist(tree, pi, srates, rates, mcp.GTR)
    aln = mcp.DeviceAlignment(codes, leaf_nums, K)
    ll, g = mcp.gradlogpdf(pd, aln)
    ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, K, mcp.GTR, pi, srates, rates)
    _check(ll, g, ll_o, g_o)
    _check(mcp.logpdf(pd, aln), None, ll_o, None)


@pytest.mark.parametrize("K,R,n_taxa,S,multi", [(7, 2, 30, 300, True), (8, 1, 17, 129, False), (13, 1, 40, 500, False),
                                                (16, 2, 12, 77, True), (20, 4, 60, 400, False), (21, 1, 25, 130, True),
                                                (24, 1, 9, 2000, False), (29, 1, 33, 260, False), (32, 2, 20, 140, True)])
@pytest.mark.parametrize("mode", [1, 0])
def test_large_alphabets_tensor_core_kernel(oracle, K, R, n_taxa, S, multi, mode):
    """6 < K <= 32 on the tile-cooperative FP64 tensor-core walk (mode 1: K padded to 8 / 16 / 24 / 32, ragged
    tiles, several rate categories, multifurcations and unary nodes) and on the runtime-K fallback (mode 0),
    both against the oracle; the tensor-core kernel's gradient is bit-reproducible."""
    rng = np.random.default_rng(1000 + K)
    tree = random_tree(n_taxa, rng, multifurcate=multi, unary=(K % 5 == 0))
    pi = rng.dirichlet(np.ones(K) * 5)
    srates = rng.uniform(0.2, 3.0, size=K * (K - 1) // 2)
    rates = mcp.discrete_gamma_rates(0.5, 0.5, R) if R > 1 else np.ones(1)
    codes, leaf_nums = simulate_codes(tree, mcp.GTR(pi, srates), pi, rates, S, rng, gap_frac=0.05)
    pd = mcp.PhyloDist(tree, pi, srates, rates, mcp.GTR)
    aln = mcp.DeviceAlignment(codes, leaf_nums, K)
    ctx = mcp.get_context()
    ctx.set_large_alphabet_mode(mode)
    try:
        ll, g = mcp.gradlogpdf(pd, aln)
        st = ctx.stats()
        assert st["block"] == (256 if mode else min(128, st["block"]))
        ll2, g2 = mcp.gradlogpdf(pd, aln)
        ll_only = mcp.logpdf(pd, aln)
    finally:
        ctx.set_large_alphabet_mode(-1)
    ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, K, mcp.GTR, pi, srates, rates)
    _check(ll, g, ll_o, g_o)
    _check(ll_only, None, ll_o, None)
    if mode:
        assert ll2 == ll and np.array_equal(g, g2)


def test_large_alphabet_batch_and_posterior(oracle):
    """The tensor-core kernel under mcp_eval_batch (trees of different sizes in one launch) and with the
    branch-length prior epilogue."""
    from mcphylo_jl_b200 import capi
    from mcphylo_jl_b200.phylodist import _tree_args
    K = 20
    ctx = capi.Context(0)
    try:
        alns, targs, want = [], [], []
        for i, (n_taxa, S) in enumerate([(11, 200), (40, 333), (5, 64)]):
            rng = np.random.default_rng(4000 + i)
            tree = random_tree(n_taxa, rng)
            pi = rng.dirichlet(np.ones(K) * 5)
            srates = rng.uniform(0.2, 3.0, size=K * (K - 1) // 2)
            codes, leaf_nums = simulate_codes(tree, mcp.GTR(pi, srates), pi, np.ones(1), S, rng, gap_frac=0.02)
            alns.append(ctx.alignment_from_codes(codes, K, leaf_nums))
            targs.append(_tree_args(mcp.PhyloDist(tree, pi, srates, [1.0], mcp.GTR))[1])
            want.append(_oracle_eval(oracle, tree, codes, leaf_nums, K, mcp.GTR, pi, srates, [1.0]))
        ll, grads = ctx.eval_batch(alns, targs, want_grad=True)
        for t in range(3):
            _check(ll[t], grads[t], *want[t])
        lp, gp = ctx.eval_posterior(alns[1], *targs[1], prior_kind=1, prior_params=[0.1])
        blv = np.asarray(targs[1][2])
        assert abs(lp - (want[1][0] + np.sum(-np.log(0.1) - blv / 0.1))) <= 1e-10 * abs(lp)
        assert np.allclose(gp, want[1][1] - 10.0, rtol=1e-8, atol=1e-8 * np.max(np.abs(want[1][1])))
        for a in alns:
            a.close()
    finally:
        ctx.close()


def test_rate_category_and_gamma_shape_gradient(oracle):
    """d logL / d rates[r] from mcp_eval_rate_gradient against central differences of the
    ORACLE's logL, and the Gamma-shape gradient through the chain rule; logL and the branch gradient of the
    same call equal the plain evaluation's."""
    rng = np.random.default_rng(99)
    tree = random_tree(30, rng, multifurcate=True)
    pi = np.array([0.1, 0.2, 0.3, 0.4])
    sr = np.array([1.0, 2.0, 1.5, 0.8, 2.5, 1.2])
    alpha = 0.7
    rates = mcp.discrete_gamma_rates(alpha, alpha, 4)
    codes, leaf_nums = simulate_codes(tree, mcp.GTR(pi, sr), pi, rates, 800, rng, gap_frac=0.02)
    aln = mcp.DeviceAlignment(codes, leaf_nums, 4)
    pd = mcp.PhyloDist(tree, pi, sr, rates, mcp.GTR)
    ll, g, rg = mcp.gradlogpdf_rates(pd, aln)
    ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, 4, mcp.GTR, pi, sr, rates)
    _check(ll, g, ll_o, g_o)
    assert rg.shape == (4,)

    def oracle_ll(r):
        return _oracle_eval(oracle, tree, codes, leaf_nums, 4, mcp.GTR, pi, sr, r, want_grad=False)[0]
    for r in range(4):
        h = 1e-5 * rates[r]
        up, dn = rates.copy(), rates.copy()
        up[r] += h
        dn[r] -= h
        fd = (oracle_ll(up) - oracle_ll(dn)) / (2 * h)
        assert abs(rg[r] - fd) <= 2e-6 * max(1.0, abs(fd)), (r, rg[r], fd)
    # Gamma shape: rates = discrete_gamma_rates(alpha, alpha, 4)
    dalpha = float(rg @ mcp.discrete_gamma_rates_dalpha(alpha, 4))
    ha = 1e-5
    fd = (oracle_ll(mcp.discrete_gamma_rates(alpha + ha, alpha + ha, 4)) - oracle_ll(mcp.discrete_gamma_rates(alpha - ha, alpha - ha, 4))) / (2 * ha)
    assert abs(dalpha - fd) <= 1e-5 * max(1.0, abs(fd)), (dalpha, fd)


def test_every_launch_shape_agrees(oracle):
    rng = np.random.default_rng(3)
    tree = random_tree(40, rng)
    pi = np.array([0.1, 0.2, 0.3, 0.4])
    sr = np.array([1.0, 2.0, 1.5, 0.8, 2.5, 1.2])
    rates = mcp.discrete_gamma_rates(0.5, 0.5, 4)
    codes, leaf_nums = simulate_codes(tree, mcp.GTR(pi, sr), pi, rates, 3000, rng)
    pd = mcp.PhyloDist(tree, pi, sr, rates, mcp.GTR)
    aln = mcp.DeviceAlignment(codes, leaf_nums, 4)
    ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, 4, mcp.GTR, pi, sr, rates)
    ctx = mcp.get_context()
    try:
        for block, ctas, cpt in [(32, 1, 1), (64, 0, 2), (128, 2, 1), (256, 0, 2), (256, 1, 1), (128, 0, 2)]:
            ctx.set_launch(block, ctas)
            ctx.set_columns_per_thread(cpt)
            ll, g = mcp.gradlogpdf(pd, aln)
            _check(ll, g, ll_o, g_o)
            assert ctx.stats()["block"] == block
            _check(mcp.logpdf(pd, aln), None, ll_o, None)
    finally:
        ctx.set_launch(0, 0)
        ctx.set_columns_per_thread(0)


@pytest.mark.parametrize("K,R,S,n_taxa", [(4, 4, 3000, 40), (2, 1, 2100, 33), (4, 2, 700, 150), (2, 2, 96, 7)])
def test_operand_ring_is_bit_identical(oracle, K, R, S, n_taxa):
    """Gradient pass with the operand ring (stored child partials fetched ahead by cp.async.bulk into shared
    memory, mcp_set_ring_mode) against the oracle and, bit for bit, against the same launch without the ring:
    one and two columns per thread, several tile widths incl. ragged last tiles and widths of one warp, trees
    with multifurcations and unary nodes (virtual leaves), rings shallower than the fetch list and empty ones."""
    rng = np.random.default_rng(500 + K + R)
    tree = random_tree(n_taxa, rng, multifurcate=True)
    pi = rng.dirichlet(np.ones(K) * 5)
    model, pi, srates = _model(K, pi, rng)
    rates = mcp.discrete_gamma_rates(0.5, 0.5, R) if R > 1 else np.ones(1)
    codes, leaf_nums = simulate_codes(tree, model(pi, srates), pi, rates, S, rng, gap_frac=0.03)
    pd = mcp.PhyloDist(tree, pi, srates, rates, model)
    aln = mcp.DeviceAlignment(codes, leaf_nums, K)
    ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, K, model, pi, srates, rates)
    ctx = mcp.get_context()
    try:
        ctx.set_scratch_mode(0)          # partials in HBM (the small-input shared-memory scratch has no ring)
        for block, cpt in [(256, 2), (256, 1), (32, 2), (96, 1), (128, 2)]:
            ctx.set_launch(block, 0)
            ctx.set_columns_per_thread(cpt)
            ctx.set_ring_mode(0)
            ll0, g0 = mcp.gradlogpdf(pd, aln)
            assert ctx.stats()["operand_ring"] == 0
            ctx.set_ring_mode(1)
            ll1, g1 = mcp.gradlogpdf(pd, aln)
            assert ctx.stats()["operand_ring"] > 0, "the ring kernel did not run"
            _check(ll1, g1, ll_o, g_o)
            assert ll1 == ll0 and np.array_equal(g1, g0), (block, cpt)
            ll2, g2 = mcp.gradlogpdf(pd, aln)            # ring phase carried over from the call before
            assert ll2 == ll1 and np.array_equal(g2, g1)
    finally:
        ctx.set_launch(0, 0)
        ctx.set_columns_per_thread(0)
        ctx.set_ring_mode(-1)
        ctx.set_scratch_mode(-1)


def test_operand_ring_batch_of_trees(oracle):
    """Several trees of different sizes in one batched launch through the ring kernel: each tree has its own
    fetch list and the ring's phase carries over from tree to tree inside a CTA."""
    rng = np.random.default_rng(77)
    pi = np.array([0.1, 0.2, 0.3, 0.4])
    sr = np.array([1.0, 2.0, 1.5, 0.8, 2.5, 1.2])
    rates = mcp.discrete_gamma_rates(0.5, 0.5, 2)
    trees, alns, want = [], [], []
    for n in (5, 23, 64, 3, 31):
        tree = random_tree(n, rng, multifurcate=True)
        codes, leaf_nums = simulate_codes(tree, mcp.GTR(pi, sr), pi, rates, 400 + 37 * n, rng, gap_frac=0.02)
        trees.append(tree)
        alns.append(mcp.DeviceAlignment(codes, leaf_nums, 4))
        want.append(_oracle_eval(oracle, tree, codes, leaf_nums, 4, mcp.GTR, pi, sr, rates))
    ctx = mcp.get_context()
    try:
        ctx.set_ring_mode(1)
        ctx.set_level_mode(0)
        ctx.set_scratch_mode(0)
        res = mcp.multi_gradlogpdf(mcp.MultiplePhyloDist(trees, pi, sr, rates, mcp.GTR), alns)
        assert ctx.stats()["operand_ring"] > 0
        for (ll, g), (ll_o, g_o) in zip(res, want):
            _check(ll, g, ll_o, g_o)
    finally:
        ctx.set_ring_mode(-1)
        ctx.set_level_mode(-1)
        ctx.set_scratch_mode(-1)


@pytest.mark.parametrize("K,R,S", [(2, 1, 1500), (3, 2, 700), (5, 1, 300)])
def test_two_columns_per_thread(oracle, K, R, S):
    """The two-columns-per-thread instantiation (automatic only for large K <= 3 problems) against
    the oracle, with a ragged last tile."""
    rng = np.random.default_rng(100 + K)
    tree = random_tree(33, rng, multifurcate=True)
    pi = rng.dirichlet(np.ones(K) * 5)
    model, pi, srates = _model(K, pi, rng)
    rates = mcp.discrete_gamma_rates(0.5, 0.5, R) if R > 1 else np.ones(1)
    codes, leaf_nums = simulate_codes(tree, model(pi, srates), pi, rates, S, rng, gap_frac=0.05)
    pd = mcp.PhyloDist(tree, pi, srates, rates, model)
    aln = mcp.DeviceAlignment(codes, leaf_nums, K)
    ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, K, model, pi, srates, rates)
    ctx = mcp.get_context()
    try:
        ctx.set_columns_per_thread(2)
        for block in (32, 128):
            ctx.set_launch(block, 0)
            ll, g = mcp.gradlogpdf(pd, aln)
            _check(ll, g, ll_o, g_o)
    finally:
        ctx.set_launch(0, 0)
        ctx.set_columns_per_thread(0)


def _shifted(model, shift):
    """A decomposition WITHOUT a null eigenvalue: P(t) = exp(-shift mu t r) * P_model(t).  Not a
    stochastic matrix, but a valid input of the ABI (and of the reference, which takes any U, D, Uinv)."""
    def f(pi, srates):
        U, D, Uinv, mu = model(pi, srates)
        return U, np.asarray(D, float) - shift, Uinv, mu
    return f


@pytest.mark.parametrize("K,R,cpt", [(2, 1, 1), (2, 1, 2), (4, 4, 1), (4, 2, 2), (3, 2, 1), (6, 1, 1)])
def test_decomposition_without_null_eigenvalue(oracle, K, R, cpt):
    """The walk kernels skip the null eigenvalue every rate matrix has (moved last by the host);
    any other decomposition must take the full-K kernels and still match the oracle.  Also checks
    the eigenvalue ORDER is irrelevant (the null one first / in the middle / last)."""
    rng = np.random.default_rng(900 + 10 * K + cpt)
    tree = random_tree(29, rng, multifurcate=True)
    pi = rng.dirichlet(np.ones(K) * 5)
    base, pi, srates = _model(K, pi, rng)
    rates = mcp.discrete_gamma_rates(0.5, 0.5, R) if R > 1 else np.ones(1)
    codes, leaf_nums = simulate_codes(tree, base(pi, srates), pi, rates, 333, rng, gap_frac=0.05)
    aln = mcp.DeviceAlignment(codes, leaf_nums, K)

    def rolled(shift_pos):
        def f(pi_, sr_):
            U, D, Uinv, mu = base(pi_, sr_)
            idx = np.roll(np.arange(K), shift_pos)
            return np.asfortranarray(U[:, idx]), np.asarray(D)[idx], np.asfortranarray(Uinv[idx, :]), mu
        return f

    ctx = mcp.get_context()
    try:
        ctx.set_level_mode(0)
        ctx.set_columns_per_thread(cpt)
        ctx.set_launch(64, 0)
        for model in (_shifted(base, 0.37), rolled(0), rolled(1), rolled(K - 1)):
            pd = mcp.PhyloDist(tree, pi, srates, rates, model)
            ll, g = mcp.gradlogpdf(pd, aln)
            ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, K, model, pi, srates, rates)
            _check(ll, g, ll_o, g_o)
            _check(mcp.logpdf(pd, aln), None, ll_o, None)
    finally:
        ctx.set_level_mode(-1)
        ctx.set_launch(0, 0)
        ctx.set_columns_per_thread(0)


def test_batch_mixing_models_with_and_without_null_eigenvalue(oracle):
    """One batch launch, several distinct models of which one has no null eigenvalue: the whole
    batch takes the full-K constant-memory kernels."""
    rng = np.random.default_rng(4242)
    K = 4
    trees, alns, expect, pis, srs = [], [], [], [], []
    models = [mcp.GTR, _shifted(mcp.GTR, 0.2), mcp.GTR]
    for i, model in enumerate(models):
        t = random_tree(15 + 4 * i, rng)
        pi = rng.dirichlet(np.ones(K) * 5)
        sr = rng.uniform(0.5, 2.5, size=6)
        codes, leaf_nums = simulate_codes(t, mcp.GTR(pi, sr), pi, np.ones(1), 500, rng)
        aln = mcp.DeviceAlignment(codes, leaf_nums, K)
        ll_o, g_o = _oracle_eval(oracle, t, codes, leaf_nums, K, model, pi, sr, [1.0])
        ll, g = mcp.gradlogpdf(mcp.PhyloDist(t, pi, sr, [1.0], model), aln)
        _check(ll, g, ll_o, g_o)
        trees.append(t); alns.append(aln); expect.append((ll_o, g_o)); pis.append(pi); srs.append(sr)
    # the batch entry point with per-tree models
    from mcphylo_jl_b200.phylodist import _device_alignment
    ctx = mcp.get_context()
    handles, flat = [], []
    for t, aln, m, p_, s_ in zip(trees, alns, models, pis, srs):
        ft = mcp.flatten(t)
        U, D, Uinv, mu = m(p_, s_)
        handles.append(_device_alignment(aln, ft.leaf_nums, K, ctx))
        flat.append((ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, np.ones(1), p_))
    lls, grads = ctx.eval_batch(handles, flat, want_grad=True)
    for ll, g, (ll_o, g_o) in zip(lls, grads, expect):
        _check(ll, g, ll_o, g_o)


def test_topology_cache_and_branch_updates(oracle):
    """Same topology, new branch lengths (the leapfrog pattern), then an NNI."""
    rng = np.random.default_rng(21)
    tree = random_tree(30, rng)
    pi = np.array([0.3, 0.7])
    codes, leaf_nums = simulate_codes(tree, mcp.Restriction(pi, []), pi, np.ones(1), 500, rng)
    aln = mcp.DeviceAlignment(codes, leaf_nums, 2)
    ctx = mcp.get_context()
    for it in range(3):
        blv = np.clip(rng.exponential(0.1, size=58), 1e-4, 1.0)
        mcp.set_branchlength_vector(tree, blv)
        pd = mcp.PhyloDist(tree, pi, [0.0], [1.0], mcp.Restriction)
        ll, g = mcp.gradlogpdf(pd, aln)
        assert ctx.stats()["schedule_rebuilt"] == (1 if it == 0 else 0)
        _check(ll, g, *_oracle_eval(oracle, tree, codes, leaf_nums, 2, mcp.Restriction, pi, [0.0], [1.0]))
    target = next(n for n in mcp.post_order(tree) if n.nchild == 2 and not n.root and n.mother.nchild == 2)
    assert mcp.NNI(tree, target) == 1
    pd = mcp.PhyloDist(tree, pi, [0.0], [1.0], mcp.Restriction)
    ll, g = mcp.gradlogpdf(pd, aln)
    assert ctx.stats()["schedule_rebuilt"] == 1
    _check(ll, g, *_oracle_eval(oracle, tree, codes, leaf_nums, 2, mcp.Restriction, pi, [0.0], [1.0]))


def test_batch_of_distinct_trees(oracle):
    """MultiplePhyloDist shape of BASELINE config 5, scaled down: distinct topologies, own data."""
    rng = np.random.default_rng(5000)
    pi = np.array([0.3, 0.7])
    trees, alns, expect = [], [], []
    for i in range(7):
        t = random_tree(12 + 3 * i, rng)
        codes, leaf_nums = simulate_codes(t, mcp.Restriction(pi, []), pi, np.ones(1), 700, rng)
        trees.append(t)
        alns.append(mcp.DeviceAlignment(codes, leaf_nums, 2))
        expect.append(_oracle_eval(oracle, t, codes, leaf_nums, 2, mcp.Restriction, pi, [0.0], [1.0]))
    mpd = mcp.MultiplePhyloDist(trees, pi, [0.0], [1.0], mcp.Restriction)
    res = mcp.multi_gradlogpdf(mpd, alns)
    for (ll, g), (ll_o, g_o) in zip(res, expect):
        _check(ll, g, ll_o, g_o)
    total = mcp.logpdf(mpd, alns)
    assert abs(total - sum(e[0] for e in expect)) <= LL_RTOL * abs(total)


def test_finite_differences_on_device():
    rng = np.random.default_rng(77)
    tree = random_tree(9, rng)
    pi = np.array([0.1, 0.2, 0.3, 0.4])
    sr = np.array([1.0, 2.0, 1.5, 0.8, 2.5, 1.2])
    rates = mcp.discrete_gamma_rates(0.5, 0.5, 4)
    codes, leaf_nums = simulate_codes(tree, mcp.GTR(pi, sr), pi, rates, 200, rng)
    aln = mcp.DeviceAlignment(codes, leaf_nums, 4)
    ll, g = mcp.gradlogpdf(mcp.PhyloDist(tree, pi, sr, rates, mcp.GTR), aln)
    blv = mcp.get_branchlength_vector(tree)
    for b in range(blv.size):
        vals = []
        for sgn in (1, -1):
            t = blv.copy()
            t[b] += sgn * 1e-6
            mcp.set_branchlength_vector(tree, t)
            vals.append(mcp.logpdf(mcp.PhyloDist(tree, pi, sr, rates, mcp.GTR), aln))
        mcp.set_branchlength_vector(tree, blv)
        assert abs((vals[0] - vals[1]) / 2e-6 - g[b]) <= 2e-6 * max(1.0, abs(g[b]))


@pytest.mark.parametrize("block,cpt,ctas", [(256, 2, 0), (256, 1, 3), (128, 2, 0)])
def test_cfg4_tree_oracle_window_at_bench_launch_shapes(oracle, block, cpt, ctas):
    """BASELINE config 4's own tree (1000 taxa, bench.py seeds, GTR + Gamma-4) on a 512-site window of its
    alignment, against the oracle, with the launch shapes the timed runs use forced: two columns per thread
    on 256-thread CTAs (every GPU count since round 2), round 1's 8-GPU shape (one column per thread, 3 CTAs
    per SM, grid 444), and a narrower tile.  The dense oracle needs 3 x 130 MB for this."""
    import bench
    w = bench.make_workload("cfg4", 512)
    codes, leaf_nums = bench.make_codes(w, 0, 512)
    assert codes.shape == (1000, 512)
    ctx = mcp.get_context()
    ctx.set_launch(block, ctas)
    ctx.set_columns_per_thread(cpt)
    ctx.set_level_mode(0)
    ctx.set_scratch_mode(0)
    try:
        pd = mcp.PhyloDist(w["tree"], w["pi"], w["srates"], w["rates"], w["model"])
        aln = mcp.DeviceAlignment(codes, leaf_nums, 4)
        ll, g = mcp.gradlogpdf(pd, aln)
        st = ctx.stats()
        assert st["block"] == block and st["columns_per_thread"] == cpt
        ll_only = mcp.logpdf(pd, aln)
    finally:
        ctx.set_launch(0, 0)
        ctx.set_columns_per_thread(0)
        ctx.set_level_mode(-1)
        ctx.set_scratch_mode(-1)
    ll_o, g_o = _oracle_eval(oracle, w["tree"], codes, leaf_nums, 4, w["model"], w["pi"], w["srates"], w["rates"])
    assert g.shape == (1998,)
    _check(ll, g, ll_o, g_o)
    _check(ll_only, None, ll_o, None)


def test_cfg5_shape_batch_oracle_window(oracle):
    """BASELINE config 5's shape -- 256 independent 100-taxon trees (bench.py seeds 5000 + i), binary sites,
    ONE mcp_eval_batch launch -- on a 256-site window per tree, every tree against the oracle
    (the MultiplePhyloDist path, /root/reference/src/distributions/Phylodist.jl:281-297)."""
    from mcphylo_jl_b200 import capi
    from mcphylo_jl_b200.phylodist import _tree_args
    import bench
    T, n_taxa, S = 256, 100, 256
    pi = bench.RESTRICTION_PI
    model_out = mcp.Restriction(pi, [])
    ctx = capi.Context(0)
    try:
        alns, targs, cases = [], [], []
        for i in range(T):
            tree = mcp.random_tree(n_taxa, np.random.default_rng(5000 + i))
            rng = np.random.default_rng(1005 * 1000 + i)
            codes, leaf_nums = mcp.simulate_codes(tree, model_out, pi, np.ones(1), S, rng, gap_frac=0.01)
            alns.append(ctx.alignment_from_codes(codes, 2, leaf_nums))
            targs.append(_tree_args(mcp.PhyloDist(tree, pi, [0.0], [1.0], mcp.Restriction))[1])
            cases.append((tree, codes, leaf_nums))
        prep = capi.PreparedBatch(ctx, alns, targs, want_grad=True)
        ll, grads = prep.eval()
        ll, grads = ll.copy(), [g.copy() for g in grads]                   # eval() returns views of reused buffers
        st = ctx.stats()
        assert st["kernel_launches"] == 3 + 1 + st["schedule_rebuilt"]      # all 256 trees in one walk launch
        for i, (tree, codes, leaf_nums) in enumerate(cases):
            ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, 2, mcp.Restriction, pi, [0.0], [1.0])
            _check(ll[i], grads[i], ll_o, g_o)
        # logL only, and a second call with new branch lengths on the cached plan
        for t, (tree, _, _) in enumerate(cases):
            prep.set_blv(t, mcp.get_branchlength_vector(tree) * 1.1)
        ll2, _ = prep.eval()
        assert ctx.stats()["schedule_rebuilt"] == 0 and np.all(ll2 != ll)
        tree, codes, leaf_nums = cases[17]
        mcp.set_branchlength_vector(tree, mcp.get_branchlength_vector(tree) * 1.1)
        ll_o, _ = _oracle_eval(oracle, tree, codes, leaf_nums, 2, mcp.Restriction, pi, [0.0], [1.0])
        _check(ll2[17], None, ll_o, None)
        for a in alns:
            a.close()
    finally:
        ctx.close()


def test_nni_pair_in_one_batched_launch(oracle):
    """The PNUTS caller pattern (/root/reference/src/samplers/tree_hamiltonian/refraction.jl:2-31,62-69):
    a gradient on the current tree, then the log-likelihoods of the tree before and after an NNI.  The two
    logpdf calls of an NNI attempt go out as ONE mcp_eval_batch launch (T = 2, same alignment), and the
    sequence gradient / pair / gradient on the accepted tree matches the oracle at cfg2 shape."""
    import copy
    from mcphylo_jl_b200 import capi
    from mcphylo_jl_b200.phylodist import _tree_args
    rng = np.random.default_rng(20242)
    tree = random_tree(50, rng)
    pi = np.array([0.3, 0.7])
    codes, leaf_nums = simulate_codes(tree, mcp.Restriction(pi, []), pi, np.ones(1), 10000, rng, gap_frac=0.01)
    ctx = capi.Context(0)
    try:
        aln = ctx.alignment_from_codes(codes, 2, leaf_nums)
        inner = [n for n in mcp.post_order(tree) if n.nchild == 2 and not n.root]
        for attempt in range(3):
            ta = _tree_args(mcp.PhyloDist(tree, pi, [0.0], [1.0], mcp.Restriction))[1]
            ll, g = ctx.eval(aln, *ta, want_grad=True)
            _check(ll, g, *_oracle_eval(oracle, tree, codes, leaf_nums, 2, mcp.Restriction, pi, [0.0], [1.0]))
            proposal = copy.deepcopy(tree)
            target = [n for n in mcp.post_order(proposal) if n.num == inner[(7 * attempt + 3) % len(inner)].num][0]
            assert mcp.NNI(proposal, target, lor=bool(attempt % 2)) == 1
            tb = _tree_args(mcp.PhyloDist(proposal, pi, [0.0], [1.0], mcp.Restriction))[1]
            pair, none = ctx.eval_batch([aln, aln], [ta, tb], want_grad=False)
            assert none is None and ctx.stats()["kernel_launches"] <= 4 + ctx.stats()["schedule_rebuilt"]   # one call, one walk
            assert pair[0] == ll or abs(pair[0] - ll) <= 1e-12 * abs(ll)
            ll_b, _ = _oracle_eval(oracle, proposal, codes, leaf_nums, 2, mcp.Restriction, pi, [0.0], [1.0], want_grad=False)
            _check(pair[1], None, ll_b, None)
            tree = proposal                                   # "accept": the next gradient runs on the new topology
            inner = [n for n in mcp.post_order(tree) if n.nchild == 2 and not n.root]
        aln.close()
    finally:
        ctx.close()


def test_large_properties_without_oracle():
    """1000 taxa x 20 000 sites (more than the dense oracle is asked to hold in a test; the oracle-window
    tests above cover this tree size): additivity over site blocks (the multi-GPU sharding identity),
    rate-category additivity, and the logL-only path."""
    rng = np.random.default_rng(20244)
    n_taxa, S = 1000, 20000
    tree = random_tree(n_taxa, rng)
    pi = np.array([0.1, 0.2, 0.3, 0.4])
    sr = np.array([1.0, 2.0, 1.5, 0.8, 2.5, 1.2])
    rates = mcp.discrete_gamma_rates(0.5, 0.5, 4)
    codes, leaf_nums = simulate_codes(tree, mcp.GTR(pi, sr), pi, rates, S, rng)
    pd = mcp.PhyloDist(tree, pi, sr, rates, mcp.GTR)
    full = mcp.DeviceAlignment(codes, leaf_nums, 4)
    ll, g = mcp.gradlogpdf(pd, full)
    assert np.isfinite(ll) and np.all(np.isfinite(g)) and g.shape == (2 * n_taxa - 2,)
    # site blocks add up
    parts = [mcp.gradlogpdf(pd, full.site_block(lo, hi)) for lo, hi in [(0, 7001), (7001, 7002), (7002, S)]]
    assert abs(sum(p[0] for p in parts) - ll) <= 1e-12 * abs(ll)
    gs = sum(p[1] for p in parts)
    assert np.all(np.abs(gs - g) <= 1e-10 * np.maximum(np.abs(g), 1e-3 * np.max(np.abs(g))))
    # rate categories are independent replicas (reference semantics: no mixing)
    per_rate = [mcp.gradlogpdf(mcp.PhyloDist(tree, pi, sr, [r], mcp.GTR), full) for r in rates]
    assert abs(sum(p[0] for p in per_rate) - ll) <= 1e-12 * abs(ll)
    # logL-only path (recycled slots) equals the gradient path's logL
    assert abs(mcp.logpdf(pd, full) - ll) <= 1e-13 * abs(ll)


def test_error_paths():
    tree, x, codes, leaf_nums, fx = golden_case("simudata")
    pd = mcp.PhyloDist(tree, fx["base_freq"], [1.0], [1.0], mcp.JC)
    bad = x.copy(order="F")
    bad[:, 3, 0] = 0.5
    with pytest.raises(mcp.capi.McpError) as ei:
        mcp.logpdf(pd, bad)
    assert ei.value.code == -4
    with pytest.raises(mcp.DimensionMismatch):
        mcp.logpdf(mcp.PhyloDist(tree, [0.5, 0.5], [1.0], [1.0], mcp.Restriction), x)
    # an alignment that lacks one of the tree's leaves
    with pytest.raises(mcp.capi.McpError) as ei:
        mcp.logpdf(pd, mcp.DeviceAlignment(codes[:-1], leaf_nums[:-1], 4))
    assert ei.value.code == -1
    # 40 states: beyond the generic kernel's limit, reported not crashed
    t2 = mcp.ParseNewick("(a:0.1,b:0.2);")
    with pytest.raises(mcp.capi.McpError) as ei:
        mcp.logpdf(mcp.PhyloDist(t2, np.full(40, 0.025), [1.0], [1.0], mcp.JC),
                   mcp.DeviceAlignment(np.zeros((2, 5), np.uint8), [1, 2], 40))
    assert ei.value.code == -3


def test_sharded_evaluator_follows_torch_streams(oracle):
    """The multi-GPU entry (mcp_eval_device on torch's current stream, result left on the device)
    at world size 1, on the default stream and on a side stream."""
    import torch
    rng = np.random.default_rng(31)
    tree = random_tree(60, rng)
    pi = np.array([0.1, 0.2, 0.3, 0.4])
    sr = np.array([1.0, 2.0, 1.5, 0.8, 2.5, 1.2])
    rates = mcp.discrete_gamma_rates(0.5, 0.5, 4)
    codes, leaf_nums = simulate_codes(tree, mcp.GTR(pi, sr), pi, rates, 5000, rng)
    pd = mcp.PhyloDist(tree, pi, sr, rates, mcp.GTR)
    aln = mcp.DeviceAlignment(codes, leaf_nums, 4)
    ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, 4, mcp.GTR, pi, sr, rates)
    ev = mcp.ShardedEvaluator(aln, 0)
    try:
        ll, g = ev.gradlogpdf(pd)
        _check(ll, g, ll_o, g_o)
        assert ev.ctx.stats()["walk_ms"] > 0
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for _ in range(3):
                ll2, g2 = ev.gradlogpdf(pd)
        assert ll2 == ll and np.array_equal(g2, g)
        assert abs(ev.logpdf(pd) - ll_o) <= LL_RTOL * abs(ll_o)
    finally:
        ev.ctx.set_stream(None)
    # shards of one alignment add up to the whole (what the all-reduce sums)
    parts = [mcp.gradlogpdf(pd, mcp.local_shard(aln, 4, r)) for r in range(4)]
    assert abs(sum(p[0] for p in parts) - ll) <= 1e-12 * abs(ll)


@pytest.mark.parametrize("K", [2, 4])
def test_extreme_branch_lengths(oracle, K):
    """Near-zero and saturated branches (no clamping anywhere, like the reference): P ~ I and
    P ~ stationary both have to survive the eigen-factored form and the power-of-two rescaling."""
    rng = np.random.default_rng(300 + K)
    tree = random_tree(40, rng)
    pi = rng.dirichlet(np.ones(K) * 5)
    model, pi, srates = _model(K, pi, rng)
    rates = mcp.discrete_gamma_rates(0.5, 0.5, 4) if K == 4 else np.ones(1)
    codes, leaf_nums = simulate_codes(tree, model(pi, srates), pi, rates, 400, rng, gap_frac=0.02)
    blv = mcp.get_branchlength_vector(tree)
    idx = rng.permutation(blv.size)
    blv[idx[:6]] = [1e-9, 1e-7, 1e-5, 30.0, 80.0, 5.0]
    mcp.set_branchlength_vector(tree, blv)
    pd = mcp.PhyloDist(tree, pi, srates, rates, model)
    aln = mcp.DeviceAlignment(codes, leaf_nums, K)
    ll, g = mcp.gradlogpdf(pd, aln)
    ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, K, model, pi, srates, rates)
    assert np.isfinite(ll) and np.all(np.isfinite(g))
    assert abs(ll - ll_o) <= LL_RTOL * abs(ll_o)
    scale = np.max(np.abs(g_o))
    assert np.all(np.abs(g - g_o) <= 1e-7 * np.maximum(np.abs(g_o), 1e-3 * scale))


def test_long_chain_does_not_underflow(oracle):
    """A 1200-taxon caterpillar with long branches: per-column likelihoods below 1e-308, which an
    unscaled double cannot hold; the exponent bookkeeping must carry them (the reference rescales
    at every node too)."""
    n = 1200
    nwk = "(" * (n - 1) + "t0000:0.4," + ",".join(f"t{i:04d}:0.4):0.3" for i in range(1, n))
    nwk = nwk[:nwk.rfind(":")] + ";"
    tree = mcp.ParseNewick(nwk)
    rng = np.random.default_rng(9)
    pi = np.array([0.1, 0.2, 0.3, 0.4])
    sr = np.array([1.0, 2.0, 1.5, 0.8, 2.5, 1.2])
    codes, leaf_nums = simulate_codes(tree, mcp.GTR(pi, sr), pi, np.ones(1), 32, rng, gap_frac=0.0)
    pd = mcp.PhyloDist(tree, pi, sr, [1.0], mcp.GTR)
    ll, g = mcp.gradlogpdf(pd, mcp.DeviceAlignment(codes, leaf_nums, 4))
    ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, 4, mcp.GTR, pi, sr, [1.0])
    assert ll_o / 32 < -720        # exp(-720) < smallest normal double
    _check(ll, g, ll_o, g_o)


@pytest.mark.parametrize("K,R", [(2, 1), (4, 4), (3, 2)])
def test_scratch_in_shared_memory_and_in_hbm_agree(oracle, K, R):
    """Small inputs keep the partials scratch in shared memory (latency path); forcing it to HBM
    must give the same numbers, and both must match the oracle."""
    rng = np.random.default_rng(400 + K)
    tree = random_tree(50, rng, multifurcate=True)
    pi = rng.dirichlet(np.ones(K) * 5)
    model, pi, srates = _model(K, pi, rng)
    rates = mcp.discrete_gamma_rates(0.5, 0.5, R) if R > 1 else np.ones(1)
    codes, leaf_nums = simulate_codes(tree, model(pi, srates), pi, rates, 1000, rng, gap_frac=0.03)
    pd = mcp.PhyloDist(tree, pi, srates, rates, model)
    aln = mcp.DeviceAlignment(codes, leaf_nums, K)
    ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, K, model, pi, srates, rates)
    ctx = mcp.get_context()
    res = {}
    try:
        for mode in (0, 1):
            ctx.set_scratch_mode(mode)
            for block in (32, 64):
                ctx.set_launch(block, 0)
                ll, g = mcp.gradlogpdf(pd, aln)
                _check(ll, g, ll_o, g_o)
                _check(mcp.logpdf(pd, aln), None, ll_o, None)
                res[(mode, block)] = (ll, g)
    finally:
        ctx.set_scratch_mode(-1)
        ctx.set_launch(0, 0)
    assert res[(0, 32)][0] == res[(1, 32)][0]          # identical arithmetic, identical logL


def test_empty_and_single_site_alignments(oracle):
    """S = 0 (the reference returns 0.0: empty sums) and S = 1."""
    tree = mcp.ParseNewick("((a:0.1,b:0.2)e:0.05,(c:0.3,d:0.1)f:0.2)g;")
    pd = mcp.PhyloDist(tree, [0.25] * 4, [1.0], [1.0, 2.0], mcp.JC)
    leaf_nums = np.array([1, 2, 3, 4], dtype=np.int32)
    ll, g = mcp.gradlogpdf(pd, mcp.DeviceAlignment(np.zeros((4, 0), np.uint8), leaf_nums, 4))
    assert ll == 0.0 and np.all(g == 0.0) and g.shape == (6,)
    codes = np.array([[0], [1], [4], [3]], dtype=np.uint8)
    ll, g = mcp.gradlogpdf(pd, mcp.DeviceAlignment(codes, leaf_nums, 4))
    ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, 4, mcp.JC, [0.25] * 4, [1.0], [1.0, 2.0])
    _check(ll, g, ll_o, g_o)


def test_config3_full_size_properties(oracle):
    """BASELINE config 3 at full size (200 taxa x 100 000 sites, GTR + Gamma4): too large for the
    dense oracle as a whole, so (i) a contiguous 400-site window is checked against the oracle and
    (ii) the full result must equal the sum over an uneven partition into site blocks, which also
    exercises different tile counts / launch shapes of the same data."""
    rng = np.random.default_rng(20243)
    tree = random_tree(200, rng)
    pi = np.array([0.1, 0.2, 0.3, 0.4])
    sr = np.array([1.0, 2.0, 1.5, 0.8, 2.5, 1.2])
    rates = mcp.discrete_gamma_rates(0.5, 0.5, 4)
    S = 100_000
    pool, leaf_nums = simulate_codes(tree, mcp.GTR(pi, sr), pi, rates, 8192, rng)
    codes = np.take(pool, rng.integers(0, 8192, size=S), axis=1)
    pd = mcp.PhyloDist(tree, pi, sr, rates, mcp.GTR)
    full = mcp.DeviceAlignment(codes, leaf_nums, 4)
    ll, g = mcp.gradlogpdf(pd, full)
    cuts = [0, 400, 33_333, 33_334, 77_001, S]
    parts = [mcp.gradlogpdf(pd, full.site_block(a, b)) for a, b in zip(cuts[:-1], cuts[1:])]
    assert abs(sum(p[0] for p in parts) - ll) <= 1e-12 * abs(ll)
    gs = sum(p[1] for p in parts)
    assert np.all(np.abs(gs - g) <= 1e-10 * np.maximum(np.abs(g), 1e-3 * np.max(np.abs(g))))
    ll_o, g_o = _oracle_eval(oracle, tree, codes[:, :400], leaf_nums, 4, mcp.GTR, pi, sr, rates)
    _check(parts[0][0], parts[0][1], ll_o, g_o)
    assert abs(mcp.logpdf(pd, full) - ll) <= 1e-13 * abs(ll)


def test_pipelined_evaluator_matches_resident(oracle):
    """Upload-overlapped evaluation in site blocks (dist.PipelinedEvaluator) vs the resident path."""
    rng = np.random.default_rng(41)
    tree = random_tree(30, rng)
    pi = np.array([0.1, 0.2, 0.3, 0.4])
    sr = np.array([1.0, 2.0, 1.5, 0.8, 2.5, 1.2])
    rates = mcp.discrete_gamma_rates(0.5, 0.5, 4)
    codes, leaf_nums = simulate_codes(tree, mcp.GTR(pi, sr), pi, rates, 3001, rng)
    pd = mcp.PhyloDist(tree, pi, sr, rates, mcp.GTR)
    ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, 4, mcp.GTR, pi, sr, rates)
    for n_blocks in (1, 3, 4):
        pipe = mcp.PipelinedEvaluator(codes, leaf_nums, 4, 0, n_blocks=n_blocks)
        try:
            for _ in range(2):                      # second call re-uploads into the same buffers
                ll, g = pipe.gradlogpdf(pd)
                _check(ll, g, ll_o, g_o)
            assert abs(pipe.logpdf(pd) - ll_o) <= LL_RTOL * abs(ll_o)
        finally:
            pipe.close()


@pytest.mark.parametrize("n_taxa,K,R,S,multi", [(10, 2, 1, 1000, False), (50, 2, 1, 3000, True), (40, 4, 4, 500, True),
                                                (2, 4, 1, 40, False), (60, 3, 2, 300, False), (25, 6, 1, 100, False)])
def test_level_parallel_kernel(oracle, n_taxa, K, R, S, multi):
    """The small-tree kernel (warps of a CTA split the ops of a tree level) forced on, against the
    oracle and against the depth-first kernel on the same input."""
    rng = np.random.default_rng(500 + n_taxa + K)
    tree = random_tree(n_taxa, rng, multifurcate=multi, unary=(K == 3))
    pi = rng.dirichlet(np.ones(K) * 5)
    model, pi, srates = _model(K, pi, rng)
    rates = mcp.discrete_gamma_rates(0.5, 0.5, R) if R > 1 else np.ones(1)
    codes, leaf_nums = simulate_codes(tree, model(pi, srates), pi, rates, S, rng, gap_frac=0.05)
    pd = mcp.PhyloDist(tree, pi, srates, rates, model)
    aln = mcp.DeviceAlignment(codes, leaf_nums, K)
    ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, K, model, pi, srates, rates)
    ctx = mcp.get_context()
    try:
        ctx.set_level_mode(1)
        ll1, g1 = mcp.gradlogpdf(pd, aln)
        assert ctx.stats()["block"] == 256
        l1 = mcp.logpdf(pd, aln)
        ctx.set_level_mode(0)
        ll0, g0 = mcp.gradlogpdf(pd, aln)
    finally:
        ctx.set_level_mode(-1)
    _check(ll1, g1, ll_o, g_o)
    _check(l1, None, ll_o, None)
    _check(ll0, g0, ll_o, g_o)


@pytest.mark.parametrize("n_taxa,K,R,S", [(50, 2, 1, 10000),      # cfg2's shape: 313 CTAs -> 40 groups, the last one of 1 CTA
                                          (30, 4, 2, 3300),       # 208 CTAs -> 26 groups, two rate categories
                                          (20, 2, 1, 6176)])      # 193 CTAs: the smallest launch that takes two levels
def test_level_kernel_two_level_final_reduction(oracle, n_taxa, K, R, S):
    """Launches of more than 192 CTAs of the fused small-tree kernel reduce their accumulator rows in two levels (groups
    of 8 CTAs, then the group rows): logL, gradient and the prior epilogue against the oracle, bit-reproducible."""
    rng = np.random.default_rng(900 + n_taxa + K)
    tree = random_tree(n_taxa, rng)
    pi = rng.dirichlet(np.ones(K) * 5)
    model, pi, srates = _model(K, pi, rng)
    rates = mcp.discrete_gamma_rates(0.5, 0.5, R) if R > 1 else np.ones(1)
    codes, leaf_nums = simulate_codes(tree, model(pi, srates), pi, rates, S, rng, gap_frac=0.02)
    pd = mcp.PhyloDist(tree, pi, srates, rates, model)
    aln = mcp.DeviceAlignment(codes, leaf_nums, K)
    ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, K, model, pi, srates, rates)
    prior = mcp.CompoundDirichlet(1.3, 0.9, 0.2, 1.4)
    vp, gp = _oracle_prior(oracle, prior, tree)
    ctx = mcp.get_context()
    try:
        ctx.set_level_mode(1)
        ll1, g1 = mcp.gradlogpdf(pd, aln)
        st = ctx.stats()
        assert st["block"] == 256 and st["grid"] > 192 and st["kernel_launches"] == 1 + st["schedule_rebuilt"]
        ll2, g2 = mcp.gradlogpdf(pd, aln)
        l1 = mcp.logpdf(pd, aln)
        lp, gl = mcp.logpdfgrad(pd, aln, prior)
        ll3, g3 = mcp.gradlogpdf(pd, aln)                  # the group counters were left at zero by every launch
    finally:
        ctx.set_level_mode(-1)
    _check(ll1, g1, ll_o, g_o)
    assert ll2 == ll1 and np.array_equal(g1, g2) and ll3 == ll1 and np.array_equal(g1, g3)
    _check(l1, None, ll_o, None)
    _check(lp, gl, ll_o + vp, g_o + gp)


# ---------------------------------------------------------------------------------------------
# likelihood + branch-length prior in one device call (mcp_eval_posterior; SURVEY.md §8f row 4,
# the body of logpdfgrad!(::Type{provided}), /root/reference/src/samplers/sampler.jl:172-190)
# ---------------------------------------------------------------------------------------------
def _oracle_prior(oracle, prior, tree):
    blv, ie = mcp.get_branchlength_vector(tree), mcp.internal_external(tree)
    if prior is None or isinstance(prior, mcp.UniformBranchLength):
        return 0.0, np.zeros(blv.size)
    if isinstance(prior, mcp.exponentialBL):
        return oracle.exponential_bl_gradlogpdf(prior.scale, blv)
    return oracle.compound_dirichlet_gradlogpdf(prior.alpha, prior.a, prior.beta, prior.c, blv, ie)


@pytest.mark.parametrize("prior", [
    mcp.CompoundDirichlet(1.0, 1.0, 0.100, 1.0),      # the reference's test prior (tree_samplers.jl:22)
    mcp.CompoundDirichlet(2.5, 0.7, 0.3, 1.9),
    mcp.exponentialBL(0.25),
    mcp.UniformBranchLength(),
    None,
])
def test_posterior_golden_tree(oracle, prior):
    tree, x, codes, leaf_nums, fx = golden_case("simudata")
    pd = mcp.PhyloDist(tree, fx["base_freq"], [1.0], [1.0], mcp.JC)
    lp, grad = mcp.logpdfgrad(pd, x, prior)
    ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, 4, mcp.JC, fx["base_freq"], [1.0], [1.0])
    vp, gp = _oracle_prior(oracle, prior, tree)
    _check(lp, grad, ll_o + vp, g_o + gp)
    lp_only, none = mcp.logpdfgrad(pd, x, prior, want_grad=False)
    assert none is None and abs(lp_only - (ll_o + vp)) <= LL_RTOL * abs(ll_o + vp)
    # the sampler's sum, piece by piece: likelihood call + prior mirror
    ll, g = mcp.gradlogpdf(pd, x)
    if prior is not None:
        v_h, g_h = mcp.gradlogpdf(prior, tree)
        assert abs(lp - (ll + v_h)) <= 1e-12 * abs(lp)
        assert np.allclose(grad, g + g_h, rtol=1e-11, atol=1e-9)


@pytest.mark.parametrize("level_mode", [0, 1])
def test_posterior_random_tree_both_kernels(oracle, level_mode):
    """Same prior epilogue in the three-launch path (finalize_results) and in the fused small-tree kernel."""
    rng = np.random.default_rng(77)
    tree = random_tree(40, rng)
    pi = rng.dirichlet(np.ones(4) * 5)
    srates = rng.uniform(0.5, 2.5, size=6)
    rates = mcp.discrete_gamma_rates(0.5, 0.5, 4)
    codes, leaf_nums = simulate_codes(tree, mcp.GTR(pi, srates), pi, rates, 500, rng, gap_frac=0.02)
    pd = mcp.PhyloDist(tree, pi, srates, rates, mcp.GTR)
    aln = mcp.DeviceAlignment(codes, leaf_nums, 4)
    ctx = mcp.get_context()
    ctx.set_level_mode(level_mode)
    try:
        prior = mcp.CompoundDirichlet(1.3, 0.9, 0.2, 1.4)
        lp, grad = mcp.logpdfgrad(pd, aln, prior)
        st = ctx.stats()
        # the ONE fused kernel with the parameters in its arguments, or: the kernel that stages the parameters from
        # pinned host memory + tables + walk + final reduction (and, the first time a topology is seen, the kernel
        # that stages its program)
        assert st["kernel_launches"] == (1 if level_mode else 4) + st["schedule_rebuilt"]
    finally:
        ctx.set_level_mode(-1)
    ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, 4, mcp.GTR, pi, srates, rates)
    vp, gp = _oracle_prior(oracle, prior, tree)
    _check(lp, grad, ll_o + vp, g_o + gp)


def test_reupload_is_ordered_against_evaluations(oracle):
    """mcp_alignment_update_codes runs on the copy stream: an evaluation enqueued BEFORE a re-upload
    must still see the old codes, one enqueued AFTER it the new ones, without any host
    synchronisation in between (per-alignment events)."""
    import torch

    from mcphylo_jl_b200 import capi

    rng = np.random.default_rng(9)
    tree = random_tree(120, rng)
    pi = np.array([0.1, 0.2, 0.3, 0.4])
    sr = np.array([1.0, 2.0, 1.5, 0.8, 2.5, 1.2])
    rates = mcp.discrete_gamma_rates(0.5, 0.5, 4)
    S = 60000                                        # big enough for the evaluation to outlast the enqueue
    codes1, leaf_nums = simulate_codes(tree, mcp.GTR(pi, sr), pi, rates, S, rng)
    codes2, _ = simulate_codes(tree, mcp.GTR(pi, sr), pi, rates, S, rng)
    assert not np.array_equal(codes1, codes2)
    ft = mcp.flatten(tree)
    U, D, Uinv, mu = mcp.GTR(pi, sr)
    targs = (ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, rates, pi)
    ctx = capi.Context(0)
    try:
        wave = ctx.wave_columns(4, ft.NN, True)
        assert wave > 0 and wave % 32 == 0
        h1 = torch.from_numpy(codes1).pin_memory()
        h2 = torch.from_numpy(codes2).pin_memory()
        aln = ctx.alignment_from_codes(codes1, 4, leaf_nums)
        ref1 = ctx.eval(aln, *targs, want_grad=True)
        out = torch.zeros((3, ft.NN), dtype=torch.float64, device="cuda:0")
        # eval(old) | upload(new) | eval(new) | upload(old) | eval(old), all enqueued back to back
        ctx.eval_device(aln, *targs, want_grad=True, d_out_ptr=out[0].data_ptr())
        aln.update_codes(h2.data_ptr())
        ctx.eval_device(aln, *targs, want_grad=True, d_out_ptr=out[1].data_ptr())
        aln.update_codes(h1.data_ptr())
        ctx.eval_device(aln, *targs, want_grad=True, d_out_ptr=out[2].data_ptr())
        ctx.synchronize()
        res = out.cpu().numpy()
        aln2 = ctx.alignment_from_codes(codes2, 4, leaf_nums)
        ref2 = ctx.eval(aln2, *targs, want_grad=True)
        assert res[0, 0] == ref1[0] and res[2, 0] == ref1[0]          # logL sums are deterministic
        assert res[1, 0] == ref2[0] and ref1[0] != ref2[0]
        assert np.allclose(res[0, 1:], ref1[1], rtol=1e-12) and np.allclose(res[1, 1:], ref2[1], rtol=1e-12)
        aln.close()
        aln2.close()
    finally:
        ctx.close()


@pytest.mark.parametrize("K,R,S", [(2, 1, 5000), (4, 4, 700)])
def test_rerooting_invariance_on_device(K, R, S):
    """Pulley principle on the CUDA path (no oracle involved): shifting length across the root and
    unrooting the tree keep logL; the two root-branch gradients are equal."""
    from synth import reroot_variants

    rng = np.random.default_rng(310 + K)
    tree = random_tree(33, rng)
    pi = rng.dirichlet(np.ones(K) * 5)
    model, srates = (mcp.Restriction, np.zeros(1)) if K == 2 else (mcp.GTR, rng.uniform(0.5, 2.5, size=6))
    rates = mcp.discrete_gamma_rates(0.5, 0.5, R) if R > 1 else np.ones(1)
    codes, leaf_nums = simulate_codes(tree, model(pi, srates), pi, rates, S, rng, gap_frac=0.05)
    aln = mcp.DeviceAlignment(codes, leaf_nums, K)

    def ev(t):
        return mcp.gradlogpdf(mcp.PhyloDist(t, pi, srates, rates, model), aln)

    ll, g = ev(tree)
    a, b = tree.children
    assert abs(g[a.num - 1] - g[b.num - 1]) <= 1e-9 * max(abs(g[a.num - 1]), 1e-3 * np.max(np.abs(g)))
    shifted, unrooted = reroot_variants(tree)
    assert abs(ev(shifted)[0] - ll) <= 1e-12 * abs(ll)
    assert unrooted is not None and abs(ev(unrooted)[0] - ll) <= 1e-12 * abs(ll)


def test_zero_length_branches_and_impossible_columns(oracle):
    """No clamping (SURVEY.md §8a edge behaviour): a branch of length exactly 0 is evaluated as
    P = I; a column that is impossible under the tree (two different states joined by zero-length
    branches) gives logL = -Inf and non-finite gradient entries, as documented in DESIGN.md §1."""
    tree = mcp.ParseNewick("(((a:0.0,b:0.0)ab:0.1,c:0.2)abc:0.05,(d:0.3,e:0.0)de:0.1);")
    pi = np.array([0.1, 0.2, 0.3, 0.4])
    sr = np.array([1.0, 2.0, 1.5, 0.8, 2.5, 1.2])
    rates = mcp.discrete_gamma_rates(0.5, 0.5, 4)
    leaves = mcp.get_leaves(tree)
    leaf_nums = np.array([n.num for n in leaves], dtype=np.int32)
    names = [n.name for n in leaves]
    rng = np.random.default_rng(5)
    S = 64
    codes = rng.integers(0, 4, size=(len(leaves), S)).astype(np.uint8)
    codes[names.index("b")] = codes[names.index("a")]          # a == b everywhere: every column is possible
    pd = mcp.PhyloDist(tree, pi, sr, rates, mcp.GTR)
    ll, g = mcp.gradlogpdf(pd, mcp.DeviceAlignment(codes, leaf_nums, 4))
    ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, 4, mcp.GTR, pi, sr, rates)
    assert np.isfinite(ll) and np.all(np.isfinite(g))
    _check(ll, g, ll_o, g_o)
    bad = codes.copy()
    bad[names.index("b"), 7] = (bad[names.index("a"), 7] + 1) % 4   # a != b at one site, joined by t = 0
    ll_bad, g_bad = mcp.gradlogpdf(pd, mcp.DeviceAlignment(bad, leaf_nums, 4))
    assert ll_bad == -np.inf
    assert not np.all(np.isfinite(g_bad))


@pytest.mark.parametrize("K,cpt", [(2, 2), (3, 1), (4, 1), (4, 2), (5, 1)])
def test_gradient_is_bit_reproducible(K, cpt):
    """The walk accumulates the branch sums without atomics (per-warp sums parked per chunk, folded in
    fixed order): repeated evaluations return identical bits, logL and every gradient component."""
    rng = np.random.default_rng(600 + K)
    tree = random_tree(60, rng)
    pi = rng.dirichlet(np.ones(K) * 5)
    model, pi, srates = _model(K, pi, rng)
    rates = mcp.discrete_gamma_rates(0.5, 0.5, 4)
    codes, leaf_nums = simulate_codes(tree, model(pi, srates), pi, rates, 5000, rng, gap_frac=0.02)
    pd = mcp.PhyloDist(tree, pi, srates, rates, model)
    aln = mcp.DeviceAlignment(codes, leaf_nums, K)
    ctx = mcp.get_context()
    try:
        ctx.set_level_mode(0)
        ctx.set_columns_per_thread(cpt)
        runs = [mcp.gradlogpdf(pd, aln) for _ in range(4)]
    finally:
        ctx.set_level_mode(-1)
        ctx.set_columns_per_thread(0)
    assert all(r[0] == runs[0][0] for r in runs)
    assert all(np.array_equal(r[1], runs[0][1]) for r in runs)


@pytest.mark.parametrize("K,R,null_eig", [(2, 1, True), (4, 4, True), (4, 2, False), (6, 1, True)])
def test_gradient_accumulator_in_global_memory(oracle, K, R, null_eig):
    """Trees beyond 4096 nodes accumulate the branch sums in the CTA's row in global memory (RED.ADD.F64)
    instead of shared memory; the mode can be forced, which is how a small tree gets to test it."""
    rng = np.random.default_rng(800 + K + R)
    tree = random_tree(45, rng, multifurcate=True)
    pi = rng.dirichlet(np.ones(K) * 5)
    base, pi, srates = _model(K, pi, rng)
    model = base if null_eig else _shifted(base, 0.25)
    rates = mcp.discrete_gamma_rates(0.5, 0.5, R) if R > 1 else np.ones(1)
    codes, leaf_nums = simulate_codes(tree, base(pi, srates), pi, rates, 1300, rng, gap_frac=0.05)
    pd = mcp.PhyloDist(tree, pi, srates, rates, model)
    aln = mcp.DeviceAlignment(codes, leaf_nums, K)
    ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, K, model, pi, srates, rates)
    ctx = mcp.get_context()
    try:
        ctx.set_level_mode(0)
        results = []
        for mode in (1, 0):
            ctx.set_accumulator_mode(mode)
            for cpt in (1, 2):          # two columns per thread falls back to one in global mode
                ctx.set_columns_per_thread(cpt)
                ll, g = mcp.gradlogpdf(pd, aln)
                _check(ll, g, ll_o, g_o)
                _check(mcp.logpdf(pd, aln), None, ll_o, None)
                results.append((ll, g))
        assert all(r[0] == results[0][0] for r in results[:2])       # logL is order-independent of the accumulator
    finally:
        ctx.set_accumulator_mode(-1)
        ctx.set_level_mode(-1)
        ctx.set_columns_per_thread(0)


def test_tree_beyond_the_shared_accumulator(oracle):
    """2100 taxa = 4199 nodes: the automatic choice moves the gradient accumulator to global memory
    (and would otherwise need 34 KB of shared memory per CTA for it)."""
    rng = np.random.default_rng(2100)
    K = 4
    tree = random_tree(2100, rng)
    pi = rng.dirichlet(np.ones(K) * 5)
    srates = rng.uniform(0.5, 2.5, size=6)
    rates = mcp.discrete_gamma_rates(0.5, 0.5, 2)
    codes, leaf_nums = simulate_codes(tree, mcp.GTR(pi, srates), pi, rates, 96, rng, gap_frac=0.02)
    pd = mcp.PhyloDist(tree, pi, srates, rates, mcp.GTR)
    aln = mcp.DeviceAlignment(codes, leaf_nums, K)
    ll, g = mcp.gradlogpdf(pd, aln)
    ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, K, mcp.GTR, pi, srates, rates)
    _check(ll, g, ll_o, g_o)
    ctx = mcp.get_context()
    try:
        ctx.set_accumulator_mode(0)        # forced shared memory still works at this size
        ll2, g2 = mcp.gradlogpdf(pd, aln)
        _check(ll2, g2, ll_o, g_o)
        assert ll2 == ll
    finally:
        ctx.set_accumulator_mode(-1)
