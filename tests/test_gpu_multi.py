"""GPU tests of the multi-device context (mcp_create_multi), the per-rank context (mcp_create_rank),
the streamed evaluation (mcp_eval_streamed) and the plan cache -- all through the C ABI, against the
CPU oracle on identical inputs.

On a one-GPU box the site-shard / reduce logic is exercised with the SAME device listed several times
(PEER and HOST reductions accept that; NCCL needs distinct devices and is tested when the box has two)."""
import numpy as np
import pytest

import mcphylo_jl_b200 as mcp
from mcphylo_jl_b200 import capi
from synth import random_tree, simulate_codes
from test_gpu_parity import _check, _model, _oracle_eval

pytestmark = pytest.mark.gpu


def _case(n_taxa, K, R, S, seed, multi=False):
    rng = np.random.default_rng(seed)
    tree = random_tree(n_taxa, rng, multifurcate=multi)
    pi = rng.dirichlet(np.ones(K) * 5)
    model, pi, srates = _model(K, pi, rng)
    rates = mcp.discrete_gamma_rates(0.5, 0.5, R) if R > 1 else np.ones(1)
    codes, leaf_nums = simulate_codes(tree, model(pi, srates), pi, rates, S, rng, gap_frac=0.03)
    return tree, pi, model, srates, rates, codes, leaf_nums


def _targs(tree, pi, model, srates, rates):
    ft = mcp.flatten(tree)
    U, D, Uinv, mu = model(np.asarray(pi, float), np.asarray(srates, float))
    return ft, (ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, rates, pi)


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("reduce", [capi.REDUCE_PEER, capi.REDUCE_HOST])
@pytest.mark.parametrize("G", [2, 3, 5])
def test_multi_context_same_device_vs_oracle(oracle, reduce, G):
    """G site shards on one GPU: every shard is evaluated by its own member context, the parts are
    summed by the group -- against the oracle on the whole alignment."""
    tree, pi, model, srates, rates, codes, leaf_nums = _case(40, 4, 4, 1237, 31 + G, multi=True)
    ctx = capi.Context(devices=[0] * G, reduce=reduce)
    try:
        assert ctx.device_count == G and ctx.reduce_mode == reduce
        aln = ctx.alignment_from_codes(codes, 4, leaf_nums)
        ft, targs = _targs(tree, pi, model, srates, rates)
        ll, g = ctx.eval(aln, *targs, want_grad=True)
        ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, 4, model, pi, srates, rates)
        _check(ll, g, ll_o, g_o)
        ll_only, _ = ctx.eval(aln, *targs, want_grad=False)
        _check(ll_only, None, ll_o, None)
        # bit-reproducible: fixed summation order in PEER and HOST modes
        ll2, g2 = ctx.eval(aln, *targs, want_grad=True)
        assert ll2 == ll and np.array_equal(g, g2)
        st = ctx.stats()
        assert st["kernel_launches"] >= G
        aln.close()
    finally:
        ctx.close()


def test_multi_context_more_devices_than_sites(oracle):
    """Shards may be empty (G > S)."""
    tree, pi, model, srates, rates, codes, leaf_nums = _case(9, 2, 1, 3, 5)
    ctx = capi.Context(devices=[0] * 5, reduce=capi.REDUCE_PEER)
    try:
        aln = ctx.alignment_from_codes(codes, 2, leaf_nums)
        ft, targs = _targs(tree, pi, model, srates, rates)
        ll, g = ctx.eval(aln, *targs, want_grad=True)
        ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, 2, model, pi, srates, rates)
        _check(ll, g, ll_o, g_o)
        aln.close()
    finally:
        ctx.close()


def test_multi_context_posterior_adds_the_prior_once(oracle):
    tree, pi, model, srates, rates, codes, leaf_nums = _case(30, 2, 1, 900, 77)
    single = capi.Context(0)
    multi = capi.Context(devices=[0, 0, 0], reduce=capi.REDUCE_PEER)
    try:
        a1 = single.alignment_from_codes(codes, 2, leaf_nums)
        a3 = multi.alignment_from_codes(codes, 2, leaf_nums)
        ft, targs = _targs(tree, pi, model, srates, rates)
        for kind, params in ((1, [0.1]), (2, [1.0, 1.0, 0.1, 1.0])):
            lp1, g1 = single.eval_posterior(a1, *targs, prior_kind=kind, prior_params=params)
            lp3, g3 = multi.eval_posterior(a3, *targs, prior_kind=kind, prior_params=params)
            assert abs(lp1 - lp3) <= 1e-12 * abs(lp1)
            assert np.allclose(g1, g3, rtol=1e-11, atol=1e-11 * np.max(np.abs(g1)))
        a1.close()
        a3.close()
    finally:
        single.close()
        multi.close()


def test_multi_context_batch_of_trees(oracle):
    """mcp_eval_batch on a multi-device context: every tree's alignment is site-sharded."""
    ctx = capi.Context(devices=[0, 0], reduce=capi.REDUCE_HOST)
    try:
        alns, trees, want = [], [], []
        for i in range(4):
            tree, pi, model, srates, rates, codes, leaf_nums = _case(12 + 3 * i, 2, 1, 300 + 17 * i, 900 + i)
            alns.append(ctx.alignment_from_codes(codes, 2, leaf_nums))
            ft, targs = _targs(tree, pi, model, srates, rates)
            trees.append(targs)
            want.append(_oracle_eval(oracle, tree, codes, leaf_nums, 2, model, pi, srates, rates))
        ll, grads = ctx.eval_batch(alns, trees, want_grad=True)
        for t in range(4):
            _check(ll[t], grads[t], *want[t])
        for a in alns:
            a.close()
    finally:
        ctx.close()


def test_multi_context_rejects_foreign_alignment():
    tree, pi, model, srates, rates, codes, leaf_nums = _case(8, 2, 1, 64, 3)
    c1, c2 = capi.Context(0), capi.Context(devices=[0, 0], reduce=capi.REDUCE_PEER)
    try:
        a1 = c1.alignment_from_codes(codes, 2, leaf_nums)
        ft, targs = _targs(tree, pi, model, srates, rates)
        with pytest.raises(capi.McpError) as ei:
            c2.eval(a1, *targs)
        assert "another context" in str(ei.value) or "not created on this" in str(ei.value)
        with pytest.raises(capi.McpError):
            c2.eval_device(a1, *targs, want_grad=True, d_out_ptr=8)
        a1.close()
    finally:
        c1.close()
        c2.close()


@pytest.mark.parametrize("reduce", [capi.REDUCE_NCCL, capi.REDUCE_PEER, capi.REDUCE_HOST, capi.REDUCE_AUTO])
def test_two_gpus_vs_oracle(oracle, reduce):
    """World size 2 through the C entry: two distinct GPUs, one context, one all-reduce per evaluation."""
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    tree, pi, model, srates, rates, codes, leaf_nums = _case(60, 4, 4, 5000, 123)
    ctx = capi.Context(devices=[0, 1], reduce=reduce)
    try:
        if reduce == capi.REDUCE_AUTO:
            assert ctx.reduce_mode == capi.REDUCE_NCCL
        aln = ctx.alignment_from_codes(codes, 4, leaf_nums)
        ft, targs = _targs(tree, pi, model, srates, rates)
        ll, g = ctx.eval(aln, *targs, want_grad=True)
        ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, 4, model, pi, srates, rates)
        _check(ll, g, ll_o, g_o)
        import torch
        pinned = torch.from_numpy(codes).pin_memory()
        ll_s, g_s = ctx.eval_streamed(pinned.data_ptr(), 4, codes.shape[1], leaf_nums, *targs, want_grad=True)
        _check(ll_s, g_s, ll_o, g_o)
        aln.close()
    finally:
        ctx.close()


@pytest.mark.parametrize("mode", ["fused", "blocks"])
@pytest.mark.parametrize("devices", [None, [0, 0, 0]])
@pytest.mark.parametrize("S", [700, 40000])
def test_streamed_evaluation_vs_oracle(oracle, monkeypatch, devices, S, mode):
    """mcp_eval_streamed: the alignment stays in (pinned) host memory and is uploaded during the call --
    one walk launch whose tiles wait for ready flags ("fused", the default for inputs beyond the small-tree
    kernel), or one launch per site block ("blocks"); repeated calls with new branch lengths and with
    CHANGED codes give the oracle's values."""
    import torch
    if mode == "blocks":
        monkeypatch.setenv("MCPHYLO_B200_STREAM_BLOCKS", "1")
    n_taxa = 24 if S > 1000 else 50
    tree, pi, model, srates, rates, codes, leaf_nums = _case(n_taxa, 4, 2, S, 404)
    ctx = capi.Context(devices=devices, reduce=capi.REDUCE_PEER) if devices else capi.Context(0)
    try:
        pinned = torch.from_numpy(codes.copy()).pin_memory()
        ft, targs = _targs(tree, pi, model, srates, rates)
        ll, g = ctx.eval_streamed(pinned.data_ptr(), 4, S, leaf_nums, *targs, want_grad=True)
        ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, 4, model, pi, srates, rates)
        _check(ll, g, ll_o, g_o)
        blocks = ctx.stream_blocks(0)
        assert blocks and blocks[0][0] == 0 and all(b[1] > b[0] for b in blocks)
        if S >= 40000 and not devices:
            # blocks: small first block, then growing; fused: the whole range is one launch
            assert len(blocks) >= 3 if mode == "blocks" else len(blocks) == 1
            tl = ctx.stream_timeline(0)
            assert len(tl) == len(blocks) and all(t[1] >= t[0] >= 0 and t[4] >= t[3] >= 0 for t in tl)
        # new data in the same host buffer + new branch lengths: nothing stale may be reused
        codes2 = codes.copy()
        codes2[:, ::3] = (codes2[:, ::3] + 1) % 4
        pinned.copy_(torch.from_numpy(codes2))
        blv = mcp.get_branchlength_vector(tree) * 1.3
        mcp.set_branchlength_vector(tree, blv)
        ft, targs = _targs(tree, pi, model, srates, rates)
        ll2, g2 = ctx.eval_streamed(pinned.data_ptr(), 4, S, leaf_nums, *targs, want_grad=True)
        ll_o2, g_o2 = _oracle_eval(oracle, tree, codes2, leaf_nums, 4, model, pi, srates, rates)
        _check(ll2, g2, ll_o2, g_o2)
        ll3, _ = ctx.eval_streamed(pinned.data_ptr(), 4, S, leaf_nums, *targs, want_grad=False)
        _check(ll3, None, ll_o2, None)
    finally:
        ctx.close()


def test_plan_cache_alternating_topologies(oracle):
    """The PNUTS pattern: gradient on tree A, logpdf on an NNI neighbour B, back to A.  With the plan
    cache the second visit of a topology rebuilds nothing (stats.schedule_rebuilt == 0) and results
    stay equal to the oracle's."""
    tree, pi, model, srates, rates, codes, leaf_nums = _case(30, 2, 1, 2000, 55)
    import copy
    tree_b = copy.deepcopy(tree)
    inner = [n for n in mcp.post_order(tree_b) if n.nchild == 2 and not n.root and n.mother is not None]
    assert mcp.NNI(tree_b, inner[len(inner) // 2]) == 1
    mcp.number_nodes(tree_b)
    ctx = capi.Context(0)
    try:
        aln = ctx.alignment_from_codes(codes, 2, leaf_nums)
        fa, ta = _targs(tree, pi, model, srates, rates)
        fb, tb = _targs(tree_b, pi, model, srates, rates)
        want_a = _oracle_eval(oracle, tree, codes, leaf_nums, 2, model, pi, srates, rates)
        want_b = _oracle_eval(oracle, tree_b, codes, leaf_nums, 2, model, pi, srates, rates)
        seq = [(ta, True, want_a), (tb, False, want_b), (ta, True, want_a), (tb, True, want_b), (ta, False, want_a)]
        rebuilt = []
        for targs, wg, want in seq:
            ll, g = ctx.eval(aln, *targs, want_grad=wg)
            _check(ll, g if wg else None, want[0], want[1] if wg else None)
            rebuilt.append(ctx.stats()["schedule_rebuilt"])
        assert rebuilt == [1, 1, 0, 1, 1]      # (A,grad) (B,ll) hit-(A,grad) (B,grad) (A,ll)
        for targs, wg, want in seq:
            ctx.eval(aln, *targs, want_grad=wg)
            assert ctx.stats()["schedule_rebuilt"] == 0
        aln.close()
    finally:
        ctx.close()


def test_retry_after_failed_evaluation(oracle):
    """an evaluation that fails after the topology was planned (bad prior parameter) must not
    leave a context that silently evaluates the NEXT call on stale device data."""
    tree, pi, model, srates, rates, codes, leaf_nums = _case(20, 2, 1, 500, 8)
    tree2, *_ = _case(20, 2, 1, 500, 9)
    ctx = capi.Context(0)
    try:
        aln = ctx.alignment_from_codes(codes, 2, leaf_nums)
        ft, targs = _targs(tree, pi, model, srates, rates)
        ctx.eval(aln, *targs, want_grad=True)
        ft2, targs2 = _targs(tree2, pi, model, srates, rates)
        with pytest.raises(capi.McpError):
            ctx.eval_posterior(aln, *targs2, prior_kind=1, prior_params=[-1.0])
        with pytest.raises(capi.McpError):
            ctx.eval_posterior(aln, *targs2, prior_kind=9, prior_params=[1.0])
        ll, g = ctx.eval_posterior(aln, *targs2, prior_kind=0, prior_params=[0.0])
        ll_o, g_o = _oracle_eval(oracle, tree2, codes, leaf_nums, 2, model, pi, srates, rates)
        _check(ll, g, ll_o, g_o)
        aln.close()
    finally:
        ctx.close()


def test_odd_tile_widths_stay_inside_the_allocation(oracle):
    """tile widths that do not divide the row stride (block 96, 160, 224; two columns per
    thread) read codes past the end of a row; the allocation carries the slack and results are unchanged."""
    tree, pi, model, srates, rates, codes, leaf_nums = _case(20, 4, 1, 1000, 14)
    ctx = capi.Context(0)
    try:
        aln = ctx.alignment_from_codes(codes, 4, leaf_nums)
        ft, targs = _targs(tree, pi, model, srates, rates)
        ll_o, g_o = _oracle_eval(oracle, tree, codes, leaf_nums, 4, model, pi, srates, rates)
        for block, cpt in ((96, 1), (96, 2), (160, 2), (224, 2), (224, 1)):
            ctx.set_launch(block, 0)
            ctx.set_columns_per_thread(cpt)
            ctx.set_level_mode(0)
            ll, g = ctx.eval(aln, *targs, want_grad=True)
            _check(ll, g, ll_o, g_o)
            assert ctx.stats()["block"] == block and ctx.stats()["columns_per_thread"] == cpt
        aln.close()
    finally:
        ctx.close()
