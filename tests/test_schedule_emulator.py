"""The host scheduler (csrc/schedule.hpp through mcp_schedule_dump) checked on CPU: its walk
program, executed by a numpy emulator of the kernel's arithmetic, must reproduce the oracle and
the reference goldens.  Also checks the structural promises the kernel relies on."""
import numpy as np
import pytest

import mcphylo_jl_b200 as mcp
from mcphylo_jl_b200 import capi
from conftest import golden_case
from emulator import run_program
from synth import random_tree, simulate_codes


def _leaf_row(ft, leaf_nums):
    lr = np.full(ft.NN, -1, dtype=np.int32)
    for i, n in enumerate(leaf_nums):
        lr[n - 1] = i
    return lr


def _emulate(oracle, tree, codes, leaf_nums, model_out, rates, pi, want_grad=True, cherries=False):
    ft = mcp.flatten(tree)
    U, D, Uinv, mu = model_out
    P, dP = oracle.transition(U, D, Uinv, mu, np.asarray(rates, float), ft.blv, want_dP=True)
    prog = capi.schedule_dump(ft.postorder_num, ft.parent_num, _leaf_row(ft, leaf_nums), want_grad, cherries=cherries)
    ll, g = run_program(prog, codes, len(D), P, dP, np.asarray(pi, float), ft.NN - 1)
    return ll, g[:ft.NN - 1], prog, ft


def test_goldens_through_the_schedule(oracle):
    tree, x, codes, leaf_nums, fx = golden_case("primates")
    ll, _, prog, ft = _emulate(oracle, tree, codes, leaf_nums, mcp.JC(fx["base_freq"], [1.0]), [1.0], fx["base_freq"])
    assert abs(ll - fx["logpdf"]) <= 1e-12 * abs(ll)
    assert prog["n_dnodes"] == ft.NN + 1            # one virtual node for the trifurcating root

    tree, x, codes, leaf_nums, fx = golden_case("simudata")
    ll, g, _, _ = _emulate(oracle, tree, codes, leaf_nums, mcp.JC(fx["base_freq"], [1.0]), [1.0], fx["base_freq"])
    assert np.max(np.abs(g - fx["grad"]) / np.abs(fx["grad"])) <= 1e-11


@pytest.mark.parametrize("n_taxa,K,R,seed", [(2, 2, 1, 0), (3, 4, 2, 1), (17, 2, 1, 2), (40, 4, 4, 3), (64, 3, 2, 4)])
def test_random_trees_match_oracle(oracle, n_taxa, K, R, seed):
    rng = np.random.default_rng(seed)
    tree = random_tree(n_taxa, rng, multifurcate=(seed % 2 == 1), unary=(seed == 3))
    pi = rng.dirichlet(np.ones(K) * 5)
    if K == 2:
        model = mcp.Restriction(pi, [])
    elif K == 4:
        model = mcp.GTR(pi, rng.uniform(0.5, 2.0, size=6))
    else:
        model = mcp.JC(pi, [])
    rates = mcp.discrete_gamma_rates(0.5, 0.5, R) if R > 1 else np.ones(1)
    codes, leaf_nums = simulate_codes(tree, model, pi, rates, 37, rng, gap_frac=0.1)
    ft = mcp.flatten(tree)
    x = oracle.codes_to_dense(codes, leaf_nums, K, ft.NN)
    U, D, Uinv, mu = model
    ll_o, g_o = oracle.felsenstein(x, ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, rates, pi, True, 1)
    ll, g, prog, _ = _emulate(oracle, tree, codes, leaf_nums, model, rates, pi)
    assert abs(ll - ll_o) <= 1e-11 * abs(ll_o)
    assert np.allclose(g, g_o, rtol=1e-9, atol=1e-9)
    # cherries recomputed in the gradient pass instead of stored: same values, fewer stored partials
    ll_c, g_c, prog_c, _ = _emulate(oracle, tree, codes, leaf_nums, model, rates, pi, cherries=True)
    assert ll_c == ll and np.array_equal(g_c, g)
    stored = lambda pr: int(np.sum((pr["post"][:, 5] & 16) != 0))
    assert stored(prog_c) == stored(prog) - prog_c["n_cherries"]
    assert prog_c["n_cherries"] == int(np.sum((prog_c["pre"][:, 5] & 3) == 3)) and prog["n_cherries"] == 0
    if n_taxa >= 17:
        assert prog_c["n_cherries"] >= 1
    # logL-only program (LIFO slots) gives the same value
    ll2, _, prog2, _ = _emulate(oracle, tree, codes, leaf_nums, model, rates, pi, want_grad=False)
    assert abs(ll2 - ll_o) <= 1e-11 * abs(ll_o)
    assert len(prog2["pre"]) == 0
    assert prog2["n_slots"] <= int(np.log2(max(n_taxa, 2))) + 2


def test_structure_of_the_program():
    rng = np.random.default_rng(11)
    tree = random_tree(200, rng)
    ft = mcp.flatten(tree)
    lr = _leaf_row(ft, ft.leaf_nums)
    prog = capi.schedule_dump(ft.postorder_num, ft.parent_num, lr, True)
    post, pre = prog["post"], prog["pre"]
    n_int = ft.NN - len(ft.leaf_nums)
    assert len(post) == n_int and len(pre) == n_int
    # at most one MEM and one REG operand per op; REG only refers to the previous op
    for i, op in enumerate(post):
        kinds = [op[5] & 3, (op[5] >> 2) & 3]
        assert kinds.count(2) <= 1 and kinds.count(1) <= 1
        if 1 in kinds:
            assert i > 0
        for kind, src in zip(kinds, (op[0], op[2])):
            if kind == 2:
                assert 0 <= src < i     # a slot is the index of an earlier op
    assert (post[-1][5] >> 5) & 1 == 1 and sum((op[5] >> 5) & 1 for op in post) == 1
    # canonical operand order, which felsenstein_walk branches on: post ops are (LEAF, LEAF),
    # (REG, LEAF) or (MEM, REG); pre families are (leaf, leaf), (KEEP, leaf) or (KEEP, PUSH), and a
    # kept pre vector is consumed by the very next op (PREM_REG), everything else comes off the LIFO
    for variant in (prog, capi.schedule_dump(ft.postorder_num, ft.parent_num, lr, False)):
        assert {(op[5] & 3, (op[5] >> 2) & 3) for op in variant["post"]} <= {(0, 0), (1, 0), (2, 1)}
    for i, op in enumerate(pre):
        fl = op[5]
        ka, kb, mk, ao, bo = fl & 3, (fl >> 2) & 3, (fl >> 8) & 3, (fl >> 10) & 3, (fl >> 12) & 3
        assert (ka, kb, ao, bo) in {(0, 0, 0, 0), (2, 0, 1, 0), (2, 2, 1, 2)}
        assert mk == (0 if i == 0 else (1 if ((pre[i - 1][5] >> 10) & 3) == 1 else 2))
    # LIFO depth of the gradient pass is logarithmic for any topology
    assert prog["n_stack"] <= int(np.log2(200)) + 1
    cat = mcp.ParseNewick("(" * 199 + "t000:1," + ",".join(f"t{i:03d}:1)" for i in range(1, 200)) + ";")
    ftc = mcp.flatten(cat)
    progc = capi.schedule_dump(ftc.postorder_num, ftc.parent_num, _leaf_row(ftc, ftc.leaf_nums), True)
    assert progc["n_stack"] == 1
    progc0 = capi.schedule_dump(ftc.postorder_num, ftc.parent_num, _leaf_row(ftc, ftc.leaf_nums), False)
    assert progc0["n_slots"] == 1


@pytest.mark.parametrize("n_taxa,seed,cherries", [(2, 0, True), (3, 1, True), (40, 2, True), (200, 3, False), (200, 4, True)])
def test_fetch_list_of_the_operand_ring(n_taxa, seed, cherries):
    """The fetch list the walk kernel's operand ring follows (mcp_schedule_fetch_list) is exactly the sequence of
    stored child partials the gradient program reads -- per family child a, then child b -- and every entry is a
    slot the post program really stores (a recomputed cherry has none and is never fetched)."""
    rng = np.random.default_rng(seed)
    tree = random_tree(n_taxa, rng, multifurcate=(seed % 2 == 1), unary=(seed == 3))
    ft = mcp.flatten(tree)
    lr = _leaf_row(ft, ft.leaf_nums)
    prog = capi.schedule_dump(ft.postorder_num, ft.parent_num, lr, True, cherries=cherries)
    fetch = capi.schedule_fetch_list(ft.postorder_num, ft.parent_num, lr, cherries=cherries)
    want = []
    for op in prog["pre"]:
        if (op[5] & 3) == 2:
            want.append(op[0])
        if ((op[5] >> 2) & 3) == 2:
            want.append(op[2])
    assert fetch.tolist() == want
    stored = {int(op[4]) for op in prog["post"] if op[5] & 16}
    assert set(want) <= stored
    # every stored post result is read exactly once by the gradient pass (the post pass may read it as well)
    assert sorted(want) == sorted(stored)
    if cherries and n_taxa >= 40:
        assert prog["n_cherries"] > 0 and len(want) == len(prog["post"]) - 1 - prog["n_cherries"]


def test_bad_trees_are_rejected():
    po = np.array([1, 2, 3], dtype=np.int32)
    with pytest.raises(capi.McpError):   # leaf without alignment row
        capi.schedule_dump(po, [3, 3, 0], [-1, 0, -1], True)
    with pytest.raises(capi.McpError):   # root not last
        capi.schedule_dump([3, 1, 2], [3, 3, 0], [0, 1, -1], True)
    with pytest.raises(capi.McpError):   # cycle / mother before child
        capi.schedule_dump(po, [2, 1, 0], [0, 1, -1], True)
    with pytest.raises(capi.McpError):
        capi.schedule_dump([1], [0], [0], True)


@pytest.mark.parametrize("n_taxa,K,seed", [(2, 2, 0), (9, 4, 1), (50, 2, 2), (33, 3, 3)])
def test_level_ordered_program(oracle, n_taxa, K, seed):
    """The level-ordered variant (small-tree kernel) has the same op semantics: executed
    sequentially by the emulator it must reproduce the oracle, and its level count is the tree
    height, not the node count."""
    rng = np.random.default_rng(50 + seed)
    tree = random_tree(n_taxa, rng, multifurcate=(seed % 2 == 1), unary=(seed == 3))
    pi = rng.dirichlet(np.ones(K) * 5)
    model = mcp.Restriction(pi, []) if K == 2 else (mcp.GTR(pi, rng.uniform(0.5, 2.0, size=6)) if K == 4 else mcp.JC(pi, []))
    codes, leaf_nums = simulate_codes(tree, model, pi, np.ones(1), 29, rng, gap_frac=0.1)
    ft = mcp.flatten(tree)
    x = oracle.codes_to_dense(codes, leaf_nums, K, ft.NN)
    U, D, Uinv, mu = model
    ll_o, g_o = oracle.felsenstein(x, ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, np.ones(1), pi, True, 1)
    P, dP = oracle.transition(U, D, Uinv, mu, np.ones(1), ft.blv, want_dP=True)
    prog = capi.schedule_dump(ft.postorder_num, ft.parent_num, _leaf_row(ft, leaf_nums), True, by_levels=True)
    ll, g = run_program(prog, codes, K, P, dP, pi, ft.NN - 1)
    assert abs(ll - ll_o) <= 1e-11 * abs(ll_o)
    assert np.allclose(g[:ft.NN - 1], g_o, rtol=1e-9, atol=1e-9)
    n_int = len(prog["post"])
    assert prog["n_slots"] == n_int and prog["n_stack"] == n_int
    assert 1 <= prog["post_levels"] <= n_int and prog["pre_levels"] == prog["post_levels"]
    if n_taxa >= 33:
        assert prog["post_levels"] < n_int // 2
    # no REG / KEEP in this variant
    assert all((op[5] & 3) != 1 and ((op[5] >> 2) & 3) != 1 for op in prog["post"])
    assert all(((op[5] >> 10) & 3) != 1 and ((op[5] >> 12) & 3) != 1 and ((op[5] >> 8) & 3) != 1 for op in prog["pre"])


@pytest.mark.parametrize("n_taxa,K,R,seed,drop_null", [(17, 2, 1, 2, True), (40, 4, 4, 3, True), (25, 4, 2, 5, False),
                                                       (30, 6, 2, 6, True)])
def test_walk_kernel_arithmetic_in_eigen_space(oracle, n_taxa, K, R, seed, drop_null):
    """The arithmetic of felsenstein_walk (csrc/device_math.cuh) emulated in numpy: transitions applied as
    L + U (expm1 * (Uinv L)) over the K-1 non-null eigen-components of the reordered decomposition
    (capi.model_reorder = what every evaluation uploads), numerator q.(dP L) formed in eigen-space.
    Reproduces the oracle; with all K components (drop_null False) likewise."""
    rng = np.random.default_rng(700 + seed)
    tree = random_tree(n_taxa, rng, multifurcate=(seed % 2 == 1), unary=(seed == 3))
    pi = rng.dirichlet(np.ones(K) * 5)
    model = mcp.Restriction(pi, []) if K == 2 else (mcp.GTR(pi, rng.uniform(0.5, 2.0, size=6)) if K == 4 else mcp.JC(pi, []))
    rates = mcp.discrete_gamma_rates(0.5, 0.5, R) if R > 1 else np.ones(1)
    codes, leaf_nums = simulate_codes(tree, model, pi, rates, 41, rng, gap_frac=0.1)
    ft = mcp.flatten(tree)
    x = oracle.codes_to_dense(codes, leaf_nums, K, ft.NN)
    U, D, Uinv, mu = model
    ll_o, g_o = oracle.felsenstein(x, ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, rates, pi, True, 1)
    Ur, Dr, Uir, null_last = capi.model_reorder(U, D, Uinv)
    assert null_last
    P, dP = oracle.transition(U, D, Uinv, mu, np.asarray(rates, float), ft.blv, want_dP=True)
    prog = capi.schedule_dump(ft.postorder_num, ft.parent_num, _leaf_row(ft, leaf_nums), True)
    eigen = dict(U=Ur, D=Dr, Uinv=Uir, mu=mu, blv=ft.blv, rates=np.asarray(rates, float), NE=K - 1 if drop_null else K)
    ll, g = run_program(prog, codes, K, P, dP, np.asarray(pi, float), ft.NN - 1, eigen=eigen)
    assert abs(ll - ll_o) <= 1e-11 * abs(ll_o)
    assert np.allclose(g[:ft.NN - 1], g_o, rtol=1e-9, atol=1e-9)
