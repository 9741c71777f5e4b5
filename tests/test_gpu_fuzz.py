"""Seeded randomised parity sweep in the GPU suite: random trees (multifurcations, unary nodes),
K in 2..20, 1-4 rate categories, ragged site counts, gap fractions and every kernel variant (tile
width, columns per thread, scratch placement, level-parallel kernel) against the CPU oracle at the
acceptance tolerances.  The sweep itself lives in tools/fuzz_parity.py (larger runs:
profiles/r1_fuzz_parity.log)."""
import importlib.util
import os

import numpy as np
import pytest

import mcphylo_jl_b200 as mcp

pytestmark = pytest.mark.gpu

_spec = importlib.util.spec_from_file_location(
    "fuzz_parity", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "fuzz_parity.py"))
fuzz = importlib.util.module_from_spec(_spec)


@pytest.mark.parametrize("seed", [101, 102])
def test_fuzz_baseline_regime(oracle, seed):
    _spec.loader.exec_module(fuzz)
    rng = np.random.default_rng(seed)
    ctx = mcp.get_context(0)
    failures = []
    try:
        for i in range(30):
            ok, desc, e_ll, e_g, verdict = fuzz.one_case(rng, ctx, i, stress=False)
            if not ok:
                failures.append(f"{verdict} {desc}")
    finally:
        ctx.set_launch(0, 0)
        ctx.set_columns_per_thread(0)
        ctx.set_scratch_mode(-1)
        ctx.set_level_mode(-1)
    assert not failures, "\n".join(failures)


@pytest.mark.parametrize("seed,index,stress", [(24, 16, False), (22, 48, False), (33, 49, True)])
def test_cases_where_the_reference_formula_is_the_inaccurate_side(oracle, seed, index, stress):
    """Three cases found by the sweep in which the CUDA path and the fp64 oracle differ by more than
    the acceptance tolerance on one gradient component (5.5e-8, 1.2e-8, 1.1e-5): slow rate categories
    (1e-4 .. 1e-8) on short branches, where the reference's P = U diag(e) Uinv is cancellation noise.
    The extended-precision arbiter sides with the CUDA path to 1e-12."""
    _spec.loader.exec_module(fuzz)
    rng = np.random.default_rng(seed)
    for _ in range(index + 1):
        c = fuzz.make_case(rng, stress)
    ft = mcp.flatten(c["tree"])
    U, D, Uinv, mu = c["model"](c["pi"], c["srates"])
    ll_x, g_x = oracle.felsenstein_extended(c["codes"], c["leaf_nums"], c["K"], ft.postorder_num, ft.parent_num,
                                            ft.blv, U, D, Uinv, mu, c["rates"], c["pi"])
    ll_x, g_x = float(ll_x), g_x.astype(np.float64)
    pd = mcp.PhyloDist(c["tree"], c["pi"], c["srates"], c["rates"], c["model"])
    ctx = mcp.get_context(0)
    try:
        for levels in (0, 1) if c["K"] <= 6 else (-1,):
            ctx.set_level_mode(levels)
            ll, g = mcp.gradlogpdf(pd, mcp.DeviceAlignment(c["codes"], c["leaf_nums"], c["K"]))
            assert abs(ll - ll_x) <= 1e-12 * abs(ll_x)
            assert np.max(np.abs(g - g_x) / np.maximum(np.abs(g_x), 1e-3 * np.max(np.abs(g_x)))) <= 1e-11
    finally:
        ctx.set_level_mode(-1)
