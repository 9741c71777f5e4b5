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
            ok, desc, e_ll, e_g = fuzz.one_case(rng, ctx, i, stress=False)
            if not ok:
                failures.append(f"{desc} logL rel {e_ll:.2e} grad rel {e_g:.2e}")
    finally:
        ctx.set_launch(0, 0)
        ctx.set_columns_per_thread(0)
        ctx.set_scratch_mode(-1)
        ctx.set_level_mode(-1)
    assert not failures, "\n".join(failures)
