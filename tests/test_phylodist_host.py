"""Host-side behaviour of the PhyloDist / MultiplePhyloDist mirror that needs no GPU: the
constructor forms and error cases the reference tests in test/distributions/phylodist.jl:17-94."""
import warnings

import numpy as np
import pytest

import mcphylo_jl_b200 as mcp
from conftest import golden_case


class _TreeVariate:            # stands in for the reference's Stochastic tree node (`.value` holds the tree)
    def __init__(self, value):
        self.value = value


@pytest.fixture(scope="module")
def primates():
    tree, x, codes, leaf_nums, fx = golden_case("primates")
    return tree, x, fx


def test_constructor_forms_are_equal(primates):
    tree, _, fx = primates
    pden = np.ones(4) / 4
    pd = mcp.PhyloDist(tree, pden, [1.0], [1.0], mcp.JC)
    s = _TreeVariate(tree)
    pd2 = mcp.PhyloDist(s, pden, [1.0], [1.0], mcp.JC)
    pd3 = mcp.PhyloDist(s, pden, 1.0, 1.0, mcp.JC)
    pd4 = mcp.PhyloDist(tree, pden, 1.0, 1.0, mcp.JC)
    pd6 = mcp.PhyloDist(tree, [0.25, 0.25, 0.25, 0.25], [1.0], [1.0], mcp.JC)
    assert all(y == pd for y in (pd2, pd3, pd4, pd6))
    assert mcp.minimum(pd) == -np.inf and mcp.maximum(pd) == np.inf
    assert mcp.size(pd) == (4, 1, 22) == tuple(fx["size"])
    assert pd != mcp.PhyloDist(tree, pden, [1.0], [2.0], mcp.JC)


def test_freek_constructor(primates):
    tree, _, _ = primates
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pd5 = mcp.PhyloDist(tree, [1.0], [1.0], mcp.freeK)
    assert pd5.substitution_model is mcp.freeK
    assert np.array_equal(pd5.base_freq, [1.0]) and pd5.nbase == 1
    # two-state freeK: equilibrium frequencies of Q = [[-a, a], [b, -b]] are (b, a) / (a + b)
    pd = mcp.PhyloDist(tree, [0.3, 0.9], [1.0], mcp.freeK)
    U, D, Uinv, mu = mcp.freeK([], [0.3, 0.9])
    Q = U @ np.diag(D) @ Uinv
    assert np.allclose(pd.base_freq @ Q, 0.0, atol=1e-14) and np.isclose(pd.base_freq.sum(), 1.0)


def test_multiple_phylodist_constructors(primates):
    tree, _, _ = primates
    trees = [tree, tree]
    pden = np.ones(4) / 4
    freqs = np.full((4, 2), 0.25)
    rates = np.ones((1, 2))
    mpd = mcp.MultiplePhyloDist(trees, freqs, rates, rates, mcp.JC)
    for other in (mcp.MultiplePhyloDist(trees, pden, [1.0], [1.0], mcp.JC),
                  mcp.MultiplePhyloDist(trees, freqs, [1.0], [1.0], mcp.JC),
                  mcp.MultiplePhyloDist(trees, freqs, rates, [1.0], mcp.JC),
                  mcp.MultiplePhyloDist(trees, freqs, [1.0], rates, mcp.JC)):
        assert other == mpd
    assert mcp.size(mpd) == (4, 1, 22, 2)
    assert mcp.minimum(mpd) == -np.inf and mcp.maximum(mpd) == np.inf
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mpd6 = mcp.MultiplePhyloDist(trees, [1.0], [1.0], mcp.freeK)
        mpd7 = mcp.MultiplePhyloDist(trees, rates, [1.0], mcp.freeK)
    assert mpd6 == mpd7
    assert all(d.substitution_model is mcp.freeK and np.array_equal(d.base_freq, [1.0]) and d.nbase == 1
               for d in mpd6.DistCollector)
    bad = np.array([[1.0, 2.0, 3.0]])
    with pytest.raises(mcp.DimensionMismatch):
        mcp.MultiplePhyloDist(trees, np.full((2, 3), 0.25), [1.0], [1.0], mcp.JC)
    with pytest.raises(mcp.DimensionMismatch):
        mcp.MultiplePhyloDist(trees, freqs, bad, [1.0], mcp.JC)
    with pytest.raises(mcp.DimensionMismatch):
        mcp.MultiplePhyloDist(trees, freqs, [1.0], bad, mcp.JC)
    with pytest.raises(mcp.DimensionMismatch):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mcp.MultiplePhyloDist(trees, bad, [1.0], mcp.freeK)


def test_device_alignment_site_blocks():
    codes = np.arange(40, dtype=np.uint8).reshape(4, 10) % 5
    aln = mcp.DeviceAlignment(codes, [1, 2, 3, 4], 4)
    assert aln.S == 10
    parts = [mcp.local_shard(aln, 3, r) for r in range(3)]
    assert [p.S for p in parts] == [4, 4, 2]
    assert np.array_equal(np.concatenate([p.codes for p in parts], axis=1), codes)


def test_pipelined_block_plan_covers_sites_in_wave_multiples():
    from mcphylo_jl_b200.dist import PipelinedEvaluator as P

    for S, wave, nb in [(1000000, 28416, 5), (1000, 28416, 5), (60000, 28416, 5), (125000, 28416, 3), (0, 100, 4),
                        (999, 1, 6), (28416 * 7, 28416, 8)]:
        b = P.plan_blocks(S, wave, nb)
        assert len(b) <= max(nb, 1)
        if S == 0:
            assert b == []
            continue
        assert b[0][0] == 0 and b[-1][1] == S
        assert all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
        assert all(hi > lo for lo, hi in b)
        assert all((hi - lo) % wave == 0 for lo, hi in b[:-1])         # only the last block is ragged
