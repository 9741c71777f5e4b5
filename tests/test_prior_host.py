"""Branch-length priors: host mirror against the reference's known answers
(/root/reference/test/distributions/treedists.jl:1-70) and finite differences."""
import numpy as np
import pytest

import mcphylo_jl_b200 as mcp

TREE17 = ("(((0:0.110833,1:0.0137979)10:0.146124,(2:0.197891,(3:0.132967,(4:0.0378759,5:0.089252)11:0.101833)"
          "12:0.184301)\n 13:0.0450774)14:0.335725,6:0.153197,(7:0.0216218,(8:0.0781687,9:0.120419)15:0.0209114)"
          "16:0.0209771);")
TREE05 = ("(((0:0.5,1:0.5)10:0.5,(2:0.5,(3:0.5,(4:0.5,5:0.5)11:0.5)12:0.5)\n 13:0.5)14:0.5,6:0.5,"
          "(7:0.5,(8:0.5,9:0.5)15:0.5)16:0.5);")


def test_uniform_branch_length():                     # treedists.jl:1-14
    tree = mcp.ParseNewick(TREE17)
    r, g = mcp.gradlogpdf(mcp.UniformBranchLength(), tree)
    assert r == 0 and np.all(g == 0) and g.size == 17
    assert mcp.logpdf(mcp.UniformBranchLength(), tree) == 0


def test_exponential_bl():                            # treedists.jl:16-30
    d = mcp.exponentialBL(1.0)
    assert d.scale == 1.0 and d.constraints is None
    tree = mcp.ParseNewick(TREE17)
    blv = mcp.get_branchlength_vector(tree)
    assert mcp.logpdf(d, tree) == pytest.approx(float(np.sum(-blv)), rel=1e-15)
    assert np.all(mcp.gradlogpdf(d, tree)[1] == -np.ones(17))


def test_compound_dirichlet_golden():                 # treedists.jl:44-69
    d = mcp.CompoundDirichlet(1.0, 1.0, 0.100, 1.0)
    assert (d.alpha, d.a, d.beta, d.c) == (1.0, 1.0, 0.1, 1.0) and d.constraints is None
    tree = mcp.ParseNewick(TREE05)
    r, g = mcp.gradlogpdf(d, tree)
    assert r == -37.39364370893438
    assert np.allclose(g, -1.9823529411764707, rtol=1e-15, atol=0)
    assert mcp.logpdf(d, tree) == r


def test_internal_external_map():
    tree = mcp.ParseNewick(TREE17)
    ie = mcp.internal_external(tree)
    assert ie.tolist() == [0] * 10 + [1] * 7          # leaves are numbered first, root (num 18) has no branch


@pytest.mark.parametrize("params", [(1.0, 1.0, 0.1, 1.0), (2.5, 0.7, 0.3, 1.9), (0.8, 1.6, 2.0, 0.4)])
def test_compound_dirichlet_gradient_is_the_derivative(params):
    d = mcp.CompoundDirichlet(*params)
    tree = mcp.ParseNewick(TREE17)
    blv = mcp.get_branchlength_vector(tree)
    ie = mcp.internal_external(tree)
    _, g = mcp.gradlogpdf(d, tree)
    for j in range(blv.size):
        h = 1e-6 * blv[j]
        up, dn = blv.copy(), blv.copy()
        up[j] += h
        dn[j] -= h
        fd = (mcp.internal_logpdf(d, up, ie) - mcp.internal_logpdf(d, dn, ie)) / (2 * h)
        assert g[j] == pytest.approx(fd, rel=1e-6, abs=1e-7)


def test_insupport():
    tree = mcp.ParseNewick(TREE17)
    d = mcp.CompoundDirichlet(1.0, 1.0, 0.1, 1.0)
    assert mcp.insupport(d, tree)
    blv = mcp.get_branchlength_vector(tree)
    blv[3] = 0.0
    mcp.set_branchlength_vector(tree, blv)
    assert not mcp.insupport(d, tree)


def test_oracle_prior_pinned_and_agrees_with_host_mirror(oracle):
    """The checker's own restatement reproduces the reference's known answer and the product's
    written-out derivative agrees with the complex-step derivative of the restated loop."""
    v, g = oracle.compound_dirichlet_gradlogpdf(1.0, 1.0, 0.1, 1.0, [0.5] * 17, [0] * 10 + [1] * 7)
    assert v == -37.39364370893438
    assert np.allclose(g, -1.9823529411764707, rtol=1e-15, atol=0)
    tree = mcp.ParseNewick(TREE17)
    blv, ie = mcp.get_branchlength_vector(tree), mcp.internal_external(tree)
    for params in [(2.5, 0.7, 0.3, 1.9), (0.8, 1.6, 2.0, 0.4)]:
        v_o, g_o = oracle.compound_dirichlet_gradlogpdf(*params, blv, ie)
        v_h, g_h = mcp.gradlogpdf(mcp.CompoundDirichlet(*params), tree)
        assert v_h == pytest.approx(v_o, rel=1e-14)
        assert np.allclose(g_h, g_o, rtol=1e-13, atol=0)
