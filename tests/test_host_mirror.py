"""Host-side mirror of the reference interface: substitution models, Gamma rates, parsers,
tree numbering.  Expected values are the reference's own (tests/golden/models.json,
parser_csv.json; sources cited inside those files)."""
import numpy as np
import pytest

import mcphylo_jl_b200 as mcp
from conftest import load_golden

M = load_golden("models")


def _same_up_to_column_sign(A, B, tol=1e-8):
    A, B = np.asarray(A), np.asarray(B)
    for j in range(A.shape[1]):
        if not (np.allclose(A[:, j], B[:, j], atol=tol) or np.allclose(A[:, j], -B[:, j], atol=tol)):
            return False
    return True


def test_restriction():
    g = M["Restriction"]
    U, D, Uinv, mu = mcp.Restriction(g["base_freq"], [0.0])
    assert np.allclose(U, g["U"]) and np.allclose(D, g["D"]) and np.allclose(Uinv, g["Uinv"])
    assert mu == pytest.approx(g["mu"], rel=1e-14)
    with pytest.raises(AssertionError):
        mcp.Restriction(np.ones(3) / 3, [0.1])


def test_jc_three_states():
    g = M["JC3"]
    U, D, Uinv, mu = mcp.JC(g["base_freq"], [0.0])
    assert np.allclose(U @ np.diag(D) @ Uinv, g["Q"], atol=1e-14)
    assert mu == pytest.approx(g["mu"], rel=1e-14)


@pytest.mark.parametrize("name", ["GTR", "freeK"])
def test_gtr_freek(name):
    g = M[name]
    if name == "GTR":
        U, D, Uinv, mu = mcp.GTR(g["base_freq"], g["rates"])
    else:
        U, D, Uinv, mu = mcp.freeK([0.0], g["rates"])
    assert np.allclose(D, g["D"], atol=1e-12)
    assert _same_up_to_column_sign(U, g["U"])
    assert _same_up_to_column_sign(np.asarray(Uinv).T, np.asarray(g["Uinv"]).T)
    assert mu == pytest.approx(g["mu"], rel=1e-13)
    # scale-free statement of the same golden
    Qg = np.asarray(g["U"]) @ np.diag(g["D"]) @ np.asarray(g["Uinv"])
    assert np.allclose(U @ np.diag(D) @ Uinv, Qg, atol=1e-12)


def test_setmatrix():
    g = M["setmatrix"]
    assert np.array_equal(mcp.setmatrix(g["in"]), np.asarray(g["out"]))
    with pytest.raises(ValueError):
        mcp.setmatrix([1.0, 2.0])


def test_gamma_rates():
    assert np.allclose(mcp.discrete_gamma_rates(*M["gamma_mean"]["args"]), M["gamma_mean"]["out"], rtol=1e-10)
    assert np.allclose(mcp.discrete_gamma_rates(*M["gamma_median"]["args"], method="median"),
                       M["gamma_median"]["out"], rtol=1e-10)
    assert np.allclose(mcp.median_boundaries(*M["median_boundaries"]["args"]), M["median_boundaries"]["out"], rtol=1e-10)
    assert np.allclose(mcp.mean_boundaries(*M["mean_boundaries"]["args"])[:3], M["mean_boundaries"]["out"], rtol=1e-10)


def test_numbering_rule():
    # leaves alphabetical (bytewise), internals in post-order, root last
    t = mcp.ParseNewick("((zeta:1,Alpha:2)in1:0.5,(beta:1,(delta:1,Beta:3):2)in2:0.1);")
    names = {n.name: n.num for n in mcp.post_order(t)}
    assert [names[k] for k in ("Alpha", "Beta", "beta", "delta", "zeta")] == [1, 2, 3, 4, 5]
    assert names["in1"] == 6 and names["in2"] == 8
    assert mcp.post_order(t)[-1].num == 9
    blv = mcp.get_branchlength_vector(t)
    assert blv[names["Beta"] - 1] == 3.0 and blv[names["in2"] - 1] == 0.1 and blv.size == 8
    ft = mcp.flatten(t)
    assert ft.parent_num[names["zeta"] - 1] == names["in1"] and ft.parent_num[8] == 0
    # round trip through our own writer
    t2 = mcp.ParseNewick(mcp.newick(t))
    assert [n.num for n in mcp.post_order(t2)] == [n.num for n in mcp.post_order(t)]
    assert np.array_equal(mcp.get_branchlength_vector(t2), blv)


def test_newick_whitespace_and_errors():
    a = mcp.ParseNewick("(a:1,\n   (b:2, c:3)\n x:4);")
    assert [n.name for n in mcp.post_order(a)][:4] == ["a", "b", "c", "x"]
    with pytest.raises(ValueError):
        mcp.ParseNewick("((a:1,b:2);")
    with pytest.raises(ValueError):
        mcp.ParseNewick("")


def test_datafortree_layout_and_numbering(tmp_path):
    g = load_golden("parser_csv")
    df = np.array([list(r) for r in g["rows"]], dtype="<U1")
    # any topology: the slot is decided by the leaf's alphabetical rank
    for nwk in ("(Welsh_N_0:1,Sardinian_N_0:1);", "(Sardinian_N_0:1,Welsh_N_0:1);"):
        tree = mcp.ParseNewick(nwk)
        x = mcp.datafortree(df, g["taxa"], tree, g["symbols"], g["gap"], g["missing"])
        assert x.shape == (2, 23, 3)
        assert np.array_equal(x[:, :, 0], np.asarray(g["slot1"]))
        assert np.array_equal(x[:, :, 1], np.asarray(g["slot2"]))
        codes, nums = mcp.codesfortree(df, g["taxa"], tree, g["symbols"], g["gap"], g["missing"])
        assert np.array_equal(mcp.dense_to_codes(x, nums), codes)
    with pytest.raises(ValueError):
        mcp.datafortree(df, g["taxa"], tree, ["0"], g["gap"], g["missing"])


def test_nexus_and_csv_parsers_agree(tmp_path):
    g = load_golden("parser_csv")
    csv = tmp_path / "m.csv"
    csv.write_text("\n".join(f"{t}," + ",".join(r) for t, r in zip(g["taxa"], g["rows"])) + "\n")
    nex = tmp_path / "m.nex"
    nex.write_text("#NEXUS\nbegin data;\n dimensions ntax=2 nchar=23;\n format missing=? gap=-;\n matrix\n"
                   + "\n".join(f"{t}    {r}" for t, r in zip(g["taxa"], g["rows"])) + "\n    ;\nend;\n")
    a = mcp.ParseNexus(str(nex))
    b = mcp.ParseCSV(str(csv), "-", "?", False)
    assert a[:5] == b[:5] and a[6] == b[6] and np.array_equal(a[5], b[5])
    assert a[0] == 2 and a[1] == 23 and a[4] == ["0", "1"]
    bad = tmp_path / "bad.nex"
    bad.write_text("not nexus\n")
    with pytest.raises(mcp.FileSyntaxError):
        mcp.ParseNexus(str(bad))


def test_model_reorder_moves_the_null_eigenvalue_last():
    """Host logic of the C library (no GPU): every evaluation works on a copy of the caller's
    eigen-decomposition with the null eigenvalue last; the product U diag(f(D)) Uinv is unchanged."""
    from mcphylo_jl_b200 import capi

    rng = np.random.default_rng(5)
    pi = rng.dirichlet(np.ones(4) * 5)
    U, D, Uinv, mu = mcp.GTR(pi, rng.uniform(0.5, 2.0, size=6))
    for shift in range(4):
        idx = np.roll(np.arange(4), shift)
        Uo, Do, Uio, null_last = capi.model_reorder(U[:, idx], D[idx], Uinv[idx, :])
        assert null_last and abs(Do[3]) <= 1e-14 * np.max(np.abs(Do)) and np.all(np.abs(Do[:3]) > 1e-3)
        assert np.allclose(Uo @ np.diag(np.exp(0.3 * Do)) @ Uio, U @ np.diag(np.exp(0.3 * D)) @ Uinv, rtol=0, atol=1e-15)
        # the other components keep their relative order
        keep = [i for i in idx if abs(D[i]) > 1e-3]
        assert np.array_equal(Do[:3], D[keep])
    # Restriction: D = [-1, 0] is already in place; a decomposition without a null eigenvalue is flagged
    U2, D2, Uinv2, _ = mcp.Restriction(np.array([0.3, 0.7]), [])
    assert capi.model_reorder(U2, D2, Uinv2)[3]
    Uo, Do, Uio, null_last = capi.model_reorder(U, D - 0.4, Uinv)
    assert not null_last and sorted(Do) == sorted(D - 0.4)
    assert np.argmin(np.abs(Do)) == 3
