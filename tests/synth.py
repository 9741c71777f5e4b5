from mcphylo_jl_b200.synthetic import random_tree, simulate_codes  # noqa: F401


def reroot_variants(tree, delta_frac=0.37):
    """Two trees with the same likelihood as `tree` under any reversible model (pulley principle):
    (1) a copy with part of the root's first child branch moved to the second child's branch;
    (2) if the first child is internal, the unrooted version: that child becomes a trifurcating root
    over its own children and the old sibling, whose branch carries the sum of the two root branches.
    Leaf names (hence leaf numbers) are unchanged, so the same alignment rows apply."""
    import copy

    import mcphylo_jl_b200 as mcp

    assert tree.nchild == 2
    shifted = copy.deepcopy(tree)
    a, b = shifted.children
    d = delta_frac * a.inc_length
    a.inc_length -= d
    b.inc_length += d
    mcp.number_nodes(shifted)
    unrooted = None
    t2 = copy.deepcopy(tree)
    a, b = t2.children
    if a.nchild == 0:
        a, b = b, a
    if a.nchild > 0:
        total = a.inc_length + b.inc_length
        t2.remove_child(b)
        t2.remove_child(a)
        a.mother = None
        a.root = True
        a.add_child(b)
        b.inc_length = total
        unrooted = a
        mcp.number_nodes(unrooted)
    return shifted, unrooted
