"""Synthetic inputs for tests and bench.py: random rooted trees by joining two
random active lineages, and alignments simulated down the tree under the evaluating model.
Host-side numpy only; nothing here is on the evaluation path."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

from .substitution_models import transition_matrix
from .tree import GeneralNode, get_leaves, number_nodes, pre_order


def random_tree(n_taxa: int, rng: np.random.Generator, multifurcate: bool = False, unary: bool = False,
                mean_bl: float = 0.1) -> GeneralNode:
    """Leaves t0000.. (so alphabetical = creation order); branch lengths Exp(mean) clipped to
    [1e-4, 1].  multifurcate: the root gets three children (and one inner node too when the tree
    is large enough); unary: one single-child node is spliced in."""
    def bl():
        return float(np.clip(rng.exponential(mean_bl), 1e-4, 1.0))

    width = max(4, len(str(n_taxa - 1)))
    active = [GeneralNode(f"t{i:0{width}d}", bl()) for i in range(n_taxa)]
    k = 0
    stop_at = 3 if (multifurcate and n_taxa >= 3) else 2
    while len(active) > stop_at:
        i, j = rng.choice(len(active), size=2, replace=False)
        a, b = active[i], active[j]
        node = GeneralNode(f"n{k:0{width}d}", bl())
        k += 1
        node.add_child(a)
        node.add_child(b)
        if multifurcate and len(active) > 6 and k == 2:
            c_idx = [m for m in range(len(active)) if m not in (i, j)][0]
            c = active[c_idx]
            node.add_child(c)
            active = [n for m, n in enumerate(active) if m not in (i, j, c_idx)]
        else:
            active = [n for m, n in enumerate(active) if m not in (i, j)]
        active.append(node)
    root = GeneralNode("root", 1.0)
    for n in active:
        root.add_child(n)
    if unary and n_taxa >= 2:
        # splice a single-child node above the first grandchild-bearing child (or any child)
        target = root.children[-1]
        mid = GeneralNode("unary", bl())
        idx = root.children.index(target)
        root.children[idx] = mid
        mid.mother = root
        mid.root = False
        target.mother = None
        mid.add_child(target)
    root.mother = None
    root.root = True
    number_nodes(root)
    return root


def simulate_codes(tree: GeneralNode, model_out, pi: Sequence[float], rates: Sequence[float], S: int,
                   rng: np.random.Generator, gap_frac: float = 0.01,
                   site_cats: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
    """(codes uint8 (n_leaves, S), leaf_nums int32).  Root state ~ pi, child state ~ row of
    P(t * rate of the site's category); categories uniform over R; a fraction of cells becomes
    gap (code K).  Rows follow get_leaves(tree)."""
    pi = np.asarray(pi, dtype=np.float64)
    K = pi.size
    rates = np.asarray(rates, dtype=np.float64)
    R = rates.size
    cats = rng.integers(0, R, size=S) if site_cats is None else site_cats
    cat_idx = [np.nonzero(cats == r)[0] for r in range(R)]
    states = {}
    cpi = np.cumsum(pi / pi.sum())
    states[id(tree)] = np.minimum((rng.random(S)[:, None] > cpi[None, :]).sum(axis=1), K - 1).astype(np.uint8)
    leaves = get_leaves(tree)
    row_of = {id(l): i for i, l in enumerate(leaves)}
    codes = np.empty((len(leaves), S), dtype=np.uint8)
    for node in pre_order(tree):
        if node is tree:
            continue
        ps = states[id(node.mother)]
        st = np.empty(S, dtype=np.uint8)
        for r in range(R):
            idx = cat_idx[r]
            if idx.size == 0:
                continue
            P = np.clip(transition_matrix(model_out, node.inc_length, rates[r]), 0.0, None)
            cum = np.cumsum(P / P.sum(axis=1, keepdims=True), axis=1)
            u = rng.random(idx.size)
            st[idx] = np.minimum((u[:, None] > cum[ps[idx]]).sum(axis=1), K - 1)
        if node.nchild == 0:
            codes[row_of[id(node)]] = st
        else:
            states[id(node)] = st
        # a mother's states can be dropped once all her children are done
        if node is node.mother.children[-1]:
            states.pop(id(node.mother), None)
    if gap_frac > 0:
        mask = rng.random(codes.shape) < gap_frac
        codes[mask] = K
    leaf_nums = np.asarray([l.num for l in leaves], dtype=np.int32)
    return codes, leaf_nums
