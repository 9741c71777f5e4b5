"""Minimal host-side tree type for the PhyloDist hot path.

The reference takes its tree type from the un-vendored dependency MCPhyloTree.jl
(compat "1.1", /root/reference/Project.toml:20,52; re-exported at
/root/reference/src/MCPhylo.jl:17-18).  Only the accessors the likelihood path
touches are restated here, under the reference's names:

  GeneralNode fields   name, mother, children, nchild, root, inc_length, num
  ParseNewick          used at /root/reference/test/likelihood/felsenstein.jl:8-14
  post_order           /root/reference/src/distributions/Phylodist.jl:108,125
  get_leaves           Phylodist.jl:116
  get_branchlength_vector / set_branchlength_vector   Phylodist.jl:115,127
  get_mother           /root/reference/src/Likelihood/LikelihoodCalculator_Node.jl:26
  find_by_name         /root/reference/src/Parser/Parser.jl:69
  number_nodes         numbering rule pinned by SURVEY.md §8c: leaves get 1..N in
                       bytewise-sorted name order, internal nodes N+1.. in post-order,
                       root last (= NN).

Nothing here runs on the device; `flatten` turns a tree into the plain int32/double
arrays the C-ABI (include/mcphylo_b200.h) takes.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Iterable, List, Optional

import numpy as np


class GeneralNode:
    """One tree node; mirrors the fields of MCPhyloTree's GeneralNode used on the path."""

    __slots__ = ("name", "mother", "children", "inc_length", "num", "root")

    def __init__(self, name: str = "no_name", inc_length: float = 1.0):
        self.name = name
        self.mother: Optional["GeneralNode"] = None
        self.children: List["GeneralNode"] = []
        self.inc_length = float(inc_length)
        self.num = 0
        self.root = True

    @property
    def nchild(self) -> int:
        return len(self.children)

    def add_child(self, child: "GeneralNode") -> None:
        child.mother = self
        child.root = False
        self.children.append(child)

    def remove_child(self, child: "GeneralNode") -> None:
        self.children.remove(child)
        child.mother = None

    def __repr__(self) -> str:  # pragma: no cover - debugging aid
        return f"GeneralNode({self.name!r}, num={self.num}, nchild={self.nchild}, t={self.inc_length})"


Node = GeneralNode


def post_order(root: GeneralNode) -> List[GeneralNode]:
    """Children (in stored order) before their mother; the root comes last."""
    # reversed "root, then children right-to-left" pre-order == left-to-right post-order
    out: List[GeneralNode] = []
    stack = [root]
    pop, push, extend = stack.pop, out.append, stack.extend
    while stack:
        node = pop()
        push(node)
        extend(node.children)
    out.reverse()
    return out


def pre_order(root: GeneralNode) -> List[GeneralNode]:
    out: List[GeneralNode] = []
    stack = [root]
    while stack:
        node = stack.pop()
        out.append(node)
        stack.extend(reversed(node.children))
    return out


def get_leaves(root: GeneralNode) -> List[GeneralNode]:
    return [n for n in post_order(root) if n.nchild == 0]


def get_mother(node: GeneralNode) -> GeneralNode:
    return node.mother


def find_by_name(root: GeneralNode, name: str) -> GeneralNode:
    for n in post_order(root):
        if n.name == name:
            return n
    raise KeyError(f"no node named {name!r}")


def find_num(root: GeneralNode, num: int) -> GeneralNode:
    for n in post_order(root):
        if n.num == num:
            return n
    raise KeyError(f"no node numbered {num}")


def number_nodes(root: GeneralNode) -> None:
    """Leaves 1..N by sorted name (bytewise), internals N+1.. in post-order, root = NN."""
    po = post_order(root)
    leaves = [n for n in po if n.nchild == 0]
    leaves.sort(key=lambda n: n.name.encode("utf-8"))
    for i, n in enumerate(leaves):
        n.num = i + 1
    k = len(leaves)
    for n in po:
        if n.nchild > 0:
            k += 1
            n.num = k


def get_branchlength_vector(root: GeneralNode) -> np.ndarray:
    """blv[node.num - 1] = node.inc_length for every non-root node (length NN-1)."""
    po = post_order(root)
    out = np.zeros(len(po), dtype=np.float64)
    out[np.array([n.num for n in po]) - 1] = [n.inc_length for n in po]
    return out[:len(po) - 1].copy()


def set_branchlength_vector(root: GeneralNode, blv: Iterable[float]) -> None:
    vals = np.asarray(blv, dtype=np.float64).tolist()
    for n in post_order(root):
        if n is not root:
            n.inc_length = vals[n.num - 1]


def tree_length(root: GeneralNode) -> float:
    return float(sum(n.inc_length for n in post_order(root) if n is not root))


# ----------------------------------------------------------------------------------------
# Newick
# ----------------------------------------------------------------------------------------

def ParseNewick(s: str) -> GeneralNode:
    """Parse one Newick string.  Whitespace/newlines anywhere are ignored, internal node
    labels are kept as names, children stay in string order, nodes are numbered by
    `number_nodes`.  A missing branch length becomes 1.0 (the node constructor default)."""
    s = "".join(s.split())
    if s.endswith(";"):
        s = s[:-1]
    if not s:
        raise ValueError("empty Newick string")
    pos = 0
    n = len(s)

    def parse_label_and_length(node: GeneralNode) -> None:
        nonlocal pos
        start = pos
        while pos < n and s[pos] not in ",():;":
            pos += 1
        label = s[start:pos]
        if label:
            node.name = label
        if pos < n and s[pos] == ":":
            pos += 1
            start = pos
            while pos < n and s[pos] not in ",();":
                pos += 1
            node.inc_length = float(s[start:pos])

    root = GeneralNode()
    cur = root
    # iterative descent so thousand-taxon caterpillars do not hit the recursion limit
    if s[pos] != "(":
        parse_label_and_length(root)
    while pos < n:
        c = s[pos]
        if c == "(":
            child = GeneralNode()
            cur.add_child(child)
            cur = child
            pos += 1
            if pos < n and s[pos] != "(":
                parse_label_and_length(cur)
        elif c == ",":
            pos += 1
            sib = GeneralNode()
            cur.mother.add_child(sib)
            cur = sib
            if pos < n and s[pos] != "(":
                parse_label_and_length(cur)
        elif c == ")":
            pos += 1
            cur = cur.mother
            if cur is None:
                raise ValueError("unbalanced parentheses in Newick string")
            parse_label_and_length(cur)
        else:
            raise ValueError(f"unexpected character {c!r} at {pos} in Newick string")
    if cur is not root:
        raise ValueError("unbalanced parentheses in Newick string")
    root.mother = None
    root.root = True
    _name_unnamed(root)
    number_nodes(root)
    return root


def _name_unnamed(root: GeneralNode) -> None:
    k = 0
    for nd in post_order(root):
        if nd.name == "no_name":
            k += 1
            nd.name = f"no_name_{k}"


def newick(root: GeneralNode) -> str:
    def rec(nd: GeneralNode) -> str:
        if nd.nchild == 0:
            return f"{nd.name}:{nd.inc_length!r}"
        inner = ",".join(rec(c) for c in nd.children)
        if nd.root:
            return f"({inner})"
        return f"({inner}){nd.name}:{nd.inc_length!r}"

    import sys

    lim = sys.getrecursionlimit()
    sys.setrecursionlimit(max(lim, 10 * len(post_order(root)) + 100))
    try:
        return rec(root) + ";"
    finally:
        sys.setrecursionlimit(lim)


# ----------------------------------------------------------------------------------------
# topology moves used by the tree samplers (caller side; only NNI is needed to replay the
# PNUTS call pattern of /root/reference/src/samplers/tree_hamiltonian/refraction.jl:35-91)
# ----------------------------------------------------------------------------------------

def NNI(root: GeneralNode, target: GeneralNode, lor: bool = True) -> int:
    """Nearest-neighbour interchange across the branch above `target`: swap one child of
    `target` with its sibling.  Node numbers are kept.  Returns 1 if a swap was made."""
    if target.root or target.nchild < 2 or target.mother is None:
        return 0
    mother = target.mother
    sibs = [c for c in mother.children if c is not target]
    if not sibs:
        return 0
    sib = sibs[0]
    ch = target.children[0] if lor else target.children[-1]
    i_s = mother.children.index(sib)
    i_c = target.children.index(ch)
    mother.children[i_s] = ch
    target.children[i_c] = sib
    ch.mother = mother
    sib.mother = target
    return 1


# ----------------------------------------------------------------------------------------
# flattening for the C-ABI
# ----------------------------------------------------------------------------------------

@dataclass
class FlatTree:
    """The integer/double arrays `mcp_eval` takes (include/mcphylo_b200.h)."""

    NN: int
    postorder_num: np.ndarray  # int32[NN], 1-based nums in post_order order, root last
    parent_num: np.ndarray     # int32[NN], indexed by num-1, 0 for the root
    blv: np.ndarray            # float64[NN-1], indexed by num-1
    leaf_nums: np.ndarray      # int32[n_leaves], 1-based, in get_leaves order
    leaf_names: List[str] = field(default_factory=list)


def flatten(root: GeneralNode) -> FlatTree:
    po = post_order(root)
    NN = len(po)
    if po[-1].num != NN:
        raise ValueError("root must carry the largest node number (run number_nodes)")
    nums = [n.num for n in po]
    parents = [n.mother.num if n.mother is not None else 0 for n in po]
    lengths = [n.inc_length for n in po]
    postorder_num = np.array(nums, dtype=np.int32)
    idx = postorder_num - 1
    parent_num = np.zeros(NN, dtype=np.int32)
    parent_num[idx] = parents
    blv_full = np.zeros(NN, dtype=np.float64)
    blv_full[idx] = lengths
    leaves = [n for n in po if not n.children]
    return FlatTree(NN, postorder_num, parent_num, blv_full[:NN - 1].copy(),
                    np.array([n.num for n in leaves], dtype=np.int32), [n.name for n in leaves])
