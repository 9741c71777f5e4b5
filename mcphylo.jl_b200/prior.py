"""Branch-length priors of the tree node and the fused "likelihood + prior" call.

Mirrors, under the reference's names,
  exponentialBL, CompoundDirichlet, UniformBranchLength   /root/reference/src/distributions/TreeDistribution.jl:8-50
  internal_logpdf, logpdf, gradlogpdf, insupport          /root/reference/src/Likelihood/Prior.jl:1-92
  logpdfgrad!(::Type{provided}, ...)                      /root/reference/src/samplers/sampler.jl:172-190

The reference differentiates `internal_logpdf` with Zygote on every leapfrog (Prior.jl:39-57); the
derivative is written out here.  `logpdf` / `gradlogpdf` of a prior alone are O(NN) host scalar
work (like the K x K eigendecomposition, they are caller-side logic, not the hot path);
`logpdfgrad` is the sampler's combined call and runs on the GPU: the prior is folded into the
final reduction of the likelihood kernel (mcp_eval_posterior, include/mcphylo_b200.h).
`internal_external` comes from the un-vendored MCPhyloTree.jl; it is restated from its use in
Prior.jl:15-23 (1 = the branch leads to an internal node, 0 = to a leaf, indexed like
get_branchlength_vector) and pinned by the golden value in test/distributions/treedists.jl:61-69.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np

from .tree import GeneralNode, get_branchlength_vector, post_order

MCP_PRIOR_NONE, MCP_PRIOR_EXPONENTIAL, MCP_PRIOR_COMPOUND_DIRICHLET = 0, 1, 2


class exponentialBL:  # noqa: N801 - reference name
    """i.i.d. Exponential(scale) branch lengths (TreeDistribution.jl:8-14)."""

    def __init__(self, scale: float, constraints=None):
        self.scale = float(scale)
        self.constraints = constraints


class CompoundDirichlet:
    """Compound Dirichlet prior of Zhang, Rannala & Yang 2012 (TreeDistribution.jl:23-39)."""

    def __init__(self, alpha: float, a: float, beta: float, c: float, constraints=None):
        self.alpha, self.a, self.beta, self.c = float(alpha), float(a), float(beta), float(c)
        self.constraints = constraints


class UniformBranchLength:
    """Improper flat prior used when no length distribution is given (TreeDistribution.jl:48)."""


LengthDistribution = (CompoundDirichlet, exponentialBL, UniformBranchLength)


def internal_external(root: GeneralNode) -> np.ndarray:
    """int64[NN-1], entry num-1 = 1 if the node below branch `num` has children, else 0."""
    po = post_order(root)
    out = np.zeros(len(po), dtype=np.int64)
    for n in po:
        if n.children:
            out[n.num - 1] = 1
    return out[:len(po) - 1].copy()


def internal_logpdf(d: CompoundDirichlet, b_lens, int_leave_map) -> float:
    """Prior.jl:1-37, same grouping of terms."""
    b = np.asarray(b_lens, dtype=np.float64)
    m = np.asarray(int_leave_map) == 1
    with np.errstate(divide="ignore", invalid="ignore"):
        logs = np.log(b)
    blen_int, blen_int_log = float(b[m].sum()), float(logs[m].sum())
    blen_leave, blen_leave_log = float(b[~m].sum()), float(logs[~m].sum())
    nterm = float((~m).sum())
    t_l = blen_int + blen_leave
    n_int = nterm - 3.0
    first = d.alpha * math.log(d.beta) - math.lgamma(d.alpha) - t_l * d.beta
    second = -math.lgamma(d.a) - math.lgamma(d.c) + math.lgamma(d.a + d.c)
    third = blen_leave_log * (d.a - 1.0) + blen_int_log * (d.a * d.c - 1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        fourth = (d.alpha - d.a * nterm - d.a * d.c * n_int) * float(np.log(np.float64(t_l)))
    return first + second + third + fourth


def prior_logpdf(d, x: GeneralNode) -> float:
    if isinstance(d, CompoundDirichlet):
        return internal_logpdf(d, get_branchlength_vector(x), internal_external(x))
    if isinstance(d, exponentialBL):
        bl = get_branchlength_vector(x)
        if np.any(bl < 0.0):
            return -math.inf            # Distributions.logpdf(Exponential, x<0) = -Inf
        return float(np.sum(-math.log(d.scale) - bl / d.scale))
    if isinstance(d, UniformBranchLength):
        return 0.0
    raise TypeError(f"not a branch-length distribution: {type(d).__name__}")


def prior_gradlogpdf(d, x: GeneralNode) -> Tuple[float, np.ndarray]:
    """(log density, d/d blv) — what `withgradient` returns in Prior.jl:39-66."""
    blv = get_branchlength_vector(x)
    if isinstance(d, CompoundDirichlet):
        ie = internal_external(x)
        nterm = float((ie == 0).sum())
        k4 = d.alpha - d.a * nterm - d.a * d.c * (nterm - 3.0)
        w = np.where(ie == 1, d.a * d.c - 1.0, d.a - 1.0)
        with np.errstate(divide="ignore", invalid="ignore"):
            g = -d.beta + np.where(w != 0.0, w / blv, 0.0) + k4 / blv.sum()
        return internal_logpdf(d, blv, ie), g
    if isinstance(d, exponentialBL):
        return prior_logpdf(d, x), np.full(blv.size, -1.0 / d.scale)
    if isinstance(d, UniformBranchLength):
        return 0.0, np.zeros(blv.size)
    raise TypeError(f"not a branch-length distribution: {type(d).__name__}")


def insupport(d, x: GeneralNode) -> bool:
    """Prior.jl:89-92."""
    bl = get_branchlength_vector(x)
    return bool(np.all(np.isfinite(bl)) and np.all(bl > 0.0))


def prior_spec(d) -> Tuple[int, list]:
    """(prior_kind, prior_params) of mcp_eval_posterior."""
    if d is None or isinstance(d, UniformBranchLength):
        return MCP_PRIOR_NONE, []
    if isinstance(d, exponentialBL):
        return MCP_PRIOR_EXPONENTIAL, [d.scale]
    if isinstance(d, CompoundDirichlet):
        return MCP_PRIOR_COMPOUND_DIRICHLET, [d.alpha, d.a, d.beta, d.c]
    raise TypeError(f"not a branch-length distribution: {type(d).__name__}")


def logpdfgrad(d, x, prior=None, device: Optional[int] = None, want_grad: bool = True):
    """logL(PhyloDist d | x) + log prior(d.tree) and the summed branch-length gradient in ONE device
    call — the body of logpdfgrad!(::Type{provided}) (sampler.jl:172-190)."""
    from .phylodist import _device_alignment, _tree_args, get_context

    ctx = get_context(device)
    ft, targs = _tree_args(d)
    aln = _device_alignment(x, ft.leaf_nums, d.nbase, ctx)
    kind, params = prior_spec(prior)
    return ctx.eval_posterior(aln, *targs, prior_kind=kind, prior_params=params, want_grad=want_grad)
