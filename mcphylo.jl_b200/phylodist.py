"""PhyloDist / MultiplePhyloDist with logpdf and gradlogpdf — the drop-in boundary.

Mirrors /root/reference/src/distributions/Phylodist.jl: constructor forms :17-100, `size`
:105, `logpdf` :107-122, `gradlogpdf` :124-138 (returns the TUPLE (logL, gradient by
node.num)), MultiplePhyloDist :143-272, its `logpdf` :281-288 and `__logpdf` :290-297.
The reference re-expands `x` and recomputes everything on the CPU in every call; here the
call flattens the tree, evaluates the substitution model's eigendecomposition on the host
and hands both to the C-ABI (capi.py -> libmcphylo_b200.so).  The leaf data `x` is uploaded
once per array object and stays resident on the GPU (the reference passes the same Array for
the life of a chain, /root/reference/src/model/dependent.jl:344-358).  The cache is keyed by the
array OBJECT: an array modified in place after its first use must be re-registered with
`release_device_cache()`.

No CPU fallback exists: without the CUDA library or a GPU these functions raise.
"""
from __future__ import annotations

import weakref
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import capi
from .substitution_models import freeK, freeK_equilibrium
from .tree import GeneralNode, flatten, post_order


class DimensionMismatch(ValueError):
    pass


class DeviceAlignment:
    """Compact leaf data (uint8 codes) as an alternative to the dense array `x`: what a caller
    with a 1000 x 1M alignment passes instead of a 64 GB one-hot array."""

    def __init__(self, codes: np.ndarray, leaf_nums: Sequence[int], K: int):
        self.codes = np.ascontiguousarray(codes, dtype=np.uint8)
        self.leaf_nums = np.asarray(leaf_nums, dtype=np.int32)
        self.K = int(K)
        self._handles = {}

    @property
    def S(self) -> int:
        return self.codes.shape[1]

    def site_block(self, lo: int, hi: int) -> "DeviceAlignment":
        """Columns [lo, hi): the shard one rank owns when sites are split across GPUs."""
        return DeviceAlignment(self.codes[:, lo:hi], self.leaf_nums, self.K)


_contexts = {}
_dense_cache = {}


def get_context(device: Optional[int] = None) -> capi.Context:
    """Lazily created per-(process, device) context; never stored inside a distribution so
    that distributions stay serialisable (SURVEY.md §5 checkpoint/resume)."""
    if device is None:
        device = _default_device
    if isinstance(device, (list, tuple)):       # several GPUs behind one context (mcp_create_multi)
        device = tuple(int(d) for d in device)
    ctx = _contexts.get(device)
    if ctx is None:
        ctx = capi.Context(devices=device) if isinstance(device, tuple) else capi.Context(device)
        _contexts[device] = ctx
    return ctx


_default_device = 0


def set_default_device(device) -> None:
    """An int, or a list of device ids: every logpdf / gradlogpdf then shards the site axis over them."""
    global _default_device
    _default_device = tuple(int(d) for d in device) if isinstance(device, (list, tuple)) else int(device)


def release_device_cache() -> None:
    """Drop cached device alignments (and contexts)."""
    for aln in list(_dense_cache.values()):
        aln[1].close()
    _dense_cache.clear()
    for ctx in list(_contexts.values()):
        ctx.close()
    _contexts.clear()


def _device_alignment(x, leaf_nums: np.ndarray, K: int, ctx: capi.Context) -> capi.Alignment:
    if isinstance(x, DeviceAlignment):
        h = x._handles.get(id(ctx))        # an alignment belongs to the context that created it
        if h is None or h.handle is None or h.ctx is not ctx:
            if x.K != K:
                raise DimensionMismatch(f"alignment has {x.K} states, distribution has {K}")
            h = ctx.alignment_from_codes(x.codes, x.K, x.leaf_nums)
            x._handles[id(ctx)] = h
        return h
    x = np.asarray(x)
    if x.ndim != 3:
        raise DimensionMismatch("x must be a (K, S, NN) array")
    # Rows are mapped through leaf_nums when the alignment is created, so the ORDER in which a topology
    # lists its leaves is irrelevant (an NNI permutes get_leaves): only the set has to match.
    key = (id(x), id(ctx))
    hit = _dense_cache.get(key)
    if hit is not None and hit[0]() is x and hit[1].handle is not None and hit[1].ctx is ctx and \
            np.array_equal(hit[2], np.sort(leaf_nums)):
        return hit[1]
    if x.shape[0] != K:
        raise DimensionMismatch(f"x has {x.shape[0]} states, distribution has {K}")
    aln = ctx.alignment_from_dense(x, leaf_nums)
    try:
        ref = weakref.ref(x, lambda _r, k=key: _dense_cache.pop(k, None))
    except TypeError:
        ref = (lambda v: (lambda: v))(x)
    _dense_cache[key] = (ref, aln, np.sort(leaf_nums))
    return aln


class PhyloDist:
    """Distribution whose likelihood is computed by Felsenstein's algorithm.

    PhyloDist(tree, base_freq, substitution_rates, rates, substitution_model)
    PhyloDist(tree, substitution_rates, rates, freeK)            (equilibrium base_freq)
    Scalars are accepted for substitution_rates / rates; `tree` may be a node or anything
    with a `.value` node (the reference's TreeVariate)."""

    def __init__(self, tree, *args):
        tree = getattr(tree, "value", tree)
        if not isinstance(tree, GeneralNode):
            raise TypeError("tree must be a GeneralNode")
        if len(args) == 3 and args[2] is freeK:
            substitution_rates, rates, substitution_model = args
            base_freq = freeK_equilibrium(np.atleast_1d(np.asarray(substitution_rates, dtype=np.float64)))
        elif len(args) == 4:
            base_freq, substitution_rates, rates, substitution_model = args
        else:
            raise TypeError("PhyloDist(tree, base_freq, substitution_rates, rates, substitution_model)")
        if not callable(substitution_model):
            raise TypeError("substitution_model must be callable")
        self.tree = tree
        self.base_freq = np.atleast_1d(np.asarray(base_freq, dtype=np.float64)).copy()
        self.substitution_rates = np.atleast_1d(np.asarray(substitution_rates, dtype=np.float64)).copy()
        self.rates = np.atleast_1d(np.asarray(rates, dtype=np.float64)).copy()
        self.substitution_model: Callable = substitution_model
        self.nbase = int(self.base_freq.size)
        self.nnodes = len(post_order(tree))

    def __eq__(self, other):
        return (isinstance(other, PhyloDist) and self.tree is other.tree
                and np.array_equal(self.base_freq, other.base_freq)
                and np.array_equal(self.substitution_rates, other.substitution_rates)
                and np.array_equal(self.rates, other.rates)
                and self.substitution_model is other.substitution_model
                and self.nbase == other.nbase and self.nnodes == other.nnodes)

    __hash__ = None

    def size(self) -> Tuple[int, int, int]:
        return (self.nbase, 1, self.nnodes)


from . import prior as _prior  # noqa: E402


def minimum(d) -> float:
    return -np.inf


def maximum(d) -> float:
    return np.inf


def size(d):
    return d.size()


def _tree_args(d: PhyloDist):
    ft = flatten(d.tree)
    U, D, Uinv, mu = d.substitution_model(d.base_freq, d.substitution_rates)
    return ft, (ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, d.rates, d.base_freq)


def logpdf(d, x, device: Optional[int] = None) -> float:
    """log-likelihood; `x` is the reference's (K, S, NN) array or a DeviceAlignment.
    With a branch-length distribution and a tree it is the prior's log density (Prior.jl:68-83)."""
    if isinstance(d, _prior.LengthDistribution):
        return _prior.prior_logpdf(d, x)
    if isinstance(d, MultiplePhyloDist):
        return float(np.sum(_multi(d, x, False, device)[0]))
    ctx = get_context(device)
    ft, targs = _tree_args(d)
    aln = _device_alignment(x, ft.leaf_nums, d.nbase, ctx)
    ll, _ = ctx.eval(aln, *targs, want_grad=False)
    return ll


def gradlogpdf(d: PhyloDist, x, device: Optional[int] = None) -> Tuple[float, np.ndarray]:
    """(logL, d logL / d branch length indexed by node.num) — a tuple, like the reference."""
    if isinstance(d, _prior.LengthDistribution):
        return _prior.prior_gradlogpdf(d, x)
    ctx = get_context(device)
    ft, targs = _tree_args(d)
    aln = _device_alignment(x, ft.leaf_nums, d.nbase, ctx)
    return ctx.eval(aln, *targs, want_grad=True)


def gradlogpdf_rates(d: PhyloDist, x, device: Optional[int] = None) -> Tuple[float, np.ndarray, np.ndarray]:
    """(logL, d logL / d branch length, d logL / d d.rates[r]) -- the rate-category gradient falls out of the
    per-category branch gradients (mcp_eval_rate_gradient); with `rates.discrete_gamma_rates_dalpha` it gives
    the Gamma-shape gradient  d logL / d alpha = rate_grad . d rates / d alpha."""
    ctx = get_context(device)
    ft, targs = _tree_args(d)
    aln = _device_alignment(x, ft.leaf_nums, d.nbase, ctx)
    return ctx.eval_rate_gradient(aln, *targs)


def gradlogpdf_model(d: PhyloDist, x, device: Optional[int] = None, with_rates: bool = False):
    """(logL, d logL / d branch length, d logL / d base_freq, d logL / d substitution_rates): the gradient with respect
    to the substitution model's own parameters next to the branch gradient (mcp_eval_model_gradient; the reference
    samples these parameters gradient-free).  base_freq entries are independent coordinates --
    they enter through the root distribution and, for GTR / Restriction, through the rate matrix; project onto the
    simplex on the caller's side.  Works for any callable substitution_model (difference quotients of the K x K
    normalised rate matrix for functions other than Restriction / JC / GTR / freeK).  with_rates appends
    d logL / d d.rates[r], taken from the same moment matrices (what gradlogpdf_rates returns)."""
    from .substitution_models import model_derivatives

    ctx = get_context(device)
    ft, targs = _tree_args(d)
    aln = _device_alignment(x, ft.leaf_nums, d.nbase, ctx)
    _, dA, dpi = model_derivatives(d.substitution_model, d.base_freq, d.substitution_rates)
    ll, grad, pg, rg = ctx.eval_model_gradient(aln, *targs, dA=dA, dpi=dpi, want_rate_grad=True)
    K = d.nbase
    if with_rates:
        return ll, grad, pg[:K], pg[K:], rg
    return ll, grad, pg[:K], pg[K:]


class MultiplePhyloDist:
    """A collection of independent PhyloDists evaluated in one batched launch."""

    def __init__(self, tree_array: Sequence[GeneralNode], *args):
        trees = [getattr(t, "value", t) for t in tree_array]
        n_t = len(trees)
        if len(args) == 3 and args[2] is freeK:
            substitution_rates, rates, model = args
            sr = _columns(substitution_rates, n_t, "substitution_rates")
            bf = np.stack([freeK_equilibrium(sr[:, i]) for i in range(n_t)], axis=1)
            rt = _columns(rates, n_t, "rates")
        elif len(args) == 4:
            base_freq, substitution_rates, rates, model = args
            bf = _columns(base_freq, n_t, "base_freq")
            sr = _columns(substitution_rates, n_t, "substitution_rates")
            rt = _columns(rates, n_t, "rates")
        else:
            raise TypeError("MultiplePhyloDist(trees, base_freq, substitution_rates, rates, model)")
        self.DistCollector: List[PhyloDist] = [PhyloDist(t, bf[:, i], sr[:, i], rt[:, i], model)
                                               for i, t in enumerate(trees)]
        self.size_array = np.asarray([d.nnodes for d in self.DistCollector], dtype=np.int64)

    def __eq__(self, other):
        return (isinstance(other, MultiplePhyloDist) and len(self.DistCollector) == len(other.DistCollector)
                and all(a == b for a, b in zip(self.DistCollector, other.DistCollector)))

    __hash__ = None

    def size(self):
        return (self.DistCollector[0].base_freq.shape[0], 1, int(self.size_array.max()), len(self.size_array))


def _columns(a, n_t: int, what: str) -> np.ndarray:
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 0:
        a = a.reshape(1, 1)
    if a.ndim == 1:
        a = a[:, None]
    if a.shape[1] == n_t:
        return a.copy()
    if a.shape[1] == 1:
        return np.repeat(a, n_t, axis=1)
    raise DimensionMismatch(f"Size of {what} and tree_array are incompatible")


def _multi(d: MultiplePhyloDist, x, want_grad: bool, device: Optional[int]):
    ctx = get_context(device)
    alns, targs = [], []
    if isinstance(x, (list, tuple)):
        xs = list(x)
    else:
        x = np.asarray(x)
        if x.ndim != 4:
            raise DimensionMismatch("x must be a (K, S, maxNN, T) array")
        # one resident slab per tree; the views keep `x` alive and are cached per tree index
        cache = _slab_cache.setdefault(id(x), (weakref.ref(x, lambda _r, k=id(x): _slab_cache.pop(k, None)), {}))[1]
        xs = []
        for ind, s in enumerate(d.size_array):
            v = cache.get((ind, int(s)))
            if v is None:
                v = np.asfortranarray(x[:, :, :int(s), ind])
                cache[(ind, int(s))] = v
            xs.append(v)
    if len(xs) != len(d.DistCollector):
        raise DimensionMismatch("number of data slabs and trees differ")
    for pd, xt in zip(d.DistCollector, xs):
        ft, ta = _tree_args(pd)
        alns.append(_device_alignment(xt, ft.leaf_nums, pd.nbase, ctx))
        targs.append(ta)
    return ctx.eval_batch(alns, targs, want_grad=want_grad)


_slab_cache = {}


def __logpdf(d: MultiplePhyloDist, x, device: Optional[int] = None):
    """Per-tree (logL, gradient) tuples, like the reference's `__logpdf`."""
    ll, grads = _multi(d, x, True, device)
    return [(float(l), g) for l, g in zip(ll, grads)]


multi_gradlogpdf = __logpdf
