"""ctypes binding of libmcphylo_b200.so (include/mcphylo_b200.h) — the same calls the Julia
glue makes with `ccall` (julia/MCPhyloB200.jl).

No fallback of any kind: a missing library raises at load time, a missing GPU raises at
`Context()` time with the library's own message.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCPHYLO_B200_LIB") or os.path.join(_HERE, "lib", "libmcphylo_b200.so")
_lib = None

SYMBOLS = [
    "mcp_abi_version", "mcp_last_error", "mcp_create", "mcp_destroy", "mcp_set_stream",
    "mcp_use_own_stream", "mcp_synchronize",
    "mcp_create_multi", "mcp_device_count", "mcp_reduce_mode", "mcp_shard_bounds", "mcp_nccl_unique_id",
    "mcp_create_rank", "mcp_eval_streamed", "mcp_stream_blocks", "mcp_stream_timeline", "mcp_host_register", "mcp_host_unregister",
    "mcp_get_stats_member", "mcp_timer_start", "mcp_timer_stop",
    "mcp_alignment_from_codes", "mcp_alignment_from_dense", "mcp_alignment_update_codes",
    "mcp_alignment_destroy",
    "mcp_eval", "mcp_eval_posterior", "mcp_eval_rate_gradient", "mcp_eval_model_gradient", "mcp_model_gradient_contract", "mcp_eval_device", "mcp_eval_batch", "mcp_get_stats", "mcp_wave_columns", "mcp_set_launch",
    "mcp_set_columns_per_thread", "mcp_set_timing", "mcp_set_scratch_mode", "mcp_set_accumulator_mode", "mcp_set_level_mode", "mcp_set_tile_order", "mcp_set_cherry_mode", "mcp_set_large_alphabet_mode", "mcp_set_ring_mode",
    "mcp_schedule_dump", "mcp_schedule_fetch_list", "mcp_model_reorder",
]


class McpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libmcphylo_b200 error {code}: {msg}")
        self.code = code


class Stats(C.Structure):
    _fields_ = [("walk_ms", C.c_double), ("device_ms", C.c_double), ("h2d_bytes", C.c_int64),
                ("d2h_bytes", C.c_int64), ("kernel_launches", C.c_int32), ("grid", C.c_int32),
                ("block", C.c_int32), ("tiles", C.c_int32), ("schedule_rebuilt", C.c_int32),
                ("scratch_bytes", C.c_int64), ("columns_per_thread", C.c_int32), ("operand_ring", C.c_int32)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_vp = C.c_void_p
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def load():
    """Load the shared library; raises if it has not been built (python -c 'import
    __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library has not been built. There is no CPU "
            "fallback. Build it with `python -c \"import __graft_entry__ as g; g.build()\"`.")
    lib = C.CDLL(LIB_PATH)
    lib.mcp_abi_version.restype = C.c_int
    lib.mcp_last_error.restype = C.c_char_p
    lib.mcp_last_error.argtypes = [_vp]
    lib.mcp_create.argtypes = [C.POINTER(_vp), C.c_int]
    lib.mcp_destroy.argtypes = [_vp]
    lib.mcp_set_stream.argtypes = [_vp, _vp]
    lib.mcp_use_own_stream.argtypes = [_vp]
    lib.mcp_synchronize.argtypes = [_vp]
    lib.mcp_set_launch.argtypes = [_vp, C.c_int, C.c_int]
    lib.mcp_set_columns_per_thread.argtypes = [_vp, C.c_int]
    lib.mcp_set_scratch_mode.argtypes = [_vp, C.c_int]
    lib.mcp_set_accumulator_mode.argtypes = [_vp, C.c_int]
    lib.mcp_set_level_mode.argtypes = [_vp, C.c_int]
    lib.mcp_set_tile_order.argtypes = [_vp, C.c_int]
    lib.mcp_set_cherry_mode.argtypes = [_vp, C.c_int]
    lib.mcp_set_large_alphabet_mode.argtypes = [_vp, C.c_int]
    lib.mcp_set_ring_mode.argtypes = [_vp, C.c_int]
    lib.mcp_alignment_from_codes.argtypes = [_vp, _vp, C.c_int, C.c_int64, _vp, C.c_int, C.POINTER(_vp)]
    lib.mcp_alignment_from_dense.argtypes = [_vp, _vp, C.c_int, C.c_int64, C.c_int, _vp, C.c_int, C.POINTER(_vp)]
    lib.mcp_alignment_destroy.argtypes = [_vp, _vp]
    lib.mcp_alignment_update_codes.argtypes = [_vp, _vp, _vp]
    eval_args = [_vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, C.c_double, _vp, C.c_int, _vp, C.c_int]
    lib.mcp_eval.argtypes = eval_args + [_dp, _vp]
    lib.mcp_eval_device.argtypes = eval_args + [_vp]
    lib.mcp_eval_rate_gradient.argtypes = eval_args[:-1] + [_dp, _vp, _vp]
    lib.mcp_eval_posterior.argtypes = eval_args[:-1] + [C.c_int, _vp, _dp, _vp]
    lib.mcp_eval_model_gradient.argtypes = eval_args[:-1] + [C.c_int, _vp, _vp, _dp, _vp, _vp, _vp, _vp]
    lib.mcp_model_gradient_contract.argtypes = [C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, C.c_double, _vp, _vp, _vp,
                                                C.c_int, _vp, _vp, _vp, _vp, _vp]
    lib.mcp_eval_batch.argtypes = [_vp, C.c_int] + [_vp] * 10 + [C.c_int, _vp, C.c_int, _vp, _vp]
    lib.mcp_get_stats.argtypes = [_vp, C.POINTER(Stats)]
    lib.mcp_get_stats_member.argtypes = [_vp, C.c_int, C.POINTER(Stats)]
    lib.mcp_timer_start.argtypes = [_vp]
    lib.mcp_timer_stop.argtypes = [_vp, _dp]
    lib.mcp_create_multi.argtypes = [C.POINTER(_vp), C.c_int, _vp, C.c_int]
    lib.mcp_device_count.argtypes = [_vp]
    lib.mcp_reduce_mode.argtypes = [_vp]
    lib.mcp_shard_bounds.argtypes = [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.mcp_nccl_unique_id.argtypes = [_vp]
    lib.mcp_create_rank.argtypes = [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, _vp]
    lib.mcp_eval_streamed.argtypes = [_vp, _vp, C.c_int, C.c_int64, _vp, C.c_int] + eval_args[2:] + [_dp, _vp]
    lib.mcp_stream_blocks.argtypes = [_vp, C.c_int, _vp, C.c_int]
    lib.mcp_stream_timeline.argtypes = [_vp, C.c_int, _vp, C.c_int]
    lib.mcp_host_register.argtypes = [_vp, C.c_size_t]
    lib.mcp_host_unregister.argtypes = [_vp]
    lib.mcp_wave_columns.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64)]
    lib.mcp_schedule_dump.argtypes = [C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp]
    lib.mcp_model_reorder.argtypes = [C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(C.c_int)]
    for name in SYMBOLS:
        if name != "mcp_last_error":
            getattr(lib, name).restype = C.c_int
    if lib.mcp_abi_version() != 2:
        raise ImportError("libmcphylo_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def model_gradient_contract(blv, U, D, Uinv, mu, rates, moments, root_w, dA, dpi=None, want_grad_check=False,
                            want_rate_grad=False):
    """Host-only second half of mcp_eval_model_gradient (no GPU): moments M[b, r, s, k] and root vector W[s] ->
    d logL / d theta_p; optionally the branch gradient re-derived from the same moments."""
    lib = load()
    blv = np.ascontiguousarray(blv, dtype=np.float64)
    rates = np.ascontiguousarray(rates, dtype=np.float64)
    D = np.ascontiguousarray(D, dtype=np.float64)
    K, R, NB = D.size, rates.size, blv.size
    Uf = np.ascontiguousarray(np.asarray(U, dtype=np.float64).ravel(order="F"))
    Uif = np.ascontiguousarray(np.asarray(Uinv, dtype=np.float64).ravel(order="F"))
    M = np.ascontiguousarray(moments, dtype=np.float64).reshape(NB, R, K, K)
    W = np.ascontiguousarray(root_w, dtype=np.float64) if root_w is not None else None
    dA = np.asarray(dA, dtype=np.float64)
    if dA.ndim == 2:
        dA = dA[:, :, None]
    n_par = dA.shape[2]
    dA_f = np.ascontiguousarray(np.stack([dA[:, :, p].ravel(order="F") for p in range(n_par)]).ravel()) if n_par else np.zeros(1)
    dpi_f = np.ascontiguousarray(np.asarray(dpi, dtype=np.float64).reshape(K, n_par).ravel(order="F")) if dpi is not None else None
    pg = np.zeros(max(n_par, 1))
    gc = np.zeros(max(NB, 1)) if want_grad_check else None
    rg = np.zeros(R) if want_rate_grad else None
    rc = lib.mcp_model_gradient_contract(K, R, NB, blv.ctypes.data, Uf.ctypes.data, D.ctypes.data, Uif.ctypes.data, float(mu),
                                         rates.ctypes.data, M.ctypes.data, W.ctypes.data if W is not None else None, n_par,
                                         dA_f.ctypes.data, dpi_f.ctypes.data if dpi_f is not None else None, pg.ctypes.data,
                                         gc.ctypes.data if want_grad_check else None, rg.ctypes.data if want_rate_grad else None)
    if rc:
        raise McpError(rc, lib.mcp_last_error(None).decode())
    out = (pg[:n_par],)
    if want_grad_check:
        out += (gc[:NB],)
    if want_rate_grad:
        out += (rg,)
    return out if len(out) > 1 else out[0]


def _f64(a, order="C"):
    return np.require(np.asarray(a, dtype=np.float64), requirements=["C"] if order == "C" else ["F"])


def _colmajor(a):
    """Flat column-major copy of a matrix (Julia layout)."""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order="F"))


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


class Alignment:
    """Leaf data resident on the GPU (mcp_alignment)."""

    def __init__(self, ctx: "Context", handle, K: int, S: int, leaf_nums: np.ndarray):
        self.ctx, self.handle, self.K, self.S = ctx, handle, K, S
        self.leaf_nums = leaf_nums

    def update_codes(self, host_ptr: int):
        """Re-upload (n_leaves, S) uint8 codes from host memory at `host_ptr` (async on the
        context's stream; pinned memory recommended)."""
        self.ctx._check(load().mcp_alignment_update_codes(self.ctx.handle, self.handle, _vp(host_ptr)))

    def close(self):
        if self.handle is not None and self.ctx.handle is not None:
            load().mcp_alignment_destroy(self.ctx.handle, self.handle)
        self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


REDUCE_AUTO, REDUCE_NCCL, REDUCE_PEER, REDUCE_HOST = 0, 1, 2, 3
REDUCE_NAMES = {REDUCE_AUTO: "auto", REDUCE_NCCL: "nccl", REDUCE_PEER: "peer", REDUCE_HOST: "host"}


def shard_bounds(S: int, n_shards: int, shard: int):
    """The site range [lo, hi) a shard owns (mcp_shard_bounds; host-only)."""
    lo, hi = C.c_int64(), C.c_int64()
    rc = load().mcp_shard_bounds(int(S), int(n_shards), int(shard), C.byref(lo), C.byref(hi))
    if rc:
        raise McpError(rc, load().mcp_last_error(None).decode())
    return lo.value, hi.value


def nccl_unique_id() -> bytes:
    """128-byte NCCL unique id for mcp_create_rank (rank 0 makes it, the launcher's channel shares it)."""
    buf = C.create_string_buffer(128)
    rc = load().mcp_nccl_unique_id(buf)
    if rc:
        raise McpError(rc, load().mcp_last_error(None).decode())
    return buf.raw


class Context:
    """One mcp_ctx: a (process, GPU) pair; `devices=[...]` makes it a (process, set of GPUs) context
    (mcp_create_multi: site-sharded alignments, one all-reduce of [logL, grad] per evaluation), and
    `rank=(n_ranks, rank, unique_id)` one rank of a multi-process group (mcp_create_rank)."""

    def __init__(self, device: int = 0, devices: Optional[Sequence[int]] = None, reduce: int = REDUCE_AUTO,
                 rank: Optional[tuple] = None):
        self.lib = load()
        h = _vp()
        if devices is not None:
            ids = _i32(list(devices))
            rc = self.lib.mcp_create_multi(C.byref(h), int(ids.size), ids.ctypes.data, int(reduce))
            device = int(ids[0])
        elif rank is not None:
            n_ranks, r, uid = rank
            assert len(uid) == 128
            rc = self.lib.mcp_create_rank(C.byref(h), int(device), int(n_ranks), int(r), C.c_char_p(uid))
        else:
            rc = self.lib.mcp_create(C.byref(h), int(device))
        if rc:
            raise McpError(rc, self.lib.mcp_last_error(None).decode())
        self.handle = h
        self.device = device

    @property
    def device_count(self) -> int:
        return int(self.lib.mcp_device_count(self.handle))

    @property
    def reduce_mode(self) -> int:
        return int(self.lib.mcp_reduce_mode(self.handle))

    def _check(self, rc: int):
        if rc:
            raise McpError(rc, self.lib.mcp_last_error(self.handle).decode())

    def close(self):
        if self.handle is not None:
            self.lib.mcp_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: Optional[int]):
        """Run on the given cudaStream_t handle (0 = CUDA default stream); None returns to the
        context's own stream."""
        if cuda_stream is None:
            self._check(self.lib.mcp_use_own_stream(self.handle))
        else:
            self._check(self.lib.mcp_set_stream(self.handle, _vp(int(cuda_stream))))

    def synchronize(self):
        self._check(self.lib.mcp_synchronize(self.handle))

    def set_launch(self, block: int = 0, ctas_per_sm: int = 0):
        self._check(self.lib.mcp_set_launch(self.handle, int(block), int(ctas_per_sm)))

    def set_columns_per_thread(self, cpt: int = 0):
        self._check(self.lib.mcp_set_columns_per_thread(self.handle, int(cpt)))

    def set_timing(self, on: bool = True):
        """Per-evaluation CUDA timing events behind stats() on / off (mcp_set_timing)."""
        self._check(self.lib.mcp_set_timing(self.handle, int(bool(on))))

    def set_level_mode(self, mode: int = -1):
        self._check(self.lib.mcp_set_level_mode(self.handle, int(mode)))

    def set_ring_mode(self, mode: int = -1):
        self._check(self.lib.mcp_set_ring_mode(self.handle, int(mode)))

    def set_large_alphabet_mode(self, mode: int = -1):
        self._check(self.lib.mcp_set_large_alphabet_mode(self.handle, int(mode)))

    def set_cherry_mode(self, mode: int = -1):
        self._check(self.lib.mcp_set_cherry_mode(self.handle, int(mode)))

    def set_tile_order(self, mode: int = -1):
        self._check(self.lib.mcp_set_tile_order(self.handle, int(mode)))

    def set_scratch_mode(self, mode: int = -1):
        self._check(self.lib.mcp_set_scratch_mode(self.handle, int(mode)))

    def set_accumulator_mode(self, mode: int = -1):
        self._check(self.lib.mcp_set_accumulator_mode(self.handle, int(mode)))

    def wave_columns(self, K: int, n_nodes: int, want_grad: bool = True) -> int:
        """Columns (sites x rates) one full wave of the persistent grid covers (mcp_wave_columns)."""
        out = C.c_int64()
        self._check(self.lib.mcp_wave_columns(self.handle, int(K), int(n_nodes), int(want_grad), C.byref(out)))
        return out.value

    def stats(self, member: Optional[int] = None) -> dict:
        s = Stats()
        if member is None:
            self._check(self.lib.mcp_get_stats(self.handle, C.byref(s)))
        else:
            self._check(self.lib.mcp_get_stats_member(self.handle, int(member), C.byref(s)))
        return s.asdict()

    def stream_timeline(self, member: int = 0):
        """Per block of the last mcp_eval_streamed call: ms since the first transfer began of
        [transfer begin, transfer end, evaluation enqueued, walk begin, walk end]."""
        buf = np.zeros((16, 5), dtype=np.float64)
        n = self.lib.mcp_stream_timeline(self.handle, int(member), buf.ctypes.data, 16)
        return buf[:max(n, 0)].round(3).tolist()

    def timer_start(self):
        self._check(self.lib.mcp_timer_start(self.handle))

    def timer_stop(self) -> float:
        """Milliseconds of device time since timer_start (CUDA events on the context's streams)."""
        ms = C.c_double()
        self._check(self.lib.mcp_timer_stop(self.handle, C.byref(ms)))
        return ms.value

    def stream_blocks(self, member: int = 0):
        """Site blocks [lo, hi) one device used in the last mcp_eval_streamed call."""
        buf = np.zeros((16, 2), dtype=np.int64)
        n = self.lib.mcp_stream_blocks(self.handle, int(member), buf.ctypes.data, 16)
        return [(int(a), int(b)) for a, b in buf[:max(n, 0)]]

    # ---- leaf data -------------------------------------------------------------------
    def alignment_from_codes(self, codes, K: int, leaf_nums) -> Alignment:
        codes = np.ascontiguousarray(np.asarray(codes, dtype=np.uint8))
        leaf_nums = _i32(leaf_nums)
        n_leaves, S = codes.shape
        assert leaf_nums.size == n_leaves
        h = _vp()
        self._check(self.lib.mcp_alignment_from_codes(self.handle, codes.ctypes.data, int(K), int(S),
                                                      leaf_nums.ctypes.data, n_leaves, C.byref(h)))
        return Alignment(self, h, int(K), int(S), leaf_nums)

    def alignment_from_dense(self, x, leaf_nums) -> Alignment:
        x = np.asarray(x, dtype=np.float64)
        K, S, NN = x.shape
        xf = x if x.flags.f_contiguous else np.asfortranarray(x)
        leaf_nums = _i32(leaf_nums)
        h = _vp()
        self._check(self.lib.mcp_alignment_from_dense(self.handle, xf.ctypes.data, K, S, NN,
                                                      leaf_nums.ctypes.data, leaf_nums.size, C.byref(h)))
        return Alignment(self, h, K, S, leaf_nums)

    # ---- evaluation ------------------------------------------------------------------
    @staticmethod
    def _pack(postorder_num, parent_num, blv, U, D, Uinv, rates, pi):
        return (_i32(postorder_num), _i32(parent_num), _f64(blv), _colmajor(U), _f64(D), _colmajor(Uinv),
                _f64(rates), _f64(pi))

    def eval(self, aln: Alignment, postorder_num, parent_num, blv, U, D, Uinv, mu, rates, pi,
             want_grad: bool = True):
        po, pa, blv, U, D, Uinv, rates, pi = self._pack(postorder_num, parent_num, blv, U, D, Uinv, rates, pi)
        NN = po.size
        assert pa.size == NN and blv.size == NN - 1 and D.size == aln.K and pi.size == aln.K
        ll = C.c_double()
        grad = np.zeros(max(NN - 1, 1), dtype=np.float64) if want_grad else None
        self._check(self.lib.mcp_eval(self.handle, aln.handle, NN, po.ctypes.data, pa.ctypes.data, blv.ctypes.data,
                                      U.ctypes.data, D.ctypes.data, Uinv.ctypes.data, float(mu),
                                      rates.ctypes.data, rates.size, pi.ctypes.data, int(want_grad),
                                      C.byref(ll), grad.ctypes.data if want_grad else None))
        return ll.value, (grad[:NN - 1] if want_grad else None)

    def eval_rate_gradient(self, aln: Alignment, postorder_num, parent_num, blv, U, D, Uinv, mu, rates, pi):
        """(logL, d logL / d blv, d logL / d rates) in one call (mcp_eval_rate_gradient)."""
        po, pa, blv, U, D, Uinv, rates, pi = self._pack(postorder_num, parent_num, blv, U, D, Uinv, rates, pi)
        NN = po.size
        ll = C.c_double()
        grad = np.zeros(max(NN - 1, 1), dtype=np.float64)
        rgrad = np.zeros(rates.size, dtype=np.float64)
        self._check(self.lib.mcp_eval_rate_gradient(self.handle, aln.handle, NN, po.ctypes.data, pa.ctypes.data,
                                                    blv.ctypes.data, U.ctypes.data, D.ctypes.data, Uinv.ctypes.data,
                                                    float(mu), rates.ctypes.data, rates.size, pi.ctypes.data,
                                                    C.byref(ll), grad.ctypes.data, rgrad.ctypes.data))
        return ll.value, grad[:NN - 1], rgrad

    def eval_model_gradient(self, aln: Alignment, postorder_num, parent_num, blv, U, D, Uinv, mu, rates, pi, dA, dpi=None,
                            want_moments: bool = False, want_rate_grad: bool = False):
        """(logL, d logL / d blv, d logL / d theta) for the substitution-model parameters whose normalised-rate-matrix
        derivatives are dA[:, :, p] (and root-frequency derivatives dpi[:, p]); with want_rate_grad also d logL / d rates[r]
        (appended 4th), with want_moments also the moment matrices M[b, r, s, k] and the root vector W[s] (appended
        last) (mcp_eval_model_gradient)."""
        po, pa, blv, U, D, Uinv, rates, pi = self._pack(postorder_num, parent_num, blv, U, D, Uinv, rates, pi)
        NN, K, R = po.size, aln.K, rates.size
        dA = np.asarray(dA, dtype=np.float64)
        if dA.ndim == 2:
            dA = dA[:, :, None]
        n_par = dA.shape[2]
        assert dA.shape[:2] == (K, K)
        dA_f = np.ascontiguousarray(np.stack([dA[:, :, p].ravel(order="F") for p in range(n_par)]).ravel()) if n_par else np.zeros(1)
        dpi_f = None
        if dpi is not None:
            dpi = np.asarray(dpi, dtype=np.float64).reshape(K, n_par)
            dpi_f = np.ascontiguousarray(dpi.ravel(order="F"))
        ll = C.c_double()
        grad = np.zeros(max(NN - 1, 1), dtype=np.float64)
        pgrad = np.zeros(max(n_par, 1), dtype=np.float64)
        mom = np.zeros((NN - 1) * R * K * K + K, dtype=np.float64) if want_moments else None
        rgrad = np.zeros(R, dtype=np.float64) if want_rate_grad else None
        self._check(self.lib.mcp_eval_model_gradient(self.handle, aln.handle, NN, po.ctypes.data, pa.ctypes.data,
                                                     blv.ctypes.data, U.ctypes.data, D.ctypes.data, Uinv.ctypes.data,
                                                     float(mu), rates.ctypes.data, R, pi.ctypes.data, n_par,
                                                     dA_f.ctypes.data, dpi_f.ctypes.data if dpi_f is not None else None,
                                                     C.byref(ll), grad.ctypes.data, pgrad.ctypes.data,
                                                     rgrad.ctypes.data if want_rate_grad else None,
                                                     mom.ctypes.data if want_moments else None))
        out = (ll.value, grad[:NN - 1], pgrad[:n_par])
        if want_rate_grad:
            out += (rgrad,)
        if want_moments:
            n_m = (NN - 1) * R * K * K
            out += (mom[:n_m].reshape(NN - 1, R, K, K), mom[n_m:])
        return out

    def eval_posterior(self, aln: Alignment, postorder_num, parent_num, blv, U, D, Uinv, mu, rates, pi,
                       prior_kind: int, prior_params, want_grad: bool = True):
        """logL + branch-length prior (and the summed gradient) in one call; the prior is added in the
        final reduction on the device (mcp_eval_posterior)."""
        po, pa, blv, U, D, Uinv, rates, pi = self._pack(postorder_num, parent_num, blv, U, D, Uinv, rates, pi)
        NN = po.size
        assert pa.size == NN and blv.size == NN - 1 and D.size == aln.K and pi.size == aln.K
        pp = _f64(list(prior_params) + [0.0] * 4)[:4].copy()
        lp = C.c_double()
        grad = np.zeros(max(NN - 1, 1), dtype=np.float64) if want_grad else None
        self._check(self.lib.mcp_eval_posterior(self.handle, aln.handle, NN, po.ctypes.data, pa.ctypes.data,
                                                blv.ctypes.data, U.ctypes.data, D.ctypes.data, Uinv.ctypes.data,
                                                float(mu), rates.ctypes.data, rates.size, pi.ctypes.data,
                                                int(prior_kind), pp.ctypes.data, C.byref(lp),
                                                grad.ctypes.data if want_grad else None))
        return lp.value, (grad[:NN - 1] if want_grad else None)

    def eval_streamed(self, codes_ptr: int, K: int, S: int, leaf_nums, postorder_num, parent_num, blv, U, D, Uinv, mu,
                      rates, pi, want_grad: bool = True):
        """One evaluation of a HOST-resident (n_leaves, S) uint8 alignment at `codes_ptr` (pinned memory
        recommended), uploaded block by block under the evaluation (mcp_eval_streamed)."""
        po, pa, blv, U, D, Uinv, rates, pi = self._pack(postorder_num, parent_num, blv, U, D, Uinv, rates, pi)
        leaf_nums = _i32(leaf_nums)
        NN = po.size
        ll = C.c_double()
        grad = np.zeros(max(NN - 1, 1), dtype=np.float64) if want_grad else None
        self._check(self.lib.mcp_eval_streamed(self.handle, _vp(codes_ptr), int(K), int(S), leaf_nums.ctypes.data,
                                               leaf_nums.size, NN, po.ctypes.data, pa.ctypes.data, blv.ctypes.data,
                                               U.ctypes.data, D.ctypes.data, Uinv.ctypes.data, float(mu),
                                               rates.ctypes.data, rates.size, pi.ctypes.data, int(want_grad),
                                               C.byref(ll), grad.ctypes.data if want_grad else None))
        return ll.value, (grad[:NN - 1] if want_grad else None)

    def eval_device(self, aln: Alignment, postorder_num, parent_num, blv, U, D, Uinv, mu, rates, pi,
                    want_grad: bool, d_out_ptr: int):
        """Result stays on the device at d_out_ptr (NN doubles); not synchronised."""
        po, pa, blv, U, D, Uinv, rates, pi = self._pack(postorder_num, parent_num, blv, U, D, Uinv, rates, pi)
        NN = po.size
        self._check(self.lib.mcp_eval_device(self.handle, aln.handle, NN, po.ctypes.data, pa.ctypes.data,
                                             blv.ctypes.data, U.ctypes.data, D.ctypes.data, Uinv.ctypes.data,
                                             float(mu), rates.ctypes.data, rates.size, pi.ctypes.data,
                                             int(want_grad), _vp(d_out_ptr)))

    def eval_batch(self, alns: Sequence[Alignment], trees: Sequence[tuple], want_grad: bool = True):
        """trees[t] = (postorder_num, parent_num, blv, U, D, Uinv, mu, rates, pi)."""
        T = len(alns)
        assert T == len(trees) and T >= 1
        packed = [self._pack(t[0], t[1], t[2], t[3], t[4], t[5], t[7], t[8]) for t in trees]
        R = packed[0][6].size
        assert all(p[6].size == R for p in packed), "all trees of a batch must share the number of rates"
        NN = _i32([p[0].size for p in packed])
        mu = _f64([t[6] for t in trees])

        def ptrs(i):
            return (C.c_void_p * T)(*[p[i].ctypes.data for p in packed])

        aln_ptrs = (C.c_void_p * T)(*[a.handle.value for a in alns])
        ll = np.zeros(T, dtype=np.float64)
        grads = [np.zeros(max(int(n) - 1, 1), dtype=np.float64) for n in NN] if want_grad else None
        gptrs = (C.c_void_p * T)(*[g.ctypes.data for g in grads]) if want_grad else None
        self._check(self.lib.mcp_eval_batch(self.handle, T, aln_ptrs, NN.ctypes.data, ptrs(0), ptrs(1), ptrs(2),
                                            ptrs(3), ptrs(4), ptrs(5), mu.ctypes.data, ptrs(6), R, ptrs(7),
                                            int(want_grad), ll.ctypes.data, gptrs))
        if want_grad:
            return ll, [g[:int(n) - 1] for g, n in zip(grads, NN)]
        return ll, None


class PreparedBatch:
    """mcp_eval_batch (T >= 1 trees) with every argument array packed ONCE: what a compiled host (the
    Julia glue keeps its arrays between leapfrog steps) pays per call is the C call itself, not Python's
    array conversions.  `set_blv(t, blv)` overwrites tree t's branch lengths in place."""

    def __init__(self, ctx: "Context", alns: Sequence[Alignment], trees: Sequence[tuple], want_grad: bool = True):
        self.ctx, self.T, self.want_grad = ctx, len(alns), bool(want_grad)
        assert self.T == len(trees) and self.T >= 1
        self._alns = list(alns)
        self._packed = [Context._pack(t[0], t[1], np.array(t[2], dtype=np.float64), t[3], t[4], t[5], t[7], t[8]) for t in trees]
        self.R = self._packed[0][6].size
        assert all(p[6].size == self.R for p in self._packed)
        self.NN = _i32([p[0].size for p in self._packed])
        self._mu = _f64([t[6] for t in trees])
        T = self.T
        self._ptrs = [(C.c_void_p * T)(*[p[i].ctypes.data for p in self._packed]) for i in range(8)]
        self._aln_ptrs = (C.c_void_p * T)(*[a.handle.value for a in alns])
        self.ll = np.zeros(T, dtype=np.float64)
        self.grads = [np.zeros(max(int(n) - 1, 1), dtype=np.float64) for n in self.NN]
        self._gptrs = (C.c_void_p * T)(*[g.ctypes.data for g in self.grads]) if want_grad else None

    def set_blv(self, t: int, blv):
        self._packed[t][2][:] = blv

    def eval(self):
        """Returns (ll[T], [grad_t]) -- views of buffers that the next call overwrites."""
        c, p = self.ctx, self._ptrs
        c._check(c.lib.mcp_eval_batch(c.handle, self.T, self._aln_ptrs, self.NN.ctypes.data, p[0], p[1], p[2], p[3],
                                      p[4], p[5], self._mu.ctypes.data, p[6], self.R, p[7], int(self.want_grad),
                                      self.ll.ctypes.data, self._gptrs))
        return self.ll, ([g[:int(n) - 1] for g, n in zip(self.grads, self.NN)] if self.want_grad else None)


def model_reorder(U, D, Uinv):
    """Host-only view of the eigen-decomposition as the kernels see it (mcp_model_reorder):
    returns (U, D, Uinv, null_last)."""
    lib = load()
    U = np.asfortranarray(U, dtype=np.float64)
    Uinv = np.asfortranarray(Uinv, dtype=np.float64)
    D = _f64(D)
    K = D.size
    assert U.shape == (K, K) and Uinv.shape == (K, K)
    Uo, Uio, Do = np.zeros((K, K), order="F"), np.zeros((K, K), order="F"), np.zeros(K)
    flag = C.c_int(0)
    rc = lib.mcp_model_reorder(K, U.ctypes.data, D.ctypes.data, Uinv.ctypes.data, Uo.ctypes.data, Do.ctypes.data,
                               Uio.ctypes.data, C.byref(flag))
    if rc:
        raise McpError(rc, lib.mcp_last_error(None).decode())
    return Uo, Do, Uio, bool(flag.value)


def schedule_fetch_list(postorder_num, parent_num, leaf_row, cherries: bool = True) -> np.ndarray:
    """Host-only: fetch list of the gradient pass's operand ring (post slots in the order they are read)."""
    lib = load()
    po, pa, lr = _i32(postorder_num), _i32(parent_num), _i32(leaf_row)
    NN = po.size
    out = np.zeros(2 * NN + 8, dtype=np.uint16)
    n = C.c_int32()
    lib.mcp_schedule_fetch_list.argtypes = [C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_int, C.POINTER(C.c_int32)]
    rc = lib.mcp_schedule_fetch_list(NN, po.ctypes.data, pa.ctypes.data, lr.ctypes.data, int(bool(cherries)), out.ctypes.data,
                                     out.size, C.byref(n))
    if rc:
        raise McpError(rc, lib.mcp_last_error(None).decode())
    return out[:n.value].copy()


def schedule_dump(postorder_num, parent_num, leaf_row, want_grad: bool, by_levels: bool = False, cherries: bool = False):
    """Host-only view of the device walk program (no GPU needed)."""
    lib = load()
    po, pa, lr = _i32(postorder_num), _i32(parent_num), _i32(leaf_row)
    NN = po.size
    cap = 2 * NN + 8
    post = np.zeros((cap, 8), dtype=np.int32)
    pre = np.zeros((cap, 8), dtype=np.int32)
    info = np.zeros(8, dtype=np.int32)
    rc = lib.mcp_schedule_dump(NN, po.ctypes.data, pa.ctypes.data, lr.ctypes.data, int(bool(want_grad)) | (2 if by_levels else 0) | (4 if cherries else 0),
                               post.ctypes.data, cap, pre.ctypes.data, cap, info.ctypes.data)
    if rc:
        raise McpError(rc, lib.mcp_last_error(None).decode())
    return {"post": post[:info[0]].copy(), "pre": pre[:info[1]].copy(), "n_slots": int(info[2]),
            "n_stack": int(info[3]), "n_dnodes": int(info[4]), "post_levels": int(info[5]),
            "pre_levels": int(info[6]), "n_cherries": int(info[7])}
