"""Site sharding across GPUs: one process per GPU (torch.distributed), each rank owns a
contiguous block of alignment columns for all nodes and rate categories, and the only exchange
per evaluation is one all-reduce (sum) of the NN doubles [logL, grad[1..NN-1]].
torch is plumbing here (process group, device tensor for the NCCL call); the evaluation itself
is the C-ABI `mcp_eval_device`, enqueued on torch's current stream so the collective follows it
in stream order without a host synchronisation in between.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np

from .phylodist import DeviceAlignment, PhyloDist, _device_alignment, _tree_args, get_context


def shard_bounds(S: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Rank g owns sites [g*ceil(S/G), min(S, (g+1)*ceil(S/G)))."""
    per = -(-S // world_size)
    lo = min(S, rank * per)
    return lo, min(S, lo + per)


def local_shard(aln: DeviceAlignment, world_size: int, rank: int) -> DeviceAlignment:
    lo, hi = shard_bounds(aln.S, world_size, rank)
    return aln.site_block(lo, hi)


def allreduce_sum(t, group=None):
    """Sum of the packed [logL, grad] vector over ranks (NCCL on device tensors, gloo on CPU
    tensors in the host-logic tests)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class ShardedEvaluator:
    """gradlogpdf / logpdf of one PhyloDist whose alignment is split over the ranks of a process
    group.  Every rank calls with the same tree and gets the same (all-reduced) result."""

    def __init__(self, local_alignment: DeviceAlignment, device: int, group=None):
        import torch

        self.torch = torch
        self.aln = local_alignment
        self.device = device
        self.group = group
        self.ctx = get_context(device)
        self._out = None
        self._pinned = None

    def _buffers(self, NN: int):
        torch = self.torch
        if self._out is None or self._out.numel() != NN:
            self._out = torch.empty(NN, dtype=torch.float64, device=f"cuda:{self.device}")
            self._pinned = torch.empty(NN, dtype=torch.float64).pin_memory()
        return self._out, self._pinned

    def evaluate(self, d: PhyloDist, want_grad: bool = True):
        ft, targs = _tree_args(d)
        return self.evaluate_flat(ft.leaf_nums, d.nbase, targs, want_grad)

    def evaluate_flat(self, leaf_nums, K: int, targs, want_grad: bool = True):
        """Same evaluation from already-flattened inputs: targs = (postorder_num, parent_num, blv,
        U, D, Uinv, mu, rates, pi), exactly the argument list of mcp_eval_device."""
        torch = self.torch
        NN = len(targs[0])
        out, pinned = self._buffers(NN)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream()
            self.ctx.set_stream(stream.cuda_stream)
            aln = _device_alignment(self.aln, leaf_nums, K, self.ctx)
            self.ctx.eval_device(aln, *targs, want_grad=want_grad, d_out_ptr=out.data_ptr())
            allreduce_sum(out, self.group)
            pinned.copy_(out, non_blocking=True)
            stream.synchronize()
        res = pinned.numpy()
        return float(res[0]), (res[1:].copy() if want_grad else None)

    def gradlogpdf(self, d: PhyloDist):
        return self.evaluate(d, True)

    def logpdf(self, d: PhyloDist) -> float:
        return self.evaluate(d, False)[0]


def sharded_sum(local_eval: Callable[[], Tuple[float, Optional[np.ndarray]]], NN: int, group=None):
    """Host-logic twin of ShardedEvaluator.evaluate for CPU tests: packs a rank-local
    (logL, grad) into the same NN-vector, all-reduces it and unpacks."""
    import torch

    ll, g = local_eval()
    t = torch.zeros(NN, dtype=torch.float64)
    t[0] = ll
    if g is not None:
        t[1:] = torch.from_numpy(np.asarray(g, dtype=np.float64))
    allreduce_sum(t, group)
    return float(t[0]), t[1:].numpy().copy()


class PipelinedEvaluator:
    """Evaluation of an alignment that has to be (re-)uploaded from host memory for the call.  The
    site axis is cut into blocks; `mcp_alignment_update_codes` puts every block's transfer on the
    context's copy stream, the evaluations follow on the compute stream and each waits only for its
    own block, so block b+1 crosses PCIe while block b is evaluated.  Block sizes are multiples of
    one wave of the persistent grid (`mcp_wave_columns`) and grow geometrically (1, 2, 4, ... waves,
    then the rest): the first evaluation starts after a fraction of a millisecond of transfer and no
    launch ends in a ragged wave.  Block results are summed on the device, all-reduced over the
    process group if there is one, and read back once.

    `codes` is the rank's (n_leaves, S) uint8 block; pinned per-block copies are made at the first
    evaluation (the wave size depends on the number of rate categories and the tree size)."""

    def __init__(self, codes: np.ndarray, leaf_nums, K: int, device: int, n_blocks: int = 5, group=None):
        import torch

        from . import capi

        self.torch = torch
        self.device, self.group, self.K = device, group, int(K)
        self.leaf_nums = np.asarray(leaf_nums, dtype=np.int32)
        self.codes = codes
        self.n_blocks = max(1, int(n_blocks))
        self.ctx = capi.Context(device)
        self.blocks = []
        self.bounds = []
        self._out = None

    @staticmethod
    def plan_blocks(S: int, wave_sites: int, n_blocks: int):
        """[(lo, hi)] site ranges: 1, 2, 4, ... waves, the last block takes the rest."""
        wave_sites = max(1, int(wave_sites))
        waves = S // wave_sites
        if n_blocks <= 1 or waves < 2:
            return [(0, S)] if S > 0 else []
        sizes, w = [], 1
        while len(sizes) < n_blocks - 1 and sum(sizes) + w < waves:
            sizes.append(w)
            w *= 2
        bounds, lo = [], 0
        for sz in sizes:
            bounds.append((lo, lo + sz * wave_sites))
            lo += sz * wave_sites
        bounds.append((lo, S))
        return bounds

    def _build(self, R: int, NN: int):
        torch = self.torch
        S = self.codes.shape[1]
        wave_sites = self.ctx.wave_columns(self.K, NN, True) // max(R, 1)
        self.bounds = self.plan_blocks(S, wave_sites, self.n_blocks)
        if len(self.bounds) > 1:
            # pin the tile width the wave size was computed for: the automatic choice narrows the
            # tiles of inputs below two waves, which a one-wave block is by construction
            self.ctx.set_launch(256 if self.K <= 6 else 128, 0)
            self.ctx.set_columns_per_thread(2 if self.K <= 4 else 1)
        for lo, hi in self.bounds:
            host = torch.from_numpy(np.ascontiguousarray(self.codes[:, lo:hi])).pin_memory()
            aln = self.ctx.alignment_from_codes(host.numpy(), self.K, self.leaf_nums)
            self.blocks.append((aln, host))
        self.codes = None

    def evaluate(self, d: PhyloDist, want_grad: bool = True, upload: bool = True):
        torch = self.torch
        ft, targs = _tree_args(d)
        NN = ft.NN
        if not self.blocks:
            self._build(len(d.rates), NN)
        if self._out is None or self._out.shape[1] != NN:
            self._out = torch.empty((len(self.blocks), NN), dtype=torch.float64, device=f"cuda:{self.device}")
            self._pinned = torch.empty(NN, dtype=torch.float64).pin_memory()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream()
            self.ctx.set_stream(stream.cuda_stream)
            # Issue order matters: host-to-device transfers are served first-in first-out, and every
            # evaluation starts with a small parameter upload of its own.  Transfer b, evaluation b,
            # transfer b+1, ...: the parameters of evaluation b queue right behind the block they
            # need anyway, and transfer b+1 then runs under the kernels of evaluation b.
            for i, (aln, host) in enumerate(self.blocks):
                if upload:
                    aln.update_codes(host.data_ptr())
                self.ctx.eval_device(aln, *targs, want_grad=want_grad, d_out_ptr=self._out[i].data_ptr())
            total = self._out.sum(dim=0)
            allreduce_sum(total, self.group)
            self._pinned.copy_(total, non_blocking=True)
            stream.synchronize()
        res = self._pinned.numpy()
        return float(res[0]), (res[1:].copy() if want_grad else None)

    def gradlogpdf(self, d: PhyloDist):
        return self.evaluate(d, True)

    def logpdf(self, d: PhyloDist) -> float:
        return self.evaluate(d, False)[0]

    def close(self):
        for aln, _ in self.blocks:
            aln.close()
        if self.ctx is not None:
            self.ctx.close()
        self.blocks, self.ctx = [], None
