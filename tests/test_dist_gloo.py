"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo process group, site blocks
per rank, one all-reduce of the packed [logL, grad] vector.  The per-shard evaluator here is
the CPU oracle (test infrastructure) standing in for the GPU — what is under test is the
partition and the reduction, which are the same code the NCCL path runs (dist.py)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import mcphylo_jl_b200 as mcp
    import oracle
    from mcphylo_jl_b200.dist import local_shard, shard_bounds, sharded_sum

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(99)           # same inputs on every rank
    tree = mcp.random_tree(14, rng, multifurcate=True)
    pi = np.array([0.1, 0.2, 0.3, 0.4])
    sr = np.array([1.0, 2.0, 1.5, 0.8, 2.5, 1.2])
    rates = mcp.discrete_gamma_rates(0.5, 0.5, 4)
    S = 101                                    # odd: ragged last block
    codes, leaf_nums = mcp.simulate_codes(tree, mcp.GTR(pi, sr), pi, rates, S, rng)
    full = mcp.DeviceAlignment(codes, leaf_nums, 4)
    ft = mcp.flatten(tree)
    U, D, Uinv, mu = mcp.GTR(pi, sr)

    def evaluate(aln):
        x = oracle.codes_to_dense(aln.codes, aln.leaf_nums, 4, ft.NN)
        return oracle.felsenstein(x, ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, rates, pi, True, 1)

    mine = local_shard(full, world, rank)
    lo, hi = shard_bounds(S, world, rank)
    assert mine.S == hi - lo
    ll, g = sharded_sum(lambda: evaluate(mine), ft.NN)
    ll_full, g_full = evaluate(full)
    ok = abs(ll - ll_full) <= 1e-12 * abs(ll_full) and np.allclose(g, g_full, rtol=1e-10, atol=1e-10)
    q.put((rank, bool(ok), lo, hi))
    dist.barrier()
    dist.destroy_process_group()


def test_site_sharding_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert (res[0][2], res[0][3], res[1][2], res[1][3]) == (0, 51, 51, 101)


def test_shard_bounds_cover_everything():
    from mcphylo_jl_b200.dist import shard_bounds
    for S in (0, 1, 7, 8, 1000, 1_000_000):
        for G in (1, 2, 4, 8):
            blocks = [shard_bounds(S, G, r) for r in range(G)]
            assert blocks[0][0] == 0 and blocks[-1][1] == S
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            assert all(0 <= hi - lo <= -(-S // G) for lo, hi in blocks)
