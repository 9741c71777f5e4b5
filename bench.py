#!/usr/bin/env python
"""bench.py — logpdf+gradient evaluations/s of the PhyloDist hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one gradlogpdf evaluation (log-likelihood + all branch-length gradients) of the
workload BASELINE.json's metric is quoted on: 1000 taxa x 1,000,000 sites, 4 states (GTR),
4 discrete-Gamma rate categories (config[3]).  With N GPUs the site axis is sharded (fixed total
work -> "scaling": "strong"); each evaluation ends in one all-reduce of [logL, gradient].

Output: ONE JSON line on rank 0 (keys documented in the task contract):
  value        evaluations/s, alignment resident in HBM (per-step host inputs are only the tree
               arrays / branch lengths / model, which every evaluation uploads anyway)
  e2e          same metric through the public API with HOST buffers: every step re-uploads the
               alignment codes from pinned host memory, flattens the tree, runs the model's
               eigendecomposition, and reads the result back
  roofline     algorithmic bytes of one evaluation / CUDA-event time of the walk kernel, against
               the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline the CPU oracle (a C/OpenMP port of the reference's loops; the reference itself is
               Julia and cannot run here) on a bounded site sample, extrapolated linearly in S

--impl reference times that same oracle port (all host threads) as the reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_taxa, sites, K, R, tree seed, data seed)  — SURVEY.md §8d
    "cfg1": (10, 1_000, 2, 1, 20241, 1001),
    "cfg2": (50, 10_000, 2, 1, 20242, 1002),
    "cfg3": (200, 100_000, 4, 4, 20243, 1003),
    "cfg4": (1000, 1_000_000, 4, 4, 20244, 1004),
    # 256 independent trees (MultiplePhyloDist / proposal batch) in one launch; seeds 5000+i
    "cfg5": (100, 50_000, 2, 1, 5000, 1005),
}
CFG5_TREES = 256
GTR_PI = np.array([0.1, 0.2, 0.3, 0.4])
GTR_EXCH = np.array([1.0, 2.0, 1.5, 0.8, 2.5, 1.2])
RESTRICTION_PI = np.array([0.3, 0.7])
SIM_POOL = 32768           # columns simulated under the model; bootstrap-resampled up to S


def make_workload(name, sites_override=None):
    import mcphylo_jl_b200 as mcp

    n_taxa, S, K, R, tseed, dseed = WORKLOADS[name]
    if sites_override:
        S = int(sites_override)
    tree = mcp.random_tree(n_taxa, np.random.default_rng(tseed))
    if K == 4:
        pi, srates, model = GTR_PI, GTR_EXCH, mcp.GTR
        rates = mcp.discrete_gamma_rates(0.5, 0.5, R)
    else:
        pi, srates, model = RESTRICTION_PI, np.zeros(1), mcp.Restriction
        rates = np.ones(1)
    return dict(name=name, tree=tree, S=S, K=K, R=R, pi=pi, srates=srates, model=model, rates=rates,
                n_taxa=n_taxa, dseed=dseed)


def make_codes(w, lo, hi):
    """Columns [lo, hi) of the synthetic alignment: SIM_POOL columns simulated down the tree under
    the evaluating model (1 % gaps), bootstrap-resampled to S columns.  Every rank generates the
    same pool and the same resampling, then keeps its block, so the global alignment does not
    depend on the number of GPUs."""
    import mcphylo_jl_b200 as mcp

    rng = np.random.default_rng(w["dseed"])
    pool_n = min(w["S"], SIM_POOL)
    pool, leaf_nums = mcp.simulate_codes(w["tree"], w["model"](w["pi"], w["srates"]), w["pi"], w["rates"],
                                         pool_n, rng, gap_frac=0.01)
    if pool_n == w["S"]:
        return np.ascontiguousarray(pool[:, lo:hi]), leaf_nums
    pick = np.random.default_rng(w["dseed"] + 7).integers(0, pool_n, size=w["S"])[lo:hi]
    return np.take(pool, pick, axis=1), leaf_nums


def algorithmic_bytes(n_taxa, S, K, R, want_grad=True):
    """every internal partial written once and read once as a child (post pass);
    pre pass reads pre[mother], the child's partial, writes pre[child]; leaves are 1-byte codes."""
    I = n_taxa - 1
    C = S * R
    b_post = C * K * 8 * (2 * I - 1) + S * n_taxa
    b_pre = C * K * 8 * 3 * (I - 1) + S * n_taxa
    return b_post + (b_pre if want_grad else 0)


def init_nccl(local_rank):
    """Process group + communicator creation with stdout pointed at stderr at the file-descriptor
    level: NCCL prints its version banner to stdout (NCCL_DEBUG=VERSION on the bench boxes), and
    stdout must carry exactly one JSON line."""
    import torch
    import torch.distributed as dist

    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        t = torch.zeros(1, device=f"cuda:{local_rank}")
        dist.all_reduce(t)                      # forces communicator creation now
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


class ClockSampler:
    """nvidia-smi style clock / throttle-reason sampling during the timed region (NVML)."""

    def __init__(self, device_index):
        self.samples, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "power_w_max": max(self.power) if self.power else None,
                "samples": len(self.samples)}


def cpu_oracle_rate(w, codes, leaf_nums, sample_sites, steps, warmup, threads=0):
    """Times the CPU oracle port on the first `sample_sites` columns and extrapolates to S
    (columns are independent, cost is linear in S).  Returns (evals/s at full S, seconds/step on
    the sample, threads)."""
    import mcphylo_jl_b200 as mcp
    import oracle

    oracle.build()
    ft = mcp.flatten(w["tree"])
    s = min(sample_sites, codes.shape[1])
    x = oracle.codes_to_dense(codes[:, :s], leaf_nums, w["K"], ft.NN)
    U, D, Uinv, mu = w["model"](w["pi"], w["srates"])
    if threads <= 0:
        # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1; the oracle sets
        # its OpenMP team size explicitly, so that default does not shrink the CPU baseline)
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        oracle.felsenstein(x, ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, w["rates"], w["pi"], True, threads)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    return 1.0 / (sec * (w["S"] / s)), sec, threads, s


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = make_workload(args.workload, args.sites)
    sample = args.cpu_sample or 2000
    codes, leaf_nums = make_codes(w, 0, min(w["S"], sample))
    rate, sec, threads, s = cpu_oracle_rate(w, codes, leaf_nums, sample, args.steps, max(args.warmup, 1))
    line = {
        "impl": "reference",
        "metric": "logpdf+gradient evaluations/s", "value": rate, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3 * (w["S"] / s),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(w, args.gpus),
        "cpu_baseline": {"value": rate, "unit": "evals/s", "cores": threads, "kind": "port",
                         "sample": f"first {s} of {w['S']} sites, full tree; {sec:.3f} s per evaluation on the "
                                   f"sample, extrapolated linearly in sites (columns are independent)"},
        "e2e": {"value": rate, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is pure Julia (no Julia toolchain here, tree library un-vendored): this arm times "
                "oracle/felsenstein_oracle.c, a loop-for-loop C/OpenMP port of the reference's CPU path",
    }
    print(json.dumps(line))


def workload_config(w, n_gpus):
    return {"workload": f"{w['name']}: {w['n_taxa']} taxa x {w['S']} sites, K={w['K']} "
                        f"({'GTR' if w['K'] == 4 else 'Restriction'}), R={w['R']} discrete-Gamma categories, "
                        f"logpdf+gradient per step",
            "n_taxa": w["n_taxa"], "sites": w["S"], "states": w["K"], "rate_categories": w["R"],
            "sharding": f"sites/{n_gpus}", "l2": "inputs larger than L2 (per-evaluation working set >> 126 MB)"}


def run_b200(args):
    import torch
    import torch.distributed as dist

    import mcphylo_jl_b200 as mcp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback exists)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        init_nccl(local_rank)
    w = make_workload(args.workload, args.sites)
    lo, hi = mcp.shard_bounds(w["S"], world, rank)
    t_gen = time.perf_counter()
    codes, leaf_nums = make_codes(w, lo, hi)
    t_gen = time.perf_counter() - t_gen
    aln = mcp.DeviceAlignment(codes, leaf_nums, w["K"])
    ev = mcp.ShardedEvaluator(aln, local_rank)
    ctx = ev.ctx
    if args.block or args.ctas_per_sm:
        ctx.set_launch(args.block, args.ctas_per_sm)
    tree = w["tree"]
    blv0 = mcp.get_branchlength_vector(tree)

    def dist_for(step):
        # new branch lengths every step (the leapfrog pattern): nothing can be reused across steps
        mcp.set_branchlength_vector(tree, blv0 * (1.0 + 1e-3 * (step % 7)))
        return mcp.PhyloDist(tree, w["pi"], w["srates"], w["rates"], w["model"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # alignment upload (one-off, reported separately)
    t_up = time.perf_counter()
    ll0, grad0 = ev.gradlogpdf(dist_for(0))
    torch.cuda.synchronize()
    t_up = time.perf_counter() - t_up

    sampler = ClockSampler(local_rank)
    stream = torch.cuda.current_stream()

    # ---- device-resident throughput -----------------------------------------------------
    # Host inputs of every step (flattened tree with that step's branch lengths + the model's
    # eigendecomposition) are prepared before the clock starts: `value` times the C-ABI evaluation
    # (parameter upload, 3 kernels, all-reduce, result download); the Python-side tree traversal of
    # the public API is part of `e2e` below.
    from mcphylo_jl_b200.phylodist import _tree_args
    prepared = []
    for i in range(max(args.steps, args.warmup)):
        d = dist_for(i)
        ft, targs = _tree_args(d)
        prepared.append((ft.leaf_nums, d.nbase, targs))
    for i in range(args.warmup):
        ev.evaluate_flat(*prepared[i], True)
    walk_ms = []
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        ll, grad = ev.evaluate_flat(*prepared[i], True)
        walk_ms.append(ctx.stats()["walk_ms"])
    e1.record(stream)
    barrier()
    sampler.stop()
    ms_total = e0.elapsed_time(e1)
    stats = ctx.stats()

    # ---- end to end through the public API with host buffers -------------------------------
    # PipelinedEvaluator: every step uploads the whole alignment from pinned host memory in site
    # blocks of 1, 2, 4, ... grid waves (block b+1 crosses PCIe on the copy stream while block b is
    # evaluated), flattens the tree, runs the eigendecomposition, evaluates, all-reduces and reads the
    # result back.
    pipe = mcp.PipelinedEvaluator(codes, leaf_nums, w["K"], local_rank, n_blocks=args.e2e_blocks)

    def e2e_step(i):
        return pipe.gradlogpdf(dist_for(i))

    for i in range(min(args.warmup, 3)):
        e2e_step(i)
    barrier()
    sampler.start()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    for i in range(args.steps):
        ll_e, grad_e = e2e_step(i)
    f1.record(stream)
    barrier()
    sampler.stop()
    ms_e2e = f0.elapsed_time(f1)

    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        NN = 2 * w["n_taxa"] - 1
        I = w["n_taxa"] - 1
        evals_per_s = args.steps / (ms_total * 1e-3)
        e2e_per_s = args.steps / (ms_e2e * 1e-3)
        local_S = hi - lo
        b_alg = algorithmic_bytes(w["n_taxa"], local_S, w["K"], w["R"], True)
        wk = float(np.mean(walk_ms)) if walk_ms and all(m > 0 for m in walk_ms) else None
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "walk_traffic.json")) as fh:
                tj = json.load(fh)
            if tj.get("workload") == w["name"] and int(tj.get("sites_per_gpu", -1)) == local_S:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass
        roofline = {"bound": "hbm", "achieved": (b_alg / (wk * 1e-3) / 1e9) if wk else None, "peak": peak,
                    "unit": "GB/s", "frac": (b_alg / (wk * 1e-3) / 1e9 / peak) if wk else None, "traffic": traffic,
                    "kernel": "felsenstein_walk<4>" if w["K"] == 4 else f"felsenstein_walk<{w['K']}>",
                    "kernel_ms": wk, "algorithmic_bytes_per_launch": b_alg,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"}
        line = {
            "metric": "logpdf+gradient evaluations/s", "value": evals_per_s, "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": f"synthetic: {min(w['S'], SIM_POOL)} columns simulated down a random {w['n_taxa']}-taxon tree "
                    f"under the evaluating model (1% gaps), bootstrap-resampled to {w['S']} sites",
            "config": workload_config(w, world),
            "site_node_updates_per_s": (I + NN - 1) * w["S"] * w["R"] * evals_per_s,
            "site_node_updates_note": "post-order node updates I*S*R plus branch-gradient updates (NN-1)*S*R per evaluation",
            "roofline": roofline,
            "e2e": {"value": e2e_per_s, "unit": "evals/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(codes.nbytes + stats["h2d_bytes"]),
                    "d2h_bytes_per_step": int(NN * 8),
                    "site_blocks": [hi - lo for lo, hi in pipe.bounds],
                    "what": f"per step: alignment codes re-uploaded from pinned host memory in {len(pipe.bounds)} site "
                            "blocks (1, 2, 4, ... grid waves) on the copy stream, each overlapped with the evaluation "
                            "of the previous block; tree flattened, eigendecomposition, mcp_eval_device per block, "
                            "all-reduce, result read back"},
            "gpu_launches": int(stats["kernel_launches"] * ((args.steps + args.warmup + 1) +
                                                            len(pipe.bounds) * (args.steps + min(args.warmup, 3)))),
            "gpu_launches_timed": int(stats["kernel_launches"] * args.steps),
            "clocks": sampler.summary(),
            "launch": {"grid": stats["grid"], "block": stats["block"], "tiles": stats["tiles"],
                       "scratch_bytes": stats["scratch_bytes"]},
            "setup": {"alignment_generate_s": t_gen, "first_eval_incl_upload_s": t_up,
                      "codes_bytes_per_gpu": int(codes.nbytes)},
            "value_what": "per step: mcp_eval_device from pre-flattened host arrays (tree arrays, branch lengths, "
                          "model uploaded every step), all-reduce of [logL, grad], result read back; alignment resident",
            "result_check": {"logL_at_initial_branch_lengths": ll0, "grad_l2_at_initial_branch_lengths":
                             float(np.linalg.norm(grad0)), "grad_finite": bool(np.all(np.isfinite(grad))),
                             "e2e_matches_resident": bool(abs(ll_e - ll) <= 1e-9 * abs(ll))},
        }
        if world == 1 and not args.no_cpu_baseline:
            sample = args.cpu_sample or 10000
            rate, sec, threads, s = cpu_oracle_rate(w, codes, leaf_nums, sample, 2, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "evals/s", "cores": threads, "kind": "port",
                                    "sample": f"first {s} of {w['S']} sites, full tree; {sec:.3f} s per evaluation "
                                              f"on the sample (OpenMP, {threads} threads), extrapolated linearly in sites"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_batch(args):
    """cfg5: T independent trees, each with its own alignment, evaluated by ONE mcp_eval_batch call
    per step (the MultiplePhyloDist path).  With N GPUs the trees are dealt round-robin to the ranks;
    there is no collective on the data path (each rank returns its own trees' results)."""
    import torch
    import torch.distributed as dist

    import mcphylo_jl_b200 as mcp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback exists)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        init_nccl(local_rank)
    n_taxa, S, K, R, tseed, dseed = WORKLOADS["cfg5"]
    if args.sites:
        S = args.sites
    T = args.trees or CFG5_TREES
    pi = RESTRICTION_PI
    model_out = mcp.Restriction(pi, [])
    mine = list(range(rank, T, world))
    trees, alns = [], []
    t_gen = time.perf_counter()
    for i in mine:
        tree = mcp.random_tree(n_taxa, np.random.default_rng(tseed + i))
        rng = np.random.default_rng(dseed * 1000 + i)
        pool_n = min(S, 8192)
        pool, leaf_nums = mcp.simulate_codes(tree, model_out, pi, np.ones(1), pool_n, rng, gap_frac=0.01)
        codes = pool if pool_n == S else np.take(pool, rng.integers(0, pool_n, size=S), axis=1)
        trees.append(tree)
        alns.append(mcp.DeviceAlignment(codes, leaf_nums, K))
    t_gen = time.perf_counter() - t_gen
    mpd = mcp.MultiplePhyloDist(trees, pi, [0.0], [1.0], mcp.Restriction)
    blv0 = [mcp.get_branchlength_vector(t) for t in trees]
    ctx = mcp.get_context(local_rank)
    if args.block or args.ctas_per_sm:
        ctx.set_launch(args.block, args.ctas_per_sm)

    def step(i):
        f = 1.0 + 1e-3 * (i % 7)
        for t, b in zip(trees, blv0):
            mcp.set_branchlength_vector(t, b * f)
        return mcp.multi_gradlogpdf(mpd, alns, device=local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    res = step(0)
    for i in range(args.warmup):
        step(i)
    sampler = ClockSampler(local_rank)
    walk_ms = []
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    for i in range(args.steps):
        res = step(i)
        walk_ms.append(ctx.stats()["walk_ms"])
    barrier()
    ms_total = (time.perf_counter() - t0) * 1e3
    sampler.stop()
    stats = ctx.stats()
    tt = torch.tensor([ms_total], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total = float(tt[0])
    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        b_alg = len(mine) * algorithmic_bytes(n_taxa, S, K, R, True)
        wk = float(np.mean(walk_ms))
        line = {
            "metric": "logpdf+gradient tree-evaluations/s (batched)", "value": T * args.steps / (ms_total * 1e-3),
            "unit": "tree-evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic: per tree 8192 columns simulated under the model, bootstrap-resampled",
            "config": {"workload": f"cfg5: {T} trees x {n_taxa} taxa x {S} binary sites (Restriction), one batched "
                                   f"logpdf+gradient call per step through MultiplePhyloDist/__logpdf",
                       "trees": T, "n_taxa": n_taxa, "sites": S, "states": K, "rate_categories": R,
                       "sharding": f"trees/{world}", "l2": "inputs larger than L2",
                       "timing": "host wall clock around the public API call (includes tree flattening in Python)"},
            "roofline": {"bound": "hbm", "achieved": b_alg / (wk * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": b_alg / (wk * 1e-3) / 1e9 / peak, "traffic": None, "kernel_ms": wk,
                         "algorithmic_bytes_per_launch": b_alg},
            "e2e": {"value": T * args.steps / (ms_total * 1e-3), "unit": "tree-evals/s",
                    "h2d_bytes_per_step": int(stats["h2d_bytes"]), "d2h_bytes_per_step": int(stats["d2h_bytes"]),
                    "what": "alignments resident; per step all tree arrays, branch lengths and models uploaded"},
            "gpu_launches": int(stats["kernel_launches"] * (args.steps + args.warmup + 1)), "clocks": sampler.summary(),
            "launch": {"grid": stats["grid"], "block": stats["block"], "tiles": stats["tiles"]},
            "setup": {"alignment_generate_s": t_gen},
            "result_check": {"sum_logL": float(sum(r[0] for r in res)), "finite": bool(all(np.all(np.isfinite(r[1])) for r in res))},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--sites", type=int, default=0, help="override the number of sites (experiments only)")
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--trees", type=int, default=0, help="cfg5: number of trees in the batch")
    ap.add_argument("--e2e-blocks", type=int, default=5, help="site blocks of the pipelined end-to-end path")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "cfg5":
        run_batch(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
