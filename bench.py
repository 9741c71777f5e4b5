#!/usr/bin/env python
"""bench.py — logpdf+gradient evaluations/s of the PhyloDist hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one gradlogpdf evaluation (log-likelihood + all branch-length gradients) of the
workload BASELINE.json's metric is quoted on: 1000 taxa x 1,000,000 sites, 4 states (GTR),
4 discrete-Gamma rate categories (config[3]).  With N GPUs the site axis is sharded (fixed total
work -> "scaling": "strong"); each evaluation ends in one all-reduce of [logL, gradient].

Everything measured goes through the C ABI (include/mcphylo_b200.h):
  --mp group (default)  ONE host process owns all N GPUs through mcp_create_multi -- the shape a Julia
                        session has.  Under torchrun rank 0 drives the N GPUs; the other ranks only take
                        part in the launcher's barriers (they hold no GPU).
  --mp ranks            one process per GPU: every torchrun rank creates mcp_create_rank on its GPU and
                        the library all-reduces over an ncclCommInitRank communicator.

Output: ONE JSON line on rank 0 (keys documented in the task contract):
  value        evaluations/s, alignment resident in HBM (per-step host inputs are only the tree
               arrays / branch lengths / model, which every evaluation uploads anyway)
  e2e          same metric through mcp_eval_streamed with HOST buffers: every step re-uploads the
               alignment codes from pinned host memory (block-pipelined inside the library), flattens the
               tree, runs the model's eigendecomposition, and reads the result back
  roofline     three fractions of the measured HBM copy bandwidth (MEASURED_PEAKS.json) for the walk kernel:
               frac (SURVEY.md 8d model of a level-scheduled implementation), frac_walk (global loads and
               stores the walk algorithm itself issues, counted from the schedule), frac_physical (DRAM
               bytes measured by ncu for this launch, profiles/walk_traffic.json)
  cpu_baseline the CPU oracle (a C/OpenMP port of the reference's loops; the reference itself is
               Julia and cannot run here) on a bounded site sample, extrapolated linearly in S
  extra        short legs of the other BASELINE configs (cfg2, cfg3, cfg5), N = 1 only

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
L2_BYTES = 126e6


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


def make_codes(w, lo, hi, pool_cap=SIM_POOL):
    """Columns [lo, hi) of the synthetic alignment: SIM_POOL columns simulated down the tree under
    the evaluating model (1 % gaps), bootstrap-resampled to S columns.  Every rank generates the
    same pool and the same resampling, then keeps its block, so the global alignment does not
    depend on the number of GPUs."""
    import mcphylo_jl_b200 as mcp

    rng = np.random.default_rng(w["dseed"])
    pool_n = min(w["S"], pool_cap)
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


def walk_bytes(tree, S, K, R, want_grad=True):
    """Global-memory bytes the walk algorithm itself issues for one evaluation, counted from the device
    program: per column and K-vector (K*8 bytes) one store per stored post result, one load
    per stored post operand, one load per stored child partial in the gradient pass, one store + one load
    per pre vector that goes through the LIFO; leaf codes once per rate category and pass.  Second children
    and kept pre vectors never leave registers; cherries (both children leaves) are recomputed from their
    codes in the gradient pass instead of being stored and re-read.  Part of this traffic (LIFO entries, partials re-read soon
    after they were written) is served by L2 and never reaches DRAM: `traffic` is the measured remainder."""
    import mcphylo_jl_b200 as mcp
    from mcphylo_jl_b200 import capi

    ft = mcp.flatten(tree)
    leaf_row = np.full(ft.NN, -1, dtype=np.int32)
    leaf_row[ft.leaf_nums - 1] = np.arange(ft.leaf_nums.size)
    sd = capi.schedule_dump(ft.postorder_num, ft.parent_num, leaf_row, want_grad, cherries=K <= 6)
    post, pre = sd["post"], sd["pre"]
    flags = post[:, 5]
    n_store = int(np.sum((flags & 16) != 0))
    n_mem_post = int(np.sum((flags & 3) == 2) + np.sum(((flags >> 2) & 3) == 2))
    vec = n_store + n_mem_post
    n_leaf_reads = int(np.sum((flags & 3) == 0) + np.sum(((flags >> 2) & 3) == 0))
    if want_grad and len(pre):
        pf = pre[:, 5]
        n_mem_pre = int(np.sum((pf & 3) == 2) + np.sum(((pf >> 2) & 3) == 2))
        n_pop = int(np.sum(((pf >> 8) & 3) == 2))
        n_push = int(np.sum(((pf >> 10) & 3) == 2) + np.sum(((pf >> 12) & 3) == 2))
        vec += n_mem_pre + n_pop + n_push
        n_leaf_reads += int(np.sum((pf & 3) == 0) + np.sum(((pf >> 2) & 3) == 0) + 2 * np.sum((pf & 3) == 3))
    return int(S * R * K * 8 * vec + S * R * n_leaf_reads)


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh)
    except Exception:
        return {}


def load_traffic(workload, sites_per_gpu):
    """ncu dram__bytes_read.sum + dram__bytes_write.sum of the walk launch for this shape, if a capture of
    this exact (workload, sites per GPU) is committed under profiles/."""
    try:
        with open(os.path.join(ROOT, "profiles", "walk_traffic.json")) as fh:
            tj = json.load(fh)
        entries = tj.get("entries") or [tj]
        for e in entries:
            if e.get("workload") == workload and int(e.get("sites_per_gpu", -1)) == int(sites_per_gpu):
                return e.get("dram_bytes_per_launch"), e.get("source")
    except Exception:
        pass
    return None, None


def load_traffic_entry(workload, sites_per_gpu):
    try:
        with open(os.path.join(ROOT, "profiles", "walk_traffic.json")) as fh:
            tj = json.load(fh)
        for e in tj.get("entries") or [tj]:
            if e.get("workload") == workload and int(e.get("sites_per_gpu", -1)) == int(sites_per_gpu):
                return e
    except Exception:
        pass
    return None


def issue_block(w, local_S, kernel_ms, clocks, sm_count=148):
    """What the walk kernel is actually bound by: neither roof of the HBM model, but the instruction stream.  Warp
    instructions per launch (ncu `smsp__inst_executed.sum` of the committed capture of this exact shape) over the issue
    slots of the timed launch: SMs x 4 schedulers x measured SM clock x live kernel time."""
    try:
        e = load_traffic_entry(w["name"], local_S)
        mhz = float((clocks or {}).get("sm_mhz") or 0.0)
        if not e or not e.get("warp_instructions") or not kernel_ms or mhz <= 0:
            return None
        slots = sm_count * 4 * mhz * 1e6 * kernel_ms * 1e-3
        return {"warp_instructions_per_launch": e["warp_instructions"], "issue_slots_per_launch": slots,
                "frac": e["warp_instructions"] / slots, "sm_mhz": mhz,
                "fp64_pipe_active_pct_under_ncu": e.get("fp64_pipe_active_pct"),
                "issue_active_pct_under_ncu": e.get("issue_active_pct"),
                "what": "share of the warp-instruction issue slots the launch used (instructions from the committed ncu capture "
                        "of this shape, time and clock measured live): the walk is instruction-issue / FP64-latency bound, "
                        "which is why frac_physical sits near 0.5"}
    except Exception:
        return None


def roofline_block(w, local_S, kernel_ms, tree, want_grad=True):
    peaks = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    b_alg = algorithmic_bytes(w["n_taxa"], local_S, w["K"], w["R"], want_grad)
    b_walk = walk_bytes(tree, local_S, w["K"], w["R"], want_grad)
    traffic, src = load_traffic(w["name"], local_S)
    sec = kernel_ms * 1e-3 if kernel_ms else None

    def gbs(b):
        return b / sec / 1e9 if sec and b is not None else None

    def frac(b):
        return gbs(b) / peak if gbs(b) is not None else None

    return {"bound": "hbm", "achieved": gbs(b_alg), "peak": peak, "unit": "GB/s", "frac": frac(b_alg),
            "frac_what": "SURVEY 8d model (level-scheduled implementation materialising every partial) / kernel time / "
                         "peak; the walk moves less than this model, so values above 1 are not a bandwidth claim",
            "traffic": traffic, "traffic_source": src,
            "frac_physical": frac(traffic), "achieved_physical": gbs(traffic),
            "walk_bytes": b_walk, "frac_walk": frac(b_walk), "achieved_walk": gbs(b_walk),
            "walk_bytes_what": "global loads+stores the walk issues (stored post results and their re-reads, LIFO "
                               "pushes/pops, leaf codes), counted from the device program; L2 absorbs part of it",
            "kernel": f"felsenstein_walk<{w['K']}>", "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": b_alg,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"}


class stdout_to_stderr:
    """NCCL prints its version banner to stdout when the communicators are created (NCCL_DEBUG=VERSION on
    the bench boxes); stdout must carry exactly one JSON line, so file descriptor 1 points at stderr
    while the library may be talking."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


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
            self._stop.wait(0.05)

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


def oracle_window(ctx, w, codes, leaf_nums, stats, sites=256):
    """Parity stamp for the bench line: the first `sites` columns evaluated by the CUDA path with the
    LAUNCH SHAPE OF THE TIMED RUN forced (tile width, columns per thread, depth-first walk kernel) and by
    the CPU oracle.  Returns the relative errors (contract: logL <= 1e-10, gradient components <= 1e-8)."""
    import mcphylo_jl_b200 as mcp
    import oracle

    oracle.build()
    tree = w["tree"]
    ft = mcp.flatten(tree)
    s = min(sites, codes.shape[1])
    sub = np.ascontiguousarray(codes[:, :s])
    U, D, Uinv, mu = w["model"](w["pi"], w["srates"])
    x = oracle.codes_to_dense(sub, leaf_nums, w["K"], ft.NN)
    ll_o, g_o = oracle.felsenstein(x, ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, w["rates"], w["pi"], True, 0)
    ctx.set_launch(int(stats["block"]), 0)
    ctx.set_columns_per_thread(int(stats["columns_per_thread"]) or 1)
    ctx.set_level_mode(0)
    ctx.set_scratch_mode(0)
    try:
        aln = ctx.alignment_from_codes(sub, w["K"], leaf_nums)
        ll, g = ctx.eval(aln, ft.postorder_num, ft.parent_num, ft.blv, U, D, Uinv, mu, w["rates"], w["pi"], want_grad=True)
        shape = ctx.stats()
        aln.close()
    finally:
        ctx.set_launch(0, 0)
        ctx.set_columns_per_thread(0)
        ctx.set_level_mode(-1)
        ctx.set_scratch_mode(-1)
    scale = np.max(np.abs(g_o))
    return {"sites": int(s), "block": int(shape["block"]), "columns_per_thread": int(shape["columns_per_thread"]),
            "operand_ring": int(shape.get("operand_ring", 0)),
            "ll_rel_err": float(abs(ll - ll_o) / abs(ll_o)),
            "grad_max_rel_err": float(np.max(np.abs(g - g_o) / np.maximum(np.abs(g_o), 1e-3 * scale))),
            "logL_cuda": float(ll), "logL_oracle": float(ll_o),
            "within_contract": bool(abs(ll - ll_o) <= 1e-10 * abs(ll_o) and
                                    np.all(np.abs(g - g_o) <= 1e-8 * np.maximum(np.abs(g_o), 1e-3 * scale)))}


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


def working_set_bytes(w, n_gpus=1):
    """Bytes one evaluation touches per GPU: the walk's own global traffic plus the resident codes."""
    return walk_bytes(w["tree"], -(-w["S"] // n_gpus), w["K"], w["R"], True) + w["n_taxa"] * -(-w["S"] // n_gpus)


def workload_config(w, n_gpus, mp="group"):
    ws = working_set_bytes(w, n_gpus)
    l2 = (f"inputs larger than L2: one evaluation moves {ws / 1e9:.1f} GB per GPU (126 MB L2), nothing survives "
          f"from one step to the next" if ws > 4 * L2_BYTES else
          f"working set {ws / 1e6:.0f} MB per GPU fits the 126 MB L2: L2 is flushed (256 MB memset) before every "
          f"timed call")
    return {"workload": f"{w['name']}: {w['n_taxa']} taxa x {w['S']} sites, K={w['K']} "
                        f"({'GTR' if w['K'] == 4 else 'Restriction'}), R={w['R']} discrete-Gamma categories, "
                        f"logpdf+gradient per step",
            "n_taxa": w["n_taxa"], "sites": w["S"], "states": w["K"], "rate_categories": w["R"],
            "sharding": f"sites/{n_gpus}", "l2": l2,
            "host": ("one process owns all GPUs (mcp_create_multi)" if mp == "group" else
                     "one process per GPU (mcp_create_rank)") if n_gpus > 1 else "one process, one GPU (mcp_create)"}


class L2Flusher:
    """Writes a buffer larger than L2 on the device between timed calls (small workloads only)."""

    def __init__(self, device):
        import torch
        self.torch = torch
        self.buf = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{device}")

    def __call__(self):
        self.buf.fill_(1)
        self.torch.cuda.synchronize()


def timed_calls(ctx, fn, steps, warmup, flush=None):
    """Runs fn(i) warmup + steps times; each timed call is bracketed by the library's device stopwatch
    (CUDA events on the evaluation streams).  Returns (per-call ms list, per-call walk-kernel ms list)."""
    for i in range(warmup):
        fn(i)
    ms, walk = [], []
    for i in range(steps):
        if flush is not None:
            flush()
        ctx.timer_start()
        fn(i)
        ms.append(ctx.timer_stop())
        walk.append(ctx.stats()["walk_ms"])
    return ms, walk


def leg_single(name, local_rank, steps, warmup, with_cpu=True):
    """Short leg of a single-tree BASELINE config (cfg2 / cfg3) on one GPU, alignment resident, through
    mcp_eval with the argument arrays of a compiled host (packed once; only branch lengths change)."""
    import mcphylo_jl_b200 as mcp
    from mcphylo_jl_b200 import capi
    from mcphylo_jl_b200.phylodist import _tree_args

    w = make_workload(name)
    codes, leaf_nums = make_codes(w, 0, w["S"])
    ctx = capi.Context(local_rank)
    aln = ctx.alignment_from_codes(codes, w["K"], leaf_nums)
    d = mcp.PhyloDist(w["tree"], w["pi"], w["srates"], w["rates"], w["model"])
    ft, targs = _tree_args(d)
    prep = capi.PreparedBatch(ctx, [aln], [targs], want_grad=True)
    blv0 = ft.blv.copy()
    small = working_set_bytes(w) <= 4 * L2_BYTES
    flush = L2Flusher(local_rank) if small else None

    def step(i):
        prep.set_blv(0, blv0 * (1.0 + 1e-3 * (i % 7)))
        return prep.eval()

    ms, walk = timed_calls(ctx, step, steps, warmup, flush)
    # the same calls back to back without the flush and with host wall-clock: API overhead on top of the device time
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    wall_ms = (time.perf_counter() - t0) * 1e3 / steps
    ms_nf, _ = timed_calls(ctx, step, steps, 0, None)
    st = ctx.stats()
    wk = float(np.median(walk))
    # what a production caller sees: mcp_set_timing(0) -- the four per-evaluation timing events behind mcp_get_stats are
    # not recorded (julia/MCPhyloB200.jl turns them off); kernel_ms above needs them, hence both runs
    ctx.set_timing(False)
    ms_nt, _ = timed_calls(ctx, step, steps, 5, flush)
    ms_nt_nf, _ = timed_calls(ctx, step, steps, 0, None)
    ctx.set_timing(True)
    out = {"config": workload_config(w, 1), "value": 1e3 / float(np.median(ms_nt)), "unit": "evals/s",
           "value_what": "1 / ms_per_call_device_no_timing_events (L2 flushed before every call when the working set fits in L2)",
           "ms_per_call_device_no_timing_events": float(np.median(ms_nt)),
           "ms_per_call_device_no_timing_events_l2_warm": float(np.median(ms_nt_nf)),
           "ms_per_call_device": float(np.median(ms)), "ms_per_call_device_l2_warm": float(np.median(ms_nf)),
           "ms_per_call_host_wall_l2_warm": wall_ms, "api_overhead_ms": wall_ms - float(np.median(ms_nf)),
           "kernel_ms": wk, "kernel_launches_per_call": st["kernel_launches"],
           "launch": {"grid": st["grid"], "block": st["block"], "columns_per_thread": st["columns_per_thread"], "tiles": st["tiles"],
                      "operand_ring": st.get("operand_ring", 0)},
           "roofline": roofline_block(w, w["S"], wk, w["tree"]),
           "oracle_window": oracle_window(ctx, w, codes, leaf_nums, st, sites=256),
           "steps": steps, "warmup": warmup}
    if small:
        out["roofline"]["note"] = "latency-bound shape (SURVEY 8d): fractions are reported for completeness, no roofline claim"
    if with_cpu:
        sample = min(w["S"], 20000 if name == "cfg2" else 5000)
        rate, sec, threads, s = cpu_oracle_rate(w, codes, leaf_nums, sample, 2, 1)
        out["cpu_baseline"] = {"value": rate, "unit": "evals/s", "cores": threads, "kind": "port",
                               "sample": f"first {s} of {w['S']} sites; {sec:.3f} s per evaluation on the sample"}
    try:    # logL + branch gradient + gradient w.r.t. the substitution-model parameters in one call
        from mcphylo_jl_b200.substitution_models import model_derivatives
        _, dA, dpi = model_derivatives(w["model"], w["pi"], w["srates"])
        mg_ms, res = [], None
        for i in range(4):
            ctx.timer_start()
            res = ctx.eval_model_gradient(aln, *targs, dA=dA, dpi=dpi)
            mg_ms.append(ctx.timer_stop())
        ll_p, g_p = ctx.eval(aln, *targs, want_grad=True)
        out["model_gradient"] = {"ms_per_call_device": float(np.median(mg_ms[1:])), "n_parameters": int(dA.shape[2]),
                                 "ll_rel_diff_to_plain": abs(res[0] - ll_p) / abs(ll_p),
                                 "grad_max_rel_diff_to_plain": float(np.max(np.abs(res[1] - g_p)) / np.max(np.abs(g_p))),
                                 "what": "mcp_eval_model_gradient: d logL / d (base_freq, substitution_rates) next to the branch gradient"}
    except Exception as ex:
        out["model_gradient"] = {"error": repr(ex)}
    aln.close()
    ctx.close()
    return out


def leg_batch(local_rank, steps, warmup, n_trees=CFG5_TREES, sites=None, with_cpu=True):
    """cfg5: T independent trees, each with its own alignment, evaluated by ONE mcp_eval_batch call per
    step (the MultiplePhyloDist path), argument arrays packed once as a compiled host would hold them."""
    import mcphylo_jl_b200 as mcp
    from mcphylo_jl_b200 import capi
    from mcphylo_jl_b200.phylodist import _tree_args

    n_taxa, S, K, R, tseed, dseed = WORKLOADS["cfg5"]
    S = sites or S
    pi = RESTRICTION_PI
    model_out = mcp.Restriction(pi, [])
    ctx = capi.Context(local_rank)
    trees, alns, targs_all, blv0 = [], [], [], []
    t_gen = time.perf_counter()
    first = None
    for i in range(n_trees):
        tree = mcp.random_tree(n_taxa, np.random.default_rng(tseed + i))
        rng = np.random.default_rng(dseed * 1000 + i)
        pool_n = min(S, 2048)
        pool, leaf_nums = mcp.simulate_codes(tree, model_out, pi, np.ones(1), pool_n, rng, gap_frac=0.01)
        codes = pool if pool_n == S else np.take(pool, rng.integers(0, pool_n, size=S), axis=1)
        if first is None:
            first = (tree, codes, leaf_nums)
        trees.append(tree)
        alns.append(ctx.alignment_from_codes(codes, K, leaf_nums))
        ft, ta = _tree_args(mcp.PhyloDist(tree, pi, [0.0], [1.0], mcp.Restriction))
        targs_all.append(ta)
        blv0.append(ft.blv.copy())
    t_gen = time.perf_counter() - t_gen
    prep = capi.PreparedBatch(ctx, alns, targs_all, want_grad=True)

    def step(i):
        f = 1.0 + 1e-3 * (i % 7)
        for t in range(n_trees):
            prep.set_blv(t, blv0[t] * f)
        return prep.eval()

    ms, walk = timed_calls(ctx, step, steps, warmup, None)
    st = ctx.stats()
    wk = float(np.median(walk))
    w = dict(name="cfg5", tree=first[0], S=S, K=K, R=R, pi=pi, srates=np.zeros(1), model=mcp.Restriction, rates=np.ones(1),
             n_taxa=n_taxa, dseed=dseed)
    rl = roofline_block(w, S, wk, first[0])
    for key in ("achieved", "frac", "walk_bytes", "frac_walk", "achieved_walk", "algorithmic_bytes_per_launch"):
        if rl.get(key) is not None:
            rl[key] = rl[key] * n_trees      # one launch processes all trees (first tree's program taken as typical)
    ll, grads = step(0)
    out = {"config": {"workload": f"cfg5: {n_trees} trees x {n_taxa} taxa x {S} binary sites (Restriction), one batched "
                                  f"logpdf+gradient call (mcp_eval_batch) per step", "trees": n_trees, "n_taxa": n_taxa,
                      "sites": S, "states": K, "rate_categories": R,
                      "l2": "inputs larger than L2 (256 alignments, 1.3 GB of codes, 30 GB of partial traffic per step)"},
           "value": n_trees * 1e3 / float(np.median(ms)), "unit": "tree-evals/s", "ms_per_step": float(np.median(ms)),
           "kernel_ms": wk, "kernel_launches_per_call": st["kernel_launches"],
           "launch": {"grid": st["grid"], "block": st["block"], "columns_per_thread": st["columns_per_thread"], "tiles": st["tiles"],
                      "operand_ring": st.get("operand_ring", 0)},
           "roofline": rl, "steps": steps, "warmup": warmup, "setup": {"alignment_generate_s": t_gen},
           "result_check": {"sum_logL": float(np.sum(ll)), "finite": bool(all(np.all(np.isfinite(g)) for g in grads))}}
    w1 = dict(w)
    out["oracle_window"] = oracle_window(ctx, w1, first[1], first[2], st, sites=256)
    if with_cpu:
        sample = 5000
        rate, sec, threads, s = cpu_oracle_rate(w1, first[1], first[2], sample, 2, 1)
        out["cpu_baseline"] = {"value": rate, "unit": "tree-evals/s", "cores": threads, "kind": "port",
                               "sample": f"one of the {n_trees} trees, first {s} of {S} sites; {sec:.3f} s on the sample"}
    for a in alns:
        a.close()
    ctx.close()
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist

    import mcphylo_jl_b200 as mcp
    from mcphylo_jl_b200 import capi
    from mcphylo_jl_b200.phylodist import _tree_args

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback exists)")
    n_gpus = max(world, 1)
    if world == 1 and args.gpus > 1:
        n_gpus = args.gpus          # plain `python bench.py --gpus N`: one process, N GPUs (group mode only)
        args.mp = "group"
    group_mode = args.mp == "group"
    if world > 1:
        # the launcher's process group is plumbing only (barriers, the NCCL unique id): gloo, no GPU
        dist.init_process_group("gloo")

    def barrier():
        if world > 1:
            dist.barrier()

    if group_mode and rank != 0:
        # rank 0 drives every GPU through ONE multi-device context; this rank only keeps the launcher's
        # barriers (before / after each timed region) company
        for _ in range(4):
            barrier()
        dist.barrier()
        dist.destroy_process_group()
        return

    reduce_mode = {"auto": capi.REDUCE_AUTO, "nccl": capi.REDUCE_NCCL, "peer": capi.REDUCE_PEER, "host": capi.REDUCE_HOST}[args.reduce]
    w = make_workload(args.workload, args.sites)
    with stdout_to_stderr():
        if group_mode:
            lo, hi = 0, w["S"]
            torch.cuda.set_device(0)
            ctx = capi.Context(devices=list(range(n_gpus)), reduce=reduce_mode) if n_gpus > 1 else capi.Context(0)
            primary = 0
        else:
            lo, hi = capi.shard_bounds(w["S"], world, rank)
            torch.cuda.set_device(local_rank)
            uid = [capi.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            ctx = capi.Context(local_rank, rank=(world, rank, uid[0]))
            primary = local_rank
    t_gen = time.perf_counter()
    codes, leaf_nums = make_codes(w, lo, hi)
    t_gen = time.perf_counter() - t_gen
    if args.block or args.ctas_per_sm:
        ctx.set_launch(args.block, args.ctas_per_sm)
    if args.cpt:
        ctx.set_columns_per_thread(args.cpt)
    tree = w["tree"]
    blv0 = mcp.get_branchlength_vector(tree)
    K = w["K"]

    def dist_for(step):
        # new branch lengths every step (the leapfrog pattern): nothing can be reused across steps
        mcp.set_branchlength_vector(tree, blv0 * (1.0 + 1e-3 * (step % 7)))
        return mcp.PhyloDist(tree, w["pi"], w["srates"], w["rates"], w["model"])

    # alignment upload (one-off, reported separately)
    t_up = time.perf_counter()
    aln = ctx.alignment_from_codes(codes, K, leaf_nums)
    ft0, targs0 = _tree_args(dist_for(0))
    with stdout_to_stderr():
        ll0, grad0 = ctx.eval(aln, *targs0, want_grad=True)
    t_up = time.perf_counter() - t_up

    sampler = ClockSampler(primary)

    # ---- device-resident throughput -----------------------------------------------------
    # Host inputs of every step (flattened tree with that step's branch lengths + the model's
    # eigendecomposition) are prepared before the clock starts: `value` times the C-ABI evaluation
    # (parameter upload, 3 kernels per device, all-reduce, result download); the Python-side tree
    # traversal of the host mirror is part of `e2e` below.
    prepared = []
    for i in range(max(args.steps, args.warmup)):
        ft, targs = _tree_args(dist_for(i))
        prepared.append(targs)
    for i in range(args.warmup):
        ctx.eval(aln, *prepared[i], want_grad=True)
    walk_ms, launches = [], 0
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    ctx.timer_start()
    t_wall = time.perf_counter()
    for i in range(args.steps):
        ll, grad = ctx.eval(aln, *prepared[i], want_grad=True)
        st = ctx.stats()
        walk_ms.append(st["walk_ms"])
        launches += st["kernel_launches"]
    t_wall = (time.perf_counter() - t_wall) * 1e3
    ms_total = ctx.timer_stop()
    torch.cuda.synchronize()
    barrier()
    sampler.stop()
    stats = ctx.stats()
    member_stats = [ctx.stats(g) for g in range(ctx.device_count)] if group_mode and n_gpus > 1 else [stats]

    # ---- end to end through the C ABI with HOST buffers --------------------------------------
    # mcp_eval_streamed: every step uploads the whole alignment from pinned host memory (inside the library:
    # each GPU's site range in blocks, block b+1 crossing PCIe while block b is evaluated), the host flattens
    # the tree and runs the eigendecomposition, the result is reduced over the GPUs and read back.
    pinned = torch.from_numpy(codes).pin_memory()
    S_local = codes.shape[1]

    def e2e_step(i):
        ft, targs = _tree_args(dist_for(i))
        return ctx.eval_streamed(pinned.data_ptr(), K, S_local, leaf_nums, *targs, want_grad=True)

    for i in range(min(args.warmup, 3)):
        e2e_step(i)
    e2e_launches = 0
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    ctx.timer_start()
    t_e2e_wall = time.perf_counter()
    for i in range(args.steps):
        ll_e, grad_e = e2e_step(i)
    t_e2e_wall = (time.perf_counter() - t_e2e_wall) * 1e3
    ms_e2e = ctx.timer_stop()
    torch.cuda.synchronize()
    barrier()
    sampler.stop()
    blocks = [ctx.stream_blocks(g) for g in range(ctx.device_count)]
    timeline = [ctx.stream_timeline(g) for g in range(ctx.device_count)]
    # per GPU and launch over a site range: parameter staging + branch tables + walk + final reduction; one more
    # kernel adds the block results when a range was evaluated in several launches
    e2e_launches = args.steps * sum(4 * len(b) + (1 if len(b) > 1 else 0) for b in blocks)

    if world > 1 and not group_mode:
        t = torch.tensor([ms_total, ms_e2e, t_wall, t_e2e_wall], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e, t_wall, t_e2e_wall = (float(v) for v in t)

    if rank == 0:
        NN = 2 * w["n_taxa"] - 1
        I = w["n_taxa"] - 1
        evals_per_s = args.steps / (ms_total * 1e-3)
        e2e_per_s = args.steps / (ms_e2e * 1e-3)
        local_S = -(-w["S"] // n_gpus)
        wk = float(np.mean(walk_ms)) if walk_ms and all(m > 0 for m in walk_ms) else None
        roofline = roofline_block(w, local_S, wk, tree)
        clocks = sampler.summary()
        roofline["issue"] = issue_block(w, local_S, wk, clocks)
        # parity stamp at the launch shape of the timed run (member 0's shape on a multi-device context)
        check_ctx = capi.Context(primary)
        window = oracle_window(check_ctx, w, codes, leaf_nums, member_stats[0], sites=256)
        check_ctx.close()
        line = {
            "metric": "logpdf+gradient evaluations/s", "value": evals_per_s, "unit": "evals/s",
            "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": f"synthetic: {min(w['S'], SIM_POOL)} columns simulated down a random {w['n_taxa']}-taxon tree "
                    f"under the evaluating model (1% gaps), bootstrap-resampled to {w['S']} sites",
            "config": workload_config(w, n_gpus, args.mp),
            "site_node_updates_per_s": (I + NN - 1) * w["S"] * w["R"] * evals_per_s,
            "site_node_updates_note": "post-order node updates I*S*R plus branch-gradient updates (NN-1)*S*R per evaluation",
            "roofline": roofline,
            "e2e": {"value": e2e_per_s, "unit": "evals/s", "ms_per_step": ms_e2e / args.steps,
                    "ms_per_step_host_wall": t_e2e_wall / args.steps,
                    "h2d_bytes_per_step": int(w["S"] * w["n_taxa"] + stats["h2d_bytes"]),
                    "d2h_bytes_per_step": int(NN * 8),
                    "site_blocks_per_gpu": [[b - a for a, b in bl] for bl in blocks],
                    "last_step_timeline_ms_gpu0": timeline[0],
                    "timeline_what": "per site block on GPU 0: transfer begin, transfer end, evaluation enqueued, walk "
                                     "begin, walk end (CUDA events, ms since the first transfer of the step began)",
                    "what": "per step through mcp_eval_streamed: alignment codes re-uploaded from pinned host memory on the "
                            "copy stream of every GPU in transfer units of 2 k .. 64 k sites, each followed by its ready flags; "
                            "ONE walk launch per GPU starts before the data has arrived, takes tiles by atomic ticket in site "
                            "order and waits per tile for the flags; tree flattened and eigendecomposition on the host; "
                            "results reduced over the GPUs, read back"},
            "ms_per_step_host_wall": t_wall / args.steps,
            "gpu_launches": int(launches + e2e_launches),
            "gpu_launches_timed": int(launches),
            "clocks": clocks,
            "launch": {"grid": stats["grid"], "block": stats["block"], "columns_per_thread": stats["columns_per_thread"],
                       "tiles": stats["tiles"], "scratch_bytes": stats["scratch_bytes"],
                       "operand_ring": stats.get("operand_ring", 0),
                       "walk_ms_per_gpu": [s["walk_ms"] for s in member_stats],
                       "reduce": capi.REDUCE_NAMES.get(ctx.reduce_mode, "none") if n_gpus > 1 else "none"},
            "setup": {"alignment_generate_s": t_gen, "upload_and_first_eval_s": t_up,
                      "codes_bytes_per_gpu": int(w["n_taxa"] * local_S)},
            "value_what": "per step: mcp_eval from pre-flattened host arrays (tree arrays, branch lengths, model uploaded "
                          "every step), reduction of [logL, grad] over the GPUs inside the library, result read back; "
                          "alignment resident; timed with CUDA events on the library's evaluation streams "
                          "(mcp_timer_start/stop), slowest GPU",
            "result_check": {"logL_at_initial_branch_lengths": ll0, "grad_l2_at_initial_branch_lengths":
                             float(np.linalg.norm(grad0)), "grad_finite": bool(np.all(np.isfinite(grad))),
                             "e2e_matches_resident": bool(abs(ll_e - ll) <= 1e-9 * abs(ll)),
                             "oracle_window": window},
        }
        if n_gpus == 1 and not args.no_cpu_baseline:
            sample = args.cpu_sample or 10000
            rate, sec, threads, s = cpu_oracle_rate(w, codes, leaf_nums, sample, 2, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "evals/s", "cores": threads, "kind": "port",
                                    "sample": f"first {s} of {w['S']} sites, full tree; {sec:.3f} s per evaluation "
                                              f"on the sample (OpenMP, {threads} threads), extrapolated linearly in sites"}
        aln.close()
        ctx.close()
        if n_gpus == 1 and not args.no_extra and args.workload == "cfg4" and not args.sites:
            extra = {}
            for name in ("cfg2", "cfg3"):
                try:
                    extra[name] = leg_single(name, primary, steps=20 if name == "cfg3" else 200, warmup=5,
                                             with_cpu=not args.no_cpu_baseline)
                except Exception as ex:      # a failing side leg must not take the headline line down
                    extra[name] = {"error": repr(ex)}
            try:
                extra["cfg5"] = leg_batch(primary, steps=5, warmup=2, with_cpu=not args.no_cpu_baseline)
            except Exception as ex:
                extra["cfg5"] = {"error": repr(ex)}
            line["extra"] = {"configs": extra,
                             "what": "short legs of the other BASELINE configs on this GPU, alignment resident, argument "
                                     "arrays packed once (compiled-host call pattern); each carries its own oracle window"}
        print(json.dumps(line))
    else:
        aln.close()
        ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_side(args):
    """--workload cfg2 / cfg3 / cfg5 on their own (N = 1): the same legs the headline line carries."""
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if args.workload == "cfg5":
        leg = leg_batch(local_rank, args.steps, max(args.warmup, 3), n_trees=args.trees or CFG5_TREES, sites=args.sites or None,
                        with_cpu=not args.no_cpu_baseline)
        metric = "logpdf+gradient tree-evaluations/s (batched)"
    else:
        leg = leg_single(args.workload, local_rank, args.steps, max(args.warmup, 3), with_cpu=not args.no_cpu_baseline)
        metric = "logpdf+gradient evaluations/s"
    line = {"metric": metric, "n_gpus": 1, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic"}
    line.update(leg)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--mp", default="group", choices=["group", "ranks"],
                    help="N > 1: one process owning all GPUs (mcp_create_multi) or one process per GPU (mcp_create_rank)")
    ap.add_argument("--reduce", default="auto", choices=["auto", "nccl", "peer", "host"])
    ap.add_argument("--sites", type=int, default=0, help="override the number of sites (experiments only)")
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the cfg2/cfg3/cfg5 side legs")
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--cpt", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--trees", type=int, default=0, help="cfg5: number of trees in the batch")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in ("cfg2", "cfg3", "cfg5") and not args.sites and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        run_side(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
