#!/usr/bin/env python
"""Where the end-to-end step spends its time: resident evaluation, streamed evaluation with and without the
host-side tree flattening, and the bare upload, each as wall-clock per call (all calls are synchronous)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--sites", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--reduce", default="auto")
    args = ap.parse_args()
    import torch
    import mcphylo_jl_b200 as mcp
    from mcphylo_jl_b200 import capi
    from mcphylo_jl_b200.phylodist import _tree_args

    w = bench.make_workload("cfg4", args.sites)
    codes, leaf_nums = bench.make_codes(w, 0, w["S"])
    red = {"auto": 0, "nccl": 1, "peer": 2, "host": 3}[args.reduce]
    with bench.stdout_to_stderr():
        ctx = capi.Context(devices=list(range(args.gpus)), reduce=red) if args.gpus > 1 else capi.Context(0)
    aln = ctx.alignment_from_codes(codes, w["K"], leaf_nums)
    pinned = torch.from_numpy(codes).pin_memory()
    d = mcp.PhyloDist(w["tree"], w["pi"], w["srates"], w["rates"], w["model"])
    ft, targs = _tree_args(d)
    out = {"gpus": args.gpus, "sites": w["S"]}

    def wall(fn, reps=args.reps, warm=2):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append((time.perf_counter() - t0) * 1e3)
        return {"median_ms": float(np.median(ts)), "min_ms": float(np.min(ts)), "all": [round(t, 3) for t in ts]}

    with bench.stdout_to_stderr():
        out["resident_eval"] = wall(lambda: ctx.eval(aln, *targs, want_grad=True))
        out["resident_walk_ms"] = [ctx.stats(g)["walk_ms"] for g in range(ctx.device_count)]
        out["flatten_and_model_host"] = wall(lambda: _tree_args(d))
        out["streamed_prepared_args"] = wall(lambda: ctx.eval_streamed(pinned.data_ptr(), w["K"], w["S"], leaf_nums, *targs, want_grad=True))
        out["streamed_timeline_ms"] = [ctx.stream_timeline(g) for g in range(ctx.device_count)]
        for mode in ("noupload", "serial"):
            os.environ["MCPHYLO_B200_STREAM_PROBE"] = mode
            out[f"streamed_probe_{mode}"] = wall(lambda: ctx.eval_streamed(pinned.data_ptr(), w["K"], w["S"], leaf_nums, *targs, want_grad=True))
        os.environ.pop("MCPHYLO_B200_STREAM_PROBE")
        out["streamed_blocks"] = [ctx.stream_blocks(g) for g in range(ctx.device_count)]
        out["streamed_with_flatten"] = wall(lambda: ctx.eval_streamed(pinned.data_ptr(), w["K"], w["S"], leaf_nums, *_tree_args(d)[1], want_grad=True))

        def upload_only():
            aln.update_codes(pinned.data_ptr())
            ctx.synchronize()
        out["upload_only_resident_alignment"] = wall(upload_only)
        # a full upload of a SECOND alignment in flight while the resident one is evaluated
        aln2 = ctx.alignment_from_codes(codes, w["K"], leaf_nums)
        rows = []
        for _ in range(4):
            ctx.synchronize()
            t0 = time.perf_counter()
            aln2.update_codes(pinned.data_ptr())
            ctx.eval(aln, *targs, want_grad=True)
            t1 = time.perf_counter()
            ctx.synchronize()
            t2 = time.perf_counter()
            rows.append({"eval_wall_ms": (t1 - t0) * 1e3, "until_upload_done_ms": (t2 - t0) * 1e3, "walk_ms": ctx.stats()["walk_ms"]})
        out["upload_concurrent_with_eval"] = rows
        aln2.close()
        out["resident_eval_after"] = wall(lambda: ctx.eval(aln, *targs, want_grad=True))
        # device stopwatch vs wall over a run of K calls
        for K in (1, 5):
            t0 = time.perf_counter()
            ctx.timer_start()
            for _ in range(K):
                ctx.eval(aln, *targs, want_grad=True)
            dev = ctx.timer_stop()
            out[f"timer_{K}_calls"] = {"device_ms": dev, "wall_ms": (time.perf_counter() - t0) * 1e3}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
