#!/usr/bin/env python
"""Replays the call pattern of MCPhylo's tree-space leapfrog on the GPU path (measurement harness,
not product code): one gradlogpdf per leapfrog plus two logpdf calls for every attempted NNI, with
the topology changing when an NNI is accepted.

The caller logic restates /root/reference/src/samplers/tree_hamiltonian/refraction.jl:2-91
(`refraction!`, `ref_NNI!`) and auxilliary.jl:113-119 (`scale_fac`, `molifier`); the likelihood
calls go through the public PhyloDist API of this repo.  BASELINE config 2 shape by default:
50 taxa x 10 000 binary sites, Restriction model.

    python tools/pnuts_replay.py --leapfrogs 300 --epsilon 0.002
"""
import argparse
import copy
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import mcphylo_jl_b200 as mcp  # noqa: E402


def molifier(x, delta):
    return np.where(x >= delta, x, (x * x + delta * delta) / (2.0 * delta))


def scale_fac(x, delta):
    return np.where(x < delta, x / delta, 1.0)


class State:
    def __init__(self, x, r, g, lf):
        self.x, self.r, self.g, self.lf = x, r, g, lf


def ref_NNI(s, tmpB, epsilon, blv, delta, logf, rng, counters, logf_pair=None):
    po = mcp.post_order(s.x)
    intext = np.zeros(len(po) - 1)
    by_num = {n.num: n for n in po}
    for n in po:
        if not n.root and n.nchild > 0:
            intext[n.num - 1] = 1
    t = 0.0
    nni = att = 0
    while tmpB.min() <= 0.0:
        timelist = tmpB / np.abs(s.r)
        ref = int(np.argmin(timelist))
        temp = epsilon - t + timelist[ref]
        blv = np.abs(blv + temp * s.r)
        s.r[ref] *= -1.0
        if intext[ref] == 1:
            mcp.set_branchlength_vector(s.x, molifier(blv, delta))
            v_copy = copy.deepcopy(s.x)
            target = next(n for n in mcp.post_order(v_copy) if n.num == ref + 1)
            made = mcp.NNI(v_copy, target, bool(rng.integers(0, 2)))
            if made and logf_pair is not None:
                # both sides of the NNI in ONE batched launch (mcp_eval_batch, two trees, same alignment)
                U_before, U_after = logf_pair(s.x, v_copy)
                counters["logpdf"] += 2
                counters["batched_pairs"] += 1
            else:
                U_before = logf(s.x)
                counters["logpdf"] += 1
                if made:
                    U_after = logf(v_copy)
                    counters["logpdf"] += 1
            if made:
                att += 1
                delta_U = 2.0 * (U_before - U_after)
                my_v = s.r[ref] ** 2
                if my_v > delta_U:
                    nni += made
                    s.r[ref] = np.sqrt(my_v - delta_U)
                    s.x = v_copy
        t = epsilon + timelist[ref]
        tmpB = blv + (epsilon - t) * s.r
    return tmpB, nni, att


def refraction(s, epsilon, logfgrad, logf, delta, rng, counters, logf_pair=None):
    blenvec = mcp.get_branchlength_vector(s.x)
    s.r += (epsilon * 0.5) * s.g
    tmpB = blenvec + epsilon * s.r
    tmpB, nni, att = ref_NNI(s, tmpB, epsilon, blenvec, delta, logf, rng, counters, logf_pair)
    blenvec = molifier(tmpB, delta)
    mcp.set_branchlength_vector(s.x, blenvec)
    lf, grad = logfgrad(s.x)
    counters["gradlogpdf"] += 1
    grad = grad * scale_fac(blenvec, delta)
    s.r += (epsilon * 0.5) * grad
    s.g = grad
    s.lf = lf
    return nni, att


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--sites", type=int, default=0)
    ap.add_argument("--leapfrogs", type=int, default=300)
    ap.add_argument("--epsilon", type=float, default=0.002)
    ap.add_argument("--delta", type=float, default=0.003)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--batch-nni", action="store_true",
                    help="evaluate the two logpdf calls of an NNI attempt as one mcp_eval_batch launch")
    ap.add_argument("--prior", action="store_true",
                    help="target = likelihood + CompoundDirichlet(1, 1, 0.1, 1) branch-length prior (the reference's "
                         "sampler test, test/samplers/tree_samplers.jl:22), evaluated in the same device call "
                         "(mcp_eval_posterior) as logpdfgrad!(::Type{provided}) would combine them")
    a = ap.parse_args()
    prior = mcp.CompoundDirichlet(1.0, 1.0, 0.100, 1.0) if a.prior else None
    w = bench.make_workload(a.workload, a.sites)
    codes, leaf_nums = bench.make_codes(w, 0, w["S"])
    aln = mcp.DeviceAlignment(codes, leaf_nums, w["K"])
    rng = np.random.default_rng(a.seed)
    ctx = mcp.get_context(0)
    dev = {"walk_ms": 0.0, "device_ms": 0.0, "rebuilds": 0, "api_s": 0.0}

    def timed(fn):
        def wrapped(*args):
            t0 = time.perf_counter()
            out = fn(*args)
            dev["api_s"] += time.perf_counter() - t0
            return out
        return wrapped

    def note():
        st = ctx.stats()
        dev["walk_ms"] += st["walk_ms"]
        dev["device_ms"] += st["device_ms"]
        dev["rebuilds"] += st["schedule_rebuilt"]

    @timed
    def logf(tree):
        pd = mcp.PhyloDist(tree, w["pi"], w["srates"], w["rates"], w["model"])
        v = mcp.logpdfgrad(pd, aln, prior, want_grad=False)[0] if prior else mcp.logpdf(pd, aln)
        note()
        return v

    @timed
    def logfgrad(tree):
        pd = mcp.PhyloDist(tree, w["pi"], w["srates"], w["rates"], w["model"])
        v = mcp.logpdfgrad(pd, aln, prior) if prior else mcp.gradlogpdf(pd, aln)
        note()
        return v

    @timed
    def logf_pair(t1, t2):
        mpd = mcp.MultiplePhyloDist([t1, t2], w["pi"], w["srates"], w["rates"], w["model"])
        v = mcp.phylodist._multi(mpd, [aln, aln], False, None)[0]
        note()
        if prior:   # the batched entry point carries no prior: add the O(NN) host mirror
            return float(v[0]) + mcp.logpdf(prior, t1), float(v[1]) + mcp.logpdf(prior, t2)
        return float(v[0]), float(v[1])

    tree = w["tree"]
    lf, g = logfgrad(tree)
    n = g.size
    counters = {"gradlogpdf": 0, "logpdf": 0, "batched_pairs": 0}
    s = State(tree, rng.standard_normal(n), g * scale_fac(mcp.get_branchlength_vector(tree), a.delta), lf)
    H0 = -s.lf + 0.5 * float(s.r @ s.r)
    for k in dev:
        dev[k] = 0
    nni = att = 0
    t0 = time.perf_counter()
    for i in range(a.leapfrogs):
        dn, da = refraction(s, a.epsilon, logfgrad, logf, a.delta, rng, counters, logf_pair if a.batch_nni else None)
        nni += dn
        att += da
    wall = time.perf_counter() - t0
    H1 = -s.lf + 0.5 * float(s.r @ s.r)
    calls = counters["gradlogpdf"] + counters["logpdf"] - counters["batched_pairs"]   # library calls
    print(json.dumps({
        "workload": f"{a.workload}: {w['n_taxa']} taxa x {w['S']} sites, K={w['K']}, R={w['R']}; tree-space leapfrog replay",
        "target": "likelihood + CompoundDirichlet(1,1,0.1,1) prior, one device call" if prior else "likelihood",
        "leapfrogs": a.leapfrogs, "epsilon": a.epsilon, "gradlogpdf_calls": counters["gradlogpdf"],
        "logpdf_calls": counters["logpdf"], "batched_nni_pairs": counters["batched_pairs"],
        "nni_attempted": att, "nni_accepted": nni,
        "schedule_rebuilds": dev["rebuilds"],
        "us_per_leapfrog_wall": wall / a.leapfrogs * 1e6,
        "us_per_call_wall": wall / calls * 1e6,
        "us_per_call_api": dev["api_s"] / calls * 1e6,
        "us_per_call_device": dev["device_ms"] / calls * 1e3,
        "us_per_call_walk_kernel": dev["walk_ms"] / calls * 1e3,
        "leapfrogs_per_s": a.leapfrogs / wall,
        "hamiltonian_drift": H1 - H0, "final_logL": s.lf,
        "note": "wall includes this harness's Python tree handling (refraction loop, deepcopy per NNI attempt); api = time inside the PhyloDist calls (PhyloDist construction, flatten, eigendecomposition, library call)",
    }))


if __name__ == "__main__":
    main()
