#!/usr/bin/env python
"""Timing of the Scann façade tree modes that score every member of the probed leaves (csrc/ivf.cu):
Scann::search_partitioned (exact distances) and Scann::search_tree_ah variant B (f32 LUT over byte codes), on a
clustered synthetic set.  One JSON line per case with the HBM roofline fraction: algorithmic bytes per query =
Σ|probed leaf| x (D*4 B rows | S B codes), measured peak from MEASURED_PEAKS.json."""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--n", type=int, default=1_000_000)
    p.add_argument("--dim", type=int, default=96)
    p.add_argument("--partitions", type=int, default=1000)
    p.add_argument("--leaves", type=int, default=16)
    p.add_argument("--nq", type=int, default=2000)
    p.add_argument("--k", type=int, default=10)
    p.add_argument("--subspaces", type=int, default=48)
    p.add_argument("--codes", type=int, default=256)
    p.add_argument("--reps", type=int, default=3)
    a = p.parse_args()
    import torch
    pkg = importlib.import_module("scann-rust_b200")
    bench = importlib.import_module("bench")
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(42)
    lat = torch.randn((2048, a.dim), generator=g, device=dev)
    x = bench.make_points(torch, a.n, a.dim, lat, 0.5, 1.0, 42, dev)
    q = bench.make_points(torch, a.nq, a.dim, lat, 0.5, 1.0, 123, dev)
    ix = pkg.indexing
    centers = ix.kmeans(x[: min(a.n, 500_000)].contiguous(), a.partitions, 10, 7)
    K = centers.shape[0]
    assign = ix.assign_partitions(x, centers, 0)
    order = torch.argsort(assign.long(), stable=True)
    counts = torch.bincount(assign.long(), minlength=K)
    off = torch.zeros((K + 1,), dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(counts, 0)
    # non-residual codebook with `codes` centroids per block (sampled rows) + byte codes by datapoint id
    ds = a.dim // a.subspaces
    sel = torch.randperm(a.n, generator=g, device=dev)[: a.codes]
    cb = x[sel].reshape(a.codes, a.subspaces, ds).permute(1, 0, 2).contiguous()       # [S, C, ds]
    codes = torch.empty((a.n, a.subspaces), dtype=torch.uint8, device=dev)
    for s0 in range(0, a.n, 1 << 18):
        xb = x[s0:s0 + (1 << 18)].reshape(-1, a.subspaces, ds)
        d = ((xb[:, :, None, :] - cb[None]) ** 2).sum(3)                                   # [m, S, C]
        codes[s0:s0 + (1 << 18)] = d.argmin(2).to(torch.uint8)
    s = pkg.LeafScanSearcher(centers, order.to(torch.int32), off, x, cb, codes)
    tok = pkg.TreePartitioner(centers, 0).partition(q, a.leaves)[0]
    members = float(counts[tok.long()].sum()) / a.nq                                       # probed members per query
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))

    def run(name, fn, bytes_per_member):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        gbs = members * a.nq * bytes_per_member / (ms / 1e3) / 1e9
        print(json.dumps({"case": name, "n": a.n, "dim": a.dim, "partitions": K, "leaves": a.leaves, "nq": a.nq,
                          "members_per_query": members, "ms_per_batch": ms, "queries_per_s": a.nq / (ms / 1e3),
                          "algorithmic_GBps": gbs, "frac_of_measured_hbm_peak": gbs / peak}), flush=True)

    run("search_partitioned[exact Dot]", lambda: s.search_partitioned(q, a.k, a.leaves, pkg.DistanceMeasure.DotProduct),
        a.dim * 4)
    run(f"search_tree_ah[variant B, {a.codes} codes]", lambda: s.search_tree_ah(q, a.k, a.leaves), a.subspaces)


if __name__ == "__main__":
    main()
