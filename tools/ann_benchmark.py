#!/usr/bin/env python
"""ANN-Benchmarks-style runner with the reference CLI's flags and report (src/bin/ann_benchmark.rs:119-133,229-299):

    python tools/ann_benchmark.py --algorithm tree-ah --num-partitions 100 --partitions-to-search 10 --num-blocks 8

builds the index through the Scann façade (scann-rust_b200/scann.py -> the C ABI), runs the reference's SEQUENTIAL
`index.search` loop (ann_benchmark.rs:174-177) and prints the same human-readable block and the same `json:` line
(`BenchmarkReport`: dataset, algorithm, distance, k, train_size, test_size, dimension, build_seconds, search_seconds,
qps, recall_at_k, index_rss_delta_bytes).  Two fields are added after the reference's: `batched_qps` (one
`search_batched` call over all test queries — how a GPU is meant to be driven) and `device_bytes_delta`.

Synthetic data is U[0,1) like the reference's (`rng.gen::<f32>()`, :409-415) but drawn from numpy's generator: the
reference's `StdRng` stream is not reproducible outside Rust, so values differ while the distribution is the same.
Ground truth = exact squared-L2 top-k (:418-431); recall = average_recall_at_k (:452-471).
"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ALGORITHMS = {"brute-force": "brute_force", "brute_force": "brute_force", "bruteforce": "brute_force",
              "partitioned": "partitioned", "hashed": "hashed", "tree-ah": "tree_ah", "tree_ah": "tree_ah",
              "treeah": "tree_ah"}
DISTANCES = {"squared-l2": "squared_l2", "squared_l2": "squared_l2", "l2": "l2", "dot-product": "dot_product",
             "dot_product": "dot_product"}  # l1 / cosine: not on the GPU path (out of scope)


def average_recall_at_k(retrieved, ground_truth, k):
    """ann_benchmark.rs:452-471"""
    if len(retrieved) == 0 or len(ground_truth) == 0 or k == 0:
        return 0.0
    n = min(len(retrieved), len(ground_truth))
    total = 0.0
    for i in range(n):
        gt = set(int(v) for v in ground_truth[i][:k])
        total += sum(1 for v in retrieved[i][:k] if int(v) in gt) / k
    return total / n


def exact_ground_truth(train, test, k):
    """ann_benchmark.rs:418-431 (stable sort by squared L2: ties keep the lower index)"""
    out = []
    tn = (train.astype(np.float64) ** 2).sum(1)
    for q in test:
        d = tn - 2.0 * (train.astype(np.float64) @ q.astype(np.float64)) + float((q.astype(np.float64) ** 2).sum())
        out.append(np.argsort(d, kind="stable")[:k].astype(np.uint32))
    return out


def current_rss_bytes():
    try:
        with open("/proc/self/statm") as f:
            return int(f.read().split()[1]) * os.sysconf("SC_PAGE_SIZE")
    except Exception:
        return None


def run(args):
    import torch

    pkg = importlib.import_module("scann-rust_b200")
    M = pkg.DistanceMeasure
    measure = {"squared_l2": M.SquaredL2, "l2": M.L2, "dot_product": M.DotProduct}[args.distance]
    if args.data_json:
        with open(args.data_json) as f:
            js = json.load(f)
        train = np.asarray(js["train"], np.float32)[:args.limit_train]
        test = np.asarray(js["test"], np.float32)[:args.limit_test]
        gt = [np.asarray(g, np.uint32)[:args.k] for g in js["neighbors"]][:len(test)]
        source = os.path.basename(args.data_json)
    else:
        rng = np.random.default_rng(args.seed)
        train = rng.random((args.synthetic_train, args.dim), dtype=np.float32)
        test = rng.random((args.synthetic_test, args.dim), dtype=np.float32)
        gt = exact_ground_truth(train, test, args.k)
        source = f"synthetic_n{args.synthetic_train}_q{args.synthetic_test}_d{args.dim}"

    torch.cuda.synchronize()
    before_rss, before_dev = current_rss_bytes(), torch.cuda.mem_get_info()[0]
    t0 = time.perf_counter()
    b = pkg.ScannBuilder().num_neighbors(args.k).distance_measure(measure)
    if args.algorithm == "brute_force":
        index = b.brute_force().build(train)
    elif args.algorithm == "partitioned":
        index = b.tree(args.num_partitions, args.partitions_to_search).build(train)
    elif args.algorithm == "hashed":
        index = b.hash(args.num_blocks).build(train)
    else:
        index = b.tree(args.num_partitions, args.partitions_to_search).hash(args.num_blocks).build(train)
    torch.cuda.synchronize()
    build_seconds = time.perf_counter() - t0
    after_rss, after_dev = current_rss_bytes(), torch.cuda.mem_get_info()[0]

    index.search(test[0], args.k)  # warm-up (module load, workspace growth)
    t0 = time.perf_counter()
    retrieved = [[i for i, _ in index.search(q, args.k)] for q in test]  # the reference's sequential loop
    search_seconds = time.perf_counter() - t0
    t0 = time.perf_counter()
    ids, _, counts = index.search_batched(test, args.k)[:3]
    batched_seconds = time.perf_counter() - t0
    report = {
        "dataset": source, "algorithm": args.algorithm, "distance": args.distance, "k": args.k,
        "train_size": int(len(train)), "test_size": int(len(test)), "dimension": int(train.shape[1]),
        "build_seconds": build_seconds, "search_seconds": search_seconds,
        "qps": len(test) / search_seconds if search_seconds > 0 else 0.0,
        "recall_at_k": average_recall_at_k(retrieved, gt, args.k),
        "index_rss_delta_bytes": (after_rss - before_rss) if before_rss is not None and after_rss is not None
        and after_rss >= before_rss else None,
        "batched_qps": len(test) / batched_seconds if batched_seconds > 0 else 0.0,
        "batched_recall_at_k": average_recall_at_k([ids[i, :counts[i]] for i in range(len(test))], gt, args.k),
        "device_bytes_delta": int(before_dev - after_dev),
    }
    print("=== ANN-Benchmarks style report ===")
    print(f"dataset: {report['dataset']}")
    print(f"algorithm: {report['algorithm']}")
    print(f"distance: {report['distance']}")
    print(f"k: {report['k']}")
    print(f"train/test/dim: {report['train_size']}/{report['test_size']}/{report['dimension']}")
    print(f"build_seconds: {report['build_seconds']:.6f}")
    print(f"search_seconds: {report['search_seconds']:.6f}")
    print(f"qps: {report['qps']:.2f}")
    print(f"recall@{report['k']}: {report['recall_at_k']:.6f}")
    print("index_rss_delta_bytes: " + (str(report["index_rss_delta_bytes"])
                                       if report["index_rss_delta_bytes"] is not None else "unavailable"))
    print("json: " + json.dumps(report))
    return report


def parse(argv=None):
    p = argparse.ArgumentParser(description="ANN-Benchmarks-style runner (flags of the reference's ann_benchmark)")
    p.add_argument("--data-json")
    p.add_argument("--algorithm", default="brute-force", type=lambda s: ALGORITHMS.get(s) or p.error(f"unsupported algorithm: {s}"))
    p.add_argument("--distance", default="squared-l2", type=lambda s: DISTANCES.get(s) or p.error(f"unsupported distance: {s}"))
    p.add_argument("--k", type=int, default=10)
    p.add_argument("--num-partitions", type=int, default=100)
    p.add_argument("--partitions-to-search", type=int, default=10)
    p.add_argument("--num-blocks", type=int, default=8)
    p.add_argument("--limit-train", type=int, default=None)
    p.add_argument("--limit-test", type=int, default=None)
    p.add_argument("--synthetic-train", type=int, default=10_000)
    p.add_argument("--synthetic-test", type=int, default=200)
    p.add_argument("--dim", type=int, default=64)
    p.add_argument("--seed", type=int, default=42)
    return p.parse_args(argv)


if __name__ == "__main__":
    run(parse())
