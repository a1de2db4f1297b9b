#!/usr/bin/env python
"""BASELINE configs[4] (C5) on ONE shard of 8: Tree-AH over synthetic PQ codes only (no raw rows, no reorder).
1B x 96 codes sharded over 8 GPUs = 125M rows x 24 B codes per GPU; K = 65,536 partitions of which this shard owns
every 8th (8,192 leaves of 15,259 rows); nibbles i.i.d. uniform generated on the device; 10,000 Gaussian queries;
leaves_to_search sweep.  Prints one JSON line per L with queries/s of this shard's step (partition + worklist +
LUT16 scan + merge of approximate candidates) and the scan's algorithmic GB/s against the measured HBM peak.
The index (3 GB of codes) is far larger than L2, unlike C3's 240 MB."""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--rows", type=int, default=125_000_000)
    p.add_argument("--partitions", type=int, default=65_536)
    p.add_argument("--world", type=int, default=8)
    p.add_argument("--dim", type=int, default=96)
    p.add_argument("--subspaces", type=int, default=48)
    p.add_argument("--nq", type=int, default=10_000)
    p.add_argument("--k", type=int, default=10)
    p.add_argument("--reorder", type=int, default=100)
    p.add_argument("--leaves", default="16,32,64,128,256")
    p.add_argument("--reps", type=int, default=3)
    a = p.parse_args()
    import torch
    pkg = importlib.import_module("scann-rust_b200")
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(42)
    K, S, D = a.partitions, a.subspaces, a.dim
    centers = torch.randn((K, D), generator=g, device=dev)
    codebook = torch.randn((S, 16, D // S), generator=g, device=dev) * 0.3
    own = torch.arange(0, K, a.world, device=dev)                       # this shard's leaves
    per = a.rows // own.numel()
    n = per * own.numel()
    sizes = torch.zeros(K, dtype=torch.int64, device=dev)
    sizes[own] = per
    off = torch.zeros(K + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(sizes, 0)
    bpp = (S + 1) // 2
    packed = torch.empty((n, bpp), dtype=torch.uint8, device=dev)
    for s0 in range(0, n, 1 << 24):
        m = min(1 << 24, n - s0)
        packed[s0:s0 + m] = torch.randint(0, 256, (m, bpp), generator=g, device=dev, dtype=torch.uint8)
    ids = torch.arange(n, dtype=torch.int32, device=dev)
    g.manual_seed(123)
    q = torch.randn((a.nq, D), generator=g, device=dev)
    cfg = pkg.TreeXHybridConfig(num_partitions=K, partitions_to_search=64, use_residuals=True,
                                pre_reorder_multiplier=a.reorder / a.k)
    s = pkg.TreeXHybridSearcher(cfg, 0).build_from_index(centers, codebook, packed, ids, off, None)
    del packed
    torch.cuda.empty_cache()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    for L in [int(v) for v in a.leaves.split(",")]:
        s.search_batched(q, a.k, partitions_to_search=L, pre_reorder_k=a.reorder)
        torch.cuda.synchronize()
        s.set_profiling(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            ids_o, d_o, c_o = s.search_batched(q, a.k, partitions_to_search=L, pre_reorder_k=a.reorder)
        e1.record()
        torch.cuda.synchronize()
        prof, _ = s.get_profile()
        s.set_profiling(False)
        by, pairs = s.last_scan_bytes()
        ms = e0.elapsed_time(e1) / a.reps
        scan_ms = prof["scan"] / a.reps
        gbs = by / (scan_ms / 1e3) / 1e9 if scan_ms > 0 else 0.0
        assert bool((d_o[:, 1:] >= d_o[:, :-1]).all())                  # sorted approximate distances
        print(json.dumps({"config": "C5 shard 1/%d" % a.world, "rows": n, "partitions": K, "leaves_to_search": L,
                          "nq": a.nq, "ms_per_step": ms, "shard_queries_per_s": a.nq / (ms / 1e3),
                          "stage_ms": {k2: v / a.reps for k2, v in prof.items()}, "pairs_on_shard": pairs,
                          "scan_algorithmic_GBps": gbs, "frac_of_measured_hbm_peak": gbs / peak,
                          "mean_results": float(c_o.float().mean())}), flush=True)


if __name__ == "__main__":
    main()
