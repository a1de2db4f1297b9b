#!/usr/bin/env python
"""Small end-to-end pass over every kernel family (a quick post-build sanity run; compute-sanitizer is closed on this pool):
Tree-AH (plain, split, filtered), partition on tcgen05, brute force f32/int8 on tcgen05 + radius + CUDA-core fallback,
leaf-scan modes, merge.  Sizes are tiny; results are checked against the oracle where cheap."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    pkg = importlib.import_module("scann-rust_b200")
    import helpers
    import oracle
    x, _ = helpers.clustered(6000, 32, 24, 0.35, 1)
    q = (x[:40] + 0.02).astype(np.float32)
    idx = helpers.build_index(oracle, x, 300, 8, iters=2)          # K = 300 >= 256: tcgen05 centroid scoring
    K = len(idx["centers"])
    s = pkg.TreeXHybridSearcher(pkg.TreeXHybridConfig(num_partitions=K, partitions_to_search=20))
    s.build_from_index(idx["centers"], idx["codebook"], idx["packed"], idx["ids"], idx["part_offsets"], x)
    ids, d, c = s.search_batched(q, 10, pre_reorder_k=50)
    assert (ids[:, 0] == np.arange(40)).mean() > 0.9
    allowed = np.random.default_rng(0).random(len(x)) < 0.3
    fi, fd, fc = s.search_with_filter(q, 10, allowed, pre_reorder_k=50)
    assert allowed[fi[fi != 0xFFFFFFFF]].all()
    qd = torch.tensor(q).cuda()
    tok = s.partition_tokens(qd, 20)
    tau = s.search_begin(qd, 10, pre_reorder_k=50, tokens=tok)
    si, sd, sc = s.search_end(tau)
    torch.cuda.synchronize()
    assert (si.cpu().numpy().view(np.uint32) == ids).all()
    mi, md, mc = pkg.merge_topk_packed(torch.stack([torch.stack([si, sd.view(torch.int32)])] * 2).contiguous())
    part = pkg.TreePartitioner(idx["centers"])
    t2, _ = part.partition(q, 20)
    assert (t2.view(np.uint32) == tok.cpu().numpy().view(np.uint32)).all()
    db = helpers.gaussian(3000, 96, 2)
    qq = helpers.gaussian(33, 96, 3)
    for m in (pkg.DistanceMeasure.SquaredL2, pkg.DistanceMeasure.DotProduct):
        bf = pkg.BruteForceSearcher(db, m)
        bi, bd, bc = bf.search_batched(qq, 10)
        assert bf.path_stats()[0] >= 1
        rc, gi, gd, gc = oracle.bf_search(db, qq, 10, oracle.SQL2 if m == pkg.DistanceMeasure.SquaredL2 else oracle.DOT)
        assert (bd.view(np.uint32) == gd.view(np.uint32)).all()
    r = pkg.BruteForceSearcher(db).search_radius(qq[0], 120.0, 512)
    codes, cal = pkg.scalar_quantize(db)
    sq = pkg.ScalarQuantizedBruteForceSearcher.from_quantized(codes, float(cal[2]), pkg.DistanceMeasure.DotProduct)
    sq.search_batched(qq, 10)
    os.environ["SCANN_BF_NO_TC"] = "1"
    pkg.BruteForceSearcher(db).search_batched(qq, 10)
    os.environ["SCANN_BF_NO_TC"] = "0"
    big = helpers.gaussian(500, 300, 4)                             # dim > 256: CUDA-core path
    pkg.BruteForceSearcher(big).search_batched(big[:5], 3)
    order = idx["ids"]
    leaf = pkg.LeafScanSearcher(idx["centers"], order, idx["part_offsets"], x)
    li, ld, lc = leaf.search_partitioned(q, 5, 4)
    assert (li[:, 0] == np.arange(40)).mean() > 0.9
    print("sanity_small ok")


if __name__ == "__main__":
    main()
