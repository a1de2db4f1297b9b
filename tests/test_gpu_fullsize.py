"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle cannot finish these sizes in
seconds): sortedness, distances equal to an independent recomputation, agreement with an independent exact top-k
(torch fp32 matmul — used here only as an independent ranking, distances are compared through a tolerance of
1e-4 relative, BASELINE north_star), idempotence, recall against exact ground truth, and the LUT16 candidates'
upper-bound property.  C1: 10k x 128 / 1k queries; C2: 1M x 128 / 10k queries (f32 Dot and int8); C3: 10M x 96
Tree-AH (K = 2000, S = 48, L = 64, R = 100)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL = 1e-4  # BASELINE.json: float distances within 1e-4 relative


def _gauss(n, d, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    return torch.randn((n, d), generator=g, device="cuda")


def _exact_topk(x, q, k, dot, chunk=2048):
    """independent exact ranking: fp32 matmul scores (tf32 disabled), top-k per query"""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        xn = (x * x).sum(1)
        ids, ds = [], []
        for s in range(0, q.shape[0], chunk):
            qq = q[s:s + chunk]
            sc = -(qq @ x.T) if dot else (qq * qq).sum(1)[:, None] + xn[None, :] - 2.0 * (qq @ x.T)
            d, i = torch.topk(sc, k, dim=1, largest=False)
            ids.append(i)
            ds.append(d)
        return torch.cat(ids), torch.cat(ds)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def _check_bf(pkg, x, q, k, measure, dot):
    bf = pkg.BruteForceSearcher(x, measure)
    ids, dists, counts = bf.search_batched(q, k)
    ids2, dists2, _ = bf.search_batched(q, k)
    torch.cuda.synchronize()
    assert bf.path_stats()[0] >= 1 and bf.path_stats()[1] == 0          # answered by the tcgen05 path
    assert (counts == k).all() and (ids == ids2).all() and (dists == dists2).all()   # idempotent
    assert (dists[:, 1:] >= dists[:, :-1]).all()                                     # sorted
    # returned distances = an independent recomputation of the same pairs
    rows = x[ids.long()]
    ref = -(rows * q[:, None, :]).sum(2) if dot else ((rows - q[:, None, :]) ** 2).sum(2)
    scale = q.norm(dim=1)[:, None] * rows.norm(dim=2) + 1.0
    assert ((dists - ref).abs() / scale).max() < REL
    # agreement with the independent exact ranking: same k-th distance, same ids away from near-ties
    gi, gd = _exact_topk(x, q, k, dot)
    assert ((dists - gd).abs() / scale).max() < REL
    rec = np.mean([(ids[i].unsqueeze(1) == gi[i].unsqueeze(0)).any(1).float().mean().item() for i in range(0, q.shape[0], 37)])
    assert rec >= 0.999, rec


def test_c1_bruteforce_sql2_full_size(gpu_lib):
    x, q = _gauss(10_000, 128, 42), _gauss(1_000, 128, 123)
    _check_bf(gpu_lib, x, q, 10, gpu_lib.DistanceMeasure.SquaredL2, False)


def test_c2_bruteforce_dot_full_size(gpu_lib):
    x, q = _gauss(1_000_000, 128, 42), _gauss(10_000, 128, 123)
    _check_bf(gpu_lib, x, q, 10, gpu_lib.DistanceMeasure.DotProduct, True)


def test_c2_scalar_quantized_int8_full_size(gpu_lib):
    x, q = _gauss(1_000_000, 128, 42), _gauss(10_000, 128, 123)
    codes, cal = gpu_lib.scalar_quantize(x)
    scale = float(cal[2])
    s = gpu_lib.ScalarQuantizedBruteForceSearcher.from_quantized(codes, scale, gpu_lib.DistanceMeasure.DotProduct)
    ids, dists, counts = s.search_batched(q, 10)
    torch.cuda.synchronize()
    assert s.path_stats()[0] >= 1 and s.path_stats()[1] == 0
    assert (counts == 10).all() and (dists[:, 1:] >= dists[:, :-1]).all()
    xs = codes.float() * scale                      # the values the reference's kernel sees: (i8)x * scale
    rows = xs[ids.long()]
    ref = -(rows * q[:, None, :]).sum(2)
    sc = q.norm(dim=1)[:, None] * rows.norm(dim=2) + 1.0
    assert ((dists - ref).abs() / sc).max() < REL
    gi, gd = _exact_topk(xs, q, 10, True)
    assert ((dists - gd).abs() / sc).max() < REL
    # levels 128..255 wrap to negative i8 (the reference's quirk): the quantiser's range is exercised on both sides
    assert int(codes.min()) < 0 < int(codes.max())


def test_c3_tree_ah_full_size(gpu_lib):
    import importlib
    import sys
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    bench = importlib.import_module("bench")
    n, dim, K, S, L, R, k, nq = 10_000_000, 96, 2000, 48, 64, 100, 10, 10_000
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(42)
    lat = torch.randn((8192, dim), generator=g, device=dev)
    x = bench.make_points(torch, n, dim, lat, 0.5, 1.0, 42, dev)
    q = bench.make_points(torch, nq, dim, lat, 0.5, 1.0, 123, dev)
    idx = gpu_lib.indexing.build_treeah_index(x, K, S, device=0)
    cfg = gpu_lib.TreeXHybridConfig(num_partitions=K, partitions_to_search=L, use_residuals=True,
                                    pre_reorder_multiplier=R / k, distance_measure=gpu_lib.DistanceMeasure.DotProduct)
    s = gpu_lib.TreeXHybridSearcher(cfg, 0).build_from_index(idx["centers"], idx["codebook"], idx["packed"], idx["ids"],
                                                            idx["part_offsets"], x)
    ids, dists, counts, (ci, cd, cc) = s.search_batched(q, k, pre_reorder_k=R, want_candidates=True)
    ids2, dists2, _ = s.search_batched(q, k, pre_reorder_k=R)
    torch.cuda.synchronize()
    assert (ids == ids2).all() and (dists == dists2).all()                      # idempotent / schedule-independent
    assert (counts == k).all() and (cc == R).all()
    assert (dists[:, 1:] >= dists[:, :-1]).all() and (cd[:, 1:] >= cd[:, :-1]).all()   # both lists sorted
    # exact reorder distances = -q.x of the returned rows (Dot), recomputed independently
    ref = -(x[ids.long()] * q[:, None, :]).sum(2)
    assert ((dists - ref).abs() / (1.0 + ref.abs())).max() < REL
    # the final top-k is the exact top-k OF THE CANDIDATES (reorder_results, tree_x_hybrid/mod.rs:342-364)
    cref = -(x[ci.long()] * q[:, None, :]).sum(2)
    best = torch.topk(cref, k, dim=1, largest=False).values
    assert ((dists - best).abs() / (1.0 + best.abs())).max() < REL
    # every candidate comes from one of the query's L closest partitions
    tok = s.partition_tokens(q[:256], L)
    owner = idx["assign"][ci[:256].long()].to(torch.int32)
    assert (owner.unsqueeze(2) == tok.unsqueeze(1)).any(2).all()
    # recall@10 against exact ground truth (BASELINE: >= 0.95 at L = 64, R = 100)
    gi, _ = _exact_topk(x, q[:1000], k, True, chunk=250)
    rec = np.mean([(ids[i].unsqueeze(1) == gi[i].unsqueeze(0)).any(1).float().mean().item() for i in range(1000)])
    assert rec >= 0.95, rec
