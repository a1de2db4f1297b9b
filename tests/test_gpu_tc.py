"""GPU tests of the tcgen05/TMEM/TMA ranking contraction (csrc/tc_gemm.cu) through its C-ABI tap.
The kernel's scores only rank candidates (every returned distance is re-scored exactly), so the check here is
against a float64 evaluation of the same bf16-rounded operands, tolerance = f32 accumulation error."""
import numpy as np
import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu


def bf16(a):
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).bfloat16().double().numpy()


def expected(q, rows, scale, want_norm):
    i8 = rows.dtype == np.int8
    qs = bf16(q * np.float32(scale)) if i8 else bf16(q)
    xs = rows.astype(np.float64) if i8 else bf16(rows)
    vals = rows.astype(np.float32) * np.float32(scale) if i8 else rows
    hx = 0.5 * (vals.astype(np.float64) ** 2).sum(1) if want_norm else np.zeros(rows.shape[0])
    return hx[None, :] - qs @ xs.T


@pytest.mark.parametrize("nq,n,dim,want_norm", [(1, 1, 1, True), (5, 100, 96, True), (130, 1000, 128, False),
                                                (257, 4099, 64, True), (300, 3000, 200, True),
                                                (77, 20_000, 256, False), (1000, 10_000, 128, True),
                                                (33, 70_001, 33, True)])
def test_tc_dense_scores(gpu_lib, nq, n, dim, want_norm):
    q = helpers.gaussian(nq, dim, 1)
    x = helpers.gaussian(n, dim, 2)
    got = gpu_lib.tc_scores(q, x, want_norm=want_norm)
    want = expected(q, x, 1.0, want_norm)
    scale = np.linalg.norm(q, axis=1)[:, None] * np.linalg.norm(x, axis=1)[None, :] + 1.0
    err = np.abs(got - want) / scale
    assert err.max() < 2e-5, f"max scaled error {err.max()}"


def test_tc_dense_scores_int8_rows(gpu_lib):
    rng = np.random.default_rng(3)
    q = helpers.gaussian(200, 128, 4)
    x = rng.integers(-128, 128, (5000, 128), dtype=np.int8)
    scale = 0.0234
    got = gpu_lib.tc_scores(q, x, scale=scale, want_norm=True)
    want = expected(q, x, scale, True)
    norm = np.linalg.norm(q, axis=1)[:, None] * np.linalg.norm(x.astype(np.float32) * scale, axis=1)[None, :] + 1.0
    assert (np.abs(got - want) / norm).max() < 2e-5


def test_tc_filter_lists(gpu_lib):
    nq, n, dim, cap = 300, 50_000, 128, 512
    q = helpers.gaussian(nq, dim, 5)
    x = helpers.gaussian(n, dim, 6)
    dense = gpu_lib.tc_scores(q, x, want_norm=True)
    thr = np.sort(dense, axis=1)[:, 100].copy()  # ~101 survivors per query
    thr[7] = -np.inf                             # nothing passes
    thr[11] = np.inf                             # everything passes (padding rows too): counted, not stored
    cand, cnt = gpu_lib.tc_scores(q, x, want_norm=True, thr=thr, cap=cap)
    assert cnt[7] == 0 and cnt[11] == (n + 127) // 128 * 128
    for i in range(nq):
        if i == 11:
            continue
        want = np.nonzero(dense[i] <= thr[i])[0]
        rows = np.sort((cand[i, :cnt[i]] & np.uint64(0xFFFFFFFF)).astype(np.int64))
        assert cnt[i] == len(want) and (rows == want).all(), f"query {i}"
    rows11 = (cand[11] & np.uint64(0xFFFFFFFF)).astype(np.int64)
    assert len(set(rows11.tolist())) == cap and rows11.max() < (n + 127) // 128 * 128


# ----------------------------------------------------------------------------- searchers on the tensor-core path
@pytest.mark.parametrize("measure", ["SquaredL2", "DotProduct"])
@pytest.mark.parametrize("n,dim,nq,k", [(100_000, 128, 500, 10), (40_000, 96, 300, 100), (20_000, 200, 64, 10)])
def test_bf_tensor_core_path_is_exact(gpu_lib, oracle, measure, n, dim, nq, k):
    db = helpers.gaussian(n, dim, 42)
    q = helpers.gaussian(nq, dim, 123)
    m = gpu_lib.DistanceMeasure[measure]
    om = {"SquaredL2": oracle.SQL2, "DotProduct": oracle.DOT}[measure]
    bf = gpu_lib.BruteForceSearcher(db, m)
    ids, dists, counts = bf.search_batched(q, k)
    tc_chunks, legacy_chunks = bf.path_stats()
    assert tc_chunks >= 1 and legacy_chunks == 0, "the tcgen05 path did not answer this batch"
    rc, oids, odists, ocounts = oracle.bf_search(db, q, k, om, nthreads=8)
    assert (counts == ocounts).all()
    # certified threshold + exact re-score: the distance lists are bit-identical, ids differ only inside exact ties
    assert (dists.view(np.uint32) == odists.view(np.uint32)).all()
    compared, mism = helpers.ids_equal_away_from_ties(ids, dists, oids, odists, counts, rel_gap=0.0)
    assert mism == 0 and compared > 0


def test_bf_clustered_rows_sorted_by_cluster(gpu_lib, oracle):
    # rows ordered cluster by cluster: the strided sample must still give a usable threshold for every query
    x, _ = helpers.clustered(60_000, 64, 30, 0.2, 9, normalize=False)
    order = np.argsort(x[:, 0], kind="stable")
    db = np.ascontiguousarray(x[order])
    q = db[::997][:50] + np.float32(0.01)
    bf = gpu_lib.BruteForceSearcher(db)
    ids, dists, counts = bf.search_batched(q, 10)
    rc, oids, odists, oc = oracle.bf_search(db, q, 10, oracle.SQL2, nthreads=8)
    assert (dists.view(np.uint32) == odists.view(np.uint32)).all()
    assert bf.path_stats()[0] >= 1


@pytest.mark.parametrize("measure", ["SquaredL2", "DotProduct"])
@pytest.mark.parametrize("order,k", [("best_last", 10), ("best_first", 10), ("best_last", 100), ("dups_across", 10),
                                     ("far_prefix", 10)])
def test_bf_two_pass_filter_row_orders(gpu_lib, oracle, measure, order, k):
    # From 32768 rows on the FILTER runs in two passes: the first eighth of the rows under the sample's bound, the rest
    # under the bound derived from the prefix survivors (brute_force.cu, bf_tighten_kernel).  Row orders that make the
    # prefix unrepresentative must not cost exactness: all good rows at the end (the prefix bound is loose), all at the
    # start (the tightened bound is as tight as it gets), exact duplicates of the best row on both sides of the cut,
    # and a prefix so far from the queries that its lists may hold fewer than k rows.
    n, dim, nq = 60_000, 64, 96
    x = helpers.gaussian(n, dim, 21)
    u = helpers.gaussian(1, dim, 22)[0]
    u /= np.linalg.norm(u)
    q = (np.float32(3.0) * u[None, :] + np.float32(0.3) * helpers.gaussian(nq, dim, 23)).astype(np.float32)
    proj = x @ u
    if order == "best_last":
        db = x[np.argsort(proj, kind="stable")]
    elif order == "best_first":
        db = x[np.argsort(-proj, kind="stable")]
    elif order == "dups_across":
        db = x.copy()
        db[:50] = np.float32(3.0) * u
        db[-50:] = np.float32(3.0) * u
    else:  # far_prefix: the first eighth is a tight far-away cluster
        db = x.copy()
        db[: n // 8] = np.float32(-40.0) * u[None, :] + np.float32(0.01) * helpers.gaussian(n // 8, dim, 24)
    db = np.ascontiguousarray(db, np.float32)
    m = gpu_lib.DistanceMeasure[measure]
    om = {"SquaredL2": oracle.SQL2, "DotProduct": oracle.DOT}[measure]
    bf = gpu_lib.BruteForceSearcher(db, m)
    ids, dists, counts = bf.search_batched(q, k)
    assert bf.path_stats() == (1, 0), "the tcgen05 path did not answer this batch"
    rc, oids, odists, ocounts = oracle.bf_search(db, q, k, om, nthreads=8)
    assert (counts == ocounts).all()
    assert (dists.view(np.uint32) == odists.view(np.uint32)).all()
    compared, mism = helpers.ids_equal_away_from_ties(ids, dists, oids, odists, counts, rel_gap=0.0)
    assert mism == 0


def test_sq8_two_pass_filter_best_rows_last(gpu_lib, oracle):
    n, dim, nq, k = 50_000, 128, 64, 10
    x = helpers.gaussian(n, dim, 31)
    u = helpers.gaussian(1, dim, 32)[0]
    u /= np.linalg.norm(u)
    q = (np.float32(3.0) * u[None, :] + np.float32(0.3) * helpers.gaussian(nq, dim, 33)).astype(np.float32)
    db = np.ascontiguousarray(x[np.argsort(x @ u, kind="stable")], np.float32)
    ocodes, ocal = oracle.sq8_quantize(db)
    s = gpu_lib.ScalarQuantizedBruteForceSearcher.from_quantized(ocodes, float(ocal[2]), gpu_lib.DistanceMeasure.DotProduct)
    ids, dists, counts = s.search_batched(q, k)
    assert s.path_stats() == (1, 0)
    rc, oids, odists, ocounts = oracle.sq8_search(ocodes, float(ocal[2]), q, k, oracle.DOT, nthreads=8)
    assert (dists.view(np.uint32) == odists.view(np.uint32)).all()
    compared, mism = helpers.ids_equal_away_from_ties(ids, dists, oids, odists, counts, rel_gap=0.0)
    assert mism == 0


def test_bf_list_overflow_falls_back(gpu_lib, oracle):
    # 9000 identical rows: every row passes any threshold -> the lists (cap 4096) overflow -> CUDA-core path
    db = np.tile(helpers.gaussian(1, 32, 1), (9000, 1)).astype(np.float32)
    db[:10] += helpers.gaussian(10, 32, 2)
    q = helpers.gaussian(8, 32, 3)
    bf = gpu_lib.BruteForceSearcher(db)
    ids, dists, counts = bf.search_batched(q, 5)
    assert bf.path_stats() == (0, 1)
    rc, oids, odists, oc = oracle.bf_search(db, q, 5, oracle.SQL2, nthreads=4)
    assert (dists.view(np.uint32) == odists.view(np.uint32)).all()


def test_bf_cuda_core_path_still_matches(gpu_lib, oracle, monkeypatch):
    monkeypatch.setenv("SCANN_BF_NO_TC", "1")
    db = helpers.gaussian(20_000, 128, 42)
    q = helpers.gaussian(100, 128, 123)
    bf = gpu_lib.BruteForceSearcher(db, gpu_lib.DistanceMeasure.DotProduct)
    ids, dists, counts = bf.search_batched(q, 10)
    assert bf.path_stats()[0] == 0
    rc, oids, odists, oc = oracle.bf_search(db, q, 10, oracle.DOT, nthreads=8)
    assert (ids == oids).mean() > 0.999


@pytest.mark.parametrize("measure", ["SquaredL2", "DotProduct"])
def test_sq8_tensor_core_path_is_exact(gpu_lib, oracle, measure):
    db = helpers.gaussian(100_000, 128, 42)
    q = helpers.gaussian(400, 128, 123)
    ocodes, ocal = oracle.sq8_quantize(db)
    m = gpu_lib.DistanceMeasure[measure]
    om = {"SquaredL2": oracle.SQL2, "DotProduct": oracle.DOT}[measure]
    s = gpu_lib.ScalarQuantizedBruteForceSearcher.from_quantized(ocodes, float(ocal[2]), m)
    ids, dists, counts = s.search_batched(q, 10)
    assert s.path_stats()[0] >= 1 and s.path_stats()[1] == 0
    rc, oids, odists, ocounts = oracle.sq8_search(ocodes, float(ocal[2]), q, 10, om, nthreads=8)
    assert (dists.view(np.uint32) == odists.view(np.uint32)).all()
    compared, mism = helpers.ids_equal_away_from_ties(ids, dists, oids, odists, counts, rel_gap=0.0)
    assert mism == 0


# ----------------------------------------------------------------------------- partitioner on the tensor-core path
@pytest.mark.parametrize("K,dim,nq,L", [(2000, 96, 1000, 64), (8192, 128, 300, 128), (300, 64, 100, 300),
                                        (5000, 200, 50, 1000), (1000, 7, 64, 10), (256, 96, 33, 500)])
def test_partition_tensor_core_bit_exact(gpu_lib, oracle, K, dim, nq, L):
    x, _ = helpers.clustered(K + nq, dim, 40, 0.4, K + dim)
    centers, q = np.ascontiguousarray(x[:K]), np.ascontiguousarray(x[K:])
    part = gpu_lib.TreePartitioner(centers)
    tokens, dists = part.partition(q, L)
    otok, odist = oracle.partition(centers, q, L)
    assert (tokens == otok).all()
    assert (dists.view(np.uint32) == odist.view(np.uint32)).all()


def test_partition_tensor_core_tied_centres_fallback(gpu_lib, oracle):
    # 700 identical centres tie at every threshold: more survivors than the shared-memory list holds
    rng = np.random.default_rng(5)
    centers = np.tile(rng.normal(0, 1, (1, 32)).astype(np.float32), (700, 1))
    centers[::7] += rng.normal(0, 1, (100, 32)).astype(np.float32)
    q = rng.normal(0, 1, (20, 32)).astype(np.float32)
    part = gpu_lib.TreePartitioner(centers)
    tokens, dists = part.partition(q, 50)
    otok, odist = oracle.partition(centers, q, 50)
    assert (tokens == otok).all() and (dists.view(np.uint32) == odist.view(np.uint32)).all()


def test_partition_cuda_core_path_still_bit_exact(gpu_lib, oracle, monkeypatch):
    monkeypatch.setenv("SCANN_PART_NO_TC", "1")
    rng = np.random.default_rng(6)
    centers = rng.normal(0, 1, (1500, 96)).astype(np.float32)
    q = rng.normal(0, 1, (100, 96)).astype(np.float32)
    tokens, dists = gpu_lib.TreePartitioner(centers).partition(q, 64)
    otok, odist = oracle.partition(centers, q, 64)
    assert (tokens == otok).all() and (dists.view(np.uint32) == odist.view(np.uint32)).all()


# ----------------------------------------------------------------------------- radius search
@pytest.mark.parametrize("measure,radius", [("SquaredL2", 110.0), ("L2", 10.5), ("DotProduct", -25.0)])
def test_bf_search_radius_matches_oracle(gpu_lib, oracle, measure, radius):
    # BruteForceSearcher::search_radius (searcher.rs:142-167): every row with distance <= radius, ascending
    db = helpers.gaussian(30_000, 96, 42)
    q = helpers.gaussian(40, 96, 123)
    om = {"SquaredL2": oracle.SQL2, "DotProduct": oracle.DOT, "L2": oracle.L2}[measure]
    bf = gpu_lib.BruteForceSearcher(db, gpu_lib.DistanceMeasure[measure])
    ids, dists, counts = bf.search_radius_batched(q, radius, max_results=2048)
    total = 0
    for i in range(len(q)):
        oi, od = oracle.bf_search_radius(db, q[i], radius, om)
        assert counts[i] == len(oi), (i, counts[i], len(oi))
        assert (dists[i, :len(oi)].view(np.uint32) == od.view(np.uint32)).all()
        assert (ids[i, :len(oi)] == oi).all() or len(set(od.tolist())) < len(od)
        assert (ids[i, len(oi):] == 0xFFFFFFFF).all()
        total += len(oi)
    assert total > 50  # the radii above select a few rows per query on this data


def test_bf_search_radius_reference_kat_and_truncation(gpu_lib):
    # src/brute_force/searcher.rs:327-340 test_brute_force_radius_search: 4 of the 5 points lie within 1.5 of the origin
    db = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1]], np.float32)
    bf = gpu_lib.BruteForceSearcher(db)
    r = bf.search_radius([0, 0, 0], 1.5)
    assert len(r) == 4 and all(d <= 1.5 for _, d in r) and r[0] == (0, 0.0)
    with pytest.raises(gpu_lib.ScannError) as e:
        bf.search_radius([0, 0, 0], 1.5, max_results=2)      # 4 rows qualify, only room for 2
    assert e.value.code == gpu_lib.capi.RESOURCE_EXHAUSTED
    with pytest.warns(UserWarning):  # the truncated (nearest max_results) rows are still delivered
        ti, td, tc = bf.search_radius_batched(np.array([[0, 0, 0]], np.float32), 1.5, max_results=2, allow_truncated=True)
    assert tc[0] == 2 and (np.diff(td[0]) >= 0).all()
    assert bf.search_radius([9, 9, 9], 0.5) == []
