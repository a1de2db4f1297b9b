"""Index build inside the library (csrc/build_index.cu): scann_kmeans_fit, scann_pq_train,
scann_treeah_build, scann_ivf_build.  The trainer is ours (the reference's RNG stream is unpinned), so it is held to
quality against a numpy Lloyd and to determinism; everything with reference semantics (assignment, residual encode,
packing) is compared with the oracle bit for bit on the arrays the build leaves behind."""
import ctypes as C

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def _inertia(x, centers):
    d = (x * x).sum(1)[:, None] - 2.0 * x @ centers.T + (centers * centers).sum(1)[None, :]
    return float(d.min(1).sum()), d.argmin(1)


def _kmeans(pkg, x, K, iters=10, seed=7, balance=0.0):
    centers = np.empty((min(K, len(x)), x.shape[1]), np.float32)
    pkg.capi.check(pkg.load().scann_kmeans_fit(pkg.capi.np_ptr(x), len(x), x.shape[1], x.shape[1], K, iters, seed,
                                               float(balance), pkg.capi.np_ptr(centers), 0, pkg.capi.HOST))
    return centers


def test_kmeans_fit_quality_and_determinism(gpu_lib):
    pkg = gpu_lib
    x, _ = helpers.clustered(20_000, 32, 40, 0.3, 3)
    c1 = _kmeans(pkg, x, 40, 12)
    c2 = _kmeans(pkg, x, 40, 12)
    assert (c1.view(np.uint32) == c2.view(np.uint32)).all(), "k-means must be deterministic for a seed"
    ref = helpers.np_kmeans(x, 40, 12, 7)
    ours, _ = _inertia(x, c1)
    theirs, _ = _inertia(x, ref)
    assert ours <= 1.25 * theirs, (ours, theirs)
    assert not (_kmeans(pkg, x, 40, 12, seed=8).view(np.uint32) == c1.view(np.uint32)).all()


def test_kmeans_fit_device_pointers_and_stride(gpu_lib):
    import torch

    pkg = gpu_lib
    x, _ = helpers.clustered(5_000, 16, 10, 0.3, 4)
    wide = np.zeros((5_000, 24), np.float32)
    wide[:, :16] = x
    host = np.empty((10, 16), np.float32)
    pkg.capi.check(pkg.load().scann_kmeans_fit(pkg.capi.np_ptr(wide), 5_000, 16, 24, 10, 8, 7, 0.0, pkg.capi.np_ptr(host), 0,
                                               pkg.capi.HOST))
    tx = torch.as_tensor(x).cuda()
    dc = pkg.indexing.kmeans(tx, 10, 8, 7)
    assert (dc.cpu().numpy().view(np.uint32) == host.view(np.uint32)).all()


def test_kmeans_fit_edges(gpu_lib):
    pkg = gpu_lib
    x = helpers.gaussian(50, 8, 1)
    one = _kmeans(pkg, x, 1, 3)
    assert np.allclose(one[0], x.astype(np.float64).mean(0), atol=1e-6)
    every = _kmeans(pkg, x, 50, 3)  # K == n: every row its own centre
    _, a = _inertia(x, every)
    assert len(set(a.tolist())) == 50
    L = pkg.load()
    c = np.empty((60, 8), np.float32)
    assert L.scann_kmeans_fit(pkg.capi.np_ptr(x), 50, 8, 8, 60, 3, 7, 0.0, pkg.capi.np_ptr(c), 0, 0) == pkg.capi.INVALID_ARGUMENT
    assert L.scann_kmeans_fit(pkg.capi.np_ptr(x), 0, 8, 8, 1, 3, 7, 0.0, pkg.capi.np_ptr(c), 0, 0) == pkg.capi.INVALID_ARGUMENT
    assert b"empty" in L.scann_last_error()


def test_kmeans_balance_splits_heavy_clusters(gpu_lib):
    pkg = gpu_lib
    rng = np.random.default_rng(5)
    # 64 tight modes of equal weight, K = 64: plain Lloyd from random rows leaves some centres covering several modes;
    # balanced training (56 Lloyd centres + 8 bisections of the heaviest cluster) must not be worse
    lat = rng.standard_normal((64, 24)).astype(np.float32) * 4
    x = (lat[rng.integers(0, 64, 64_000)] + 0.2 * rng.standard_normal((64_000, 24))).astype(np.float32)
    plain = _kmeans(pkg, x, 64, 12, balance=0.0)
    bal = _kmeans(pkg, x, 64, 12, balance=2.0)
    _, a0 = _inertia(x, plain)
    _, a1 = _inertia(x, bal)
    m0, m1 = np.bincount(a0, minlength=64).max(), np.bincount(a1, minlength=64).max()
    assert m1 <= m0 and m1 <= 3.2 * 1000, (m0, m1)
    assert _inertia(x, bal)[0] <= _inertia(x, plain)[0] * 1.25


def test_pq_train_matches_numpy_quality_and_rejects_bad_shapes(gpu_lib):
    pkg = gpu_lib
    x = helpers.gaussian(8_000, 16, 2)
    cb = np.empty((4, 16, 4), np.float32)
    L = pkg.load()
    pkg.capi.check(L.scann_pq_train(pkg.capi.np_ptr(x), 8_000, 16, 16, None, None, 0, 4, 10, 42, pkg.capi.np_ptr(cb), 0, 0))
    for s in range(4):
        sub = np.ascontiguousarray(x[:, s * 4:(s + 1) * 4])
        ours, _ = _inertia(sub, cb[s])
        theirs, _ = _inertia(sub, helpers.np_kmeans(sub, 16, 10, 42 + s))
        assert ours <= 1.15 * theirs
    assert L.scann_pq_train(pkg.capi.np_ptr(x), 8_000, 16, 16, None, None, 0, 5, 10, 42, pkg.capi.np_ptr(cb), 0, 0) == \
        pkg.capi.INVALID_ARGUMENT
    assert b"divisible" in L.scann_last_error()
    # residual form: x - centers[assign]
    centers = _kmeans(pkg, x, 8, 5)
    _, assign = _inertia(x, centers)
    assign = assign.astype(np.uint32)
    cb2 = np.empty((4, 16, 4), np.float32)
    pkg.capi.check(L.scann_pq_train(pkg.capi.np_ptr(x), 8_000, 16, 16, pkg.capi.np_ptr(centers), pkg.capi.np_ptr(assign), 8, 4,
                                    10, 42, pkg.capi.np_ptr(cb2), 0, 0))
    resid = x - centers[assign]
    direct = np.empty((4, 16, 4), np.float32)
    pkg.capi.check(L.scann_pq_train(pkg.capi.np_ptr(resid), 8_000, 16, 16, None, None, 0, 4, 10, 42, pkg.capi.np_ptr(direct), 0, 0))
    assert (cb2.view(np.uint32) == direct.view(np.uint32)).all()


def test_treeah_build_searches_like_the_oracle_on_its_own_arrays(gpu_lib, oracle):
    """scann_treeah_build → search; then the same index rebuilt OUTSIDE the library with the oracle's assignment /
    encode / packing from the centres and codebook the trainer produced must give the same results."""
    pkg = gpu_lib
    x, _ = helpers.clustered(30_000, 32, 48, 0.35, 11)
    q = (x[:64] + 0.01).astype(np.float32)
    cfg = pkg.TreeXHybridConfig(num_partitions=24, partitions_to_search=6, pre_reorder_multiplier=8.0)
    cfg.hash_config.num_subspaces = 16
    s = pkg.TreeXHybridSearcher(cfg).build(x, train_rows=0, kmeans_iters=8, seed=7)
    ids, dists, counts = s.search_batched(q, 10)[:3]
    assert (counts == 10).all() and (np.diff(dists, axis=1) >= 0).all()
    bf = pkg.BruteForceSearcher(x)
    ei, ed, _ = bf.search_batched(q, 10)
    assert helpers.recall(ids, ei, 10) >= 0.9
    # exact distances of what it returned
    d64 = ((q[:, None, :].astype(np.float64) - x[ids].astype(np.float64)) ** 2).sum(-1)
    assert np.abs(d64 - dists).max() <= 1e-5 * (1 + d64.max())
    # the trainer's centres / codebook, the oracle's index arrays
    centers = _kmeans(pkg, x, 24, 8, 7, balance=3.0)
    assign = oracle.partition(centers, x, 1)[0][:, 0].astype(np.uint32)
    cb = np.empty((16, 16, 2), np.float32)
    pkg.capi.check(pkg.load().scann_pq_train(pkg.capi.np_ptr(x), len(x), 32, 32, pkg.capi.np_ptr(centers), pkg.capi.np_ptr(assign),
                                             24, 16, 8, 42, pkg.capi.np_ptr(cb), 0, 0))
    order = np.argsort(assign, kind="stable").astype(np.uint32)
    off = np.concatenate([[0], np.cumsum(np.bincount(assign, minlength=24))]).astype(np.uint64)
    packed = oracle.pack4(oracle.pq_encode_residual(cb, x[order], centers, assign[order]))
    s2 = pkg.TreeXHybridSearcher(cfg).build_from_index(centers, cb, packed, order, off, x)
    ids2, dists2, counts2 = s2.search_batched(q, 10)[:3]
    assert (counts2 == counts).all()
    assert (dists2.view(np.uint32) == dists.view(np.uint32)).all(), "library build != oracle-assembled index"
    assert (ids2 == ids).mean() > 0.999


def test_treeah_build_flat_hasher_and_set_reorder(gpu_lib):
    pkg = gpu_lib
    x = helpers.gaussian(4_000, 16, 9)
    q = x[:8] + 0.001
    h = C.c_void_p()
    L = pkg.load()
    pkg.capi.check(L.scann_treeah_build(pkg.capi.np_ptr(x), 4_000, 16, 16, 1, 8, 0, 10, 42, 0, pkg.capi.SQL2, 1, 0, 0, C.byref(h)))
    ids = np.empty((8, 5), np.uint32)
    d = np.empty((8, 5), np.float32)
    cnt = np.empty(8, np.uint32)
    qq = np.ascontiguousarray(q, np.float32)

    def run():
        pkg.capi.check(L.scann_treeah_search(h, pkg.capi.np_ptr(qq), 8, 16, 1, 200, 5, pkg.capi.np_ptr(ids), pkg.capi.np_ptr(d),
                                             pkg.capi.np_ptr(cnt), None, None, None, 0, None))
        return ids.copy(), d.copy()

    i1, d1 = run()                       # reorder on: exact SqL2, the query's own row first
    assert (i1[:, 0] == np.arange(8)).all() and d1[:, 0].max() < 1e-4
    pkg.capi.check(L.scann_treeah_set_reorder(h, 0))
    i0, d0 = run()                       # AsymmetricHasher::search: approximate distances
    assert not np.allclose(d0, d1)
    pkg.capi.check(L.scann_treeah_set_reorder(h, 1))
    i2, d2 = run()
    assert (i2 == i1).all() and (d2.view(np.uint32) == d1.view(np.uint32)).all()
    L.scann_treeah_destroy(h)
    bad = C.c_void_p()
    assert L.scann_treeah_build(pkg.capi.np_ptr(x), 4_000, 16, 16, 4, 5, 0, 10, 42, 1, 0, 1, 0, 0, C.byref(bad)) == \
        pkg.capi.INVALID_ARGUMENT and not bad.value
    assert L.scann_treeah_build(pkg.capi.np_ptr(x), 0, 16, 16, 4, 8, 0, 10, 42, 1, 0, 1, 0, 0, C.byref(bad)) == \
        pkg.capi.INVALID_ARGUMENT
    assert b"empty dataset" in L.scann_last_error()


def test_ivf_build_is_exact_inside_probed_leaves(gpu_lib):
    pkg = gpu_lib
    x, _ = helpers.clustered(12_000, 24, 30, 0.3, 13)
    q = (x[100:132] + 0.01).astype(np.float32)
    h = C.c_void_p()
    L = pkg.load()
    pkg.capi.check(L.scann_ivf_build(pkg.capi.np_ptr(x), 12_000, 24, 24, 20, 8, 7, 0, 0, C.byref(h)))
    ids = np.empty((32, 10), np.uint32)
    d = np.empty((32, 10), np.float32)
    cnt = np.empty(32, np.uint32)
    pkg.capi.check(L.scann_ivf_search(h, 0, pkg.capi.np_ptr(q), 32, 24, 20, 10, pkg.capi.SQL2, -1, pkg.capi.np_ptr(ids),
                                      pkg.capi.np_ptr(d), pkg.capi.np_ptr(cnt), 0, None))
    L.scann_ivf_destroy(h)
    ei, ed, _ = pkg.BruteForceSearcher(x).search_batched(q, 10)  # probing every leaf = brute force
    assert (d.view(np.uint32) == ed.view(np.uint32)).all() and (ids == ei).mean() > 0.999
