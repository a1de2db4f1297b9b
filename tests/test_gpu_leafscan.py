"""GPU parity of the Scann façade's tree modes that score every member of the probed leaves (scann_ivf_*):
Scann::search_partitioned, Scann::search_tree_ah ("variant B", f32 LookupTable over byte codes) and the flat
AsymmetricHasher f32-LUT scoring, against the oracle's restatements (oracle/scann_oracle.cpp a11/a13)."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def _ivf_index(oracle, n, dim, K, seed):
    x, _ = helpers.clustered(n, dim, 32, 0.35, seed, normalize=False)
    centers = helpers.np_kmeans(x, K, 6, seed)
    assign = oracle.partition(centers, x, 1)[0][:, 0].astype(np.uint32)
    order = np.argsort(assign, kind="stable").astype(np.uint32)
    off = np.concatenate([[0], np.cumsum(np.bincount(assign, minlength=len(centers)))]).astype(np.uint64)
    return x, centers, order, off


@pytest.mark.parametrize("measure", ["SquaredL2", "DotProduct", "L2"])
@pytest.mark.parametrize("n,dim,K,L,k,nq", [(20_000, 64, 40, 6, 10, 64), (3000, 17, 9, 20, 25, 20)])
def test_search_partitioned_matches_oracle(gpu_lib, oracle, measure, n, dim, K, L, k, nq):
    x, centers, order, off = _ivf_index(oracle, n, dim, K, 11)
    q = (x[:nq] + 0.05).astype(np.float32)
    om = {"SquaredL2": oracle.SQL2, "DotProduct": oracle.DOT, "L2": oracle.L2}[measure]
    s = gpu_lib.LeafScanSearcher(centers, order, off, x)
    ids, dists, counts = s.search_partitioned(q, k, L, gpu_lib.DistanceMeasure[measure])
    rc, oids, odists, ocounts = oracle.scann_partitioned(centers, off, order, x, q, L, k, om, nthreads=8)
    assert rc == 0 and (counts == ocounts).all()
    assert (dists.view(np.uint32) == odists.view(np.uint32)).all()   # single-pair AVX2 order restated bit for bit
    assert (ids == oids).all()                                        # stable sort: ties keep candidate order


@pytest.mark.parametrize("C,dim,S", [(16, 32, 8), (256, 32, 8), (16, 96, 48), (64, 30, 6)])
def test_search_tree_ah_variant_b_matches_oracle(gpu_lib, oracle, C, dim, S):
    n, K, L, k = 15_000, 24, 5, 10
    x, centers, order, off = _ivf_index(oracle, n, dim, K, 12)
    ds = dim // S
    rng = np.random.default_rng(C)
    cb = np.stack([x[rng.choice(n, C, replace=False)][:, s * ds:(s + 1) * ds] for s in range(S)]).astype(np.float32)
    codes = oracle.pq_encode(cb, x)                                   # [n, S] bytes, by datapoint id
    q = (x[100:164] + 0.02).astype(np.float32)
    s = gpu_lib.LeafScanSearcher(centers, order, off, x, cb, codes)
    ids, dists, counts = s.search_tree_ah(q, k, L)
    rc, oids, odists, ocounts = oracle.scann_tree_ah(centers, off, order, cb, codes, x, q, L, k, nthreads=8)
    assert rc == 0 and (counts == ocounts).all()
    assert (dists.view(np.uint32) == odists.view(np.uint32)).all()   # sequential f32 LUT sums: bit-identical
    assert (ids == oids).all()
    # with Scann::search_impl's post-hoc exact reorder of the k results (utils/reordering.rs:23-54)
    ids2, dists2, _ = s.search_tree_ah(q, k, L, reorder=gpu_lib.DistanceMeasure.SquaredL2)
    rc, oids2, odists2, _ = oracle.scann_tree_ah(centers, off, order, cb, codes, x, q, L, k,
                                                 reorder_measure=oracle.SQL2, nthreads=8)
    assert (ids2 == oids2).all() and (dists2.view(np.uint32) == odists2.view(np.uint32)).all()


def test_flat_asymmetric_hasher_f32_lut_scoring(gpu_lib, oracle):
    # AsymmetricHasher::search (hasher.rs:162-185) with the 256-code default: K = 1, L = 1.  The reference keeps
    # its k best in a FastTopNeighbors; away from exact ties that is the k smallest, which is what is compared.
    n, dim, S, C, k = 8000, 24, 6, 256, 10
    x = helpers.gaussian(n, dim, 21)
    rng = np.random.default_rng(1)
    cb = np.stack([x[rng.choice(n, C, replace=False)][:, s * 4:(s + 1) * 4] for s in range(S)]).astype(np.float32)
    codes = oracle.pq_encode(cb, x)
    q = helpers.gaussian(40, dim, 22)
    centers = np.zeros((1, dim), np.float32)
    s = gpu_lib.LeafScanSearcher(centers, np.arange(n, dtype=np.uint32), np.array([0, n], np.uint64), None, cb, codes)
    ids, dists, counts = s.search_tree_ah(q, k, 1)
    rc, oids, odists, ocounts = oracle.ah_search(cb, codes, q, k, lut16=False, nthreads=8)
    assert (counts == ocounts).all()
    assert (np.sort(dists, 1).view(np.uint32) == np.sort(odists, 1).view(np.uint32)).all()
    compared, mism = helpers.ids_equal_away_from_ties(ids, dists, oids, odists, counts, rel_gap=0.0)
    assert mism == 0


def test_asymmetric_hasher_search_with_reordering(gpu_lib, oracle):
    # hasher.rs:188-229: top pre_reorder_k by the f32 LUT -> exact SqL2 -> first k
    n, dim, S, C, k, pre = 6000, 24, 6, 64, 10, 80
    x = helpers.gaussian(n, dim, 31)
    rng = np.random.default_rng(2)
    cb = np.stack([x[rng.choice(n, C, replace=False)][:, s * 4:(s + 1) * 4] for s in range(S)]).astype(np.float32)
    codes = oracle.pq_encode(cb, x)
    q = helpers.gaussian(30, dim, 32)
    s = gpu_lib.LeafScanSearcher(np.zeros((1, dim), np.float32), np.arange(n, dtype=np.uint32),
                                 np.array([0, n], np.uint64), x, cb, codes)
    ids, dists, counts = s.search_with_reordering(q, k, pre)
    rc, oids, odists, ocounts = oracle.ah_search(cb, codes, q, k, lut16=False, raw=x, pre_k=pre, nthreads=8)
    assert rc == 0 and (counts == ocounts).all()
    compared, mism = helpers.ids_equal_away_from_ties(ids, dists, oids, odists, counts, rel_gap=0.0)
    same = ids == oids
    assert mism == 0 and same.mean() > 0.97            # candidate-set ties at the pre_reorder cut-off may differ
    assert (dists[same].view(np.uint32) == odists[same].view(np.uint32)).all()


def test_leafscan_errors_and_facade(gpu_lib, oracle):
    x, centers, order, off = _ivf_index(oracle, 2000, 8, 5, 13)
    s = gpu_lib.LeafScanSearcher(centers, order, off, x)
    with pytest.raises(gpu_lib.ScannError) as e:
        s.search_partitioned(np.zeros((1, 3), np.float32), 5, 2)
    assert e.value.code == gpu_lib.capi.INVALID_ARGUMENT
    with pytest.raises(gpu_lib.ScannError) as e:
        s.search_tree_ah(x[:2], 5, 2)                                # no hasher in this index
    assert e.value.code == gpu_lib.capi.FAILED_PRECONDITION
    sc = gpu_lib.ScannBuilder().num_neighbors(5).tree(8, 3).build(x)  # tree only -> SearchMode::Partitioned
    assert sc.search_mode == gpu_lib.SearchMode.Partitioned
    ids, dists, counts = sc.search_batched(x[:32])
    assert ids.shape == (32, 5) and (np.asarray(ids)[:, 0] == np.arange(32)).all() and (np.asarray(dists)[:, 0] == 0).all()


def test_ivf_create_refuses_member_ids_outside_the_dataset(gpu_lib, oracle):
    """index arrays are validated at create — a member id >= the number of datapoints would be read out of
    bounds by the scan kernels (the reference skips such entries through dataset.get)."""
    x, centers, order, off = _ivf_index(oracle, 2000, 16, 8, 5)
    bad = order.copy()
    bad[7] = 2000  # one id past the end
    with pytest.raises(gpu_lib.ScannError) as e:
        gpu_lib.LeafScanSearcher(centers, bad, off, x)
    assert e.value.code == gpu_lib.capi.INVALID_ARGUMENT and "member ids" in e.value.message
    gpu_lib.LeafScanSearcher(centers, order, off, x).close()  # the well-formed index is accepted
