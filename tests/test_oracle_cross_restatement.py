"""The C++ oracle against a second, independent Python restatement of the reference's approximate stage
(tests/ref_restatement.py): candidate ids, approximate distances and their ORDER (FastTopNeighbors slot replacement, stable
sorts, LUT16 rounding) must agree bit for bit, including inside exact ties of the integer scores."""
import numpy as np
import pytest

import helpers
import ref_restatement as rr


@pytest.mark.parametrize("n,dim,K,S,L,R,seed", [(1500, 16, 6, 8, 3, 25, 1), (900, 12, 5, 6, 5, 40, 2), (600, 8, 4, 8, 2, 7, 3)])
def test_oracle_candidates_equal_python_restatement(oracle, n, dim, K, S, L, R, seed):
    x, _ = helpers.clustered(n, dim, 8, 0.4, seed, normalize=False)
    x = np.round(x * 4) / 4  # coarse grid: many exactly tied integer scores, so tie order is really exercised
    x = x.astype(np.float32)
    idx = helpers.build_index(oracle, x, K, S, seed=seed, iters=4)
    q = (x[::max(1, n // 10)][:10] + np.float32(0.125)).astype(np.float32)
    rc, oids, odists, ocounts, ocand, ocd, ocn = oracle.treex_search(
        idx["centers"], idx["codebook"], idx["part_offsets"], idx["ids"], idx["packed"], x, q, L, R, 5, lut16=True,
        want_candidates=True)
    assert rc == 0
    ties = 0
    for i in range(len(q)):
        want = rr.approx_candidates(idx["centers"], idx["codebook"], idx["part_offsets"], idx["ids"], idx["packed"], q[i],
                                    L, R)
        c = int(ocn[i])
        assert c == len(want)
        wd = np.array([d for _, d in want], np.float32)
        wi = np.array([j for j, _ in want], np.uint32)
        assert (ocd[i, :c].view(np.uint32) == wd.view(np.uint32)).all(), "approximate distances differ"
        assert (ocand[i, :c] == wi).all(), "candidate ids / tie order differ"
        ties += int((np.diff(wd) == 0).sum())
    assert ties > 0, "the case was meant to contain exact ties"


def test_fast_top_neighbors_restatement_matches_reference_unit_tests():
    # src/brute_force/top_k.rs tests (test_fast_top_neighbors*): capacity 3, pushes 5,3,7,4 -> {3,4,5}
    t = rr.FastTopNeighbors(3)
    for i, d in enumerate([5.0, 3.0, 7.0, 4.0, 6.0]):
        t.push(i, np.float32(d))
    assert [i for i, _ in t.results()] == [1, 3, 0]
    z = rr.FastTopNeighbors(0)
    z.push(1, np.float32(1.0))
    assert z.results() == []


@pytest.mark.parametrize("k,n,levels,seed", [(1, 50, 3, 0), (5, 200, 4, 1), (10, 500, 6, 2), (32, 400, 5, 3), (100, 300, 7, 4)])
def test_oracle_topk_heap_order_equals_python_binary_heap(oracle, k, n, levels, seed):
    """TopK (BinaryHeap of (OrderedFloat, index)): accepted flags, kept set and the ORDER results() returns — which inside
    ties is the heap's internal array order — against a Python emulation of std's BinaryHeap sift algorithms."""
    rng = np.random.default_rng(seed)
    dists = rng.integers(0, levels, n).astype(np.float32)  # few distinct values: ties everywhere
    ids = rng.permutation(n).astype(np.uint32)
    oi, od, acc = oracle.topk_run(k, ids, dists)
    t = rr.TopK(k)
    pacc = [t.push(int(i), d) for i, d in zip(ids, dists)]
    res = t.results()
    assert pacc == acc.tolist()
    assert [i for i, _ in res] == oi.tolist()
    assert [np.float32(d) for _, d in res] == od.tolist()


def test_python_binary_heap_is_a_heap():
    rng = np.random.default_rng(9)
    h = rr.RustBinaryHeap()
    vals = [(float(v), int(i)) for i, v in enumerate(rng.integers(0, 20, 300))]
    for v in vals:
        h.push(v)
    out = [h.pop() for _ in range(len(vals))]
    assert out == sorted(vals, reverse=True)


@pytest.mark.parametrize("seed,scale", [(0, 1.0), (1, 37.5), (2, 1e-3)])
def test_oracle_sq8_quantiser_equals_python_restatement(oracle, seed, scale):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((300, 7)) * scale).astype(np.float32)
    x[5, 3] = np.float32(40.0 * scale)  # an outlier beyond 3 sigma: clipped
    codes, cal = oracle.sq8_quantize(x)
    pc, (lo, hi, sc, inv) = rr.sq8_quantize(x)
    assert (np.array([lo, hi, sc, inv], np.float32).view(np.uint32) == cal.view(np.uint32)).all()
    assert (codes == pc).all()
    assert codes.min() < 0  # values above 127 wrap to negative i8, as in the reference


# ----------------------------------------------------------------------------- the AVX2 + FMA distance kernels
def test_exact_fma_helper_rounds_like_ieee():
    # against float64 evaluation where that is exact enough to decide, and hand-made halfway cases
    rng = np.random.default_rng(5)
    for _ in range(300):
        a, b, c = (np.float32(v) for v in rng.standard_normal(3) * 10.0 ** rng.integers(-3, 4))
        exact = rr.Fraction(float(a)) * rr.Fraction(float(b)) + rr.Fraction(float(c))
        got = rr.fma32(a, b, c)
        lo, hi = np.nextafter(got, np.float32(-np.inf)), np.nextafter(got, np.float32(np.inf))
        err = abs(rr.Fraction(float(got)) - exact)
        assert err <= abs(rr.Fraction(float(lo)) - exact) and err <= abs(rr.Fraction(float(hi)) - exact)
    one, ulp = np.float32(1.0), np.float32(2.0 ** -23)
    # 1 + ulp/2 is a tie: to even (1.0); 1 + 3 ulp/2 is a tie: to even (1 + 2 ulp)
    assert rr.fma32(np.float32(0.5), ulp, one) == one
    assert rr.fma32(np.float32(1.5), ulp, one) == np.float32(1.0 + 2.0 ** -22)
    # a product that the unfused sequence rounds differently: (1 + 2^-12)^2 - 1 = 2^-11 + 2^-24 (the square's 2^-24 is a
    # tie at 1.0's precision and is rounded away before the subtraction)
    x = np.float32(1.0 + 2.0 ** -12)
    fused, unfused = rr.fma32(x, x, np.float32(-1.0)), np.float32(np.float32(x * x) - np.float32(1.0))
    assert unfused == np.float32(2.0 ** -11) and fused == np.float32(2.0 ** -11 + 2.0 ** -24)
    assert rr.f32_from_fraction(rr.Fraction(1, 2 ** 149)) == np.float32(1e-45)  # smallest subnormal


@pytest.mark.parametrize("dim", [1, 7, 8, 9, 16, 33, 96, 100])
@pytest.mark.parametrize("measure", ["dot", "sql2"])
def test_oracle_one_to_many_equals_exact_fma_restatement(oracle, dim, measure):
    # 11 rows: the 3-row (Dot) and 4-row (SqL2) batches AND their remainder rows; magnitudes spread over six decades so
    # that a different summation order or an unfused multiply-add would change low bits
    rng = np.random.default_rng(100 + dim)
    q = (rng.standard_normal(dim) * 10.0 ** rng.integers(-3, 3, dim)).astype(np.float32)
    db = (rng.standard_normal((11, dim)) * 10.0 ** rng.integers(-3, 3, (11, dim))).astype(np.float32)
    om = oracle.DOT if measure == "dot" else oracle.SQL2
    got = oracle.one_to_many(q, db, om)
    want = rr.one_to_many_f32(q, db, measure)
    assert (got.view(np.uint32) == want.view(np.uint32)).all()
    # and the order matters on this data: a plain left-to-right f32 sum differs somewhere
    if dim >= 16:
        naive = np.array([np.float32(sum((np.float32(a) * np.float32(b) if measure == "dot" else
                                          np.float32(np.float32(a - b) * np.float32(a - b))) for a, b in zip(q, x)))
                          for x in db], np.float32)
        naive = -naive if measure == "dot" else naive
        assert (naive.view(np.uint32) != want.view(np.uint32)).any()


@pytest.mark.parametrize("dim", [5, 8, 24, 43, 128])
@pytest.mark.parametrize("measure", ["dot", "sql2"])
def test_oracle_int8_one_to_many_equals_exact_fma_restatement(oracle, dim, measure):
    rng = np.random.default_rng(200 + dim)
    q = (rng.standard_normal(dim) * 3).astype(np.float32)
    db = rng.integers(-128, 128, (9, dim), dtype=np.int8)
    inv = np.float32(0.0234375 * 1.37)
    om = oracle.DOT if measure == "dot" else oracle.SQL2
    got = oracle.one_to_many_i8(q, db, float(inv), om)
    want = rr.one_to_many_i8(q, db, inv, measure)
    assert (got.view(np.uint32) == want.view(np.uint32)).all()


@pytest.mark.parametrize("measure", ["sql2", "dot"])
@pytest.mark.parametrize("n,dim,K,S,L,R,k,seed", [(1200, 16, 6, 8, 3, 30, 10, 11), (700, 24, 5, 12, 5, 12, 12, 12)])
def test_oracle_tree_ah_end_to_end_equals_python_restatement(oracle, measure, n, dim, K, S, L, R, k, seed):
    # the whole TreeXHybridSearcher::search of the restatement (partition -> residual LUT16 scan -> per-leaf
    # FastTopNeighbors -> stable merge -> exact reorder with the AVX2 + FMA single-pair kernel -> stable sort -> top-k)
    # against the oracle's final (id, distance) lists, bit for bit and in order; the grid makes exact distance ties real
    x, _ = helpers.clustered(n, dim, 8, 0.4, seed, normalize=False)
    x = (np.round(x * 2) / 2).astype(np.float32)
    idx = helpers.build_index(oracle, x, K, S, seed=seed, iters=4)
    q = (x[::max(1, n // 8)][:8] + np.float32(0.25)).astype(np.float32)
    om = oracle.SQL2 if measure == "sql2" else oracle.DOT
    rc, oids, odists, ocounts = oracle.treex_search(idx["centers"], idx["codebook"], idx["part_offsets"], idx["ids"],
                                                    idx["packed"], x, q, L, R, k, lut16=True, reorder_measure=om)
    assert rc == 0
    ties = 0
    for i in range(len(q)):
        cand = rr.approx_candidates(idx["centers"], idx["codebook"], idx["part_offsets"], idx["ids"], idx["packed"], q[i],
                                    L, R)
        want = rr.reorder_results(x, q[i], [j for j, _ in cand], k, measure)
        c = int(ocounts[i])
        assert c == len(want)
        wd = np.array([d for _, d in want], np.float32)
        wi = np.array([j for j, _ in want], np.uint32)
        assert (odists[i, :c].view(np.uint32) == wd.view(np.uint32)).all(), "exact distances differ"
        assert (oids[i, :c] == wi).all(), "ids / tie order differ"
        ties += int((np.diff(wd) == 0).sum())
    assert ties > 0 or measure == "dot", "the SqL2 case was meant to contain exact distance ties"


@pytest.mark.parametrize("measure", ["sql2", "l2", "dot"])
@pytest.mark.parametrize("n,dim,k", [(300, 12, 10), (150, 20, 40), (7, 5, 10)])
def test_oracle_brute_force_end_to_end_equals_python_restatement(oracle, measure, n, dim, k):
    # BruteForceSearcher::search: one-to-many kernel -> N heap pushes -> drain_sorted.  Grid data: many exactly equal
    # distances, so the ids the heap keeps at the cut-off and the ORDER inside ties are exercised; k > n is clamped
    rng = np.random.default_rng(n + dim)
    db = rng.integers(-2, 3, (n, dim)).astype(np.float32)
    q = rng.integers(-2, 3, (6, dim)).astype(np.float32)
    om = {"sql2": oracle.SQL2, "l2": oracle.L2, "dot": oracle.DOT}[measure]
    rc, oids, odists, ocounts = oracle.bf_search(db, q, k, om, nthreads=2)
    assert rc == 0
    ties = 0
    for i in range(len(q)):
        want = rr.bf_search(db, q[i], k, measure)
        c = int(ocounts[i])
        assert c == len(want) == min(k, n)
        wd = np.array([d for _, d in want], np.float32)
        assert (odists[i, :c].view(np.uint32) == wd.view(np.uint32)).all()
        assert oids[i, :c].tolist() == [j for j, _ in want], "ids / tie order differ"
        ties += int((np.diff(wd) == 0).sum())
    assert ties > 0


@pytest.mark.parametrize("measure", ["sql2", "dot"])
def test_oracle_sq8_search_end_to_end_equals_python_restatement(oracle, measure):
    # quantiser (sequential f64 statistics, 3-sigma clipping, wrap to i8) -> sign-extending int8 kernels -> TopK
    rng = np.random.default_rng(77)
    db = (rng.integers(-6, 7, (200, 16)) * 0.25).astype(np.float32)
    q = (rng.integers(-6, 7, (5, 16)) * 0.25).astype(np.float32)
    codes, (lo, hi, sc, inv) = rr.sq8_quantize(db)
    ocodes, ocal = oracle.sq8_quantize(db)
    assert (codes == ocodes).all() and np.float32(sc).view(np.uint32) == ocal[2].view(np.uint32)
    om = oracle.SQL2 if measure == "sql2" else oracle.DOT
    rc, oids, odists, ocounts = oracle.sq8_search(ocodes, float(ocal[2]), q, 10, om, nthreads=2)
    assert rc == 0
    for i in range(len(q)):
        want = rr.sq8_search(codes, sc, q[i], 10, measure)
        wd = np.array([d for _, d in want], np.float32)
        assert (odists[i].view(np.uint32) == wd.view(np.uint32)).all()
        assert oids[i].tolist() == [j for j, _ in want], "ids / tie order differ"


# ----------------------------------------------------------------------------- PQ pieces and the flat AsymmetricHasher
@pytest.mark.parametrize("S,C,ds", [(8, 16, 2), (5, 16, 3), (4, 256, 4)])
def test_oracle_pq_encode_pack_lut_equal_python_restatement(oracle, S, C, ds):
    rng = np.random.default_rng(S * 100 + C)
    cb = (rng.integers(-3, 4, (S, C, ds)) * 0.5).astype(np.float32)  # coarse grid: equidistant codewords exist
    cb[:, 3] = cb[:, 1]                                               # and exact duplicates: the LOWER code must win
    x = (rng.integers(-3, 4, (60, S * ds)) * 0.5).astype(np.float32)
    ocodes = oracle.pq_encode(cb, x)
    want = np.array([rr.pq_encode(cb, r) for r in x], np.uint8)
    assert (ocodes == want).all()
    assert not (want == 3).any()
    if C == 16:
        assert (oracle.pack4(ocodes) == np.array([rr.pack4(list(r)) for r in want], np.uint8)).all()
    q = (rng.standard_normal(S * ds) * 1.7).astype(np.float32)
    olut = oracle.lut_f32(cb, q)
    wlut = np.array(rr.lut_f32(cb, q), np.float32)
    assert (olut.view(np.uint32) == wlut.view(np.uint32)).all()
    od = oracle.lut_f32_scan(olut, ocodes)
    wd = np.array([rr.lut_f32_distance(rr.lut_f32(cb, q), r) for r in want], np.float32)
    assert (od.view(np.uint32) == wd.view(np.uint32)).all()


@pytest.mark.parametrize("k,pre_k", [(10, 0), (5, 25), (40, 40)])
def test_oracle_asymmetric_hasher_equals_python_restatement(oracle, k, pre_k):
    # flat AH: f32 LUT -> scan of all N -> FastTopNeighbors(k) [-> exact SqL2 re-rank of pre_reorder_k candidates]
    rng = np.random.default_rng(k + pre_k)
    S, C, ds, n = 6, 16, 2, 250
    cb = (rng.integers(-3, 4, (S, C, ds)) * 0.5).astype(np.float32)
    x = (rng.integers(-3, 4, (n, S * ds)) * 0.5).astype(np.float32)
    codes = oracle.pq_encode(cb, x)
    q = (rng.integers(-3, 4, (5, S * ds)) * 0.5).astype(np.float32)
    rc, oids, odists, ocounts = oracle.ah_search(cb, codes, q, k, lut16=False, raw=x if pre_k else None, pre_k=pre_k,
                                                 nthreads=2)
    assert rc == 0
    for i in range(len(q)):
        want = (rr.ah_search_with_reordering(cb, codes, x, q[i], k, pre_k) if pre_k else rr.ah_search(cb, codes, q[i], k))
        c = int(ocounts[i])
        assert c == len(want)
        wd = np.array([d for _, d in want], np.float32)
        assert (odists[i, :c].view(np.uint32) == wd.view(np.uint32)).all()
        assert oids[i, :c].tolist() == [j for j, _ in want], "ids / tie order differ"


# ----------------------------------------------------------------------------- the Scann facade's tree modes
def _facade_case(oracle, seed, n=400, dim=12, K=7, S=6):
    rng = np.random.default_rng(seed)
    x = (rng.integers(-4, 5, (n, dim)) * 0.5).astype(np.float32)
    centers = x[rng.permutation(n)[:K]].copy()
    tok = np.array([rr.partition(centers, r, 1)[0] for r in x])
    order = np.argsort(tok, kind="stable").astype(np.uint32)
    off = np.zeros(K + 1, np.uint64)
    off[1:] = np.cumsum(np.bincount(tok, minlength=K))
    cb = (rng.integers(-3, 4, (S, 16, dim // S)) * 0.5).astype(np.float32)
    q = (rng.integers(-4, 5, (6, dim)) * 0.5).astype(np.float32)
    return x, centers, off, order, cb, q


@pytest.mark.parametrize("measure", ["sql2", "l2", "dot"])
def test_oracle_scann_search_partitioned_equals_python_restatement(oracle, measure):
    x, centers, off, order, cb, q = _facade_case(oracle, 5)
    om = {"sql2": oracle.SQL2, "l2": oracle.L2, "dot": oracle.DOT}[measure]
    L, k = 3, 15
    rc, oids, odists, ocounts = oracle.scann_partitioned(centers, off, order, x, q, L, k, om, nthreads=2)
    assert rc == 0
    for i in range(len(q)):
        want = rr.scann_search_partitioned(centers, off, order, x, q[i], L, k, measure)
        c = int(ocounts[i])
        assert c == len(want)
        wd = np.array([d for _, d in want], np.float32)
        assert (odists[i, :c].view(np.uint32) == wd.view(np.uint32)).all()
        assert oids[i, :c].tolist() == [j for j, _ in want], "ids / tie order differ"


@pytest.mark.parametrize("reorder", [None, "sql2", "dot"])
def test_oracle_scann_search_tree_ah_equals_python_restatement(oracle, reorder):
    x, centers, off, order, cb, q = _facade_case(oracle, 6)
    codes = oracle.pq_encode(cb, x)
    L, k = 3, 12
    om = -1 if reorder is None else {"sql2": oracle.SQL2, "dot": oracle.DOT}[reorder]
    rc, oids, odists, ocounts = oracle.scann_tree_ah(centers, off, order, cb, codes, x if reorder else None, q, L, k,
                                                     reorder_measure=om, nthreads=2)
    assert rc == 0
    for i in range(len(q)):
        want = rr.scann_search_tree_ah(centers, off, order, cb, codes, q[i], L, k)
        if reorder:
            want = rr.reordering_helper(x, q[i], want, k, reorder)
        c = int(ocounts[i])
        assert c == len(want)
        wd = np.array([d for _, d in want], np.float32)
        assert (odists[i, :c].view(np.uint32) == wd.view(np.uint32)).all()
        assert oids[i, :c].tolist() == [j for j, _ in want], "ids / tie order differ"


# ----------------------------------------------------------------------------- KMeansTree::search_leaves
@pytest.mark.parametrize("seed,dim,fanout,depth_max,k", [(1, 3, 4, 4, 1), (2, 2, 5, 3, 3), (7, 3, 4, 5, 8), (4, 1, 6, 3, 2), (9, 2, 3, 6, 20)])
def test_oracle_kmtree_search_leaves_equals_python_restatement(oracle, seed, dim, fanout, depth_max, k):
    import test_kmtree
    rng = np.random.default_rng(seed)
    centers, depth, cb, cc, ch = test_kmtree.random_tree(rng, dim, fanout, depth_max)
    centers = np.round(centers).astype(np.float32)  # integer grid: equal child distances, so the stable orders matter
    q = np.round(rng.standard_normal((12, dim))).astype(np.float32)
    nodes, dists, depths, counts = oracle.kmtree_search_leaves(centers, depth, cb, cc, ch, q, k)
    ties = 0
    for i in range(len(q)):
        want = rr.kmtree_search_leaves(centers, depth, cb, cc, ch, q[i], k)
        c = int(counts[i])
        assert c == len(want)
        assert nodes[i, :c].tolist() == [n for n, _, _ in want]
        wd = np.array([d for _, d, _ in want], np.float32)
        assert (dists[i, :c].view(np.uint32) == wd.view(np.uint32)).all()
        assert depths[i, :c].tolist() == [d for _, _, d in want]
        alld = [float(rr.sqdist_seq(q[i], c)) for c in centers]
        ties += len(alld) - len(set(alld))
    assert ties > 0, "the grid was meant to give equal centre distances (stable child order / final sort exercised)"


# ----------------------------------------------------------------------------- restrict filter, radius search
@pytest.mark.parametrize("keep", [0.05, 0.5, 0.97])
def test_oracle_search_with_filter_equals_python_restatement(oracle, keep):
    n, dim, K, S, L, R, k, seed = 900, 16, 5, 8, 4, 20, 8, 21
    x, _ = helpers.clustered(n, dim, 8, 0.4, seed, normalize=False)
    x = (np.round(x * 2) / 2).astype(np.float32)
    idx = helpers.build_index(oracle, x, K, S, seed=seed, iters=4)
    q = (x[::max(1, n // 6)][:6] + np.float32(0.25)).astype(np.float32)
    mask = np.random.default_rng(int(keep * 100)).random(n) < keep
    bits = np.packbits(mask, bitorder="little")
    rc, oids, odists, ocounts, ocand, ocd, ocn = oracle.treex_search(
        idx["centers"], idx["codebook"], idx["part_offsets"], idx["ids"], idx["packed"], x, q, L, R, k, lut16=True,
        want_candidates=True, allow=bits)
    assert rc == 0
    allowed = set(np.nonzero(mask)[0].tolist())
    for i in range(len(q)):
        cand = rr.approx_candidates(idx["centers"], idx["codebook"], idx["part_offsets"], idx["ids"], idx["packed"], q[i],
                                    L, R, allow=allowed)
        c = int(ocn[i])
        assert c == len(cand) and ocand[i, :c].tolist() == [j for j, _ in cand]
        assert all(j in allowed for j in ocand[i, :c].tolist())
        want = rr.reorder_results(x, q[i], [j for j, _ in cand], k, "sql2")
        m = int(ocounts[i])
        assert oids[i, :m].tolist() == [j for j, _ in want]
        assert (odists[i, :m].view(np.uint32) == np.array([d for _, d in want], np.float32).view(np.uint32)).all()


@pytest.mark.parametrize("measure,radius", [("sql2", 16.0), ("l2", 4.0), ("dot", -2.0), ("sql2", -1.0)])
def test_oracle_search_radius_equals_python_restatement(oracle, measure, radius):
    rng = np.random.default_rng(31)
    db = rng.integers(-2, 3, (250, 9)).astype(np.float32)
    q = rng.integers(-2, 3, 9).astype(np.float32)
    om = {"sql2": oracle.SQL2, "l2": oracle.L2, "dot": oracle.DOT}[measure]
    oi, od = oracle.bf_search_radius(db, q, radius, om)
    want = rr.bf_search_radius(db, q, radius, measure)
    assert oi.tolist() == [j for j, _ in want]
    assert (od.view(np.uint32) == np.array([d for _, d in want], np.float32).view(np.uint32)).all()
    assert (len(want) > 10) == (radius > 0 or measure == "dot")  # the cases really select something (or nothing)


# ----------------------------------------------------------------------------- randomised sweep
@pytest.mark.parametrize("seed", range(12))
def test_oracle_equals_python_restatement_random_sweep(oracle, seed):
    """Random small configurations (sizes, dims with and without an 8-lane tail, k around and above n, grid coarseness):
    brute force (three measures) and Tree-AH end to end, ids + distances + tie order."""
    rng = np.random.default_rng(1000 + seed)
    n, dim = int(rng.integers(3, 120)), int(rng.choice([3, 8, 11, 16, 20]))
    k = int(rng.integers(1, n + 5))
    step = float(rng.choice([0.25, 0.5, 1.0]))
    db = (rng.integers(-3, 4, (n, dim)) * step).astype(np.float32)
    q = (rng.integers(-3, 4, (3, dim)) * step).astype(np.float32)
    for measure, om in (("sql2", oracle.SQL2), ("l2", oracle.L2), ("dot", oracle.DOT)):
        rc, oids, odists, ocounts = oracle.bf_search(db, q, k, om, nthreads=1)
        for i in range(len(q)):
            want = rr.bf_search(db, q[i], k, measure)
            c = int(ocounts[i])
            assert c == len(want)
            assert oids[i, :c].tolist() == [j for j, _ in want]
            assert (odists[i, :c].view(np.uint32) == np.array([d for _, d in want], np.float32).view(np.uint32)).all()
    # Tree-AH on a separate index: S divides dim, K partitions by nearest of K random rows
    S = int(rng.choice([2, 4]))
    dim2 = S * int(rng.integers(1, 4))
    n2 = int(rng.integers(40, 300))
    x = (rng.integers(-3, 4, (n2, dim2)) * step).astype(np.float32)
    K = int(rng.integers(2, 6))
    idx = helpers.build_index(oracle, x, K, S, seed=seed, iters=3)
    L, R, kk = int(rng.integers(1, K + 1)), int(rng.integers(1, 40)), int(rng.integers(1, 15))
    qq = (rng.integers(-3, 4, (3, dim2)) * step).astype(np.float32)
    for measure, om in (("sql2", oracle.SQL2), ("dot", oracle.DOT)):
        rc, oids, odists, ocounts = oracle.treex_search(idx["centers"], idx["codebook"], idx["part_offsets"], idx["ids"],
                                                        idx["packed"], x, qq, L, R, kk, lut16=True, reorder_measure=om)
        assert rc == 0
        for i in range(len(qq)):
            cand = rr.approx_candidates(idx["centers"], idx["codebook"], idx["part_offsets"], idx["ids"], idx["packed"],
                                        qq[i], L, R)
            want = rr.reorder_results(x, qq[i], [j for j, _ in cand], kk, measure)
            c = int(ocounts[i])
            assert c == len(want)
            assert oids[i, :c].tolist() == [j for j, _ in want]
            assert (odists[i, :c].view(np.uint32) == np.array([d for _, d in want], np.float32).view(np.uint32)).all()
