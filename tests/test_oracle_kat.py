"""Pins the CPU oracle against every known-answer test the reference holds for the hot path
(SURVEY.md §8c).  Each test names the reference test it ports (paths relative to /root/reference).
CPU only."""
import numpy as np
import pytest


def approx(a, b, tol=1e-6):
    return abs(float(a) - float(b)) < tol


# ------------------------------------------------------------------ src/simd/tests.rs:189-231
def test_simd_one_to_many_dot(oracle):
    q = np.arange(1, 9, dtype=np.float32)
    db = np.stack([np.full(8, 1.0), np.full(8, 2.0), np.full(8, 0.5)]).astype(np.float32)
    r = oracle.one_to_many(q, db, oracle.DOT)
    assert approx(r[0], -36.0) and approx(r[1], -72.0) and approx(r[2], -18.0)


def test_simd_one_to_many_sql2(oracle):
    q = np.arange(1, 9, dtype=np.float32)
    db = np.stack([q, np.zeros(8), q + 1]).astype(np.float32)
    r = oracle.one_to_many(q, db, oracle.SQL2)
    assert approx(r[0], 0.0) and approx(r[1], 204.0) and approx(r[2], 8.0)


# ------------------------------------------------------------------ src/simd/tests.rs:237-264
def test_simd_lut16_batch_portable(oracle):
    lut = np.zeros((2, 16), np.uint8)
    lut[0] = np.arange(16)
    lut[1] = 15 - np.arange(16)
    packed = np.array([[0x00], [0x11], [0x0F], [0xF0]], np.uint8)
    r = oracle.lut16_scan_f32(packed, lut, 2)
    assert list(r) == [15.0, 15.0, 30.0, 0.0]
    assert list(oracle.lut16_scan_u32(packed, lut, 2)) == [15, 15, 30, 0]


# ------------------------------------------------------------------ src/hashes/lut16_simd.rs:306-375
def test_lut16_simd_quantization_roundtrip(oracle):
    t0 = np.arange(16, dtype=np.float32)
    t1 = 15 - t0
    l8, bias, mult = oracle.lut16_quantize(np.stack([t0, t1]))
    assert abs(oracle.lut16_distance_single([0, 0], l8, bias, mult) - 15.0) < 0.1
    assert abs(oracle.lut16_distance_single([15, 15], l8, bias, mult) - 15.0) < 0.1
    assert abs(oracle.lut16_distance_single([5, 10], l8, bias, mult) - 10.0) < 0.1
    # quantiser details: range 15 → scale 17 → entries i*17, multiplier 1/17
    assert list(l8[0]) == [17 * i for i in range(16)]
    assert bias == 0.0 and mult == np.float32(1.0) / np.float32(17.0)


def test_lut16_simd_batch_computation(oracle):
    l8, bias, mult = oracle.lut16_quantize(np.arange(16, dtype=np.float32)[None, :])
    packed = np.array([[0x00], [0x05], [0x0A], [0x0F]], np.uint8)
    r = oracle.lut16_distances(packed, l8, 1, bias, mult)
    for got, want in zip(r, [0.0, 5.0, 10.0, 15.0]):
        assert abs(got - want) < 0.1


def test_lut16_simd_two_subspace_packed(oracle):
    t0 = np.arange(16, dtype=np.float32)
    l8, bias, mult = oracle.lut16_quantize(np.stack([t0, 15 - t0]))
    packed = np.array([[0x00], [0x55], [0x0F], [0xF0]], np.uint8)
    r = oracle.lut16_distances(packed, l8, 2, bias, mult)
    for got, want in zip(r, [15.0, 15.0, 30.0, 0.0]):
        assert abs(got - want) < 0.5


# lut16_simd.rs:377-411 batch == single
def test_lut16_simd_batch_equals_single(oracle):
    t = np.arange(16, dtype=np.float32) * 2.0
    l8, bias, mult = oracle.lut16_quantize(np.stack([t, t]))
    n = 100
    lo = np.arange(n) % 16
    hi = (np.arange(n) + 5) % 16
    packed = (lo | (hi << 4)).astype(np.uint8)[:, None]
    r = oracle.lut16_distances(packed, l8, 2, bias, mult)
    for i in range(n):
        assert abs(r[i] - oracle.lut16_distance_single([lo[i], hi[i]], l8, bias, mult)) < 0.5


def test_lut16_degenerate_range(oracle):
    # lut16_simd.rs:63-72: range < 1e-10 → scale 1, multiplier 1, bias = the value
    l8, bias, mult = oracle.lut16_quantize(np.full((3, 16), 2.5, np.float32))
    assert (l8 == 0).all() and bias == 2.5 and mult == 1.0


# ------------------------------------------------------------------ src/hashes/lut16.rs:312-328
def test_lut16_packed_codes_roundtrip(oracle):
    codes = np.array([[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11]], np.uint8)
    packed = oracle.pack4(codes)
    assert packed.shape == (3, 2)
    assert packed[0, 0] == 0x10 and packed[0, 1] == 0x32
    assert (oracle.unpack4(packed, 4) == codes).all()


def test_lut16_packed_codes_odd_subspaces(oracle):
    codes = np.array([[1, 2, 3], [15, 0, 9]], np.uint8)
    packed = oracle.pack4(codes)
    assert packed.shape == (2, 2) and packed[0, 1] == 0x03 and packed[1, 1] == 0x09
    assert (oracle.unpack4(packed, 3) == codes).all()


# ------------------------------------------------------------------ src/hashes/lut16.rs:339-366
def test_lut16_lookup_tables_from_query(oracle):
    cb = np.zeros((2, 16, 2), np.float32)
    cb[0, :, 0] = np.arange(16)
    cb[1, :, 1] = np.arange(16)
    q = np.array([5.0, 0.0, 0.0, 5.0], np.float32)
    lf = oracle.lut_f32(cb, q)
    assert lf[0, 5] + lf[1, 5] < 0.01
    assert abs(lf[0, 0] + lf[1, 0] - 50.0) < 0.01


# ------------------------------------------------------------------ src/distance_measures/one_to_many.rs:380-430
def test_one_to_many_squared_l2(oracle):
    q = np.array([1, 2, 3], np.float32)
    db = np.array([[1, 2, 3], [2, 3, 4], [0, 0, 0]], np.float32)
    r = oracle.one_to_many(q, db, oracle.SQL2)
    assert approx(r[0], 0) and approx(r[1], 3) and approx(r[2], 14)


def test_one_to_many_dot_product(oracle):
    q = np.array([1, 2, 3], np.float32)
    db = np.array([[1, 1, 1], [2, 2, 2]], np.float32)
    r = oracle.one_to_many(q, db, oracle.DOT)
    assert approx(r[0], -6) and approx(r[1], -12)


def test_one_to_many_strided(oracle):
    q = np.array([1, 2], np.float32)
    data = np.array([1, 2, 0, 0, 3, 4, 0, 0], np.float32)
    r = oracle.one_to_many_flat(q, data, 4, 2, oracle.SQL2)
    assert approx(r[0], 0) and approx(r[1], 8)


# ------------------------------------------------------------------ one_to_many_asymmetric.rs:411-451
def test_int8_dot_product(oracle):
    q = np.arange(1, 9, dtype=np.float32)
    db = np.array([[127] * 8, [0] * 8], np.int8)
    r = oracle.one_to_many_i8(q, db, 1.0 / 127.0, oracle.DOT)
    assert abs(r[0] + 36.0) < 1e-3 and abs(r[1]) < 1e-3


def test_int8_squared_l2(oracle):
    q = np.arange(1, 9, dtype=np.float32)
    db = np.array([[127] * 8, [0] * 8], np.int8)
    r = oracle.one_to_many_i8(q, db, 1.0 / 127.0, oracle.SQL2)
    assert abs(r[1] - 204.0) < 0.5


def test_int8_sign_extension_quirk(oracle):
    # the kernel sign-extends the stored byte (levels 128..255 read as -128..-1)
    q = np.ones(8, np.float32)
    db = np.array([[-1] * 8], np.int8)  # stored level 255
    r = oracle.one_to_many_i8(q, db, 0.5, oracle.DOT)
    assert approx(r[0], 4.0)  # -(8 * 1 * (-1*0.5))


# ------------------------------------------------------------------ src/brute_force/top_k.rs:399-465
def test_top_k_basic(oracle):
    ids, dists, acc = oracle.topk_run(3, [0, 1, 2, 3, 4], [5.0, 3.0, 7.0, 4.0, 6.0])
    assert list(acc) == [True, True, True, True, False]
    assert list(ids) == [1, 3, 0] and list(dists) == [3.0, 4.0, 5.0]


def test_top_k_empty(oracle):
    ids, dists, _ = oracle.topk_run(5, [], [])
    assert len(ids) == 0


def test_fast_top_neighbors(oracle):
    ids, dists = oracle.ftn_run(3, [0, 1, 2, 3], [5.0, 3.0, 7.0, 2.0])
    assert len(ids) == 3 and ids[0] == 3 and ids[1] == 1


def test_fast_top_neighbors_batch(oracle):
    ids, dists = oracle.ftn_run(3, [0, 1, 2, 3, 4], [5.0, 3.0, 7.0, 1.0, 4.0], batch_mode=True)
    assert list(ids) == [3, 1, 4]


def test_fixed_top_k_stress_sequence(oracle):
    # top_k.rs:492-515 (FixedTopK; same admission rule as TopK): 10 smallest of (i*7)%100
    d = [float((i * 7) % 100) for i in range(100)]
    ids, dists, _ = oracle.topk_run(10, list(range(100)), d)
    assert len(ids) == 10 and (np.diff(dists) >= 0).all() and dists[-1] < 10.0


def test_top_k_tie_eviction_largest_index_first(oracle):
    # (OrderedFloat, idx) lexicographic max-heap: among equal d the largest idx is evicted first
    ids, dists, _ = oracle.topk_run(2, [0, 1, 2, 3], [1.0, 1.0, 1.0, 0.5])
    assert set(ids) == {0, 3} or set(ids) == {1, 3}
    assert 2 not in set(ids)


# ------------------------------------------------------------------ src/brute_force/searcher.rs:280-377
DS5 = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1]], np.float32)


def test_brute_force_search(oracle):
    rc, ids, dists, counts = oracle.bf_search(DS5, np.zeros((1, 3)), 3, oracle.SQL2)
    assert rc == 0 and counts[0] == 3 and ids[0, 0] == 0 and abs(dists[0, 0]) < 1e-6


def test_brute_force_search_all_sorted(oracle):
    rc, ids, dists, counts = oracle.bf_search(DS5, np.full((1, 3), 0.5), 5, oracle.SQL2)
    assert counts[0] == 5 and (np.diff(dists[0]) >= 0).all()


def test_brute_force_dot(oracle):
    db = np.array([[1, 0], [0, 1], [1, 1]], np.float32)
    rc, ids, dists, counts = oracle.bf_search(db, np.array([[1.0, 0.0]]), 3, oracle.DOT)
    assert counts[0] == 3 and dists[0, 0] <= dists[0, 1]
    assert list(dists[0]) == [-1.0, -1.0, 0.0]


def test_brute_force_radius(oracle):
    ids, dists = oracle.bf_search_radius(DS5, np.zeros(3), 1.5, oracle.SQL2)
    assert len(ids) == 4


def test_brute_force_batched(oracle):
    rc, ids, dists, counts = oracle.bf_search(DS5, np.array([[0, 0, 0], [1, 1, 1]], np.float32), 2, oracle.SQL2,
                                              nthreads=2)
    assert list(counts) == [2, 2] and ids[1, 0] == 4


def test_brute_force_empty_dataset(oracle):
    rc, ids, dists, counts = oracle.bf_search(np.zeros((0, 3), np.float32), np.array([[1, 2, 3.0]]), 5,
                                              oracle.SQL2)
    assert rc == 0 and counts[0] == 0


def test_brute_force_dimension_mismatch(oracle):
    rc, *_ = oracle.bf_search(DS5, np.array([[1.0, 2.0]]), 5, oracle.SQL2)
    assert rc == oracle.INVALID_ARGUMENT


def test_brute_force_k_clamped(oracle):
    rc, ids, dists, counts = oracle.bf_search(DS5, np.zeros((1, 3)), 50, oracle.SQL2)
    assert counts[0] == 5


# tests/unit_tests.rs:191-260 exact match k=1
def test_unit_brute_force_exact_match(oracle):
    rc, ids, dists, counts = oracle.bf_search(DS5, np.zeros((1, 3)), 1, oracle.SQL2)
    assert counts[0] == 1 and ids[0, 0] == 0 and abs(dists[0, 0]) < 1e-6


# tests/unit_tests.rs:145-181 distance KATs that lie on the path
def test_unit_distances(oracle):
    assert approx(oracle.pair_distance(oracle.L2, [0, 0], [3, 4]), 5.0)
    assert approx(oracle.pair_distance(oracle.SQL2, [0, 0], [3, 4]), 25.0)
    assert approx(oracle.pair_distance(oracle.DOT, [1, 2, 3], [4, 5, 6]), -32.0)


# ------------------------------------------------------------------ tests/stress_tests.rs:325-363 (property)
def test_stress_recall_verification_property(oracle):
    rng = np.random.default_rng(42)
    db = rng.random((1000, 32), dtype=np.float32)
    q = np.random.default_rng(123).random((1, 32), dtype=np.float32)
    rc, ids, dists, counts = oracle.bf_search(db, q, 10, oracle.SQL2)
    alld = np.array([oracle.pair_distance(oracle.SQL2, q[0], db[i]) for i in range(1000)], np.float32)
    order = np.argsort(alld, kind="stable")
    assert list(ids[0]) == list(order[:10])
    assert np.abs(dists[0] - alld[order[:10]]).max() < 1e-5


# ------------------------------------------------------------------ src/brute_force/scalar_quantized.rs:424-513
def test_scalar_quantized_search(oracle):
    db = np.array([[0, 0, 0], [10, 0, 0], [0, 10, 0], [0, 0, 10], [10, 10, 10]], np.float32)
    codes, cal = oracle.sq8_quantize(db)
    rc, ids, dists, counts = oracle.sq8_search(codes, cal[2], np.zeros((1, 3)), 3, oracle.SQL2)
    assert counts[0] == 3 and ids[0, 0] == 0


def test_scalar_quantized_dot(oracle):
    db = np.array([[10, 0], [0, 10], [10, 10]], np.float32)
    codes, cal = oracle.sq8_quantize(db)
    rc, ids, dists, counts = oracle.sq8_search(codes, cal[2], np.array([[1.0, 0.0]]), 3, oracle.DOT)
    assert counts[0] == 3


def test_scalar_quantized_accuracy_vs_float(oracle):
    db = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9], [1.1, 2.1, 3.1], [10, 0, 0]], np.float32)
    q = np.array([[1.0, 2.0, 3.0]], np.float32)
    _, fids, _, _ = oracle.bf_search(db, q, 3, oracle.SQL2)
    codes, cal = oracle.sq8_quantize(db)
    _, qids, _, _ = oracle.sq8_search(codes, cal[2], q, 3, oracle.SQL2)
    assert fids[0, 0] == qids[0, 0]


# ------------------------------------------------------------------ src/quantization/scalar.rs:411-431
def test_scalar_quantizer_basic(oracle):
    cal = oracle.sq8_calibrate([-1.0, 1.0, 0.0, 0.5])
    assert approx(cal[0], -1.0) and approx(cal[1], 1.0)
    qv = oracle.sq8_quantize_value(0.5, cal)
    assert abs(0.5 - oracle.sq8_dequantize_value(qv, cal)) < 0.02
    # 0.5 → level round(1.5*127.5)=191 → stored as i8 wraps to -65
    assert qv == -65


# src/quantization/scalar.rs:433-454
def test_quantized_dataset(oracle):
    db = np.array([[1, 2, 3], [4, 5, 6], [-1, 0, 1]], np.float32)
    codes, cal = oracle.sq8_quantize(db)
    assert codes.shape == (3, 3)
    for j, want in enumerate([4.0, 5.0, 6.0]):
        assert abs(oracle.sq8_dequantize_value(int(codes[1, j]), cal) - want) < 1.0


# src/quantization/mod.rs:151-163
def test_quantization_stats(oracle):
    st = oracle.sq8_stats(np.array([[1, 2, 3], [4, 5, 6]], np.float32))
    assert st[0] == 1.0 and st[1] == 6.0 and abs(st[2] - 3.5) < 0.01
    assert abs(st[3] - np.std([1, 2, 3, 4, 5, 6], ddof=1)) < 1e-6


# ------------------------------------------------------------------ src/utils/reordering.rs:102-122
def test_reordering(oracle):
    raw = np.array([[0, 0], [1, 0], [2, 0], [3, 0]], np.float32)
    ids, dists = oracle.reorder(raw, oracle.SQL2, np.zeros(2), [2, 1, 3, 0], 3)
    assert list(ids) == [0, 1, 2]


# ------------------------------------------------------------------ src/hashes/lut.rs:281-294 (int8 LUT ≈ f32 LUT)
def test_lut_int8_close_to_f32(oracle):
    rng = np.random.default_rng(0)
    cb = rng.random((4, 16, 4), dtype=np.float32)
    q = np.full(16, 0.5, np.float32)
    lf = oracle.lut_f32(cb, q)
    l8, bias, mult = oracle.lut16_quantize(lf)
    codes = np.array([[0, 1, 2, 3]], np.uint8)
    f = oracle.lut_f32_scan(lf, codes)[0]
    g = oracle.lut16_distance_single(codes[0], l8, bias, mult)
    assert abs(f - g) < 2.0


# ------------------------------------------------------------------ dataset.rs:90-96 stride rule
def test_dense_stride(oracle):
    assert oracle.dense_stride(128) == 128 and oracle.dense_stride(96) == 96 and oracle.dense_stride(3) == 16


# ------------------------------------------------------------------ partition: stable tie → lower centre id
def test_partition_order_and_ties(oracle):
    centers = np.array([[1, 0], [0, 1], [5, 5], [0, -1]], np.float32)
    tokens, dists = oracle.partition(centers, np.zeros((1, 2)), 3)
    assert list(tokens[0]) == [0, 1, 3] and list(dists[0]) == [1.0, 1.0, 1.0]
    tokens, _ = oracle.partition(centers, np.zeros((1, 2)), 10)
    assert list(tokens[0][:4]) == [0, 1, 3, 2] and tokens[0][4] == 0xFFFFFFFF


# ------------------------------------------------------------------ structural tests (tree_x_hybrid/mod.rs:436-468)
def test_treex_search_structural(oracle):
    rng = np.random.default_rng(1)
    n, dim, K, S = 500, 32, 10, 8
    x = np.sin(np.arange(n)[:, None] * np.arange(dim)[None, :] / 100.0).astype(np.float32)
    centers = x[rng.choice(n, K, replace=False)]
    assign = oracle.partition(centers, x, 1)[0][:, 0]
    cb = rng.normal(0, 0.3, (S, 16, dim // S)).astype(np.float32)
    order = np.argsort(assign, kind="stable")
    part_off = np.concatenate([[0], np.cumsum(np.bincount(assign, minlength=K))]).astype(np.uint64)
    codes = oracle.pq_encode_residual(cb, x[order], centers, assign[order])
    q = np.sin(np.arange(dim) / 10.0).astype(np.float32)[None, :]
    for lut16 in (False, True):
        c = oracle.pack4(codes) if lut16 else codes
        rc, ids, dists, counts = oracle.treex_search(centers, cb, part_off, order.astype(np.uint32), c, x, q, 3, 30,
                                                     10, lut16=lut16)
        assert rc == 0 and counts[0] == 10 and (np.diff(dists[0]) >= 0).all()
        # reordered distances are exact SqL2
        for i, d in zip(ids[0], dists[0]):
            assert d == np.float32(oracle.pair_distance(oracle.SQL2, q[0], x[i]))


# ----------------------------------------------------------------------------- restrict filter (oracle side)
def test_oracle_search_with_filter_properties(oracle):
    """TreeXHybridSearcher::search_with_filter (tree_x_hybrid/mod.rs:245-250, 327-332) in the oracle: an all-allowed
    filter is the plain search; a real filter only returns allowed datapoints and equals the plain search of an index
    whose partitions contain only the allowed rows (the reference skips filtered-out rows before the per-leaf top-k)."""
    import helpers
    x, _ = helpers.clustered(4000, 16, 12, 0.35, 3)
    q = (x[:20] + 0.03).astype(np.float32)
    idx = helpers.build_index(oracle, x, 10, 4)
    args = (idx["centers"], idx["codebook"], idx["part_offsets"], idx["ids"], idx["packed"], x, q, 4, 30, 10)
    rc, ids0, d0, c0 = oracle.treex_search(*args, lut16=True)
    rc, ids1, d1, c1 = oracle.treex_search(*args, lut16=True, allow=np.packbits(np.ones(4000, bool), bitorder="little"))
    assert (ids0 == ids1).all() and (d0.view(np.uint32) == d1.view(np.uint32)).all()
    allowed = np.random.default_rng(1).random(4000) < 0.4
    rc, ids2, d2, c2 = oracle.treex_search(*args, lut16=True, allow=np.packbits(allowed, bitorder="little"))
    assert allowed[ids2[ids2 != 0xFFFFFFFF]].all()
    keep = allowed[idx["ids"]]
    off = idx["part_offsets"].astype(np.int64)
    cnt = np.array([keep[off[i]:off[i + 1]].sum() for i in range(len(off) - 1)])
    off2 = np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint64)
    rc, ids3, d3, c3 = oracle.treex_search(idx["centers"], idx["codebook"], off2, np.ascontiguousarray(idx["ids"][keep]),
                                           np.ascontiguousarray(idx["packed"][keep]), x, q, 4, 30, 10, lut16=True)
    assert (ids2 == ids3).all() and (d2.view(np.uint32) == d3.view(np.uint32)).all() and (c2 == c3).all()
