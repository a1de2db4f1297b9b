"""CPU-only tests: the C-ABI library loads and exports every symbol include/scann_b200.h declares, fails
loudly without a GPU (no CPU fallback), and the host-side logic (config arithmetic, sharding, oracle
end-to-end consistency) behaves like the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol(pkg):
    path = pkg.build_lib.build()
    assert os.path.exists(path)
    lib = C.CDLL(path)
    header = open(os.path.join(ROOT, "include", "scann_b200.h")).read()
    declared = set(re.findall(r"\b(scann_[a-z0-9_]+)\s*\(", header))
    declared -= {"scann_status"}
    assert declared == set(pkg.capi.SYMBOLS), declared ^ set(pkg.capi.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"
    assert pkg.load().scann_version() >= 100


def test_no_cpu_fallback_without_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.ScannError) as e:
        pkg.BruteForceSearcher(np.zeros((4, 4), np.float32))
    assert e.value.code == pkg.capi.UNAVAILABLE
    # straight through the C ABI too
    h = C.c_void_p()
    x = np.zeros((4, 4), np.float32)
    st = pkg.load().scann_bf_create(x.ctypes.data_as(C.c_void_p), 4, 4, 4, 0, 0, 0, C.byref(h))
    assert st == pkg.capi.UNAVAILABLE and not h.value
    assert b"no CPU fallback" in pkg.load().scann_last_error()


def test_product_package_never_imports_the_oracle():
    pkgdir = os.path.join(ROOT, "scann-rust_b200")
    for dp, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "libscann_oracle" not in src, f


def test_pre_reorder_k_matches_rust_cast(pkg):
    cfg = pkg.TreeXHybridConfig()
    assert cfg.pre_reorder_k(10) == 30           # (10 as f32 * 3.0) as usize
    cfg.pre_reorder_multiplier = 10.0
    assert cfg.pre_reorder_k(10) == 100
    cfg.pre_reorder_multiplier = 0.05
    assert cfg.pre_reorder_k(10) == 0            # truncation toward zero
    cfg.pre_reorder_multiplier = 2.55
    assert cfg.pre_reorder_k(3) == 7


def test_builder_mode_selection(pkg):
    b = pkg.ScannBuilder().num_neighbors(7).tree(100, 10).hash(8).reorder(50)
    assert b.config.num_partitions == 100 and b.config.hash_num_blocks == 8 and b.config.reorder_num_candidates == 50
    b2 = pkg.ScannBuilder().partitioned(10, 2).hashed(4)  # README / north-star spellings
    assert b2.config.num_partitions == 10 and b2.config.hash_num_blocks == 4


def test_ragged_query_batch_rejected(pkg):
    from importlib import import_module
    s = import_module("scann-rust_b200.searchers")
    with pytest.raises(pkg.ScannError) as e:
        s._Batch([[1.0, 2.0], [1.0]], 0)
    assert e.value.code == pkg.capi.INVALID_ARGUMENT


def test_shard_index_round_robin_inside_partitions(pkg, oracle):
    x = helpers.gaussian(2000, 16, 3)
    idx = helpers.build_index(oracle, x, 9, 4)
    world = 4
    seen = []
    for r in range(world):
        sh = pkg.indexing.shard_index(idx, r, world)
        off = sh["part_offsets"].astype(np.int64)
        full = idx["part_offsets"].astype(np.int64)
        for leaf in range(9):
            members = idx["ids"][full[leaf]:full[leaf + 1]]
            assert (sh["ids"][off[leaf]:off[leaf + 1]] == members[r::world]).all()
        seen.append(sh["ids"])
    assert sorted(np.concatenate(seen).tolist()) == list(range(2000))


def test_oracle_sharded_search_merges_to_unsharded_superset(oracle, pkg):
    """SURVEY §8e caveat: per-shard top-R is taken over 1/world of each leaf, so the union of shard results
    contains the single-shard result set; after the (dist, id) merge the top-k can only improve."""
    x, _ = helpers.clustered(6000, 32, 16, 0.3, 2)
    q = x[:40] + 0.02
    idx = helpers.build_index(oracle, x, 12, 8)
    L, R, k = 4, 40, 10
    _, ids, dists, _ = oracle.treex_search(idx["centers"], idx["codebook"], idx["part_offsets"], idx["ids"],
                                           idx["packed"], x, q, L, R, k)
    parts_i, parts_d = [], []
    for r in range(2):
        sh = pkg.indexing.shard_index(idx, r, 2)
        _, i2, d2, _ = oracle.treex_search(sh["centers"], sh["codebook"], sh["part_offsets"], sh["ids"], sh["packed"],
                                           x, q, L, R, k)
        parts_i.append(i2)
        parts_d.append(d2)
    pi, pd = np.stack(parts_i), np.stack(parts_d)
    merged_d = np.sort(pd.transpose(1, 0, 2).reshape(len(q), -1), axis=1)[:, :k]
    assert (merged_d <= dists + 1e-7).all()


def test_oracle_treex_lut16_vs_f32lut_recall(oracle):
    """The LUT16 composition tracks the reference's f32-LUT TreeX closely after reorder."""
    x, _ = helpers.clustered(8000, 32, 32, 0.3, 6)
    q = x[:50] + 0.01
    idx = helpers.build_index(oracle, x, 16, 8)
    _, a, _, _ = oracle.treex_search(idx["centers"], idx["codebook"], idx["part_offsets"], idx["ids"], idx["codes"],
                                     x, q, 6, 100, 10, lut16=False)
    _, b, _, _ = oracle.treex_search(idx["centers"], idx["codebook"], idx["part_offsets"], idx["ids"], idx["packed"],
                                     x, q, 6, 100, 10, lut16=True)
    assert helpers.recall(a, b, 10) > 0.95
    _, gt, _, _ = oracle.bf_search(x, q, 10, oracle.SQL2)
    assert helpers.recall(b, gt, 10) > 0.6


def test_oracle_scann_modes(oracle):
    x = helpers.gaussian(1500, 16, 8)
    q = helpers.gaussian(10, 16, 9)
    idx = helpers.build_index(oracle, x, 8, 4, use_residuals=False)
    rc, ids, dists, counts = oracle.scann_partitioned(idx["centers"], idx["part_offsets"], idx["ids"], x, q, 8, 5,
                                                      oracle.SQL2)
    _, gt, gd, _ = oracle.bf_search(x, q, 5, oracle.SQL2)
    assert (ids == gt).all()  # all partitions searched == brute force
    codes_by_id = oracle.pq_encode(idx["codebook"], x)
    rc, ids2, d2, c2 = oracle.scann_tree_ah(idx["centers"], idx["part_offsets"], idx["ids"], idx["codebook"],
                                            codes_by_id, x, q, 3, 5)
    assert rc == 0 and (c2 == 5).all() and (np.diff(d2, axis=1) >= 0).all()


def test_search_batched_with_params_groups_queries(pkg):
    """Searcher::search_batched_with_params (src/searcher.rs:164-169) host logic: queries sharing their parameters form
    one batch, results come back in query order, and a length mismatch is InvalidArgument (searcher.rs:233-237)."""
    from importlib import import_module
    s = import_module("scann-rust_b200.searchers")
    calls = []

    class Stub(s._Handle):  # no GPU: search_batched is replaced by a recorder with a deterministic answer
        def search_batched(self, queries, k, pre_reorder_k=None):
            q = np.asarray(queries, np.float32)
            calls.append((len(q), k, pre_reorder_k))
            ids = (q[:, :1].astype(np.uint32) * 100 + np.arange(k, dtype=np.uint32)[None, :])
            return ids, ids.astype(np.float32), np.full(len(q), k, np.uint32)

    st = Stub()
    q = np.arange(5, dtype=np.float32)[:, None] * np.ones((1, 3), np.float32)
    P = s.SearchParameters
    params = [P().with_num_neighbors(2), P().with_num_neighbors(4), P().with_num_neighbors(2),
              P().with_num_neighbors(2).with_pre_reordering_neighbors(50), P()]
    out = st.search_batched_with_params(q, params, default_k=3)
    assert sorted(calls) == [(1, 2, 50), (1, 3, None), (1, 4, None), (2, 2, None)]
    assert [len(r) for r in out] == [2, 4, 2, 2, 3]
    assert [r[0][0] for r in out] == [0, 100, 200, 300, 400]          # query order preserved
    with pytest.raises(pkg.ScannError) as e:
        st.search_batched_with_params(q, params[:3])
    assert e.value.code == pkg.capi.INVALID_ARGUMENT
