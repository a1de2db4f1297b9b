"""Round-2 GPU parity tests (through the C ABI, against the CPU oracle on the same seeded inputs):
  * the certified bf16 ranking threshold on adversarial low-dimensional data (values at bf16 rounding midpoints),
  * scann_treeah_create_ex layout flags (raw rows by position / borrowed raw),
  * bit-exact SQ8 calibration and codes."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


# ----------------------------------------------------------------------------- bf16 certification
def _bf16_midpoints(rng, shape):
    """f32 values that sit exactly half-way between two bf16 neighbours (worst case of the RN operand rounding),
    with random sign and magnitude, so that |q~.x~ - q.x| reaches its bound 2^-7 |q||x| in low dimension."""
    mant = rng.integers(0, 128, shape).astype(np.uint32)           # 7 explicit bf16 mantissa bits
    expo = rng.integers(126, 128, shape).astype(np.uint32)         # magnitudes in [0.5, 2): keeps the lists short
    sign = rng.integers(0, 2, shape).astype(np.uint32)
    bits = (sign << 31) | (expo << 23) | (mant << 16) | np.uint32(0x8000)   # + half an ulp of bf16
    return bits.view(np.float32)


@pytest.mark.parametrize("dim", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("measure", ["SquaredL2", "DotProduct"])
def test_bf_exact_on_bf16_midpoint_data(gpu_lib, oracle, dim, measure):
    rng = np.random.default_rng(1000 + dim)
    n, nq, k = 20_000, 256, 10
    db = _bf16_midpoints(rng, (n, dim))
    q = _bf16_midpoints(rng, (nq, dim))
    m = gpu_lib.DistanceMeasure[measure]
    om = {"SquaredL2": oracle.SQL2, "DotProduct": oracle.DOT}[measure]
    bf = gpu_lib.BruteForceSearcher(db, m)
    ids, dists, counts = bf.search_batched(q, k)
    tc, legacy = bf.path_stats()
    assert tc > 0 or legacy > 0
    if dim >= 2:  # dim 1 has so many near-ties inside 2*eps that the list may overflow to the (equally exact) legacy path
        assert tc > 0, "the tensor-core ranking path did not run"
    rc, oids, odists, ocounts = oracle.bf_search(db, q, k, om, nthreads=8)
    assert (counts == ocounts).all()
    # the k distances are the exact k smallest (ids may differ only inside exact ties, which this data has many of)
    assert (dists.view(np.uint32) == odists.view(np.uint32)).all()
    compared, mism = helpers.ids_equal_away_from_ties(ids, dists, oids, odists, counts)
    assert mism == 0


@pytest.mark.parametrize("dim", [2, 8])
def test_partition_exact_on_bf16_midpoint_centres(gpu_lib, oracle, dim):
    rng = np.random.default_rng(7 + dim)
    K, nq, L = 1024, 300, 16
    cen = _bf16_midpoints(rng, (K, dim))
    q = _bf16_midpoints(rng, (nq, dim))
    part = gpu_lib.TreePartitioner(cen)
    tok, dd = part.partition(q, L)
    otok, odd = oracle.partition(cen, q, L, nthreads=8)
    assert (dd.view(np.uint32) == odd.view(np.uint32)).all()
    same = tok == otok
    # tokens equal wherever the distance is unique in the list (duplicated centres tie exactly)
    for i in range(nq):
        for j in range(L):
            if not same[i, j]:
                assert (odd[i] == odd[i, j]).sum() > 1


# ----------------------------------------------------------------------------- create_ex flags
def test_treeah_raw_by_position_and_borrowed(gpu_lib, oracle):
    import torch

    x, _ = helpers.clustered(30_000, 32, 48, 0.35, 3)
    q = (x[:200] + 0.01).astype(np.float32)
    idx = helpers.build_index(oracle, x, 24, 16)
    cfg = gpu_lib.TreeXHybridConfig(num_partitions=24, partitions_to_search=6)
    a = gpu_lib.TreeXHybridSearcher(cfg).build_from_index(idx["centers"], idx["codebook"], idx["packed"], idx["ids"],
                                                          idx["part_offsets"], x)
    ia, da, ca = a.search_batched(q, 10, pre_reorder_k=60)
    # the same index with raw rows in index-row order and GLOBAL ids shifted by 1000 (a shard's view)
    ids = idx["ids"].astype(np.int64)
    raw_pos = x[ids]
    dev = torch.device("cuda", 0)
    t = lambda v, dt: torch.as_tensor(v).to(dev, dt).contiguous()
    b = gpu_lib.TreeXHybridSearcher(cfg).build_from_index(
        t(idx["centers"], torch.float32), t(idx["codebook"], torch.float32), t(idx["packed"], torch.uint8),
        t(ids + 1000, torch.int32), t(idx["part_offsets"].astype(np.int64), torch.int64), t(raw_pos, torch.float32),
        raw_by_position=True, borrow_raw=True)
    ib, db_, cb_ = b.search_batched(t(q, torch.float32), 10, pre_reorder_k=60)
    torch.cuda.synchronize()
    assert (cb_.cpu().numpy() == ca).all()
    assert (db_.cpu().numpy().view(np.uint32) == da.view(np.uint32)).all()
    assert (ib.cpu().numpy().view(np.uint32) == ia + 1000).all()


def test_treeah_out_of_range_raw_rows_are_skipped(gpu_lib, oracle):
    """ids that point outside the raw array (malformed / sharded index) are dropped, as dataset.get(idx) does in the
    reference (tree_x_hybrid/mod.rs:352-356), instead of being read out of bounds."""
    x, _ = helpers.clustered(5_000, 16, 16, 0.35, 5)
    q = x[:50].copy()
    idx = helpers.build_index(oracle, x, 8, 8)
    cfg = gpu_lib.TreeXHybridConfig(num_partitions=8, partitions_to_search=8)
    s = gpu_lib.TreeXHybridSearcher(cfg).build_from_index(idx["centers"], idx["codebook"], idx["packed"], idx["ids"],
                                                          idx["part_offsets"], x[:2500])
    ids, dists, counts = s.search_batched(q, 10, pre_reorder_k=100)
    for i in range(len(q)):
        c = int(counts[i])
        assert (ids[i, :c] < 2500).all()
        assert (ids[i, c:] == 0xFFFFFFFF).all()


# ----------------------------------------------------------------------------- SQ8 calibration bit-exact (a4)
@pytest.mark.parametrize("n,dim,seed", [(10_000, 128, 1), (100_003, 96, 2), (777, 3, 3), (1_000_000, 128, 42)])
def test_sq8_quantize_bit_exact(gpu_lib, oracle, n, dim, seed):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((n, dim), dtype=np.float32) * np.float32(1.7) + np.float32(0.3))
    codes, cal = gpu_lib.scalar_quantize(x)
    ocodes, ocal = oracle.sq8_quantize(x)
    assert (np.asarray(cal).view(np.uint32) == np.asarray(ocal, np.float32).view(np.uint32)).all(), (cal, ocal)
    assert (codes == ocodes).all()
