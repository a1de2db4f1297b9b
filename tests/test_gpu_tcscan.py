"""Tensor-core LUT16 scan (csrc/tcscan.cu, tcgen05.mma kind::i8 over one-hot expanded codes) against the CPU oracle.
The scan is forced on with SCANN_SCAN_TC=1 and held to the same stage-wise, tie-proof checks as the register-LUT kernel
(test_gpu_parity._treeah_case): the R approximate candidates (LUT16 u32 sums dequantised) must be bit-identical to the
oracle's, the candidate sets equal below the cut-off tie, and the exact reorder of them bit-identical."""
import os

import numpy as np
import pytest

import helpers
from test_gpu_parity import _check_stages, _treeah_case

pytestmark = pytest.mark.gpu


@pytest.fixture()
def tc_on():
    old = os.environ.get("SCANN_SCAN_TC")
    os.environ["SCANN_SCAN_TC"] = "1"
    yield
    if old is None:
        os.environ.pop("SCANN_SCAN_TC", None)
    else:
        os.environ["SCANN_SCAN_TC"] = old


@pytest.mark.parametrize("n,dim,K,S,nq,L,R", [
    (30_000, 32, 12, 16, 300, 4, 50),      # S = 16: two table atoms
    (120_000, 96, 60, 48, 700, 16, 100),   # C3 geometry scaled down (six table atoms + the threshold atom)
    (150_000, 128, 300, 64, 500, 24, 100), # C4 geometry scaled down (eight atoms: the A ring wraps inside a tile)
    (60_000, 64, 8, 32, 1000, 8, 100),     # every query probes every leaf: 1000 pairs per leaf = 8 query groups
    (20_000, 96, 40, 48, 37, 6, 20),       # few queries: groups of < 16 pairs (N = 16 MMAs)
])
def test_tc_scan_matches_oracle(gpu_lib, oracle, tc_on, n, dim, K, S, nq, L, R):
    s = _treeah_case(gpu_lib, oracle, n=n, dim=dim, K=K, S=S, nq=nq, L=L, R=R, k=10,
                     measure=gpu_lib.DistanceMeasure.SquaredL2, seed=11, min_recall=0.99)
    tc, lut = s.path_stats()
    assert tc >= 1 and lut == 0, (tc, lut)


def test_tc_scan_same_results_as_register_lut_kernel(gpu_lib, oracle):
    """Same index, same batch: SCANN_SCAN_TC=1 and =0 return identical ids, distances and candidate lists."""
    x, _ = helpers.clustered(200_000, 96, 64, 0.35, 21)
    q, _ = helpers.clustered(2000, 96, 64, 0.35, 21)
    q = (q + 0.03 * helpers.gaussian(2000, 96, 22)).astype(np.float32)
    idx = helpers.build_index(oracle, x, 50, 48)
    cfg = gpu_lib.TreeXHybridConfig(num_partitions=50, partitions_to_search=12,
                                    distance_measure=gpu_lib.DistanceMeasure.DotProduct)
    s = gpu_lib.TreeXHybridSearcher(cfg).build_from_index(idx["centers"], idx["codebook"], idx["packed"], idx["ids"],
                                                          idx["part_offsets"], x)
    out = {}
    for mode in ("0", "1"):
        os.environ["SCANN_SCAN_TC"] = mode
        out[mode] = s.search_batched(q, 10, pre_reorder_k=100, want_candidates=True)
    os.environ.pop("SCANN_SCAN_TC", None)
    tc, lut = s.path_stats()
    assert tc == 1 and lut == 1
    (i0, d0, c0, (ci0, cd0, cc0)), (i1, d1, c1, (ci1, cd1, cc1)) = out["0"], out["1"]
    assert (c0 == c1).all() and (cc0 == cc1).all()
    assert (cd0.view(np.uint32) == cd1.view(np.uint32)).all() and (ci0 == ci1).all()
    assert (d0.view(np.uint32) == d1.view(np.uint32)).all() and (i0 == i1).all()


def test_tc_scan_flagged_queries_fall_back(gpu_lib, oracle, tc_on):
    """Queries without a bound (closest leaf smaller than R) and queries whose candidate list overflows (massively
    duplicated points below the bound) are re-done by the register-LUT kernel: results still match the oracle."""
    rng = np.random.default_rng(5)
    dim, S, K = 32, 16, 6
    base = rng.normal(0, 1, (K, dim)).astype(np.float32) * 4
    sizes = [3000, 40, 2500, 12, 3000, 2000]  # two leaves smaller than R = 60
    x = np.concatenate([base[c] + 0.3 * rng.normal(0, 1, (m, dim)).astype(np.float32) for c, m in enumerate(sizes)])
    x[:2000] = x[0]  # 2000 identical points: every one of them is below the bound of a query near them
    idx = helpers.build_index(oracle, x.astype(np.float32), K, S)
    x = x.astype(np.float32)
    q = np.concatenate([x[:40] + 0.01, base[1][None] + 0.05 * rng.normal(0, 1, (30, dim)),
                        base[3][None] + 0.05 * rng.normal(0, 1, (30, dim)),
                        x[5000:5100] + 0.02]).astype(np.float32)
    L, R, k = 3, 60, 10
    rc, *ora = oracle.treex_search(idx["centers"], idx["codebook"], idx["part_offsets"], idx["ids"], idx["packed"], x, q,
                                   L, R, k, lut16=True, nthreads=8, want_candidates=True)
    cfg = gpu_lib.TreeXHybridConfig(num_partitions=K, partitions_to_search=L)
    s = gpu_lib.TreeXHybridSearcher(cfg).build_from_index(idx["centers"], idx["codebook"], idx["packed"], idx["ids"],
                                                          idx["part_offsets"], x)
    ids, dists, counts, (ci, cd, cc) = s.search_batched(q, k, pre_reorder_k=R, want_candidates=True)
    assert s.path_stats()[0] == 1
    _check_stages(oracle, x, oracle.SQL2, q, k, (ids, dists, counts, ci, cd, cc), tuple(ora))
