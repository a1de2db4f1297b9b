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
