"""N > 1 host logic on CPU: world_size-2 gloo ranks shard one index round-robin inside the partitions
(indexing.shard_index), search their shard (the CPU oracle stands in for the per-rank GPU searcher — this test
covers sharding, the all-gather layout and the merge order, not the kernels), all-gather (id, distance) pairs
through scann-rust_b200.distributed and merge by (distance, id) — the contract of scann_merge_topk."""
import importlib
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _merge_reference(gi, gd, k):
    """numpy statement of scann_merge_topk: [parts][nq][k] → [nq][k] by (distance, id), padding ignored."""
    parts, nq, _ = gi.shape
    oi = np.full((nq, k), 0xFFFFFFFF, np.uint32)
    od = np.full((nq, k), np.inf, np.float32)
    for q in range(nq):
        cand = sorted((float(gd[p, q, j]), int(gi[p, q, j])) for p in range(parts) for j in range(k)
                      if gi[p, q, j] != 0xFFFFFFFF)
        for j, (d, i) in enumerate(cand[:k]):
            oi[q, j], od[q, j] = i, d
    return oi, od


def _worker(rank, world, init_file, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    import oracle

    pkg = importlib.import_module("scann-rust_b200")
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    x, _ = helpers.clustered(6000, 32, 16, 0.3, 2)
    q = (x[:48] + 0.02).astype(np.float32)
    idx = helpers.build_index(oracle, x, 12, 8)
    L, R, k = 4, 40, 10
    sh = pkg.indexing.shard_index(idx, rank, world)
    _, ids, dists, _ = oracle.treex_search(sh["centers"], sh["codebook"], sh["part_offsets"], sh["ids"], sh["packed"], x, q,
                                           L, R, k)
    gi, gd = pkg.distributed.all_gather_results(torch.from_numpy(ids.view(np.int32)), torch.from_numpy(dists))
    oi, od = _merge_reference(gi.numpy().view(np.uint32), gd.numpy(), k)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ids=oi, dists=od, local_ids=ids, local_dists=dists)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_search_and_merge():
    world = 2
    with tempfile.TemporaryDirectory() as d:
        init_file = os.path.join(d, "init")
        mp.spawn(_worker, args=(world, init_file, d), nprocs=world, join=True)
        r0, r1 = np.load(os.path.join(d, "rank0.npz")), np.load(os.path.join(d, "rank1.npz"))
    # every rank ends with the same merged result
    assert (r0["ids"] == r1["ids"]).all() and (r0["dists"] == r1["dists"]).all()
    # the merge is the (distance, id)-ordered top-k of the union of the local lists
    gi = np.stack([r0["local_ids"], r1["local_ids"]])
    gd = np.stack([r0["local_dists"], r1["local_dists"]])
    oi, od = _merge_reference(gi, gd, 10)
    assert (oi == r0["ids"]).all()
    assert (np.diff(r0["dists"], axis=1) >= 0).all()
    # sharded result is at least as good as the unsharded one (SURVEY §8e caveat)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    import oracle
    x, _ = helpers.clustered(6000, 32, 16, 0.3, 2)
    q = (x[:48] + 0.02).astype(np.float32)
    idx = helpers.build_index(oracle, x, 12, 8)
    _, ids1, d1, _ = oracle.treex_search(idx["centers"], idx["codebook"], idx["part_offsets"], idx["ids"], idx["packed"], x,
                                         q, 4, 40, 10)
    assert (r0["dists"] <= d1 + 1e-7).all()
