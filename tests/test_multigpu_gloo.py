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


# ---------------------------------------------------------------------------------------------------------------
# two_phase_search plumbing (token slices all-gathered, bounds min-reduced) with a CPU stand-in for the GPU searcher
class _StubSearcher:
    """Implements the searcher protocol of distributed.two_phase_search on the CPU: tokens via the oracle's
    partitioner, bounds = rank-dependent floats, results = what the oracle finds on this rank's shard."""

    class config:
        partitions_to_search = 4

    def __init__(self, oracle, sh, x, rank):
        self.oracle, self.sh, self.x, self.rank = oracle, sh, x, rank
        self.seen = {}

    def partition_tokens(self, queries, L):
        tok, _ = self.oracle.partition(self.sh["centers"], queries.numpy(), L)
        return torch.from_numpy(tok.view(np.int32).copy())

    def search_begin(self, queries, k, partitions_to_search=None, pre_reorder_k=None, tokens=None):
        self.seen["tokens"] = tokens.numpy().copy()
        self.q, self.k, self.R = queries.numpy(), k, pre_reorder_k
        nq = queries.shape[0]
        return torch.arange(nq, dtype=torch.float32) + (100.0 if self.rank == 0 else -1.0) * (torch.arange(nq) % 2)

    def search_end(self, tau):
        self.seen["tau"] = tau.numpy().copy()
        _, ids, dists, counts = self.oracle.treex_search(self.sh["centers"], self.sh["codebook"], self.sh["part_offsets"],
                                                         self.sh["ids"], self.sh["packed"], self.x, self.q, 4, self.R, self.k)
        return torch.from_numpy(ids.view(np.int32)), torch.from_numpy(dists), torch.from_numpy(counts.view(np.int32))


def _worker2(rank, world, init_file, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    import oracle

    pkg = importlib.import_module("scann-rust_b200")
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    x, _ = helpers.clustered(6000, 32, 16, 0.3, 2)
    q = torch.from_numpy((x[:48] + 0.02).astype(np.float32))
    idx = helpers.build_index(oracle, x, 12, 8)
    stub = _StubSearcher(oracle, pkg.indexing.shard_index(idx, rank, world), x, rank)
    ids, dists, counts = pkg.distributed.two_phase_search(stub, q, 10, pre_reorder_k=40)
    np.savez(os.path.join(out_dir, f"tp{rank}.npz"), tokens=stub.seen["tokens"], tau=stub.seen["tau"], ids=ids.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_phase_plumbing_gloo():
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker2, args=(world, os.path.join(d, "init"), d), nprocs=world, join=True)
        r0, r1 = np.load(os.path.join(d, "tp0.npz")), np.load(os.path.join(d, "tp1.npz"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    import oracle
    x, _ = helpers.clustered(6000, 32, 16, 0.3, 2)
    idx = helpers.build_index(oracle, x, 12, 8)
    otok, _ = oracle.partition(idx["centers"], (x[:48] + 0.02).astype(np.float32), 4)
    # every rank saw the tokens of the WHOLE batch (its own slice + the gathered ones), in query order
    assert (r0["tokens"].view(np.uint32) == otok).all() and (r1["tokens"].view(np.uint32) == otok).all()
    # and the element-wise minimum of the two ranks' bounds
    n = np.arange(48, dtype=np.float32)
    want = np.minimum(n + 100.0 * (np.arange(48) % 2), n - 1.0 * (np.arange(48) % 2))
    assert (r0["tau"] == want).all() and (r1["tau"] == want).all()
