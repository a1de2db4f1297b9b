"""GPU tests of the two-phase (split) Tree-AH search for a sharded index: scann_treeah_search_begin scans every
query's closest leaf and returns the bounds it proves; after a min-reduction over the shards,
scann_treeah_search_end scans the rest under the global bounds.  Two shards are emulated on one GPU (sequentially —
no kernel waits on another)."""
import numpy as np
import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu


def _index(oracle, n=60_000, dim=32, K=48, S=16, seed=3):
    x, _ = helpers.clustered(n, dim, 64, 0.35, seed)
    return x, helpers.build_index(oracle, x, K, S)


def _searcher(pkg, idx, x, L, part_offsets=None, ids=None, packed=None):
    s = pkg.TreeXHybridSearcher(pkg.TreeXHybridConfig(num_partitions=len(idx["centers"]), partitions_to_search=L))
    s.build_from_index(idx["centers"], idx["codebook"], packed if packed is not None else idx["packed"],
                       ids if ids is not None else idx["ids"],
                       part_offsets if part_offsets is not None else idx["part_offsets"], x)
    return s


def test_split_search_single_shard_equals_plain_search(gpu_lib, oracle):
    x, idx = _index(oracle)
    q = torch.tensor(x[:500] + 0.01).cuda()
    s = _searcher(gpu_lib, idx, x, 8)
    ids, dists, counts = s.search_batched(q, 10, pre_reorder_k=60)
    tau = s.search_begin(q, 10, pre_reorder_k=60)
    assert tau.shape == (500,) and bool(torch.isfinite(tau).all())
    ids2, dists2, counts2 = s.search_end(tau)
    torch.cuda.synchronize()
    assert (ids.cpu() == ids2.cpu()).all() and (dists.cpu() == dists2.cpu()).all() and (counts.cpu() == counts2.cpu()).all()
    # no bounds at all (tau = None) is also the plain search
    s.search_begin(q, 10, pre_reorder_k=60)
    ids3, dists3, _ = s.search_end(None)
    assert (ids.cpu() == ids3.cpu()).all()
    # tau is a valid bound: at least R candidates of the full search lie at or below it
    _, _, _, (ci, cd, cc) = s.search_batched(q, 10, pre_reorder_k=60, want_candidates=True)
    cd, cc, tau = cd.cpu().numpy(), cc.cpu().numpy(), tau.cpu().numpy()
    for i in range(500):
        if cc[i] == 60:
            assert cd[i, 59] <= tau[i]


def test_split_search_with_gathered_tokens(gpu_lib, oracle):
    x, idx = _index(oracle)
    q = torch.tensor(x[2000:2512] + 0.01).cuda()
    s = _searcher(gpu_lib, idx, x, 8)
    ids, dists, counts = s.search_batched(q, 10, pre_reorder_k=60)
    # four "ranks" partition a quarter of the batch each; the concatenation is the all-gather
    tokens = torch.cat([s.partition_tokens(q[r * 128:(r + 1) * 128], 8) for r in range(4)])
    otok, _ = oracle.partition(idx["centers"], q.cpu().numpy(), 8)
    assert (tokens.cpu().numpy().view(np.uint32) == otok).all()
    tau = s.search_begin(q, 10, pre_reorder_k=60, tokens=tokens)
    ids2, dists2, counts2 = s.search_end(tau)
    torch.cuda.synchronize()
    assert (ids.cpu() == ids2.cpu()).all() and (dists.cpu() == dists2.cpu()).all()


def test_split_search_protocol_errors(gpu_lib, oracle):
    x, idx = _index(oracle, n=5000, K=8, S=8)
    s = _searcher(gpu_lib, idx, x, 4)
    q = torch.tensor(x[:16]).cuda()
    with pytest.raises(gpu_lib.ScannError) as e:
        gpu_lib.capi.check(gpu_lib.capi.load().scann_treeah_search_end(s._h, None, None, None, None, None))  # no begin
    assert e.value.code == gpu_lib.capi.FAILED_PRECONDITION
    s.search_begin(q, 5)
    with pytest.raises(gpu_lib.ScannError) as e:
        s.search_begin(q, 5)
    assert e.value.code == gpu_lib.capi.FAILED_PRECONDITION
    ids, dists, counts = s.search_end(None)
    assert ids.shape == (16, 5)
    with pytest.raises(gpu_lib.ScannError):
        s.search_begin(x[:16], 5)  # host queries are not accepted by the split path


@pytest.mark.parametrize("world", [2, 4])
def test_two_phase_sharded_search_is_deterministic_superset(gpu_lib, oracle, world):
    x, idx = _index(oracle)
    K = len(idx["centers"])
    L, R, k = 12, 80, 10
    qn = (x[1000:1600] + 0.01).astype(np.float32)
    q = torch.tensor(qn).cuda()
    full = _searcher(gpu_lib, idx, x, L)
    fi, fd, fc, (ci, cd, cc) = full.search_batched(q, k, pre_reorder_k=R, want_candidates=True)
    # whole partitions per shard (bench.py --shard partition)
    off = idx["part_offsets"].astype(np.int64)
    shards = []
    for r in range(world):
        keep = np.zeros(len(idx["ids"]), bool)
        cnt = np.zeros(K, np.int64)
        for leaf in range(K):
            if leaf % world == r:
                keep[off[leaf]:off[leaf + 1]] = True
                cnt[leaf] = off[leaf + 1] - off[leaf]
        po = np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint64)
        shards.append(_searcher(gpu_lib, idx, x, L, po, np.ascontiguousarray(idx["ids"][keep]),
                                np.ascontiguousarray(idx["packed"][keep])))

    def run():
        taus = [s.search_begin(q, k, pre_reorder_k=R) for s in shards]
        tau = torch.stack(taus).min(0).values  # = all_reduce(MIN)
        outs = [s.search_end(tau) for s in shards]
        gi = torch.stack([o[0] for o in outs])
        gd = torch.stack([o[1] for o in outs])
        mi, md, mc = gpu_lib.merge_topk(gi, gd)
        torch.cuda.synchronize()
        return mi.cpu().numpy().view(np.uint32), md.cpu().numpy(), tau.cpu().numpy()

    mi, md, tau = run()
    mi2, md2, _ = run()
    assert (mi == mi2).all() and (md.view(np.uint32) == md2.view(np.uint32)).all()  # deterministic
    # the global bound is valid: the R-th best approximate distance of the unsharded search is <= tau
    cdn, ccn = cd.cpu().numpy(), cc.cpu().numpy()
    assert all(cdn[i, R - 1] <= tau[i] for i in range(len(qn)) if ccn[i] == R)
    # superset: every exact distance of the merged sharded result is <= the unsharded one at the same rank
    fdn = fd.cpu().numpy()
    assert (md <= fdn + 0.0).all()
    # and the plain (unsplit) sharded search gives the same merged result
    outs = [s.search_batched(q, k, pre_reorder_k=R) for s in shards]
    pi, pd, _ = gpu_lib.merge_topk(torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs]))
    assert (pd.cpu().numpy() <= fdn).all()
    rc, gt, _, _ = oracle.bf_search(x, qn, k, oracle.SQL2, nthreads=8)
    assert helpers.recall(mi, gt, k) >= helpers.recall(fi.cpu().numpy().view(np.uint32), gt, k) - 1e-9


def test_split_search_abort_and_cross_thread_end(gpu_lib, oracle):
    """Between begin and end the handle is busy, not locked: a plain search is refused, abort gives the handle back,
    and end may run on another thread than begin (a held std::mutex would make that undefined behaviour)."""
    import threading

    x, idx = _index(oracle, n=5000, K=8, S=8)
    s = _searcher(gpu_lib, idx, x, 4)
    q = torch.tensor(x[:16]).cuda()
    ref = s.search_batched(q, 5)
    torch.cuda.synchronize()
    s.search_begin(q, 5)
    with pytest.raises(gpu_lib.ScannError) as e:
        s.search_batched(q, 5)
    assert e.value.code == gpu_lib.capi.FAILED_PRECONDITION
    s.search_abort()
    again = s.search_batched(q, 5)          # usable again
    torch.cuda.synchronize()
    assert (again[0] == ref[0]).all()
    tau = s.search_begin(q, 5)
    out = {}

    def finish():
        torch.cuda.set_device(0)
        out["r"] = s.search_end(tau)
        torch.cuda.synchronize()

    t = threading.Thread(target=finish)
    t.start()
    t.join()
    assert (out["r"][0] == ref[0]).all() and (out["r"][1] == ref[1]).all()


def test_device_calls_on_two_streams_are_ordered(gpu_lib, oracle):
    """SCANN_DEVICE calls return with their kernels enqueued; the handle's workspace is shared, so a call on another
    stream must start after them (StreamOrder event).  Alternate two streams without any host synchronisation."""
    x, idx = _index(oracle, n=20000, K=16, S=8)
    s = _searcher(gpu_lib, idx, x, 6)
    qa = torch.tensor(x[:512]).cuda()
    qb = torch.tensor(x[512:1024]).cuda()
    ra = s.search_batched(qa, 10)
    rb = s.search_batched(qb, 10)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for it in range(6):
        with torch.cuda.stream(s1 if it % 2 == 0 else s2):
            outs.append(s.search_batched(qa if it % 2 == 0 else qb, 10))
    torch.cuda.synchronize()
    for it, o in enumerate(outs):
        want = ra if it % 2 == 0 else rb
        assert (o[0] == want[0]).all() and (o[1] == want[1]).all()


def test_token_prefetcher_pipeline_gives_identical_results(gpu_lib, oracle, tmp_path):
    """distributed.TokenPrefetcher: the partition slice + token all-gather of the next batch run on a side stream; the
    results of a stream of batches must equal the un-pipelined two_phase_search (world of one NCCL rank)."""
    import torch.distributed as dist

    x, idx = _index(oracle, n=20000, K=16, S=8)
    s = _searcher(gpu_lib, idx, x, 6)
    batches = [torch.tensor(x[i * 256:(i + 1) * 256] + 0.01).cuda() for i in range(4)]
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", init_method=f"file://{tmp_path}/pg", rank=0, world_size=1)
    try:
        D = gpu_lib.distributed
        plain = [D.two_phase_search(s, b, 10, pre_reorder_k=60) for b in batches]
        torch.cuda.synchronize()
        pf = D.TokenPrefetcher(s, idx["centers"], 6)
        pf.prefetch(batches[0])
        piped = []
        for i, b in enumerate(batches):
            piped.append(D.two_phase_search(s, b, 10, pre_reorder_k=60, prefetcher=pf,
                                            next_queries=batches[i + 1] if i + 1 < len(batches) else None))
        torch.cuda.synchronize()
        for a, b in zip(plain, piped):
            assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and (a[2] == b[2]).all()
        # a batch that was not prefetched falls back to the inline partition
        other = D.two_phase_search(s, batches[2], 10, pre_reorder_k=60, prefetcher=pf)
        torch.cuda.synchronize()
        assert (other[0] == plain[2][0]).all()
        pf.close()
    finally:
        if created:
            dist.destroy_process_group()
