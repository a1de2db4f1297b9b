"""Multi-GPU plumbing: one process per GPU, the index sharded (whole partitions per rank in bench.py;
``indexing.shard_index`` implements the rows-round-robin-inside-partitions alternative), every rank searches the whole
query batch on its shard and returns its local top-k (exact distances, global ids).

  TokenPrefetcher      partition slice + token all-gather of the next batch on a side stream (software pipeline)
  two_phase_search     the Tree-AH step on a sharded index: token slices all-gathered, closest-leaf bounds all-reduced
                       (MIN), scan under the global bounds (include/scann_b200.h scann_treeah_search_begin/_end)
  exchange_and_merge   ONE all-gather of the packed [2, nq, k] result buffer (80 B per query per rank at k = 10) and a
                       k-way merge kernel ordered by (distance, id)
  sharded_search       both, for any searcher (falls back to a plain local search + exchange)

torch.distributed is only the transport (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations


def all_gather_results(ids, dists, group=None):
    """ids/dists: [nq, k] torch tensors on this rank → ([world, nq, k], [world, nq, k])."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    nq, k = int(ids.shape[0]), int(ids.shape[1])
    # concatenation along dim 0 is the layout every backend (nccl, gloo) accepts; viewed as [world, nq, k]
    gi = torch.empty((world * nq, k), dtype=ids.dtype, device=ids.device)
    gd = torch.empty((world * nq, k), dtype=dists.dtype, device=dists.device)
    dist.all_gather_into_tensor(gi, ids.contiguous(), group=group)
    dist.all_gather_into_tensor(gd, dists.contiguous(), group=group)
    return gi.view(world, nq, k), gd.view(world, nq, k)


_PHASE_MARKS = None  # set to a list to collect CUDA events: start, tokens, gathered, begin, reduced, end (per step)


def exchange_and_merge(ids, dists, group=None):
    """This rank's [nq, k] results → merged (ids, dists, counts) on every rank.  When ids and dists are the two
    halves of one packed [2, nq, k] buffer (what the searchers return for CUDA queries) the exchange is ONE all-gather
    and the merge reads the gathered buffer in place; otherwise two all-gathers + the plain merge."""
    import torch
    import torch.distributed as dist

    from . import searchers

    world = dist.get_world_size(group)
    nq, k = int(ids.shape[0]), int(ids.shape[1])
    packed = None
    if ids.is_cuda and ids.dtype == torch.int32 and dists.dtype == torch.float32 and ids.is_contiguous() \
            and dists.is_contiguous() and ids.untyped_storage().data_ptr() == dists.untyped_storage().data_ptr() \
            and ids.storage_offset() == 0 and dists.storage_offset() == nq * k:
        packed = torch.empty(0, dtype=torch.int32, device=ids.device).set_(ids.untyped_storage(), 0, (2 * nq * k,))
    if packed is None:
        gi, gd = all_gather_results(ids, dists, group)
        return searchers.merge_topk(gi, gd, ids.device.index or 0)
    gathered = torch.empty((world * 2 * nq * k,), dtype=torch.int32, device=ids.device)
    dist.all_gather_into_tensor(gathered, packed, group=group)
    return searchers.merge_topk_packed(gathered.view(world, 2, nq, k), ids.device.index or 0)


class TokenPrefetcher:
    """Software pipeline for the sharded step: the partition slice + token all-gather of the NEXT batch run on a side
    stream (own TreePartitioner handle, hence own workspace) while the current batch is inside search_end, so the two
    latency-bound phases (~0.13 ms at 8 GPUs) leave the critical path of a stream of batches.

        pf = TokenPrefetcher(searcher, centers, L)
        for i, q in enumerate(batches):
            ids, dists, cnt = two_phase_search(searcher, q, k, prefetcher=pf,
                                               next_queries=batches[i + 1] if i + 1 < len(batches) else None)

    Every rank must issue the same sequence of calls (the collectives are matched by order).  Tokens are bit-identical
    to the inline path: both run TreePartitioner::partition over the same centres (scann_part_select /
    scann_treeah_partition share their kernels)."""

    def __init__(self, searcher, centers, partitions_to_search=None, group=None):
        import torch

        from . import searchers

        self.group = group
        self.L = int(partitions_to_search if partitions_to_search is not None else searcher.config.partitions_to_search)
        self.part = searchers.TreePartitioner(centers, searcher.device)
        self.stream = torch.cuda.Stream(device=searcher.device)
        self._slot = None  # (key, tokens, event, keepalive)

    @staticmethod
    def _key(queries):
        return (queries.data_ptr(), tuple(queries.shape))

    def prefetch(self, queries):
        """enqueue partition slice + all-gather for `queries` (torch CUDA [nq, dim], nq divisible by the world size)"""
        import torch
        import torch.distributed as dist

        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        nq = int(queries.shape[0])
        per = nq // world
        if per * world != nq:
            self._slot = None
            return
        self.stream.wait_stream(torch.cuda.current_stream(queries.device))  # the batch may still be being produced
        with torch.cuda.stream(self.stream):
            mine, _ = self.part.partition(queries[rank * per:(rank + 1) * per], self.L)
            tokens = torch.empty((nq, self.L), dtype=torch.int32, device=queries.device)
            dist.all_gather_into_tensor(tokens, mine, group=self.group)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._slot = (self._key(queries), tokens, ev, (queries, mine))

    def take(self, queries):
        """the prefetched tokens of exactly this batch (the current stream then waits for them), else None"""
        import torch

        slot, self._slot = self._slot, None
        if slot is None or slot[0] != self._key(queries):
            return None
        cur = torch.cuda.current_stream(queries.device)
        cur.wait_event(slot[2])
        slot[1].record_stream(cur)  # allocated on the side stream, consumed on this one
        return slot[1]

    def close(self):
        self._slot = None
        self.part.close()


def two_phase_search(searcher, queries, k: int, group=None, partitions_to_search=None, pre_reorder_k=None,
                     prefetcher=None, next_queries=None):
    """Tree-AH on a sharded index, one step on this rank (queries: the whole batch, torch CUDA [nq, dim]):
      1. partition nq/world queries here, all-gather the tokens (the stage is not repeated on every GPU);
      2. search_begin: probe every query's closest leaf on this shard → bounds; all-reduce MIN over the shards;
      3. search_end: scan under the global bounds, merge, exact reorder → this shard's top-k."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nq = int(queries.shape[0])
    marks = _PHASE_MARKS  # optional CUDA-event marks between the phases (bench.py --phase-times)

    def mark():
        if marks is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append(e)

    mark()
    tokens = prefetcher.take(queries) if prefetcher is not None else None
    per = (nq + world - 1) // world
    if tokens is not None:
        mark()  # keeps the phase table aligned: partition slice and all-gather cost nothing on this stream
    elif world > 1 and per * world == nq:  # equal slices only (all_gather_into_tensor); otherwise partition locally
        L = int(partitions_to_search if partitions_to_search is not None else searcher.config.partitions_to_search)
        mine = searcher.partition_tokens(queries[rank * per:(rank + 1) * per], L)
        mark()
        tokens = torch.empty((nq, L), dtype=torch.int32, device=queries.device)
        dist.all_gather_into_tensor(tokens, mine, group=group)
    mark()
    tau = searcher.search_begin(queries, k, partitions_to_search, pre_reorder_k, tokens=tokens)
    mark()
    try:
        dist.all_reduce(tau, op=dist.ReduceOp.MIN, group=group)
    except Exception:
        if hasattr(searcher, "search_abort"):
            searcher.search_abort()  # the handle is busy between begin and end: give it back
        raise
    mark()
    out = searcher.search_end(tau)
    mark()
    if prefetcher is not None and next_queries is not None:
        prefetcher.prefetch(next_queries)  # runs beside search_end's kernels; issued after this batch's all-reduce
    return out


def sharded_search(searcher, queries, k: int, group=None, **kw):
    """Local search on this rank's shard + all-gather + GPU merge → (ids, dists, counts) on every rank."""
    import torch.distributed as dist

    from . import searchers

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return searcher.search_batched(queries, k, **kw)
    if hasattr(searcher, "search_begin") and getattr(queries, "is_cuda", False):
        # Tree-AH on a sharded index: every shard scans the queries' closest leaves first, the bounds they prove are
        # min-reduced over the shards (4 B per query), and the bulk of the scan runs under the global bounds
        ids, dists, counts = two_phase_search(searcher, queries, k, group, **kw)
    else:
        ids, dists, counts = searcher.search_batched(queries, k, **kw)
    return exchange_and_merge(ids, dists, group)
