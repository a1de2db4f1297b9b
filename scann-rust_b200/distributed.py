"""Multi-GPU plumbing: one process per GPU, the index sharded (whole partitions per rank in bench.py;
``indexing.shard_index`` implements the rows-round-robin-inside-partitions alternative), every rank searches the whole
query batch on its shard and returns its local top-k (exact distances, global ids).

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


def two_phase_search(searcher, queries, k: int, group=None, partitions_to_search=None, pre_reorder_k=None):
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
    tokens = None
    per = (nq + world - 1) // world
    if world > 1 and per * world == nq:  # equal slices only (all_gather_into_tensor); otherwise partition locally
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
