"""Index build for Tree-AH on the GPU (outside the timed search path).

Training (k-means for the partition centres and the per-subspace residual codebooks) is our own — the
reference's trainer draws from an unpinned `rand::StdRng` (src/utils/random.rs:20-46) so its centroids are
not reproducible anyway.  What IS the reference's semantics, and is done by the library's exact kernels:
  * point → partition assignment = TreePartitioner::partition(x, 1)     (tree_partitioner.rs:196-229)
  * residual = x - centre[assign], codes = Codebook::encode(residual)   (tree_x_hybrid/mod.rs:177-189,
                                                                          hashes/codebook.rs:82-95)
  * PackedCodes4Bit::from_codes packing                                  (hashes/lut16.rs:43-61)
torch is used for the training matmuls only.
"""
from __future__ import annotations

import numpy as np

from . import searchers


def _torch():
    import torch
    return torch


def kmeans(x, K: int, iters: int = 20, seed: int = 7, chunk: int = 262144):
    """Lloyd's k-means on a torch tensor [n, d]; returns centres [K, d] (f32).  Empty clusters are
    re-seeded from the points farthest from their centre."""
    torch = _torch()
    n, d = x.shape
    K = min(K, n)
    g = torch.Generator(device=x.device)
    g.manual_seed(seed)
    perm = torch.randperm(n, generator=g, device=x.device)[:K]
    centers = x[perm].clone().float()
    for _ in range(iters):
        sums = torch.zeros((K, d), dtype=torch.float32, device=x.device)
        cnts = torch.zeros((K,), dtype=torch.float32, device=x.device)
        cn = (centers * centers).sum(1)
        far_val, far_idx = None, None
        for s in range(0, n, chunk):
            xb = x[s:s + chunk].float()
            dist = cn[None, :] - 2.0 * (xb @ centers.t())
            md, a = dist.min(1)
            sums.index_add_(0, a, xb)
            cnts.index_add_(0, a, torch.ones_like(md))
            md = md + (xb * xb).sum(1)
            v, i = md.max(0)
            if far_val is None or v > far_val:
                far_val, far_idx = v, i + s
        nonempty = cnts > 0
        new_centers = torch.where(nonempty[:, None], sums / cnts.clamp(min=1.0)[:, None], centers)
        empty = (~nonempty).nonzero().flatten()
        if empty.numel() > 0:
            repl = torch.randint(0, n, (empty.numel(),), generator=g, device=x.device)
            new_centers[empty] = x[repl].float()
        centers = new_centers
    return centers.contiguous()


def train_codebook(residuals, S: int, num_codes: int = 16, iters: int = 20, seed: int = 42):
    """Per-subspace k-means (Codebook::train, hashes/codebook.rs:146-202: subspace s uses seed+s) on
    residuals [n, D] → codebook [S, num_codes, ds]."""
    torch = _torch()
    n, D = residuals.shape
    ds = D // S
    cb = torch.empty((S, num_codes, ds), dtype=torch.float32, device=residuals.device)
    for s in range(S):
        cb[s] = kmeans(residuals[:, s * ds:(s + 1) * ds].contiguous(), num_codes, iters, seed + s)
    return cb.contiguous()


def assign_partitions(x, centers, device: int = 0, chunk: int = 1 << 20):
    """assign[i] = TreePartitioner::partition(x[i], 1).tokens[0] with the reference's exact arithmetic."""
    torch = _torch()
    part = searchers.TreePartitioner(centers, device)
    out = torch.empty((x.shape[0],), dtype=torch.int32, device=x.device)
    for s in range(0, x.shape[0], chunk):
        tokens, _ = part.partition(x[s:s + chunk], 1)
        out[s:s + chunk] = tokens[:, 0]
    torch.cuda.synchronize()
    part.close()
    return out


def build_treeah_index(x, K: int, S: int, train_sample: int = 1_000_000, kmeans_iters: int = 20, seed: int = 7,
                       use_residuals: bool = True, device: int = 0, encode_chunk: int = 1 << 21):
    """x: torch CUDA tensor [N, D] f32.  Returns a dict of torch CUDA tensors:
    centers [K,D], codebook [S,16,ds], packed [N, ceil(S/2)] (grouped by partition), ids [N] (i32),
    part_offsets [K+1] (i64), assign [N]."""
    torch = _torch()
    N, D = x.shape
    assert D % S == 0
    g = torch.Generator(device=x.device)
    g.manual_seed(seed)
    ns = min(train_sample, N)
    sample = x[torch.randperm(N, generator=g, device=x.device)[:ns]] if ns < N else x
    centers = kmeans(sample, K, kmeans_iters, seed)
    K = centers.shape[0]
    assign = assign_partitions(x, centers, device)
    if use_residuals:
        a_s = assign_partitions(sample, centers, device) if ns < N else assign
        resid = sample - centers[a_s.long()]
    else:
        resid = sample
    codebook = train_codebook(resid, S, 16, kmeans_iters, 42)
    del resid
    order = torch.argsort(assign.long(), stable=True)
    counts = torch.bincount(assign.long(), minlength=K)
    part_offsets = torch.zeros((K + 1,), dtype=torch.int64, device=x.device)
    part_offsets[1:] = torch.cumsum(counts, 0)
    bpp = (S + 1) // 2
    packed = torch.empty((N, bpp), dtype=torch.uint8, device=x.device)
    for s in range(0, N, encode_chunk):
        idx = order[s:s + encode_chunk]
        xb = x[idx].contiguous()
        ab = assign[idx].contiguous()
        packed[s:s + encode_chunk] = searchers.pq_encode(codebook, xb, centers if use_residuals else None,
                                                         ab if use_residuals else None, device)
    torch.cuda.synchronize()
    return {
        "centers": centers, "codebook": codebook, "packed": packed, "ids": order.to(torch.int32).contiguous(),
        "part_offsets": part_offsets, "assign": assign,
    }


def shard_index(index: dict, rank: int, world: int):
    """Row-shards an index round-robin INSIDE each partition: rank r keeps members
    r, r+world, r+2*world, … of every leaf.  Works on torch tensors (any device) or numpy arrays."""
    off = index["part_offsets"]
    is_np = isinstance(off, np.ndarray)
    if is_np:
        off64 = off.astype(np.int64)
        n = int(off64[-1])
        pos = np.arange(n, dtype=np.int64)
        leaf = np.searchsorted(off64, pos, side="right") - 1
        keep = ((pos - off64[leaf]) % world) == rank
        counts = np.bincount(leaf[keep], minlength=len(off64) - 1)
        new_off = np.zeros(len(off64), np.uint64)
        new_off[1:] = np.cumsum(counts)
        return {**index, "packed": index["packed"][keep], "ids": index["ids"][keep], "part_offsets": new_off}
    torch = _torch()
    off64 = off.to(torch.int64)
    n = int(off64[-1].item())
    pos = torch.arange(n, dtype=torch.int64, device=off.device)
    leaf = torch.searchsorted(off64, pos, right=True) - 1
    keep = ((pos - off64[leaf]) % world) == rank
    counts = torch.bincount(leaf[keep], minlength=off64.numel() - 1)
    new_off = torch.zeros_like(off64)
    new_off[1:] = torch.cumsum(counts, 0)
    return {**index, "packed": index["packed"][keep].contiguous(), "ids": index["ids"][keep].contiguous(),
            "part_offsets": new_off}
