"""Index build for Tree-AH on the GPU (outside the timed search path).

Training (k-means for the partition centres and the per-subspace residual codebooks) is our own — the
reference's trainer draws from an unpinned `rand::StdRng` (src/utils/random.rs:20-46) so its centroids are
not reproducible anyway.  What IS the reference's semantics, and is done by the library's exact kernels:
  * point → partition assignment = TreePartitioner::partition(x, 1)     (tree_partitioner.rs:196-229)
  * residual = x - centre[assign], codes = Codebook::encode(residual)   (tree_x_hybrid/mod.rs:177-189,
                                                                          hashes/codebook.rs:82-95)
  * PackedCodes4Bit::from_codes packing                                  (hashes/lut16.rs:43-61)
Training runs in the library too (scann_kmeans_fit / scann_pq_train, csrc/build_index.cu); torch only holds the device
buffers.
"""
from __future__ import annotations

import numpy as np

from . import capi, searchers


def _torch():
    import torch
    return torch


def kmeans(x, K: int, iters: int = 20, seed: int = 7, chunk: int = 0, balance_ratio: float = 0.0):
    """KMeans::fit (src/trees/kmeans.rs:166-432) on the GPU through the C ABI (scann_kmeans_fit, csrc/build_index.cu):
    Lloyd's algorithm with the library's exact nearest-centre assignment (tensor-core scoring + exact re-score) and f64
    cluster sums.  x: torch CUDA tensor [n, d] f32; returns centres [min(K, n), d].  `chunk` is ignored (kept for
    callers of the old torch trainer); balance_ratio > 1 splits clusters heavier than that multiple of n/K."""
    torch = _torch()
    x = x.float().contiguous()
    n, d = x.shape
    K = min(K, n)
    centers = torch.empty((K, d), dtype=torch.float32, device=x.device)
    dev = x.device.index or 0
    capi.require_gpu()
    capi.check(capi.load().scann_kmeans_fit(x.data_ptr(), n, d, d, K, iters, seed, float(balance_ratio), centers.data_ptr(),
                                            dev, capi.DEVICE))
    return centers


def train_codebook(residuals, S: int, num_codes: int = 16, iters: int = 20, seed: int = 42):
    """Codebook::train (hashes/codebook.rs:146-202: subspace s uses seed+s) on residuals [n, D] → codebook
    [S, 16, ds], through the C ABI (scann_pq_train)."""
    torch = _torch()
    if num_codes != 16:
        raise capi.ScannError(capi.INVALID_ARGUMENT, "the GPU trainer builds 16-code (LUT16) codebooks")
    r = residuals.float().contiguous()
    n, D = r.shape
    cb = torch.empty((S, 16, D // S), dtype=torch.float32, device=r.device)
    dev = r.device.index or 0
    capi.require_gpu()
    capi.check(capi.load().scann_pq_train(r.data_ptr(), n, D, D, None, None, 0, S, iters, seed, cb.data_ptr(), dev,
                                          capi.DEVICE))
    return cb


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
