"""Test-side helpers: seeded synthetic data and a small CPU index builder.  The builder trains with
numpy k-means (our own: the reference's trainer RNG is unpinned) and uses the ORACLE for the parts that
have reference semantics (assignment, residual encode, nibble packing), so the same index arrays can be
fed to both the oracle and the GPU library."""
from __future__ import annotations

import numpy as np


def gaussian(n, dim, seed):
    return np.random.default_rng(seed).standard_normal((n, dim), dtype=np.float32)


def clustered(n, dim, n_latent, spread, seed, normalize=True):
    """SURVEY §8d C3-style mixture: latent centres ~N(0,1), point = centre + spread*N(0,1), L2-normalised."""
    rng = np.random.default_rng(seed)
    lat = rng.standard_normal((n_latent, dim), dtype=np.float32)
    which = rng.integers(0, n_latent, n)
    x = lat[which] + np.float32(spread) * rng.standard_normal((n, dim), dtype=np.float32)
    if normalize:
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32), lat


def np_kmeans(x, K, iters=10, seed=7):
    rng = np.random.default_rng(seed)
    K = min(K, x.shape[0])
    centers = x[rng.choice(x.shape[0], K, replace=False)].astype(np.float32).copy()
    for _ in range(iters):
        d = (x * x).sum(1)[:, None] - 2.0 * x @ centers.T + (centers * centers).sum(1)[None, :]
        a = d.argmin(1)
        for c in range(K):
            m = a == c
            if m.any():
                centers[c] = x[m].mean(0)
    return centers.astype(np.float32)


def build_index(oracle, x, K, S, seed=7, use_residuals=True, iters=8):
    """→ dict(centers [K,D], codebook [S,16,ds], codes [n,S] u8 and packed [n,ceil(S/2)] grouped by
    partition, ids [n] u32, part_offsets [K+1] u64, assign [n])"""
    n, D = x.shape
    ds = D // S
    centers = np_kmeans(x, K, iters, seed)
    K = centers.shape[0]
    assign = oracle.partition(centers, x, 1)[0][:, 0].astype(np.uint32)
    resid = x - centers[assign] if use_residuals else x
    cb = np.stack([np_kmeans(np.ascontiguousarray(resid[:, s * ds:(s + 1) * ds]), 16, iters, 42 + s)
                   for s in range(S)]).astype(np.float32)
    if cb.shape[1] < 16:  # tiny datasets: pad the codebook with far-away centroids
        pad = np.full((S, 16 - cb.shape[1], ds), 1e6, np.float32)
        cb = np.concatenate([cb, pad], 1)
    order = np.argsort(assign, kind="stable").astype(np.uint32)
    counts = np.bincount(assign, minlength=K)
    part_offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
    if use_residuals:
        codes = oracle.pq_encode_residual(cb, x[order], centers, assign[order])
    else:
        codes = oracle.pq_encode(cb, x[order])
    packed = oracle.pack4(codes)
    return dict(centers=centers, codebook=cb, codes=codes, packed=packed, ids=order, part_offsets=part_offsets,
                assign=assign)


def recall(ids_a, ids_b, k):
    """mean |a ∩ b| / k over queries (ann_benchmark.rs:452-471 average_recall_at_k)."""
    hit = 0
    for a, b in zip(ids_a, ids_b):
        hit += len(set(int(v) for v in a[:k]) & set(int(v) for v in b[:k]))
    return hit / (len(ids_a) * k)


def ids_equal_away_from_ties(ids_a, dists_a, ids_b, dists_b, counts, rel_gap=1e-6):
    """Neighbour ids must match wherever the oracle's distance is separated from its neighbours in the list
    (BASELINE: 'neighbour ids away from exact ties').  Returns (#compared, #mismatches)."""
    compared = mism = 0
    for q in range(ids_a.shape[0]):
        c = int(counts[q])
        for j in range(c):
            d = dists_b[q, j]
            tol = rel_gap * max(1.0, abs(float(d)))
            tied = (j > 0 and abs(dists_b[q, j - 1] - d) <= tol) or (j + 1 < c and abs(dists_b[q, j + 1] - d) <= tol)
            if j == c - 1:
                tied = True  # the cut-off position can tie with the first excluded element
            if tied:
                continue
            compared += 1
            if ids_a[q, j] != ids_b[q, j]:
                mism += 1
    return compared, mism
