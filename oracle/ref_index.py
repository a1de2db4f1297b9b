"""ORACLE SIDE — TEST / BENCH INFRASTRUCTURE ONLY.

Index build for ``bench.py --impl reference`` with plain torch ops, so that the reference arm never builds, loads or
calls the product library (libscann_b200.so).  Same algorithms as scann-rust_b200/indexing.py uses for the GPU arm
(Lloyd k-means for the partition centres and the per-subspace residual codebooks; assignment = nearest centre by
squared L2, lower id on ties; Codebook::encode = per-subspace nearest codeword, lower code on ties,
hashes/codebook.rs:82-95; PackedCodes4Bit::from_codes = low nibble for the even subspace, hashes/lut16.rs:43-61),
written with torch matmul / argmin instead of the exact CUDA kernels.  Ties and last-bit rounding can differ from the
exact kernels; the arm only TIMES the CPU search on the index built here.
"""
from __future__ import annotations


def _chunk(K):
    return max(4096, min(262144, (1 << 31) // (4 * max(K, 1))))


def _lloyd(torch, x, K, iters, seed, chunk=None):
    n, d = x.shape
    K = min(K, n)
    chunk = chunk or _chunk(K)
    g = torch.Generator(device=x.device)
    g.manual_seed(seed)
    centers = x[torch.randperm(n, generator=g, device=x.device)[:K]].clone().float()
    for _ in range(iters):
        sums = torch.zeros((K, d), dtype=torch.float32, device=x.device)
        cnts = torch.zeros((K,), dtype=torch.float32, device=x.device)
        cn = (centers * centers).sum(1)
        for s in range(0, n, chunk):
            xb = x[s:s + chunk].float()
            a = (cn[None, :] - 2.0 * (xb @ centers.t())).argmin(1)
            sums.index_add_(0, a, xb)
            cnts.index_add_(0, a, torch.ones((xb.shape[0],), dtype=torch.float32, device=x.device))
        nonempty = cnts > 0
        new_centers = torch.where(nonempty[:, None], sums / cnts.clamp(min=1.0)[:, None], centers)
        empty = (~nonempty).nonzero().flatten()
        if empty.numel() > 0:
            new_centers[empty] = x[torch.randint(0, n, (empty.numel(),), generator=g, device=x.device)].float()
        centers = new_centers
    return centers.contiguous()


def kmeans(torch, x, K, iters=20, seed=7, chunk=None, balance_ratio=0.0):
    """balance_ratio > 1: the same scheme as the product trainer (csrc/build_index.cu kmeans_device) — Lloyd places
    K - K/8 centres, then the heaviest cluster is bisected (2-means on its own rows) until K centres exist."""
    import heapq

    n = x.shape[0]
    K = min(K, n)
    reserve = K // 8
    if not balance_ratio > 1.0 or reserve == 0 or n // K < 4:
        return _lloyd(torch, x, K, iters, seed, chunk)
    K0 = K - reserve
    c = _lloyd(torch, x, K0, iters, seed, chunk)
    a = assign(torch, x, c, chunk)
    order = torch.argsort(a, stable=True)
    bounds = torch.cumsum(torch.bincount(a, minlength=K0), 0).tolist()
    members = [order[(bounds[i - 1] if i else 0):bounds[i]] for i in range(K0)]
    cs = [c[i] for i in range(K0)]
    heap = [(-int(m.numel()), i) for i, m in enumerate(members)]
    heapq.heapify(heap)
    while len(cs) < K and heap and heap[0][0] <= -2:
        _, i = heapq.heappop(heap)
        rows = x[members[i]].float()
        sub = _lloyd(torch, rows, 2, 6, seed + 7919 * len(cs))
        side = assign(torch, rows, sub)
        m0, m1 = members[i][side == 0], members[i][side == 1]
        if m0.numel() == 0 or m1.numel() == 0:
            heapq.heappush(heap, (0, i))
            continue
        cs[i], members[i] = sub[0], m0
        cs.append(sub[1])
        members.append(m1)
        heapq.heappush(heap, (-int(m0.numel()), i))
        heapq.heappush(heap, (-int(m1.numel()), len(cs) - 1))
    while len(cs) < K:
        cs.append(x[len(cs) % n].float())
    return torch.stack(cs).contiguous()


def train_codebook(torch, residuals, S, num_codes=16, iters=20, seed=42):
    n, D = residuals.shape
    ds = D // S
    cb = torch.empty((S, num_codes, ds), dtype=torch.float32, device=residuals.device)
    for s in range(S):
        cb[s] = kmeans(torch, residuals[:, s * ds:(s + 1) * ds].contiguous(), num_codes, iters, seed + s)
    return cb.contiguous()


def assign(torch, x, centers, chunk=None):
    chunk = chunk or _chunk(centers.shape[0])
    out = torch.empty((x.shape[0],), dtype=torch.int64, device=x.device)
    cn = (centers * centers).sum(1)
    for s in range(0, x.shape[0], chunk):
        out[s:s + chunk] = (cn[None, :] - 2.0 * (x[s:s + chunk] @ centers.t())).argmin(1)
    return out


def encode_packed(torch, x, centers, a, codebook, chunk=262144):
    """rows x (already in index order) with their partition a -> PackedCodes4Bit rows [n, ceil(S/2)] u8"""
    S, ncodes, ds = codebook.shape
    bpp = (S + 1) // 2
    out = torch.empty((x.shape[0], bpp), dtype=torch.uint8, device=x.device)
    for s in range(0, x.shape[0], chunk):
        r = (x[s:s + chunk] - centers[a[s:s + chunk]]).view(-1, S, 1, ds)
        d = ((r - codebook[None]) ** 2).sum(-1)            # [m, S, 16]
        codes = d.argmin(-1).to(torch.uint8)               # [m, S], lower code on ties
        if S % 2:
            codes = torch.cat([codes, torch.zeros_like(codes[:, :1])], 1)
        out[s:s + chunk] = codes[:, 0::2] | (codes[:, 1::2] << 4)
    return out


def build_treeah(torch, x, K, S, train_sample=1_000_000, iters=20, seed=7, balance_ratio=0.0):
    """-> dict of torch tensors on x.device: centers, codebook, packed (grouped by partition), ids (i64), off (i64)"""
    n = x.shape[0]
    g = torch.Generator(device=x.device)
    g.manual_seed(seed)
    ns = min(train_sample, n)
    sample = x[torch.randperm(n, generator=g, device=x.device)[:ns]].contiguous() if ns < n else x
    centers = kmeans(torch, sample, K, iters, seed, balance_ratio=balance_ratio)
    K = centers.shape[0]
    a_s = assign(torch, sample, centers)
    codebook = train_codebook(torch, sample - centers[a_s], S, 16, iters, 42)
    del sample, a_s
    a = assign(torch, x, centers)
    order = torch.argsort(a, stable=True)
    counts = torch.bincount(a, minlength=K)
    off = torch.zeros((K + 1,), dtype=torch.int64, device=x.device)
    off[1:] = torch.cumsum(counts, 0)
    packed = torch.empty((n, (S + 1) // 2), dtype=torch.uint8, device=x.device)
    for s0 in range(0, n, 1 << 20):
        idx = order[s0:s0 + (1 << 20)]
        packed[s0:s0 + (1 << 20)] = encode_packed(torch, x[idx], centers, a[idx], codebook)
    return {"centers": centers, "codebook": codebook, "packed": packed, "ids": order, "off": off}
