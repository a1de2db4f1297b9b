"""TEST INFRASTRUCTURE.  A SECOND, independent restatement of the reference's hot path in plain Python / numpy scalars,
written from the Rust source (not from oracle/scann_oracle.cpp), used to cross-check the C++ oracle on small cases: if two
independent restatements agree bit for bit — distances, ids AND tie order — a transcription mistake in either would have
to be made twice.  (Neither is output of the reference itself: the crate cannot be built in this image.)  First the
approximate stage of Tree-AH; further down the AVX2 + FMA distance kernels with an exact fused multiply-add, the exact
reorder, the brute-force searchers, the PQ pieces, the flat AsymmetricHasher, the facade's tree modes and KMeansTree —
each block names the Rust lines it follows.

Follows, line for line:
  TreePartitioner::partition          src/partitioning/tree_partitioner.rs:175-229  (sequential f32 sum of d*d, stable sort)
  residual                            src/tree_x_hybrid/mod.rs:307-315              (q - centroid per dimension)
  Lut16LookupTables::from_query       src/hashes/lut16.rs:151-173                   (squared_l2 per code, sequential)
  Lut16SimdTables::from_float_tables  src/hashes/lut16_simd.rs:39-90                (global min/max, 255/range, round, as u8)
  lut16_distances_batch_portable      src/simd/dispatch.rs:259-295                  (low nibble first, u32 sum, as f32)
  dequantisation                      src/hashes/lut16_simd.rs:134-140              (sum * multiplier + bias * S)
  FastTopNeighbors::push / results    src/brute_force/top_k.rs:333-383              (first slot holding the max is replaced)
  search_with_filter merge            src/tree_x_hybrid/mod.rs:282-291              (flatten, stable sort, truncate)
"""
import numpy as np

F = np.float32


def sqdist_seq(a, b):
    s = F(0.0)
    for x, y in zip(a, b):
        d = F(F(x) - F(y))
        s = F(s + F(d * d))
    return s


def partition(centers, q, L):
    d = [sqdist_seq(q, c) for c in centers]
    order = sorted(range(len(centers)), key=lambda i: (d[i], i))  # sort_by_key(OrderedFloat) is stable: ties keep index order
    return order[:min(L, len(order))]


def rust_round(x):
    """f32::round: half away from zero"""
    x = F(x)
    return F(np.floor(x + F(0.5))) if x >= 0 else F(np.ceil(x - F(0.5)))


def as_u8(x):
    """Rust `as u8`: saturating, NaN -> 0"""
    if x != x:
        return 0
    return int(min(max(float(x), 0.0), 255.0))


def lut16_tables(codebook, resid):
    """codebook [S][16][ds] -> (u8 tables [S][16], bias, multiplier)"""
    S, C, ds = codebook.shape
    ft = np.empty((S, C), np.float32)
    for s in range(S):
        for c in range(C):
            ft[s, c] = sqdist_seq(resid[s * ds:(s + 1) * ds], codebook[s, c])
    gmin, gmax = F(np.finfo(np.float32).max), F(np.finfo(np.float32).min)
    for v in ft.flat:
        gmin = min(gmin, F(v))
        gmax = max(gmax, F(v))
    rng = F(gmax - gmin)
    if rng < F(1e-10):
        mult, scale = F(1.0), F(1.0)
    else:
        scale = F(F(255.0) / rng)
        mult = F(F(1.0) / scale)
    q8 = np.empty((S, C), np.uint8)
    for s in range(S):
        for c in range(C):
            q8[s, c] = as_u8(rust_round(F(F(ft[s, c] - gmin) * scale)))
    return q8, gmin, mult


class FastTopNeighbors:
    def __init__(self, capacity):
        self.cap = capacity
        self.idx, self.dist = [], []

    def push(self, index, distance):
        if len(self.idx) < self.cap:
            self.idx.append(index)
            self.dist.append(distance)
        elif self.cap > 0:
            mi, md = 0, self.dist[0]
            for i in range(1, len(self.dist)):
                if self.dist[i] > md:
                    md, mi = self.dist[i], i
            if distance < md:
                self.idx[mi], self.dist[mi] = index, distance

    def results(self):
        return sorted(zip(self.idx, self.dist), key=lambda t: t[1])  # stable


def approx_candidates(centers, codebook, part_offsets, ids, packed, q, L, R, use_residuals=True, allow=None):
    """-> [(datapoint index, approximate distance)] in the order search_with_filter hands to reorder_results.
    allow: optional RestrictFilter as a set of allowed datapoint ids (checked before the push, tree_x_hybrid/mod.rs:325-331)"""
    S = codebook.shape[0]
    bpp = (S + 1) // 2
    allr = []
    for leaf in partition(centers, q, L):
        resid = np.array([F(F(a) - F(b)) for a, b in zip(q, centers[leaf])], np.float32) if use_residuals else q
        q8, bias, mult = lut16_tables(codebook, resid)
        bias_total = F(bias * F(S))
        top = FastTopNeighbors(R)
        for row in range(int(part_offsets[leaf]), int(part_offsets[leaf + 1])):
            if allow is not None and int(ids[row]) not in allow:
                continue
            total, sub = 0, 0
            for b in range(bpp):
                byte = int(packed[row, b])
                if sub < S:
                    total += int(q8[sub, byte & 0x0F])
                    sub += 1
                if sub < S:
                    total += int(q8[sub, (byte >> 4) & 0x0F])
                    sub += 1
            top.push(int(ids[row]), F(F(F(total) * mult) + bias_total))
        allr.extend(top.results())
    allr.sort(key=lambda t: t[1])  # stable
    return allr[:R]


class RustBinaryHeap:
    """std::collections::BinaryHeap<(OrderedFloat<f32>, u32)> (max-heap), the algorithms of library/alloc binary_heap:
    push = sift_up(0, old_len); pop = swap the last element into the root, sift_down_to_bottom(0), then sift_up."""

    def __init__(self):
        self.data = []

    def __len__(self):
        return len(self.data)

    def peek(self):
        return self.data[0] if self.data else None

    def _sift_up(self, start, pos):
        d = self.data
        elt = d[pos]
        while pos > start:
            parent = (pos - 1) // 2
            if elt <= d[parent]:
                break
            d[pos] = d[parent]
            pos = parent
        d[pos] = elt
        return pos

    def push(self, item):
        self.data.append(item)
        self._sift_up(0, len(self.data) - 1)

    def pop(self):
        d = self.data
        if not d:
            return None
        item = d.pop()
        if d:
            item, d[0] = d[0], item
            end, start, pos = len(d), 0, 0
            elt = d[pos]
            child = 2 * pos + 1
            while child <= max(end - 2, 0) and child + 1 < end:
                if d[child] <= d[child + 1]:
                    child += 1
                d[pos] = d[child]
                pos = child
                child = 2 * pos + 1
            if child == end - 1:
                d[pos] = d[child]
                pos = child
            d[pos] = elt
            self._sift_up(start, pos)
        return item


class TopK:
    """src/brute_force/top_k.rs:19-112: heap of (distance, index); results() = heap.iter() order, stable-sorted by distance"""

    def __init__(self, k):
        self.k = k
        self.heap = RustBinaryHeap()

    def push(self, index, distance):
        distance = float(np.float32(distance))
        if len(self.heap) < self.k:
            self.heap.push((distance, index))
            return True
        top = self.heap.peek()
        if top is not None and distance < top[0]:
            self.heap.pop()
            self.heap.push((distance, index))
            return True
        return False

    def results(self):
        return sorted(((i, d) for d, i in self.heap.data), key=lambda t: t[1])


def sq8_quantize(db, num_std_devs=3.0, bits=8):
    """QuantizationStats::from_dataset (src/quantization/mod.rs:77-110: sequential f64 sums, sample variance) +
    ScalarQuantizer::calibrate (scalar.rs:103-130, the non-symmetric std-dev branch) + quantize_value (:161-165)
    -> (codes i8 [n, dim], (min_value, max_value, scale, inv_scale))"""
    mn, mx = F(np.finfo(np.float32).max), F(np.finfo(np.float32).min)
    s = s2 = 0.0
    cnt = 0
    for row in db:
        for v in row:
            v = F(v)
            mn, mx = min(mn, v), max(mx, v)
            s += float(v)
            s2 += float(v) * float(v)
            cnt += 1
    mean = F(s / cnt) if cnt > 0 else F(0.0)
    var = F((s2 - s * s / cnt) / (cnt - 1)) if cnt > 1 else F(0.0)
    std = F(np.sqrt(var))
    levels = (1 << bits) - 1
    rng0 = F(F(num_std_devs) * std)
    lo, hi = max(F(mean - rng0), mn), min(F(mean + rng0), mx)
    rng = F(hi - lo)
    if rng > F(1e-10):
        scale, inv = F(rng / F(levels)), F(F(levels) / rng)
    else:
        scale, inv = F(1.0), F(1.0)
    codes = np.empty(db.shape, np.int8)
    for i, row in enumerate(db):
        for j, v in enumerate(row):
            c = min(max(F(v), lo), hi)
            qi = int(rust_round(F(F(c - lo) * inv)))
            qi = min(max(qi, 0), levels)
            codes[i, j] = np.array(qi, np.int64).astype(np.uint8).view(np.int8)  # `as i8` wraps 128..255 to -128..-1
    return codes, (lo, hi, scale, inv)


# ----------------------------------------------------------------------------------------------------------------------
# The AVX2 + FMA distance kernels, lane for lane, with the fused multiply-add evaluated EXACTLY (rational arithmetic, one
# rounding to f32) — written from the Rust source:
#   horizontal_sum_f32_avx2                       src/simd/x86.rs:31-44   ((v0+v4) + (v1+v5)) + ((v2+v6) + (v3+v7))
#   dot_product_avx2 / squared_l2_avx2            src/simd/x86.rs:72-96, 139-165
#   one_to_many_{dot_product,squared_l2}_avx2     src/simd/x86.rs:195-346 (3 / 4 rows in flight: same arithmetic per row)
#   one_to_many_int8_float_{dot_product,squared_l2}_avx2   src/distance_measures/one_to_many_asymmetric.rs:77-144, 207-261
#   horizontal_sum_avx (hadd, hadd)               src/distance_measures/one_to_many_asymmetric.rs:383-399 (same tree)
from fractions import Fraction  # noqa: E402


def f32_from_fraction(fr):
    """Round an exact rational to the nearest f32, ties to even (normal and subnormal range; overflow -> inf)."""
    if fr == 0:
        return F(0.0)
    neg = fr < 0
    a = -fr if neg else fr
    e = a.numerator.bit_length() - a.denominator.bit_length()
    if Fraction(2) ** e > a:
        e -= 1
    if Fraction(2) ** (e + 1) <= a:
        e += 1
    qexp = max(e, -126) - 23  # exponent of the last mantissa bit
    scaled = a / (Fraction(2) ** qexp)
    n = scaled.numerator // scaled.denominator
    rem = scaled - n
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and (n & 1)):
        n += 1
    if n * Fraction(2) ** qexp >= Fraction(2) ** 128:
        v = float("inf")
    else:
        v = float(Fraction(n) * Fraction(2) ** qexp)  # <= 25 significant bits: exact in f64
    return F(-v if neg else v)


def fma32(a, b, c):
    """_mm256_fmadd_ps per lane: a * b + c with ONE rounding"""
    a, b, c = F(a), F(b), F(c)
    if not (np.isfinite(a) and np.isfinite(b) and np.isfinite(c)):
        return F(a * b + c)
    return f32_from_fraction(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))


def hsum8(v):
    s = [F(v[i] + v[i + 4]) for i in range(4)]      # extractf128 + add_ps
    return F(F(s[0] + s[1]) + F(s[2] + s[3]))       # movehdup/add, movehl/add_ss  (= hadd, hadd)


def _row_f32(q, x, measure):
    dim = len(q)
    chunks = dim // 8
    acc = [F(0.0)] * 8
    for c in range(chunks):
        for l in range(8):
            j = 8 * c + l
            if measure == "dot":
                acc[l] = fma32(q[j], x[j], acc[l])
            else:
                d = F(F(q[j]) - F(x[j]))
                acc[l] = fma32(d, d, acc[l])
    r = hsum8(acc)
    for j in range(8 * chunks, dim):  # scalar tail: multiply, then add (Rust does not contract)
        if measure == "dot":
            r = F(r + F(F(q[j]) * F(x[j])))
        else:
            d = F(F(q[j]) - F(x[j]))
            r = F(r + F(d * d))
    return F(-r) if measure == "dot" else r


def one_to_many_f32(q, db, measure):
    return np.array([_row_f32(q, x, measure) for x in db], np.float32)


def one_to_many_i8(q, db_i8, inv_mul, measure):
    out = []
    inv = F(inv_mul)
    for x in db_i8:
        dim = len(q)
        chunks = dim // 8
        acc = [F(0.0)] * 8
        for c in range(chunks):
            for l in range(8):
                j = 8 * c + l
                xs = F(F(int(x[j])) * inv)  # cvtepi8 -> cvtepi32_ps -> mul_ps: sign-extended, no offset
                if measure == "dot":
                    acc[l] = fma32(q[j], xs, acc[l])
                else:
                    d = F(F(q[j]) - xs)
                    acc[l] = fma32(d, d, acc[l])
        r = hsum8(acc)
        for j in range(8 * chunks, dim):
            xs = F(F(int(x[j])) * inv)
            if measure == "dot":
                r = F(r + F(F(q[j]) * xs))
            else:
                d = F(F(q[j]) - xs)
                r = F(r + F(d * d))
        out.append(F(-r) if measure == "dot" else r)
    return np.array(out, np.float32)


def reorder_results(raw, q, cand_ids, k, measure):
    """TreeXHybridSearcher::reorder_results (src/tree_x_hybrid/mod.rs:342-364): exact distance of every candidate by the
    single-pair kernel (DistanceMeasure::distance -> squared_l2_avx2 / -dot_product_avx2, src/distance_measures/mod.rs:70-81,
    one_to_one.rs:156-214,464-469 — the same lane arithmetic as the one-to-many kernels), candidates whose id has no row
    are skipped (dataset.get), stable sort by distance only, truncate."""
    ex = []
    for i in cand_ids:
        if int(i) < len(raw):
            d = _row_f32(q, raw[int(i)], "dot" if measure == "dot" else "sql2")
            if measure == "l2":
                d = F(np.sqrt(d))
            ex.append((int(i), d))
    ex.sort(key=lambda t: t[1])  # sort_by(partial_cmp): stable, ties keep the candidate (= approximate) order
    return ex[:k]


def bf_search(db, q, k, measure):
    """BruteForceSearcher::search_impl (src/brute_force/searcher.rs:96-139): all distances by the one-to-many kernel (sqrt
    pass for L2), N pushes through TopK in index order, drain_sorted; k clamped to N (:91)."""
    d = one_to_many_f32(q, db, "dot" if measure == "dot" else "sql2")
    if measure == "l2":
        d = np.sqrt(d).astype(np.float32)
    t = TopK(min(k, len(db)))
    for i, v in enumerate(d):
        t.push(i, v)
    return t.results()


def sq8_search(codes, scale, q, k, measure):
    """ScalarQuantizedBruteForceSearcher::search_impl (src/brute_force/scalar_quantized.rs:187-246): the int8 one-to-many
    kernels with inv_multiplier = quantizer.scale(), stride = dim, then TopK as above."""
    d = one_to_many_i8(q, codes, scale, "dot" if measure == "dot" else "sql2")
    if measure == "l2":
        d = np.sqrt(d).astype(np.float32)
    t = TopK(min(k, len(codes)))
    for i, v in enumerate(d):
        t.push(i, v)
    return t.results()


# ----------------------------------------------------------------------------------------------------------------------
# Product-quantisation pieces and the flat AsymmetricHasher, from the Rust source:
#   SubspaceCodebook::encode / compute_distances   src/hashes/codebook.rs:82-115   (strict '<': lowest code on ties)
#   Codebook::encode                               src/hashes/codebook.rs:205-216
#   PackedCodes4Bit::from_codes                    src/hashes/lut16.rs:43-61       (low nibble = even subspace)
#   LookupTable::from_query / compute_distance     src/hashes/lut.rs:47-82         (sequential f32 sum)
#   AsymmetricHasher::search / search_with_reordering   src/hashes/hasher.rs:162-229 (exact SqL2 hard-coded for the re-rank)
def pq_encode(codebook, x):
    S, C, ds = codebook.shape
    codes = []
    for s in range(S):
        sub = x[s * ds:(s + 1) * ds]
        best, best_d = 0, F(np.inf)
        for c in range(C):
            d = sqdist_seq(sub, codebook[s, c])
            if d < best_d:
                best_d, best = d, c
        codes.append(best)
    return codes


def pack4(point_codes):
    out = []
    for i in range(0, len(point_codes), 2):
        lo = point_codes[i] & 0x0F
        hi = ((point_codes[i + 1] & 0x0F) << 4) if i + 1 < len(point_codes) else 0
        out.append(lo | hi)
    return out


def lut_f32(codebook, q):
    S, C, ds = codebook.shape
    return [[sqdist_seq(q[s * ds:(s + 1) * ds], codebook[s, c]) for c in range(C)] for s in range(S)]


def lut_f32_distance(lut, codes):
    s = F(0.0)
    for sub, code in enumerate(codes):
        s = F(s + lut[sub][int(code)])
    return s


def ah_search(codebook, codes, q, k):
    lut = lut_f32(codebook, q)
    top = FastTopNeighbors(k)
    for i, c in enumerate(codes):
        top.push(i, lut_f32_distance(lut, c))
    return top.results()


def ah_search_with_reordering(codebook, codes, raw, q, k, pre_reorder_k):
    cand = ah_search(codebook, codes, q, pre_reorder_k)
    return reorder_results(raw, q, [i for i, _ in cand], k, "sql2")


# ----------------------------------------------------------------------------------------------------------------------
# The Scann facade's tree modes, from src/scann.rs:215-294 and src/utils/reordering.rs:23-54
def _measure_distance(q, x, measure):
    d = _row_f32(q, x, "dot" if measure == "dot" else "sql2")
    return F(np.sqrt(d)) if measure == "l2" else d


def scann_search_partitioned(centers, part_offsets, part_ids, raw, q, L, k, measure):
    """search_partitioned: members of the L closest partitions in token order, exact distance each by the single-pair
    kernel, stable sort by distance, truncate"""
    res = []
    for leaf in partition(centers, q, L):
        for idx in part_ids[int(part_offsets[leaf]):int(part_offsets[leaf + 1])]:
            if int(idx) < len(raw):
                res.append((int(idx), _measure_distance(q, raw[int(idx)], measure)))
    res.sort(key=lambda t: t[1])
    return res[:k]


def scann_search_tree_ah(centers, part_offsets, part_ids, codebook, codes_by_id, q, L, k):
    """search_tree_ah ("variant B"): ONE f32 LUT of the un-residualised query, every member of the L closest partitions
    scored by LookupTable::compute_distance, stable sort, truncate"""
    lut = lut_f32(codebook, q)
    res = []
    for leaf in partition(centers, q, L):
        for idx in part_ids[int(part_offsets[leaf]):int(part_offsets[leaf + 1])]:
            res.append((int(idx), lut_f32_distance(lut, codes_by_id[int(idx)])))
    res.sort(key=lambda t: t[1])
    return res[:k]


def reordering_helper(raw, q, results, k, measure):
    """ReorderingHelper::reorder: exact distance of the given results, stable sort, first k"""
    ex = [(i, _measure_distance(q, raw[i], measure)) for i, _ in results if i < len(raw)]
    ex.sort(key=lambda t: t[1])
    return ex[:k]


def kmtree_search_leaves(centers, depth, child_begin, child_count, children, q, k):
    """KMeansTree::search_leaves (src/trees/kmeans_tree.rs:302-355) on a tree flattened in preorder: depth-first, children
    in stable ascending order of their sequential squared distance, a leaf is always pushed, every level stops descending
    once 2k leaves are collected, stable sort by distance, first k.  -> [(node, distance, depth)]"""
    results = []

    def rec(node):
        d = sqdist_seq(q, centers[node])
        if int(child_count[node]) == 0:
            results.append((int(node), d, int(depth[node])))
            return
        kids = [int(children[int(child_begin[node]) + i]) for i in range(int(child_count[node]))]
        order = sorted(((i, sqdist_seq(q, centers[c])) for i, c in enumerate(kids)), key=lambda t: t[1])
        for i, _ in order:
            rec(kids[i])
            if len(results) >= 2 * k:
                break

    rec(0)
    results.sort(key=lambda t: t[1])
    return results[:k]


def bf_search_radius(db, q, radius, measure):
    """BruteForceSearcher::search_radius (src/brute_force/searcher.rs:142-167): every row with d <= radius, stable sort"""
    d = one_to_many_f32(q, db, "dot" if measure == "dot" else "sql2")
    if measure == "l2":
        d = np.sqrt(d).astype(np.float32)
    res = [(i, v) for i, v in enumerate(d) if v <= F(radius)]
    res.sort(key=lambda t: t[1])
    return res
