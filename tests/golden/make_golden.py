#!/usr/bin/env python
"""Generates the committed golden vectors of tests/golden/*.npz.

Provenance: the reference (sunbains/scann-rust) is a Rust crate that cannot be built in this image and ships no
golden-vector files, so these vectors are outputs of the CPU ORACLE (oracle/scann_oracle.cpp — the restatement that is
pinned against the reference's own known-answer tests, tests/test_oracle_kat.py) on small seeded inputs.  They freeze
that behaviour: tests/test_golden.py checks that the oracle still reproduces them (CPU) and that the CUDA path matches
them (GPU), independently of the live oracle-vs-GPU comparisons in the other test files; the second, independent restatement
(tests/ref_restatement.py) must reproduce them as well, without the oracle in the loop.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import helpers
    import oracle
    oracle.build()
    # --- LUT16: u8 tables + integer accumulators
    rng = np.random.default_rng(2024)
    S, ds = 12, 2
    cb = rng.normal(0, 0.4, (S, 16, ds)).astype(np.float32)
    q = rng.normal(0, 1, (6, S * ds)).astype(np.float32)
    luts, bias, mult = [], [], []
    for i in range(len(q)):
        l8, b, m, _ = oracle.lut16_build(cb, q[i], None)
        luts.append(l8)
        bias.append(b)
        mult.append(m)
    codes = rng.integers(0, 16, (700, S), dtype=np.uint8)
    packed = oracle.pack4(codes)
    sums = np.stack([oracle.lut16_scan_u32(packed, l8, S) for l8 in luts])
    np.savez_compressed(os.path.join(HERE, "lut16_small.npz"), codebook=cb, queries=q, lut8=np.stack(luts),
                        bias=np.array(bias, np.float32), mult=np.array(mult, np.float32), packed=packed, sums=sums)
    # --- brute force f32 / int8
    db = helpers.gaussian(900, 24, 11)
    qq = helpers.gaussian(12, 24, 12)
    out = {"db": db, "queries": qq}
    for name, m in (("sql2", oracle.SQL2), ("dot", oracle.DOT)):
        rc, ids, dists, counts = oracle.bf_search(db, qq, 7, m)
        out[f"{name}_ids"], out[f"{name}_dists"] = ids, dists
    codes8, cal = oracle.sq8_quantize(db)
    rc, ids, dists, counts = oracle.sq8_search(codes8, float(cal[2]), qq, 7, oracle.DOT)
    out.update(sq8_codes=codes8, sq8_cal=cal, sq8_dot_ids=ids, sq8_dot_dists=dists)
    np.savez_compressed(os.path.join(HERE, "bf_small.npz"), **out)
    # --- partition + Tree-AH
    x, _ = helpers.clustered(2500, 16, 12, 0.35, 5)
    tq = (x[:10] + 0.03).astype(np.float32)
    idx = helpers.build_index(oracle, x, 9, 4)
    tok, tdist = oracle.partition(idx["centers"], tq, 4)
    # (a) R larger than the probed leaves: no cut-off, so the candidate set and the final result are tie-free
    rc, ids, dists, counts = oracle.treex_search(idx["centers"], idx["codebook"], idx["part_offsets"], idx["ids"],
                                                 idx["packed"], x, tq, 3, 2000, 8, lut16=True)
    # (b) R = 30: the sorted approximate candidate distances (tie-independent) of a real cut-off
    rc, _, _, _, cand, cand_d, cand_n = oracle.treex_search(idx["centers"], idx["codebook"], idx["part_offsets"],
                                                            idx["ids"], idx["packed"], x, tq, 4, 30, 8, lut16=True,
                                                            want_candidates=True)
    np.savez_compressed(os.path.join(HERE, "treeah_small.npz"), x=x, queries=tq, centers=idx["centers"],
                        codebook=idx["codebook"], part_offsets=idx["part_offsets"], ids=idx["ids"], packed=idx["packed"],
                        tokens=tok, token_dists=tdist, out_ids=ids, out_dists=dists, out_counts=counts,
                        cand_dists=cand_d, cand_counts=cand_n)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
