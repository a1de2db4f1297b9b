import importlib, sys, os
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
pkg = importlib.import_module("scann-rust_b200")
import oracle, helpers
x, _ = helpers.clustered(20_000, 32, 32, 0.35, 1)
q = (x[:32] + 0.02).astype(np.float32)
idx = helpers.build_index(oracle, x, 16, 8)
s = pkg.TreeXHybridSearcher(pkg.TreeXHybridConfig(num_partitions=16, partitions_to_search=4))
s.build_from_index(idx["centers"], idx["codebook"], idx["packed"], idx["ids"], idx["part_offsets"], x)
ids, dists, counts, (ci, cd, cc) = s.search_batched(q, 10, pre_reorder_k=50, want_candidates=True)
rc, oids, odists, ocounts, ocand, ocd, ocn = oracle.treex_search(idx["centers"], idx["codebook"], idx["part_offsets"], idx["ids"], idx["packed"], x, q, 4, 50, 10, lut16=True, want_candidates=True)
print('recall', helpers.recall(ids, oids, 10), 'cand dist equal', (cd.view(np.uint32) == ocd.view(np.uint32)).all())
for i in range(32):
    if set(ids[i]) != set(oids[i]):
        print('q', i, 'gpu', ids[i], dists[i]); print('    orc', oids[i], odists[i])
        print('   cand cutoff gpu', cd[i][-3:], 'orc', ocd[i][-3:], 'cand set diff', set(ci[i]) ^ set(ocand[i]))
