#!/usr/bin/env python
"""Debug aid: one tensor-core-scan parity case (tests/test_gpu_parity._treeah_case) with SCANN_SCAN_TC=1."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["SCANN_SCAN_TC"] = "1"
import oracle
from test_gpu_parity import _treeah_case
pkg = importlib.import_module("scann-rust_b200")
n, dim, K, S, nq, L, R = (int(v) for v in sys.argv[1:8])
s = _treeah_case(pkg, oracle, n=n, dim=dim, K=K, S=S, nq=nq, L=L, R=R, k=10, measure=pkg.DistanceMeasure.SquaredL2,
                 seed=11, min_recall=0.99)
print("OK", sys.argv[1:], s.path_stats())
