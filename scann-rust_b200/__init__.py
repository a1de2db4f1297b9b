"""scann-rust_b200 — B200-native (sm_100a) batched-search hot path of sunbains/scann-rust.

The product is ``lib/libscann_b200.so`` (hand-written CUDA behind the C ABI of ``include/scann_b200.h``);
this package is the host-side mirror of the reference's searcher API over that ABI plus the GPU index
builder.  There is no CPU fallback: anything that computes raises ``ScannError(UNAVAILABLE)`` when the
CUDA library or a device is missing.  (The directory name has a hyphen; import it with
``importlib.import_module("scann-rust_b200")``.)
"""
from . import build as build_lib  # noqa: F401
from . import capi, distributed, indexing, scann, searchers  # noqa: F401
from .capi import ScannError, device_count, load  # noqa: F401
from .scann import Scann, ScannBuilder, ScannConfig, SearchMode  # noqa: F401
from .searchers import (AsymmetricHasher, KMeansTree, KMeansTreeConfig, LeafScanSearcher, AsymmetricHasherConfig, BruteForceSearcher, DistanceMeasure,  # noqa: F401
                        ScalarQuantizedBruteForceSearcher, ScalarQuantizedConfig, SearchParameters, TreePartitioner, TreeXHybridConfig,
                        TreeXHybridSearcher, lut16_build, lut16_scan, merge_topk, merge_topk_packed, pq_encode, results_to_lists,
                        scalar_quantize, tc_scores)
