"""Error conventions of the C ABI that can be observed WITHOUT a device (argument checks come before the device check):
status = the reference's ErrorCode ordinal (src/error.rs:10-45), message = the reference's own text where it has one
(SURVEY.md section 8b).  CPU test; no compute entry point gets as far as a kernel."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def L(pkg):
    import importlib
    return importlib.import_module("scann-rust_b200.capi").load()


def _arrs():
    f = np.zeros(64, np.float32)
    return f, np.zeros(64, np.uint8), np.zeros(64, np.uint32), np.zeros(8, np.uint64)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


INVALID_ARGUMENT, FAILED_PRECONDITION, UNIMPLEMENTED = 3, 9, 12

CASES = [
    # (id, symbol, args builder, status, message fragment, reference line the convention comes from)
    ("bf_null_dataset", "scann_bf_create", lambda f, b, u, o, h: (None, 4, 8, 8, 0, 0, 0, C.byref(h)),
     INVALID_ARGUMENT, "bad dataset arguments", "brute_force/searcher.rs:34-55"),
    ("bf_stride_lt_dim", "scann_bf_create", lambda f, b, u, o, h: (_ptr(f), 4, 8, 4, 0, 0, 0, C.byref(h)),
     INVALID_ARGUMENT, "bad dataset arguments", "data_format/dataset.rs:90-96"),
    ("bf_measure_off_path", "scann_bf_create", lambda f, b, u, o, h: (_ptr(f), 4, 8, 8, 9, 0, 0, C.byref(h)),
     UNIMPLEMENTED, "outside the GPU hot path", "distance_measures/mod.rs:70-81"),
    ("bf_search_unbuilt", "scann_bf_search", lambda f, b, u, o, h: (None, _ptr(f), 1, 8, 1, None, None, None, 0, None),
     FAILED_PRECONDITION, "not built", "partitioning/tree_partitioner.rs:197-198"),
    ("sq8_null_codes", "scann_sq8_create", lambda f, b, u, o, h: (None, 4, 8, 1.0, 0, 0, 0, C.byref(h)),
     INVALID_ARGUMENT, "bad dataset arguments", "brute_force/scalar_quantized.rs:99-116"),
    ("sq8_quantize_null", "scann_sq8_quantize", lambda f, b, u, o, h: (None, 4, 8, 8, _ptr(b), _ptr(f), 0, 0),
     INVALID_ARGUMENT, "bad arguments", "quantization/scalar.rs:195-226"),
    ("part_empty", "scann_part_create", lambda f, b, u, o, h: (_ptr(f), 0, 8, 0, 0, C.byref(h)),
     INVALID_ARGUMENT, "Cannot partition empty dataset", "partitioning/tree_partitioner.rs:50"),
    ("part_select_unbuilt", "scann_part_select", lambda f, b, u, o, h: (None, _ptr(f), 1, 8, 1, None, None, 0, None),
     FAILED_PRECONDITION, "Partitioner not built", "partitioning/tree_partitioner.rs:197-198"),
    ("treeah_dim_not_divisible", "scann_treeah_create",
     lambda f, b, u, o, h: (_ptr(f), 2, 10, _ptr(f), 4, _ptr(b), _ptr(u), _ptr(o), 8, None, 0, 0, 1, 0, 0, 0, C.byref(h)),
     INVALID_ARGUMENT, "Dimensionality 10 must be divisible by num_subspaces 4", "hashes/codebook.rs:154-159"),
    ("treeah_too_many_subspaces", "scann_treeah_create",
     lambda f, b, u, o, h: (_ptr(f), 2, 1028, _ptr(f), 257, _ptr(b), _ptr(u), _ptr(o), 8, None, 0, 0, 1, 0, 0, 0, C.byref(h)),
     INVALID_ARGUMENT, "num_subspaces 257 outside 1..256", "DESIGN.md limits"),
    ("treeah_empty", "scann_treeah_create",
     lambda f, b, u, o, h: (_ptr(f), 0, 8, _ptr(f), 4, _ptr(b), _ptr(u), _ptr(o), 8, None, 0, 0, 1, 0, 0, 0, C.byref(h)),
     INVALID_ARGUMENT, "Cannot build from empty dataset", "tree_x_hybrid/mod.rs:132-134"),
    ("treeah_null_codebook", "scann_treeah_create",
     lambda f, b, u, o, h: (_ptr(f), 2, 8, None, 4, _ptr(b), _ptr(u), _ptr(o), 8, None, 0, 0, 1, 0, 0, 0, C.byref(h)),
     INVALID_ARGUMENT, "NULL index array", "-"),
    ("treeah_search_unbuilt", "scann_treeah_search",
     lambda f, b, u, o, h: (None, _ptr(f), 1, 8, 1, 1, 1, None, None, None, None, None, None, 0, None),
     FAILED_PRECONDITION, "not built", "tree_x_hybrid/mod.rs:240-250"),
    ("treeah_search_end_unbuilt", "scann_treeah_search_end", lambda f, b, u, o, h: (None, None, None, None, None, None),
     FAILED_PRECONDITION, "not built", "-"),
    ("treeah_abort_unbuilt", "scann_treeah_search_abort", lambda f, b, u, o, h: (None,),
     FAILED_PRECONDITION, "not built", "-"),
    ("treeah_filter_unbuilt", "scann_treeah_set_filter", lambda f, b, u, o, h: (None, None, 0, 0),
     FAILED_PRECONDITION, "not built", "-"),
    ("treeah_build_dim_not_divisible", "scann_treeah_build",
     lambda f, b, u, o, h: (_ptr(f), 8, 6, 6, 2, 4, 100, 5, 7, 1, 0, 0, 0, 0, C.byref(h)),
     INVALID_ARGUMENT, "Dimensionality 6 must be divisible by num_subspaces 4", "hashes/codebook.rs:154-159"),
    ("treeah_build_empty", "scann_treeah_build",
     lambda f, b, u, o, h: (_ptr(f), 0, 8, 8, 2, 4, 100, 5, 7, 1, 0, 0, 0, 0, C.byref(h)),
     INVALID_ARGUMENT, "Cannot build from empty dataset", "tree_x_hybrid/mod.rs:132-134"),
    ("kmeans_more_clusters_than_rows", "scann_kmeans_fit", lambda f, b, u, o, h: (_ptr(f), 4, 8, 8, 9, 5, 7, 0.0, _ptr(f), 0, 0),
     INVALID_ARGUMENT, "num_clusters 9 outside 1..4", "trees/kmeans.rs:166-200"),
    ("ivf_search_unbuilt", "scann_ivf_search",
     lambda f, b, u, o, h: (None, 0, _ptr(f), 1, 8, 1, 1, 0, 0, None, None, None, 0, None),
     FAILED_PRECONDITION, "not built", "scann.rs:215-226"),
    ("kmtree_search_unbuilt", "scann_kmtree_search_leaves",
     lambda f, b, u, o, h: (None, _ptr(f), 1, 8, 1, None, None, None, None, 0, None),
     FAILED_PRECONDITION, "Tree not built", "trees/kmeans_tree.rs:302-316"),
    ("pq_encode_bad_shape", "scann_pq_encode", lambda f, b, u, o, h: (_ptr(f), 4, 0, _ptr(f), 2, 8, None, None, _ptr(b), 0, 0),
     INVALID_ARGUMENT, "bad S/ds/stride", "hashes/codebook.rs:205-216"),
    ("lut16_build_bad_shape", "scann_lut16_build",
     lambda f, b, u, o, h: (_ptr(f), 0, 2, _ptr(f), 1, None, _ptr(b), _ptr(f), _ptr(f), 0, 0),
     INVALID_ARGUMENT, "bad S/ds", "hashes/lut16.rs:151-173"),
    ("lut16_scan_bad_shape", "scann_lut16_scan", lambda f, b, u, o, h: (_ptr(b), 4, 0, _ptr(b), _ptr(u), 0, 0),
     INVALID_ARGUMENT, "num_subspaces 0 outside 1..256", "simd/dispatch.rs:246-295"),
    ("tc_scores_dim_too_large", "scann_tc_scores",
     lambda f, b, u, o, h: (_ptr(f), 1, 300, _ptr(f), 0, 1, 300, 1.0, 0, None, _ptr(f), None, 0, None, 0),
     INVALID_ARGUMENT, "dimension 300 too large for the tensor-core path", "DESIGN.md limits"),
    ("merge_topk_null", "scann_merge_topk", lambda f, b, u, o, h: (None, None, 0, 1, 1, None, None, None, 0, 0, None),
     INVALID_ARGUMENT, "NULL buffer", "-"),
    ("merge_topk_packed_null", "scann_merge_topk_packed", lambda f, b, u, o, h: (None, 2, 1, 1, None, None, None, 0, None),
     INVALID_ARGUMENT, "NULL buffer", "-"),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_argument_errors_come_before_the_device_check(L, case):
    _, sym, build, status, fragment, _ref = case
    f, b, u, o = _arrs()
    h = C.c_void_p()
    rc = getattr(L, sym)(*build(f, b, u, o, h))
    msg = (L.scann_last_error() or b"").decode("utf-8", "replace")
    assert rc == status, f"{sym}: status {rc} ({msg})"
    assert fragment in msg, f"{sym}: message {msg!r}"
    assert not h.value, "no handle may be returned with an error status"


def test_last_error_is_thread_local(L):
    import threading
    f, b, u, o = _arrs()
    h = C.c_void_p()
    assert L.scann_part_create(_ptr(f), 0, 8, 0, 0, C.byref(h)) == INVALID_ARGUMENT
    seen = {}

    def other():
        seen["before"] = (L.scann_last_error() or b"").decode()
        L.scann_lut16_scan(_ptr(b), 4, 0, _ptr(b), _ptr(u), 0, 0)
        seen["after"] = (L.scann_last_error() or b"").decode()

    t = threading.Thread(target=other)
    t.start()
    t.join()
    assert "Cannot partition" not in seen["before"] and "num_subspaces 0" in seen["after"]
    assert "Cannot partition empty dataset" in (L.scann_last_error() or b"").decode()  # this thread's message is untouched
