// Host-side launchers shared between translation units (all take device pointers, enqueue on `s`).
#pragma once

#include "common.cuh"

namespace scann {

// partition.cu — exact centroid scoring + top-L (src/partitioning/tree_partitioner.rs:175-229)
//   centersT: [dim][K] transposed centres; scratch: [nq][K] f32; tokens [nq][L]; dists [nq][L] or null
scann_status launch_partition(const float* centersT, size_t K, size_t dim, const float* queries, size_t nq, size_t L,
                              uint32_t* tokens, float* dists, float* scratch, cudaStream_t s);
void launch_transpose(const float* in, size_t rows, size_t cols, float* out, cudaStream_t s);

// select.cu — final exact re-score + (dist, id) order, and the multi-GPU merge
//   cand: [nq][kc] candidate row ids (0xFFFFFFFF = empty); writes ids/dists [nq][k], counts [nq]
scann_status launch_rescore_topk(const RescoreParams& rp, const uint32_t* cand, size_t nq, size_t kc, size_t k,
                                 uint32_t* ids, float* dists, uint32_t* counts, cudaStream_t s);
scann_status launch_merge_topk(const uint32_t* ids_in, const float* dists_in, size_t parts, size_t nq, size_t k,
                               uint32_t* ids_out, float* dists_out, uint32_t* counts_out, cudaStream_t s);

// runtime.cu — accumulation mode of the LUT16 scan for a table of S subspaces (lut16_device.cuh scan_block):
// default 3 (IDP.2A, needs S <= 128) else 2; SCANN_ACC_MODE=0|1|2|3 overrides (tuning / parity tests)
int scan_acc_mode(int S);

}  // namespace scann
