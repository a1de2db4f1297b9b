// scann_b200.hpp — header-only C++ mirror of the reference crate's searcher API over the C ABI
// (scann_b200.h).  The reference is compiled Rust and there is no Rust toolchain in this image, so this is
// the compiled-language host side above the ABI: same type and method names, same argument meaning and
// error behaviour as the Rust types (citations relative to /root/reference):
//
//   scann::BruteForceSearcher                 src/brute_force/searcher.rs:18-209
//   scann::ScalarQuantizedBruteForceSearcher  src/brute_force/scalar_quantized.rs:82-348
//   scann::TreePartitioner                    src/partitioning/tree_partitioner.rs:18-250
//   scann::TreeXHybridSearcher (+Config)      src/tree_x_hybrid/mod.rs:23-380
//   scann::LeafScanSearcher                   src/scann.rs:215-294 (Scann::search_partitioned / search_tree_ah)
//   scann::Result<T> / ScannError / ErrorCode src/error.rs:10-147
//
// `Result<T>` carries {code, message} like `Result<T, ScannError>`; nothing throws across the ABI.
#pragma once

#include <cstdint>
#include <limits>
#include <string>
#include <utility>
#include <vector>

#include "scann_b200.h"

namespace scann {

enum class ErrorCode : int32_t {
  Ok = 0, Cancelled, Unknown, InvalidArgument, DeadlineExceeded, NotFound, AlreadyExists, PermissionDenied,
  ResourceExhausted, FailedPrecondition, Aborted, OutOfRange, Unimplemented, Internal, Unavailable, DataLoss,
  Unauthenticated
};

struct ScannError {
  ErrorCode code = ErrorCode::Ok;
  std::string message;
};

template <class T>
struct Result {
  T value{};
  ScannError error;
  bool ok() const { return error.code == ErrorCode::Ok; }
};

inline ScannError make_error(scann_status st) {
  ScannError e;
  e.code = static_cast<ErrorCode>(st);
  if (st != SCANN_OK) e.message = scann_last_error();
  return e;
}

enum class DistanceMeasure : int { SquaredL2 = SCANN_SQL2, L2 = SCANN_L2, DotProduct = SCANN_DOT };

using DatapointIndex = uint32_t;                                    // src/types.rs:10
using NNResultsVector = std::vector<std::pair<DatapointIndex, float>>;  // src/types.rs:20

namespace detail {
inline std::vector<NNResultsVector> unflatten(const std::vector<uint32_t>& ids, const std::vector<float>& dists,
                                              const std::vector<uint32_t>& counts, size_t k) {
  std::vector<NNResultsVector> out(counts.size());
  for (size_t q = 0; q < counts.size(); ++q) {
    out[q].reserve(counts[q]);
    for (uint32_t j = 0; j < counts[q]; ++j) out[q].emplace_back(ids[q * k + j], dists[q * k + j]);
  }
  return out;
}
// &[Vec<f32>] → contiguous [nq*dim]; ragged batches are an InvalidArgument like a per-query dim mismatch
inline bool flatten(const std::vector<std::vector<float>>& queries, std::vector<float>* flat, size_t* dim) {
  *dim = queries.empty() ? 0 : queries[0].size();
  flat->clear();
  flat->reserve(queries.size() * *dim);
  for (const auto& q : queries) {
    if (q.size() != *dim) return false;
    flat->insert(flat->end(), q.begin(), q.end());
  }
  return true;
}
}  // namespace detail

// ------------------------------------------------------------------------------------------------
class BruteForceSearcher {
 public:
  // BruteForceSearcher::new(dataset, distance_measure); `data` is DenseDataset::raw_data() with `stride`
  static Result<BruteForceSearcher> create(const float* data, size_t n, size_t dim, size_t stride,
                                           DistanceMeasure measure, int device = 0) {
    Result<BruteForceSearcher> r;
    r.error = make_error(scann_bf_create(data, n, dim, stride, static_cast<int>(measure), device, SCANN_HOST, &r.value.h_));
    return r;
  }
  BruteForceSearcher() = default;
  BruteForceSearcher(BruteForceSearcher&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
  BruteForceSearcher& operator=(BruteForceSearcher&& o) noexcept {
    if (this != &o) { scann_bf_destroy(h_); h_ = o.h_; o.h_ = nullptr; }
    return *this;
  }
  BruteForceSearcher(const BruteForceSearcher&) = delete;
  ~BruteForceSearcher() { scann_bf_destroy(h_); }

  // search_batched(&[Vec<f32>], k) (searcher.rs:170-208)
  Result<std::vector<NNResultsVector>> search_batched(const std::vector<std::vector<float>>& queries, size_t k) const {
    Result<std::vector<NNResultsVector>> r;
    if (queries.empty()) return r;  // Ok(vec![])
    std::vector<float> flat;
    size_t dim;
    if (!detail::flatten(queries, &flat, &dim)) {
      r.error = {ErrorCode::InvalidArgument, "Query dimensionality does not match dataset dimensionality"};
      return r;
    }
    size_t nq = queries.size();
    std::vector<uint32_t> ids(nq * k), counts(nq);
    std::vector<float> dists(nq * k);
    r.error = make_error(scann_bf_search(h_, flat.data(), nq, dim, k, ids.data(), dists.data(), counts.data(),
                                         SCANN_HOST, nullptr));
    if (r.ok()) r.value = detail::unflatten(ids, dists, counts, k);
    return r;
  }
  // search(&[f32], k) (searcher.rs:77-93)
  Result<NNResultsVector> search(const std::vector<float>& query, size_t k) const {
    auto b = search_batched({query}, k);
    Result<NNResultsVector> r;
    r.error = b.error;
    if (b.ok() && !b.value.empty()) r.value = std::move(b.value[0]);
    return r;
  }
  // search_radius(&[T], radius) (searcher.rs:142-167): every row with distance <= radius, ascending.  max_results
  // bounds the output (ResourceExhausted when more rows qualify).
  Result<NNResultsVector> search_radius(const std::vector<float>& query, float radius, size_t max_results = 1024) const {
    Result<NNResultsVector> r;
    std::vector<uint32_t> ids(max_results), counts(1);
    std::vector<float> dists(max_results);
    r.error = make_error(scann_bf_search_radius(h_, query.data(), 1, query.size(), radius, max_results, ids.data(),
                                                dists.data(), counts.data(), SCANN_HOST, nullptr));
    if (r.ok()) r.value = detail::unflatten(ids, dists, counts, max_results)[0];
    return r;
  }

 private:
  scann_bf* h_ = nullptr;
};

// ------------------------------------------------------------------------------------------------
class ScalarQuantizedBruteForceSearcher {
 public:
  // ScalarQuantizedBruteForceSearcher::new(&dataset, config): quantise (QuantizedDataset::from_dataset) + build
  static Result<ScalarQuantizedBruteForceSearcher> create(const float* data, size_t n, size_t dim, size_t stride,
                                                          DistanceMeasure measure, int device = 0) {
    Result<ScalarQuantizedBruteForceSearcher> r;
    std::vector<int8_t> codes(n * dim);
    float cal[4];
    r.error = make_error(scann_sq8_quantize(data, n, dim, stride, codes.data(), cal, device, SCANN_HOST));
    if (!r.ok()) return r;
    r.value.scale_ = cal[2];
    r.error = make_error(scann_sq8_create(codes.data(), n, dim, cal[2], static_cast<int>(measure), device, SCANN_HOST,
                                          &r.value.h_));
    return r;
  }
  // from_quantized(quantized_dataset, distance_measure)
  static Result<ScalarQuantizedBruteForceSearcher> from_quantized(const int8_t* codes, size_t n, size_t dim, float scale,
                                                                  DistanceMeasure measure, int device = 0) {
    Result<ScalarQuantizedBruteForceSearcher> r;
    r.value.scale_ = scale;
    r.error = make_error(scann_sq8_create(codes, n, dim, scale, static_cast<int>(measure), device, SCANN_HOST, &r.value.h_));
    return r;
  }
  ScalarQuantizedBruteForceSearcher() = default;
  ScalarQuantizedBruteForceSearcher(ScalarQuantizedBruteForceSearcher&& o) noexcept : h_(o.h_), scale_(o.scale_) { o.h_ = nullptr; }
  ScalarQuantizedBruteForceSearcher& operator=(ScalarQuantizedBruteForceSearcher&& o) noexcept {
    if (this != &o) { scann_sq8_destroy(h_); h_ = o.h_; scale_ = o.scale_; o.h_ = nullptr; }
    return *this;
  }
  ScalarQuantizedBruteForceSearcher(const ScalarQuantizedBruteForceSearcher&) = delete;
  ~ScalarQuantizedBruteForceSearcher() { scann_sq8_destroy(h_); }

  Result<std::vector<NNResultsVector>> search_batched(const std::vector<std::vector<float>>& queries, size_t k) const {
    Result<std::vector<NNResultsVector>> r;
    if (queries.empty()) return r;
    std::vector<float> flat;
    size_t dim;
    if (!detail::flatten(queries, &flat, &dim)) {
      r.error = {ErrorCode::InvalidArgument, "Query dimensionality does not match dataset dimensionality"};
      return r;
    }
    size_t nq = queries.size();
    std::vector<uint32_t> ids(nq * k), counts(nq);
    std::vector<float> dists(nq * k);
    r.error = make_error(scann_sq8_search(h_, flat.data(), nq, dim, k, ids.data(), dists.data(), counts.data(),
                                          SCANN_HOST, nullptr));
    if (r.ok()) r.value = detail::unflatten(ids, dists, counts, k);
    return r;
  }
  float scale() const { return scale_; }

 private:
  scann_sq8* h_ = nullptr;
  float scale_ = 1.0f;
};

// ------------------------------------------------------------------------------------------------
struct PartitionResult {  // src/partitioning/partitioner.rs:12-21
  std::vector<uint32_t> tokens;
  std::vector<float> distances;
};

class TreePartitioner {
 public:
  static Result<TreePartitioner> from_centers(const float* centers, size_t K, size_t dim, int device = 0) {
    Result<TreePartitioner> r;
    r.value.dim_ = dim;
    r.error = make_error(scann_part_create(centers, K, dim, device, SCANN_HOST, &r.value.h_));
    return r;
  }
  TreePartitioner() = default;
  TreePartitioner(TreePartitioner&& o) noexcept : h_(o.h_), dim_(o.dim_) { o.h_ = nullptr; }
  TreePartitioner& operator=(TreePartitioner&& o) noexcept {
    if (this != &o) { scann_part_destroy(h_); h_ = o.h_; dim_ = o.dim_; o.h_ = nullptr; }
    return *this;
  }
  TreePartitioner(const TreePartitioner&) = delete;
  ~TreePartitioner() { scann_part_destroy(h_); }

  // Partitioner::partition(query, num_partitions) (tree_partitioner.rs:196-229)
  Result<PartitionResult> partition(const std::vector<float>& query, size_t num_partitions) const {
    Result<PartitionResult> r;
    if (!h_) {
      r.error = {ErrorCode::FailedPrecondition, "Partitioner not built"};
      return r;
    }
    r.value.tokens.resize(num_partitions);
    r.value.distances.resize(num_partitions);
    r.error = make_error(scann_part_select(h_, query.data(), 1, query.size(), num_partitions, r.value.tokens.data(),
                                           r.value.distances.data(), SCANN_HOST, nullptr));
    if (r.ok()) {  // drop the padding when num_partitions > K
      size_t m = 0;
      while (m < num_partitions && r.value.tokens[m] != 0xFFFFFFFFu) ++m;
      r.value.tokens.resize(m);
      r.value.distances.resize(m);
    }
    return r;
  }

 private:
  scann_part* h_ = nullptr;
  size_t dim_ = 0;
};

// ------------------------------------------------------------------------------------------------
struct TreeXHybridConfig {  // src/tree_x_hybrid/mod.rs:23-78
  size_t num_partitions = 100;
  size_t partitions_to_search = 10;
  size_t num_subspaces = 8;       // hash_config.num_subspaces (num_codes is 16 on the LUT16 path)
  bool use_residuals = true;
  float pre_reorder_multiplier = 3.0f;
  DistanceMeasure distance_measure = DistanceMeasure::SquaredL2;  // reorder measure (hard-wired SqL2 in the reference, :124)
  size_t pre_reorder_k(size_t k) const {  // (k as f32 * multiplier) as usize (:263)
    float v = static_cast<float>(k) * pre_reorder_multiplier;
    return v <= 0.0f ? 0 : static_cast<size_t>(v);
  }
};

class TreeXHybridSearcher {
 public:
  explicit TreeXHybridSearcher(TreeXHybridConfig cfg = {}) : cfg_(cfg) {}
  TreeXHybridSearcher(TreeXHybridSearcher&& o) noexcept : cfg_(o.cfg_), h_(o.h_) { o.h_ = nullptr; }
  TreeXHybridSearcher(const TreeXHybridSearcher&) = delete;
  ~TreeXHybridSearcher() { scann_treeah_destroy(h_); }

  // the arrays TreeXHybridSearcher::build leaves behind (:131-209)
  ScannError build_from_index(const float* centers, size_t K, size_t dim, const float* codebook, size_t S,
                              const uint8_t* packed, const uint32_t* ids, const uint64_t* part_offsets, size_t n,
                              const float* raw, size_t num_raw, size_t stride, int device = 0) {
    scann_treeah_destroy(h_);
    h_ = nullptr;
    return make_error(scann_treeah_create(centers, K, dim, codebook, S, packed, ids, part_offsets, n, raw, num_raw, stride,
                                          cfg_.use_residuals ? 1 : 0, static_cast<int>(cfg_.distance_measure), device,
                                          SCANN_HOST, &h_));
  }

  Result<std::vector<NNResultsVector>> search_batched(const std::vector<std::vector<float>>& queries, size_t k) const {
    Result<std::vector<NNResultsVector>> r;
    if (queries.empty()) return r;
    std::vector<float> flat;
    size_t dim;
    if (!detail::flatten(queries, &flat, &dim)) {
      r.error = {ErrorCode::InvalidArgument, "Query dimensionality mismatch"};
      return r;
    }
    size_t nq = queries.size();
    std::vector<uint32_t> ids(nq * k), counts(nq);
    std::vector<float> dists(nq * k);
    r.error = make_error(scann_treeah_search(h_, flat.data(), nq, dim, cfg_.partitions_to_search, cfg_.pre_reorder_k(k), k,
                                             ids.data(), dists.data(), counts.data(), nullptr, nullptr, nullptr,
                                             SCANN_HOST, nullptr));
    if (r.ok()) r.value = detail::unflatten(ids, dists, counts, k);
    return r;
  }
  // search(&[f32], k) (:240-242)
  Result<NNResultsVector> search(const std::vector<float>& query, size_t k) const {
    auto b = search_batched({query}, k);
    Result<NNResultsVector> r;
    r.error = b.error;
    if (b.ok() && !b.value.empty()) r.value = std::move(b.value[0]);
    return r;
  }
  // search_with_filter(&[f32], k, Some(filter)) (:245-250): `allowed[i]` = RestrictFilter::is_allowed(i)
  Result<NNResultsVector> search_with_filter(const std::vector<float>& query, size_t k,
                                             const std::vector<bool>& allowed) const {
    std::vector<uint8_t> bits((allowed.size() + 7) / 8, 0);
    for (size_t i = 0; i < allowed.size(); ++i)
      if (allowed[i]) bits[i >> 3] |= static_cast<uint8_t>(1u << (i & 7));
    Result<NNResultsVector> r;
    r.error = make_error(scann_treeah_set_filter(h_, bits.data(), allowed.size(), SCANN_HOST));
    if (!r.ok()) return r;
    r = search(query, k);
    scann_treeah_set_filter(h_, nullptr, 0, SCANN_HOST);
    return r;
  }
  const TreeXHybridConfig& config() const { return cfg_; }

 private:
  TreeXHybridConfig cfg_;
  scann_treeah* h_ = nullptr;
};

// The Scann facade's tree modes that score every member of the probed leaves (src/scann.rs:215-294):
//   search_partitioned <- Scann::search_partitioned (exact distances inside the L closest leaves)
//   search_tree_ah     <- Scann::search_tree_ah ("variant B": one f32 LookupTable of the query over byte codes)
class LeafScanSearcher {
 public:
  LeafScanSearcher() = default;
  LeafScanSearcher(LeafScanSearcher&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
  LeafScanSearcher(const LeafScanSearcher&) = delete;
  ~LeafScanSearcher() { scann_ivf_destroy(h_); }

  // centers [K*dim]; ids [n] grouped by partition; part_offsets [K+1]; raw [num_raw*stride] or null;
  // codebook [S*C*ds] + codes_by_id [num_raw*S] or both null
  ScannError build_from_index(const float* centers, size_t K, size_t dim, const uint32_t* ids,
                              const uint64_t* part_offsets, size_t n, const float* raw, size_t num_raw, size_t stride,
                              const float* codebook = nullptr, size_t S = 0, size_t C = 0,
                              const uint8_t* codes_by_id = nullptr, int device = 0) {
    scann_ivf_destroy(h_);
    h_ = nullptr;
    return make_error(scann_ivf_create(centers, K, dim, ids, part_offsets, n, raw, num_raw, stride, codebook, S, C,
                                       codes_by_id, device, SCANN_HOST, &h_));
  }
  Result<std::vector<NNResultsVector>> search_partitioned(const std::vector<std::vector<float>>& queries, size_t k,
                                                          size_t partitions_to_search,
                                                          DistanceMeasure m = DistanceMeasure::SquaredL2) const {
    return run(0, queries, k, partitions_to_search, m, -1);
  }
  // reorder >= 0: ReorderingHelper::reorder of the k results with that measure (scann.rs:198-209)
  Result<std::vector<NNResultsVector>> search_tree_ah(const std::vector<std::vector<float>>& queries, size_t k,
                                                      size_t partitions_to_search, int reorder = -1) const {
    return run(1, queries, k, partitions_to_search, DistanceMeasure::SquaredL2, reorder);
  }

 private:
  Result<std::vector<NNResultsVector>> run(int mode, const std::vector<std::vector<float>>& queries, size_t k, size_t L,
                                           DistanceMeasure m, int reorder) const {
    Result<std::vector<NNResultsVector>> r;
    if (queries.empty()) return r;
    std::vector<float> flat;
    size_t dim;
    if (!detail::flatten(queries, &flat, &dim)) {
      r.error = {ErrorCode::InvalidArgument, "Query dimensionality mismatch"};
      return r;
    }
    size_t nq = queries.size();
    std::vector<uint32_t> ids(nq * k), counts(nq);
    std::vector<float> dists(nq * k);
    r.error = make_error(scann_ivf_search(h_, mode, flat.data(), nq, dim, L, k, static_cast<int>(m), reorder, ids.data(),
                                          dists.data(), counts.data(), SCANN_HOST, nullptr));
    if (r.ok()) r.value = detail::unflatten(ids, dists, counts, k);
    return r;
  }
  scann_ivf* h_ = nullptr;
};

}  // namespace scann
