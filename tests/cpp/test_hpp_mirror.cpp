// Drives the C++ host mirror (include/scann_b200.hpp) with the reference's brute-force / partitioner unit
// tests (src/brute_force/searcher.rs:280-377, src/partitioning/tree_partitioner.rs:289-304).
// Built by tests/test_cpp_mirror.py with g++ against libscann_b200.so; run on a GPU box.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "scann_b200.hpp"

#define CHECK(cond)                                                      \
  do {                                                                   \
    if (!(cond)) {                                                       \
      std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      return 1;                                                          \
    }                                                                    \
  } while (0)

int main() {
  using namespace scann;
  const float ds5[5 * 3] = {0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 1, 1};
  auto bf = BruteForceSearcher::create(ds5, 5, 3, 3, DistanceMeasure::SquaredL2);
  CHECK(bf.ok());
  auto r = bf.value.search({0.f, 0.f, 0.f}, 3);  // test_brute_force_search
  CHECK(r.ok() && r.value.size() == 3 && r.value[0].first == 0 && std::fabs(r.value[0].second) < 1e-6f);
  auto all = bf.value.search({0.5f, 0.5f, 0.5f}, 5);  // test_brute_force_search_all
  CHECK(all.ok() && all.value.size() == 5);
  for (size_t i = 1; i < all.value.size(); ++i) CHECK(all.value[i].second >= all.value[i - 1].second);
  auto batch = bf.value.search_batched({{0.f, 0.f, 0.f}, {1.f, 1.f, 1.f}}, 2);  // test_brute_force_batched
  CHECK(batch.ok() && batch.value.size() == 2 && batch.value[0].size() == 2 && batch.value[1][0].first == 4);
  auto bad = bf.value.search({1.f, 2.f}, 5);  // test_brute_force_dimension_mismatch
  CHECK(!bad.ok() && bad.error.code == ErrorCode::InvalidArgument);
  auto clamp = bf.value.search({0.f, 0.f, 0.f}, 50);  // k clamped to n
  CHECK(clamp.ok() && clamp.value.size() == 5);
  auto empty = BruteForceSearcher::create(nullptr, 0, 3, 3, DistanceMeasure::SquaredL2);  // test_brute_force_empty_dataset
  CHECK(empty.ok());
  auto er = empty.value.search({1.f, 2.f, 3.f}, 5);
  CHECK(er.ok() && er.value.empty());

  auto rad = bf.value.search_radius({0.f, 0.f, 0.f}, 1.5f);  // test_brute_force_radius_search (searcher.rs:327-340)
  CHECK(rad.ok() && rad.value.size() == 4 && rad.value[0].first == 0);
  for (auto& pr : rad.value) CHECK(pr.second <= 1.5f);

  const float dot3[3 * 2] = {1, 0, 0, 1, 1, 1};
  auto dp = BruteForceSearcher::create(dot3, 3, 2, 2, DistanceMeasure::DotProduct);
  CHECK(dp.ok());
  auto dr = dp.value.search({1.f, 0.f}, 3);  // test_brute_force_search_dot_product
  CHECK(dr.ok() && dr.value.size() == 3 && dr.value[0].second == -1.0f && dr.value[2].second == 0.0f);

  const float centers[4 * 2] = {1, 0, 0, 1, 5, 5, 0, -1};
  auto part = TreePartitioner::from_centers(centers, 4, 2);
  CHECK(part.ok());
  auto pr = part.value.partition({0.f, 0.f}, 3);
  CHECK(pr.ok() && pr.value.tokens.size() == 3 && pr.value.tokens[0] == 0 && pr.value.tokens[1] == 1 &&
        pr.value.tokens[2] == 3);
  auto pall = part.value.partition({0.f, 0.f}, 10);  // more than K → K results
  CHECK(pall.ok() && pall.value.tokens.size() == 4);
  TreePartitioner unbuilt;
  auto pu = unbuilt.partition({0.f, 0.f}, 1);
  CHECK(!pu.ok() && pu.error.code == ErrorCode::FailedPrecondition);

  TreeXHybridConfig cfg;
  CHECK(cfg.pre_reorder_k(10) == 30);

  // Scann::search_partitioned through the mirror: 2 leaves, exact distances inside the closest one
  const float lc[2 * 2] = {0, 0, 10, 10};
  const float pts[6 * 2] = {0, 0, 1, 0, 0, 2, 10, 10, 11, 10, 9, 9};
  const uint32_t lids[6] = {0, 1, 2, 3, 4, 5};
  const uint64_t loff[3] = {0, 3, 6};
  LeafScanSearcher leaf;
  CHECK(leaf.build_from_index(lc, 2, 2, lids, loff, 6, pts, 6, 2).code == ErrorCode::Ok);
  auto lp = leaf.search_partitioned({{0.2f, 0.f}}, 5, 1);
  CHECK(lp.ok() && lp.value.size() == 1 && lp.value[0].size() == 3 && lp.value[0][0].first == 0 &&
        lp.value[0][1].first == 1 && lp.value[0][2].first == 2);
  auto lp2 = leaf.search_partitioned({{10.f, 10.f}}, 2, 2);
  CHECK(lp2.ok() && lp2.value[0].size() == 2 && lp2.value[0][0].first == 3 && lp2.value[0][0].second == 0.0f);
  auto lbad = leaf.search_tree_ah({{0.f, 0.f}}, 2, 1);  // no hasher in this index
  CHECK(!lbad.ok() && lbad.error.code == ErrorCode::FailedPrecondition);
  // ---- native index build + the Scann facade (src/scann.rs tests :438-520: build, search, sorted results) ----
  {
    const size_t n = 6000, dim = 32, nc = 12;
    std::vector<float> data(n * dim), cen(nc * dim);
    uint64_t st = 88172645463325252ull;
    auto rnd = [&]() {  // xorshift, uniform in [-1, 1)
      st ^= st << 13;
      st ^= st >> 7;
      st ^= st << 17;
      return static_cast<float>(static_cast<double>(st >> 11) / 9007199254740992.0 * 2.0 - 1.0);
    };
    for (auto& v : cen) v = 4.0f * rnd();
    for (size_t i = 0; i < n; ++i)
      for (size_t d = 0; d < dim; ++d) data[i * dim + d] = cen[(i % nc) * dim + d] + 0.3f * rnd();
    std::vector<std::vector<float>> qs;
    for (size_t q = 0; q < 40; ++q) {
      std::vector<float> v(dim);
      for (size_t d = 0; d < dim; ++d) v[d] = data[(q * 37) * dim + d] + 0.01f * rnd();
      qs.push_back(v);
    }
    auto exact = Scann::brute_force(data.data(), n, dim, dim);
    CHECK(exact.ok() && exact.value.search_mode() == SearchMode::BruteForce);
    auto truth = exact.value.search_batched(qs, 10);
    CHECK(truth.ok() && truth.value.size() == 40 && truth.value[0].size() == 10);
    auto recall = [&](const std::vector<NNResultsVector>& got) {
      size_t hit = 0;
      for (size_t q = 0; q < got.size(); ++q)
        for (auto& a : got[q])
          for (auto& b : truth.value[q]) hit += a.first == b.first;
      return static_cast<double>(hit) / (10.0 * got.size());
    };
    auto tree_ah = ScannBuilder().num_neighbors(10).tree(16, 6).hash(16).reorder(60).build(data.data(), n, dim, dim);
    CHECK(tree_ah.ok() && tree_ah.value.search_mode() == SearchMode::TreeAH);
    auto ta = tree_ah.value.search_batched(qs);
    CHECK(ta.ok() && ta.value.size() == 40);
    for (auto& r1 : ta.value) {
      CHECK(r1.size() == 10);
      for (size_t i = 1; i < r1.size(); ++i) CHECK(r1[i].second >= r1[i - 1].second);
    }
    CHECK(recall(ta.value) >= 0.9);
    auto parted = Scann::partitioned(data.data(), n, dim, dim, 16, 6);
    CHECK(parted.ok() && parted.value.search_mode() == SearchMode::Partitioned);
    auto pa = parted.value.search_batched(qs, 10);
    CHECK(pa.ok() && recall(pa.value) >= 0.95);
    auto hashed = ScannBuilder().hashed(16).reorder(800).build(data.data(), n, dim, dim);  // flat PQ: a mode (500 rows) is one code cell
    CHECK(hashed.ok() && hashed.value.search_mode() == SearchMode::Hashed);
    auto ha = hashed.value.search_batched(qs, 10);
    CHECK(ha.ok() && recall(ha.value) >= 0.9);
    auto none = Scann::brute_force(nullptr, 0, dim, dim);  // "Dataset cannot be empty"
    CHECK(!none.ok() && none.error.code == ErrorCode::InvalidArgument);

    AsymmetricHasherConfig hc;
    hc.num_subspaces = 16;
    AsymmetricHasher hasher(hc);
    CHECK(hasher.search(qs[0], 5).ok() && hasher.search(qs[0], 5).value.empty());  // unbuilt -> Ok(vec![])
    CHECK(hasher.build(data.data(), n, dim, dim).code == ErrorCode::Ok);
    auto hs = hasher.search(qs[0], 5);
    CHECK(hs.ok() && hs.value.size() == 5);
    auto hr = hasher.search_with_reordering(qs[0], 5, 800);
    CHECK(hr.ok() && hr.value.size() == 5 && hr.value[0].first == truth.value[0][0].first);
    CHECK(std::fabs(hr.value[0].second - truth.value[0][0].second) <= 1e-5f * (1.0f + truth.value[0][0].second));
    auto hb = hasher.search_batched(qs, 5);
    CHECK(hb.ok() && hb.value.size() == 40);

    TreeXHybridConfig tcfg;
    tcfg.num_partitions = 16;
    tcfg.partitions_to_search = 6;
    tcfg.num_subspaces = 16;
    TreeXHybridSearcher tx(tcfg);
    CHECK(tx.build(data.data(), n, dim, dim).code == ErrorCode::Ok);
    std::vector<SearchParameters> params(qs.size());
    for (size_t i = 0; i < params.size(); ++i) params[i].with_num_neighbors(i % 2 ? 5 : 10).with_pre_reordering_neighbors(50);
    auto wp = tx.search_batched_with_params(qs, params);
    CHECK(wp.ok() && wp.value.size() == qs.size() && wp.value[0].size() == 10 && wp.value[1].size() == 5);
    params.pop_back();
    auto wbad = tx.search_batched_with_params(qs, params);
    CHECK(!wbad.ok() && wbad.error.code == ErrorCode::InvalidArgument);
  }
  // ---- KMeansTree (src/trees/kmeans_tree.rs tests :425-441): build, size, sorted search_leaves ----
  {
    std::vector<float> pts;
    const float cx[3] = {0.f, 10.f, 0.f}, cy[3] = {0.f, 10.f, 10.f};
    for (int c = 0; c < 3; ++c)
      for (int i = 0; i < 20; ++i) {
        pts.push_back(cx[c] + i * 0.1f);
        pts.push_back(cy[c] + i * 0.05f);
      }
    KMeansTreeConfig kc;
    kc.num_children = 3;
    KMeansTree tree(kc);
    CHECK(tree.search_leaves({0.f, 0.f}, 2).ok() && tree.search_leaves({0.f, 0.f}, 2).value.empty());
    CHECK(tree.build(pts.data(), 60, 2, 2).code == ErrorCode::Ok);
    CHECK(tree.size() == 60 && tree.num_leaves() >= 1);
    auto hits = tree.search_leaves({0.f, 0.f}, 2);
    CHECK(hits.ok() && !hits.value.empty());
    for (size_t i = 1; i < hits.value.size(); ++i) CHECK(hits.value[i].distance >= hits.value[i - 1].distance);
    CHECK(tree.build(nullptr, 0, 2, 2).code == ErrorCode::InvalidArgument);
  }
  std::puts("hpp mirror ok");
  return 0;
}
