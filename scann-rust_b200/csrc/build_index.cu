// Index build on the GPU: the training / build side of the searchers, so that the drop-in covers
// TreeXHybridSearcher::build (src/tree_x_hybrid/mod.rs:131-209), TreePartitioner::build
// (src/partitioning/tree_partitioner.rs:48-98 -> KMeans::fit, src/trees/kmeans.rs:166-432) and Codebook::train
// (src/hashes/codebook.rs:146-202) without leaving the library.
//
// What has reference semantics is done by the exact kernels of the search path: point -> partition assignment is
// TreePartitioner::partition(x, 1) (tensor-core scoring + exact sequential re-score, partition.cu), residual encode is
// Codebook::encode (taps.cu pq_encode_kernel), packing is PackedCodes4Bit::from_codes.  The k-means itself is OURS: the
// reference seeds k-means++ from an unpinned rand::StdRng (src/utils/random.rs:20-46), so its centroids cannot be
// reproduced by anybody; here it is Lloyd's algorithm from K distinct pseudo-random rows, with
//   assignment   the same exact nearest-centre kernel the searcher uses (lower id on ties),
//   update       per-cluster sums in f64 (atomicAdd on double: the f32 result does not depend on the summation order
//                beyond the last bit), empty clusters re-seeded from a hashed row.
// Everything runs on the legacy default stream of the device and synchronises before returning.
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "kernels.h"

namespace scann {

namespace {

constexpr float kBalanceRatio = 3.0f;  // partition centres of the in-library builds are trained balanced (kmeans_device)

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// rows[i] = src[perm(i)], perm(i) = (a * i + b) mod n with gcd(a, n) = 1: m distinct pseudo-random rows
__global__ void gather_affine_kernel(const float* __restrict__ src, size_t n, size_t dim, size_t stride, uint64_t a,
                                     uint64_t b, size_t m, float* __restrict__ dst) {
  const size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= m * dim) return;
  const size_t i = t / dim, d = t - i * dim;
  const size_t r = static_cast<size_t>((static_cast<unsigned __int128>(a) * i + b) % n);
  dst[t] = src[r * stride + d];
}

__global__ void compact_rows_kernel(const float* __restrict__ src, size_t n, size_t dim, size_t stride, size_t col0,
                                    size_t width, float* __restrict__ dst) {
  const size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n * width) return;
  const size_t i = t / width, d = t - i * width;
  dst[t] = src[i * stride + col0 + d];
}

// one warp per point: sums[assign][d] += x[d] (f64), cnt[assign] += 1
__global__ void __launch_bounds__(256) kmeans_accum_kernel(const float* __restrict__ x, size_t n, size_t dim,
                                                           const uint32_t* __restrict__ assign, size_t K,
                                                           double* __restrict__ sums, uint32_t* __restrict__ cnt) {
  const size_t i = static_cast<size_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const uint32_t c = assign[i];
  if (c >= K) return;
  for (size_t d = lane; d < dim; d += 32) atomicAdd(&sums[static_cast<size_t>(c) * dim + d], static_cast<double>(x[i * dim + d]));
  if (lane == 0) atomicAdd(&cnt[c], 1u);
}

__global__ void kmeans_update_kernel(const double* __restrict__ sums, const uint32_t* __restrict__ cnt, size_t K,
                                     size_t dim, const float* __restrict__ x, size_t n, uint64_t salt,
                                     float* __restrict__ centers) {
  const size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= K * dim) return;
  const size_t c = t / dim, d = t - c * dim;
  const uint32_t m = cnt[c];
  if (m > 0) centers[t] = static_cast<float>(sums[t] / static_cast<double>(m));
  else centers[t] = x[(splitmix64(salt + c) % n) * dim + d];  // empty cluster: re-seed from a hashed row
}

__global__ void residual_kernel(const float* __restrict__ x, size_t n, size_t dim, const float* __restrict__ centers,
                                const uint32_t* __restrict__ assign, float* __restrict__ out) {
  const size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n * dim) return;
  const size_t i = t / dim, d = t - i * dim;
  out[t] = __fsub_rn(x[t], centers[static_cast<size_t>(assign[i]) * dim + d]);  // tree_x_hybrid/mod.rs:181-185
}

// Nearest centre for a SMALL centre set (the 16 codewords of one PQ subspace): one thread per row, centres in shared
// memory, the partitioner's exact arithmetic (sequential d = x - c, sum += d*d, never fused; ties -> lower id), so the
// result equals TreePartitioner::partition(x, 1) without its per-row selection machinery.
__global__ void __launch_bounds__(256) nearest_small_kernel(const float* __restrict__ x, size_t n, int dim,
                                                            const float* __restrict__ centers, int K,
                                                            uint32_t* __restrict__ assign) {
  extern __shared__ float cs[];  // [K][dim]
  for (int i = threadIdx.x; i < K * dim; i += blockDim.x) cs[i] = centers[i];
  __syncthreads();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* row = x + i * dim;
  float best = __int_as_float(0x7F800000);
  uint32_t arg = 0;
  for (int c = 0; c < K; ++c) {
    float acc = 0.0f;
    for (int d = 0; d < dim; ++d) {
      const float df = __fsub_rn(row[d], cs[c * dim + d]);
      acc = __fadd_rn(acc, __fmul_rn(df, df));
    }
    if (acc < best || c == 0) {
      best = acc;
      arg = static_cast<uint32_t>(c);
    }
  }
  assign[i] = arg;
}

__global__ void iota_kernel(uint32_t* __restrict__ v, size_t n) {
  const size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t < n) v[t] = static_cast<uint32_t>(t);
}

__global__ void hist_kernel(const uint32_t* __restrict__ keys, size_t n, size_t K, unsigned long long* __restrict__ cnt) {
  const size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t < n && keys[t] < K) atomicAdd(&cnt[keys[t]], 1ull);
}

__global__ void gather_rows_kernel(const float* __restrict__ src, size_t dim, const uint32_t* __restrict__ rows, size_t m,
                                   float* __restrict__ dst) {
  const size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= m * dim) return;
  const size_t i = t / dim, d = t - i * dim;
  dst[t] = src[static_cast<size_t>(rows[i]) * dim + d];
}

__global__ void fill_kernel(float* __restrict__ v, size_t n, float value) {
  const size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t < n) v[t] = value;
}

inline unsigned grid_for(size_t n, unsigned block = 256) { return static_cast<unsigned>((n + block - 1) / block); }

uint64_t coprime_multiplier(uint64_t seed, uint64_t n) {
  if (n <= 1) return 1;
  for (uint64_t t = 0;; ++t) {
    uint64_t a = (splitmix64(seed + t) % n) | 1ull;
    uint64_t g = a, h = n;
    while (h) {
      const uint64_t r = g % h;
      g = h;
      h = r;
    }
    if (g == 1) return a;
  }
}

// nearest centre of every row (TreePartitioner::partition(x, 1).tokens[0]); x contiguous device [n][dim].
// *part: created on first use, refreshed with the current centres afterwards (scann_part_update), destroyed by the caller.
scann_status assign_rows(scann_part** part, const float* x, size_t n, size_t dim, const float* centers, size_t K,
                         uint32_t* assign, int device) {
  SCANN_CUDA(cudaDeviceSynchronize());  // centres / rows were produced on the default stream; the partitioner has its own
  if (*part == nullptr) SCANN_TRY(scann_part_create(centers, K, dim, device, SCANN_DEVICE, part));
  else SCANN_TRY(scann_part_update(*part, centers, SCANN_DEVICE));
  scann_status st = scann_part_select(*part, x, n, dim, 1, assign, nullptr, SCANN_DEVICE, nullptr);
  if (st == SCANN_OK && cudaDeviceSynchronize() != cudaSuccess) st = cuda_fail(cudaGetLastError(), "assign", __FILE__, __LINE__);
  return st;
}
scann_status assign_rows(const float* x, size_t n, size_t dim, const float* centers, size_t K, uint32_t* assign, int device) {
  scann_part* part = nullptr;
  scann_status st = assign_rows(&part, x, n, dim, centers, K, assign, device);
  scann_part_destroy(part);
  return st;
}
struct PartOwner {  // destroys the partitioner on every exit path
  scann_part* p = nullptr;
  ~PartOwner() { scann_part_destroy(p); }
};

struct LloydScratch {  // sized once for the largest (n, K) of a training run
  DevBuf<uint32_t> assign, cnt;
  DevBuf<double> sums;
  scann_status reserve(size_t n, size_t K, size_t dim) {
    if (assign.n < n) SCANN_TRY(assign.alloc(n));
    if (cnt.n < K) SCANN_TRY(cnt.alloc(K));
    if (sums.n < K * dim) SCANN_TRY(sums.alloc(K * dim));
    return SCANN_OK;
  }
};

// Lloyd's algorithm on contiguous device rows from K distinct pseudo-random rows; centers: device [K][dim] (K <= n).
// sc.assign holds the assignment of the LAST iteration's assignment step (i.e. against the centres before the last
// update) on return.
scann_status lloyd_device(const float* x, size_t n, size_t dim, size_t K, int iters, uint64_t seed, float* centers,
                          int device, LloydScratch& sc) {
  const uint64_t a = coprime_multiplier(seed, n), b = splitmix64(seed ^ 0xA5A5A5A5ull) % n;
  gather_affine_kernel<<<grid_for(K * dim), 256>>>(x, n, dim, dim, a, b, K, centers);
  SCANN_CUDA(cudaGetLastError());
  SCANN_TRY(sc.reserve(n, K, dim));
  PartOwner part;
  const bool small = K <= 64 && K * dim <= 4096;
  for (int it = 0; it < iters; ++it) {
    if (small) {
      nearest_small_kernel<<<grid_for(n), 256, K * dim * sizeof(float)>>>(x, n, static_cast<int>(dim), centers,
                                                                          static_cast<int>(K), sc.assign.p);
      SCANN_CUDA(cudaGetLastError());
    } else {
      SCANN_TRY(assign_rows(&part.p, x, n, dim, centers, K, sc.assign.p, device));
    }
    SCANN_CUDA(cudaMemsetAsync(sc.sums.p, 0, K * dim * sizeof(double)));
    SCANN_CUDA(cudaMemsetAsync(sc.cnt.p, 0, K * sizeof(uint32_t)));
    kmeans_accum_kernel<<<grid_for(n, 8), 256>>>(x, n, dim, sc.assign.p, K, sc.sums.p, sc.cnt.p);
    kmeans_update_kernel<<<grid_for(K * dim), 256>>>(sc.sums.p, sc.cnt.p, K, dim, x, n,
                                                     splitmix64(seed + 1000003ull * (it + 1)), centers);
    SCANN_CUDA(cudaGetLastError());
  }
  return SCANN_OK;
}

// nearest centre against the CURRENT centres (the partitioner's exact arithmetic either way)
scann_status nearest_device(const float* x, size_t n, size_t dim, const float* centers, size_t K, uint32_t* assign,
                            int device) {
  if (K <= 64 && K * dim <= 4096) {
    nearest_small_kernel<<<grid_for(n), 256, K * dim * sizeof(float)>>>(x, n, static_cast<int>(dim), centers,
                                                                        static_cast<int>(K), assign);
    SCANN_CUDA(cudaGetLastError());
    return SCANN_OK;
  }
  return assign_rows(x, n, dim, centers, K, assign, device);
}

// k-means on contiguous device rows; centers: device [K][dim] (K <= n).
// balance_ratio <= 1: plain Lloyd.  balance_ratio > 1: BALANCED training — Lloyd places K - K/8 centres, then the
// heaviest cluster is bisected (2-means on its own rows) again and again until K centres exist.  High-dimensional data
// has hubs (a centre near the bulk of the data collects the tails of many modes); Lloyd never leaves that optimum and
// re-seeding inside a hub only shaves it, while bisection always halves the current worst leaf.  On the C5 data model
// the largest leaf falls from 13-16x to ~2.5x the mean (it must stay below the scan's in-leaf position limit).
scann_status kmeans_device(const float* x, size_t n, size_t dim, size_t K, int iters, uint64_t seed, float balance_ratio,
                           float* centers, int device) {
  LloydScratch sc;
  const size_t reserve = K / 8;
  if (!(balance_ratio > 1.0f) || reserve == 0 || n / K < 4) {
    SCANN_TRY(lloyd_device(x, n, dim, K, iters, seed, centers, device, sc));
    SCANN_CUDA(cudaDeviceSynchronize());
    return SCANN_OK;
  }
  const size_t K0 = K - reserve;
  SCANN_TRY(lloyd_device(x, n, dim, K0, iters, seed, centers, device, sc));
  SCANN_TRY(nearest_device(x, n, dim, centers, K0, sc.assign.p, device));
  std::vector<uint32_t> h_assign(n);
  SCANN_CUDA(cudaMemcpy(h_assign.data(), sc.assign.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  std::vector<std::vector<uint32_t>> members(K);
  for (size_t i = 0; i < n; ++i)
    if (h_assign[i] < K0) members[h_assign[i]].push_back(static_cast<uint32_t>(i));
  // max-heap on (rows, cluster); a cluster that cannot be split any more re-enters with weight 0
  std::vector<std::pair<size_t, uint32_t>> heap;
  for (size_t c = 0; c < K0; ++c) heap.emplace_back(members[c].size(), static_cast<uint32_t>(c));
  std::make_heap(heap.begin(), heap.end());
  size_t biggest = 0;
  for (size_t c = 0; c < K0; ++c) biggest = std::max(biggest, members[c].size());
  DevBuf<uint32_t> d_idx;
  DevBuf<float> d_rows, d_two;
  SCANN_TRY(d_idx.alloc(std::max<size_t>(biggest, 2)));
  SCANN_TRY(d_rows.alloc(std::max<size_t>(biggest, 2) * dim));
  SCANN_TRY(d_two.alloc(2 * dim));
  LloydScratch sc2;
  std::vector<uint32_t> side;
  size_t have = K0;
  while (have < K && !heap.empty() && heap.front().first >= 2) {
    std::pop_heap(heap.begin(), heap.end());
    const uint32_t c = heap.back().second;
    heap.pop_back();
    std::vector<uint32_t>& mem = members[c];
    const size_t m = mem.size();
    SCANN_CUDA(cudaMemcpy(d_idx.p, mem.data(), m * sizeof(uint32_t), cudaMemcpyHostToDevice));
    gather_rows_kernel<<<grid_for(m * dim), 256>>>(x, dim, d_idx.p, m, d_rows.p);
    SCANN_CUDA(cudaGetLastError());
    SCANN_TRY(lloyd_device(d_rows.p, m, dim, 2, 6, seed + 7919ull * have, d_two.p, device, sc2));
    SCANN_TRY(nearest_device(d_rows.p, m, dim, d_two.p, 2, sc2.assign.p, device));
    side.resize(m);
    SCANN_CUDA(cudaMemcpy(side.data(), sc2.assign.p, m * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    std::vector<uint32_t> m0, m1;
    for (size_t i = 0; i < m; ++i) (side[i] == 0 ? m0 : m1).push_back(mem[i]);
    if (m0.empty() || m1.empty()) {  // identical rows: not splittable
      heap.emplace_back(0, c);
      std::push_heap(heap.begin(), heap.end());
      continue;
    }
    SCANN_CUDA(cudaMemcpy(centers + static_cast<size_t>(c) * dim, d_two.p, dim * sizeof(float), cudaMemcpyDeviceToDevice));
    SCANN_CUDA(cudaMemcpy(centers + have * dim, d_two.p + dim, dim * sizeof(float), cudaMemcpyDeviceToDevice));
    members[have] = std::move(m1);
    mem = std::move(m0);
    heap.emplace_back(mem.size(), c);
    std::push_heap(heap.begin(), heap.end());
    heap.emplace_back(members[have].size(), static_cast<uint32_t>(have));
    std::push_heap(heap.begin(), heap.end());
    ++have;
  }
  if (have < K) {  // nothing left to split (duplicate rows): the remaining centres sit on hashed rows
    const uint64_t a = coprime_multiplier(seed ^ 0xF111ull, n), b = splitmix64(seed ^ 0xF222ull) % n;
    gather_affine_kernel<<<grid_for((K - have) * dim), 256>>>(x, n, dim, dim, a, b, K - have, centers + have * dim);
    SCANN_CUDA(cudaGetLastError());
  }
  SCANN_CUDA(cudaDeviceSynchronize());
  return SCANN_OK;
}

// rows of the caller (host or device, any stride) -> contiguous device [n][dim]; *out points at the caller's memory
// when it already is that
scann_status contiguous_device_rows(const float* x, size_t n, size_t dim, size_t stride, int memspace, DevBuf<float>* own,
                                    const float** out) {
  if (memspace == SCANN_DEVICE && stride == dim) {
    *out = x;
    return SCANN_OK;
  }
  SCANN_TRY(own->alloc(n * dim));
  if (memspace == SCANN_DEVICE) {
    compact_rows_kernel<<<grid_for(n * dim), 256>>>(x, n, dim, stride, 0, dim, own->p);
    SCANN_CUDA(cudaGetLastError());
  } else {
    SCANN_CUDA(cudaMemcpy2D(own->p, dim * sizeof(float), x, stride * sizeof(float), dim * sizeof(float), n,
                            cudaMemcpyHostToDevice));
  }
  *out = own->p;
  return SCANN_OK;
}

// Codebook::train on contiguous device rows (already residuals when the index uses them); codebook: device [S][16][ds]
scann_status pq_train_device(const float* rows, size_t n, size_t dim, size_t S, int iters, uint64_t seed, float* codebook,
                             int device) {
  const size_t ds = dim / S, codes = std::min<size_t>(16, n);
  DevBuf<float> sub;
  SCANN_TRY(sub.alloc(n * ds));
  if (codes < 16) {  // tiny datasets: the unused codewords sit far away and are never chosen
    fill_kernel<<<grid_for(S * 16 * ds), 256>>>(codebook, S * 16 * ds, 1e6f);
    SCANN_CUDA(cudaGetLastError());
  }
  for (size_t s = 0; s < S; ++s) {
    compact_rows_kernel<<<grid_for(n * ds), 256>>>(rows, n, ds, dim, s * ds, ds, sub.p);
    SCANN_CUDA(cudaGetLastError());
    SCANN_TRY(kmeans_device(sub.p, n, ds, codes, iters, seed + s, 0.0f, codebook + s * 16 * ds, device));  // seed + s: codebook.rs:177
  }
  return SCANN_OK;
}

}  // namespace

// kmtree.cu builds its nodes with the same trainer: centres of `x` (contiguous device rows) and, when assign_out is not
// null, the final nearest-centre assignment against those centres
scann_status kmeans_rows_device(const float* x, size_t n, size_t dim, size_t K, int iters, uint64_t seed,
                                float balance_ratio, float* centers, uint32_t* assign_out, int device) {
  SCANN_TRY(kmeans_device(x, n, dim, K, iters, seed, balance_ratio, centers, device));
  if (assign_out) SCANN_TRY(assign_rows(x, n, dim, centers, K, assign_out, device));
  return SCANN_OK;
}

}  // namespace scann

extern "C" {

scann_status scann_kmeans_fit(const float* x, size_t n, size_t dim, size_t stride, size_t K, int iters, uint64_t seed,
                              float balance_ratio, float* centers, int device, int memspace) {
  using namespace scann;
  SCANN_REQUIRE(x && centers, SCANN_INVALID_ARGUMENT, "NULL buffer");
  SCANN_REQUIRE(n > 0 && dim > 0 && stride >= dim, SCANN_INVALID_ARGUMENT, "Cannot cluster empty dataset");
  SCANN_REQUIRE(K >= 1 && K <= n, SCANN_INVALID_ARGUMENT, "num_clusters %zu outside 1..%zu", K, n);
  SCANN_REQUIRE(iters >= 0, SCANN_INVALID_ARGUMENT, "negative iteration count");
  SCANN_TRY(check_device(device));
  DeviceGuard g(device);
  DevBuf<float> own, d_centers;
  const float* rows = nullptr;
  SCANN_TRY(contiguous_device_rows(x, n, dim, stride, memspace, &own, &rows));
  float* c = centers;
  if (memspace == SCANN_HOST) {
    SCANN_TRY(d_centers.alloc(K * dim));
    c = d_centers.p;
  }
  SCANN_TRY(kmeans_device(rows, n, dim, K, iters, seed, balance_ratio, c, device));
  if (memspace == SCANN_HOST) SCANN_CUDA(cudaMemcpy(centers, c, K * dim * sizeof(float), cudaMemcpyDeviceToHost));
  return SCANN_OK;
}

scann_status scann_pq_train(const float* x, size_t n, size_t dim, size_t stride, const float* centers,
                            const uint32_t* assign, size_t num_centers, size_t S, int iters, uint64_t seed, float* codebook,
                            int device, int memspace) {
  using namespace scann;
  SCANN_REQUIRE(x && codebook, SCANN_INVALID_ARGUMENT, "NULL buffer");
  SCANN_REQUIRE(n > 0 && dim > 0 && stride >= dim, SCANN_INVALID_ARGUMENT, "Cannot train on empty dataset");
  SCANN_REQUIRE(S >= 1 && dim % S == 0, SCANN_INVALID_ARGUMENT, "Dimensionality %zu must be divisible by num_subspaces %zu",
                dim, S);  // codebook.rs:154-159
  SCANN_REQUIRE((centers == nullptr) == (assign == nullptr), SCANN_INVALID_ARGUMENT, "centers and assign go together");
  SCANN_TRY(check_device(device));
  DeviceGuard g(device);
  DevBuf<float> own, resid, d_cb, d_cen;
  DevBuf<uint32_t> d_as;
  const float* rows = nullptr;
  SCANN_TRY(contiguous_device_rows(x, n, dim, stride, memspace, &own, &rows));
  if (centers) {
    const float* cen = centers;
    const uint32_t* as = assign;
    if (memspace == SCANN_HOST) {
      SCANN_TRY(d_cen.upload(centers, num_centers * dim, SCANN_HOST, 0));
      SCANN_TRY(d_as.upload(assign, n, SCANN_HOST, 0));
      cen = d_cen.p;
      as = d_as.p;
    }
    SCANN_TRY(resid.alloc(n * dim));
    residual_kernel<<<grid_for(n * dim), 256>>>(rows, n, dim, cen, as, resid.p);
    SCANN_CUDA(cudaGetLastError());
    rows = resid.p;
  }
  float* cb = codebook;
  if (memspace == SCANN_HOST) {
    SCANN_TRY(d_cb.alloc(S * 16 * (dim / S)));
    cb = d_cb.p;
  }
  SCANN_TRY(pq_train_device(rows, n, dim, S, iters, seed, cb, device));
  SCANN_CUDA(cudaDeviceSynchronize());
  if (memspace == SCANN_HOST) SCANN_CUDA(cudaMemcpy(codebook, cb, S * 16 * (dim / S) * sizeof(float), cudaMemcpyDeviceToHost));
  return SCANN_OK;
}

scann_status scann_treeah_build(const float* x, size_t n, size_t dim, size_t stride, size_t K, size_t S, size_t train_rows,
                                int kmeans_iters, uint64_t seed, int use_residuals, int reorder_measure, int keep_raw,
                                int device, int memspace, scann_treeah** out) {
  using namespace scann;
  SCANN_REQUIRE(out != nullptr, SCANN_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  SCANN_REQUIRE(x != nullptr && n > 0 && dim > 0 && stride >= dim, SCANN_INVALID_ARGUMENT,
                "Cannot build from empty dataset");  // tree_x_hybrid/mod.rs:132-134
  SCANN_REQUIRE(K >= 1, SCANN_INVALID_ARGUMENT, "num_partitions must be >= 1");
  SCANN_REQUIRE(S >= 1 && S <= 256 && dim % S == 0, SCANN_INVALID_ARGUMENT,
                "Dimensionality %zu must be divisible by num_subspaces %zu", dim, S);
  SCANN_REQUIRE(n < 0xFFFFFFFFull, SCANN_INVALID_ARGUMENT, "index too large for u32 ids");
  SCANN_TRY(check_device(device));
  DeviceGuard g(device);
  DevBuf<float> own, sample, centers, codebook, resid, rows_sorted;
  DevBuf<uint32_t> assign, assign_s, ids, assign_sorted, ids_sorted;
  DevBuf<uint8_t> packed, cub_tmp;
  DevBuf<unsigned long long> cnt, d_off;
  const float* rows = nullptr;
  SCANN_TRY(contiguous_device_rows(x, n, dim, stride, memspace, &own, &rows));
  // 1. training sample (distinct pseudo-random rows), partition centres, residual codebook
  const size_t m = std::min(train_rows > 0 ? train_rows : n, n);
  const float* srows = rows;
  if (m < n) {
    SCANN_TRY(sample.alloc(m * dim));
    gather_affine_kernel<<<grid_for(m * dim), 256>>>(rows, n, dim, dim, coprime_multiplier(seed ^ 0x5EEDull, n),
                                                     splitmix64(seed ^ 0x1234ull) % n, m, sample.p);
    SCANN_CUDA(cudaGetLastError());
    srows = sample.p;
  }
  const size_t Keff = std::min(K, m);
  SCANN_TRY(centers.alloc(Keff * dim));
  SCANN_TRY(kmeans_device(srows, m, dim, Keff, kmeans_iters, seed, kBalanceRatio, centers.p, device));
  SCANN_TRY(codebook.alloc(S * 16 * (dim / S)));
  const float* trows = srows;
  if (use_residuals) {
    SCANN_TRY(assign_s.alloc(m));
    SCANN_TRY(assign_rows(srows, m, dim, centers.p, Keff, assign_s.p, device));
    SCANN_TRY(resid.alloc(m * dim));
    residual_kernel<<<grid_for(m * dim), 256>>>(srows, m, dim, centers.p, assign_s.p, resid.p);
    SCANN_CUDA(cudaGetLastError());
    trows = resid.p;
  }
  SCANN_TRY(pq_train_device(trows, m, dim, S, kmeans_iters, 42, codebook.p, device));
  resid.free_();
  sample.free_();
  assign_s.free_();
  // 2. every row -> its partition (TreePartitioner::partition(x, 1)), rows grouped by partition (stable: ascending id)
  SCANN_TRY(assign.alloc(n));
  SCANN_TRY(assign_rows(rows, n, dim, centers.p, Keff, assign.p, device));
  SCANN_TRY(ids.alloc(n));
  SCANN_TRY(assign_sorted.alloc(n));
  SCANN_TRY(ids_sorted.alloc(n));
  iota_kernel<<<grid_for(n), 256>>>(ids.p, n);
  int bits = 1;
  while ((size_t(1) << bits) < Keff) ++bits;
  size_t tmp_bytes = 0;
  SCANN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, assign.p, assign_sorted.p, ids.p, ids_sorted.p,
                                             static_cast<int>(n), 0, bits));
  SCANN_TRY(cub_tmp.alloc(tmp_bytes));
  SCANN_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp_bytes, assign.p, assign_sorted.p, ids.p, ids_sorted.p,
                                             static_cast<int>(n), 0, bits));
  SCANN_TRY(cnt.alloc(Keff));
  SCANN_CUDA(cudaMemset(cnt.p, 0, Keff * sizeof(unsigned long long)));
  hist_kernel<<<grid_for(n), 256>>>(assign.p, n, Keff, cnt.p);
  SCANN_CUDA(cudaGetLastError());
  std::vector<unsigned long long> hcnt(Keff);
  SCANN_CUDA(cudaMemcpy(hcnt.data(), cnt.p, Keff * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  std::vector<uint64_t> off(Keff + 1, 0);
  for (size_t l = 0; l < Keff; ++l) off[l + 1] = off[l] + hcnt[l];
  SCANN_TRY(d_off.upload(reinterpret_cast<const unsigned long long*>(off.data()), Keff + 1, SCANN_HOST, 0));
  ids.free_();
  assign.free_();
  cub_tmp.free_();
  // 3. residual PQ codes of the rows in index order (Codebook::encode + PackedCodes4Bit::from_codes), in chunks
  const size_t bpp = (S + 1) / 2, chunk = std::min<size_t>(n, size_t(1) << 21);
  SCANN_TRY(packed.alloc(n * bpp));
  SCANN_TRY(rows_sorted.alloc(chunk * dim));
  for (size_t r0 = 0; r0 < n; r0 += chunk) {
    const size_t mc = std::min(chunk, n - r0);
    gather_rows_kernel<<<grid_for(mc * dim), 256>>>(rows, dim, ids_sorted.p + r0, mc, rows_sorted.p);
    SCANN_CUDA(cudaGetLastError());
    SCANN_TRY(scann_pq_encode(codebook.p, S, dim / S, rows_sorted.p, mc, dim, use_residuals ? centers.p : nullptr,
                              use_residuals ? assign_sorted.p + r0 : nullptr, packed.p + r0 * bpp, device, SCANN_DEVICE));
  }
  rows_sorted.free_();
  // 4. the searcher
  return scann_treeah_create_ex(centers.p, Keff, dim, codebook.p, S, packed.p, ids_sorted.p,
                                reinterpret_cast<const uint64_t*>(d_off.p), n, keep_raw ? rows : nullptr, keep_raw ? n : 0,
                                dim, use_residuals, reorder_measure, 0u, device, SCANN_DEVICE, out);
}

// Scann::init_partitioning (src/scann.rs:140-152) on the GPU: k-means centres over the dataset, every row assigned by
// TreePartitioner::partition(x, 1), partition_indices = rows grouped by partition in ascending id -> the searcher of
// Scann::search_partitioned (scann_ivf_search mode 0).
scann_status scann_ivf_build(const float* x, size_t n, size_t dim, size_t stride, size_t K, int kmeans_iters, uint64_t seed,
                             int device, int memspace, scann_ivf** out) {
  using namespace scann;
  SCANN_REQUIRE(out != nullptr, SCANN_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  SCANN_REQUIRE(x != nullptr && n > 0 && dim > 0 && stride >= dim, SCANN_INVALID_ARGUMENT, "Dataset cannot be empty");
  SCANN_REQUIRE(K >= 1 && n < 0xFFFFFFFFull, SCANN_INVALID_ARGUMENT, "bad num_partitions / dataset size");
  SCANN_TRY(check_device(device));
  DeviceGuard g(device);
  DevBuf<float> own, sample, centers;
  DevBuf<uint32_t> assign, ids, assign_sorted, ids_sorted;
  DevBuf<uint8_t> cub_tmp;
  DevBuf<unsigned long long> cnt, d_off;
  const float* rows = nullptr;
  SCANN_TRY(contiguous_device_rows(x, n, dim, stride, memspace, &own, &rows));
  const size_t m = std::min<size_t>(n, 1000000);
  const float* srows = rows;
  if (m < n) {
    SCANN_TRY(sample.alloc(m * dim));
    gather_affine_kernel<<<grid_for(m * dim), 256>>>(rows, n, dim, dim, coprime_multiplier(seed ^ 0x5EEDull, n),
                                                     splitmix64(seed ^ 0x1234ull) % n, m, sample.p);
    SCANN_CUDA(cudaGetLastError());
    srows = sample.p;
  }
  const size_t Keff = std::min(K, m);
  SCANN_TRY(centers.alloc(Keff * dim));
  SCANN_TRY(kmeans_device(srows, m, dim, Keff, kmeans_iters, seed, kBalanceRatio, centers.p, device));
  sample.free_();
  SCANN_TRY(assign.alloc(n));
  SCANN_TRY(assign_rows(rows, n, dim, centers.p, Keff, assign.p, device));
  SCANN_TRY(ids.alloc(n));
  SCANN_TRY(assign_sorted.alloc(n));
  SCANN_TRY(ids_sorted.alloc(n));
  iota_kernel<<<grid_for(n), 256>>>(ids.p, n);
  int bits = 1;
  while ((size_t(1) << bits) < Keff) ++bits;
  size_t tmp_bytes = 0;
  SCANN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, assign.p, assign_sorted.p, ids.p, ids_sorted.p,
                                             static_cast<int>(n), 0, bits));
  SCANN_TRY(cub_tmp.alloc(tmp_bytes));
  SCANN_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp_bytes, assign.p, assign_sorted.p, ids.p, ids_sorted.p,
                                             static_cast<int>(n), 0, bits));
  SCANN_TRY(cnt.alloc(Keff));
  SCANN_CUDA(cudaMemset(cnt.p, 0, Keff * sizeof(unsigned long long)));
  hist_kernel<<<grid_for(n), 256>>>(assign.p, n, Keff, cnt.p);
  SCANN_CUDA(cudaGetLastError());
  std::vector<unsigned long long> hcnt(Keff);
  SCANN_CUDA(cudaMemcpy(hcnt.data(), cnt.p, Keff * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  std::vector<uint64_t> off(Keff + 1, 0);
  for (size_t l = 0; l < Keff; ++l) off[l + 1] = off[l] + hcnt[l];
  SCANN_TRY(d_off.upload(reinterpret_cast<const unsigned long long*>(off.data()), Keff + 1, SCANN_HOST, 0));
  SCANN_CUDA(cudaDeviceSynchronize());
  return scann_ivf_create(centers.p, Keff, dim, ids_sorted.p, reinterpret_cast<const uint64_t*>(d_off.p), n, rows, n, dim,
                          nullptr, 0, 0, nullptr, device, SCANN_DEVICE, out);
}

}  // extern "C"
