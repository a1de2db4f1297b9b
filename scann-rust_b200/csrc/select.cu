// Final exact re-score + ordering shared by the brute-force searchers, and the multi-GPU k-way merge.
#include "kernels.h"

namespace scann {

// One CTA per query.  cand: [nq][kc] row ids (0xFFFFFFFF = empty).  Every candidate row gets its exact
// distance in the reference's AVX2+FMA order (8 lanes per row, see exact_pair_distance), then the
// candidates are ordered by (distance, id) and the first k are written.
//   f32 rows: BruteForceSearcher::compute_distances  (src/brute_force/searcher.rs:113-139)
//   i8 rows : ScalarQuantizedBruteForceSearcher::compute_distances (scalar_quantized.rs:204-246)
// Order inside exact distance ties: ascending id.  (The reference's TopK evicts the largest id first
// among equal distances, top_k.rs:66-81, so it also keeps the lowest ids; the order in which
// drain_sorted lists tied entries is heap-array order and is not pinned by any reference test.)
template <bool I8>
__global__ void __launch_bounds__(256) rescore_topk_kernel(const RescoreParams rp, const uint32_t* __restrict__ cand,
                                                           int kc, int k, uint32_t* __restrict__ ids,
                                                           float* __restrict__ dists, uint32_t* __restrict__ counts) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int p2 = next_pow2(kc < 1 ? 1 : kc);
  uint64_t* keys = reinterpret_cast<uint64_t*>(sm);   // [p2]
  float* qs = reinterpret_cast<float*>(keys + p2);    // [dim]
  const int tid = threadIdx.x;
  const size_t q = blockIdx.x;
  const int dim = static_cast<int>(rp.dim);
  for (int d = tid; d < dim; d += 256) qs[d] = rp.queries[q * rp.dim + d];
  for (int j = tid; j < p2; j += 256) keys[j] = ~0ull;
  __syncthreads();
  const uint32_t* cq = cand + q * kc;
  const int grp = tid >> 3, sub = tid & 7;
  for (int j0 = 0; j0 < kc; j0 += 32) {
    int j = j0 + grp;
    uint32_t id = j < kc ? cq[j] : 0xFFFFFFFFu;
    bool valid = id != 0xFFFFFFFFu;
    size_t rid = valid ? id : 0;
    const void* row = I8 ? static_cast<const void*>(rp.raw_i8 + rid * rp.stride)
                         : static_cast<const void*>(rp.raw + rid * rp.stride);
    float d = exact_pair_distance<I8>(qs, row, dim, rp.measure, rp.scale, sub);
    if (valid && sub == 0) keys[j] = (static_cast<uint64_t>(f32_key(d)) << 32) | id;
  }
  __syncthreads();
  block_bitonic_sort<256>(keys, p2);
  int m = 0;
  for (int j = tid; j < k; j += 256) {
    uint64_t key = j < p2 ? keys[j] : ~0ull;
    bool ok = key != ~0ull;
    ids[q * k + j] = ok ? static_cast<uint32_t>(key & 0xFFFFFFFFu) : 0xFFFFFFFFu;
    dists[q * k + j] = ok ? key_f32(static_cast<uint32_t>(key >> 32)) : __int_as_float(0x7F800000);
  }
  if (tid == 0) {
    int lim = k < p2 ? k : p2;
    for (int j = 0; j < lim; ++j)
      if (keys[j] != ~0ull) ++m;
    counts[q] = static_cast<uint32_t>(m);
  }
}

scann_status launch_rescore_topk(const RescoreParams& rp, const uint32_t* cand, size_t nq, size_t kc, size_t k,
                                 uint32_t* ids, float* dists, uint32_t* counts, cudaStream_t s) {
  if (nq == 0) return SCANN_OK;
  SCANN_REQUIRE(kc <= 4096, SCANN_INVALID_ARGUMENT, "too many candidates per query (%zu)", kc);
  int p2 = next_pow2(kc < 1 ? 1 : static_cast<int>(kc));
  size_t smem = static_cast<size_t>(p2) * 8 + rp.dim * 4 + 16;
  if (rp.raw_i8) {
    if (smem > 48 * 1024)
      SCANN_CUDA(cudaFuncSetAttribute(rescore_topk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    rescore_topk_kernel<true><<<static_cast<unsigned>(nq), 256, smem, s>>>(rp, cand, static_cast<int>(kc),
                                                                          static_cast<int>(k), ids, dists, counts);
  } else {
    if (smem > 48 * 1024)
      SCANN_CUDA(cudaFuncSetAttribute(rescore_topk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    rescore_topk_kernel<false><<<static_cast<unsigned>(nq), 256, smem, s>>>(rp, cand, static_cast<int>(kc),
                                                                           static_cast<int>(k), ids, dists, counts);
  }
  SCANN_CUDA(cudaGetLastError());
  return SCANN_OK;
}

// Variable-length variant for the tensor-core ranking path (tc_gemm.cu FILTER lists): lists[q][0 .. min(cnt[q], cap))
// hold (ordered score key << 32 | row) entries in arbitrary order.  One CTA per query:
//   1. two-level certification (qn != nullptr): the list contains every row whose ranking score v is at most the
//      filter threshold, hence the k rows with the smallest v overall.  With v_(k) the k-th smallest v of the list and
//      eps the bound on |v - exact| (brute_force.cu), a row with v > v_(k) + 2*eps is beaten by those k rows whatever
//      the rounding did — only the rows within 2*eps of v_(k) are re-scored (k + a few tens instead of ~900);
//   2. exact distances of the survivors in the reference's AVX2 order, (distance, id) order, first k written.
template <bool I8>
__global__ void __launch_bounds__(256) rescore_lists_kernel(const RescoreParams rp,
                                                            const unsigned long long* __restrict__ lists,
                                                            const uint32_t* __restrict__ cnt, int cap, int k,
                                                            uint32_t n_rows, float radius, uint32_t* __restrict__ flag,
                                                            const float* __restrict__ qn, float xmax2,
                                                            uint32_t* __restrict__ ids, float* __restrict__ dists,
                                                            uint32_t* __restrict__ counts) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int tid = threadIdx.x;
  const size_t q = blockIdx.x;
  const int c = static_cast<int>(min(cnt[q], static_cast<uint32_t>(cap)));
  const int p2cap = next_pow2(cap < 1 ? 1 : cap);
  uint64_t* keys = reinterpret_cast<uint64_t*>(sm);              // [p2cap]
  uint32_t* sel = reinterpret_cast<uint32_t*>(keys + p2cap);     // [cap] rows to re-score
  uint32_t* hist = sel + cap;                                    // [264]
  float* qs = reinterpret_cast<float*>(hist + 264);              // [dim]
  const int dim = static_cast<int>(rp.dim);
  const float inf = __int_as_float(0x7F800000);
  for (int d = tid; d < dim; d += 256) qs[d] = rp.queries[q * rp.dim + d];
  const unsigned long long* lq = lists + q * static_cast<size_t>(cap);
  for (int j = tid; j < c; j += 256) {
    // padding rows of the operand tiles (row >= n_rows) can enter a list when thr = +inf or hx is skipped (Dot): they
    // must not take part in the k-th-score selection
    const unsigned long long e = lq[j];
    keys[j] = static_cast<uint32_t>(e & 0xFFFFFFFFull) < n_rows ? e : ~0ull;
  }
  __syncthreads();
  float thr2 = inf;
  if (qn != nullptr && c > k) {  // uniform over the CTA
    const uint64_t T = block_radix_threshold<256>(keys, c, k, hist);
    const float vk = key_f32(static_cast<uint32_t>(T >> 32));
    const float nqr = sqrtf(qn[q]), nx = sqrtf(xmax2);
    const float eps = tc_rank_eps(nqr, nx, I8);
    thr2 = vk + 2.0f * eps;
    thr2 = thr2 + fabsf(thr2) * 1e-6f;
    if (!(thr2 == thr2)) thr2 = inf;
  }
  if (tid == 0) hist[258] = 0;
  __syncthreads();
  for (int j = tid; j < c; j += 256) {
    const uint64_t key = keys[j];
    const uint32_t row = static_cast<uint32_t>(key & 0xFFFFFFFFull);
    const float v = key_f32(static_cast<uint32_t>(key >> 32));
    // rows >= n_rows: padding rows of the operand tiles can enter a list when thr = +inf or hx is skipped (Dot)
    if (row < n_rows && !(v > thr2)) sel[atomicAdd(&hist[258], 1u)] = row;
  }
  __syncthreads();
  const int m2 = static_cast<int>(hist[258]);
  const int p2 = next_pow2(m2 < 1 ? 1 : m2);
  for (int j = tid; j < p2; j += 256) keys[j] = ~0ull;
  __syncthreads();
  const int grp = tid >> 3, sub = tid & 7;
  for (int j0 = 0; j0 < m2; j0 += 32) {
    const int j = j0 + grp;
    const bool valid = j < m2;
    const uint32_t id = valid ? sel[j] : 0u;
    const void* row = I8 ? static_cast<const void*>(rp.raw_i8 + static_cast<size_t>(id) * rp.stride)
                         : static_cast<const void*>(rp.raw + static_cast<size_t>(id) * rp.stride);
    const float d = exact_pair_distance<I8>(qs, row, dim, rp.measure, rp.scale, sub);
    // radius search (searcher.rs:142-167): keep d <= radius only (radius = +inf for the top-k searches)
    const bool in_radius = radius == inf ? true : d <= radius;
    if (valid && sub == 0 && in_radius) keys[j] = (static_cast<uint64_t>(f32_key(d)) << 32) | id;
  }
  __syncthreads();
  block_bitonic_sort<256>(keys, p2);
  for (int j = tid; j < k; j += 256) {
    const uint64_t key = j < p2 ? keys[j] : ~0ull;
    const bool ok = key != ~0ull;
    ids[q * k + j] = ok ? static_cast<uint32_t>(key & 0xFFFFFFFFu) : 0xFFFFFFFFu;
    dists[q * k + j] = ok ? key_f32(static_cast<uint32_t>(key >> 32)) : inf;
  }
  if (tid == 0) {
    int m = 0;
    const int lim = k < p2 ? k : p2;
    for (int j = 0; j < lim; ++j)
      if (keys[j] != ~0ull) ++m;
    counts[q] = static_cast<uint32_t>(m);
    if (flag && k < p2 && keys[k] != ~0ull) atomicOr(flag, 2u);  // more than k entries qualify (radius search)
  }
}

scann_status launch_rescore_lists(const RescoreParams& rp, const unsigned long long* lists, const uint32_t* cnt,
                                  size_t nq, size_t cap, size_t k, size_t n_rows, uint32_t* ids, float* dists,
                                  uint32_t* counts, cudaStream_t s, float radius, uint32_t* flag, const float* qn,
                                  float xmax2) {
  if (nq == 0) return SCANN_OK;
  SCANN_REQUIRE(cap >= 1 && cap <= 8192, SCANN_INVALID_ARGUMENT, "candidate list capacity %zu out of range", cap);
  const size_t smem = static_cast<size_t>(next_pow2(static_cast<int>(cap))) * 8 + cap * 4 + 264 * 4 + rp.dim * 4 + 16;
  if (rp.raw_i8) {
    SCANN_CUDA(cudaFuncSetAttribute(rescore_lists_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(smem)));
    rescore_lists_kernel<true><<<static_cast<unsigned>(nq), 256, smem, s>>>(
        rp, lists, cnt, static_cast<int>(cap), static_cast<int>(k), static_cast<uint32_t>(n_rows), radius, flag, qn, xmax2,
        ids, dists, counts);
  } else {
    SCANN_CUDA(cudaFuncSetAttribute(rescore_lists_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(smem)));
    rescore_lists_kernel<false><<<static_cast<unsigned>(nq), 256, smem, s>>>(
        rp, lists, cnt, static_cast<int>(cap), static_cast<int>(k), static_cast<uint32_t>(n_rows), radius, flag, qn, xmax2,
        ids, dists, counts);
  }
  SCANN_CUDA(cudaGetLastError());
  return SCANN_OK;
}

// ---- multi-GPU merge: [parts][nq][k] -> [nq][k] by (distance, id) -----------------
__global__ void __launch_bounds__(128) merge_topk_kernel(const uint32_t* __restrict__ ids_in,
                                                         const float* __restrict__ dists_in, int parts, size_t nq,
                                                         size_t part_stride, int k, uint32_t* __restrict__ ids_out,
                                                         float* __restrict__ dists_out,
                                                         uint32_t* __restrict__ counts_out) {
  extern __shared__ __align__(16) uint8_t sm[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(sm);
  const int total = parts * k;
  const int p2 = next_pow2(total < 1 ? 1 : total);
  const size_t q = blockIdx.x;
  const int tid = threadIdx.x;
  for (int i = tid; i < p2; i += 128) {
    uint64_t key = ~0ull;
    if (i < total) {
      int part = i / k, j = i - part * k;
      size_t src = static_cast<size_t>(part) * part_stride + q * k + j;
      uint32_t id = ids_in[src];
      if (id != 0xFFFFFFFFu) key = (static_cast<uint64_t>(f32_key(dists_in[src])) << 32) | id;
    }
    keys[i] = key;
  }
  __syncthreads();
  block_bitonic_sort<128>(keys, p2);
  for (int j = tid; j < k; j += 128) {
    uint64_t key = keys[j];
    bool ok = key != ~0ull;
    ids_out[q * k + j] = ok ? static_cast<uint32_t>(key & 0xFFFFFFFFu) : 0xFFFFFFFFu;
    dists_out[q * k + j] = ok ? key_f32(static_cast<uint32_t>(key >> 32)) : __int_as_float(0x7F800000);
  }
  if (tid == 0 && counts_out) {
    int m = 0;
    for (int j = 0; j < k; ++j)
      if (keys[j] != ~0ull) ++m;
    counts_out[q] = static_cast<uint32_t>(m);
  }
}

scann_status launch_merge_topk(const uint32_t* ids_in, const float* dists_in, size_t parts, size_t nq, size_t k,
                               uint32_t* ids_out, float* dists_out, uint32_t* counts_out, cudaStream_t s,
                               size_t part_stride) {
  if (part_stride == 0) part_stride = nq * k;
  if (nq == 0 || k == 0) return SCANN_OK;
  SCANN_REQUIRE(parts * k <= 8192, SCANN_INVALID_ARGUMENT, "parts*k = %zu too large for the merge kernel", parts * k);
  int p2 = next_pow2(static_cast<int>(parts * k));
  size_t smem = static_cast<size_t>(p2) * 8;
  if (smem > 48 * 1024)
    SCANN_CUDA(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(smem)));
  merge_topk_kernel<<<static_cast<unsigned>(nq), 128, smem, s>>>(ids_in, dists_in, static_cast<int>(parts), nq,
                                                                 part_stride, static_cast<int>(k), ids_out, dists_out,
                                                                 counts_out);
  SCANN_CUDA(cudaGetLastError());
  return SCANN_OK;
}

}  // namespace scann

extern "C" {

scann_status scann_merge_topk(const uint32_t* ids_in, const float* dists_in, size_t parts, size_t nq, size_t k,
                              uint32_t* ids_out, float* dists_out, uint32_t* counts_out, int device, int memspace,
                              void* stream) {
  using namespace scann;
  if (nq == 0 || k == 0) return SCANN_OK;
  SCANN_REQUIRE(ids_in && dists_in && ids_out && dists_out && parts >= 1, SCANN_INVALID_ARGUMENT, "NULL buffer");
  SCANN_TRY(check_device(device));
  DeviceGuard g(device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (memspace == SCANN_DEVICE)
    return launch_merge_topk(ids_in, dists_in, parts, nq, k, ids_out, dists_out, counts_out, s);
  DevBuf<uint32_t> d_ii, d_io, d_c;
  DevBuf<float> d_di, d_do;
  SCANN_TRY(d_ii.upload(ids_in, parts * nq * k, SCANN_HOST, s));
  SCANN_TRY(d_di.upload(dists_in, parts * nq * k, SCANN_HOST, s));
  SCANN_TRY(d_io.alloc(nq * k));
  SCANN_TRY(d_do.alloc(nq * k));
  SCANN_TRY(d_c.alloc(nq));
  SCANN_TRY(launch_merge_topk(d_ii.p, d_di.p, parts, nq, k, d_io.p, d_do.p, d_c.p, s));
  SCANN_CUDA(cudaMemcpyAsync(ids_out, d_io.p, nq * k * 4, cudaMemcpyDeviceToHost, s));
  SCANN_CUDA(cudaMemcpyAsync(dists_out, d_do.p, nq * k * 4, cudaMemcpyDeviceToHost, s));
  if (counts_out) SCANN_CUDA(cudaMemcpyAsync(counts_out, d_c.p, nq * 4, cudaMemcpyDeviceToHost, s));
  SCANN_CUDA(cudaStreamSynchronize(s));
  return SCANN_OK;
}


scann_status scann_merge_topk_packed(const uint32_t* packed, size_t parts, size_t nq, size_t k, uint32_t* ids_out,
                                     float* dists_out, uint32_t* counts_out, int device, void* stream) {
  using namespace scann;
  if (nq == 0 || k == 0) return SCANN_OK;
  SCANN_REQUIRE(packed && ids_out && dists_out && parts >= 1, SCANN_INVALID_ARGUMENT, "NULL buffer");
  SCANN_TRY(check_device(device));
  DeviceGuard g(device);
  return launch_merge_topk(packed, reinterpret_cast<const float*>(packed + nq * k), parts, nq, k, ids_out, dists_out,
                           counts_out, static_cast<cudaStream_t>(stream), 2 * nq * k);
}

}  // extern "C"
