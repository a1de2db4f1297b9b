// TreePartitioner, query side (src/partitioning/tree_partitioner.rs:175-229).
//
// center_dist_kernel: squared-L2 from every query to every centre in the reference's exact order —
// sequential over dimensions, d = q - c, sum += d*d, never fused (compute_center_distances /
// squared_distance, :175-192).  The arithmetic is K*D*3 flops per query (C3: 0.58 MFLOP) so it runs on
// the CUDA cores bit-exactly instead of going through a tensor-core GEMM plus re-score; at a 10k
// batch it is ~1 % of the Tree-AH step.  Layout: centres transposed to [D][K] so consecutive threads
// read consecutive centres (coalesced), 8 queries per CTA staged in shared memory as [D][8] so one
// centre value is reused for 8 queries from two LDS.128.
//
// part_select_kernel: the reference sorts all K (dist, id) pairs with a stable sort and takes the
// first L (:206-222).  Stable order by OrderedFloat == ascending (dist, id), so an exact radix select
// of the L smallest 64-bit keys (f32_key(dist) << 32 | id) followed by a bitonic sort of the winners
// returns the identical token list.
#include "kernels.h"

namespace scann {

constexpr int kPartQ = 8;

__global__ void __launch_bounds__(256) center_dist_kernel(const float* __restrict__ centersT, int K, int dim,
                                                          const float* __restrict__ queries, int nq,
                                                          float* __restrict__ out) {
  extern __shared__ __align__(16) float qsT[];  // [dim][8]
  const int q0 = blockIdx.x * kPartQ;
  for (int i = threadIdx.x; i < dim * kPartQ; i += blockDim.x) {
    int d = i / kPartQ, qi = i % kPartQ;
    qsT[i] = (q0 + qi < nq) ? queries[static_cast<size_t>(q0 + qi) * dim + d] : 0.0f;
  }
  __syncthreads();
  const float4* qs4 = reinterpret_cast<const float4*>(qsT);
  for (int c = blockIdx.y * blockDim.x + threadIdx.x; c < K; c += gridDim.y * blockDim.x) {
    float acc[kPartQ];
#pragma unroll
    for (int i = 0; i < kPartQ; ++i) acc[i] = 0.0f;
    for (int d = 0; d < dim; ++d) {
      const float cv = __ldg(centersT + static_cast<size_t>(d) * K + c);
      const float4 a = qs4[d * 2], b = qs4[d * 2 + 1];
      const float qv[kPartQ] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < kPartQ; ++i) {
        float df = __fsub_rn(qv[i], cv);
        acc[i] = __fadd_rn(acc[i], __fmul_rn(df, df));
      }
    }
#pragma unroll
    for (int i = 0; i < kPartQ; ++i)
      if (q0 + i < nq) out[static_cast<size_t>(q0 + i) * K + c] = acc[i];
  }
}

constexpr int kSelChunk = 2048;

__global__ void __launch_bounds__(256) part_select_kernel(const float* __restrict__ dist, int K, int L,
                                                          uint32_t* __restrict__ tokens, float* __restrict__ dists) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int R = L < K ? L : K;
  const int p2 = next_pow2(R < 1 ? 1 : R);
  uint64_t* buf = reinterpret_cast<uint64_t*>(sm);
  uint64_t* out = buf + (R + kSelChunk);
  uint32_t* hist = reinterpret_cast<uint32_t*>(out + p2);
  const size_t q = blockIdx.x;
  const float* row = dist + q * K;
  auto gen = [&](int i) -> uint64_t { return (static_cast<uint64_t>(f32_key(row[i])) << 32) | static_cast<uint32_t>(i); };
  int m = block_topr_sorted<256, kSelChunk>(gen, K, R, buf, out, hist);
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    if (j < m) {
      uint64_t k = out[j];
      tokens[q * L + j] = static_cast<uint32_t>(k & 0xFFFFFFFFu);
      if (dists) dists[q * L + j] = key_f32(static_cast<uint32_t>(k >> 32));
    } else {
      tokens[q * L + j] = 0xFFFFFFFFu;
      if (dists) dists[q * L + j] = __int_as_float(0x7F800000);
    }
  }
}

__global__ void transpose_kernel(const float* __restrict__ in, size_t rows, size_t cols, float* __restrict__ out) {
  __shared__ float tile[32][33];
  size_t c0 = static_cast<size_t>(blockIdx.x) * 32, r0 = static_cast<size_t>(blockIdx.y) * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    size_t r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? in[r * cols + c] : 0.0f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    size_t c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[c * rows + r] = tile[threadIdx.x][j];
  }
}

void launch_transpose(const float* in, size_t rows, size_t cols, float* out, cudaStream_t s) {
  dim3 grid(static_cast<unsigned>((cols + 31) / 32), static_cast<unsigned>((rows + 31) / 32));
  transpose_kernel<<<grid, dim3(32, 8), 0, s>>>(in, rows, cols, out);
}

scann_status launch_partition(const float* centersT, size_t K, size_t dim, const float* queries, size_t nq, size_t L,
                              uint32_t* tokens, float* dists, float* scratch, cudaStream_t s) {
  if (nq == 0 || L == 0) return SCANN_OK;
  SCANN_REQUIRE(L <= 1024, SCANN_INVALID_ARGUMENT, "partitions_to_search %zu > 1024 unsupported", L);
  {
    unsigned gy = static_cast<unsigned>((K + 255) / 256);
    if (gy > 64) gy = 64;
    dim3 grid(static_cast<unsigned>((nq + kPartQ - 1) / kPartQ), gy);
    size_t smem = dim * kPartQ * sizeof(float);
    if (smem > 48 * 1024)
      SCANN_CUDA(cudaFuncSetAttribute(center_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    center_dist_kernel<<<grid, 256, smem, s>>>(centersT, static_cast<int>(K), static_cast<int>(dim), queries,
                                               static_cast<int>(nq), scratch);
    SCANN_CUDA(cudaGetLastError());
  }
  {
    int R = static_cast<int>(L < K ? L : K);
    int p2 = next_pow2(R < 1 ? 1 : R);
    size_t smem = (static_cast<size_t>(R) + kSelChunk + p2) * sizeof(uint64_t) + 264 * sizeof(uint32_t);
    part_select_kernel<<<static_cast<unsigned>(nq), 256, smem, s>>>(scratch, static_cast<int>(K), static_cast<int>(L),
                                                                    tokens, dists);
    SCANN_CUDA(cudaGetLastError());
  }
  return SCANN_OK;
}

}  // namespace scann

// ------------------------------------------------------------------------------------------ C ABI
struct scann_part {
  int device = 0;
  size_t K = 0, dim = 0;
  scann::DevBuf<float> centersT;
  scann::Workspace ws;
  std::mutex mu;
  cudaStream_t stream = nullptr;
};

extern "C" {

scann_status scann_part_create(const float* centers, size_t K, size_t dim, int device, int memspace,
                               scann_part** out) {
  using namespace scann;
  SCANN_REQUIRE(out != nullptr, SCANN_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  SCANN_REQUIRE(centers != nullptr && K > 0 && dim > 0, SCANN_INVALID_ARGUMENT,
                "Cannot partition empty dataset");  // tree_partitioner.rs:49-51
  SCANN_TRY(check_device(device));
  DeviceGuard g(device);
  scann_part* h = new scann_part();
  h->device = device;
  h->K = K;
  h->dim = dim;
  scann_status st = SCANN_OK;
  do {
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
      st = cuda_fail(cudaGetLastError(), "cudaStreamCreate", __FILE__, __LINE__);
      break;
    }
    DevBuf<float> tmp;
    if ((st = tmp.upload(centers, K * dim, memspace, h->stream)) != SCANN_OK) break;
    if ((st = h->centersT.alloc(K * dim)) != SCANN_OK) break;
    launch_transpose(tmp.p, K, dim, h->centersT.p, h->stream);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) {
      st = cuda_fail(cudaGetLastError(), "sync", __FILE__, __LINE__);
      break;
    }
  } while (0);
  if (st != SCANN_OK) {
    scann_part_destroy(h);
    return st;
  }
  *out = h;
  return SCANN_OK;
}

scann_status scann_part_select(scann_part* h, const float* queries, size_t nq, size_t qdim, size_t L,
                               uint32_t* tokens, float* dists, int memspace, void* stream) {
  using namespace scann;
  SCANN_REQUIRE(h != nullptr, SCANN_FAILED_PRECONDITION, "Partitioner not built");  // tree_partitioner.rs:197-198
  if (nq == 0 || L == 0) return SCANN_OK;
  SCANN_REQUIRE(queries && tokens, SCANN_INVALID_ARGUMENT, "NULL buffer");
  SCANN_REQUIRE(qdim == h->dim, SCANN_INVALID_ARGUMENT, "Query dimensionality %zu does not match dataset dimensionality %zu",
                qdim, h->dim);
  SCANN_REQUIRE(L <= 1024, SCANN_INVALID_ARGUMENT, "partitions_to_search %zu > 1024 unsupported", L);
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard g(h->device);
  cudaStream_t s = memspace == SCANN_DEVICE ? static_cast<cudaStream_t>(stream)
                                             : (stream ? static_cast<cudaStream_t>(stream) : h->stream);
  // bounded scratch: process the batch in chunks of queries
  size_t chunk = (size_t(256) << 20) / (h->K * sizeof(float));
  if (chunk < 1) chunk = 1;
  if (chunk > nq) chunk = nq;
  size_t need = Workspace::padded(chunk * h->K * sizeof(float));
  if (memspace == SCANN_HOST)
    need += Workspace::padded(chunk * h->dim * 4) + Workspace::padded(chunk * L * 4) * 2;
  SCANN_TRY(h->ws.reserve(need));
  float* scratch = h->ws.take<float>(chunk * h->K);
  float* dq = nullptr;
  uint32_t* dtok = nullptr;
  float* ddist = nullptr;
  if (memspace == SCANN_HOST) {
    dq = h->ws.take<float>(chunk * h->dim);
    dtok = h->ws.take<uint32_t>(chunk * L);
    ddist = h->ws.take<float>(chunk * L);
  }
  for (size_t q0 = 0; q0 < nq; q0 += chunk) {
    size_t nqc = nq - q0 < chunk ? nq - q0 : chunk;
    const float* qin = queries + q0 * h->dim;
    if (memspace == SCANN_HOST) {
      SCANN_CUDA(cudaMemcpyAsync(dq, qin, nqc * h->dim * 4, cudaMemcpyHostToDevice, s));
      SCANN_TRY(launch_partition(h->centersT.p, h->K, h->dim, dq, nqc, L, dtok, ddist, scratch, s));
      SCANN_CUDA(cudaMemcpyAsync(tokens + q0 * L, dtok, nqc * L * 4, cudaMemcpyDeviceToHost, s));
      if (dists) SCANN_CUDA(cudaMemcpyAsync(dists + q0 * L, ddist, nqc * L * 4, cudaMemcpyDeviceToHost, s));
      SCANN_CUDA(cudaStreamSynchronize(s));
    } else {
      SCANN_TRY(launch_partition(h->centersT.p, h->K, h->dim, qin, nqc, L, tokens + q0 * L,
                                 dists ? dists + q0 * L : nullptr, scratch, s));
    }
  }
  return SCANN_OK;
}

void scann_part_destroy(scann_part* h) {
  if (!h) return;
  {
    scann::DeviceGuard g(h->device);
    h->ws.release();
    h->centersT.free_();
    if (h->stream) cudaStreamDestroy(h->stream);
  }
  delete h;
}

}  // extern "C"
