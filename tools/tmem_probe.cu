// Stand-alone probe: how fast can warps read accumulators out of TENSOR MEMORY (tcgen05.ld 32x32b.x32), alone and while
// the tensor core is busy?  This is the epilogue bound of tc_score_kernel (tc_gemm.cu: a 256 x 128 f32 accumulator tile =
// 128 KB has to be read per 1024 tensor cycles at K = 128) and of tc_scan_kernel (tcscan.cu: 64 KB per 1600 cycles).
//   W      loader warps per CTA (4 / 8 / 16; warp w reads TMEM lane quadrant w & 3)
//   DEPTH  tcgen05.ld instructions in flight before tcgen05.wait::ld (1 / 2 / 4)
//   mma    0: loads only; 1: one thread issues tcgen05.mma kind::i8 M128 x N128 x K32 back to back into other columns
// One CTA per SM on every SM.  Prints bytes / clock / SM of TMEM reads and clocks per MMA under load (64 alone).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_probe.bin tools/tmem_probe.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (long long it = 0; !mbar_try(bar, parity); ++it)
    if (it > 100000000LL) {
      printf("mbarrier timeout\n");
      __trap();
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_i8_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return static_cast<uint64_t>((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// out[blockIdx.x] = {loader cycles (max over the warps), MMAs issued while the loaders ran, issuer cycles, sink}
template <int W, int DEPTH>
__global__ void __launch_bounds__(32 * (W + 1), 1) tmem_probe_kernel(int iters, int mma, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;          // 128 x 128 B (contents irrelevant)
  uint8_t* sB = smem + 16384;  // 128 x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 32768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  volatile uint32_t* done = tmem_slot + 1;
  unsigned long long* cyc_max = reinterpret_cast<unsigned long long*>(tmem_slot + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (tid == 0) {
    mbar_init(smem_u32(bars), 1);
    mbar_init(smem_u32(bars + 1), 1);
    *done = 0;
    *cyc_max = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < W) {
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t col0 = static_cast<uint32_t>(warp >> 2) * 32u;  // W = 16: four warps per quadrant, 32 columns apart
    uint32_t sink = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      uint32_t v[DEPTH][32];
#pragma unroll
      for (int d = 0; d < DEPTH; ++d)
        tc_ld32_nowait(tmem + lane_base + ((col0 + static_cast<uint32_t>(it * DEPTH + d) * 32u) & 255u), v[d]);
      tc_wait_ld();
#pragma unroll
      for (int d = 0; d < DEPTH; ++d)
#pragma unroll
        for (int j = 0; j < 32; ++j) sink ^= v[d][j];
    }
    const long long t1 = clock64();
    if (lane == 0) {
      atomicMax(cyc_max, static_cast<unsigned long long>(t1 - t0));
      atomicAdd(const_cast<uint32_t*>(done), 1u);
    }
    if (sink == 0x12345679u) out[4 * blockIdx.x + 3] = sink;
  } else if (lane == 0) {
    unsigned long long n_mma = 0;
    const long long t0 = clock64();
    if (mma) {
      const uint32_t idesc = (2u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
      const uint64_t ad = umma_desc(smem_u32(sA)), bd = umma_desc(smem_u32(sB));
      uint32_t batch = 0;
      while (*done < W) {
        if (batch >= 2) mbar_wait(smem_u32(bars + (batch & 1)), ((batch >> 1) - 1) & 1u);
        tc_fence_after();
        for (int r = 0; r < 4; ++r)
          for (int k4 = 0; k4 < 4; ++k4) mma_i8_ss(tmem + 256, ad + 2u * k4, bd + 2u * k4, idesc, 1u);
        tc_commit(smem_u32(bars + (batch & 1)));
        n_mma += 16;
        ++batch;
      }
      // drain
      if (batch >= 1) mbar_wait(smem_u32(bars + ((batch - 1) & 1)), ((batch - 1) >> 1) & 1u);
      if (batch >= 2) mbar_wait(smem_u32(bars + (batch & 1)), ((batch >> 1) - 1) & 1u);
    } else {
      while (*done < W) {
      }
    }
    const long long t1 = clock64();
    out[4 * blockIdx.x + 1] = n_mma;
    out[4 * blockIdx.x + 2] = static_cast<unsigned long long>(t1 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) out[4 * blockIdx.x + 0] = *cyc_max;
  if (warp == W) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at line %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

template <int W, int DEPTH>
int run(int grid, int mma, unsigned long long* d_out) {
  const int iters = 4096 / DEPTH;  // 4096 loads of 4 KB per warp
  const int smem = 1024 + 32768 + 128;
  CK(cudaFuncSetAttribute(tmem_probe_kernel<W, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaMemset(d_out, 0, grid * 32));
  tmem_probe_kernel<W, DEPTH><<<grid, 32 * (W + 1), smem>>>(iters, mma, d_out);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<unsigned long long> h(grid * 4);
  CK(cudaMemcpy(h.data(), d_out, grid * 32, cudaMemcpyDeviceToHost));
  double cyc = 0, mm = 0, icyc = 0;
  for (int b = 0; b < grid; ++b) {
    cyc += double(h[4 * b]);
    mm += double(h[4 * b + 1]);
    icyc += double(h[4 * b + 2]);
  }
  cyc /= grid;
  const double bytes = double(W) * 4096.0 * 4096.0;
  printf("W=%2d depth=%d mma=%d grid=%3d: %.1f B/clk/SM TMEM read (%.0f clk per 4 KB warp-load)", W, DEPTH, mma, grid,
         bytes / cyc, cyc / 4096.0);
  if (mma) printf(", %.1f clk per MMA under load (64 alone)", icyc / (mm > 0 ? mm : 1));
  printf("\n");
  return 0;
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  unsigned long long* d_out;
  CK(cudaMalloc(&d_out, 1024 * 32));
  for (int mma = 0; mma < 2; ++mma) {
    if (run<4, 1>(sms, mma, d_out)) return 1;
    if (run<4, 2>(sms, mma, d_out)) return 1;
    if (run<4, 4>(sms, mma, d_out)) return 1;
    if (run<8, 1>(sms, mma, d_out)) return 1;
    if (run<8, 2>(sms, mma, d_out)) return 1;
    if (run<8, 4>(sms, mma, d_out)) return 1;
    if (run<16, 1>(sms, mma, d_out)) return 1;
    if (run<16, 2>(sms, mma, d_out)) return 1;
  }
  if (run<8, 2>(1, 0, d_out)) return 1;
  return 0;
}
