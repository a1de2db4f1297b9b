// Query x row contraction on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
// Replaces the inner loops of
//   BruteForceSearcher::compute_distances            (src/brute_force/searcher.rs:113-139, x86.rs:195-346)
//   ScalarQuantizedBruteForceSearcher::compute_distances (src/brute_force/scalar_quantized.rs:204-246)
//   TreePartitioner::compute_center_distances        (src/partitioning/tree_partitioner.rs:175-194)
// as a RANKING device only: the kernel produces the approximate score
//        v(q, x) = hx[x] - q~ . x~          q~, x~ = operands rounded to bf16, f32 accumulation in TMEM
// (hx = |x|^2 / 2 for the L2 family so that v = (|q - x|^2 - |q|^2) / 2, hx = 0 for Dot so that v = -q.x).  The
// callers keep the best k + margin rows by v, re-score those exactly in the reference's own summation order
// and certify the margin against the bf16 error bound (brute_force.cu); no value computed here is returned.
//
// Kernel shape (one persistent CTA per SM, 576 threads):
//   warp 0      TMA producer   queries -> resident A tiles, rows -> STAGES-deep ring of B tiles
//                              (cp.async.bulk.tensor.2d, 128-byte swizzle, mbarrier complete_tx)
//   warp 1      MMA issuer     one elected lane: tcgen05.mma.cta_group::1.kind::f16, M = 128, N = 128, K = 16;
//                              MT query tiles share every B tile; accumulators double-buffered in TMEM;
//                              tcgen05.commit releases the B slot and publishes the accumulator
//   warps 2-17  epilogue       tcgen05.ld 32x32b.x32 (thread = one query row), then either
//                              DENSE : v -> out[q][row]                                  (first rows / centroids)
//                              FILTER: v <= thr[q] -> append (score key, row) to the query's candidate list
// A work unit is (query super-tile of MT*128 queries, run of row tiles).  Units that run at the same time
// share the row run, so the rows are read from HBM once and from L2 by the other CTAs.
#include <cuda.h>
#include <cuda_bf16.h>

#include <stdlib.h>

#include <algorithm>

#include "kernels.h"

namespace scann {

namespace {

constexpr int kTcBM = 128;       // queries per MMA = TMEM lanes
constexpr int kTcBN = 128;       // rows per B tile = accumulator columns
constexpr int kTcAtomK = 64;     // bf16 elements per 128-byte swizzle atom
constexpr int kTcAtomBytes = kTcBM * 128;  // one (128 rows x 128 B) atom tile
constexpr int kTcEpiWarps = 16;  // 4 per TMEM lane quadrant: the epilogue is latency-bound, warps hide it
constexpr int kTcThreads = 32 * (2 + kTcEpiWarps);
constexpr int kTcTmemCols = 512;
// FILTER staging: every epilogue warp queues its survivors (row, lane) in shared memory and flushes the queue with
// one global atomic per entry, all lanes at once — a survivor costs a shared-memory atomic instead of a serialised
// round trip to L2 (at ~1 survivor per 1024 scores nearly every 32x32 chunk has one).
constexpr int kWqLane = 8;                           // queue slots of ONE lane (= one query) of an epilogue warp
constexpr int kWqFlushAt = 6;                        // the warp writes its queues out when some lane holds this many
constexpr int kWqBytes = kWqLane * 32 * 8 + 16;      // [slot][lane] (row, raw w bits) (+ pad)

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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
// shared-memory queue store by 32-bit shared address (a pointer that has been through a non-inlined call is generic to
// the compiler, and generic stores are what it then emits)
__device__ __forceinline__ void sts_v2(uint32_t saddr, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(saddr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {  // one lane of the (converged) warp
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}

// K-major operand tile in the canonical 128-byte-swizzle layout (rows of 128 B, 8-row groups 1024 B apart):
// start address >> 4, LBO (ignored for swizzled K-major) = 1, SBO = 1024 >> 4, descriptor version 1 (sm_100),
// layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return static_cast<uint64_t>((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, N = 128, M = 128
constexpr uint32_t kIdescBf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((kTcBN >> 3) << 17) | ((kTcBM >> 4) << 24);

struct TcArgs {
  uint32_t nq;              // valid queries (rows of A beyond nq are zero padding)
  uint32_t q_tiles;         // query super-tiles of MT * 128
  uint32_t row0;            // first row of this launch (multiple of 128)
  uint32_t n_tiles;         // 128-row tiles in this launch
  uint32_t tile_stride;     // tile t covers rows row0 + t * tile_stride * 128 .. + 128 (1 = contiguous; > 1 = sample)
  uint32_t tiles_per_unit;  // row tiles per work unit
  uint32_t n_units;         // ceil(n_tiles / tiles_per_unit)
  const float* hx;          // [>= row0 + n_tiles * 128]
  int filter;               // 0 = DENSE, 1 = FILTER
  float* dense;             // DENSE: [nq][ld], column = row - row0
  size_t ld;
  const float* thr;         // FILTER: [nq] keep v <= thr
  unsigned long long* cand; // FILTER: [nq][cap] (ordered score key << 32 | row) in arbitrary order
  uint32_t cap;
  uint32_t* cand_cnt;       // FILTER: [nq] appended (may exceed cap = overflow)
  int no_hx;                // FILTER with hx == 0 for every real row (Dot): v = -acc, the hx loads and subtractions are
                            // skipped (padding rows then score 0 and may enter a list: the re-score ignores ids >= n)
};

// FILTER survivors.  A lane of an epilogue warp is one query, and it queues its hits in slots of its OWN (slot-major
// shared memory, no atomics, nothing warp-wide: a hit costs a compare, an address and a store).  Once per tile the warp
// votes; when some lane holds kWqFlushAt entries every lane reserves its list slots with ONE atomic and writes its
// entries out.  Both routines are single copies, called: inlined at every use they were a fifth of the epilogue's code.
__device__ __forceinline__ unsigned long long tc_list_entry(uint32_t w_bits, uint32_t row) {
  return (static_cast<unsigned long long>(f32_key(-__uint_as_float(w_bits))) << 32) | row;  // v = -w
}
__device__ __noinline__ void tc_flush(uint32_t* __restrict__ cand_cnt, unsigned long long* __restrict__ cand,
                                      uint32_t cap, uint32_t q, uint32_t wq_lane_s, uint32_t n) {
  if (n == 0) return;
  const uint32_t base = atomicAdd(cand_cnt + q, n);
  for (uint32_t i = 0; i < n && base + i < cap; ++i) {
    uint32_t row, bits;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(row), "=r"(bits) : "r"(wq_lane_s + 256u * i) : "memory");
    cand[static_cast<size_t>(q) * cap + base + i] = tc_list_entry(bits, row);
  }
}
// a lane's slots are full (more than kWqLane - kWqFlushAt hits of one query in one tile): straight to the query's list
__device__ __noinline__ void tc_emit_direct(uint32_t* __restrict__ cand_cnt, unsigned long long* __restrict__ cand,
                                            uint32_t cap, uint32_t q, uint32_t w_bits, uint32_t row) {
  const uint32_t slot = atomicAdd(cand_cnt + q, 1u);
  if (slot < cap) cand[static_cast<size_t>(q) * cap + slot] = tc_list_entry(w_bits, row);
}

template <int MT, int KA, int STAGES>
__global__ void __launch_bounds__(kTcThreads, 1)
    tc_score_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128-byte swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                   // [MT][KA] atom tiles
  uint8_t* sB = sA + MT * KA * kTcAtomBytes;            // [STAGES][KA] atom tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + STAGES * KA * kTcAtomBytes);
  // barrier indices
  constexpr int kFull = 0, kEmpty = STAGES, kAFull = 2 * STAGES, kAEmpty = 2 * STAGES + 1, kTFull = 2 * STAGES + 2,
                kTEmpty = 2 * STAGES + 4, kNumBars = 2 * STAGES + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);
  uint8_t* wq_base = reinterpret_cast<uint8_t*>(tmem_slot + 4);  // [kTcEpiWarps][kWqBytes]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * static_cast<uint32_t>(i); };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(bar(kFull + i), 1);
      mbar_init(bar(kEmpty + i), 1);
    }
    mbar_init(bar(kAFull), 1);
    mbar_init(bar(kAEmpty), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(kTFull + i), 1);
      mbar_init(bar(kTEmpty + i), kTcEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTcTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t total_units = a.q_tiles * a.n_units;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (elect_one()) {
      uint32_t stage = 0, phase = 0, aphase = 0;
      for (uint32_t u = blockIdx.x; u < total_units; u += gridDim.x) {
        const uint32_t qt = u % a.q_tiles, nu = u / a.q_tiles;
        const uint32_t t0 = nu * a.tiles_per_unit;
        const uint32_t t1 = min(t0 + a.tiles_per_unit, a.n_tiles);
        mbar_wait(bar(kAEmpty), aphase ^ 1);  // MMAs of the previous unit have finished reading A
        mbar_expect_tx(bar(kAFull), MT * KA * kTcAtomBytes);
        for (int mt = 0; mt < MT; ++mt)
          for (int ka = 0; ka < KA; ++ka)
            tma_load_2d(smem_u32(sA + (mt * KA + ka) * kTcAtomBytes), &tmA, bar(kAFull), ka * kTcAtomK,
                        static_cast<int>((qt * MT + mt) * kTcBM));
        aphase ^= 1;
        for (uint32_t t = t0; t < t1; ++t) {
          mbar_wait(bar(kEmpty + stage), phase ^ 1);
          mbar_expect_tx(bar(kFull + stage), KA * kTcAtomBytes);
          for (int ka = 0; ka < KA; ++ka)
            tma_load_2d(smem_u32(sB + (stage * KA + ka) * kTcAtomBytes), &tmB, bar(kFull + stage), ka * kTcAtomK,
                        static_cast<int>(a.row0 + t * a.tile_stride * kTcBN));
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    // elect.sync instead of "lane == 0": the compiler then knows that a single thread runs the loop and keeps the
    // descriptors on the uniform datapath — the 16 UTCHMMA of a tile back to back.  With "lane == 0" every tcgen05.mma
    // was wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop (167 of them in the object).
    if (elect_one()) {
      uint32_t stage = 0, phase = 0, aphase = 0, as = 0, asphase = 0;
      for (uint32_t u = blockIdx.x; u < total_units; u += gridDim.x) {
        const uint32_t nu = u / a.q_tiles;
        const uint32_t t0 = nu * a.tiles_per_unit;
        const uint32_t t1 = min(t0 + a.tiles_per_unit, a.n_tiles);
        mbar_wait(bar(kAFull), aphase);
        aphase ^= 1;
        for (uint32_t t = t0; t < t1; ++t) {
          mbar_wait(bar(kTEmpty + as), asphase ^ 1);  // epilogue has drained this accumulator
          mbar_wait(bar(kFull + stage), phase);       // TMA has landed this B tile
          tc_fence_after();
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint32_t d = tmem_base + (as * MT + mt) * kTcBN;
#pragma unroll
            for (int ka = 0; ka < KA; ++ka) {
              const uint64_t ad = umma_desc(smem_u32(sA + (mt * KA + ka) * kTcAtomBytes));
              const uint64_t bd = umma_desc(smem_u32(sB + (stage * KA + ka) * kTcAtomBytes));
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4)  // 16 bf16 = 32 bytes per MMA: +2 in the (address >> 4) field
                tc_mma_bf16(d, ad + 2u * k4, bd + 2u * k4, kIdescBf16, (ka | k4) != 0 ? 1u : 0u);
            }
          }
          tc_commit(bar(kEmpty + stage));  // B slot free once these MMAs retire
          tc_commit(bar(kTFull + as));     // accumulator ready
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
          if (++as == 2) {
            as = 0;
            asphase ^= 1;
          }
        }
        tc_commit(bar(kAEmpty));  // A may be overwritten once everything issued so far has retired
      }
    }
  } else {
    // ===================================================================== epilogue warps
    const int e = warp - 2;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const int grp = e >> 2;     // 0..3: MT = 2 -> (query tile, column half); MT = 1 -> column quarter
    const int mt = MT == 2 ? grp >> 1 : 0;
    constexpr int ncols = MT == 2 ? kTcBN / 2 : kTcBN / 4;
    const int c0 = (MT == 2 ? (grp & 1) : grp) * ncols;
    constexpr int nchunks = ncols / 32;
    const uint32_t wq_lane_s = smem_u32(wq_base + e * kWqBytes) + 8u * static_cast<uint32_t>(lane);  // slot i: + 256 i
    uint32_t wq_n = 0;  // entries this lane holds
    uint32_t as = 0, asphase = 0;
    for (uint32_t u = blockIdx.x; u < total_units; u += gridDim.x) {
      const uint32_t qt = u % a.q_tiles, nu = u / a.q_tiles;
      const uint32_t t0 = nu * a.tiles_per_unit;
      const uint32_t t1 = min(t0 + a.tiles_per_unit, a.n_tiles);
      const uint32_t q = (qt * MT + mt) * kTcBM + quad * 32 + lane;
      const bool qvalid = q < a.nq;
      float nthr = __int_as_float(0x7F800000);  // +inf: nothing passes
      if (a.filter && qvalid) nthr = -a.thr[q];
      for (uint32_t t = t0; t < t1; ++t) {
        if (!a.no_hx && t + 1 < t1 && lane < ncols / 32)  // pull the next tile's hx lines into L1 while this tile is filtered
          asm volatile("prefetch.global.L1 [%0];" ::"l"(a.hx + a.row0 + (t + 1) * a.tile_stride * kTcBN + c0 + lane * 32));
        mbar_wait(bar(kTFull + as), asphase);
        tc_fence_after();
        const uint32_t row_tile = a.row0 + t * a.tile_stride * kTcBN + c0;  // first row of this warp's columns
        const uint32_t col_tile = t * kTcBN + c0;                            // DENSE output column of that row
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + (as * MT + mt) * kTcBN + c0;
#pragma unroll 1
        for (int ch = 0; ch < nchunks; ++ch) {
          const int cc = ch * 32;
          uint32_t vr[32];
          tc_ld32(taddr + cc, vr);
          // w = acc - hx = -v for every score of the chunk (IEEE subtraction is symmetric: -(acc - hx) is bit for bit
          // hx - acc), so one copy of the filter code serves Dot (hx == 0: w is the raw accumulator) and the L2 family
          float w[32];
          if (a.no_hx) {
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) w[j] = __uint_as_float(vr[j]);
          } else {
            const float4* h4 = reinterpret_cast<const float4*>(a.hx + row_tile + cc);
            float h[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 x = __ldg(h4 + j);
              h[4 * j + 0] = x.x;
              h[4 * j + 1] = x.y;
              h[4 * j + 2] = x.z;
              h[4 * j + 3] = x.w;
            }
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) w[j] = __fsub_rn(__uint_as_float(vr[j]), h[j]);
          }
          if (ch + 1 == nchunks) {
            // the whole accumulator is in registers: hand it back to the MMA warp before the last chunk's work
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(kTEmpty + as));
          }
          if (!a.filter) {
            if (qvalid) {
              float4* o = reinterpret_cast<float4*>(a.dense + static_cast<size_t>(q) * a.ld + col_tile + cc);
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] = make_float4(-w[4 * j], -w[4 * j + 1], -w[4 * j + 2], -w[4 * j + 3]);
            }
            continue;
          }
          // FILTER: v <= thr is w >= -thr.  A max tree per 8 columns; only the lanes with a hit (about one group in
          // three has one, at C2) leave the straight path, and nothing in it is a warp-wide step.
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            float m = w[8 * g8];
#pragma unroll
            for (int j = 1; j < 8; ++j) m = fmaxf(m, w[8 * g8 + j]);
            if (m >= nthr) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (w[8 * g8 + j] >= nthr) {
                  const uint32_t row = row_tile + cc + 8 * g8 + j;
                  if (wq_n < static_cast<uint32_t>(kWqLane)) {
                    sts_v2(wq_lane_s + 256u * wq_n, row, __float_as_uint(w[8 * g8 + j]));
                    ++wq_n;
                  } else {
                    tc_emit_direct(a.cand_cnt, a.cand, a.cap, q, __float_as_uint(w[8 * g8 + j]), row);
                  }
                }
              }
            }
          }
        }
        if (a.filter && __any_sync(0xFFFFFFFFu, wq_n >= static_cast<uint32_t>(kWqFlushAt))) {
          tc_flush(a.cand_cnt, a.cand, a.cap, q, wq_lane_s, wq_n);
          wq_n = 0;
        }
        if (++as == 2) {
          as = 0;
          asphase ^= 1;
        }
      }
      if (a.filter) {  // the lane -> query mapping changes with the unit
        tc_flush(a.cand_cnt, a.cand, a.cap, q, wq_lane_s, wq_n);
        wq_n = 0;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTcTmemCols) : "memory");
  }
}

// ---- operand preparation -------------------------------------------------------------------------
// f32 / i8 rows -> bf16 [rows_pad][Kpad] (zero padded) + hx.  One warp per row.
//   hx = 0.5 * |x|^2 of the values the exact kernels see (f32 rows; (i8) * scale for SQ8) for the L2 family,
//   0 for Dot, +inf for padding rows (never selected).
template <bool I8>
__global__ void __launch_bounds__(256) tc_prep_rows_kernel(const void* __restrict__ src, size_t n, size_t dim,
                                                           size_t stride, size_t kpad, size_t rows_pad, float scale,
                                                           int want_norm, __nv_bfloat16* __restrict__ dst,
                                                           float* __restrict__ hx, float* __restrict__ norm_max_key) {
  const size_t r = static_cast<size_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows_pad) return;
  float s = 0.0f;
  for (size_t d = lane; d < kpad; d += 32) {
    float v = 0.0f;
    if (r < n && d < dim)
      v = I8 ? static_cast<float>(static_cast<const int8_t*>(src)[r * stride + d])
             : static_cast<const float*>(src)[r * stride + d];
    dst[r * kpad + d] = __float2bfloat16_rn(v);  // i8 values are exact in bf16
    const float f = I8 ? v * scale : v;
    s = fmaf(f, f, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  if (lane == 0) {
    hx[r] = r < n ? (want_norm ? 0.5f * s : 0.0f) : __int_as_float(0x7F800000);
    if (r < n && norm_max_key) atomicMax(reinterpret_cast<uint32_t*>(norm_max_key), __float_as_uint(s));  // s >= 0
  }
}

// f32 queries -> bf16 [q_pad][Kpad] + |q|^2.  One warp per query.
__global__ void __launch_bounds__(256) tc_prep_queries_kernel(const float* __restrict__ q, size_t nq, size_t dim,
                                                              size_t kpad, size_t q_pad, float scale,
                                                              __nv_bfloat16* __restrict__ dst, float* __restrict__ qn) {
  const size_t r = static_cast<size_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= q_pad) return;
  float s = 0.0f;
  for (size_t d = lane; d < kpad; d += 32) {
    const float v = (r < nq && d < dim) ? q[r * dim + d] : 0.0f;
    dst[r * kpad + d] = __float2bfloat16_rn(v * scale);  // SQ8: the row scale rides on the query
    s = fmaf(v, v, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  if (lane == 0) qn[r] = s;
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// bf16 [rows][kpad] row-major; box = 64 elements (128 B) x 128 rows; 128-byte swizzle; OOB rows read as zero
scann_status make_map(CUtensorMap* map, const void* base, size_t rows, size_t kpad) {
  EncodeTiledFn fn = encode_fn();
  SCANN_REQUIRE(fn != nullptr, SCANN_UNAVAILABLE, "cuTensorMapEncodeTiled is not available in this driver");
  cuuint64_t dims[2] = {kpad, rows};
  cuuint64_t strides[1] = {kpad * 2};
  cuuint32_t box[2] = {kTcAtomK, kTcBM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SCANN_REQUIRE(r == CUDA_SUCCESS, SCANN_INTERNAL, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  return SCANN_OK;
}

}  // namespace

// u8 [rows][row_bytes] row-major (row_bytes a multiple of 16); box = 128 bytes x 128 rows; 128-byte swizzle; columns /
// rows beyond the tensor read as zero (tcscan.cu: the per-group LUT tiles)
scann_status tc_make_map_u8(void* map, const void* base, size_t rows, size_t row_bytes) {
  EncodeTiledFn fn = encode_fn();
  SCANN_REQUIRE(fn != nullptr, SCANN_UNAVAILABLE, "cuTensorMapEncodeTiled is not available in this driver");
  cuuint64_t dims[2] = {row_bytes, rows};
  cuuint64_t strides[1] = {row_bytes};
  cuuint32_t box[2] = {128, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(static_cast<CUtensorMap*>(map), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SCANN_REQUIRE(r == CUDA_SUCCESS, SCANN_INTERNAL, "cuTensorMapEncodeTiled(u8) failed (%d)", static_cast<int>(r));
  return SCANN_OK;
}

namespace {

template <int MT, int KA>
scann_status launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& a, int sms, cudaStream_t s) {
  constexpr int kBudget = 225 * 1024 - kTcEpiWarps * kWqBytes;  // operand tiles
  constexpr int kStagesRaw = (kBudget - MT * KA * kTcAtomBytes) / (KA * kTcAtomBytes);
  constexpr int STAGES = kStagesRaw > 6 ? 6 : kStagesRaw;
  static_assert(STAGES >= 2, "operand tiles do not fit");
  const size_t smem = 1024 + static_cast<size_t>(MT + STAGES) * KA * kTcAtomBytes + (2 * STAGES + 6) * 8 + 16 +
                      kTcEpiWarps * kWqBytes;
  auto kern = tc_score_kernel<MT, KA, STAGES>;
  SCANN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const uint32_t total = a.q_tiles * a.n_units;
  const unsigned grid = static_cast<unsigned>(std::min<uint32_t>(total, static_cast<uint32_t>(sms)));
  kern<<<grid, kTcThreads, smem, s>>>(tmA, tmB, a);
  SCANN_CUDA(cudaGetLastError());
  return SCANN_OK;
}

}  // namespace

size_t tc_kpad(size_t dim) { return (dim + kTcAtomK - 1) / kTcAtomK * kTcAtomK; }
bool tc_supported(size_t dim) { return dim >= 1 && tc_kpad(dim) <= 256; }
size_t tc_rows_pad(size_t n) { return (n + kTcBN - 1) / kTcBN * kTcBN; }
size_t tc_queries_pad(size_t nq, size_t dim) {
  const size_t m = (tc_kpad(dim) <= 128 ? 2 : 1) * kTcBM;
  return (nq + m - 1) / m * m;
}

scann_status tc_prepare_rows(const void* rows, bool i8, size_t n, size_t dim, size_t stride, float scale,
                             bool want_norm, void* dst_bf16, float* hx, float* norm_max, cudaStream_t s) {
  const size_t kpad = tc_kpad(dim), rows_pad = tc_rows_pad(n);
  if (rows_pad == 0) return SCANN_OK;
  if (norm_max) SCANN_CUDA(cudaMemsetAsync(norm_max, 0, sizeof(float), s));
  const unsigned grid = static_cast<unsigned>((rows_pad + 7) / 8);
  if (i8)
    tc_prep_rows_kernel<true><<<grid, 256, 0, s>>>(rows, n, dim, stride, kpad, rows_pad, scale, want_norm ? 1 : 0,
                                                   static_cast<__nv_bfloat16*>(dst_bf16), hx, norm_max);
  else
    tc_prep_rows_kernel<false><<<grid, 256, 0, s>>>(rows, n, dim, stride, kpad, rows_pad, scale, want_norm ? 1 : 0,
                                                    static_cast<__nv_bfloat16*>(dst_bf16), hx, norm_max);
  SCANN_CUDA(cudaGetLastError());
  return SCANN_OK;
}

scann_status tc_prepare_queries(const float* q, size_t nq, size_t dim, float scale, void* dst_bf16, float* qn,
                                cudaStream_t s) {
  const size_t kpad = tc_kpad(dim), q_pad = tc_queries_pad(nq, dim);
  if (q_pad == 0) return SCANN_OK;
  tc_prep_queries_kernel<<<static_cast<unsigned>((q_pad + 7) / 8), 256, 0, s>>>(
      q, nq, dim, kpad, q_pad, scale, static_cast<__nv_bfloat16*>(dst_bf16), qn);
  SCANN_CUDA(cudaGetLastError());
  return SCANN_OK;
}

scann_status launch_tc_scores(const TcScoreParams& p, cudaStream_t s) {
  SCANN_REQUIRE(tc_supported(p.dim), SCANN_INVALID_ARGUMENT, "dimension %zu too large for the tensor-core path", p.dim);
  SCANN_REQUIRE(p.row0 % kTcBN == 0, SCANN_INTERNAL, "row0 must be a multiple of %d", kTcBN);
  if (p.nq == 0 || p.nrows == 0) return SCANN_OK;
  const size_t kpad = tc_kpad(p.dim);
  const int ka = static_cast<int>(kpad / kTcAtomK);
  const int mt = ka <= 2 ? 2 : 1;
  CUtensorMap tmA, tmB;
  SCANN_TRY(make_map(&tmA, p.q_bf16, tc_queries_pad(p.nq, p.dim), kpad));
  SCANN_TRY(make_map(&tmB, p.rows_bf16, p.rows_pad_total, kpad));
  TcArgs a;
  a.nq = static_cast<uint32_t>(p.nq);
  a.q_tiles = static_cast<uint32_t>(tc_queries_pad(p.nq, p.dim) / (mt * kTcBM));
  a.row0 = static_cast<uint32_t>(p.row0);
  a.n_tiles = static_cast<uint32_t>((p.nrows + kTcBN - 1) / kTcBN);
  a.tile_stride = static_cast<uint32_t>(p.tile_stride < 1 ? 1 : p.tile_stride);
  // enough units to fill the machine a few times over; runs of >= 8 tiles amortise the A load, but only while the
  // machine stays full (small problems — a slice of a batch against a few thousand centroids — need the parallelism)
  uint32_t want_units = static_cast<uint32_t>(std::max(1, 4 * p.sms / static_cast<int>(a.q_tiles)));
  uint32_t tpu = (a.n_tiles + want_units - 1) / want_units;
  if (tpu < 8) {
    const uint32_t t8 = std::min<uint32_t>(8, a.n_tiles);
    if (a.q_tiles * ((a.n_tiles + t8 - 1) / t8) >= 2u * static_cast<uint32_t>(p.sms)) tpu = t8;
  }
  a.tiles_per_unit = tpu;
  a.n_units = (a.n_tiles + tpu - 1) / tpu;
  a.hx = p.hx;
  a.filter = p.filter ? 1 : 0;
  a.dense = p.dense;
  a.ld = p.ld;
  a.thr = p.thr;
  a.cand = p.cand;
  a.cap = static_cast<uint32_t>(p.cap);
  a.cand_cnt = p.cand_cnt;
  a.no_hx = (p.filter && p.hx_is_zero) ? 1 : 0;
  if (mt == 2) {
    if (ka == 1) return launch_tc<2, 1>(tmA, tmB, a, p.sms, s);
    return launch_tc<2, 2>(tmA, tmB, a, p.sms, s);
  }
  if (ka == 3) return launch_tc<1, 3>(tmA, tmB, a, p.sms, s);
  return launch_tc<1, 4>(tmA, tmB, a, p.sms, s);
}

}  // namespace scann

// ------------------------------------------------------------------------------------------ C ABI (parity tap)
extern "C" {

scann_status scann_tc_scores(const float* queries, size_t nq, size_t dim, const void* rows, int rows_i8, size_t n,
                             size_t stride, float scale, int want_norm, const float* thr, float* dense,
                             uint64_t* cand, size_t cap, uint32_t* cand_cnt, int device) {
  using namespace scann;
  SCANN_REQUIRE(queries && rows && nq > 0 && n > 0 && dim > 0 && stride >= dim, SCANN_INVALID_ARGUMENT, "bad arguments");
  SCANN_REQUIRE(thr ? (cand && cand_cnt && cap > 0) : dense != nullptr, SCANN_INVALID_ARGUMENT, "missing output");
  SCANN_REQUIRE(tc_supported(dim), SCANN_INVALID_ARGUMENT, "dimension %zu too large for the tensor-core path", dim);
  SCANN_TRY(check_device(device));
  DeviceGuard g(device);
  const size_t kpad = tc_kpad(dim), rpad = tc_rows_pad(n), qpad = tc_queries_pad(nq, dim);
  const size_t esz = rows_i8 ? 1 : 4;
  DevBuf<float> d_q, d_hx, d_qn, d_dense, d_thr;
  DevBuf<uint8_t> d_rows;
  DevBuf<uint16_t> d_rb, d_qb;
  DevBuf<unsigned long long> d_cand;
  DevBuf<uint32_t> d_cnt;
  cudaStream_t s = 0;
  SCANN_TRY(d_q.upload(queries, nq * dim, SCANN_HOST, s));
  SCANN_TRY(d_rows.upload(static_cast<const uint8_t*>(rows), n * stride * esz, SCANN_HOST, s));
  SCANN_TRY(d_rb.alloc(rpad * kpad));
  SCANN_TRY(d_qb.alloc(qpad * kpad));
  SCANN_TRY(d_hx.alloc(rpad));
  SCANN_TRY(d_qn.alloc(qpad));
  SCANN_TRY(tc_prepare_rows(d_rows.p, rows_i8 != 0, n, dim, stride, scale, want_norm != 0, d_rb.p, d_hx.p, nullptr, s));
  SCANN_TRY(tc_prepare_queries(d_q.p, nq, dim, rows_i8 ? scale : 1.0f, d_qb.p, d_qn.p, s));
  TcScoreParams p;
  p.q_bf16 = d_qb.p;
  p.nq = nq;
  p.dim = dim;
  p.rows_bf16 = d_rb.p;
  p.rows_pad_total = rpad;
  p.hx = d_hx.p;
  p.row0 = 0;
  p.nrows = n;
  p.filter = thr != nullptr;
  p.dense = nullptr;
  p.ld = rpad;
  p.thr = nullptr;
  p.cand = nullptr;
  p.cap = cap;
  p.cand_cnt = nullptr;
  p.sms = sm_count(device);
  if (thr) {
    SCANN_TRY(d_thr.upload(thr, nq, SCANN_HOST, s));
    SCANN_TRY(d_cand.alloc(nq * cap));
    SCANN_TRY(d_cnt.alloc(nq));
    SCANN_CUDA(cudaMemsetAsync(d_cnt.p, 0, nq * 4, s));
    p.thr = d_thr.p;
    p.cand = d_cand.p;
    p.cand_cnt = d_cnt.p;
  } else {
    SCANN_TRY(d_dense.alloc(nq * rpad));
    p.dense = d_dense.p;
  }
  SCANN_TRY(launch_tc_scores(p, s));
  if (thr) {
    SCANN_CUDA(cudaMemcpyAsync(cand, d_cand.p, nq * cap * 8, cudaMemcpyDeviceToHost, s));
    SCANN_CUDA(cudaMemcpyAsync(cand_cnt, d_cnt.p, nq * 4, cudaMemcpyDeviceToHost, s));
  } else {
    SCANN_CUDA(cudaMemcpy2DAsync(dense, n * 4, d_dense.p, rpad * 4, n * 4, nq, cudaMemcpyDeviceToHost, s));
  }
  SCANN_CUDA(cudaStreamSynchronize(s));
  return SCANN_OK;
}

}  // extern "C"
