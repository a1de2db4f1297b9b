// Micro-benchmark: issue throughput of the integer instructions the LUT16 scan is built from, alone and
// mixed, on one B200.  Each warp runs K independent dependency chains; results are reported as
// warp-instructions per cycle per SM (4 SMSPs).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 4096

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t a0, uint32_t one, uint32_t sh, long long* cyc) {
  uint32_t x[CHAINS], y[CHAINS];
  unsigned long long w[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { x[i] = a0 + threadIdx.x * 7 + i; y[i] = a0 * 3 + i; w[i] = x[i]; }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (MODE == 0) {  // PRMT
        asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(sh));
      } else if (MODE == 1) {  // IMAD
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(y[i]), "r"(one));
      } else if (MODE == 2) {  // IMAD.WIDE
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(y[i]), "r"(one));
      } else if (MODE == 3) {  // IDP4A
        asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(y[i]), "r"(one));
      } else if (MODE == 4) {  // IMAD.HI
        asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(y[i]), "r"(sh));
      } else if (MODE == 5) {  // LOP3
        asm volatile("lop3.b32 %0, %0, %1, %2, 0xCA;" : "+r"(x[i]) : "r"(y[i]), "r"(sh));
      } else if (MODE == 6) {  // PRMT + IMAD 1:1
        asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(sh));
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(y[i]) : "r"(y[i]), "r"(one));
      } else if (MODE == 7) {  // PRMT + IMAD.WIDE 1:1
        asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(sh));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(y[i]), "r"(one));
      } else if (MODE == 8) {  // PRMT + IDP4A 1:1
        asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(sh));
        asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(y[i]) : "r"(x[i]), "r"(one));
      } else if (MODE == 9) {  // LEA.HI-style shifted add (what mode 2 of the scan uses)
        x[i] += y[i] >> 8;
        asm volatile("" : "+r"(x[i]));
      } else if (MODE == 10) {  // 3 ALU (PRMT,PRMT,LOP3) + 3 FMA-pipe candidates (IMAD.WIDE, IDP4A, IDP4A)
        uint32_t lo, hi, r;
        asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(lo) : "r"(x[i]), "r"(y[i]), "r"(sh));
        asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(y[i]), "r"(x[i]), "r"(sh));
        asm volatile("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(r) : "r"(one), "r"(lo), "r"(hi));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(r), "r"(one));
        asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(r), "r"(one));
        asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(y[i]) : "r"(r), "r"(sh));
      } else if (MODE == 11) {  // current scan mix: PRMT,PRMT,LOP3,LOP3(mask),IMAD,LEA.HI
        uint32_t lo, hi, r, e;
        asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(lo) : "r"(x[i]), "r"(y[i]), "r"(sh));
        asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(y[i]), "r"(x[i]), "r"(sh));
        asm volatile("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(r) : "r"(one), "r"(lo), "r"(hi));
        e = r & 0x00FF00FFu;
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(e), "r"(one));
        y[i] += r >> 8;
        asm volatile("" : "+r"(y[i]));
      } else if (MODE == 12) {  // PRMT,PRMT,LOP3,LOP3(mask),IMAD,IMAD.WIDE
        uint32_t lo, hi, r, e;
        asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(lo) : "r"(x[i]), "r"(y[i]), "r"(sh));
        asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(y[i]), "r"(x[i]), "r"(sh));
        asm volatile("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(r) : "r"(one), "r"(lo), "r"(hi));
        e = r & 0x00FF00FFu;
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(e), "r"(one));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(r), "r"(one));
      } else if (MODE == 13) {  // PRMT,PRMT,LOP3 + 4 IDP4A
        uint32_t lo, hi, r;
        asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(lo) : "r"(x[i]), "r"(y[i]), "r"(sh));
        asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(y[i]), "r"(x[i]), "r"(sh));
        asm volatile("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(r) : "r"(one), "r"(lo), "r"(hi));
        uint32_t lo32 = (uint32_t)w[i], hi32 = (uint32_t)(w[i] >> 32);
        asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(r), "r"(one));
        asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(y[i]) : "r"(r), "r"(sh));
        asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(lo32) : "r"(r), "r"(a0));
        asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(hi32) : "r"(r), "r"(sh));
        w[i] = ((unsigned long long)hi32 << 32) | lo32;
      } else if (MODE == 14) {  // scan mode 3 now: PRMT,PRMT,LOP3 + 2 IDP2A per 4 lookups
        uint32_t lo, hi, r;
        asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(lo) : "r"(x[i]), "r"(y[i]), "r"(sh));
        asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(y[i]), "r"(x[i]), "r"(sh));
        asm volatile("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(r) : "r"(one), "r"(lo), "r"(hi));
        asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(a0), "r"(r));
        asm volatile("dp2a.hi.u32.u32 %0, %1, %2, %0;" : "+r"(y[i]) : "r"(a0), "r"(r));
      } else if (MODE == 15) {  // candidate: PRMT,PRMT + 4 IDP2A with per-byte weights (no select LOP3)
        uint32_t lo, hi;
        asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(lo) : "r"(x[i]), "r"(y[i]), "r"(sh));
        asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(y[i]), "r"(x[i]), "r"(sh));
        uint32_t lo32 = (uint32_t)w[i], hi32 = (uint32_t)(w[i] >> 32);
        asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(lo32) : "r"(a0), "r"(lo));
        asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(lo32) : "r"(one), "r"(hi));
        asm volatile("dp2a.hi.u32.u32 %0, %1, %2, %0;" : "+r"(hi32) : "r"(a0), "r"(lo));
        asm volatile("dp2a.hi.u32.u32 %0, %1, %2, %0;" : "+r"(hi32) : "r"(one), "r"(hi));
        x[i] ^= lo; y[i] ^= hi;
        w[i] = ((unsigned long long)hi32 << 32) | lo32;
      } else if (MODE == 16) {  // IDP2A alone
        asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(y[i]), "r"(one));
      }
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += x[i] + y[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter, int warps_per_sm) {
  int nsm = 148, threads = 256, blocks_per_sm = warps_per_sm / 8;
  int grid = nsm * blocks_per_sm;
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, (size_t)grid * threads * 4);
  cudaMalloc(&cyc, grid * 8);
  k<MODE><<<grid, threads>>>(out, 12345u, 1u, 0x3210u, cyc);
  cudaDeviceSynchronize();
  k<MODE><<<grid, threads>>>(out, 12345u, 1u, 0x3210u, cyc);
  cudaDeviceSynchronize();
  long long* h = new long long[grid];
  cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
  double mx = 0; for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
  double winstr = (double)ITERS * CHAINS * per_iter * warps_per_sm;  // warp-instructions per SM
  printf("%-44s warps/SM=%2d  %.3f warp-instr/cyc/SM  (%.2f cyc per chain-step per SMSP-warp-set)\n", name, warps_per_sm,
         winstr / mx, mx / ((double)ITERS * CHAINS) / (warps_per_sm / 4.0));
  cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc); delete[] h;
}

int main() {
  for (int w : {8, 16}) {
    run<0>("PRMT", 1, w);
    run<5>("LOP3", 1, w);
    run<9>("SHF+IADD / LEA.HI", 1, w);
    run<1>("IMAD", 1, w);
    run<2>("IMAD.WIDE", 1, w);
    run<3>("IDP4A", 1, w);
    run<4>("IMAD.HI", 1, w);
    run<6>("PRMT+IMAD", 2, w);
    run<7>("PRMT+IMAD.WIDE", 2, w);
    run<8>("PRMT+IDP4A", 2, w);
    run<11>("scan mix now: 2PRMT+LOP3+LOP3+IMAD+LEA.HI", 6, w);
    run<12>("2PRMT+LOP3+LOP3+IMAD+IMAD.WIDE", 6, w);
    run<10>("2PRMT+LOP3+IMAD.WIDE+2IDP4A", 6, w);
    run<13>("2PRMT+LOP3+4IDP4A", 7, w);
    run<16>("IDP2A", 1, w);
    run<14>("scan mode 3: 2PRMT+LOP3+2IDP2A", 5, w);
    run<15>("candidate: 2PRMT+4IDP2A(+2LOP3 xor)", 8, w);
  }
  return 0;
}
