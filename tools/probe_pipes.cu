// Stand-alone probe: issue throughput (cycles per warp instruction per SM sub-partition) of the CUDA-core
// instructions the fused kernel's epilogues are made of.  8 independent dependency chains per thread.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define CHAINS 8
#define REPS 64
#define LOOPS 64

template <int OP>
__device__ __forceinline__ void step(uint32_t (&r)[CHAINS], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) {
    if (OP == 0) asm volatile("add.f32 %0, %0, %1;" : "+r"(r[i]) : "r"(k0));                    // FADD reg
    if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(k0), "r"(k1));     // FFMA reg
    if (OP == 2) asm volatile("fma.rn.f32 %0, %0, 0f3F7FF000, %1;" : "+r"(r[i]) : "r"(k0));       // FFMA imm (multiplier)
    if (OP == 3) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(k0), "r"(k1));   // HFMA2 reg
    if (OP == 4) asm volatile("{.reg .b32 c; mov.b32 c, 0x3BFF3BFF; fma.rn.f16x2 %0, %0, %1, c;}" : "+r"(r[i]) : "r"(k0));   // HFMA2 imm addend
    if (OP == 5) asm volatile("add.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(k0));                   // HADD2
    if (OP == 6) asm volatile("{.reg .f32 a, b; mov.b32 a, %0; mov.b32 b, %1; cvt.rn.f16x2.f32 %0, a, b;}" : "+r"(r[i]) : "r"(k0));   // F2FP
    if (OP == 7) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(r[i]) : "r"(k0));            // PRMT
    if (OP == 8) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(k0), "r"(k1)); // LOP3
    if (OP == 9) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(r[i]));                        // MUFU.EX2 fp32
    if (OP == 10) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r[i]));                         // 2 x MUFU.EX2.F16 + PRMT
    if (OP == 11) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(k0));                  // HMNMX2
    if (OP == 12) asm volatile("max.f32 %0, %0, %1;" : "+r"(r[i]) : "r"(k0));                    // FMNMX
    if (OP == 13) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(k0), "r"(k1));    // IMAD
    if (OP == 14) asm volatile("mul.f32 %0, %0, %1;" : "+r"(r[i]) : "r"(k0));                    // FMUL
  }
}

template <int OP>
__global__ void __launch_bounds__(512, 1) k(uint32_t* out, long long* cyc, int n_warps, uint32_t k0, uint32_t k1) {
  const int warp = threadIdx.x >> 5;
  uint32_t r[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) r[i] = threadIdx.x * 7 + i;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < n_warps) {
#pragma unroll 1
    for (int l = 0; l < LOOPS; ++l) {
#pragma unroll
      for (int rep = 0; rep < REPS; ++rep) step<OP>(r, k0, k1);
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc ^= r[i];
  out[threadIdx.x] = acc;
  if ((threadIdx.x & 31) == 0) cyc[warp] = t1 - t0;
}

template <int OP>
void run(const char* name, uint32_t* out, long long* cyc) {
  for (int nw : {4, 8, 16}) {
    long long h[16];
    for (int it = 0; it < 2; ++it) {
      k<OP><<<1, 512>>>(out, cyc, nw, 0x3C003C00u, 0x00010001u);
      cudaDeviceSynchronize();
    }
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double n = (double)CHAINS * REPS * LOOPS;
    printf("%-22s %d warp(s)/SMSP: %.2f cycles per warp instruction per warp -> %.2f per SMSP\n", name, nw / 4, h[0] / n, h[0] / n / (nw / 4));
  }
}

int main() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 512 * 4); cudaMalloc(&cyc, 16 * 8);
  run<0>("FADD reg", out, cyc);  run<1>("FFMA reg", out, cyc);  run<2>("FFMA imm-mul", out, cyc);
  run<14>("FMUL reg", out, cyc);
  run<3>("HFMA2 reg", out, cyc); run<4>("HFMA2 imm-add", out, cyc); run<5>("HADD2", out, cyc);
  run<6>("F2FP.F16.F32.PACK", out, cyc); run<7>("PRMT", out, cyc); run<8>("LOP3", out, cyc);
  run<9>("MUFU.EX2 f32", out, cyc); run<10>("ex2.f16x2 (2 MUFU+PRMT)", out, cyc); run<11>("HMNMX2", out, cyc);
  run<12>("FMNMX", out, cyc); run<13>("IMAD", out, cyc);
  return 0;
}
