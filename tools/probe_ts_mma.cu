// Stand-alone probe: tcgen05.mma with the A operand in TENSOR MEMORY (the ".ts" form) -- checks the packing this repo
// assumes (row r = TMEM lane r, K elements 2k / 2k + 1 in the low / high half of 32-bit column k, a K = 16 step = 8
// columns) against a CPU product, and times it next to the shared-memory-A form.  Build + run (GPU box):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I beso_b200/csrc tools/probe_ts_mma.cu -o tools/build/probe_ts_mma
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "umma.cuh"

using namespace beso::umma;

constexpr int M = 128, N = 128, K = 64;

__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st_u32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// a: [M][K] fp16, b: [N][K] fp16, out: [M][N] fp32.  A goes to TMEM columns [256, 256 + K/2), B to a SW128 atom, D at column 0.
__global__ void __launch_bounds__(128, 1) probe(const __half* a, const __half* b, float* out, long long* cyc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(sm);
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar[2];
  const int warp = threadIdx.x >> 5, row = threadIdx.x;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar[0]), 1); mbar_init(smem_u32(&bar[1]), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  // B -> smem atom [N x 64] (rows n, K-major SW128); A also as a smem atom at +16 KB for the SS comparison
  for (int idx = threadIdx.x; idx < N * 8; idx += 128) {
    const int r = idx >> 3, ch = idx & 7;
    *reinterpret_cast<uint4*>(sm + sw128_offset(r, ch)) = *reinterpret_cast<const uint4*>(b + r * K + ch * 8);
    *reinterpret_cast<uint4*>(sm + 16384 + sw128_offset(r, ch)) = *reinterpret_cast<const uint4*>(a + r * K + ch * 8);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16);
  // A row -> TMEM: column k holds elements (2k, 2k + 1) as a packed half2 (low half = even k)
  uint32_t pk[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) pk[k] = *reinterpret_cast<const uint32_t*>(a + row * K + 2 * k);
  tmem_st_u32x32(lane_base + 256, pk);
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t idesc = idesc_f16_m128(N);
  if (threadIdx.x == 0) {
    const uint64_t b_desc = smem_desc_sw128(sbase), a_desc = smem_desc_sw128(sbase + 16384);
    long long t0 = clock64();
    for (int j = 0; j < 4; ++j) mma_ts(tm + 0, tm + 256 + 8 * j, b_desc + 2u * j, idesc, j ? 1u : 0u);
    mma_commit(smem_u32(&bar[0]));
    mbar_wait(smem_u32(&bar[0]), 0);
    long long t1 = clock64();
    for (int j = 0; j < 4; ++j) mma_bf16(tm + 128, a_desc + 2u * j, b_desc + 2u * j, idesc, j ? 1u : 0u);
    mma_commit(smem_u32(&bar[1]));
    mbar_wait(smem_u32(&bar[1]), 0);
    long long t2 = clock64();
    // throughput: 64 back-to-back MMAs of each form
    for (int i = 0; i < 64; ++i) mma_ts(tm + 384, tm + 256 + 8 * (i & 3), b_desc + 2u * (i & 3), idesc, 1u);
    mma_commit(smem_u32(&bar[0]));
    mbar_wait(smem_u32(&bar[0]), 1);
    long long t3 = clock64();
    for (int i = 0; i < 64; ++i) mma_bf16(tm + 384, a_desc + 2u * (i & 3), b_desc + 2u * (i & 3), idesc, 1u);
    mma_commit(smem_u32(&bar[1]));
    mbar_wait(smem_u32(&bar[1]), 1);
    long long t4 = clock64();
    // the precise mode's pattern: 4 MMAs with A in shared memory, 4 with A in tensor memory, 4 in shared memory, ...
    for (int i = 0; i < 64; ++i) {
      if ((i >> 2) % 3 == 1) mma_ts(tm + 384, tm + 256 + 8 * (i & 3), b_desc + 2u * (i & 3), idesc, 1u);
      else mma_bf16(tm + 384, a_desc + 2u * (i & 3), b_desc + 2u * (i & 3), idesc, 1u);
    }
    mma_commit(smem_u32(&bar[0]));
    mbar_wait(smem_u32(&bar[0]), 0);
    long long t5 = clock64();
    cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4;
  }
  __syncthreads();
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32], w[32];
    tmem_ld32(lane_base + c0, v);
    tmem_ld32(lane_base + 128 + c0, w);
    tmem_wait_ld();
    for (int i = 0; i < 32; ++i) { out[row * N + c0 + i] = v[i]; out[M * N + row * N + c0 + i] = w[i]; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  std::vector<__half> a(M * K), b(N * K);
  srand(1);
  for (auto& x : a) x = __float2half((rand() % 2001 - 1000) / 1000.0f);
  for (auto& x : b) x = __float2half((rand() % 2001 - 1000) / 1000.0f);
  __half *da, *db; float* dout; long long* dc;
  cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dout, 2 * M * N * 4); cudaMalloc(&dc, 64);
  cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  probe<<<1, 128, 65536>>>(da, db, dout, dc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> out(2 * M * N); long long cyc[5];
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(cyc, dc, 40, cudaMemcpyDeviceToHost);
  double e_ts = 0, e_ss = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)__half2float(a[m * K + k]) * (double)__half2float(b[n * K + k]);
      e_ts = fmax(e_ts, fabs(out[m * N + n] - ref));
      e_ss = fmax(e_ss, fabs(out[M * N + m * N + n] - ref));
    }
  printf("A from TMEM: max |err| %.3e   A from shared memory: max |err| %.3e   (K = %d products of |x| <= 1)\n", e_ts, e_ss, K);
  printf("latency of 4 MMAs + commit: TS %lld cyc, SS %lld cyc;  64 MMAs N=%d: TS %.1f cyc/MMA, SS %.1f cyc/MMA, alternating 4 SS / 4 TS / 4 SS %.1f cyc/MMA\n",
         cyc[0], cyc[1], N, cyc[2] / 64.0, cyc[3] / 64.0, cyc[4] / 64.0);
  return 0;
}
