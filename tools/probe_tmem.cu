// Stand-alone probe: tcgen05.ld throughput / latency from TMEM with 1, 4 and 8 reading warps, and the
// cost of the generic->async proxy fence.  Build + run (GPU box):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I beso_b200/csrc tools/probe_tmem.cu -o tools/build/probe_tmem
#include <cstdio>
#include <cuda_runtime.h>

#include "umma.cuh"

using namespace beso::umma;

__device__ __forceinline__ float consume(const float (&v)[32]) {
  float a = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) a += v[i];
  return a;
}

// mode 0: 64 loads of 32 columns, 4 in flight between waits.  mode 1: dependent chain (ld, wait) x 16.
__global__ void __launch_bounds__(256, 1) probe(long long* out, float* sink, int n_warps, int mode) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t lane_base = tm + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < n_warps) {
    if (mode == 0) {
      float a[32], b[32], c[32], d[32];
#pragma unroll 1
      for (int it = 0; it < 16; ++it) {
        const uint32_t col = (uint32_t)((it & 3) * 128);
        tmem_ld32(lane_base + col, a);
        tmem_ld32(lane_base + col + 32, b);
        tmem_ld32(lane_base + col + 64, c);
        tmem_ld32(lane_base + col + 96, d);
        tmem_wait_ld();
        acc += consume(a) + consume(b) + consume(c) + consume(d);
      }
    } else if (mode == 1) {
      float a[32];
#pragma unroll 1
      for (int it = 0; it < 16; ++it) {
        tmem_ld32(lane_base + (uint32_t)((it & 15) * 32), a);
        tmem_wait_ld();
        acc += a[it];
      }
    } else {
      // fence.proxy.async cost: 16 x (st.shared, fence)
      __shared__ float buf[256];
#pragma unroll 1
      for (int it = 0; it < 16; ++it) {
        buf[threadIdx.x] = acc + it;
        fence_async_smem();
      }
      acc += buf[(threadIdx.x + 1) & 255];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  const long long t2 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  if (acc == 123.456f) sink[threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* out;
  float* sink;
  cudaMalloc(&out, 64);
  cudaMalloc(&sink, 4096);
  const int cfgs[][2] = {{1, 0}, {4, 0}, {8, 0}, {1, 1}, {4, 1}, {8, 1}, {1, 2}, {8, 2}};
  for (auto& c : cfgs) {
    long long h[2] = {0, 0};
    for (int rep = 0; rep < 3; ++rep) {
      probe<<<1, 256>>>(out, sink, c[0], c[1]);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    }
    if (c[1] == 0) {
      const double bytes = 64.0 * 4096 * c[0];
      printf("tmem ld x32, %d warps, 4 in flight: warp0 %lld cyc, all %lld cyc -> %.1f B/clk/SM (%.1f B/clk/warp)\n", c[0], h[0], h[1],
             bytes / h[1], bytes / h[1] / c[0]);
    } else if (c[1] == 1) {
      printf("tmem ld x32 dependent chain, %d warps: %.1f cyc per (ld + wait)\n", c[0], h[1] / 16.0);
    } else {
      printf("st.shared + fence.proxy.async, %d warps: %.1f cyc per iteration\n", c[0], h[1] / 16.0);
    }
  }
  return 0;
}
