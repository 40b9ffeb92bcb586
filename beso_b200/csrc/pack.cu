// Weight re-layout kernels.  PRECISE mode reads every Linear weight as a transposed fp32 image
// Wt[k][n] = W[n][k] (nn.Linear stores (out, in) row-major) so that a warp's lanes read
// consecutive output columns coalesced.  The four GEMM weights of a block are additionally tiled by
// 64-column groups, [ceil(N/64)][K][64] (zero padded): the warp that owns a column group walks its K
// rows at a constant 256-byte stride, so the k-unrolled loads need one pointer and immediate offsets.
#include "common.cuh"

namespace beso {
namespace {

// dst[k * ld_dst + col0 + n] = src[n * K + k]; 32x32 tiles through shared memory.
__global__ void transpose_kernel(const float* __restrict__ src, int N, int K, float* __restrict__ dst,
                                 int ld_dst, int col0) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    tile[i][threadIdx.x] = (n < N && k < K) ? src[(size_t)n * K + k] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    if (k < K && n < N) dst[(size_t)k * ld_dst + col0 + n] = tile[threadIdx.x][i];
  }
}

// tiled variant: column c = col0 + n of the [K][Ntot] image goes to dst[((c / 64) * K + k) * 64 + c % 64]
__global__ void transpose_tiled_kernel(const float* __restrict__ src, int N, int K, float* __restrict__ dst, int col0) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    tile[i][threadIdx.x] = (n < N && k < K) ? src[(size_t)n * K + k] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    if (k < K && n < N) {
      const int c = col0 + n;
      dst[((size_t)(c >> 6) * K + k) * 64 + (c & 63)] = tile[threadIdx.x][i];
    }
  }
}

__global__ void copy_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

}  // namespace

int pack_transpose(const float* src, int N, int K, float* dst, int ld_dst, int col0, cudaStream_t s) {
  dim3 grid((K + 31) / 32, (N + 31) / 32), block(32, 8);
  transpose_kernel<<<grid, block, 0, s>>>(src, N, K, dst, ld_dst, col0);
  ++g_kernel_launches;
  BESO_CUDA(cudaGetLastError());
  return BESO_OK;
}

int pack_transpose_tiled(const float* src, int N, int K, float* dst, int col0, cudaStream_t s) {
  dim3 grid((K + 31) / 32, (N + 31) / 32), block(32, 8);
  transpose_tiled_kernel<<<grid, block, 0, s>>>(src, N, K, dst, col0);
  ++g_kernel_launches;
  BESO_CUDA(cudaGetLastError());
  return BESO_OK;
}

int pack_copy(const float* src, float* dst, int64_t n, cudaStream_t s) {
  const int block = 256;
  const int grid = (int)((n + block - 1) / block < 1024 ? (n + block - 1) / block : 1024);
  copy_kernel<<<grid > 0 ? grid : 1, block, 0, s>>>(src, dst, (long long)n);
  ++g_kernel_launches;
  BESO_CUDA(cudaGetLastError());
  return BESO_OK;
}

}  // namespace beso
