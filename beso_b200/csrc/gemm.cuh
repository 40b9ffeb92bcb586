// Hand-written tcgen05 GEMM for the training path (train.cu): fp32 tensors in HBM, bf16 operands on the
// tensor cores, fp32 accumulation in TMEM.  See gemm.cu for the kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace beso {

// C[M][N] = (A . B^T + bias) .* mul + resid (+ C), all fp32 row-major in global memory:
//   A element (m, k):  a_kmajor ? A[m * lda + k] : A[k * lda + m]
//   B element (n, k):  b_kmajor ? B[n * ldb + k] : B[k * ldb + n]
// prec = 2: both operands are split into three bf16 images inside the kernel (24 mantissa bits) and every product is
// six MMAs (all cross terms down to 2^-24: the fp32-parity mode); prec = 1: two images, three MMAs (16 bits);
// prec = 0: one bf16 MMA.
struct GemmArgs {
  const float* A; int lda; int a_kmajor;
  const float* B; int ldb; int b_kmajor;
  float* C; int ldc;
  int M, N, K;
  const float* bias;          // per output column, or null
  const float* mul; int ldm;  // element-wise factor on (acc + bias) -- a dropout mask carrying 1 / (1 - p) -- or null
  const float* resid; int ldr;   // added element-wise after that (the residual stream), or null
  int accumulate;             // C += result
  float* gelu_out; int ldg;   // not null: C receives the pre-activation (acc + bias) and gelu_out = erf-GELU of it
  int prec;
};

struct GemmWs {               // split-K partial sums (deterministic reduction order)
  float* partial = nullptr;
  size_t floats = 0;
};

int gemm_run(const GemmArgs& a, GemmWs& ws, int sm_count, cudaStream_t st);
void gemm_ws_free(GemmWs& ws);

}  // namespace beso
