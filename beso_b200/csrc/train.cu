// Training path: GCDenoiser.loss (score_wrappers.py:45-79) forward + hand-derived backward, and the
// data-parallel gradient exchange (one NCCL all-reduce of the flat fp32 gradient over NVLink).
//
// Every dense product (forward, data gradient, weight gradient) runs on the tcgen05 tensor cores through the
// hand-written GEMM of gemm.cu: fp32 tensors in HBM, three bf16 operand images and six MMAs per product by
// default (fp32-parity mode, the one the gradient goldens pin), two images / three MMAs with
// BESO_FLAG_TRAIN_SPLIT2, one bf16 MMA per product with BESO_FLAG_TRAIN_FAST.  Bias, residual add and erf-GELU ride in the GEMM epilogues.  Everything else --
// embeddings + interleave, LayerNorm forward/backward, causal attention forward/backward, GELU backward,
// bias / column reductions, dropout, the Karras pre-conditioned loss -- is hand-written below in fp32.
// Gradients are written into ONE flat buffer in nn.Module.parameters() order, which is what the
// all-reduce operates on.
//
// Dropout (score_gpts.py:37-38,72,79,109) and the element-wise goal mask of CFG training (:360-371) take their
// masks from the caller, who draws them with the reference's torch calls in the reference's op order.
#include <math.h>
#include <nccl.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "gemm.cuh"
#include "plan.cuh"

namespace beso {
namespace {

constexpr int kTB = 256;
inline int blocks_for(size_t n, int per = kTB) { return (int)((n + per - 1) / per); }
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct Dims { int B, t, T, G, obs, act, d, H, hs, L, F, M; float sigma_data; };

// ---- embeddings (score_gpts.py:284-337) ---------------------------------------------------------------
// X[r][c] for row r = (b, tok); also writes the network input x_in = (a + n*sigma) * c_in  [B,t,act].
__global__ void embed_fwd_kernel(Dims D, const float* __restrict__ state, const float* __restrict__ action,
                                 const float* __restrict__ goal, const float* __restrict__ noise,
                                 const float* __restrict__ sigma, const float* __restrict__ goal_keep, int pred_last,
                                 const float* __restrict__ pos, const float* __restrict__ tokw, const float* __restrict__ tokb,
                                 const float* __restrict__ sigw, const float* __restrict__ sigb,
                                 const float* __restrict__ actw, const float* __restrict__ actb,
                                 const float* __restrict__ drop_mask, float* __restrict__ X, float* __restrict__ xin) {
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= (size_t)D.M * D.d) return;
  const int c = (int)(idx % D.d);
  const int r = (int)(idx / D.d), b = r / D.T, tok = r - b * D.T;
  const float sg = sigma[b];
  float v;
  if (tok == 0) {
    v = fmaf(logf(sg) / 4.0f, sigw[c], sigb[c]);
  } else if (tok <= D.G) {
    const int g = tok - 1;
    const float* in = goal + ((size_t)b * D.G + g) * D.obs;
    const float* kp = goal_keep ? goal_keep + ((size_t)b * D.G + g) * D.obs : nullptr;
    float a = 0.f;
    for (int k = 0; k < D.obs; ++k) a = fmaf(kp ? in[k] * kp[k] : in[k], tokw[(size_t)c * D.obs + k], a);
    v = a + tokb[c] + pos[(size_t)g * D.d + c];
  } else {
    const int j = tok - 1 - D.G, step = j >> 1;
    float a = 0.f;
    if ((j & 1) == 0) {
      const float* in = state + ((size_t)b * D.t + step) * D.obs;
      for (int k = 0; k < D.obs; ++k) a = fmaf(in[k], tokw[(size_t)c * D.obs + k], a);
      v = a + tokb[c];
    } else {
      const float c_in = 1.0f / sqrtf(sg * sg + D.sigma_data * D.sigma_data);
      const size_t o = ((size_t)b * D.t + step) * D.act;
      for (int k = 0; k < D.act; ++k) {
        const float nz = (pred_last && step != D.t - 1) ? 0.f : noise[o + k];
        const float xi = (action[o + k] + nz * sg) * c_in;
        if (c == 0) xin[o + k] = xi;
        a = fmaf(xi, actw[(size_t)c * D.act + k], a);
      }
      v = a + actb[c];
    }
    v += pos[(size_t)(D.G + step) * D.d + c];
  }
  if (drop_mask) v *= drop_mask[idx];                   // self.drop(input_seq), score_gpts.py:338
  X[idx] = v;
}

// ---- LayerNorm (eps 1e-5, biased variance): warp per row ------------------------------------------------
__global__ void ln_fwd_kernel(const float* __restrict__ X, int M, int d, const float* __restrict__ w,
                              const float* __restrict__ b, float* __restrict__ Y, float* __restrict__ stats) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= M) return;
  const float* x = X + (size_t)r * d;
  float s = 0.f;
  for (int c = lane; c < d; c += 32) s += x[c];
  const float mean = wsum(s) / d;
  float v = 0.f;
  for (int c = lane; c < d; c += 32) { const float t = x[c] - mean; v = fmaf(t, t, v); }
  const float rstd = 1.0f / sqrtf(wsum(v) / d + 1e-5f);
  for (int c = lane; c < d; c += 32) Y[(size_t)r * d + c] = (x[c] - mean) * rstd * w[c] + b[c];
  if (lane == 0) { stats[2 * r] = mean; stats[2 * r + 1] = rstd; }
}
// dX (+)= LN'(dY);  dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * w
__global__ void ln_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ X, const float* __restrict__ stats,
                              const float* __restrict__ w, int M, int d, float* __restrict__ dX, int accumulate,
                              float* __restrict__ dyxh) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= M) return;
  const float mean = stats[2 * r], rstd = stats[2 * r + 1];
  const float* x = X + (size_t)r * d;
  const float* dy = dY + (size_t)r * d;
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < d; c += 32) { const float g = dy[c] * w[c], xh = (x[c] - mean) * rstd; s1 += g; s2 = fmaf(g, xh, s2); }
  s1 = wsum(s1) / d; s2 = wsum(s2) / d;
  for (int c = lane; c < d; c += 32) {
    const float g = dy[c] * w[c], xh = (x[c] - mean) * rstd;
    const float v = rstd * (g - s1 - xh * s2);
    dX[(size_t)r * d + c] = accumulate ? dX[(size_t)r * d + c] + v : v;
    dyxh[(size_t)r * d + c] = dy[c] * xh;          // column sums of this give d(loss)/d(ln.weight)
  }
}
// dw[c] = sum_r dy * xhat, db[c] = sum_r dy : one block per 32 columns, rows strided over warps
__global__ void ln_param_grad_kernel(const float* __restrict__ dY, const float* __restrict__ X,
                                     const float* __restrict__ stats, int M, int d, float* __restrict__ dw,
                                     float* __restrict__ db) {
  __shared__ float sw[8][33], sb[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x, wy = threadIdx.y;
  float aw = 0.f, ab = 0.f;
  if (c < d)
    for (int r = wy; r < M; r += 8) {
      const float dy = dY[(size_t)r * d + c];
      aw = fmaf(dy, (X[(size_t)r * d + c] - stats[2 * r]) * stats[2 * r + 1], aw);
      ab += dy;
    }
  sw[wy][threadIdx.x] = aw; sb[wy][threadIdx.x] = ab;
  __syncthreads();
  if (wy == 0 && c < d) {
    for (int i = 1; i < 8; ++i) { aw += sw[i][threadIdx.x]; ab += sb[i][threadIdx.x]; }
    dw[c] = aw; db[c] = ab;
  }
}

// ---- bias add / column sums -------------------------------------------------------------------------------
__global__ void bias_add_kernel(float* __restrict__ C, const float* __restrict__ bias, size_t M, int N, int ldc) {
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= M * N) return;
  const size_t r = idx / N; const int c = (int)(idx % N);
  C[r * ldc + c] += bias[c];
}
// dst = src + bias (row broadcast), N % 4 == 0: the residual stream copy and the bias of the following
// beta = 1 GEMM in one pass
__global__ void copy_add_bias_kernel(float* __restrict__ dst, const float* __restrict__ src, const float* __restrict__ bias,
                                     size_t n4, int N) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = reinterpret_cast<const float4*>(src)[i];
  const float4 b = *reinterpret_cast<const float4*>(bias + (int)((i * 4) % N));
  v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  reinterpret_cast<float4*>(dst)[i] = v;
}
// dst = a .* m (dropout masks carry the 1 / (1 - p) scale)
__global__ void mul_kernel(float* __restrict__ dst, const float* __restrict__ a, const float* __restrict__ m, size_t n4) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = reinterpret_cast<const float4*>(a)[i];
  const float4 w = reinterpret_cast<const float4*>(m)[i];
  v.x *= w.x; v.y *= w.y; v.z *= w.z; v.w *= w.w;
  reinterpret_cast<float4*>(dst)[i] = v;
}
// Column sums (bias gradients) in two deterministic stages: partial[rb][c] over row block rb, then the sum over rb.
// Stage 1: block = 32 columns x 8 row lanes, grid = (N / 32, kColsumRowBlocks): each warp reads 128 contiguous
// bytes of a row, four rows in flight per thread, ~1200 blocks keep every SM busy (HBM-bound pass over A).
constexpr int kColsumRowBlocks = 148;
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ A, int M, int N, int lda,
                                                             float* __restrict__ partial) {
  __shared__ float sh[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x, ty = threadIdx.y;
  const int stride = kColsumRowBlocks * 8;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (c < N) {
    int r = blockIdx.y * 8 + ty;
    for (; r + 3 * stride < M; r += 4 * stride) {
      a0 += A[(size_t)r * lda + c];
      a1 += A[(size_t)(r + stride) * lda + c];
      a2 += A[(size_t)(r + 2 * stride) * lda + c];
      a3 += A[(size_t)(r + 3 * stride) * lda + c];
    }
    for (; r < M; r += stride) a0 += A[(size_t)r * lda + c];
  }
  sh[ty][threadIdx.x] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (ty == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
    partial[(size_t)blockIdx.y * N + c] = t;
  }
}
__global__ void colsum_reduce_kernel(const float* __restrict__ partial, int N, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  float a = 0.f;
  for (int rb = 0; rb < kColsumRowBlocks; ++rb) a += partial[(size_t)rb * N + c];
  out[c] = a;
}

// ---- erf-GELU -----------------------------------------------------------------------------------------------
// U += bias (kept for the backward pass), G = gelu(U): one pass over the 4d-wide hidden activations
__global__ void bias_gelu_fwd_kernel(float* __restrict__ U, const float* __restrict__ bias, float* __restrict__ G, size_t n4, int N) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 u = reinterpret_cast<float4*>(U)[i];
  const float4 b = *reinterpret_cast<const float4*>(bias + (int)((i * 4) % N));
  u.x += b.x; u.y += b.y; u.z += b.z; u.w += b.w;
  reinterpret_cast<float4*>(U)[i] = u;
  float4 g;
  g.x = 0.5f * u.x * (1.0f + erff(u.x * 0.70710678118654752440f));
  g.y = 0.5f * u.y * (1.0f + erff(u.y * 0.70710678118654752440f));
  g.z = 0.5f * u.z * (1.0f + erff(u.z * 0.70710678118654752440f));
  g.w = 0.5f * u.w * (1.0f + erff(u.w * 0.70710678118654752440f));
  reinterpret_cast<float4*>(G)[i] = g;
}
__device__ __forceinline__ float gelu_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}
__global__ void gelu_bwd_kernel(const float* __restrict__ U, float* __restrict__ dG, size_t n4) {   // in place: dU
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 u = reinterpret_cast<const float4*>(U)[i];
  float4 g = reinterpret_cast<float4*>(dG)[i];
  g.x *= gelu_grad(u.x); g.y *= gelu_grad(u.y); g.z *= gelu_grad(u.z); g.w *= gelu_grad(u.w);
  reinterpret_cast<float4*>(dG)[i] = g;
}

// ---- causal attention, one CTA per (sequence, head) -------------------------------------------------------------
// QKV [M][3d] (q | k | v); P [B][H][T][T] saved; Y [M][d].  T <= 64, head size <= 64.  Q, K, V of the head are
// staged in shared memory with coalesced loads (the three Linear biases are added on the way in and the biased
// values written back for the backward pass), then T x T scores, the row softmax and P V are small dot products
// out of shared memory.  HBM-bound: every activation byte is read once and written once.
constexpr int kAttnThreads = 128;
__host__ __device__ inline int attn_lds(int hs) { return ((hs + 3) & ~3) + 4; }   // row pitch: 16-byte rows, conflict-free float4 reads
__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, fmaf(a.x, b.x, acc))));
}
// out[i0..i0+1][j0..j0+1] (+)= A rows i0, i0+1 . B rows j0, j0+1 over hs4 float4 columns (2 x 2 register tile)
__device__ __forceinline__ void tile2x2(const float* A, const float* Bm, int lds, int hs4, int i0, int j0, int T, float (&o)[2][2]) {
  const int i1 = min(i0 + 1, T - 1), j1 = min(j0 + 1, T - 1);
  const float4* a0 = reinterpret_cast<const float4*>(A + i0 * lds);
  const float4* a1 = reinterpret_cast<const float4*>(A + i1 * lds);
  const float4* b0 = reinterpret_cast<const float4*>(Bm + j0 * lds);
  const float4* b1 = reinterpret_cast<const float4*>(Bm + j1 * lds);
  for (int e = 0; e < hs4; ++e) {
    const float4 x0 = a0[e], x1 = a1[e], y0 = b0[e], y1 = b1[e];
    o[0][0] = dot4(x0, y0, o[0][0]); o[0][1] = dot4(x0, y1, o[0][1]);
    o[1][0] = dot4(x1, y0, o[1][0]); o[1][1] = dot4(x1, y1, o[1][1]);
  }
}
__global__ void __launch_bounds__(kAttnThreads)
attn_fwd_kernel(Dims D, float* __restrict__ QKV, const float* __restrict__ bq, const float* __restrict__ bk,
                const float* __restrict__ bv, const float* __restrict__ drop_mask, float* __restrict__ P, float* __restrict__ Y) {
  extern __shared__ __align__(16) float sh[];
  const int T = D.T, hs = D.hs, ld = 3 * D.d, lds = attn_lds(hs), hs4 = (hs + 3) >> 2, ldt = T + 1;
  float* q = sh;                       // [T][lds], columns hs .. lds-1 zero
  float* k = q + T * lds;
  float* v = k + T * lds;
  float* sc = v + T * lds;             // [T][T + 1]
  const int b = blockIdx.x / D.H, h = blockIdx.x % D.H;
  const size_t row0 = (size_t)b * T;
  // hs % 4 == 0 (checked on the host): 16-byte loads, all of a thread's loads issued before its stores (the
  // write-back aliases the source, so a load-store-load loop would serialise on the memory latency)
  const int n4 = 3 * T * hs4;
  for (int base = 0; base < n4; base += 4 * kAttnThreads) {
    float4 val[4];
    float4* g[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = base + u * kAttnThreads + threadIdx.x;
      g[u] = nullptr;
      if (idx < n4) {
        const int part = idx / (T * hs4), r = (idx / hs4) % T, e4 = idx % hs4;
        g[u] = reinterpret_cast<float4*>(QKV + (row0 + r) * ld + part * D.d + h * hs) + e4;
        val[u] = *g[u];
        const float4 bb = reinterpret_cast<const float4*>((part == 0 ? bq : part == 1 ? bk : bv) + h * hs)[e4];
        val[u].x += bb.x; val[u].y += bb.y; val[u].z += bb.z; val[u].w += bb.w;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = base + u * kAttnThreads + threadIdx.x;
      if (idx < n4) {
        const int part = idx / (T * hs4), r = (idx / hs4) % T, e4 = idx % hs4;
        *g[u] = val[u];
        reinterpret_cast<float4*>((part == 0 ? q : part == 1 ? k : v) + r * lds)[e4] = val[u];
      }
    }
  }
  for (int r = threadIdx.x; r < 3 * T; r += kAttnThreads)      // zero the 4 pad columns of every row
    *reinterpret_cast<float4*>(q + r * lds + hs4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const float scale = 1.0f / sqrtf((float)hs);
  const int T2 = (T + 1) >> 1;
  for (int idx = threadIdx.x; idx < T2 * T2; idx += kAttnThreads) {
    const int i0 = (idx / T2) * 2, j0 = (idx % T2) * 2;
    if (j0 > i0 + 1) continue;                          // fully masked tile
    float o[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    tile2x2(q, k, lds, hs4, i0, j0, T, o);
#pragma unroll
    for (int di = 0; di < 2; ++di)
#pragma unroll
      for (int dj = 0; dj < 2; ++dj) {
        const int i = i0 + di, j = j0 + dj;
        if (i < T && j < T) sc[i * ldt + j] = (j <= i) ? o[di][dj] * scale : -INFINITY;
      }
  }
  __syncthreads();
  float* Pm = P + ((size_t)b * D.H + h) * T * T;
  for (int i = threadIdx.x >> 5; i < T; i += kAttnThreads / 32) {        // one warp per row
    const int lane = threadIdx.x & 31;
    const float a0 = (lane < T && lane <= i) ? sc[i * ldt + lane] : -INFINITY;
    const float a1 = (lane + 32 < T && lane + 32 <= i) ? sc[i * ldt + lane + 32] : -INFINITY;
    float mx = fmaxf(a0, a1);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float p0 = (lane <= i) ? expf(a0 - mx) : 0.f, p1 = (lane + 32 <= i) ? expf(a1 - mx) : 0.f;
    const float inv = 1.0f / wsum(p0 + p1);
    p0 *= inv; p1 *= inv;
    // P (before attn_drop) is what the softmax backward needs; P V uses the dropped-out probabilities
    const float* dm = drop_mask ? drop_mask + ((size_t)b * D.H + h) * T * T + (size_t)i * T : nullptr;
    if (lane < T) { Pm[i * T + lane] = p0; sc[i * ldt + lane] = dm ? p0 * dm[lane] : p0; }
    if (lane + 32 < T) { Pm[i * T + lane + 32] = p1; sc[i * ldt + lane + 32] = dm ? p1 * dm[lane + 32] : p1; }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < T * hs4; idx += kAttnThreads) {      // (row, 4 columns) per thread
    const int i = idx / hs4, e4 = idx % hs4;
    float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j <= i; ++j) {
      const float pj = sc[i * ldt + j];
      const float4 vv = reinterpret_cast<const float4*>(v + j * lds)[e4];
      y.x = fmaf(pj, vv.x, y.x); y.y = fmaf(pj, vv.y, y.y); y.z = fmaf(pj, vv.z, y.z); y.w = fmaf(pj, vv.w, y.w);
    }
    reinterpret_cast<float4*>(Y + (row0 + i) * D.d + h * hs)[e4] = y;
  }
}
inline size_t attn_fwd_smem(int T, int hs) { return ((size_t)3 * T * attn_lds(hs) + (size_t)T * (T + 1)) * sizeof(float); }

__global__ void __launch_bounds__(kAttnThreads)
attn_bwd_kernel(Dims D, const float* __restrict__ QKV, const float* __restrict__ P, const float* __restrict__ drop_mask,
                const float* __restrict__ dY, float* __restrict__ dQKV) {
  extern __shared__ __align__(16) float sh[];
  const int T = D.T, hs = D.hs, ld = 3 * D.d, lds = attn_lds(hs), hs4 = (hs + 3) >> 2, ldt = T + 1;
  float* q = sh;                       // [T][lds]
  float* k = q + T * lds;
  float* v = k + T * lds;
  float* dy = v + T * lds;
  float* Pm = dy + T * lds;            // [T][T + 1]
  float* dS = Pm + T * ldt;            // [T][T + 1]
  float* Pd = dS + T * ldt;            // [T][T + 1]: attn_drop(P) (only with a dropout mask)
  const int b = blockIdx.x / D.H, h = blockIdx.x % D.H;
  const size_t row0 = (size_t)b * T;
  const float scale = 1.0f / sqrtf((float)hs);
  for (int idx = threadIdx.x; idx < 4 * T * hs4; idx += kAttnThreads) {
    const int part = idx / (T * hs4), r = (idx / hs4) % T, e4 = idx % hs4;
    const float4 val = part < 3 ? reinterpret_cast<const float4*>(QKV + (row0 + r) * ld + part * D.d + h * hs)[e4]
                                : reinterpret_cast<const float4*>(dY + (row0 + r) * D.d + h * hs)[e4];
    reinterpret_cast<float4*>((part == 0 ? q : part == 1 ? k : part == 2 ? v : dy) + r * lds)[e4] = val;
  }
  for (int r = threadIdx.x; r < 4 * T; r += kAttnThreads)      // zero the 4 pad columns of every row
    *reinterpret_cast<float4*>(q + r * lds + hs4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* Pg = P + ((size_t)b * D.H + h) * T * T;
  const float* dm = drop_mask ? drop_mask + ((size_t)b * D.H + h) * T * T : nullptr;
  for (int idx = threadIdx.x; idx < T * T; idx += kAttnThreads) {
    Pm[(idx / T) * ldt + idx % T] = Pg[idx];
    if (dm) Pd[(idx / T) * ldt + idx % T] = Pg[idx] * dm[idx];
  }
  if (!dm) Pd = Pm;
  __syncthreads();
  // dP[i][j] = dY_i . V_j ;  dS = P * (dP - sum_j dP * P) * scale
  const int T2 = (T + 1) >> 1;
  for (int idx = threadIdx.x; idx < T2 * T2; idx += kAttnThreads) {
    const int i0 = (idx / T2) * 2, j0 = (idx % T2) * 2;
    float o[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    if (j0 <= i0 + 1) tile2x2(dy, v, lds, hs4, i0, j0, T, o);
#pragma unroll
    for (int di = 0; di < 2; ++di)
#pragma unroll
      for (int dj = 0; dj < 2; ++dj) {
        const int i = i0 + di, j = j0 + dj;
        if (i < T && j < T) dS[i * ldt + j] = (j <= i) ? (dm ? o[di][dj] * dm[i * T + j] : o[di][dj]) : 0.f;
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T; i += kAttnThreads) {
    float sdp = 0.f;
    for (int j = 0; j <= i; ++j) sdp = fmaf(dS[i * ldt + j], Pm[i * ldt + j], sdp);
    for (int j = 0; j < T; ++j) dS[i * ldt + j] = (j <= i) ? Pm[i * ldt + j] * (dS[i * ldt + j] - sdp) * scale : 0.f;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < T * hs4; idx += kAttnThreads) {      // (row, 4 columns) per thread
    const int i = idx / hs4, e4 = idx % hs4;
    float4 dq = make_float4(0.f, 0.f, 0.f, 0.f), dk = dq, dv = dq;
    for (int j = 0; j <= i; ++j) {
      const float s_ = dS[i * ldt + j];
      const float4 kk = reinterpret_cast<const float4*>(k + j * lds)[e4];
      dq.x = fmaf(s_, kk.x, dq.x); dq.y = fmaf(s_, kk.y, dq.y); dq.z = fmaf(s_, kk.z, dq.z); dq.w = fmaf(s_, kk.w, dq.w);
    }
    for (int r = i; r < T; ++r) {
      const float s_ = dS[r * ldt + i], p_ = Pd[r * ldt + i];
      const float4 qq = reinterpret_cast<const float4*>(q + r * lds)[e4];
      const float4 dd = reinterpret_cast<const float4*>(dy + r * lds)[e4];
      dk.x = fmaf(s_, qq.x, dk.x); dk.y = fmaf(s_, qq.y, dk.y); dk.z = fmaf(s_, qq.z, dk.z); dk.w = fmaf(s_, qq.w, dk.w);
      dv.x = fmaf(p_, dd.x, dv.x); dv.y = fmaf(p_, dd.y, dv.y); dv.z = fmaf(p_, dd.z, dv.z); dv.w = fmaf(p_, dd.w, dv.w);
    }
    float* o = dQKV + (row0 + i) * ld + h * hs;
    reinterpret_cast<float4*>(o)[e4] = dq;
    reinterpret_cast<float4*>(o + D.d)[e4] = dk;
    reinterpret_cast<float4*>(o + 2 * D.d)[e4] = dv;
  }
}
inline size_t attn_bwd_smem(int T, int hs) { return ((size_t)4 * T * attn_lds(hs) + (size_t)3 * T * (T + 1)) * sizeof(float); }

__global__ void gather_action_rows_kernel(Dims D, const float* __restrict__ X, float* __restrict__ HA) {
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= (size_t)D.B * D.t * D.d) return;
  const int c = (int)(idx % D.d);
  const size_t it = idx / D.d;
  const int b = (int)(it / D.t), step = (int)(it % D.t);
  HA[idx] = X[((size_t)b * D.T + 1 + D.G + 2 * step + 1) * D.d + c];
}
__global__ void scatter_action_rows_kernel(Dims D, const float* __restrict__ dHA, float* __restrict__ dX) {
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= (size_t)D.M * D.d) return;
  const int c = (int)(idx % D.d);
  const int r = (int)(idx / D.d), b = r / D.T, tok = r - b * D.T, j = tok - 1 - D.G;
  dX[idx] = (j >= 0 && (j & 1)) ? dHA[((size_t)b * D.t + (j >> 1)) * D.d + c] : 0.f;
}
// target = (a - c_skip * x_noised) / c_out ; loss = mean((pred - target)^2) ; dpred = 2 (pred - target) / N
__global__ void loss_kernel(Dims D, const float* __restrict__ pred, const float* __restrict__ action,
                            const float* __restrict__ noise, const float* __restrict__ sigma, int pred_last,
                            float* __restrict__ dpred, float* __restrict__ partial) {
  __shared__ float s[kTB / 32];
  const size_t n = (size_t)D.B * D.t * D.act;
  const float invN = 1.0f / (float)(pred_last ? (size_t)D.B * D.act : n);
  float acc = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t it = i / D.act;
    const int b = (int)(it / D.t), step = (int)(it % D.t);
    const float sg = sigma[b], sd = D.sigma_data;
    const float nz = (pred_last && step != D.t - 1) ? 0.f : noise[i];
    const float xn = action[i] + nz * sg;
    const float den = sg * sg + sd * sd;
    const float c_skip = sd * sd / den, c_out = sg * sd / sqrtf(den);
    const float target = (action[i] - c_skip * xn) / c_out;
    float diff = pred[i] - target;
    if (pred_last && step != D.t - 1) diff = 0.f;
    acc = fmaf(diff, diff, acc);
    dpred[i] = 2.0f * diff * invN;
  }
  acc = wsum(acc);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < kTB / 32; ++i) t += s[i]; partial[blockIdx.x] = t * invN; }
}
__global__ void sum_partials_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
  float a = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) a += partial[i];
  a = wsum(a);
  if (threadIdx.x == 0) *out = a;
}

// ---- embedding backward --------------------------------------------------------------------------------------------
// dpos[p][c] = sum over b of dX rows at position p (goal token p < G; state and action token of step p - G)
// d pos_emb[p][c] = sum over sequences of the rows that use position p (one goal row, or a state and an action
// row): partial sums over kPosChunks sequence chunks, then colsum_reduce-style reduction (deterministic).
constexpr int kPosChunks = 64;
__global__ void pos_grad_partial_kernel(Dims D, const float* __restrict__ dX, int n_pos, float* __restrict__ partial) {
  const int c = threadIdx.x, p = blockIdx.x, chunk = blockIdx.y;      // blockDim.x == d
  float a = 0.f;
  if (p < D.G + D.t) {
    for (int b = chunk; b < D.B; b += kPosChunks) {
      const size_t r0 = (size_t)b * D.T;
      if (p < D.G) a += dX[(r0 + 1 + p) * D.d + c];
      else { const int step = p - D.G; a += dX[(r0 + 1 + D.G + 2 * step) * D.d + c] + dX[(r0 + 2 + D.G + 2 * step) * D.d + c]; }
    }
  }
  partial[((size_t)chunk * n_pos + p) * D.d + c] = a;
}
__global__ void pos_grad_reduce_kernel(const float* __restrict__ partial, int n, float* __restrict__ dpos) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  float a = 0.f;
  for (int chunk = 0; chunk < kPosChunks; ++chunk) a += partial[(size_t)chunk * n + idx];
  dpos[idx] = a;
}
// gathers for the embedding weight gradients: rows of kind 0 = sigma token, 1 = state + goal tokens, 2 = action tokens
__global__ void gather_embed_rows_kernel(Dims D, int kind, const float* __restrict__ dX, const float* __restrict__ state,
                                         const float* __restrict__ goal, const float* __restrict__ goal_keep,
                                         const float* __restrict__ xin, const float* __restrict__ sigma,
                                         float* __restrict__ dRows, float* __restrict__ inRows) {
  const int per = kind == 0 ? 1 : (kind == 1 ? D.G + D.t : D.t);
  const int kin = kind == 0 ? 1 : (kind == 1 ? D.obs : D.act);
  const size_t n_rows = (size_t)D.B * per;
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= n_rows * (size_t)(D.d + kin)) return;
  const size_t row = idx / (D.d + kin);
  const int c = (int)(idx % (D.d + kin));
  const int b = (int)(row / per), q = (int)(row % per);
  int tok;
  if (kind == 0) tok = 0;
  else if (kind == 1) tok = q < D.G ? 1 + q : 1 + D.G + 2 * (q - D.G);
  else tok = 2 + D.G + 2 * q;
  if (c < D.d) { dRows[row * D.d + c] = dX[((size_t)b * D.T + tok) * D.d + c]; return; }
  const int k = c - D.d;
  float v;
  if (kind == 0) v = logf(sigma[b]) / 4.0f;
  else if (kind == 1) {
    if (q < D.G) { const size_t o = ((size_t)b * D.G + q) * D.obs + k; v = goal_keep ? goal[o] * goal_keep[o] : goal[o]; }
    else v = state[((size_t)b * D.t + (q - D.G)) * D.obs + k];
  } else v = xin[((size_t)b * D.t + q) * D.act + k];
  inRows[row * kin + k] = v;
}

size_t align4(size_t n) { return (n + 3) & ~size_t(3); }
// scratch for the two-stage reductions: column sums (64 row blocks x up to 4096 columns), position gradients
constexpr size_t kPartialFloats = (size_t)148 * 2048;

}  // namespace

}  // namespace beso

// ================================ NCCL communicator (data-parallel gradient exchange) ============================
struct beso_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
  cudaStream_t stream = nullptr;        // the exchange runs here, behind the backward kernels of the compute stream
  std::vector<cudaEvent_t> events;      // one per gradient bucket + the join event
  int next_event = 0;
};

namespace beso {
namespace {
__global__ void scale_kernel(float* x, size_t n, float s) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] *= s;
}
int nccl_fail(ncclResult_t r, const char* what) {
  set_error(std::string("NCCL error: ") + ncclGetErrorString(r) + " at " + what);
  return BESO_E_NCCL;
}
}  // namespace
#define BESO_NCCL(expr) do { ncclResult_t _r = (expr); if (_r != ncclSuccess) return nccl_fail(_r, #expr); } while (0)

// One gradient bucket [off, off + n) of the flat buffer is final on the compute stream `st`: all-reduce(sum) it on the
// communicator's stream (then * scale), behind an event, while `st` goes on with the next layer's backward.
// A second range (off2, n2) rides in the same NCCL group (one launch).  scale == 1 / world uses the collective's own
// averaging (ncclAvg); any other scale is a separate pass over the bucket.
static int sync_bucket(beso_comm* c, float* grad, size_t off, size_t n, float scale, cudaStream_t st, size_t off2 = 0, size_t n2 = 0) {
  if (!c || c->world <= 1 || n == 0) return BESO_OK;
  if (c->next_event >= (int)c->events.size()) { set_error("internal: out of gradient-bucket events"); return BESO_E_INVALID; }
  cudaEvent_t ev = c->events[c->next_event++];
  BESO_CUDA(cudaEventRecord(ev, st));
  BESO_CUDA(cudaStreamWaitEvent(c->stream, ev, 0));
  const bool avg = fabsf(scale * (float)c->world - 1.0f) < 1e-6f;
  const ncclRedOp_t op = avg ? ncclAvg : ncclSum;
  BESO_NCCL(ncclGroupStart());
  BESO_NCCL(ncclAllReduce(grad + off, grad + off, n, ncclFloat, op, c->comm, c->stream));
  if (n2) BESO_NCCL(ncclAllReduce(grad + off2, grad + off2, n2, ncclFloat, op, c->comm, c->stream));
  BESO_NCCL(ncclGroupEnd());
  if (!avg && scale != 1.0f) {
    for (int k = 0; k < (n2 ? 2 : 1); ++k) {
      const size_t o = k ? off2 : off, cnt = k ? n2 : n;
      const int grid = (int)((cnt + 255) / 256 < 1024 ? (cnt + 255) / 256 : 1024);
      scale_kernel<<<grid, 256, 0, c->stream>>>(grad + o, cnt, scale);
      ++g_kernel_launches;
    }
    BESO_CUDA(cudaGetLastError());
  }
  return BESO_OK;
}
// the compute stream waits for every bucket issued since the last join
static int join_buckets(beso_comm* c, cudaStream_t st) {
  if (!c || c->world <= 1) return BESO_OK;
  cudaEvent_t ev = c->events.back();
  BESO_CUDA(cudaEventRecord(ev, c->stream));
  BESO_CUDA(cudaStreamWaitEvent(st, ev, 0));
  c->next_event = 0;
  return BESO_OK;
}

// ================================ training workspace =======================================================
struct TrainWs {
  float* buf = nullptr;
  size_t floats = 0;
  GemmWs gemm{};
  int sm_count = 0;
};

static int ensure_ws(TrainWs*& ws, size_t floats) {
  if (!ws) ws = new TrainWs();
  if (!ws->sm_count) {
    int dev = 0;
    BESO_CUDA(cudaGetDevice(&dev));
    BESO_CUDA(cudaDeviceGetAttribute(&ws->sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  if (floats > ws->floats) {
    if (ws->buf) cudaFree(ws->buf);
    ws->buf = nullptr; ws->floats = 0;
    BESO_CUDA(cudaMalloc(&ws->buf, floats * sizeof(float)));
    ws->floats = floats;
  }
  return BESO_OK;
}
void train_ws_free(TrainWs* ws) {
  if (!ws) return;
  if (ws->buf) cudaFree(ws->buf);
  gemm_ws_free(ws->gemm);
  delete ws;
}

int train_gemm(TrainWs*& ws, const GemmArgs& a, cudaStream_t st) {
  int rc = ensure_ws(ws, 0);
  if (rc) return rc;
  return gemm_run(a, ws->gemm, ws->sm_count, st);
}

int train_loss_fwd_bwd(TrainWs*& ws, const beso_model_desc& m, const float* const* prm, const float* state,
                       const float* action, const float* goal, const float* noise, const float* sigma,
                       const float* goal_keep, const beso_dropout_masks* drop, float* loss_out, float* grad, int B,
                       uint32_t flags, cudaStream_t st, beso_comm* comm, float grad_scale) {
  if (!m.linear_output) { set_error("training path supports linear_output models only"); return BESO_E_UNSUPPORTED; }
  Dims D;
  D.B = B; D.t = m.window; D.G = m.goal_conditioned ? m.goal_len : 0; D.T = 1 + D.G + 2 * D.t; D.obs = m.obs_dim;
  D.act = m.act_dim; D.d = m.d; D.H = m.n_heads; D.hs = m.d / m.n_heads; D.L = m.n_layers; D.F = 4 * m.d;
  D.M = B * D.T; D.sigma_data = m.sigma_data;
  if (D.T > 64) { set_error("training path supports at most 64 tokens per sequence"); return BESO_E_UNSUPPORTED; }
  const int pred_last = (flags & BESO_FLAG_PRED_LAST) ? 1 : 0;
  const int prec = (flags & BESO_FLAG_TRAIN_FAST) ? 0 : ((flags & BESO_FLAG_TRAIN_SPLIT2) ? 1 : 2);
  const int d = D.d, F = D.F, M = D.M, L = D.L;
  const size_t Md = (size_t)M * d, MF = (size_t)M * F, nP = (size_t)B * D.H * D.T * D.T, nBt = (size_t)B * D.t;
  // ---- carve the workspace ----
  const size_t per_layer = 5 * align4(Md) + align4(3 * Md) + align4(nP) + 2 * align4(MF) + 2 * align4(2 * (size_t)M);
  const size_t n_gather = (size_t)B * (D.G + D.t);
  const size_t total = (size_t)L * per_layer + 2 * align4(Md) + align4(2 * (size_t)M) + align4(nBt * d) + 3 * align4(nBt * D.act) +
                       5 * align4(Md) + align4(MF) + align4(3 * Md) + align4(n_gather * (size_t)(d + (D.obs > D.act ? D.obs : D.act))) + kPartialFloats + 4096;
  int rc = ensure_ws(ws, total);
  if (rc) return rc;
  float* p = ws->buf;
  auto take = [&](size_t n) { float* q = p; p += align4(n); return q; };
  struct Layer { float *Xin, *H1, *QKV, *P, *Y, *Xmid, *H2, *U, *Gg, *st1, *st2; };
  std::vector<Layer> A(L);
  for (int l = 0; l < L; ++l) {
    A[l].Xin = take(Md); A[l].H1 = take(Md); A[l].QKV = take(3 * Md); A[l].P = take(nP); A[l].Y = take(Md);
    A[l].Xmid = take(Md); A[l].H2 = take(Md); A[l].U = take(MF); A[l].Gg = take(MF); A[l].st1 = take(2 * (size_t)M);
    A[l].st2 = take(2 * (size_t)M);
  }
  float *XL = take(Md), *HF = take(Md), *stf = take(2 * (size_t)M), *HA = take(nBt * d), *pred = take(nBt * D.act),
        *dpred = take(nBt * D.act), *xin = take(nBt * D.act);
  float *dX = take(Md), *dT = take(Md), *dH = take(Md), *dY = take(Md), *dXm = take(Md), *dBig = take(MF), *dQKV = take(3 * Md);
  float *gRows = take(n_gather * (size_t)(d + (D.obs > D.act ? D.obs : D.act))), *partial = take(kPartialFloats);

  // ---- parameter pointers and gradient slots (parameters() order) ----
  std::vector<size_t> goff;
  size_t acc_off = 0;
  const int n_params = 3 + 16 * L + 8;
  {
    beso_model_desc mm = m;
    for (int i = 0; i < n_params; ++i) { goff.push_back(acc_off); acc_off += (size_t)beso_param_numel(&mm, i); }
  }
  auto W = [&](int i) { return prm[i]; };
  auto Gp = [&](int i) { return grad ? grad + goff[i] : nullptr; };
  auto lp = [&](int l, int k) { return 3 + 16 * l + k; };
  const int pt = 3 + 16 * L;
  const float* m_embed = drop ? drop->embed : nullptr;
  auto m_attn = [&](int l) -> const float* { return (drop && drop->attn) ? drop->attn[l] : nullptr; };
  auto m_res1 = [&](int l) -> const float* { return (drop && drop->resid_attn) ? drop->resid_attn[l] : nullptr; };
  auto m_res2 = [&](int l) -> const float* { return (drop && drop->resid_mlp) ? drop->resid_mlp[l] : nullptr; };
#define LAUNCH(kernel, grid, block, smem, ...)                       \
  do {                                                               \
    kernel<<<grid, block, smem, st>>>(__VA_ARGS__);                  \
    ++g_kernel_launches;                                             \
  } while (0)
  // C[M][N] = op(A) op(B) on the tensor cores.  ta: A is stored [K][M] (lda), tb: B is stored [N][K] (ldb).
  struct Epi { const float* bias = nullptr; const float* resid = nullptr; int ldr = 0; const float* mul = nullptr; int ldm = 0;
               int accumulate = 0; float* gelu = nullptr; int ldg = 0; };
  auto gemm = [&](bool ta, bool tb, int Mm, int Nn, int Kk, const float* Am, int lda, const float* Bm, int ldb, float* Cm, int ldc,
                  const Epi& e) -> int {
    GemmArgs a{};
    a.A = Am; a.lda = lda; a.a_kmajor = ta ? 0 : 1;
    a.B = Bm; a.ldb = ldb; a.b_kmajor = tb ? 1 : 0;
    a.C = Cm; a.ldc = ldc; a.M = Mm; a.N = Nn; a.K = Kk;
    a.bias = e.bias; a.resid = e.resid; a.ldr = e.ldr; a.mul = e.mul; a.ldm = e.ldm; a.accumulate = e.accumulate;
    a.gelu_out = e.gelu; a.ldg = e.ldg; a.prec = prec;
    return gemm_run(a, ws->gemm, ws->sm_count, st);
  };
#define GEMM(...) do { if ((rc = gemm(__VA_ARGS__))) return rc; } while (0)
  const Epi none{};

  // ============================== forward ==============================
  if (attn_bwd_smem(D.T, D.hs) > 48 * 1024) {
    BESO_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_fwd_smem(D.T, D.hs)));
    BESO_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_bwd_smem(D.T, D.hs)));
  }
  if ((d & 3) || (F & 3) || (D.hs & 3)) { set_error("training path needs embed_dim % 4 == 0 and head size % 4 == 0"); return BESO_E_UNSUPPORTED; }
  LAUNCH(embed_fwd_kernel, blocks_for(Md), kTB, 0, D, state, action, goal, noise, sigma, goal_keep, pred_last, W(0), W(1), W(2),
         W(pt + 2), W(pt + 3), W(pt + 4), W(pt + 5), m_embed, A[0].Xin, xin);
  const int ln_grid = (M + 7) / 8;
  for (int l = 0; l < L; ++l) {
    Layer& a = A[l];
    LAUNCH(ln_fwd_kernel, ln_grid, 256, 0, a.Xin, M, d, W(lp(l, 0)), W(lp(l, 1)), a.H1, a.st1);
    // q | k | v column blocks (reference parameter order: key, query, value)
    GEMM(false, true, M, d, d, a.H1, d, W(lp(l, 6)), d, a.QKV, 3 * d, none);
    GEMM(false, true, M, d, d, a.H1, d, W(lp(l, 4)), d, a.QKV + d, 3 * d, none);
    GEMM(false, true, M, d, d, a.H1, d, W(lp(l, 8)), d, a.QKV + 2 * d, 3 * d, none);
    // q / k / v biases are added (and written back) by the attention kernel's tile load
    LAUNCH(attn_fwd_kernel, B * D.H, kAttnThreads, attn_fwd_smem(D.T, D.hs), D, a.QKV, W(lp(l, 7)), W(lp(l, 5)), W(lp(l, 9)), m_attn(l), a.P, a.Y);
    { Epi e; e.bias = W(lp(l, 11)); e.resid = a.Xin; e.ldr = d; e.mul = m_res1(l); e.ldm = d;     // X_mid = X_in + drop(Y Wp^T + b_proj)
      GEMM(false, true, M, d, d, a.Y, d, W(lp(l, 10)), d, a.Xmid, d, e); }
    LAUNCH(ln_fwd_kernel, ln_grid, 256, 0, a.Xmid, M, d, W(lp(l, 2)), W(lp(l, 3)), a.H2, a.st2);
    { Epi e; e.bias = W(lp(l, 13)); e.gelu = a.Gg; e.ldg = F;                                     // U = H2 W1^T + b1 (kept), G = gelu(U)
      GEMM(false, true, M, F, d, a.H2, d, W(lp(l, 12)), d, a.U, F, e); }
    float* Xnext = (l + 1 < L) ? A[l + 1].Xin : XL;
    { Epi e; e.bias = W(lp(l, 15)); e.resid = a.Xmid; e.ldr = d; e.mul = m_res2(l); e.ldm = d;    // X_next = X_mid + drop(G W2^T + b_2)
      GEMM(false, true, M, d, F, a.Gg, F, W(lp(l, 14)), F, Xnext, d, e); }
  }
  LAUNCH(ln_fwd_kernel, ln_grid, 256, 0, XL, M, d, W(pt), W(pt + 1), HF, stf);
  LAUNCH(gather_action_rows_kernel, blocks_for(nBt * d), kTB, 0, D, HF, HA);
  { Epi e; e.bias = W(pt + 7);
    GEMM(false, true, (int)nBt, D.act, d, HA, d, W(pt + 6), d, pred, D.act, e); }
  const int lgrid = (int)(blocks_for(nBt * D.act) < 1024 ? blocks_for(nBt * D.act) : 1024);
  LAUNCH(loss_kernel, lgrid, kTB, 0, D, pred, action, noise, sigma, pred_last, dpred, partial);
  LAUNCH(sum_partials_kernel, 1, 32, 0, partial, lgrid, loss_out);
  BESO_CUDA(cudaGetLastError());
  if (!grad) return BESO_OK;

  // ============================== backward ==============================
  // column sums (bias gradients): two deterministic stages, bandwidth-bound, full-chip parallel
  auto colsum = [&](const float* Am, int rows, int N, int lda, float* out) -> int {
    if ((size_t)N * kColsumRowBlocks > kPartialFloats) { set_error("internal: column-sum scratch too small"); return BESO_E_INVALID; }
    LAUNCH(colsum_partial_kernel, dim3((N + 31) / 32, kColsumRowBlocks), dim3(32, 8), 0, Am, rows, N, lda, partial);
    LAUNCH(colsum_reduce_kernel, (N + 127) / 128, 128, 0, partial, N, out);
    return BESO_OK;
  };
#define COLSUM(...) do { if ((rc = colsum(__VA_ARGS__))) return rc; } while (0)
  // LayerNorm backward: dX (+)= ..., then weight / bias gradients from column sums
  auto ln_bwd = [&](const float* dYv, const float* Xv, const float* stv, const float* wv, int accumulate, float* dw, float* db) -> int {
    LAUNCH(ln_bwd_kernel, ln_grid, 256, 0, dYv, Xv, stv, wv, M, d, dX, accumulate, dT);
    int r2 = colsum(dT, M, d, d, dw);
    if (r2) return r2;
    return colsum(dYv, M, d, d, db);
  };
#define LNBWD(...) do { if ((rc = ln_bwd(__VA_ARGS__))) return rc; } while (0)
  // gradient of a dropped-out residual branch: dX .* mask (the branch output was scaled by the mask)
  auto branch_grad = [&](const float* mask) -> const float* {
    if (!mask) return dX;
    LAUNCH(mul_kernel, blocks_for(Md / 4), kTB, 0, dXm, dX, mask, Md / 4);
    return dXm;
  };
  Epi acc1; acc1.accumulate = 1;
  // head
  GEMM(true, false, D.act, d, (int)nBt, dpred, D.act, HA, d, Gp(pt + 6), d, none);
  COLSUM(dpred, (int)nBt, D.act, D.act, Gp(pt + 7));
  GEMM(false, false, (int)nBt, d, D.act, dpred, D.act, W(pt + 6), d, HA, d, none);               // HA <- dHA
  LAUNCH(scatter_action_rows_kernel, blocks_for(Md), kTB, 0, D, HA, dH);                         // dH <- dHF
  LNBWD(dH, XL, stf, W(pt), 0, Gp(pt), Gp(pt + 1));                                              // dX = d loss / d X_L
  for (int l = L - 1; l >= 0; --l) {
    Layer& a = A[l];
    // ---- MLP branch: X_next = X_mid + drop(gelu(H2 W1^T + b1) W2^T + b2) ----
    const float* dB = branch_grad(m_res2(l));
    GEMM(true, false, d, F, M, dB, d, a.Gg, F, Gp(lp(l, 14)), F, none);                            // dW2 = dB^T G
    COLSUM(dB, M, d, d, Gp(lp(l, 15)));
    GEMM(false, false, M, F, d, dB, d, W(lp(l, 14)), F, dBig, F, none);                            // dG = dB W2
    LAUNCH(gelu_bwd_kernel, blocks_for(MF / 4), kTB, 0, a.U, dBig, MF / 4);                       // dU
    GEMM(true, false, F, d, M, dBig, F, a.H2, d, Gp(lp(l, 12)), d, none);                          // dW1 = dU^T H2
    COLSUM(dBig, M, F, F, Gp(lp(l, 13)));
    GEMM(false, false, M, d, F, dBig, F, W(lp(l, 12)), d, dH, d, none);                            // dH2 = dU W1
    LNBWD(dH, a.Xmid, a.st2, W(lp(l, 2)), 1, Gp(lp(l, 2)), Gp(lp(l, 3)));                          // dX = d/dX_mid
    // ---- attention branch: X_mid = X_in + drop(Y Wp^T + bp) ----
    dB = branch_grad(m_res1(l));
    GEMM(true, false, d, d, M, dB, d, a.Y, d, Gp(lp(l, 10)), d, none);                             // dWp
    COLSUM(dB, M, d, d, Gp(lp(l, 11)));
    GEMM(false, false, M, d, d, dB, d, W(lp(l, 10)), d, dY, d, none);                              // dY = dB Wp
    LAUNCH(attn_bwd_kernel, B * D.H, kAttnThreads, attn_bwd_smem(D.T, D.hs), D, a.QKV, a.P, m_attn(l), dY, dQKV);
    GEMM(true, false, d, d, M, dQKV, 3 * d, a.H1, d, Gp(lp(l, 6)), d, none);                       // dWq
    GEMM(true, false, d, d, M, dQKV + d, 3 * d, a.H1, d, Gp(lp(l, 4)), d, none);                   // dWk
    GEMM(true, false, d, d, M, dQKV + 2 * d, 3 * d, a.H1, d, Gp(lp(l, 8)), d, none);               // dWv
    COLSUM(dQKV, M, d, 3 * d, Gp(lp(l, 7)));
    COLSUM(dQKV + d, M, d, 3 * d, Gp(lp(l, 5)));
    COLSUM(dQKV + 2 * d, M, d, 3 * d, Gp(lp(l, 9)));
    GEMM(false, false, M, d, d, dQKV, 3 * d, W(lp(l, 6)), d, dH, d, none);                         // dH1
    GEMM(false, false, M, d, d, dQKV + d, 3 * d, W(lp(l, 4)), d, dH, d, acc1);
    GEMM(false, false, M, d, d, dQKV + 2 * d, 3 * d, W(lp(l, 8)), d, dH, d, acc1);
    LNBWD(dH, a.Xin, a.st1, W(lp(l, 0)), 1, Gp(lp(l, 0)), Gp(lp(l, 1)));                           // dX = d/dX_in
    // data parallel: this block's 16 gradient tensors are final -- their all-reduce overlaps the blocks below
    if ((rc = sync_bucket(comm, grad, goff[lp(l, 0)], goff[lp(l + 1, 0)] - goff[lp(l, 0)], grad_scale, st))) return rc;
  }
  // ---- embeddings ----
  if (m_embed) LAUNCH(mul_kernel, blocks_for(Md / 4), kTB, 0, dX, dX, m_embed, Md / 4);           // X_0 = drop(embeddings)
  const int n_pos = D.G + D.t + 1;
  if ((size_t)kPosChunks * n_pos * d > kPartialFloats || d > 1024) { set_error("internal: position-gradient scratch too small"); return BESO_E_INVALID; }
  LAUNCH(pos_grad_partial_kernel, dim3(n_pos, kPosChunks), d, 0, D, dX, n_pos, partial);
  LAUNCH(pos_grad_reduce_kernel, blocks_for((size_t)n_pos * d), kTB, 0, partial, n_pos * d, Gp(0));
  struct Kind { int kind, per, kin, pw, pb; };
  const Kind kinds[3] = {{0, 1, 1, pt + 2, pt + 3}, {1, D.G + D.t, D.obs, 1, 2}, {2, D.t, D.act, pt + 4, pt + 5}};
  for (const Kind& k : kinds) {
    const size_t rows = (size_t)B * k.per;
    float* dRows = gRows;
    float* inRows = gRows + rows * d;
    LAUNCH(gather_embed_rows_kernel, blocks_for(rows * (d + k.kin)), kTB, 0, D, k.kind, dX, state, goal, goal_keep, xin, sigma,
           dRows, inRows);
    GEMM(true, false, d, k.kin, (int)rows, dRows, d, inRows, k.kin, Gp(k.pw), k.kin, none);        // dW = dRows^T in
    COLSUM(dRows, (int)rows, d, d, Gp(k.pb));
  }
  BESO_CUDA(cudaGetLastError());
  if (comm && comm->world > 1) {        // embeddings (first 3 tensors) and the tail (ln_f, sigma / action embeddings, head)
    if ((rc = sync_bucket(comm, grad, 0, goff[3], grad_scale, st, goff[pt], acc_off - goff[pt]))) return rc;
    if ((rc = join_buckets(comm, st))) return rc;
  }
#undef LAUNCH
#undef GEMM
#undef COLSUM
#undef LNBWD
  return BESO_OK;
}

}  // namespace beso

// ================================ NCCL gradient exchange: C ABI ================================================
using beso::nccl_fail;
using beso::scale_kernel;

extern "C" {

int beso_comm_unique_id(char* out128) {
  if (!out128) { beso::set_error("null buffer"); return BESO_E_INVALID; }
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  BESO_NCCL(ncclGetUniqueId(&id));
  memcpy(out128, &id, 128);
  return BESO_OK;
}

int beso_comm_init(int rank, int world, const char* unique_id128, int device, beso_comm** out) {
  if (!out || !unique_id128 || world < 1 || rank < 0 || rank >= world) { beso::set_error("bad comm arguments"); return BESO_E_INVALID; }
  BESO_CUDA(cudaSetDevice(device));
  ncclUniqueId id;
  memcpy(&id, unique_id128, 128);
  beso_comm* c = new beso_comm();
  c->rank = rank; c->world = world; c->device = device;
  ncclResult_t r = ncclCommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) { delete c; return nccl_fail(r, "ncclCommInitRank"); }
  BESO_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->events.resize(beso::kMaxLayers + 4);
  for (auto& e : c->events) BESO_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  *out = c;
  return BESO_OK;
}

int beso_comm_destroy(beso_comm* c) {
  if (!c) return BESO_OK;
  if (c->comm) ncclCommDestroy(c->comm);
  for (auto& e : c->events) cudaEventDestroy(e);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return BESO_OK;
}

// One all-reduce(sum) of the flat gradient, then *scale (1/world for the global-batch mean).
int beso_allreduce_grads(beso_comm* c, float* flat_grad_dev, size_t n, float scale, void* stream) {
  if (!c || !flat_grad_dev) { beso::set_error("null comm or buffer"); return BESO_E_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  if (c->world > 1) BESO_NCCL(ncclAllReduce(flat_grad_dev, flat_grad_dev, n, ncclFloat, ncclSum, c->comm, st));
  if (scale != 1.0f) {
    const int grid = (int)((n + 255) / 256 < 2048 ? (n + 255) / 256 : 2048);
    scale_kernel<<<grid > 0 ? grid : 1, 256, 0, st>>>(flat_grad_dev, n, scale);
    ++beso::g_kernel_launches;
    BESO_CUDA(cudaGetLastError());
  }
  return BESO_OK;
}

}  // extern "C"
