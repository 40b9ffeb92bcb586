// PRECISE mode: the whole GCDenoiser -> DiffusionGPT forward and the DDIM / Euler / Heun sample
// loop as ONE fused fp32 CUDA-core kernel.  Every activation lives in shared memory for the whole
// launch; weights are streamed from L2 as pre-transposed [K][N] fp32 images (coalesced float4).
//
// Reference semantics (paths relative to beso/agents/diffusion_agents/k_diffusion/):
//   pre-conditioning   score_wrappers.py:31-43, 81-96
//   score-GPT forward  score_gpts.py:272-358 (attention :50-80, block :96-115)
//   samplers           gc_sampling.py:167-213 (Euler), 259-314 (Heun), 895-924 (DDIM)
//   CFG mix            classifier_free_sampler.py:35-49
//
// Generic in (obs, act, W, G, d, L, H): also serves the shipped checkpoints (d=360 / d=240).
#include "common.cuh"

namespace beso {
namespace {

constexpr int kThreads = 512;   // 16 warps, 4 per scheduler: at most 128 registers per thread
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ float gelu_erf(float x) {           // nn.GELU() default (score_gpts.py:107)
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

enum { EPI_STORE = 0, EPI_GELU = 1, EPI_RESID = 2 };

// Shared-memory accesses of the GEMM go through explicit ld.shared / st.shared: through a plain pointer the
// compiler emits generic loads with 64-bit address arithmetic and cannot tell they never alias the weights.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f2(uint32_t a, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}

// out[r][n] (=|+=) bias[n] + sum_k A[r][k] * W[k][n]       r < R, n < N (N % 2 == 0, K % 4 == 0)
// A, out in shared memory; bias in global memory; Wt = the weight tiled by 64-column groups, [ceil(N/64)][K][64]
// (L2-resident).  A warp owns an RT-row x 64-column item, each lane RT x 2 accumulators, so a weight element
// fetched from L2 feeds RT rows (with R <= 24 and RT = 12 the weight set crosses the L2 -> SM link twice per
// evaluation).  Weights are requested one k-quad ahead and moved into place at the end of the quad (the moves pin
// the wait there; left free, the scheduler sinks the loads next to their first use); the A rows are reloaded in
// place, also a quad ahead; the other warps of the scheduler cover what latency is left.
// When a GEMM has fewer items than half the warps (the two N = d GEMMs) and the caller passes a scratch buffer,
// K is split in two: the second half's partial sums go to the scratch and are added after a CTA barrier.
template <int EPI, int RT>
__device__ __noinline__ void gemm_rows_rt(const float* A, int lda, int R, const float* __restrict__ Wt,
                                          const float* __restrict__ bias, int K, int N, float* out, int ldo,
                                          float* scratch, int lds) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncg = (N + 63) >> 6, nrg = (R + RT - 1) / RT, tiles = ncg * nrg;
  const int KS = (scratch != nullptr && tiles * 2 <= kWarps && (K & 7) == 0) ? 2 : 1;
  const int Kh = K / KS;
  const uint32_t a_s = smem_addr(A), o_s = smem_addr(out), p_s = scratch ? smem_addr(scratch) : 0u;
  for (int item = warp; item < tiles * KS; item += kWarps) {
    const int ks = item / tiles, tile = item - ks * tiles;
    const int cg = tile % ncg, rg = tile / ncg;
    const int k0 = ks * Kh, k1 = k0 + Kh;
    const int n0 = cg * 64 + lane * 2;
    const bool active = n0 < N;
    const int r0 = rg * RT;
    uint32_t ar[RT];
#pragma unroll
    for (int i = 0; i < RT; ++i) ar[i] = a_s + (uint32_t)(min(r0 + i, R - 1) * lda) * 4u;
    float acc[RT][2];
#pragma unroll
    for (int i = 0; i < RT; ++i) acc[i][0] = acc[i][1] = 0.f;
    // this lane's column pair of weight row k is wq[k * 32] (padding columns of the last group are zero)
    const float2* wq = reinterpret_cast<const float2*>(Wt + (size_t)cg * K * 64) + lane;
    auto ldw4 = [&](float2 (&w)[4], int k) {             // rows k .. k+3 (256 bytes apart)
#pragma unroll
      for (int i = 0; i < 4; ++i) w[i] = __ldg(wq + (size_t)(k + i) * 32);
    };
    // one k-quad: row i's 8 FMAs, then row i's A registers are reloaded for the NEXT quad (kn), so a single set of
    // A registers gives a whole quad of distance between the shared-memory load and its use
    float4 a[RT];
    auto quad = [&](const float2 (&w)[4], int kn) {
#pragma unroll
      for (int i = 0; i < RT; ++i) {
        const float4 v = a[i];
        acc[i][0] = fmaf(v.x, w[0].x, acc[i][0]); acc[i][1] = fmaf(v.x, w[0].y, acc[i][1]);
        acc[i][0] = fmaf(v.y, w[1].x, acc[i][0]); acc[i][1] = fmaf(v.y, w[1].y, acc[i][1]);
        acc[i][0] = fmaf(v.z, w[2].x, acc[i][0]); acc[i][1] = fmaf(v.z, w[2].y, acc[i][1]);
        acc[i][0] = fmaf(v.w, w[3].x, acc[i][0]); acc[i][1] = fmaf(v.w, w[3].y, acc[i][1]);
        a[i] = lds_f4(ar[i] + (uint32_t)kn * 4u);
      }
    };
    float2 w[4], v[4];
    ldw4(w, k0);
#pragma unroll
    for (int i = 0; i < RT; ++i) a[i] = lds_f4(ar[i] + (uint32_t)k0 * 4u);
    int k = k0;
#pragma unroll 1
    for (; k + 8 <= k1; k += 4) {
      ldw4(v, k + 4);
      quad(w, k + 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) w[i] = v[i];
    }
    quad(w, k);                                           // the last quad (its reload of A is a harmless repeat)
    if (active) {
      if (ks == 0) {
        const float2 b = __ldg(reinterpret_cast<const float2*>(bias + n0));
#pragma unroll
        for (int i = 0; i < RT; ++i) {
          const int r = r0 + i;
          if (r < R) {
            float2 y = make_float2(acc[i][0] + b.x, acc[i][1] + b.y);
            const uint32_t o = o_s + (uint32_t)(r * ldo + n0) * 4u;
            if (EPI == EPI_GELU) {
              y.x = gelu_erf(y.x); y.y = gelu_erf(y.y);
            } else if (EPI == EPI_RESID) {
              const float2 x = lds_f2(o);
              y.x += x.x; y.y += x.y;
            }
            sts_f2(o, y);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < RT; ++i) {
          const int r = r0 + i;
          if (r < R) sts_f2(p_s + (uint32_t)(r * lds + n0) * 4u, make_float2(acc[i][0], acc[i][1]));
        }
      }
    }
  }
  if (KS == 2) {                                          // CTA-uniform
    __syncthreads();
    const int n2 = N >> 1;
    for (int idx = threadIdx.x; idx < R * n2; idx += kThreads) {
      const int r = idx / n2, c = (idx - r * n2) * 2;
      const uint32_t o = o_s + (uint32_t)(r * ldo + c) * 4u;
      const float2 x = lds_f2(o), p = lds_f2(p_s + (uint32_t)(r * lds + c) * 4u);
      sts_f2(o, make_float2(x.x + p.x, x.y + p.y));
    }
  }
}

// Row-tile height by the number of padded rows it costs (R = 23 -> 2 x 12, R = 32 -> 4 x 8).
template <int EPI>
__device__ __forceinline__ void gemm_rows(const float* A, int lda, int R, const float* __restrict__ Wt,
                                          const float* __restrict__ bias, int K, int N, float* out, int ldo,
                                          float* scratch = nullptr, int lds = 0) {
  const int pad12 = (R + 11) / 12 * 12 - R, pad8 = (R + 7) / 8 * 8 - R;
  if (pad12 <= pad8) gemm_rows_rt<EPI, 12>(A, lda, R, Wt, bias, K, N, out, ldo, scratch, lds);
  else gemm_rows_rt<EPI, 8>(A, lda, R, Wt, bias, K, N, out, ldo, scratch, lds);
}

// nn.LayerNorm(d), eps 1e-5, biased variance: one warp per row.
__device__ __noinline__ void layer_norm_rows(const float* X, int ldx, int R, int d, const float* __restrict__ w,
                                const float* __restrict__ b, float* out, int ldo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < R; r += kWarps) {
    const float* x = X + (size_t)r * ldx;
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s += x[c];
    const float mean = warp_sum(s) / (float)d;
    float v = 0.f;
    for (int c = lane; c < d; c += 32) { const float t = x[c] - mean; v = fmaf(t, t, v); }
    const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)d + 1e-5f);
    for (int c = lane; c < d; c += 32) out[(size_t)r * ldo + c] = (x[c] - mean) * rstd * __ldg(w + c) + __ldg(b + c);
  }
}

// Causal softmax attention over the tokens of each sequence (score_gpts.py:69-79).
// qkv: [R][ldq] with q | k | v at column 0 | d | 2d; y -> out[R][ldo] heads side by side.
__device__ __noinline__ void attention_rows(const float* qkv, int ldq, int ns, int T, int d, int H, int hs,
                               float* out, int ldo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float scale = 1.0f / sqrtf((float)hs);
  const int items = ns * H * T;
  for (int it = warp; it < items; it += kWarps) {
    const int i = it % T, h = (it / T) % H, s = it / (T * H);
    const float* q = qkv + (size_t)(s * T + i) * ldq + h * hs;
    float sc[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = lane + 32 * u;
      float a = -INFINITY;
      if (j <= i) {
        const float* k = qkv + (size_t)(s * T + j) * ldq + d + h * hs;
        a = 0.f;
        for (int e = 0; e < hs; e += 4) {
          const float4 qa = *reinterpret_cast<const float4*>(q + e);
          const float4 ka = *reinterpret_cast<const float4*>(k + e);
          a = fmaf(qa.x, ka.x, a); a = fmaf(qa.y, ka.y, a); a = fmaf(qa.z, ka.z, a); a = fmaf(qa.w, ka.w, a);
        }
        a *= scale;
      }
      sc[u] = a;
    }
    const float mx = warp_max(fmaxf(sc[0], sc[1]));
    float p0 = (lane <= i) ? expf(sc[0] - mx) : 0.f;
    float p1 = (lane + 32 <= i) ? expf(sc[1] - mx) : 0.f;
    const float inv = 1.0f / warp_sum(p0 + p1);
    p0 *= inv; p1 *= inv;
    for (int e0 = 0; e0 < hs; e0 += 32) {      // warp-uniform trip count: every lane shuffles
      const int e = e0 + lane;
      const bool on = e < hs;
      float y = 0.f;
      for (int j = 0; j <= i; ++j) {
        const float pj = __shfl_sync(0xffffffffu, (j < 32) ? p0 : p1, j & 31);
        if (on) y = fmaf(pj, qkv[(size_t)(s * T + j) * ldq + 2 * d + h * hs + e], y);
      }
      if (on) out[(size_t)(s * T + i) * ldo + h * hs + e] = y;
    }
  }
}

struct Smem {
  float *X, *Hb, *Big, *state, *goal, *xcur, *dC, *dU, *x2, *d1, *d2, *h3, *sig;
  int ldb;
};

struct Ctx {
  const SimtModel& m;
  int ns, t, T, R;      // sequences in this CTA, observed steps, tokens per sequence, rows
};

// One DiffusionGPT evaluation (+ GCDenoiser pre-conditioning unless inner) for the CTA's sequences.
//   xin [ns][t][act] un-scaled noisy actions, sig[ns] noise levels  ->  dout [ns][t][act]
__device__ __noinline__ void eval_model(const Ctx& c, const Smem& sm, const float* xin, const float* sig, bool uncond,
                           bool inner, float* dout) {
  const SimtModel& m = c.m;
  const int d = m.d, T = c.T, R = c.R, G = m.G, t = c.t;
  const float sd = m.sigma_data;
  // ---- embeddings + positions + interleave (score_gpts.py:284-337) -------------------------
  for (int idx = threadIdx.x; idx < R * d; idx += kThreads) {
    const int r = idx / d, col = idx - r * d;
    const int s = r / T, tok = r - s * T;
    const float sg = sig[s];
    float v;
    if (tok == 0) {
      v = fmaf(logf(sg) / 4.0f, __ldg(m.sigw + col), __ldg(m.sigb + col));
    } else if (tok <= G) {
      const int g = tok - 1;
      v = __ldg(m.tokb + col);
      if (!uncond) {
        const float* in = sm.goal + (size_t)(s * G + g) * m.obs;
        float a = 0.f;
        for (int k = 0; k < m.obs; ++k) a = fmaf(in[k], __ldg(m.tokw + (size_t)k * d + col), a);
        v += a;
      }
      v += __ldg(m.pos + (size_t)g * d + col);
    } else {
      const int j = tok - 1 - G, step = j >> 1;
      float a = 0.f;
      if ((j & 1) == 0) {
        const float* in = sm.state + (size_t)(s * t + step) * m.obs;
        for (int k = 0; k < m.obs; ++k) a = fmaf(in[k], __ldg(m.tokw + (size_t)k * d + col), a);
        v = a + __ldg(m.tokb + col);
      } else {
        const float c_in = inner ? 1.0f : 1.0f / sqrtf(sg * sg + sd * sd);
        const float* in = xin + (size_t)(s * t + step) * m.act;
        for (int k = 0; k < m.act; ++k) a = fmaf(in[k] * c_in, __ldg(m.actw + (size_t)k * d + col), a);
        v = a + __ldg(m.actb + col);
      }
      v += __ldg(m.pos + (size_t)(G + step) * d + col);
    }
    sm.X[(size_t)r * d + col] = v;
  }
  __syncthreads();
  // ---- transformer blocks (score_gpts.py:112-115) ------------------------------------------
  for (int l = 0; l < m.L; ++l) {
    const SimtLayer& w = m.layer[l];
    layer_norm_rows(sm.X, d, R, d, w.ln1w, w.ln1b, sm.Hb, d);
    __syncthreads();
    gemm_rows<EPI_STORE>(sm.Hb, d, R, w.wqkv, w.bqkv, d, 3 * d, sm.Big, sm.ldb);
    __syncthreads();
    attention_rows(sm.Big, sm.ldb, c.ns, T, d, m.H, m.hs, sm.Hb, d);
    __syncthreads();
    gemm_rows<EPI_RESID>(sm.Hb, d, R, w.wproj, w.bproj, d, d, sm.X, d, sm.Big, sm.ldb);     // qkv is dead: split-K scratch
    __syncthreads();
    layer_norm_rows(sm.X, d, R, d, w.ln2w, w.ln2b, sm.Hb, d);
    __syncthreads();
    gemm_rows<EPI_GELU>(sm.Hb, d, R, w.w1, w.b1, d, 4 * d, sm.Big, sm.ldb);
    __syncthreads();
    gemm_rows<EPI_RESID>(sm.Big, sm.ldb, R, w.w2, w.b2, 4 * d, d, sm.X, d, sm.Hb, d);      // LN2 output is dead
    __syncthreads();
  }
  // ---- ln_f, action-token gather, head, pre-conditioning (score_gpts.py:341-354) -----------
  layer_norm_rows(sm.X, d, R, d, m.lnfw, m.lnfb, sm.Hb, d);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (m.linear_out) {
    for (int it = warp; it < c.ns * t; it += kWarps) {
      const int s = it / t, step = it - s * t;
      const float* hrow = sm.Hb + (size_t)(s * T + 1 + G + 2 * step + 1) * d;
      const float sg = sig[s];
      const float den = sg * sg + sd * sd;
      for (int a = 0; a < m.act; ++a) {
        float p = 0.f;
        for (int col = lane; col < d; col += 32) p = fmaf(hrow[col], __ldg(m.hw0 + (size_t)col * m.act_pad + a), p);
        p = warp_sum(p);
        if (lane == 0) {
          float f = p + __ldg(m.hb0 + a);
          if (!inner) {
            const float c_skip = sd * sd / den, c_out = sg * sd / sqrtf(den);
            f = __fadd_rn(__fmul_rn(f, c_out), __fmul_rn(xin[(size_t)it * m.act + a], c_skip));
          }
          dout[(size_t)it * m.act + a] = f;
        }
      }
    }
  } else {
    // action_pred = Linear(d,100) -> SiLU -> Linear(100,act)   (score_gpts.py:186-190)
    float* hid = sm.Big;   // [ns*t][hid_pad]
    for (int idx = threadIdx.x; idx < c.ns * t * m.hid; idx += kThreads) {
      const int it = idx / m.hid, u = idx - it * m.hid;
      const int s = it / t, step = it - s * t;
      const float* hrow = sm.Hb + (size_t)(s * T + 1 + G + 2 * step + 1) * d;
      float p = __ldg(m.hb0 + u);
      for (int col = 0; col < d; ++col) p = fmaf(hrow[col], __ldg(m.hw0 + (size_t)col * m.hid_pad + u), p);
      hid[(size_t)it * m.hid_pad + u] = p / (1.0f + expf(-p));
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < c.ns * t * m.act; idx += kThreads) {
      const int it = idx / m.act, a = idx - it * m.act;
      const int s = it / t;
      float f = __ldg(m.hb1 + a);
      for (int u = 0; u < m.hid; ++u) f = fmaf(hid[(size_t)it * m.hid_pad + u], __ldg(m.hw1 + (size_t)u * m.act_pad + a), f);
      if (!inner) {
        const float sg = sig[s], den = sg * sg + sd * sd;
        const float c_skip = sd * sd / den, c_out = sg * sd / sqrtf(den);
        f = __fadd_rn(__fmul_rn(f, c_out), __fmul_rn(xin[idx], c_skip));
      }
      dout[idx] = f;
    }
  }
  __syncthreads();
}

// model(...) as the samplers see it: plain, uncond, or the classifier-free mix of both.
__device__ void eval_wrapped(const Ctx& c, const Smem& sm, const float* xin, const float* sig,
                             uint32_t flags, float lambda, float* dout) {
  const bool inner = flags & BESO_FLAG_INNER;
  if (flags & BESO_FLAG_CFG) {
    eval_model(c, sm, xin, sig, false, inner, sm.dC);
    eval_model(c, sm, xin, sig, true, inner, sm.dU);
    const int n = c.ns * c.t * c.m.act;
    for (int i = threadIdx.x; i < n; i += kThreads)   // out_uncond + lambda * (out - out_uncond)
      dout[i] = __fadd_rn(sm.dU[i], __fmul_rn(lambda, __fsub_rn(sm.dC[i], sm.dU[i])));
    __syncthreads();
  } else {
    eval_model(c, sm, xin, sig, (flags & BESO_FLAG_UNCOND) != 0, inner, dout);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
simt_denoise_kernel(const __grid_constant__ SimtModel m, const __grid_constant__ SampleArgs sa,
                    const float* __restrict__ state, const float* __restrict__ goal,
                    const float* __restrict__ action, const float* __restrict__ sigma,
                    float* __restrict__ out, int B, int t, int S, uint32_t flags, float lambda, int ldb) {
  extern __shared__ __align__(16) float smem_f[];
  const int T = 1 + m.G + 2 * t;
  const int seq0 = blockIdx.x * S;
  const int ns = min(S, B - seq0);
  if (ns <= 0) return;
  Smem sm;
  float* p = smem_f;
  auto take = [&](size_t n) { float* q = p; p += (n + 3) & ~size_t(3); return q; };
  sm.ldb = ldb;
  sm.X = take((size_t)S * T * m.d);
  sm.Hb = take((size_t)S * T * m.d);
  sm.Big = take((size_t)S * T * ldb);
  sm.state = take((size_t)S * t * m.obs);
  sm.goal = take((size_t)S * max(m.G, 1) * m.obs);
  const int nx = S * t * m.act;
  sm.xcur = take(nx); sm.dC = take(nx); sm.dU = take(nx); sm.x2 = take(nx); sm.d1 = take(nx); sm.d2 = take(nx); sm.h3 = take(nx);
  sm.sig = take(S);
  Ctx c{m, ns, t, T, ns * T};

  for (int i = threadIdx.x; i < ns * t * m.obs; i += kThreads) {
    const float v = state[(size_t)seq0 * t * m.obs + i];
    sm.state[i] = sa.in_tab ? io_scale(v, sa.in_tab, m.obs, i % m.obs) : v;
  }
  for (int i = threadIdx.x; i < ns * m.G * m.obs; i += kThreads) {
    float v = goal[(size_t)seq0 * m.G * m.obs + i];
    if (sa.in_tab) v = io_scale(v, sa.in_tab, m.obs, i % m.obs);
    if (sa.goal_keep) v = __fmul_rn(v, __ldg(sa.goal_keep + i % m.obs));
    sm.goal[i] = v;
  }
  const int n = ns * t * m.act;
  for (int i = threadIdx.x; i < n; i += kThreads) sm.xcur[i] = action[(size_t)seq0 * t * m.act + i];
  float* gout = out + (size_t)seq0 * t * m.act;

  if (sa.n_steps == 0) {                      // GCDenoiser.forward, per-sequence sigma
    for (int i = threadIdx.x; i < ns; i += kThreads) sm.sig[i] = sigma[seq0 + i];
    __syncthreads();
    eval_wrapped(c, sm, sm.xcur, sm.sig, flags, lambda, sm.d1);
    for (int i = threadIdx.x; i < n; i += kThreads) gout[i] = sm.d1[i];
    return;
  }
  for (int step = 0; step < sa.n_steps; ++step) {
    const float s_hat = sa.sig[step];         // gamma = 0  ->  sigma_hat = sigma_i * 1
    const float s_next = sa.sig[step + 1];
    __syncthreads();
    for (int i = threadIdx.x; i < ns; i += kThreads) sm.sig[i] = s_hat;
    __syncthreads();
    eval_wrapped(c, sm, sm.xcur, sm.sig, flags, lambda, sm.d1);       // d1 <- denoised
    if (sa.sampler == BESO_SAMPLER_DDIM) {    // gc_sampling.py:921-923
      const float ca = sa.ca[step], ce = sa.ce[step];
      for (int i = threadIdx.x; i < n; i += kThreads)
        sm.xcur[i] = __fsub_rn(__fmul_rn(ca, sm.xcur[i]), __fmul_rn(ce, sm.d1[i]));
    } else if (sa.sampler == BESO_SAMPLER_LMS) {        // gc_sampling.py:454-465; history d_{i-1..i-3} in x2, d2, h3
      const float c0 = sa.ca[step], c1 = sa.ce[step], c2 = sa.c1[step], c3 = sa.c2[step];
      for (int i = threadIdx.x; i < n; i += kThreads) {
        const float d = __fdiv_rn(__fsub_rn(sm.xcur[i], sm.d1[i]), s_hat);
        float acc = __fmul_rn(c0, d);
        if (c1 != 0.0f) acc = __fadd_rn(acc, __fmul_rn(c1, sm.x2[i]));
        if (c2 != 0.0f) acc = __fadd_rn(acc, __fmul_rn(c2, sm.d2[i]));
        if (c3 != 0.0f) acc = __fadd_rn(acc, __fmul_rn(c3, sm.h3[i]));
        sm.xcur[i] = __fadd_rn(sm.xcur[i], acc);
        sm.h3[i] = sm.d2[i]; sm.d2[i] = sm.x2[i]; sm.x2[i] = d;
      }
    } else if (sa.sampler == BESO_SAMPLER_TWO_STAGE) {  // coefficient program (include/beso_b200.h)
      const float a1 = sa.ca[step], b1 = sa.ce[step], sb = sa.sigb[step], su = sa.su[step];
      const float* nz = su != 0.0f ? sa.noise + (size_t)step * sa.noise_stride + (size_t)seq0 * t * m.act : nullptr;
      if (sb == 0.0f) {
        for (int i = threadIdx.x; i < n; i += kThreads) {
          float xe = fmaf(a1, sm.xcur[i], b1 * sm.d1[i]);
          if (nz) xe = fmaf(su, __ldg(nz + i), xe);
          sm.xcur[i] = xe;
        }
      } else {
        for (int i = threadIdx.x; i < n; i += kThreads) sm.x2[i] = fmaf(a1, sm.xcur[i], b1 * sm.d1[i]);
        __syncthreads();
        for (int i = threadIdx.x; i < ns; i += kThreads) sm.sig[i] = sb;
        __syncthreads();
        eval_wrapped(c, sm, sm.x2, sm.sig, flags, lambda, sm.d2);
        const float a2 = sa.c1[step], b2 = sa.c2[step], c2 = sa.c3[step];
        for (int i = threadIdx.x; i < n; i += kThreads) {
          float xe = fmaf(a2, sm.xcur[i], fmaf(b2, sm.x2[i], c2 * sm.d2[i]));
          if (nz) xe = fmaf(su, __ldg(nz + i), xe);
          sm.xcur[i] = xe;
        }
      }
    } else if (sa.sampler == BESO_SAMPLER_DPMPP_2M) {   // gc_sampling.py:726-735; d2 keeps old_denoised
      const float ca = sa.ca[step], ce = sa.ce[step], c1 = sa.c1[step], c2 = sa.c2[step];
      for (int i = threadIdx.x; i < n; i += kThreads) {
        const float den = sm.d1[i];
        const float dd = (c2 != 0.0f) ? __fsub_rn(__fmul_rn(c1, den), __fmul_rn(c2, sm.d2[i])) : den;
        sm.xcur[i] = __fsub_rn(__fmul_rn(ca, sm.xcur[i]), __fmul_rn(ce, dd));
        sm.d2[i] = den;
      }
    } else {
      // Euler ancestral (gc_sampling.py:216-256) steps down to sigma_down and adds sigma_up of fresh noise
      const bool anc = sa.sampler == BESO_SAMPLER_EULER_ANCESTRAL;
      const float dt = __fsub_rn(anc ? sa.ca[step] : s_next, s_hat);
      const bool heun2 = (sa.sampler == BESO_SAMPLER_HEUN) && (s_next != 0.0f);
      const bool add_noise = anc && sa.ca[step] > 0.0f;
      const float* nz = add_noise ? sa.noise + (size_t)step * sa.noise_stride + (size_t)seq0 * t * m.act : nullptr;
      for (int i = threadIdx.x; i < n; i += kThreads) {
        const float dd = __fdiv_rn(__fsub_rn(sm.xcur[i], sm.d1[i]), s_hat);   // to_d
        sm.d1[i] = dd;
        float xe = __fadd_rn(sm.xcur[i], __fmul_rn(dd, dt));
        if (add_noise) xe = __fadd_rn(xe, __fmul_rn(__ldg(nz + i), sa.ce[step]));
        if (heun2) sm.x2[i] = xe; else sm.xcur[i] = xe;
      }
      if (heun2) {                            // gc_sampling.py:304-310
        __syncthreads();
        for (int i = threadIdx.x; i < ns; i += kThreads) sm.sig[i] = s_next;
        __syncthreads();
        eval_wrapped(c, sm, sm.x2, sm.sig, flags, lambda, sm.d2);
        for (int i = threadIdx.x; i < n; i += kThreads) {
          const float d2 = __fdiv_rn(__fsub_rn(sm.x2[i], sm.d2[i]), s_next);
          const float dp = __fdiv_rn(__fadd_rn(sm.d1[i], d2), 2.0f);
          sm.xcur[i] = __fadd_rn(sm.xcur[i], __fmul_rn(dp, dt));
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += kThreads) {
    float v = sm.xcur[i];
    if (sa.clip) v = io_clip(v, sa.clip, m.act, i % m.act);
    gout[i] = v;
    if (sa.unscaled) sa.unscaled[(size_t)seq0 * t * m.act + i] = sa.out_tab ? io_scale(v, sa.out_tab, m.act, i % m.act) : v;
  }
}

inline int round_ldb(int n) {   // smallest ld >= n with ld % 8 == 4: conflict-free float4 rows
  int ld = (n + 3) & ~3;
  while ((ld & 7) != 4) ld += 4;
  return ld;
}

size_t simt_smem_bytes(const SimtModel& m, int t, int S, int* ldb_out) {
  const int T = 1 + m.G + 2 * t;
  const int ldb = round_ldb(4 * m.d);
  auto r4 = [](size_t n) { return (n + 3) & ~size_t(3); };
  size_t f = 2 * r4((size_t)S * T * m.d) + r4((size_t)S * T * ldb) + r4((size_t)S * t * m.obs) +
             r4((size_t)S * (m.G > 0 ? m.G : 1) * m.obs) + 7 * r4((size_t)S * t * m.act) + r4(S);
  if (ldb_out) *ldb_out = ldb;
  return f * sizeof(float);
}

}  // namespace

int simt_plan_launch(const SimtModel& m, int t, int max_smem, SimtLaunch* out) {
  const int T = 1 + m.G + 2 * t;
  if (T > 64) { set_error("precise mode supports at most 64 tokens per sequence"); return BESO_E_UNSUPPORTED; }
  int S = 0;
  for (int s = 1; s <= 8; ++s) {
    if (simt_smem_bytes(m, t, s, nullptr) <= (size_t)max_smem && s * T <= 64) S = s; else break;
  }
  if (S == 0) { set_error("model too large for the shared-memory resident precise kernel"); return BESO_E_UNSUPPORTED; }
  out->S = S;
  out->t = t;
  out->smem_bytes = simt_smem_bytes(m, t, S, nullptr);
  return BESO_OK;
}

int simt_launch(const SimtModel& m, const SimtLaunch& L, const SampleArgs& sa, const float* state,
                const float* goal, const float* action_or_x, const float* sigma, float* out,
                cudaStream_t stream) {
  int ldb = 0;
  const size_t smem = simt_smem_bytes(m, L.t, L.S, &ldb);
  static size_t configured = 0;
  if (smem > configured) {
    BESO_CUDA(cudaFuncSetAttribute(simt_denoise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int grid = (L.B + L.S - 1) / L.S;
  simt_denoise_kernel<<<grid, kThreads, smem, stream>>>(m, sa, state, goal, action_or_x, sigma, out, L.B, L.t,
                                                        L.S, L.flags, L.cond_lambda, ldb);
  ++g_kernel_launches;
  BESO_CUDA(cudaGetLastError());
  return BESO_OK;
}

}  // namespace beso
