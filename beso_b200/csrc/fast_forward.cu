// FAST mode: GCDenoiser -> DiffusionGPT forward and the DDIM / Euler / Heun sample loop as ONE
// persistent, warp-specialised sm_100a kernel.  bf16 operands on tcgen05 tensor cores, fp32
// accumulation in TMEM, fp32 LayerNorm / softmax / GELU / residual / pre-conditioning.
//
// Reference semantics (beso/agents/diffusion_agents/k_diffusion/): score_wrappers.py:31-43,81-96;
// score_gpts.py:50-80,96-115,272-358; gc_sampling.py:167-213,259-314,895-924;
// classifier_free_sampler.py:35-49.
//
// Design (DESIGN.md has the long form)
//  * A CTA owns a 128-row tile = floor(128 / T) whole sequences (T tokens each) for the whole
//    launch: all layers and all sampler steps.  Row r of the tile is TMEM lane r.
//  * The fp32 residual stream X (128 x 256) lives in TMEM columns [0,256) and never leaves it:
//    the attention-projection and MLP-down GEMMs accumulate straight into X (residual add for
//    free); their biases are added by the next LayerNorm pass (tcgen05.ld -> add -> tcgen05.st).
//  * TMEM columns [256,512) are two 128-column scratch accumulators (per-head QKV, FC1 chunks).
//  * Every GEMM B operand comes from one linear "weight tape" in HBM/L2, pre-swizzled into the
//    UMMA K-major SWIZZLE_128B image and ordered exactly as the MMA warp consumes it; a producer
//    thread streams it with cp.async.bulk (TMA engine) through a 4 x 16 KB mbarrier ring.
//  * One elected thread issues every tcgen05.mma, driven by a host-built "fill program" (one
//    entry per ring fill: A operand, TMEM column, N, barriers to wait on / commit to).
//  * 8 compute warps (2 per TMEM lane quadrant) do the LayerNorms, the QKV drain, causal
//    attention with mma.sync on bf16 Q/K/V staged in shared memory, erf-GELU, the action head
//    read-out, the Karras pre-conditioning and the sampler update.
//  * Embeddings (state / goal / action / sigma / position / biases) are one K=128 GEMM: the A
//    operand carries the raw inputs plus one-hot token-position columns whose B rows hold
//    (bias + pos_emb) split into bf16 hi + lo, so X starts exact to ~2^-17.
#include <cuda_bf16.h>
#include <string.h>

#include <vector>

#include "fast.cuh"
#include "umma.cuh"

namespace beso {

using namespace umma;

namespace {

// ---- fixed geometry (d = 256, 4 heads of 64) ---------------------------------------------------
constexpr int kD = 256, kH = 4, kHS = 64, kFF = 1024;
constexpr int kRows = 128;
constexpr int kThreads = 384;             // warp 0 producer, warp 1 MMA + TMEM alloc, 2-3 idle, 4-11 compute
constexpr int kComputeWarp0 = 4, kComputeThreads = 256;
constexpr int kMaxTokens = 24;            // one-hot columns available in the misc atom
constexpr int kOneHot0 = 16;              // first one-hot column of the misc atom
constexpr int kMaxAct = 13, kMaxObs = 64;

// shared-memory map (bytes from the 1024-aligned base)
constexpr uint32_t kSmA = 0;                        // 4 atoms [128 x 64] bf16: LN output / embedding input
constexpr uint32_t kSmRing = 65536;                 // 4 slots x 16 KB weight ring
constexpr uint32_t kSlotBytes = 16384, kSlots = 4;
constexpr uint32_t kSmU = 131072;                   // union region
constexpr uint32_t kQkvStride = 400;                // bytes per row of the bf16 Q|K|V staging (192 + 8 pad)
constexpr uint32_t kSmQkv = kSmU;                   // [128][400 B]
constexpr uint32_t kSmY = kSmU + 51200;             // attention output atom [128 x 64] bf16
constexpr uint32_t kSmH0 = kSmU, kSmH1 = kSmU + 32768;   // FC1 output chunks, 2 atoms each
constexpr uint32_t kSmVecA = kSmU + 67584;          // 198656: [pend | ln_w | ln_b | bqkv(768)] fp32
constexpr uint32_t kVecAFloats = 1536;
constexpr uint32_t kSmVecM = kSmVecA + kVecAFloats * 4;   // [bproj | ln2_w | ln2_b | b1(1024)]
constexpr uint32_t kVecMFloats = 1792;
constexpr uint32_t kSmStats = kSmVecM + kVecMFloats * 4;  // [2][128] float2
constexpr uint32_t kSmProg = kSmStats + 2048;
constexpr int kProgEntries = 4 + 104 + 4;        // embedding | one layer (identical for all) | head
constexpr uint32_t kXFloats = 832;
constexpr uint32_t kSmBars = 230400;                // 32 mbarriers + tmem pointer
constexpr uint32_t kSmemBytes = kSmBars + 512;

// barrier ids used by the fill program
enum { B_A_READY = 0, B_X_DONE, B_ACC_FULL0, B_ACC_FULL1, B_ACC_EMPTY0, B_ACC_EMPTY1, B_OP_READY0, B_OP_READY1,
       B_OP_EMPTY0, B_OP_EMPTY1, B_Y_READY, B_Y_EMPTY, B_FULL0, B_EMPTY0 = B_FULL0 + 4, B_COUNT = B_EMPTY0 + 4 };
constexpr int kAttnWarps = 10;            // 8 compute warps + warps 2-3 help with attention
constexpr uint8_t kNone = 0xF;

// TMEM columns
constexpr uint32_t kColX = 0, kColS0 = 256, kColS1 = 384;

struct __align__(8) Fill {     // one ring fill = rows x 64 bf16 of B operand + the MMAs that consume it
  uint16_t a_off16;            // A atom, shared-memory offset / 16
  uint16_t d_col;              // TMEM column of D
  uint8_t n8;                  // N / 8 (rows of the fill)
  uint8_t acc;                 // bit 0: accumulate into D on the first k-step; bit 1: this fill and the next one
                               // sit in adjacent ring slots and are consumed by ONE wider MMA per k-step
  uint8_t waits;               // two 4-bit barrier ids to wait on before issuing (0xF = none)
  uint8_t commits;             // two 4-bit barrier ids to commit to afterwards
};

struct FastParams {
  const uint8_t* tape;         // per-eval weight tape
  Fill prog[kProgEntries];     // fill program: embedding | one layer | head (kernel-parameter space)
  const float* vec;            // per layer: vecA (1536) | vecM (1792); then final vecA (1536)
  int n_fills, L, G, obs, act, T, t, S, n_tiles, B, evals;
  uint32_t flags;
  float lambda, sigma_data;
  const float *state, *goal, *xin, *sigma;
  float* out;
  float* trace;                // optional debug dump of X after every LayerNorm pass (tile 0, eval 0)
  long long* timeline;         // optional clock64 stamps of block 0, second evaluation (see tools/timeline_fast.py)
};

// ================================ device code =====================================================
__device__ __forceinline__ int prog_index(int f, int n_fills) {
  if (f < 4) return f;
  if (f >= n_fills - 4) return 108 + (f - (n_fills - 4));
  return 4 + (f - 4) % 104;
}
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void attn_sync() { asm volatile("bar.sync 2, 320;" ::: "memory"); }   // compute + helper warps
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}

// Bounded wait: a protocol bug must not hang the GPU.  ~4 s at 2 GHz, then trap with the barrier id.
__device__ __noinline__ void wait_timeout(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000ll) {
      printf("beso fast kernel: mbarrier id %u parity %u timed out (block %d thread %d)\n",
             (bar & 0x3FF) / 8, parity, blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void spin_wait(uint32_t bar, uint32_t parity) {
  for (int i = 0; i < 64; ++i) if (mbar_try_wait(bar, parity)) return;
  wait_timeout(bar, parity);
}

// erf-GELU without erff(): with z = |x| / sqrt(2), 0.5 * erfc(z) = 2^q(z) to 2.1e-6 absolute for a
// degree-5 q fitted on [0, 6] (q -> -inf beyond), so
//   gelu(x) = x * Phi(x) = max(x, 0) - |x| * 2^q(z)          |error| <= 5.9e-7 absolute
// 9 instructions (5 FFMA Horner, 1 MUFU.EX2) instead of ~46 for erff().
__device__ __forceinline__ float gelu_fast(float x) {
  const float ax = fabsf(x);
  const float z = ax * 0.70710678118654752440f;
  float q = fmaf(-2.784754615e-03f, z, 2.889863029e-02f);
  q = fmaf(q, z, -1.476386487e-01f);
  q = fmaf(q, z, -9.191277623e-01f);
  q = fmaf(q, z, -1.627753854e+00f);
  q = fmaf(q, z, -1.000006080e+00f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q));
  return fmaf(-ax, e, fmaxf(x, 0.f));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct Compute {
  uint8_t* sm;
  uint32_t sbase, tmem;
  int wq, lane, hf, row, ctid;       // TMEM lane quadrant, lane, column half, tile row, compute thread id
  uint32_t phases;                   // parity bit per barrier id this role waits on
  long long* tl;                     // timeline cursor (nullptr = off)
  __device__ void stamp() { if (tl != nullptr) *tl++ = clock64(); }
  __device__ uint32_t bar(int id) const { return sbase + kSmBars + id * 8; }
  __device__ void wait(int id) { spin_wait(bar(id), (phases >> id) & 1u); phases ^= 1u << id; }
  __device__ void arrive(int id) const { __syncwarp(); if (lane == 0) mbar_arrive(bar(id)); }
  __device__ uint32_t lane_addr(uint32_t col) const { return tmem + ((uint32_t)(wq * 32) << 16) + col; }
};

// A <- bf16(LayerNorm(X + pend) * w + b).   vec = [pend | w | b] in shared memory.
// X (TMEM) is only read: every projection / MLP bias is added to X up front by the embedding GEMM and
// `pend` holds minus the biases that are not due yet at this point of the network (see fast_pack).
__device__ void ln_pass(Compute& c, const float* vec, const FastParams& p, int trace_slot) {
  float va[32], vb[32];
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
  const int col0 = c.hf * 128;
  const bool tracing = p.trace != nullptr && blockIdx.x == 0 && trace_slot >= 0;
  auto pass1 = [&](float (&v)[32], int col) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 pd = *reinterpret_cast<const float4*>(vec + col + i);
      const float a0 = v[i] + pd.x, a1 = v[i + 1] + pd.y, a2 = v[i + 2] + pd.z, a3 = v[i + 3] + pd.w;
      s0 += a0; s1 += a1; s2 += a2; s3 += a3;
      q0 = fmaf(a0, a0, q0); q1 = fmaf(a1, a1, q1); q2 = fmaf(a2, a2, q2); q3 = fmaf(a3, a3, q3);
      if (tracing) {
        float* tr = p.trace + ((size_t)trace_slot * kRows + c.row) * kD + col + i;
        tr[0] = a0; tr[1] = a1; tr[2] = a2; tr[3] = a3;
      }
    }
  };
  // the next chunk's TMEM load is in flight while the current one is reduced
  tmem_ld32(c.lane_addr(kColX + col0), va);
  tmem_wait_ld();
  tmem_ld32(c.lane_addr(kColX + col0 + 32), vb);
  pass1(va, col0);
  tmem_wait_ld();
  tmem_ld32(c.lane_addr(kColX + col0 + 64), va);
  pass1(vb, col0 + 32);
  tmem_wait_ld();
  tmem_ld32(c.lane_addr(kColX + col0 + 96), vb);
  pass1(va, col0 + 64);
  tmem_wait_ld();
  tmem_ld32(c.lane_addr(kColX + col0), va);          // first chunk of the second pass
  pass1(vb, col0 + 96);
  const float sum = (s0 + s1) + (s2 + s3), sq = (q0 + q1) + (q2 + q3);
  float2* stats = reinterpret_cast<float2*>(c.sm + kSmStats);
  stats[c.hf * kRows + c.row] = make_float2(sum, sq);
  compute_sync();
  const float2 o = stats[(c.hf ^ 1) * kRows + c.row];
  const float mean = (sum + o.x) * (1.0f / kD);
  const float var = fmaxf((sq + o.y) * (1.0f / kD) - mean * mean, 0.f);
  const float rstd = rsqrtf(var + 1e-5f);
  const float nmr = -mean * rstd;
  const float* w = vec + kD;
  const float* b = vec + 2 * kD;
  auto pass2 = [&](float (&v)[32], int col) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 pd = *reinterpret_cast<const float4*>(vec + col + i);
      const float4 ww = *reinterpret_cast<const float4*>(w + col + i);
      const float4 bb = *reinterpret_cast<const float4*>(b + col + i);
      v[i] = fmaf(fmaf(v[i] + pd.x, rstd, nmr), ww.x, bb.x);
      v[i + 1] = fmaf(fmaf(v[i + 1] + pd.y, rstd, nmr), ww.y, bb.y);
      v[i + 2] = fmaf(fmaf(v[i + 2] + pd.z, rstd, nmr), ww.z, bb.z);
      v[i + 3] = fmaf(fmaf(v[i + 3] + pd.w, rstd, nmr), ww.w, bb.w);
    }
    uint8_t* atom = c.sm + kSmA + (col >> 6) * 16384;
    const int chunk0 = (col & 63) >> 3;
#pragma unroll
    for (int q = 0; q < 4; ++q) st_chunk(atom, c.row, chunk0 + q, v + q * 8);
  };
  tmem_wait_ld();
  tmem_ld32(c.lane_addr(kColX + col0 + 32), vb);
  pass2(va, col0);
  tmem_wait_ld();
  tmem_ld32(c.lane_addr(kColX + col0 + 64), va);
  pass2(vb, col0 + 32);
  tmem_wait_ld();
  tmem_ld32(c.lane_addr(kColX + col0 + 96), vb);
  pass2(va, col0 + 64);
  tmem_wait_ld();
  pass2(vb, col0 + 96);
  fence_async_smem();
  tc_fence_before();
  c.arrive(B_A_READY);
  compute_sync();          // stats buffer may be rewritten by the next pass only after everyone read it
}

// Accumulator of head h (Q|K at S0, V at S1[0:64)) -> + bias -> bf16 Q|K|V staging rows.
__device__ void drain_qkv(Compute& c, const float* bqkv_h) {
  float v[32];
#pragma unroll 1
  for (int ch = 0; ch < 3; ++ch) {
    const int col = c.hf * 96 + ch * 32;                 // 0..191 within [Q_h | K_h | V_h]
    tmem_ld32(c.lane_addr(kColS0 + col), v);
    tmem_wait_ld();
    uint8_t* dst = c.sm + kSmQkv + (uint32_t)c.row * kQkvStride + col * 2;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 u;
      const float4 b0 = *reinterpret_cast<const float4*>(bqkv_h + col + q * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(bqkv_h + col + q * 8 + 4);
      u.x = pack_bf16x2(v[q * 8 + 0] + b0.x, v[q * 8 + 1] + b0.y);
      u.y = pack_bf16x2(v[q * 8 + 2] + b0.z, v[q * 8 + 3] + b0.w);
      u.z = pack_bf16x2(v[q * 8 + 4] + b1.x, v[q * 8 + 5] + b1.y);
      u.w = pack_bf16x2(v[q * 8 + 6] + b1.z, v[q * 8 + 7] + b1.w);
      *reinterpret_cast<uint4*>(dst + q * 16) = u;
    }
  }
  tc_fence_before();
}

// Causal softmax(Q K^T) V for every sequence of the tile, one warp per sequence, mma.sync bf16.
// Q is pre-scaled by 1/sqrt(hs) (folded into the packed weights).  Output -> Y atom (SW128 A layout).
__device__ void attention_head(uint8_t* sm, uint32_t sbase, int awarp, int lane, int S, int T) {
  const uint32_t qkv = sbase + kSmQkv;
  const int MT = (T + 15) >> 4;                  // 16-row query tiles == 16-key steps
  for (int item = awarp; item < S * MT; item += kAttnWarps) {
    const int mt = MT - 1 - item / S, s = item % S;   // later query tiles see more keys: schedule them first
    const int row0 = s * T;
    {
      // ---- S = Q K^T over keys [0, 16*(mt+1)) (causal: later key tiles are fully masked) ----
      const int nkt = mt + 1;                    // 16-key steps needed
      float sc[2][2][4];                         // [key16][n8][frag]
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) sc[a][b][0] = sc[a][b][1] = sc[a][b][2] = sc[a][b][3] = 0.f;
      uint32_t qa[4][4];
      {
        const int r = min(row0 + mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, kRows - 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) ldmatrix_x4(qkv + r * kQkvStride + (k * 16 + (lane >> 4) * 8) * 2, qa[k]);
      }
#pragma unroll
      for (int kt = 0; kt < 2; ++kt) {
        if (kt < nkt) {
          const int r = min(row0 + kt * 16 + (lane & 7) + (lane >> 4) * 8, kRows - 1);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint32_t kb[4];
            ldmatrix_x4(qkv + r * kQkvStride + (64 + k * 16 + ((lane >> 3) & 1) * 8) * 2, kb);
            mma_16816(sc[kt][0], qa[k], kb[0], kb[1]);
            mma_16816(sc[kt][1], qa[k], kb[2], kb[3]);
          }
        }
      }
      // ---- mask + softmax (rows i0 = lane/4 and i0 + 8 of this query tile) ----
      const int i_lo = mt * 16 + (lane >> 2), i_hi = i_lo + 8;
      float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
      for (int kt = 0; kt < 2; ++kt)
#pragma unroll
        for (int nb = 0; nb < 2; ++nb)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int j = kt * 16 + nb * 8 + (lane & 3) * 2 + e;
            const bool on = kt < nkt;
            if (!(on && j <= i_lo)) sc[kt][nb][e] = -INFINITY;
            if (!(on && j <= i_hi)) sc[kt][nb][2 + e] = -INFINITY;
            mx_lo = fmaxf(mx_lo, sc[kt][nb][e]);
            mx_hi = fmaxf(mx_hi, sc[kt][nb][2 + e]);
          }
      mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1)); mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
      mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1)); mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
      float sum_lo = 0.f, sum_hi = 0.f;
      uint32_t pa[2][4];                         // P as A fragments, one per 16-key step
#pragma unroll
      for (int kt = 0; kt < 2; ++kt) {
        float pv[2][4];
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            pv[nb][e] = __expf(sc[kt][nb][e] - mx_lo);
            pv[nb][2 + e] = __expf(sc[kt][nb][2 + e] - mx_hi);
            sum_lo += pv[nb][e]; sum_hi += pv[nb][2 + e];
          }
        }
        pa[kt][0] = pack_bf16x2(pv[0][0], pv[0][1]);   // (row lo, keys 0-7)
        pa[kt][1] = pack_bf16x2(pv[0][2], pv[0][3]);   // (row hi, keys 0-7)
        pa[kt][2] = pack_bf16x2(pv[1][0], pv[1][1]);   // (row lo, keys 8-15)
        pa[kt][3] = pack_bf16x2(pv[1][2], pv[1][3]);   // (row hi, keys 8-15)
      }
      sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 1); sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 2);
      sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 1); sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 2);
      // ---- O = P V ----
      float o[8][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
      for (int kt = 0; kt < 2; ++kt) {
        if (kt < nkt) {
          const int r = min(row0 + kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, kRows - 1);
#pragma unroll
          for (int np = 0; np < 4; ++np) {       // pairs of 8-wide output column tiles
            uint32_t vb[4];
            ldmatrix_x4_trans(qkv + r * kQkvStride + (128 + np * 16 + (lane >> 4) * 8) * 2, vb);
            mma_16816(o[np * 2], pa[kt], vb[0], vb[1]);
            mma_16816(o[np * 2 + 1], pa[kt], vb[2], vb[3]);
          }
        }
      }
      const float inv_lo = 1.0f / sum_lo, inv_hi = 1.0f / sum_hi;
      uint8_t* y = sm + kSmY;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const int e = n * 8 + (lane & 3) * 2;    // column within the head
        if (i_lo < T) {
          const uint32_t r = row0 + i_lo;
          *reinterpret_cast<uint32_t*>(y + sw128_offset(r, e >> 3) + (e & 7) * 2) = pack_bf16x2(o[n][0] * inv_lo, o[n][1] * inv_lo);
        }
        if (i_hi < T) {
          const uint32_t r = row0 + i_hi;
          *reinterpret_cast<uint32_t*>(y + sw128_offset(r, e >> 3) + (e & 7) * 2) = pack_bf16x2(o[n][2] * inv_hi, o[n][3] * inv_hi);
        }
      }
    }
  }
}

// FC1 chunk accumulator (buffer b) -> + b1 -> erf-GELU -> bf16 -> H[b] (two K atoms).
__device__ void drain_gelu(Compute& c, int b, const float* b1c) {
  float v[32];
  uint8_t* atom = c.sm + (b ? kSmH1 : kSmH0) + c.hf * 16384;
#pragma unroll 1
  for (int ch = 0; ch < 2; ++ch) {
    const int col = c.hf * 64 + ch * 32;
    tmem_ld32(c.lane_addr((b ? kColS1 : kColS0) + col), v);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 bb = *reinterpret_cast<const float4*>(b1c + col + i);
      v[i] = gelu_fast(v[i] + bb.x); v[i + 1] = gelu_fast(v[i + 1] + bb.y);
      v[i + 2] = gelu_fast(v[i + 2] + bb.z); v[i + 3] = gelu_fast(v[i + 3] + bb.w);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) st_chunk(atom, c.row, ch * 4 + q, v + q * 8);
  }
  fence_async_smem();
  tc_fence_before();
}

// Per-thread description of the embedding-input task it owns for the whole tile: one (row, atom).
struct EmbedTask {
  const float* src;      // obs atom: state / goal vector of this row (nullptr = zeros)
  int row, atom, vs, tok, xoff;   // xoff >= 0: action row, offset of its act values in the tile's x buffer
  bool valid;
};
__device__ EmbedTask make_embed_task(const Compute& c, const FastParams& p, int tile) {
  const bool cfg = (p.flags & BESO_FLAG_CFG) != 0;
  EmbedTask e;
  e.row = c.ctid & (kRows - 1);
  e.atom = c.ctid >> 7;
  e.vs = e.row / p.T;
  e.tok = e.row - e.vs * p.T;
  const int ls = cfg ? (e.vs >> 1) : e.vs;
  const int seq = tile * (cfg ? p.S / 2 : p.S) + ls;
  e.valid = e.vs < p.S && seq < p.B;
  e.src = nullptr;
  e.xoff = -1;
  if (e.valid) {
    const bool uncond = cfg ? ((e.vs & 1) != 0) : ((p.flags & BESO_FLAG_UNCOND) != 0);
    const int j = e.tok - 1 - p.G;
    if (e.tok >= 1 && e.tok <= p.G) { if (!uncond) e.src = p.goal + ((size_t)seq * p.G + (e.tok - 1)) * p.obs; }
    else if (j >= 0 && (j & 1) == 0) e.src = p.state + ((size_t)seq * p.t + (j >> 1)) * p.obs;
    else if (j >= 0) e.xoff = (ls * p.t + (j >> 1)) * p.act;
  }
  return e;
}

// A <- embedding-GEMM input rows: [obs atom | misc atom] (see file header), atoms 2..3 untouched.
// One (row, atom) per thread; all global loads of a row are issued before any is used.
__device__ void build_embed_input(const Compute& c, const FastParams& p, const EmbedTask& e, const float* xsrc,
                                  const float* sigv) {
  uint8_t* atom = c.sm + kSmA + e.atom * 16384;
  if (e.atom == 0) {
    float v[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = 0.f;
    if (e.src != nullptr) {
      if ((p.obs & 3) == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (i * 4 < p.obs) {
            const float4 f = __ldg(reinterpret_cast<const float4*>(e.src) + i);
            v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) if (i < p.obs) v[i] = __ldg(e.src + i);
      }
    }
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) st_chunk(atom, e.row, ch, v + ch * 8);
  } else {
    const bool inner = (p.flags & BESO_FLAG_INNER) != 0;
    const float sg = e.valid ? sigv[e.vs] : 1.0f;
    const float c_in = inner ? 1.0f : 1.0f / sqrtf(sg * sg + p.sigma_data * p.sigma_data);
    const float cn = logf(sg) * 0.25f;
    const float cn_hi = __bfloat162float(__float2bfloat16_rn(cn));
    const int hot = kOneHot0 + 2 * e.tok;
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = ch * 8 + i;
        float x = 0.f;
        if (e.valid) {
          if (k < kOneHot0) {
            if (e.xoff >= 0 && k < p.act) x = xsrc[e.xoff + k] * c_in;
            else if (e.tok == 0 && k >= p.act && k < p.act + 3) x = (k == p.act + 1) ? (cn - cn_hi) : cn_hi;
          } else if (k == hot || k == hot + 1) {
            x = 1.0f;
          }
        }
        v[i] = x;
      }
      st_chunk(atom, e.row, ch, v);
    }
  }
  fence_async_smem();
  tc_fence_before();
  c.arrive(B_A_READY);
}

__device__ void load_vec_async(const Compute& c, uint32_t dst_off, const float* src, int nfloats) {
  for (int i = c.ctid * 4; i < nfloats; i += kComputeThreads * 4) cp_async16(c.sbase + dst_off + i * 4, src + i);
  cp_async_commit();
}

__global__ void __launch_bounds__(kThreads, 1)
fast_sample_kernel(const __grid_constant__ FastParams p, const __grid_constant__ SampleArgs sa) {
  extern __shared__ uint8_t smem_raw[];
  // dynamic shared memory is at least 16-byte aligned; the operand tiles need 1024
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kSmBars + B_COUNT * 8);

  if (threadIdx.x == 0) {
    for (int i = 0; i < B_COUNT; ++i) {
      const bool by_warps = (i == B_A_READY || i == B_ACC_EMPTY0 || i == B_ACC_EMPTY1 || i == B_OP_READY0 || i == B_OP_READY1);
      mbar_init(sbase + kSmBars + i * 8, i == B_Y_READY ? kAttnWarps : (by_warps ? 8 : 1));
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  for (uint32_t i = threadIdx.x; i < (kSmVecA - kSmU) / 16; i += kThreads)      // padding rows must stay finite
    reinterpret_cast<uint4*>(sm + kSmU)[i] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    // ======================= weight-tape producer =======================
    // The whole warp runs the (warp-uniform) loop so that addresses stay in uniform registers;
    // one elected lane issues the copies.
    uint32_t g = 0;
    for (int it = 0; it < my_tiles * p.evals; ++it) {
      uint32_t off = 0;
      for (int f = 0; f < p.n_fills; ++f, ++g) {
        const uint32_t slot = g & (kSlots - 1), par = (g >> 2) & 1u;
        const uint32_t bytes = (uint32_t)p.prog[prog_index(f, p.n_fills)].n8 * 8u * 128u;
        const uint32_t full = sbase + kSmBars + (B_FULL0 + slot) * 8;
        spin_wait(sbase + kSmBars + (B_EMPTY0 + slot) * 8, par ^ 1u);
        if (elect_one()) {
          if (p.timeline != nullptr && blockIdx.x == 0 && it == 1) p.timeline[3 * p.n_fills + f] = clock64();
          mbar_expect_tx(full, bytes);
          bulk_g2s(sbase + kSmRing + slot * kSlotBytes, p.tape + off, bytes, full);
        }
        __syncwarp();
        off += bytes;
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    // Warp-uniform control flow and operands (fill program in kernel-parameter space); one elected
    // lane issues tcgen05.mma / tcgen05.commit.
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    uint32_t phases = (1u << B_ACC_EMPTY0) | (1u << B_ACC_EMPTY1);   // "empty" barriers pass the first time
    uint32_t g = 0;
    for (int it = 0; it < my_tiles * p.evals; ++it) {
      for (int f = 0; f < p.n_fills;) {
        const Fill e = p.prog[prog_index(f, p.n_fills)];
        const bool pair = (e.acc & 2) != 0;            // the next fill sits in the adjacent slot: one wide MMA
        const Fill e2 = p.prog[prog_index(pair ? f + 1 : f, p.n_fills)];
        const bool tl_on = p.timeline != nullptr && blockIdx.x == 0 && it == 1 && lane == 0;
        if (tl_on) p.timeline[3 * f] = clock64();
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const uint32_t id = (e.waits >> (4 * w)) & 0xF;
          if (id != kNone) { spin_wait(sbase + kSmBars + id * 8, (phases >> id) & 1u); phases ^= 1u << id; }
        }
        const uint32_t slot = g & (kSlots - 1), par = (g >> 2) & 1u;
        if (tl_on) p.timeline[3 * f + 1] = clock64();
        spin_wait(sbase + kSmBars + (B_FULL0 + slot) * 8, par);
        if (pair) spin_wait(sbase + kSmBars + (B_FULL0 + slot + 1) * 8, par);
        if (tl_on) p.timeline[3 * f + 2] = clock64();
        tc_fence_after();
        const uint64_t a_desc = smem_desc_sw128(sbase + (uint32_t)e.a_off16 * 16u);
        const uint64_t b_desc = smem_desc_sw128(sbase + kSmRing + slot * kSlotBytes);
        const uint32_t idesc = idesc_bf16_m128(((uint32_t)e.n8 + (pair ? (uint32_t)e2.n8 : 0u)) * 8u);
        const uint32_t d_addr = tm + e.d_col;
        const uint32_t c0 = e2.commits & 0xF, c1 = (e2.commits >> 4) & 0xF;
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            mma_bf16(d_addr, a_desc + 2u * j, b_desc + 2u * j, idesc, ((e.acc & 1) | j) ? 1u : 0u);
          mma_commit(sbase + kSmBars + (B_EMPTY0 + slot) * 8);
          if (pair) mma_commit(sbase + kSmBars + (B_EMPTY0 + slot + 1) * 8);
          if (c0 != kNone) mma_commit(sbase + kSmBars + c0 * 8);
          if (c1 != kNone) mma_commit(sbase + kSmBars + c1 * 8);
        }
        __syncwarp();
        if (tl_on && pair) { p.timeline[3 * f + 3] = p.timeline[3 * f + 4] = p.timeline[3 * f + 5] = clock64(); }
        f += pair ? 2 : 1;
        g += pair ? 2 : 1;
      }
    }
  } else if (warp < kComputeWarp0) {
    // ======================= attention helper warps (2, 3) =======================
    uint32_t y_phase = 1;
    for (int it = 0; it < my_tiles * p.evals * p.L * kH; ++it) {
      attn_sync();
      spin_wait(sbase + kSmBars + B_Y_EMPTY * 8, y_phase);
      y_phase ^= 1u;
      attention_head(sm, sbase, 8 + (warp - 2), lane, p.S, p.T);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(sbase + kSmBars + B_Y_READY * 8);
      attn_sync();
    }
  } else {
    // ======================= compute warps =======================
    Compute c;
    c.sm = sm; c.sbase = sbase; c.tmem = tmem;
    c.ctid = threadIdx.x - kComputeWarp0 * 32;
    c.lane = lane; c.wq = warp & 3; c.hf = (warp - kComputeWarp0) >> 2;
    c.row = c.wq * 32 + lane;
    c.phases = (1u << B_OP_EMPTY0) | (1u << B_OP_EMPTY1) | (1u << B_Y_EMPTY);
    float* vecA = reinterpret_cast<float*>(sm + kSmVecA);
    float* vecM = reinterpret_cast<float*>(sm + kSmVecM);
    float* xbuf = reinterpret_cast<float*>(sm + kSmProg + kProgEntries * 8);
    float* xcur = xbuf, *d1 = xbuf + kXFloats, *x2 = xbuf + 2 * kXFloats, *dU = xbuf + 3 * kXFloats;
    float* sigv = xbuf + 4 * kXFloats;                       // per virtual sequence noise level
    const bool cfg = (p.flags & BESO_FLAG_CFG) != 0;
    const bool inner = (p.flags & BESO_FLAG_INNER) != 0;
    const int nls = cfg ? p.S / 2 : p.S;                     // sequences per tile
    const int n_x = nls * p.t * p.act;
    const size_t layer_stride = kVecAFloats + kVecMFloats;
    load_vec_async(c, kSmVecA, p.vec, kVecAFloats);
    load_vec_async(c, kSmVecM, p.vec + kVecAFloats, kVecMFloats);

    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int seq0 = tile * nls;
      const int ns = min(nls, p.B - seq0);
      const EmbedTask etask = make_embed_task(c, p, tile);
      compute_sync();                                        // previous tile's x fully written out
      for (int i = c.ctid; i < n_x; i += kComputeThreads)
        xcur[i] = (i < ns * p.t * p.act) ? p.xin[(size_t)seq0 * p.t * p.act + i] : 0.f;
      int step = 0;
      bool second = false;
      for (int ev = 0; ev < p.evals; ++ev) {
        const float s_hat = sa.n_steps ? sa.sig[step] : 0.f;
        const float s_next = sa.n_steps ? sa.sig[step + 1] : 0.f;
        const float s_eval = second ? s_next : s_hat;
        for (int i = c.ctid; i < p.S; i += kComputeThreads) {
          const int ls = cfg ? (i >> 1) : i;
          sigv[i] = sa.n_steps ? s_eval : ((seq0 + ls < p.B) ? __ldg(p.sigma + seq0 + ls) : 1.0f);
        }
        compute_sync();
        const float* xsrc = second ? x2 : xcur;
        c.tl = (p.timeline != nullptr && blockIdx.x == 0 && c.ctid == 0 && tile == 0 && ev == 1) ? p.timeline + 4 * p.n_fills : nullptr;
        c.stamp();
        build_embed_input(c, p, etask, xsrc, sigv);
        c.stamp();

        for (int l = 0; l < p.L; ++l) {
          // ---------------- attention half ----------------
          cp_async_wait<0>();
          compute_sync();                                   // vecA(l) (and vecM(l)) landed for everyone
          c.wait(B_X_DONE);
          tc_fence_after();
          c.stamp();
          ln_pass(c, vecA, p, (ev == 0 && tile == (int)blockIdx.x) ? 2 * l : -1);
          c.stamp();
          for (int h = 0; h < kH; ++h) {
            c.wait(B_ACC_FULL0);
            tc_fence_after();
            c.stamp();
            drain_qkv(c, vecA + 3 * kD + h * 192);
            c.arrive(B_ACC_EMPTY0);
            attn_sync();                                    // Q|K|V of this head visible to all 10 attention warps
            c.wait(B_Y_EMPTY);                              // previous head's Y consumed by its proj MMAs
            c.stamp();
            attention_head(sm, sbase, c.ctid >> 5, lane, p.S, p.T);
            fence_async_smem();
            c.arrive(B_Y_READY);
            c.stamp();
            attn_sync();                                    // staging may be overwritten by the next drain
            c.stamp();
          }
          // vecA is free: prefetch the next layer's (or the final block)
          load_vec_async(c, kSmVecA, p.vec + (size_t)(l + 1) * layer_stride, kVecAFloats);
          // ---------------- MLP half ----------------
          c.wait(B_X_DONE);
          tc_fence_after();
          c.stamp();
          ln_pass(c, vecM, p, (ev == 0 && tile == (int)blockIdx.x) ? 2 * l + 1 : -1);
          c.stamp();
          for (int ch = 0; ch < 8; ++ch) {
            const int b = ch & 1;
            c.wait(b ? B_ACC_FULL1 : B_ACC_FULL0);
            tc_fence_after();
            c.wait(b ? B_OP_EMPTY1 : B_OP_EMPTY0);          // H[b] consumed by FC2(ch-2)
            c.stamp();
            drain_gelu(c, b, vecM + 3 * kD + ch * 128);
            c.stamp();
            c.arrive(b ? B_ACC_EMPTY1 : B_ACC_EMPTY0);
            c.arrive(b ? B_OP_READY1 : B_OP_READY0);
          }
          compute_sync();                                   // everyone done with vecM(l)
          const int nl = (l + 1 < p.L) ? l + 1 : 0;
          load_vec_async(c, kSmVecM, p.vec + (size_t)nl * layer_stride + kVecAFloats, kVecMFloats);
        }
        // ---------------- ln_f + action head + pre-conditioning + sampler update ----------------
        cp_async_wait<1>();                                 // final vecA block (vecM(0) may still fly)
        compute_sync();
        c.wait(B_X_DONE);
        tc_fence_after();
        c.stamp();
        ln_pass(c, vecA, p, (ev == 0 && tile == (int)blockIdx.x) ? 2 * p.L : -1);
        c.stamp();
        c.wait(B_ACC_FULL0);
        tc_fence_after();
        c.stamp();
        float pr[16];
        tmem_ld16(c.lane_addr(kColS0), pr);
        tmem_wait_ld();
        tc_fence_before();
        c.arrive(B_ACC_EMPTY0);
        const float* hb = vecA + 3 * kD;
        const int vs = c.row / p.T, tok = c.row - vs * p.T;
        const int j = tok - 1 - p.G;
        const int ls = cfg ? (vs >> 1) : vs;
        const bool act_row = (c.hf == 0) && vs < p.S && tok > p.G && (j & 1) && (ls < ns);
        const int xo = (ls * p.t + (j >> 1)) * p.act;
        float dval[kMaxAct];
        if (act_row) {
          const float sg = sigv[vs];
          const float den = sg * sg + p.sigma_data * p.sigma_data;
          const float c_skip = p.sigma_data * p.sigma_data / den, c_out = sg * p.sigma_data / sqrtf(den);
#pragma unroll
          for (int a = 0; a < kMaxAct; ++a) {
            if (a < p.act) {
              const float f = pr[a] + hb[a];
              dval[a] = inner ? f : __fadd_rn(__fmul_rn(f, c_out), __fmul_rn(xsrc[xo + a], c_skip));
            }
          }
          if (cfg && (vs & 1)) {
#pragma unroll
            for (int a = 0; a < kMaxAct; ++a) if (a < p.act) dU[xo + a] = dval[a];
          }
        }
        if (cfg) compute_sync();
        if (act_row && !(cfg && (vs & 1))) {
#pragma unroll
          for (int a = 0; a < kMaxAct; ++a) {
            if (a < p.act) {
              float D = dval[a];
              if (cfg) D = __fadd_rn(dU[xo + a], __fmul_rn(p.lambda, __fsub_rn(D, dU[xo + a])));
              const int i = xo + a;
              if (sa.n_steps == 0) {
                p.out[(size_t)seq0 * p.t * p.act + i] = D;
              } else if (sa.sampler == BESO_SAMPLER_DDIM) {
                xcur[i] = __fsub_rn(__fmul_rn(sa.ca[step], xcur[i]), __fmul_rn(sa.ce[step], D));
              } else {
                const float dt = __fsub_rn(s_next, s_hat);
                if (!second) {
                  const float dd = __fdiv_rn(__fsub_rn(xcur[i], D), s_hat);
                  const float xe = __fadd_rn(xcur[i], __fmul_rn(dd, dt));
                  if (sa.sampler == BESO_SAMPLER_HEUN && s_next != 0.0f) { d1[i] = dd; x2[i] = xe; } else xcur[i] = xe;
                } else {
                  const float d2 = __fdiv_rn(__fsub_rn(x2[i], D), s_next);
                  xcur[i] = __fadd_rn(xcur[i], __fmul_rn(__fdiv_rn(__fadd_rn(d1[i], d2), 2.0f), dt));
                }
              }
            }
          }
        }
        if (sa.n_steps) {
          if (!second && sa.sampler == BESO_SAMPLER_HEUN && s_next != 0.0f) second = true;
          else { second = false; ++step; }
        }
        // vecA is free again: first block of the next evaluation
        compute_sync();
        load_vec_async(c, kSmVecA, p.vec, kVecAFloats);
      }
      if (sa.n_steps) {
        compute_sync();
        for (int i = c.ctid; i < ns * p.t * p.act; i += kComputeThreads) p.out[(size_t)seq0 * p.t * p.act + i] = xcur[i];
      }
    }
    cp_async_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ================================ weight packing ===================================================
struct PackTile {       // one [rows x 64] bf16 SW128 sub-tile of the tape from a row-major fp32 matrix
  const float* src; int ld, row0, col0, rows, valid_rows, valid_cols; float scale; uint32_t dst;
};
__global__ void pack_tiles_kernel(const PackTile* tiles, uint8_t* tape) {
  const PackTile t = tiles[blockIdx.x];
  for (int idx = threadIdx.x; idx < t.rows * 8; idx += blockDim.x) {
    const int r = idx >> 3, chunk = idx & 7;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = chunk * 8 + i;
      v[i] = (r < t.valid_rows && k < t.valid_cols) ? t.src[(size_t)(t.row0 + r) * t.ld + t.col0 + k] * t.scale : 0.f;
    }
    st_chunk(tape + t.dst, r, chunk, v);
  }
}

struct EmbSrc {
  const float *pos, *tokw, *tokb, *sigw, *sigb, *actw, *actb;
  const float* resid_bias[2 * kMaxLayers];     // attn.proj.bias and mlp.2.bias of every layer
  int obs, act, G, W, L;
};
// Embedding GEMM B operand: W_emb[n][k], n < 256, k < 128 (atom 0 = obs, atom 1 = misc), as 8 fills
// [128 rows x 64] in (k-atom, row-half) order.
__global__ void pack_emb_kernel(EmbSrc s, uint8_t* tape) {
  const int fill = blockIdx.x;                 // 0..3: atom = fill >> 1, rows (fill & 1) * 128 ..
  const int atom = fill >> 1, n0 = (fill & 1) * 128;
  for (int idx = threadIdx.x; idx < 128 * 8; idx += blockDim.x) {
    const int r = idx >> 3, chunk = idx & 7, n = n0 + r;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = chunk * 8 + i;
      float x = 0.f;
      if (atom == 0) {
        if (k < s.obs) x = s.tokw[(size_t)n * s.obs + k];
      } else if (k < s.act) {
        x = s.actw[(size_t)n * s.act + k];
      } else if (k < s.act + 3) {
        const float w = s.sigw[n];
        const float hi = __bfloat162float(__float2bfloat16_rn(w));
        x = (k == s.act + 2) ? (w - hi) : hi;    // A holds [cn_hi, cn_lo, cn_hi]
      } else if (k >= kOneHot0 && k < kOneHot0 + 2 * kMaxTokens) {
        const int tok = (k - kOneHot0) >> 1;
        float tbl;
        if (tok == 0) tbl = s.sigb[n];
        else if (tok <= s.G) tbl = s.tokb[n] + s.pos[(size_t)(tok - 1) * kD + n];
        else {
          const int j = tok - 1 - s.G, step = j >> 1;
          tbl = (step < s.W) ? ((j & 1) ? s.actb[n] : s.tokb[n]) + s.pos[(size_t)(s.G + step) * kD + n] : 0.f;
        }
        for (int i2 = 0; i2 < 2 * s.L; ++i2) tbl += s.resid_bias[i2][n];   // all residual-branch biases, up front
        const float hi = __bfloat162float(__float2bfloat16_rn(tbl));
        x = ((k - kOneHot0) & 1) ? (tbl - hi) : hi;
      }
      v[i] = x;
    }
    st_chunk(tape + (size_t)fill * 16384, r, chunk, v);
  }
}

// pend vectors: minus the residual-branch biases that the embedding GEMM added too early.
//   LN1(l): -(sum_{l' >= l} bproj + b2)     LN2(l): LN1(l) + bproj_l      ln_f: 0
__global__ void pack_pend_kernel(EmbSrc s, float* vec, uint32_t layer_stride, uint32_t m_off) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= kD) return;
  float r = 0.f;
  vec[(size_t)s.L * layer_stride + n] = 0.f;
  for (int l = s.L - 1; l >= 0; --l) {
    const float bp = s.resid_bias[2 * l][n], b2 = s.resid_bias[2 * l + 1][n];
    vec[(size_t)l * layer_stride + m_off + n] = -(r + b2);
    r += bp + b2;
    vec[(size_t)l * layer_stride + n] = -r;
  }
}

struct VecCopy { const float* src; uint32_t dst; int n; float scale; };
__global__ void pack_vec_kernel(const VecCopy* cp, float* vec) {
  const VecCopy c = cp[blockIdx.x];
  for (int i = threadIdx.x; i < c.n; i += blockDim.x) vec[c.dst + i] = c.src ? c.src[i] * c.scale : 0.f;
}

// ---- fill program (host) ---------------------------------------------------------------------------
uint8_t pair(uint8_t a, uint8_t b) { return (uint8_t)((a & 0xF) | ((b & 0xF) << 4)); }

std::vector<Fill> build_program() {
  std::vector<Fill> v;
  auto add = [&](uint32_t a_off, uint32_t d_col, int n, int acc, uint8_t w0, uint8_t w1, uint8_t c0, uint8_t c1) {
    Fill f;
    f.a_off16 = (uint16_t)(a_off / 16); f.d_col = (uint16_t)d_col; f.n8 = (uint8_t)(n / 8); f.acc = (uint8_t)acc;
    f.waits = pair(w0, w1); f.commits = pair(c0, c1);
    v.push_back(f);
  };
  // embedding: X = A_emb (atoms 0,1) * W_emb^T
  for (int kb = 0; kb < 2; ++kb)
    for (int half = 0; half < 2; ++half)
      add(kSmA + kb * 16384, kColX + half * 128, 128, (kb > 0) | (half == 0 ? 2 : 0), (kb == 0 && half == 0) ? B_A_READY : kNone, kNone,
          (kb == 1 && half == 1) ? B_X_DONE : kNone, kNone);
  auto qkv = [&](int h) {
    for (int kb = 0; kb < 4; ++kb) {
      add(kSmA + kb * 16384, kColS0, 128, (kb > 0) | 2, kb == 0 ? B_ACC_EMPTY0 : kNone, (kb == 0 && h == 0) ? B_A_READY : kNone, kNone, kNone);
      add(kSmA + kb * 16384, kColS1, 64, kb > 0, kNone, kNone, kb == 3 ? B_ACC_FULL0 : kNone, kNone);
    }
  };
  auto proj = [&](int h) {
    add(kSmY, kColX, 128, 1 | 2, B_Y_READY, kNone, kNone, kNone);
    add(kSmY, kColX + 128, 128, 1, kNone, kNone, B_Y_EMPTY, h == 3 ? B_X_DONE : kNone);
  };
  auto fc1 = [&](int c) {
    const int b = c & 1;
    for (int kb = 0; kb < 4; ++kb)
      add(kSmA + kb * 16384, b ? kColS1 : kColS0, 128, kb > 0, kb == 0 ? (b ? B_ACC_EMPTY1 : B_ACC_EMPTY0) : kNone,
          (kb == 0 && c == 0) ? B_A_READY : kNone, kb == 3 ? (b ? B_ACC_FULL1 : B_ACC_FULL0) : kNone, kNone);
  };
  auto fc2 = [&](int c) {
    const int b = c & 1;
    for (int kb = 0; kb < 2; ++kb)
      for (int half = 0; half < 2; ++half)
        add((b ? kSmH1 : kSmH0) + kb * 16384, kColX + half * 128, 128, 1 | (half == 0 ? 2 : 0),
            (kb == 0 && half == 0) ? (b ? B_OP_READY1 : B_OP_READY0) : kNone, kNone,
            (kb == 1 && half == 1) ? (b ? B_OP_EMPTY1 : B_OP_EMPTY0) : kNone, (kb == 1 && half == 1 && c == 7) ? B_X_DONE : kNone);
  };
  {                                 // one transformer block; every layer replays it
    qkv(0); qkv(1); proj(0); qkv(2); proj(1); qkv(3); proj(2); proj(3);
    fc1(0); fc1(1); fc2(0);
    for (int c = 2; c < 8; ++c) { fc1(c); fc2(c - 1); }
    fc2(7);
  }
  for (int kb = 0; kb < 4; ++kb)   // action head, N = 16
    add(kSmA + kb * 16384, kColS0, 16, kb > 0, kb == 0 ? B_A_READY : kNone, kb == 0 ? B_ACC_EMPTY0 : kNone,
        kb == 3 ? B_ACC_FULL0 : kNone, kNone);
  return v;
}

}  // namespace

// ================================ host API =========================================================
bool fast_supported(const beso_model_desc& m) {
  const int G = m.goal_conditioned ? m.goal_len : 0;
  const int T = 1 + G + 2 * m.window;
  return m.d == kD && m.n_heads == kH && m.linear_output && m.n_layers <= kMaxLayers && m.obs_dim <= kMaxObs &&
         m.act_dim <= kMaxAct && T <= kMaxTokens;
}

int fast_seqs_per_tile(const beso_model_desc& m, int t) {
  if (!fast_supported(m)) return 0;
  const int G = m.goal_conditioned ? m.goal_len : 0;
  return kRows / (1 + G + 2 * t);
}

void fast_free(FastWeights& w) {
  if (w.tape) cudaFree(w.tape);
  if (w.vec) cudaFree(w.vec);
  w.tape = nullptr; w.vec = nullptr;
}

// Parameter indices in parameters() order
static int p_layer(int l, int k) { return 3 + l * 16 + k; }   // k: 0 ln1w 1 ln1b 2 ln2w 3 ln2b 4 key.w 5 key.b 6 query.w 7 query.b
                                                               //    8 value.w 9 value.b 10 proj.w 11 proj.b 12 mlp0.w 13 mlp0.b 14 mlp2.w 15 mlp2.b

int fast_pack(FastWeights& w, const beso_model_desc& m, const float* const* prm, cudaStream_t st) {
  const int L = m.n_layers, G = m.goal_conditioned ? m.goal_len : 0;
  const std::vector<Fill> prog = build_program();
  if ((int)prog.size() != kProgEntries) { set_error("internal: fill program size"); return BESO_E_INVALID; }
  size_t tape_bytes = 0;
  for (int i = 0; i < kProgEntries; ++i)
    tape_bytes += (size_t)prog[i].n8 * 8 * 128 * ((i >= 4 && i < 108) ? L : 1);
  const size_t vec_floats = (size_t)L * (kVecAFloats + kVecMFloats) + kVecAFloats;
  const size_t prog_bytes = kProgEntries * sizeof(Fill);
  if (!w.tape) {
    // tape | program | (scratch tables for the pack kernels)
    BESO_CUDA(cudaMalloc(&w.tape, tape_bytes + prog_bytes + (1 << 20)));
    BESO_CUDA(cudaMalloc(&w.vec, vec_floats * sizeof(float)));
    w.tape_bytes = tape_bytes; w.vec_floats = vec_floats;
  }
  uint8_t* tape = reinterpret_cast<uint8_t*>(w.tape);
  uint8_t* scratch = tape + tape_bytes + prog_bytes;
  BESO_CUDA(cudaMemcpyAsync(tape + tape_bytes, prog.data(), prog_bytes, cudaMemcpyHostToDevice, st));

  // ---- tape sub-tiles, in program order ----
  std::vector<PackTile> tiles;
  uint32_t off = 4 * 16384;                                   // embedding fills are written by pack_emb_kernel
  auto tile = [&](const float* src, int ld, int row0, int col0, int rows, int vrows, float scale) {
    tiles.push_back({src, ld, row0, col0, rows, vrows, 64, scale, off});
    off += rows * 128;
  };
  const float qscale = 0.125f;                                // 1 / sqrt(64): exact power of two
  for (int l = 0; l < L; ++l) {
    const float *wk = prm[p_layer(l, 4)], *wq = prm[p_layer(l, 6)], *wv = prm[p_layer(l, 8)], *wp = prm[p_layer(l, 10)];
    const float *w1 = prm[p_layer(l, 12)], *w2 = prm[p_layer(l, 14)];
    auto qkv = [&](int h) {
      for (int kb = 0; kb < 4; ++kb) {
        tile(wq, kD, h * 64, kb * 64, 64, 64, qscale);
        tile(wk, kD, h * 64, kb * 64, 64, 64, 1.f);
        tile(wv, kD, h * 64, kb * 64, 64, 64, 1.f);
      }
    };
    auto proj = [&](int h) { for (int half = 0; half < 2; ++half) for (int s = 0; s < 2; ++s) tile(wp, kD, half * 128 + s * 64, h * 64, 64, 64, 1.f); };
    auto fc1 = [&](int c) { for (int kb = 0; kb < 4; ++kb) for (int s = 0; s < 2; ++s) tile(w1, kD, c * 128 + s * 64, kb * 64, 64, 64, 1.f); };
    auto fc2 = [&](int c) {
      for (int kb = 0; kb < 2; ++kb) for (int half = 0; half < 2; ++half) for (int s = 0; s < 2; ++s)
        tile(w2, kFF, half * 128 + s * 64, c * 128 + kb * 64, 64, 64, 1.f);
    };
    qkv(0); qkv(1); proj(0); qkv(2); proj(1); qkv(3); proj(2); proj(3);
    fc1(0); fc1(1); fc2(0);
    for (int c = 2; c < 8; ++c) { fc1(c); fc2(c - 1); }
    fc2(7);
  }
  const int p_tail = 3 + 16 * L;                              // ln_f.w, ln_f.b, sigma_emb.w/b, action_emb.w/b, action_pred.w/b
  for (int kb = 0; kb < 4; ++kb) tile(prm[p_tail + 6], kD, 0, kb * 64, 16, m.act_dim, 1.f);
  if (off != tape_bytes) { set_error("internal: tape layout mismatch"); return BESO_E_INVALID; }
  if (tiles.size() * sizeof(PackTile) > (1 << 19)) { set_error("internal: pack table too large"); return BESO_E_INVALID; }
  BESO_CUDA(cudaMemcpyAsync(scratch, tiles.data(), tiles.size() * sizeof(PackTile), cudaMemcpyHostToDevice, st));
  pack_tiles_kernel<<<(unsigned)tiles.size(), 128, 0, st>>>(reinterpret_cast<const PackTile*>(scratch), tape);
  ++g_kernel_launches;
  BESO_CUDA(cudaGetLastError());
  EmbSrc es{};
  es.pos = prm[0]; es.tokw = prm[1]; es.tokb = prm[2]; es.sigw = prm[p_tail + 2]; es.sigb = prm[p_tail + 3];
  es.actw = prm[p_tail + 4]; es.actb = prm[p_tail + 5];
  es.obs = m.obs_dim; es.act = m.act_dim; es.G = G; es.W = m.window; es.L = L;
  for (int l = 0; l < L; ++l) { es.resid_bias[2 * l] = prm[p_layer(l, 11)]; es.resid_bias[2 * l + 1] = prm[p_layer(l, 15)]; }
  pack_emb_kernel<<<4, 256, 0, st>>>(es, tape);
  ++g_kernel_launches;
  BESO_CUDA(cudaGetLastError());

  // ---- fp32 vectors ----
  std::vector<VecCopy> vc;
  for (int l = 0; l < L; ++l) {
    const uint32_t a = (uint32_t)(l * (kVecAFloats + kVecMFloats)), mo = a + kVecAFloats;
    vc.push_back({prm[p_layer(l, 0)], a + kD, kD, 1.f});
    vc.push_back({prm[p_layer(l, 1)], a + 2 * kD, kD, 1.f});
    for (int h = 0; h < kH; ++h) {
      vc.push_back({prm[p_layer(l, 7)] + h * 64, a + 3 * kD + h * 192, 64, qscale});
      vc.push_back({prm[p_layer(l, 5)] + h * 64, a + 3 * kD + h * 192 + 64, 64, 1.f});
      vc.push_back({prm[p_layer(l, 9)] + h * 64, a + 3 * kD + h * 192 + 128, 64, 1.f});
    }
    vc.push_back({prm[p_layer(l, 2)], mo + kD, kD, 1.f});
    vc.push_back({prm[p_layer(l, 3)], mo + 2 * kD, kD, 1.f});
    vc.push_back({prm[p_layer(l, 13)], mo + 3 * kD, kFF, 1.f});
  }
  const uint32_t fa = (uint32_t)(L * (kVecAFloats + kVecMFloats));
  vc.push_back({prm[p_tail], fa + kD, kD, 1.f});
  vc.push_back({prm[p_tail + 1], fa + 2 * kD, kD, 1.f});
  vc.push_back({prm[p_tail + 7], fa + 3 * kD, m.act_dim, 1.f});
  vc.push_back({nullptr, fa + 3 * kD + (uint32_t)m.act_dim, 768 - m.act_dim, 1.f});
  uint8_t* scratch2 = scratch + (1 << 19);
  BESO_CUDA(cudaMemcpyAsync(scratch2, vc.data(), vc.size() * sizeof(VecCopy), cudaMemcpyHostToDevice, st));
  pack_vec_kernel<<<(unsigned)vc.size(), 128, 0, st>>>(reinterpret_cast<const VecCopy*>(scratch2), w.vec);
  pack_pend_kernel<<<1, kD, 0, st>>>(es, w.vec, kVecAFloats + kVecMFloats, kVecAFloats);
  g_kernel_launches += 2;
  BESO_CUDA(cudaGetLastError());
  BESO_CUDA(cudaStreamSynchronize(st));                       // the host tables above go out of scope
  return BESO_OK;
}

static float* g_trace = nullptr;
static long long* g_timeline = nullptr;
void fast_set_trace(float* trace_dev) { g_trace = trace_dev; }
void fast_set_timeline(long long* dev) { g_timeline = dev; }

int fast_launch(const FastWeights& w, const beso_model_desc& m, int sm_count, const SampleArgs& sa,
                const float* state, const float* goal, const float* x, const float* sigma, float* out,
                int B, int t, uint32_t flags, float lambda, cudaStream_t st) {
  if (!w.tape) { set_error("fast weights not packed"); return BESO_E_NOT_PACKED; }
  FastParams p{};
  const int L = m.n_layers;
  const size_t n_fills = 4 + (size_t)L * 104 + 4;
  p.tape = reinterpret_cast<const uint8_t*>(w.tape);
  {
    const std::vector<Fill> prog = build_program();
    memcpy(p.prog, prog.data(), sizeof(p.prog));
  }
  p.vec = w.vec;
  p.n_fills = (int)n_fills; p.L = L; p.G = m.goal_conditioned ? m.goal_len : 0; p.obs = m.obs_dim; p.act = m.act_dim;
  p.t = t; p.T = 1 + p.G + 2 * t;
  p.S = kRows / p.T;
  const bool cfg = flags & BESO_FLAG_CFG;
  if (cfg) p.S &= ~1;                                         // cond / uncond pairs share a tile
  if (p.S < 1) { set_error("sequence does not fit a 128-row tile"); return BESO_E_UNSUPPORTED; }
  const int per_tile = cfg ? p.S / 2 : p.S;
  p.n_tiles = (B + per_tile - 1) / per_tile;
  p.B = B;
  p.evals = 1;
  if (sa.n_steps) {
    p.evals = sa.n_steps;
    if (sa.sampler == BESO_SAMPLER_HEUN)
      for (int i = 0; i < sa.n_steps; ++i) if (sa.sig[i + 1] != 0.0f) ++p.evals;
  }
  p.flags = flags; p.lambda = lambda; p.sigma_data = m.sigma_data;
  p.state = state; p.goal = goal; p.xin = x; p.sigma = sigma; p.out = out;
  p.trace = g_trace;
  p.timeline = g_timeline;
  static bool configured = false;
  if (!configured) {
    BESO_CUDA(cudaFuncSetAttribute(fast_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes + 1024));
    configured = true;
  }
  const int grid = p.n_tiles < sm_count ? p.n_tiles : sm_count;
  fast_sample_kernel<<<grid, kThreads, kSmemBytes + 1024, st>>>(p, sa);
  ++g_kernel_launches;
  BESO_CUDA(cudaGetLastError());
  return BESO_OK;
}

}  // namespace beso
