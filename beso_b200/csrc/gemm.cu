// Training GEMMs on the tcgen05 tensor cores (replaces the library SGEMM of round 1).
//
//   C[M][N] = A . B^T  (+ bias + residual | + C | erf-GELU second output),   fp32 in HBM, fp32 accumulate in TMEM
//
// One persistent, warp-specialised CTA per SM:
//   warps 4-11  producers: read the fp32 operand tiles from global memory (coalesced along whichever dimension is
//               contiguous: the "NT" forward products, the "NN" data gradients and the "TN" weight gradients all
//               arrive here), convert to bf16 -- in the fp32-parity mode to a bf16 hi image and a bf16 lo image of
//               the remainder -- and store them as K-major SWIZZLE_128B operand tiles ([rows x 64], the layout the
//               forward kernel uses), through a ring of shared-memory stages guarded by mbarriers;
//   warp 12     issues tcgen05.mma (M = 128, N = 128 / 256, K = 16 per instruction; parity mode: hi.hi + lo.hi + hi.lo
//               into the same accumulator) and commits stages / accumulators;
//   warps 0-3   epilogue: TMEM -> registers -> a padded shared-memory transpose -> coalesced 128-byte row segments
//               with bias, residual, accumulate or GELU applied on the way; two accumulators (2 x BN TMEM columns)
//               so the epilogue of tile i overlaps the main loop of tile i + 1.
// Weight gradients contract over the B * T rows (K ~ 10^5) into small outputs: split-K over CTAs into partial
// buffers and a second kernel that adds them in a fixed order (deterministic).
//
// Why bf16 and not fp16 here: gradients of a batch-mean loss are ~1e-6 and below, outside fp16's range; bf16 keeps
// fp32's exponent.  Operand precision is bought with images: x = h + m + l (three bf16 values, 24 mantissa bits,
// exact remainders) keeps every cross term down to 2^-24 (hh, hm, mh, mm, hl, lh: six MMAs); x = h + m (three MMAs)
// carries 16 bits; one image is plain bf16 mixed-precision arithmetic.  Measured on the gradient goldens: within
// 1.5e-6 of each tensor's scale with three images (the default, fp32-parity mode) and 1.5e-5 with two; against fp64
// products of random matrices 2.5e-6 / 6e-6 of the output scale (the floor is the tensor core's fp32 accumulation).
#include <cuda_bf16.h>

#include <string>

#include "common.cuh"
#include "gemm.cuh"
#include "umma.cuh"

namespace beso {

using namespace umma;

namespace {

constexpr int kBM = 128, kBK = 64;
constexpr int kEpiWarps = 4, kProdWarps = 8, kMmaWarpG = 12, kThreadsG = 13 * 32;
constexpr int kProdThreads = kProdWarps * 32;
constexpr uint32_t kStagePitch = 144;                 // bytes per row of the epilogue transpose buffer (32 floats + 4 pad)
constexpr uint32_t kStageBufBytes = 32 * kStagePitch; // per epilogue warp

// IMG = bf16 images per operand: 1 (x ~ h), 2 (x ~ h + m: 16 mantissa bits), 3 (x ~ h + m + l: 24 bits = fp32)
// BRES: the B operand of this CTA's column tile (all of K <= 256) is converted once and stays resident in shared
// memory; the pipeline stages then carry A tiles only.  The GEMMs with K = embed_dim (QKV, projection, FC1 and the
// data gradients through W2 / the d x d weights) stream the activations exactly once this way instead of re-reading
// and re-converting the same 256 x 256 weight tile for every 128-row tile.
constexpr uint32_t kResKb = 4;                        // k-blocks the resident B operand holds (K <= 256)
template <int BN, int IMG, bool BRES>
struct Cfg {
  static constexpr uint32_t a_bytes = kBM * 128u, b_bytes = BN * 128u;
  static constexpr uint32_t images = (uint32_t)IMG;
  static constexpr uint32_t b_res = BRES ? kResKb * images * b_bytes : 0u;       // [kb][image][BN x 64]
  static constexpr uint32_t stage_bytes = images * (a_bytes + (BRES ? 0u : b_bytes));
  static constexpr uint32_t stages_raw = (196608u - b_res) / stage_bytes;
  static constexpr uint32_t stages = stages_raw > 6u ? 6u : stages_raw;
  static constexpr uint32_t sm_stage0 = b_res;
  static constexpr uint32_t sm_epi = b_res + stages * stage_bytes;
  static constexpr uint32_t sm_bars = sm_epi + kEpiWarps * kStageBufBytes;
  static constexpr uint32_t smem = sm_bars + 256u;
};

struct Tile { int m0, n0, kb0, kb1, ks; };
__device__ __forceinline__ Tile tile_of(int idx, int nt, int mt, int kb_total, int kb_per) {
  Tile t;
  const int n_t = idx % nt, r = idx / nt;
  const int m_t = r % mt;
  t.ks = r / mt;
  t.m0 = m_t * kBM; t.n0 = n_t;                     // n0 is scaled by BN at the use site
  t.kb0 = t.ks * kb_per;
  t.kb1 = min(kb_total, t.kb0 + kb_per);
  return t;
}

__device__ __forceinline__ bool elect_one_g() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}
// Bounded wait: a protocol bug must trap, not hang the GPU.
__device__ __noinline__ void wait_slow(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  uint32_t polls = 0;
  while (!mbar_try_wait_sleep(bar, parity)) {
    if ((++polls & 63u) == 0 && clock64() - t0 > 8000000000ll) {
      printf("beso gemm kernel: mbarrier at %u parity %u timed out (block %d thread %d)\n", bar, parity, blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  wait_slow(bar, parity);
}

// 8 consecutive K values -> one 16-byte chunk of each bf16 image: h = bf16(x), m = bf16(x - h), l = bf16(x - h - m)
// (the remainders are exact in fp32); image i of the tile lies i * img_delta bytes after the first
template <int IMG>
__device__ __forceinline__ void store_chunk(uint32_t addr, uint32_t img_delta, const float (&v)[8]) {
  float r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = v[i];
#pragma unroll
  for (int im = 0; im < IMG; ++im) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 b = __floats2bfloat162_rn(r[2 * i], r[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&b);
      if (im + 1 < IMG) {
        const float2 f = __bfloat1622float2(b);
        r[2 * i] -= f.x; r[2 * i + 1] -= f.y;
      }
    }
    sts128(addr + (uint32_t)im * img_delta, w[0], w[1], w[2], w[3]);
  }
}

// One operand tile [ROWS x 64] (element (row, k)) by the 256 producer threads, in two halves so that a whole pipeline
// stage of global loads (ROWS / 32 tasks of 8 values per thread: 96 KB per SM for a 128 + 256 row stage) is in flight
// while the producers wait for their shared-memory slot -- the memory latency is paid once per stage, not per batch.
//   KMAJOR: element at src[(row0 + row) * ld + k]     (k contiguous: 8 lanes read the 256 bytes of one row)
//  !KMAJOR: element at src[k * ld + row0 + row]       (rows contiguous: a warp reads 32 rows of one k, 128 bytes)
// Rows >= n_rows and k >= K are zero.  vec: 16-byte loads are legal (ld % 4 == 0, base 16-byte aligned).
template <int ROWS, bool KMAJOR>
__device__ __forceinline__ void tile_issue(float (&v)[ROWS / 32][8], const float* __restrict__ src, int ld, int row0, int n_rows,
                                           int k0, int K, bool vec, int ptid) {
  constexpr int kTasks = ROWS / 32;
  if (KMAJOR) {
#pragma unroll
    for (int u = 0; u < kTasks; ++u) {
      const int task = ptid + u * kProdThreads, row = task >> 3, k = k0 + (task & 7) * 8;
      const bool rv = row0 + row < n_rows;
      const float* g = src + (size_t)(row0 + row) * ld + k;
      if (rv && vec && k + 8 <= K) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(g)), b = __ldg(reinterpret_cast<const float4*>(g) + 1);
        v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w; v[u][4] = b.x; v[u][5] = b.y; v[u][6] = b.z; v[u][7] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[u][j] = (rv && k + j < K) ? __ldg(g + j) : 0.f;
      }
    }
  } else {
    const int pw = ptid >> 5, lane = ptid & 31;            // task u of warp pw: 32-row group u, k chunk pw
    const int k = k0 + pw * 8;
#pragma unroll
    for (int u = 0; u < kTasks; ++u) {
      const int row = u * 32 + lane;
      const bool rv = row0 + row < n_rows;
      const float* g = src + (size_t)k * ld + row0 + row;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[u][j] = (rv && k + j < K) ? __ldg(g + (size_t)j * ld) : 0.f;
    }
  }
}
template <int ROWS, bool KMAJOR, int IMG>
__device__ __forceinline__ void tile_store(const float (&v)[ROWS / 32][8], uint32_t smem_hi, uint32_t img_delta, int ptid) {
  constexpr int kTasks = ROWS / 32;
#pragma unroll
  for (int u = 0; u < kTasks; ++u) {
    uint32_t row, chunk;
    if (KMAJOR) { const int task = ptid + u * kProdThreads; row = (uint32_t)(task >> 3); chunk = (uint32_t)(task & 7); }
    else { row = (uint32_t)(u * 32 + (ptid & 31)); chunk = (uint32_t)(ptid >> 5); }
    store_chunk<IMG>(smem_hi + sw128_offset(row, chunk), img_delta, v[u]);
  }
}

struct KArgs {
  GemmArgs g;
  int mt, nt, ksplit, kb_total, kb_per, n_tiles;
  int vec_a, vec_b, vec_c;
  float* partial;          // split-K: [ksplit][M][N]
};

__device__ __forceinline__ float gelu_erf_g(float u) { return 0.5f * u * (1.0f + erff(u * 0.70710678118654752440f)); }

template <int BN, int IMG, bool AK, bool BK, bool BRES>
__global__ void __launch_bounds__(kThreadsG, 1) gemm_kernel(const __grid_constant__ KArgs ka) {
  using C = Cfg<BN, IMG, BRES>;
  static_assert(C::stages >= 2, "pipeline needs two stages");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bars = sbase + C::sm_bars;
  // barrier map: full[s] at s, empty[s] at 8 + s, acc_full[a] at 16 + a, acc_empty[a] at 18 + a
  auto bar_full = [&](uint32_t s) { return bars + s * 8u; };
  auto bar_empty = [&](uint32_t s) { return bars + (8u + s) * 8u; };
  auto bar_accf = [&](uint32_t a) { return bars + (16u + a) * 8u; };
  auto bar_acce = [&](uint32_t a) { return bars + (18u + a) * 8u; };
  const uint32_t bar_bres = bars + 20u * 8u;          // resident B operand complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + C::sm_bars + 21 * 8);
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < C::stages; ++s) { mbar_init(bar_full(s), kProdWarps); mbar_init(bar_empty(s), 1); }
    for (uint32_t a = 0; a < 2; ++a) { mbar_init(bar_accf(a), 1); mbar_init(bar_acce(a), kEpiWarps); }
    mbar_init(bar_bres, kProdWarps);
    fence_barrier_init();
  }
  if (warp == kMmaWarpG) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const GemmArgs& g = ka.g;

  if (warp >= kEpiWarps && warp < kMmaWarpG) {
    // ===================================== producers =====================================
    // The global loads of the next kDepth pipeline steps are in flight in registers while the producers wait for
    // shared-memory slots and convert: with a resident B operand a step is one 32 KB A tile, so three steps
    // (96 registers per thread, 96 KB per SM) are kept in flight; with A + B per step (96 registers) one.
    constexpr int kDepth = BRES ? 3 : 1;
    const int ptid = threadIdx.x - kEpiWarps * 32;
    float va[kDepth][kBM / 32][8], vb[BRES ? 1 : kDepth][BN / 32][8];
    struct Cur { int idx, kb; Tile t; bool have; };
    auto start = [&]() {
      Cur c;
      c.idx = blockIdx.x;
      c.have = c.idx < ka.n_tiles;
      c.t = tile_of(c.have ? c.idx : 0, ka.nt, ka.mt, ka.kb_total, ka.kb_per);
      c.kb = c.t.kb0;
      return c;
    };
    auto advance = [&](Cur& c) {
      if (++c.kb >= c.t.kb1) {
        c.idx += gridDim.x;
        c.have = c.idx < ka.n_tiles;
        if (c.have) { c.t = tile_of(c.idx, ka.nt, ka.mt, ka.kb_total, ka.kb_per); c.kb = c.t.kb0; }
      }
    };
    Cur ci = start(), cs = ci;
    if constexpr (BRES) {
      // every tile of this CTA has the same column tile (grid % nt == 0): its B operand, once
      if (ci.have) {
        for (int rk = 0; rk < ka.kb_total; ++rk) {
          tile_issue<BN, BK>(vb[0], g.B, g.ldb, ci.t.n0 * BN, g.N, rk * kBK, g.K, ka.vec_b != 0, ptid);
          tile_store<BN, BK, IMG>(vb[0], sbase + (uint32_t)rk * C::images * C::b_bytes, C::b_bytes, ptid);
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_bres);
    }
#pragma unroll
    for (int d = 0; d < kDepth; ++d) {
      if (ci.have) {
        tile_issue<kBM, AK>(va[d], g.A, g.lda, ci.t.m0, g.M, ci.kb * kBK, g.K, ka.vec_a != 0, ptid);
        if constexpr (!BRES) tile_issue<BN, BK>(vb[d], g.B, g.ldb, ci.t.n0 * BN, g.N, ci.kb * kBK, g.K, ka.vec_b != 0, ptid);
        advance(ci);
      }
    }
    uint32_t it = 0;
    while (cs.have) {
#pragma unroll
      for (int d = 0; d < kDepth; ++d) {
        if (!cs.have) break;
        const uint32_t s = it % C::stages, par = (it / C::stages) & 1u;
        wait_bar(bar_empty(s), par ^ 1u);
        const uint32_t st = sbase + C::sm_stage0 + s * C::stage_bytes;
        tile_store<kBM, AK, IMG>(va[d], st, C::a_bytes, ptid);
        if constexpr (!BRES) tile_store<BN, BK, IMG>(vb[d], st + C::images * C::a_bytes, C::b_bytes, ptid);
        // No proxy fence here: fence.proxy.async compiles to MEMBAR.ALL.CTA, which would wait for this thread's
        // prefetched global loads of the following steps and serialise the pipeline on the memory latency.  The
        // stores are released by the mbarrier arrive below and the MMA warp fences the proxies after its acquire.
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full(s));
        ++it;
        advance(cs);
        if (ci.have) {
          tile_issue<kBM, AK>(va[d], g.A, g.lda, ci.t.m0, g.M, ci.kb * kBK, g.K, ka.vec_a != 0, ptid);
          if constexpr (!BRES) tile_issue<BN, BK>(vb[d], g.B, g.ldb, ci.t.n0 * BN, g.N, ci.kb * kBK, g.K, ka.vec_b != 0, ptid);
          advance(ci);
        }
      }
    }
  } else if (warp == kMmaWarpG) {
    // ===================================== MMA issuer =====================================
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    constexpr uint32_t idesc = idesc_bf16_m128(BN);
    uint32_t it = 0, tcount = 0;
    if constexpr (BRES) wait_bar(bar_bres, 0);
    for (int idx = blockIdx.x; idx < ka.n_tiles; idx += gridDim.x, ++tcount) {
      const Tile t = tile_of(idx, ka.nt, ka.mt, ka.kb_total, ka.kb_per);
      const uint32_t acc = tcount & 1u, apar = (tcount >> 1) & 1u;
      wait_bar(bar_acce(acc), apar ^ 1u);
      tc_fence_after();
      const uint32_t d_addr = tm + acc * BN;
      for (int kb = t.kb0; kb < t.kb1; ++kb, ++it) {
        const uint32_t s = it % C::stages, par = (it / C::stages) & 1u;
        wait_bar(bar_full(s), par);
        fence_async_smem();                 // generic-proxy operand stores (acquired above) -> async proxy (tcgen05.mma)
        tc_fence_after();
        const uint32_t st = sbase + C::sm_stage0 + s * C::stage_bytes;
        const uint64_t a_hi = smem_desc_sw128(st);
        const uint64_t b_hi = smem_desc_sw128(BRES ? sbase + (uint32_t)kb * C::images * C::b_bytes : st + C::images * C::a_bytes);
        if (elect_one_g()) {
#pragma unroll
          for (uint32_t j = 0; j < 4; ++j) {
            // image pairs (a, b) with a + b < IMG: every cross term down to 2^-8 IMG of the product
#pragma unroll
            for (uint32_t ia = 0; ia < (uint32_t)IMG; ++ia)
#pragma unroll
              for (uint32_t ib = 0; ia + ib < (uint32_t)IMG; ++ib)
                mma_bf16(d_addr, a_hi + ia * (C::a_bytes >> 4) + 2u * j, b_hi + ib * (C::b_bytes >> 4) + 2u * j, idesc,
                         (kb > t.kb0 || j > 0 || ia > 0 || ib > 0) ? 1u : 0u);
          }
          mma_commit(bar_empty(s));
          if (kb + 1 == t.kb1) mma_commit(bar_accf(acc));
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================== epilogue =====================================
    const uint32_t stage_s = sbase + C::sm_epi + (uint32_t)warp * kStageBufBytes;
    uint32_t tcount = 0;
    for (int idx = blockIdx.x; idx < ka.n_tiles; idx += gridDim.x, ++tcount) {
      const Tile t = tile_of(idx, ka.nt, ka.mt, ka.kb_total, ka.kb_per);
      const uint32_t acc = tcount & 1u, apar = (tcount >> 1) & 1u;
      wait_bar(bar_accf(acc), apar);
      tc_fence_after();
      const uint32_t t_addr = tmem + ((uint32_t)(warp * 32) << 16) + acc * BN;
      const int n_base = t.n0 * BN;
      float* outp = ka.partial ? ka.partial + (size_t)t.ks * g.M * g.N : g.C;
      const int ldo = ka.partial ? g.N : g.ldc;
      const bool plain = ka.partial != nullptr;
      // this thread's part of the coalesced phase: row (lane >> 3) + 4 i of the warp's 32, columns 4 (lane & 7) .. + 3 of
      // every 32-column chunk; everything that does not depend on i or c is computed once per tile
      const int r0 = lane >> 3, cq = (lane & 7) * 4;
      const int m_first = t.m0 + warp * 32 + r0;
      const bool use_bias = !plain && g.bias != nullptr, use_mul = !plain && g.mul != nullptr, use_res = !plain && g.resid != nullptr,
                 use_acc = !plain && g.accumulate != 0, use_gelu = !plain && g.gelu_out != nullptr;
      const uint32_t rd_s = stage_s + (uint32_t)r0 * kStagePitch + (uint32_t)cq * 4u;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        if (n_base + c * 32 >= g.N) break;                      // warp-uniform: the rest of the tile is padding
        float v[32];
        tmem_ld32(t_addr + c * 32, v);
        tmem_wait_ld();
        if (c == BN / 32 - 1 || n_base + (c + 1) * 32 >= g.N) {  // last read of this accumulator: hand it back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acce(acc));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(stage_s + (uint32_t)lane * kStagePitch + j * 16, __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                 __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
        __syncwarp();
        // coalesced phase: 8 lanes cover the 128 bytes of a row segment, 4 rows per instruction
        const int n = n_base + c * 32 + cq;
        float4 x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[i].x), "=f"(x[i].y), "=f"(x[i].z), "=f"(x[i].w)
                       : "r"(rd_s + (uint32_t)i * 4u * kStagePitch));
        __syncwarp();                                            // the staging buffer may be overwritten by the next chunk
        if (n < g.N) {
          if (ka.vec_c && n + 4 <= g.N) {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (use_bias) b4 = __ldg(reinterpret_cast<const float4*>(g.bias + n));
            float* cp = outp + (size_t)m_first * ldo + n;
            const float* mp = use_mul ? g.mul + (size_t)m_first * g.ldm + n : nullptr;
            const float* rp = use_res ? g.resid + (size_t)m_first * g.ldr + n : nullptr;
            float* gp = use_gelu ? g.gelu_out + (size_t)m_first * g.ldg + n : nullptr;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (m_first + 4 * i < g.M) {
                float4 y = make_float4(x[i].x + b4.x, x[i].y + b4.y, x[i].z + b4.z, x[i].w + b4.w);
                if (use_mul) { const float4 q = __ldg(reinterpret_cast<const float4*>(mp + (size_t)(4 * i) * g.ldm)); y.x *= q.x; y.y *= q.y; y.z *= q.z; y.w *= q.w; }
                if (use_res) { const float4 q = __ldg(reinterpret_cast<const float4*>(rp + (size_t)(4 * i) * g.ldr)); y.x += q.x; y.y += q.y; y.z += q.z; y.w += q.w; }
                float4* o = reinterpret_cast<float4*>(cp + (size_t)(4 * i) * ldo);
                if (use_acc) { const float4 q = *o; y.x += q.x; y.y += q.y; y.z += q.z; y.w += q.w; }
                *o = y;
                if (use_gelu)
                  *reinterpret_cast<float4*>(gp + (size_t)(4 * i) * g.ldg) = make_float4(gelu_erf_g(y.x), gelu_erf_g(y.y), gelu_erf_g(y.z), gelu_erf_g(y.w));
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {                        // static indices: x stays in registers
              const int m = m_first + 4 * i;
              if (m >= g.M) continue;
              const float xv[4] = {x[i].x, x[i].y, x[i].z, x[i].w};
              float* cp = outp + (size_t)m * ldo + n;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (n + e >= g.N) break;
                float y = xv[e];
                if (use_bias) y += __ldg(g.bias + n + e);
                if (use_mul) y *= __ldg(g.mul + (size_t)m * g.ldm + n + e);
                if (use_res) y += __ldg(g.resid + (size_t)m * g.ldr + n + e);
                if (use_acc) y += cp[e];
                cp[e] = y;
                if (use_gelu) g.gelu_out[(size_t)m * g.ldg + n + e] = gelu_erf_g(y);
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarpG) tmem_dealloc(tmem, 512);
}

// C = sum_ks partial[ks] (+ bias + resid | + C), fixed order
__global__ void splitk_reduce_kernel(GemmArgs g, const float* __restrict__ partial, int ksplit) {
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t total = (size_t)g.M * g.N;
  if (idx >= total) return;
  const int m = (int)(idx / g.N), n = (int)(idx % g.N);
  float a = 0.f;
  for (int ks = 0; ks < ksplit; ++ks) a += partial[(size_t)ks * total + idx];
  if (g.bias) a += g.bias[n];
  if (g.mul) a *= g.mul[(size_t)m * g.ldm + n];
  if (g.resid) a += g.resid[(size_t)m * g.ldr + n];
  float* cp = g.C + (size_t)m * g.ldc + n;
  if (g.accumulate) a += *cp;
  *cp = a;
}

template <int BN, int IMG, bool BRES>
int launch(const KArgs& ka, int grid, cudaStream_t st) {
  using C = Cfg<BN, IMG, BRES>;
  const int smem = (int)C::smem + 1024;
#define BESO_GEMM_CASE(AK, BK)                                                                                      \
  do {                                                                                                              \
    static bool cfgd = false;                                                                                       \
    if (!cfgd) {                                                                                                    \
      BESO_CUDA(cudaFuncSetAttribute(gemm_kernel<BN, IMG, AK, BK, BRES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
      cfgd = true;                                                                                                  \
    }                                                                                                               \
    gemm_kernel<BN, IMG, AK, BK, BRES><<<grid, kThreadsG, smem, st>>>(ka);                                          \
  } while (0)
  if (ka.g.a_kmajor && ka.g.b_kmajor) BESO_GEMM_CASE(true, true);
  else if (ka.g.a_kmajor) BESO_GEMM_CASE(true, false);
  else if (ka.g.b_kmajor) BESO_GEMM_CASE(false, true);
  else BESO_GEMM_CASE(false, false);
#undef BESO_GEMM_CASE
  ++g_kernel_launches;
  BESO_CUDA(cudaGetLastError());
  return BESO_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

void gemm_ws_free(GemmWs& ws) {
  if (ws.partial) cudaFree(ws.partial);
  ws.partial = nullptr; ws.floats = 0;
}

int gemm_run(const GemmArgs& a, GemmWs& ws, int sm_count, cudaStream_t st) {
  if (a.M < 1 || a.N < 1 || a.K < 1 || !a.A || !a.B || !a.C) { set_error("gemm: bad arguments"); return BESO_E_INVALID; }
  KArgs ka{};
  ka.g = a;
  if (a.prec < 0 || a.prec > 2) { set_error("gemm: prec must be 0, 1 or 2"); return BESO_E_INVALID; }
  // resident-B mode: K <= 256, enough row tiles to keep the chip busy, one or two images (see Cfg); its column tile
  // is 256 wide with one image and 128 with two so that the resident operand stays within 128 KB
  const bool bres = a.K <= (int)(kResKb * kBK) && a.prec < 2 && (a.M + kBM - 1) / kBM >= 2 * sm_count;
  // three images per operand fill shared memory twice as fast: 128-wide tiles keep two pipeline stages
  const int bn = (a.N > 128 && a.prec < 2 && !(bres && a.prec == 1)) ? 256 : 128;
  ka.mt = (a.M + kBM - 1) / kBM;
  ka.nt = (a.N + bn - 1) / bn;
  ka.kb_total = (a.K + kBK - 1) / kBK;
  // split-K when the output has far fewer tiles than the chip has SMs and the contraction is long
  const int out_tiles = ka.mt * ka.nt;
  int ksplit = 1;
  if (!bres && !a.gelu_out && out_tiles * 2 <= sm_count && ka.kb_total >= 16) {
    ksplit = sm_count / out_tiles;
    const int max_split = ka.kb_total / 4;                    // at least 4 k-blocks per split
    if (ksplit > max_split) ksplit = max_split;
    if (ksplit < 1) ksplit = 1;
  }
  ka.kb_per = (ka.kb_total + ksplit - 1) / ksplit;
  ksplit = (ka.kb_total + ka.kb_per - 1) / ka.kb_per;         // no empty splits
  ka.ksplit = ksplit;
  ka.n_tiles = out_tiles * ksplit;
  ka.vec_a = a.a_kmajor ? (a.lda % 4 == 0 && aligned16(a.A)) : 0;
  ka.vec_b = a.b_kmajor ? (a.ldb % 4 == 0 && aligned16(a.B)) : 0;
  ka.partial = nullptr;
  if (ksplit > 1) {
    const size_t need = (size_t)ksplit * a.M * a.N;
    if (need > ws.floats) {
      if (ws.partial) cudaFree(ws.partial);
      ws.partial = nullptr; ws.floats = 0;
      BESO_CUDA(cudaMalloc(&ws.partial, need * sizeof(float)));
      ws.floats = need;
    }
    ka.partial = ws.partial;
    ka.vec_c = (a.N % 4 == 0);
  } else {
    ka.vec_c = a.ldc % 4 == 0 && aligned16(a.C) && (!a.resid || (a.ldr % 4 == 0 && aligned16(a.resid))) &&
               (!a.mul || (a.ldm % 4 == 0 && aligned16(a.mul))) &&
               (!a.bias || aligned16(a.bias)) && (!a.gelu_out || (a.ldg % 4 == 0 && aligned16(a.gelu_out)));
  }
  int grid = ka.n_tiles < sm_count ? ka.n_tiles : sm_count;
  int rc;
  if (bres) {
    grid = (sm_count / ka.nt) * ka.nt;                         // every CTA keeps one column tile: grid % nt == 0
    if (grid < ka.nt) grid = ka.nt;
    rc = a.prec ? launch<128, 2, true>(ka, grid, st) : (bn == 256 ? launch<256, 1, true>(ka, grid, st) : launch<128, 1, true>(ka, grid, st));
  } else if (bn == 256) rc = a.prec ? launch<256, 2, false>(ka, grid, st) : launch<256, 1, false>(ka, grid, st);
  else rc = a.prec == 2 ? launch<128, 3, false>(ka, grid, st) : (a.prec ? launch<128, 2, false>(ka, grid, st) : launch<128, 1, false>(ka, grid, st));
  if (rc) return rc;
  if (ksplit > 1) {
    const size_t total = (size_t)a.M * a.N;
    splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a, ws.partial, ksplit);
    ++g_kernel_launches;
    BESO_CUDA(cudaGetLastError());
  }
  return BESO_OK;
}

}  // namespace beso
