// FAST mode: GCDenoiser -> DiffusionGPT forward and the sample loop (DDIM / Euler / Heun / Euler-ancestral /
// DPM-Solver++(2M) / two-stage coefficient programs) as ONE persistent, warp-specialised sm_100a kernel.  fp16 operands on tcgen05 tensor cores (bf16 for the
// embedding GEMM), fp32 accumulation in TMEM, fp32 LayerNorm statistics / softmax / residual /
// pre-conditioning, packed-fp16 GELU and LayerNorm scaling.
//
// Reference semantics (beso/agents/diffusion_agents/k_diffusion/): score_wrappers.py:31-43,81-96;
// score_gpts.py:50-80,96-115,272-358; gc_sampling.py:167-413,703-736,895-1016;
// classifier_free_sampler.py:35-49.
//
// Design (DESIGN.md has the long form)
//  * A CTA owns a 128-row tile = floor(128 / T) whole sequences (T tokens each) for the whole
//    launch: all layers and all sampler steps.  Row r of the tile is TMEM lane r.
//  * The fp32 residual stream X (128 x 256) lives in TMEM columns [0,256) and never leaves it:
//    the attention-projection and MLP-down GEMMs accumulate straight into X (residual add for
//    free); their biases are added once, up front, by the embedding GEMM and the LayerNorm passes subtract
//    the ones that are not due yet.
//  * TMEM columns [256,512) are two 128-column scratch accumulators (per-head QKV, FC1 chunks).
//  * Every GEMM B operand comes from one linear "weight tape" in HBM/L2, pre-swizzled into the
//    UMMA K-major SWIZZLE_128B image and ordered exactly as the MMA warp consumes it; a producer
//    warp streams it with cp.async.bulk (TMA engine) through a ring of two 32 KB stages.
//  * One elected thread issues every tcgen05.mma from a straight-line schedule (compile-time shapes,
//    operand offsets and barrier ids); the cta_group::2 variant still walks a group table in shared memory.
//  * 8 compute warps (2 per TMEM lane quadrant) do the LayerNorms, the QKV drain, causal
//    attention with mma.sync on fp16 Q/K/V staged in shared memory (with two helper warps), erf-GELU,
//    the action head read-out, the Karras pre-conditioning and the sampler update.
//  * Embeddings (state / goal / action / sigma / position / biases) are one K=128 GEMM: the A
//    operand carries the raw inputs plus one-hot token-position columns whose B rows hold
//    (bias + pos_emb) split into bf16 hi + lo, so X starts exact to ~2^-17.
//
// Variants of the one kernel template (fast_sample_kernel<CG, DBG, MC, PREC, HSP, GEO>):
//  * GEO 0 (G256): embed_dim <= 256, everything above.  GEO 1 (G384): embed_dim <= 384 (the reference's kitchen models):
//    X takes 384 TMEM columns, ONE scratch accumulator ([Q|K] job + V job, FC1 without ping-pong), 16 KB ring slots,
//    the sampler's x buffers in a global scratch.  GEO 2 (G256P): the precise mode on full 128-row tiles.
//  * PREC: fp32-equivalent arithmetic from fp16 hi / lo operand images -- stacked on the MMA row dimension (64
//    sequence rows per tile; GEO 0 / 1) or as three MMAs per product with the lo image of the LayerNorm output read
//    as a tensor-memory A operand (GEO 2).  fast_launch picks the layout per batch.
//  * HSP: padded head size (32 / 64); CG / MC: CTA-pair MMAs / multicast weight ring (GEO 0, fp16 mode, opt-in).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <type_traits>
#include <vector>

#include "fast.cuh"
#include "umma.cuh"

namespace beso {

using namespace umma;

namespace {

// ---- geometry: embed_dim padded to 256 columns, head dims padded to 32 or 64 and packed 64 columns per
// attention pass (one 64-wide head, or two 32-wide heads), hidden width padded to 1024 ----------------
constexpr int kD = 256, kH = 4, kFF = 1024;
constexpr int kMaxPass = 6;               // attention passes per layer (n_heads * padded head size / 64)
constexpr int kRows = 128;
constexpr int kThreads = 384;             // warps 0-7 compute, 8-9 attention helpers, 10 producer, 11 MMA + TMEM alloc
// The warp scheduler prefers higher warp ids among eligible warps, so the two latency-critical
// single-warp roles get the highest ids of their sub-partitions.
constexpr int kProducerWarp = 10, kMmaWarp = 11, kHelperWarp0 = 8;
constexpr int kComputeWarp0 = 0, kComputeThreads = 256;
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
constexpr uint32_t kSmVecA = kSmU + 67584;          // 198656: [pend(256) | bq: 192 per pass, up to 6 passes | spare] fp32
constexpr uint32_t kVecBq = 256;                    // float index of the Q biases inside vecA
constexpr uint32_t kVecAFloats = 1536;
constexpr uint32_t kSmVecM = kSmVecA + kVecAFloats * 4;   // [pend(256) | b1 / 4 as fp16 (512 floats) | b1 fp32 (1024, PREC)]
constexpr uint32_t kVecB1H = 256, kVecB1F = 768;   // float indices of the two b1 images inside vecM
constexpr uint32_t kVecMFloats = 1792;
constexpr uint32_t kSmStats = kSmVecM + kVecMFloats * 4;  // [2][128] float2
constexpr uint32_t kSmProg = kSmStats + 2048;
constexpr int kGroupEntries = 2 + 52 + 4;        // device-side group table in shared memory (uint4 each)
constexpr uint32_t kXFloats = 832;
constexpr uint32_t kSmBars = 230400;                // 32 mbarriers + tmem pointer
constexpr uint32_t kSmemBytes = kSmBars + 512;

// barrier ids used by the fill program
enum { B_A_READY = 0, B_X_DONE, B_ACC_FULL0, B_ACC_FULL1, B_ACC_EMPTY0, B_ACC_EMPTY1, B_OP_READY0, B_OP_READY1,
       B_OP_EMPTY0, B_OP_EMPTY1, B_Y_READY, B_Y_EMPTY, B_OP_READY0B, B_OP_READY1B, B_FULL0, B_EMPTY0 = B_FULL0 + 4, B_PFULL0 = B_EMPTY0 + 4, B_COUNT = B_PFULL0 + 4 };
constexpr int kAttnWarps = 10;            // 8 compute warps + warps 2-3 help with attention
constexpr uint32_t kNone = 0xF;

// TMEM columns
constexpr uint32_t kColX = 0, kColS0 = 256, kColS1 = 384;

// ---- the two geometries of the kernel --------------------------------------------------------------------------
// G256: embed_dim <= 256 (everything above).  G384: embed_dim <= 384 (the reference's kitchen models: d = 360, 6 heads
// of 60; configs/franka_kitchen_main_config.yaml:38,58).  The wide geometry keeps the SAME roles, barriers and drains but
//  * X takes 384 of the 512 TMEM columns, so there is ONE 128-column scratch accumulator: an attention pass is a
//    [Q|K] job (N = 128) and a V job (N = 64), FC1 chunks do not ping-pong (FC2 of the previous chunk runs while the
//    current one is drained), every GEMM into X is three N = 128 column blocks;
//  * the A operand is six K atoms (96 KB), the ring is 16 KB slots (three; five during the MLP half, see RingW), the
//    hidden width is padded to 1536 (12 chunks), and the sampler's x / d buffers live in a global scratch.
struct G256 {
  static constexpr int DP = 256, NA = 4, FF = 1024, NCH = 8;
  static constexpr uint32_t ColS0 = 256, ColS1 = 384;
  static constexpr uint32_t SmRing = kSmRing, SmU = kSmU, SmQkv = kSmQkv, SmY = kSmY, SmH0 = kSmH0, SmH1 = kSmH1;
  static constexpr uint32_t SmVecA = kSmVecA, VecBq = kVecBq, BqStride = 192, VecAFloats = kVecAFloats;
  static constexpr uint32_t SmVecM = kSmVecM, VecB1H = kVecB1H, VecB1F = kVecB1F, VecMFloats = kVecMFloats;
  static constexpr uint32_t SmStats = kSmStats, SmProg = kSmProg, SmSigv = kSmProg + 1024 + 4 * kXFloats * 4;
  static constexpr bool XGlobal = false, HSingle = false;
  static constexpr uint32_t ColAL = 0, SmYLo = 0, SmHLo = 0;          // (G256P only)
  static constexpr uint32_t MlpSlots = 3, SmXSlot = 0;
};
struct G384 {
  static constexpr int DP = 384, NA = 6, FF = 1536, NCH = 12;
  static constexpr uint32_t ColS0 = 384, ColS1 = 384;
  static constexpr uint32_t SmRing = 98304, SmU = 147456;            // A: 6 x 16 KB | ring: 3 x 16 KB | union
  static constexpr uint32_t SmQkv = SmU, SmY = SmU + 51200, SmH0 = SmU, SmH1 = SmU + 32768;
  static constexpr uint32_t SmVecA = SmU + 67584, VecBq = 384, BqStride = 64, VecAFloats = 768;   // [pend(384) | bq: 64 per pass]
  // [pend(384) | b1: fp16 image / 4 (768 floats, fast) or fp32 image (1536 floats, PREC)]
  static constexpr uint32_t SmVecM = SmVecA + VecAFloats * 4, VecB1H = 384, VecB1F = 384, VecMFloats = 1920;
  static constexpr uint32_t SmStats = SmVecM + VecMFloats * 4, SmProg = SmStats + 2048, SmSigv = SmProg;
  static constexpr bool XGlobal = true, HSingle = false;
  static constexpr uint32_t ColAL = 0, SmYLo = 0, SmHLo = 0;          // (G256P only)
  static constexpr uint32_t MlpSlots = 3, SmXSlot = 0;                // no spare shared memory in the MLP half
};
// G256P: the precise mode for embed_dim <= 256 on FULL 128-row tiles ("P128").  Every product is three MMAs,
//     A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T          (fp16 images, fp32 accumulate; the lo.lo term is below 2^-22)
// instead of the stacked layout's two MMAs per 64 sequence rows: 25 % fewer tensor-pipe cycles per row and, because the
// weights now stream once per 128 rows instead of once per 64, half the weight traffic per row.  What makes it fit:
//  * the lo image of the LayerNorm output (the A operand of QKV and FC1, K = 256) lives in TENSOR memory -- 128 columns
//    of packed fp16 pairs, read by tcgen05.mma with a TMEM A operand -- so TMEM is X 256 | one scratch accumulator 128 |
//    A_lo 128, and the schedule is the single-accumulator one of G384;
//  * the lo images of the small A operands (embedding input, attention output Y, hidden chunk H) are shared-memory atoms;
//    H is single-buffered (hi + lo = 64 KB), the ring is three 16 KB slots;
//  * attention runs per 64-row half of the tile through the 64-row hi / lo staging buffer of the stacked mode: the warps
//    that own rows 64..127 keep their Q | K | V in registers while the first half is processed.  Sequences never straddle
//    row 64: a tile holds 2 x floor(64 / T) sequences.
struct G256P {
  static constexpr int DP = 256, NA = 4, FF = 1024, NCH = 8;
  static constexpr uint32_t ColS0 = 256, ColS1 = 256, ColAL = 384;
  static constexpr uint32_t SmRing = 65536, SmU = 114688;             // A_hi: 4 x 16 KB | ring: 3 x 16 KB | union
  static constexpr uint32_t SmQkv = SmU, SmY = SmU + 51200, SmYLo = SmY + 16384;
  static constexpr uint32_t SmH0 = SmU, SmH1 = SmU, SmHLo = SmU + 32768;
  static constexpr uint32_t SmVecA = SmU + 83968, VecBq = kVecBq, BqStride = 192, VecAFloats = kVecAFloats;
  static constexpr uint32_t SmVecM = SmVecA + VecAFloats * 4, VecB1H = kVecB1H, VecB1F = kVecB1F, VecMFloats = kVecMFloats;
  static constexpr uint32_t SmStats = SmVecM + VecMFloats * 4, SmProg = SmStats + 2048, SmSigv = SmProg + 1024 + 4 * kXFloats * 4;
  static constexpr bool XGlobal = false, HSingle = true;
  // The MLP half needs 64 KB of the 82 KB union region (H hi + lo): a fourth ring slot lives behind H while it runs --
  // with three slots the weight stream of the MLP half is latency-bound (two 16 KB requests in flight).  The bytes belong
  // to the Y atoms during the attention half: the producer takes the slot only after that half's last projection.
  static constexpr uint32_t MlpSlots = 4, SmXSlot = SmU + 65536;
};
static_assert(G256P::SmXSlot + 16384 <= G256P::SmVecA, "extra ring slot must fit behind H");
static_assert(G256P::SmVecA == kSmVecA && G256P::SmSigv + 512 <= kSmBars, "P128 geometry: vectors and x buffers sit where G256 has them");
static_assert(G384::SmSigv + 512 <= kSmBars, "wide geometry does not fit shared memory");
// G384 weight ring: 16 KB slots, one ring group per slot; full barriers B_FULL0 + s, empty barriers B_WEMPTY0 + s (s < 6).
// Producer and MMA issuer replay the same slot sequence; n_slots may differ between phases of the schedule.
constexpr uint32_t kWSlots = 3;
constexpr int B_WEMPTY0 = B_FULL0 + 6;
struct RingW {
  uint32_t cur, par;                 // next slot; bit s of par = parity of slot s's next use
  __device__ __forceinline__ uint32_t begin(uint32_t n_slots) { if (cur >= n_slots) cur = 0; return cur; }
  __device__ __forceinline__ uint32_t parity(uint32_t s) const { return (par >> s) & 1u; }
  __device__ __forceinline__ void end(uint32_t s) { par ^= 1u << s; cur = s + 1; }
  // slots 0 .. kWSlots - 1: the ring proper; further slots (MLP half only, G::MlpSlots) live in the part of the union
  // region that the MLP half leaves free (G::SmXSlot)
  template <class G> static __device__ __forceinline__ uint32_t slot_off(uint32_t s) {
    return s < kWSlots ? G::SmRing + s * 16384u : G::SmXSlot + (s - kWSlots) * 16384u;
  }
};

struct FastParams {
  const uint8_t* tape;         // per-eval weight tape
  const float* vec;            // per layer: vecA (1536) | vecM (1792); then final vecA (1536)
  int n_fills, L, G, obs, act, T, t, S, n_tiles, B, evals;
  int npass, hsp, d_true;      // attention passes per layer, padded head size (32 / 64), true embed_dim
  uint32_t flags;
  float lambda, sigma_data, inv_d;
  const float *state, *goal, *xin, *sigma;
  float* out;
  float* xscratch;             // G384: per-CTA x / d1 / x2 / dU buffers (4 x kXFloats floats each)
  float* trace;                // optional debug dump of X after every LayerNorm pass (tile 0, eval 0)
  long long* timeline;         // optional clock64 stamps of block 0, second evaluation (see tools/timeline_fast.py)
};

// ================================ device code =====================================================
// One ring group = the B operand of one k-block of one GEMM job (a fill or two adjacent fills) plus the
// four K=16 MMAs that consume it.  walk_eval() enumerates the groups of one model evaluation in the
// order the tape stores them; producer, MMA issuer and the pair forwarder all replay it.
struct Group {
  uint32_t a_off, d_col, n, acc;     // A atom (smem offset), TMEM column of D, N, accumulate on first k-step
  uint32_t w0, w1, c0, c1;           // barrier ids to wait on before / commit to after (kNone = none)
  bool pair;                         // two fills (two adjacent 16 KB ring slots when CG = 1)
  bool kk2;                          // two K blocks (8 MMAs): A atoms a_off, a_off + 16 KB; also two fills
  bool bf16;                         // operands are bf16 (embedding GEMM), else fp16
};
// Job types of the per-evaluation schedule, decoded arithmetically (no tables on the issue path).
enum { J_EMB = 0, J_QKV, J_PROJ, J_FC1, J_FC2, J_HEAD };
__device__ __forceinline__ uint32_t job_groups(uint32_t type) {
  return type == J_QKV || type == J_HEAD ? 4u : (type == J_PROJ ? 1u : 2u);
}
// layer-local job number (0..23) -> (type, index): Q0 Q1 P0 Q2 P1 Q3 P2 P3 | F1_0 F1_1 F2_0 (F1_c F2_c-1)c=2..7 F2_7
__device__ __forceinline__ void layer_job(uint32_t lj, uint32_t& type, uint32_t& idx) {
  if (lj < 8) {
    const uint32_t code = (0x76352410u >> (4 * lj)) & 0xFu;     // nibble = (proj ? 4 : 0) | head
    type = (code & 4u) ? J_PROJ : J_QKV;
    idx = code & 3u;
  } else {
    const uint32_t m = lj - 8;
    if (m < 3) { type = m == 2 ? J_FC2 : J_FC1; idx = m == 1 ? 1u : 0u; }
    else if (m == 15) { type = J_FC2; idx = 7; }
    else { const uint32_t t = m - 3, c = 2 + (t >> 1); type = (t & 1) ? J_FC2 : J_FC1; idx = (t & 1) ? c - 1 : c; }
  }
}
__device__ __forceinline__ Group make_group(uint32_t type, uint32_t idx, uint32_t kb) {
  Group q;
  q.w0 = q.w1 = q.c0 = q.c1 = kNone;
  q.pair = true; q.kk2 = false; q.acc = kb > 0; q.bf16 = type == J_EMB;
  q.a_off = kSmA + kb * 16384;
  if (type == J_EMB) {                  // X = A_emb W_emb^T
    q.d_col = kColX; q.n = 256;
    if (kb == 0) q.w0 = B_A_READY;
    if (kb == 1) q.c0 = B_X_DONE;
  } else if (type == J_QKV) {           // head idx: [Q|K|V] (192 columns) of this head
    q.d_col = kColS0; q.n = 192;
    if (kb == 0) { q.w0 = B_ACC_EMPTY0; if (idx == 0) q.w1 = B_A_READY; }
    if (kb == 3) q.c0 = B_ACC_FULL0;
  } else if (type == J_PROJ) {          // X += Y_h Wproj[:, h]^T
    q.a_off = kSmY; q.d_col = kColX; q.n = 256; q.acc = 1;
    q.w0 = B_Y_READY; q.c0 = B_Y_EMPTY;
    if (idx == 3) q.c1 = B_X_DONE;
  } else if (type == J_FC1) {           // hidden chunk idx (128 wide), two K blocks per group
    const uint32_t b = idx & 1;
    q.a_off = kSmA + kb * 32768; q.d_col = b ? kColS1 : kColS0; q.n = 128; q.pair = false; q.kk2 = true;
    if (kb == 0) { q.w0 = B_ACC_EMPTY0 + b; if (idx == 0) q.w1 = B_A_READY; }
    if (kb == 1) q.c0 = B_ACC_FULL0 + b;
  } else if (type == J_FC2) {           // X += H_idx W2[:, chunk]^T
    const uint32_t b = idx & 1;
    q.a_off = (b ? kSmH1 : kSmH0) + kb * 16384; q.d_col = kColX; q.n = 256; q.acc = 1;
    q.w0 = (kb == 0 ? B_OP_READY0 : B_OP_READY0B) + b;     // one barrier per K atom of H (a parity wait must
                                                           // never fall two phases behind)
    if (kb == 1) { q.c0 = B_OP_EMPTY0 + b; if (idx == 7) q.c1 = B_X_DONE; }
  } else {                              // action head, N = 16
    q.d_col = kColS0; q.n = 16; q.pair = false;
    if (kb == 0) { q.w0 = B_A_READY; q.w1 = B_ACC_EMPTY0; }
    if (kb == 3) q.c0 = B_ACC_FULL0;
  }
  return q;
}
// Calls f(group) for every ring group of one evaluation, from ONE call site (compact code for the
// single-warp roles, everything stays in the uniform datapath).
template <class F>
__device__ __forceinline__ void walk_eval(int L, F&& f) {
  const uint32_t n_jobs = 2u + 24u * (uint32_t)L;
  uint32_t lj = 0;
#pragma unroll 1
  for (uint32_t j = 0; j < n_jobs; ++j) {
    uint32_t type, idx = 0;
    if (j == 0) type = J_EMB;
    else if (j == n_jobs - 1) type = J_HEAD;
    else { layer_job(lj, type, idx); lj = lj == 23 ? 0 : lj + 1; }
    const uint32_t nk = job_groups(type);
#pragma unroll 1
    for (uint32_t kb = 0; kb < nk; ++kb) f(make_group(type, idx, kb));
  }
}

__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// compute + helper warps.  Out of line on purpose: the two roles then execute the SAME bar instruction (one program
// counter), which is what compute-sanitizer's synccheck expects of the participants of a barrier -- inlined, the tool
// reports the legal "same named barrier from two call sites" pattern as divergence.
__device__ __noinline__ void attn_sync() { asm volatile("bar.sync 2, 320;" ::: "memory"); }
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
// (try_wait itself suspends the warp for a hardware-defined interval, the clock is read every 64 polls.)
__device__ __noinline__ void wait_timeout(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  uint32_t polls = 0;
  while (!mbar_try_wait_sleep(bar, parity)) {
    if ((++polls & 63u) == 0 && clock64() - t0 > 8000000000ll) {
      printf("beso fast kernel: mbarrier id %u parity %u timed out (block %d thread %d)\n",
             (bar & 0x3FF) / 8, parity, blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __noinline__ void wait_timeout_cluster(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 8000000000ll) {
      printf("beso fast kernel: (cluster) mbarrier id %u parity %u timed out (block %d thread %d)\n",
             (bar & 0x3FF) / 8, parity, blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void spin_wait_cluster(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;     // keep the fast path tiny: the issue loops live in the I-cache
  wait_timeout_cluster(bar, parity);
}
__device__ __forceinline__ void spin_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  wait_timeout(bar, parity);
}

// erf-GELU without erff() and without the special-function unit, two elements per instruction in packed
// fp16.  The FC1 weights and bias are packed with a factor 1/4 and the FC2 weights with a factor 4, so the
// drain sees xs = x / 4 and produces gelu(x) / 4.  With u = min(|xs|, 1.375) and T(a) = 0.5 erfc(a / sqrt 2):
//   gelu(x) / 4 = max(xs, 0) - |xs| T(4 u),        T(4 u) = p(u)^8,   p a degree-5 polynomial
// p is fitted to T^(1/8) (a slowly varying, Gaussian-like function); the 8th power is three squarings.
// Evaluated in fp16 the result is within 1.4x of the rounding floor of the exact GELU rounded to fp16 (mean
// |error| 9.3e-5 vs 6.5e-5 over [-10, 10], in units of the unscaled GELU).
// Why this form: on B200 HFMA2 / HMUL2 issue at one warp instruction per two cycles per scheduler and
// MUFU.EX2 at one per eight, and ex2.approx.f16x2 is two MUFUs -- the exponential of the previous
// formulation (max(x,0) - |x| 2^q(|x|)) cost as much as its degree-5 polynomial and serialised behind it
// (profiles/r1_probe_pipes.txt, tools/probe_gelu.cu).  13 packed instructions per pair, no MUFU.
__device__ __forceinline__ __half2 h2const(float v) { return __float2half2_rn(v); }
constexpr float kGeluInScale = 0.25f, kGeluOutScale = 4.0f;
template <int NP>
__device__ __forceinline__ void gelu2_vec(__half2 (&x)[NP]) {
  __half2 u[NP], p[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) u[i] = __hmin2(__habs2(x[i]), h2const(1.375f));
#pragma unroll
  for (int i = 0; i < NP; ++i) p[i] = __hfma2(h2const(-1.97688124e-01f), u[i], h2const(5.00786336e-01f));
#pragma unroll
  for (int i = 0; i < NP; ++i) p[i] = __hfma2(p[i], u[i], h2const(-7.39247727e-02f));
#pragma unroll
  for (int i = 0; i < NP; ++i) p[i] = __hfma2(p[i], u[i], h2const(-5.06604987e-01f));
#pragma unroll
  for (int i = 0; i < NP; ++i) p[i] = __hfma2(p[i], u[i], h2const(-3.66090489e-01f));
#pragma unroll
  for (int i = 0; i < NP; ++i) p[i] = __hfma2(p[i], u[i], h2const(9.17009012e-01f));
#pragma unroll
  for (int i = 0; i < NP; ++i) p[i] = __hmul2(p[i], p[i]);
#pragma unroll
  for (int i = 0; i < NP; ++i) p[i] = __hmul2(p[i], p[i]);
#pragma unroll
  for (int i = 0; i < NP; ++i) p[i] = __hmul2(p[i], p[i]);
#pragma unroll
  for (int i = 0; i < NP; ++i) x[i] = __hfma2(__hneg2(__habs2(x[i])), p[i], __hmax2(x[i], h2const(0.f)));
}
__device__ __forceinline__ uint32_t h2bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 bits2h(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
// (a, b) -> packed fp16 hi pair and the packed fp16 rounding of the remainders
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  hi = h2bits(h);
  lo = h2bits(__floats2half2_rn(a - hf.x, b - hf.y));
}

__device__ __forceinline__ float ex2f(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
  return e;
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
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct Compute {
  uint8_t* sm;
  uint32_t sbase, tmem;
  int wq, lane, hf, row, ctid;       // TMEM lane quadrant, lane, column half, tile row (= TMEM lane), compute thread id
  int srow, is_lo;                   // sequence row this thread works on; PREC: 1 = this TMEM lane holds the "lo" image
  uint32_t row_off, rx4;             // row * 128 and (row & 7) << 4: SW128 address of chunk k = row_off + ((k << 4) ^ rx4)
  uint32_t phases;                   // parity bit per barrier id this role waits on
  __device__ uint32_t bar(int id) const { return sbase + kSmBars + id * 8; }
  __device__ void wait(int id) { spin_wait(bar(id), (phases >> id) & 1u); phases ^= 1u << id; }
  __device__ void wait2(int a, int b) {          // two barriers, tests issued back to back
    const uint32_t pa = (phases >> a) & 1u, pb = (phases >> b) & 1u;
    const bool ra = mbar_try_wait(bar(a), pa), rb = mbar_try_wait(bar(b), pb);
    if (!ra) wait_timeout(bar(a), pa);
    if (!rb) wait_timeout(bar(b), pb);
    phases ^= (1u << a) | (1u << b);
  }
  int cg;                            // CTAs per MMA group (1 or 2)
  // compute -> MMA barriers live in the leader CTA (rank 0) of the pair
  __device__ void arrive(int id) const {
    __syncwarp();
    if (lane == 0) { if (cg == 2) mbar_arrive_cluster(bar(id), 0); else mbar_arrive(bar(id)); }
  }
  __device__ uint32_t lane_addr(uint32_t col) const { return tmem + ((uint32_t)(wq * 32) << 16) + col; }
  // shared address of 16-byte chunk k of this thread's row in the SW128 atom at shared address `atom`
  __device__ uint32_t chunk_addr(uint32_t atom, uint32_t k) const { return atom + row_off + ((k << 4) ^ rx4); }
};

// A <- fp16(LayerNorm0(X + pend)), LayerNorm0 = (x - mean) * rstd.   vec_s = shared address of pend (fp32).
// The LayerNorm weight / bias are folded into the Linear that consumes A (see fast_pack).
// X (TMEM) is only read: every projection / MLP bias is added to X up front by the embedding GEMM and
// `pend` holds minus the biases that are not due yet at this point of the network.
template <bool DBG>
__device__ __noinline__ void ln_pass(const Compute c, uint32_t vec_s, float inv_d, float* trace_row) {
  // X is read from TMEM once: the 128 values of this thread (pending biases added) wait for the row
  // statistics as 64 packed fp16 pairs in registers; the statistics themselves are fp32 of the unrounded
  // values.  The normalisation is one packed HFMA2 per pair, straight into the fp16 A operand.
  float va[32], vb[32];
  __half2 keep[64];
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
  const int col0 = c.hf * 128;
  auto pass1 = [&](float (&v)[32], int ch) {
    const int col = col0 + ch * 32;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 pd = lds128f_ro(vec_s + (uint32_t)(col + i) * 4u);
      const float a0 = v[i] + pd.x, a1 = v[i + 1] + pd.y, a2 = v[i + 2] + pd.z, a3 = v[i + 3] + pd.w;
      s0 += a0; s1 += a1; s2 += a2; s3 += a3;
      q0 = fmaf(a0, a0, q0); q1 = fmaf(a1, a1, q1); q2 = fmaf(a2, a2, q2); q3 = fmaf(a3, a3, q3);
      keep[ch * 16 + (i >> 1)] = __floats2half2_rn(a0, a1);
      keep[ch * 16 + (i >> 1) + 1] = __floats2half2_rn(a2, a3);
      if (DBG) {
        if (trace_row != nullptr) {
          float* tr = trace_row + col + i;
          tr[0] = a0; tr[1] = a1; tr[2] = a2; tr[3] = a3;
        }
      }
    }
  };
  // the next chunk's TMEM load is in flight while the current one is reduced
  tmem_ld32(c.lane_addr(kColX + col0), va);
  tmem_ld32(c.lane_addr(kColX + col0 + 32), vb);
  tmem_wait_ld();
  pass1(va, 0);
  tmem_ld32(c.lane_addr(kColX + col0 + 64), va);
  pass1(vb, 1);
  tmem_ld32(c.lane_addr(kColX + col0 + 96), vb);
  tmem_wait_ld();
  pass1(va, 2);
  pass1(vb, 3);
  const float sum = (s0 + s1) + (s2 + s3), sq = (q0 + q1) + (q2 + q3);
  float2* stats = reinterpret_cast<float2*>(c.sm + kSmStats);
  stats[c.hf * kRows + c.row] = make_float2(sum, sq);
  compute_sync();
  const float2 o = stats[(c.hf ^ 1) * kRows + c.row];
  // columns >= embed_dim of X are exactly zero (zero weight rows, zero pend), so the sums run over the true lanes
  const float mean = (sum + o.x) * inv_d;
  const float var = fmaxf((sq + o.y) * inv_d - mean * mean, 0.f);
  const float rstd = rsqrtf(var + 1e-5f);
  const __half2 r2 = __float2half2_rn(rstd), n2 = __float2half2_rn(-mean * rstd);
  const uint32_t a_s = c.sbase + kSmA + (uint32_t)c.hf * 32768u;       // this thread's two K atoms
#pragma unroll
  for (int ch8 = 0; ch8 < 16; ++ch8) {             // 16 chunks of 8 columns = 16 bytes of fp16 each
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = h2bits(__hfma2(keep[ch8 * 4 + k], r2, n2));
    sts128(c.chunk_addr(a_s + (uint32_t)(ch8 >> 3) * 16384u, ch8 & 7), w[0], w[1], w[2], w[3]);
  }
  fence_async_smem();
  tc_fence_before();
  c.arrive(B_A_READY);
  // no trailing barrier: every path to the next LayerNorm pass (which rewrites the stats buffer) goes through
  // another barrier of all compute warps (attention syncs, the end-of-MLP sync, the epilogue syncs)
}

// G384: 192 columns per thread.  Two sweeps over X in TMEM (statistics, then normalise in fp32 and round once) instead of
// parking the row in registers: 96 packed pairs plus the load buffers do not fit the 168-register budget, and a TMEM
// sweep costs a few hundred cycles.
template <bool DBG>
__device__ __noinline__ void ln_pass_w(const Compute c, uint32_t vec_s, float inv_d, float* trace_row) {
  float va[32], vb[32];
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
  const int col0 = c.hf * 192;
  auto pass1 = [&](float (&v)[32], int ch) {
    const int col = col0 + ch * 32;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 pd = lds128f_ro(vec_s + (uint32_t)(col + i) * 4u);
      const float a0 = v[i] + pd.x, a1 = v[i + 1] + pd.y, a2 = v[i + 2] + pd.z, a3 = v[i + 3] + pd.w;
      s0 += a0; s1 += a1; s2 += a2; s3 += a3;
      q0 = fmaf(a0, a0, q0); q1 = fmaf(a1, a1, q1); q2 = fmaf(a2, a2, q2); q3 = fmaf(a3, a3, q3);
      if (DBG) {
        if (trace_row != nullptr) {
          float* tr = trace_row + col + i;
          tr[0] = a0; tr[1] = a1; tr[2] = a2; tr[3] = a3;
        }
      }
    }
  };
  tmem_ld32(c.lane_addr(kColX + col0), va);
  tmem_ld32(c.lane_addr(kColX + col0 + 32), vb);
  tmem_wait_ld();
#pragma unroll
  for (int ch = 0; ch < 6; ch += 2) {
    pass1(va, ch);
    if (ch + 2 < 6) tmem_ld32(c.lane_addr(kColX + col0 + (ch + 2) * 32), va);
    pass1(vb, ch + 1);
    if (ch + 2 < 6) { tmem_ld32(c.lane_addr(kColX + col0 + (ch + 3) * 32), vb); tmem_wait_ld(); }
  }
  const float sum = (s0 + s1) + (s2 + s3), sq = (q0 + q1) + (q2 + q3);
  float2* stats = reinterpret_cast<float2*>(c.sm + G384::SmStats);
  stats[c.hf * kRows + c.row] = make_float2(sum, sq);
  // the second sweep's first loads fly while the statistics are exchanged
  tmem_ld32(c.lane_addr(kColX + col0), va);
  tmem_ld32(c.lane_addr(kColX + col0 + 32), vb);
  compute_sync();
  const float2 o = stats[(c.hf ^ 1) * kRows + c.row];
  const float mean = (sum + o.x) * inv_d;
  const float var = fmaxf((sq + o.y) * inv_d - mean * mean, 0.f);
  const float rstd = rsqrtf(var + 1e-5f), nm = -mean * rstd;
  auto pass2 = [&](float (&v)[32], int ch) {
    const int col = col0 + ch * 32;                      // 32 columns = chunks k0 .. k0 + 3 of one K atom
    const uint32_t atom = c.sbase + kSmA + (uint32_t)(col >> 6) * 16384u, k0 = (uint32_t)(col & 63) >> 3;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // (volatile loads: identical to the first sweep's, they must not be merged with them and kept live across it)
      const float4 p0 = lds128f_v(vec_s + (uint32_t)(col + k * 8) * 4u), p1 = lds128f_v(vec_s + (uint32_t)(col + k * 8 + 4) * 4u);
      const float* e = v + k * 8;
      sts128(c.chunk_addr(atom, k0 + k),
             pack_f16x2(fmaf(e[0] + p0.x, rstd, nm), fmaf(e[1] + p0.y, rstd, nm)), pack_f16x2(fmaf(e[2] + p0.z, rstd, nm), fmaf(e[3] + p0.w, rstd, nm)),
             pack_f16x2(fmaf(e[4] + p1.x, rstd, nm), fmaf(e[5] + p1.y, rstd, nm)), pack_f16x2(fmaf(e[6] + p1.z, rstd, nm), fmaf(e[7] + p1.w, rstd, nm)));
    }
  };
  tmem_wait_ld();
#pragma unroll
  for (int ch = 0; ch < 6; ch += 2) {
    pass2(va, ch);
    if (ch + 2 < 6) tmem_ld32(c.lane_addr(kColX + col0 + (ch + 2) * 32), va);
    pass2(vb, ch + 1);
    if (ch + 2 < 6) { tmem_ld32(c.lane_addr(kColX + col0 + (ch + 3) * 32), vb); tmem_wait_ld(); }
  }
  fence_async_smem();
  tc_fence_before();
  c.arrive(B_A_READY);
}

// Accumulator of head h (Q|K at S0, V at S1[0:64)) -> fp16 Q|K|V staging rows.  Only Q gets its bias here: the
// K bias shifts every score of a query row by the same amount (softmax-invariant) and the V bias passes
// through the softmax average unchanged, so it is folded into the projection bias at pack time.
// PARTS: bit 0 = Q, bit 1 = K, bit 2 = V are in the accumulator.  7: one [Q|K|V] job (G256).  G384 has a single 128-column
// scratch accumulator: 3 = the [Q|K] job, 4 = the V job (V then sits at the start of the accumulator).
template <class G, int PARTS>
__device__ __noinline__ void drain_qkv(const Compute c, uint32_t bq_s) {
  // Each thread takes 32 columns of Q, of K and of V of its row (so that both column halves carry the same
  // share of the Q bias).  All TMEM reads first, then the accumulator is handed back to the MMA warp (QKV of
  // the next head can start) while this thread still converts and stores.
  float v0[32], v1[32], v2[32];
  const int colb = c.hf * 32;                            // within each 64-column block of [Q_h | K_h | V_h]
  if constexpr ((PARTS & 1) != 0) tmem_ld32(c.lane_addr(G::ColS0 + colb), v0);
  if constexpr ((PARTS & 2) != 0) tmem_ld32(c.lane_addr(G::ColS0 + 64 + colb), v1);
  if constexpr ((PARTS & 4) != 0) tmem_ld32(c.lane_addr(G::ColS0 + (PARTS == 7 ? 128 : 0) + colb), v2);
  tmem_wait_ld();
  tc_fence_before();
  c.arrive(B_ACC_EMPTY0);
  const uint32_t dst0 = c.sbase + G::SmQkv + (uint32_t)c.row * kQkvStride + (uint32_t)colb * 2u;
  if constexpr ((PARTS & 1) != 0) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) {                      // Q (+ bias)
    const float4 b0 = lds128f_ro(bq_s + (uint32_t)(colb + i) * 4u);
    v0[i] += b0.x; v0[i + 1] += b0.y; v0[i + 2] += b0.z; v0[i + 3] += b0.w;
  }
  }
  auto emit = [&](const float (&v)[32], uint32_t dst) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      sts128(dst + q * 16, pack_f16x2(v[q * 8 + 0], v[q * 8 + 1]), pack_f16x2(v[q * 8 + 2], v[q * 8 + 3]),
             pack_f16x2(v[q * 8 + 4], v[q * 8 + 5]), pack_f16x2(v[q * 8 + 6], v[q * 8 + 7]));
  };
  if constexpr ((PARTS & 1) != 0) emit(v0, dst0);
  if constexpr ((PARTS & 2) != 0) emit(v1, dst0 + 128);
  if constexpr ((PARTS & 4) != 0) emit(v2, dst0 + 256);
}

// Causal softmax(Q K^T) V for every sequence of the tile, one warp per (sequence, 16-query tile), mma.sync fp16.
// Q is pre-scaled by log2(e) / sqrt(hs) (folded into the packed weights).  Output -> Y atom (SW128 A layout).
// NKT = number of 16-key steps this query tile sees (causal: later key tiles are fully masked); compile-time so
// that a 16-token sequence does not pay for the masked second key step.
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// HI = false: rows 8..15 of the diagonal tile lie beyond the sequence (e.g. tokens 24..31 of a 23-token sequence):
// as queries their softmax is skipped and their P rows are zero; as keys (columns 8..15 of the last key step)
// their scores are neither computed nor exponentiated.
// HSP = padded head size: 64 (one head per pass) or 32 (two heads per pass, `co` = column offset of this head inside
// the 64-column blocks of Q, K, V and Y).
// SPLIT (the precise mode): staging rows [0, 64) hold the fp16 hi image of Q|K|V and rows [64, 128) the lo image;
// every product is three mma.sync (hi.hi + lo.hi + hi.lo, fp32 accumulate), the probabilities are split the same way
// and the output goes to the hi / lo rows of the Y atom.
// NH = 2: two warps share one item -- both compute the scores and the softmax, each the P V product of half of the
// output columns (`half`); used when a tile has so few (sequence, query tile) items that warps would idle.
// LAY: 0 = single fp16 image; 1 = split, stacked operand rows (SPLIT above); 2 = split, one 64-row half of a 128-row
// tile (P128): Y row = yrow0 + staging row, the lo image of Y is a separate atom.
template <class G, int NKT, bool HI, int HSP, int LAY, int NH>
// ybar != 0: mbarrier (shared address) / parity to pass before Y is written -- "the previous pass's projection MMAs have
// read Y" -- for the schedules that let that projection run under this pass's attention (single-accumulator geometries).
__device__ __forceinline__ void attention_item(uint32_t sbase, int lane, int row0, int mt, int T, int co, int half, int yrow0,
                                               uint32_t ybar, uint32_t ypar) {
  constexpr bool SPLIT = LAY != 0;
  const uint32_t qkv = sbase + G::SmQkv + (uint32_t)co * 2u;
  constexpr int KS = HSP / 16;                   // 16-wide k steps over the head dimension
  constexpr int kLastRow = SPLIT ? 63 : kRows - 1;
  constexpr uint32_t kLo = 64u * kQkvStride;     // byte offset of the lo image
  // ---- S = Q K^T ----
  float sc[NKT][2][4];
#pragma unroll
  for (int a = 0; a < NKT; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) sc[a][b][0] = sc[a][b][1] = sc[a][b][2] = sc[a][b][3] = 0.f;
  uint32_t qa[KS][4], ql[SPLIT ? KS : 1][4];
  {
    const int r = min(row0 + mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, kLastRow);
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      ldmatrix_x4(qkv + r * kQkvStride + (k * 16 + (lane >> 4) * 8) * 2, qa[k]);
      if constexpr (SPLIT) ldmatrix_x4(qkv + kLo + r * kQkvStride + (k * 16 + (lane >> 4) * 8) * 2, ql[k]);
    }
  }
#pragma unroll
  for (int kt = 0; kt < NKT; ++kt) {
    const int r = min(row0 + kt * 16 + (lane & 7) + (lane >> 4) * 8, kLastRow);
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      uint32_t kb[4];
      ldmatrix_x4(qkv + r * kQkvStride + (64 + k * 16 + ((lane >> 3) & 1) * 8) * 2, kb);
      mma_16816(sc[kt][0], qa[k], kb[0], kb[1]);
      if (HI || kt + 1 < NKT) mma_16816(sc[kt][1], qa[k], kb[2], kb[3]);
      if constexpr (SPLIT) {
        mma_16816(sc[kt][0], ql[k], kb[0], kb[1]);                              // lo . hi
        if (HI || kt + 1 < NKT) mma_16816(sc[kt][1], ql[k], kb[2], kb[3]);
        ldmatrix_x4(qkv + kLo + r * kQkvStride + (64 + k * 16 + ((lane >> 3) & 1) * 8) * 2, kb);
        mma_16816(sc[kt][0], qa[k], kb[0], kb[1]);                              // hi . lo
        if (HI || kt + 1 < NKT) mma_16816(sc[kt][1], qa[k], kb[2], kb[3]);
      }
    }
  }
  // ---- mask + softmax (rows i0 = lane/4 and i0 + 8 of this query tile) ----
  const int i_lo = mt * 16 + (lane >> 2), i_hi = i_lo + 8;
  float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
  for (int kt = 0; kt < NKT; ++kt)
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (!HI && kt + 1 == NKT && nb == 1) continue;      // keys beyond the sequence
        const int j = kt * 16 + nb * 8 + (lane & 3) * 2 + e;
        if (j > i_lo) sc[kt][nb][e] = -INFINITY;
        mx_lo = fmaxf(mx_lo, sc[kt][nb][e]);
        if (HI) {
          if (j > i_hi) sc[kt][nb][2 + e] = -INFINITY;
          mx_hi = fmaxf(mx_hi, sc[kt][nb][2 + e]);
        }
      }
  mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1)); mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
  if (HI) { mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1)); mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2)); }
  // the scores are in log2 units: p = 2^(s - max), normalised before the P V product so that the output needs
  // no scaling
  float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
  for (int kt = 0; kt < NKT; ++kt)
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (!HI && kt + 1 == NKT && nb == 1) continue;
        sc[kt][nb][e] = ex2f(sc[kt][nb][e] - mx_lo);
        sum_lo += sc[kt][nb][e];
        if (HI) { sc[kt][nb][2 + e] = ex2f(sc[kt][nb][2 + e] - mx_hi); sum_hi += sc[kt][nb][2 + e]; }
      }
  sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 1); sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 2);
  if (HI) { sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 1); sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 2); }
  const float inv_lo = __frcp_rn(sum_lo), inv_hi = HI ? __frcp_rn(sum_hi) : 0.f;
  uint32_t pa[NKT][4], pl[SPLIT ? NKT : 1][4];   // P as A fragments, one per 16-key step (and the lo image)
#pragma unroll
  for (int kt = 0; kt < NKT; ++kt) {
    if constexpr (SPLIT) {
      split2(sc[kt][0][0] * inv_lo, sc[kt][0][1] * inv_lo, pa[kt][0], pl[kt][0]);
      split2(sc[kt][0][2] * inv_hi, sc[kt][0][3] * inv_hi, pa[kt][1], pl[kt][1]);
      split2(sc[kt][1][0] * inv_lo, sc[kt][1][1] * inv_lo, pa[kt][2], pl[kt][2]);
      split2(sc[kt][1][2] * inv_hi, sc[kt][1][3] * inv_hi, pa[kt][3], pl[kt][3]);
      if (!HI) { pa[kt][1] = pa[kt][3] = pl[kt][1] = pl[kt][3] = 0u; }
      if (!(HI || kt + 1 < NKT)) { pa[kt][2] = pl[kt][2] = 0u; }
    } else {
    pa[kt][0] = pack_f16x2(sc[kt][0][0] * inv_lo, sc[kt][0][1] * inv_lo);   // (row lo, keys 0-7)
    pa[kt][1] = HI ? pack_f16x2(sc[kt][0][2] * inv_hi, sc[kt][0][3] * inv_hi) : 0u;   // (row hi, keys 0-7)
    pa[kt][2] = (HI || kt + 1 < NKT) ? pack_f16x2(sc[kt][1][0] * inv_lo, sc[kt][1][1] * inv_lo) : 0u;   // (row lo, keys 8-15)
    pa[kt][3] = HI ? pack_f16x2(sc[kt][1][2] * inv_hi, sc[kt][1][3] * inv_hi) : 0u;   // (row hi, keys 8-15)
    }
  }
  // ---- O = P V ----
  constexpr int NO = HSP / 8 / NH;               // 8-wide output column tiles of this warp
  const int nfirst = half * NO;                  // first of them inside the head
  float o[NO][4];
#pragma unroll
  for (int n = 0; n < NO; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
  for (int kt = 0; kt < NKT; ++kt) {
    const int r = min(row0 + kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, kLastRow);
#pragma unroll
    for (int np = 0; np < NO / 2; ++np) {        // pairs of 8-wide output column tiles
      uint32_t vb[4];
      ldmatrix_x4_trans(qkv + r * kQkvStride + (128 + (nfirst + np * 2) * 8 + (lane >> 4) * 8) * 2, vb);
      mma_16816(o[np * 2], pa[kt], vb[0], vb[1]);
      mma_16816(o[np * 2 + 1], pa[kt], vb[2], vb[3]);
      if constexpr (SPLIT) {
        mma_16816(o[np * 2], pl[kt], vb[0], vb[1]);                             // lo . hi
        mma_16816(o[np * 2 + 1], pl[kt], vb[2], vb[3]);
        ldmatrix_x4_trans(qkv + kLo + r * kQkvStride + (128 + (nfirst + np * 2) * 8 + (lane >> 4) * 8) * 2, vb);
        mma_16816(o[np * 2], pa[kt], vb[0], vb[1]);                             // hi . lo
        mma_16816(o[np * 2 + 1], pa[kt], vb[2], vb[3]);
      }
    }
  }
  // column e = n * 8 + (lane & 3) * 2 of the head: 16-byte chunk n of the row, bytes (lane & 3) * 4 within it
  uint32_t r_lo = (uint32_t)(row0 + i_lo), r_hi = (uint32_t)(row0 + i_hi);
  if constexpr (LAY == 1) {                       // sequence row -> its hi operand row (the lo row is 16 rows below)
    r_lo = ((r_lo >> 4) << 5) | (r_lo & 15u);
    r_hi = ((r_hi >> 4) << 5) | (r_hi & 15u);
  }
  if constexpr (LAY == 2) { r_lo += (uint32_t)yrow0; r_hi += (uint32_t)yrow0; }
  constexpr uint32_t kYLo = LAY == 2 ? G::SmYLo - G::SmY : 2048u;   // byte distance of the lo image of a Y element
  const uint32_t y_lo = sbase + G::SmY + r_lo * 128u + (uint32_t)(lane & 3) * 4u, x_lo = (r_lo & 7u) << 4;
  const uint32_t y_hi = sbase + G::SmY + r_hi * 128u + (uint32_t)(lane & 3) * 4u, x_hi = (r_hi & 7u) << 4;
  const uint32_t n0 = ((uint32_t)co >> 3) + (uint32_t)nfirst;   // first 16-byte chunk of this warp's columns inside the Y row
  if (ybar != 0u) spin_wait(ybar, ypar);          // (stays complete for the rest of the pass: later items pass at once)
  if (i_lo < T) {
#pragma unroll
    for (int n = 0; n < NO; ++n) {
      const uint32_t a = y_lo + (((n0 + (uint32_t)n) << 4) ^ x_lo);
      if constexpr (SPLIT) { uint32_t h, l; split2(o[n][0], o[n][1], h, l); sts32(a, h); sts32(a + kYLo, l); }
      else sts32(a, pack_f16x2(o[n][0], o[n][1]));
    }
  }
  if (HI && i_hi < T) {
#pragma unroll
    for (int n = 0; n < NO; ++n) {
      const uint32_t a = y_hi + (((n0 + (uint32_t)n) << 4) ^ x_hi);
      if constexpr (SPLIT) { uint32_t h, l; split2(o[n][2], o[n][3], h, l); sts32(a, h); sts32(a + kYLo, l); }
      else sts32(a, pack_f16x2(o[n][2], o[n][3]));
    }
  }
}
template <class G, int HSP, int LAY, int NH>
// Items [item0, item1) of the S * MT * NSUB items of a pass (heaviest first), dealt to n_slots slots.
__device__ __forceinline__ void attention_items(uint32_t sbase, int slot, int n_slots, int half, int lane, int S, int T, int yrow0,
                                                uint32_t ybar, uint32_t ypar, int item0 = 0, int item1 = 1 << 30) {
  constexpr int NSUB = 64 / HSP;
  const int MT = (T + 15) >> 4;                  // 16-row query tiles == 16-key steps
  item1 = min(item1, S * MT * NSUB);
  for (int item = item0 + slot; item < item1; item += n_slots) {
    const int sub = item % NSUB, it2 = item / NSUB;
    const int mt = MT - 1 - it2 / S, s = it2 % S;     // later query tiles see more keys: schedule them first
    const bool hi = mt * 16 + 8 < T;              // any of the query rows 8..15 of this tile inside the sequence?
    const int co = sub * HSP;
    if (mt == 0) { if (hi) attention_item<G, 1, true, HSP, LAY, NH>(sbase, lane, s * T, 0, T, co, half, yrow0, ybar, ypar); else attention_item<G, 1, false, HSP, LAY, NH>(sbase, lane, s * T, 0, T, co, half, yrow0, ybar, ypar); }
    else { if (hi) attention_item<G, 2, true, HSP, LAY, NH>(sbase, lane, s * T, mt, T, co, half, yrow0, ybar, ypar); else attention_item<G, 2, false, HSP, LAY, NH>(sbase, lane, s * T, mt, T, co, half, yrow0, ybar, ypar); }
  }
}
template <class G, int HSP, bool SPLIT>
__device__ __forceinline__ void attention_head_t(uint32_t sbase, int awarp, int lane, int S, int T, uint32_t ybar, uint32_t ypar) {
  // The precise mode's tiles hold at most 64 rows = 4 items for the 10 attention warps: pairs of warps share an item.
  // (Only there: the single-pass fp16 kernel has 8-10 items, and a second item body would only grow its image.)
  if constexpr (SPLIT) {
    if (S * ((T + 15) >> 4) * (64 / HSP) * 2 <= kAttnWarps) {
      attention_items<G, HSP, SPLIT ? 1 : 0, 2>(sbase, awarp >> 1, kAttnWarps / 2, awarp & 1, lane, S, T, 0, ybar, ypar);
      return;
    }
  }
  attention_items<G, HSP, SPLIT ? 1 : 0, 1>(sbase, awarp, kAttnWarps, 0, lane, S, T, 0, ybar, ypar);
}
// The padded head size is a template parameter of the kernel: only the attention code of the model's head size is in
// the kernel image (the fused kernel is ~13 k instructions; its hot paths have to stay resident in the instruction cache).
template <class G, int HSP>
__device__ __noinline__ void attention_head(uint32_t sbase, int awarp, int lane, int S, int T, uint32_t ybar, uint32_t ypar) {
  attention_head_t<G, HSP, false>(sbase, awarp, lane, S, T, ybar, ypar);
}

// FC1 chunk accumulator (buffer b) -> + b1 -> erf-GELU (packed fp16) -> H[b] (two K atoms, fp16).
// b1h_s = shared address of this chunk's 128 biases as fp16.
// tb = scratch accumulator the chunk sits in, b = H buffer it goes to (G256: the same index; G384: tb = 0 always).
template <class G>
__device__ __noinline__ void drain_gelu(const Compute c, int tb, int b, uint32_t b1h_s) {
  // Both 32-column pieces of this thread are read first (one exposed TMEM round trip per chunk) and the
  // accumulator is released at once; then 8 pairs at a time go through the GELU.  FC2's first k-block only
  // needs K atom 0 of H, which is signalled as soon as it is written.
  float va[32], vb[32];
  const uint32_t s_col = (tb ? G::ColS1 : G::ColS0) + c.hf * 32;
  const uint32_t h_s = c.sbase + (b ? G::SmH1 : G::SmH0);
  tmem_ld32(c.lane_addr(s_col), va);
  tmem_ld32(c.lane_addr(s_col + 64), vb);
  tmem_wait_ld();
  tc_fence_before();
  c.arrive(tb ? B_ACC_EMPTY1 : B_ACC_EMPTY0);
  auto emit = [&](const float* v, uint32_t col, uint32_t atom, uint32_t k0) {
    const uint4 bb0 = lds128_ro(b1h_s + col * 2u), bb1 = lds128_ro(b1h_s + col * 2u + 16u);
    __half2 x[8];
    x[0] = __hadd2(__floats2half2_rn(v[0], v[1]), bits2h(bb0.x));
    x[1] = __hadd2(__floats2half2_rn(v[2], v[3]), bits2h(bb0.y));
    x[2] = __hadd2(__floats2half2_rn(v[4], v[5]), bits2h(bb0.z));
    x[3] = __hadd2(__floats2half2_rn(v[6], v[7]), bits2h(bb0.w));
    x[4] = __hadd2(__floats2half2_rn(v[8], v[9]), bits2h(bb1.x));
    x[5] = __hadd2(__floats2half2_rn(v[10], v[11]), bits2h(bb1.y));
    x[6] = __hadd2(__floats2half2_rn(v[12], v[13]), bits2h(bb1.z));
    x[7] = __hadd2(__floats2half2_rn(v[14], v[15]), bits2h(bb1.w));
    gelu2_vec<8>(x);
    sts128(c.chunk_addr(atom, k0), h2bits(x[0]), h2bits(x[1]), h2bits(x[2]), h2bits(x[3]));
    sts128(c.chunk_addr(atom, k0 + 1), h2bits(x[4]), h2bits(x[5]), h2bits(x[6]), h2bits(x[7]));
  };
  emit(va, c.hf * 32, h_s, c.hf * 4);
  emit(va + 16, c.hf * 32 + 16, h_s, c.hf * 4 + 2);
  fence_async_smem();                     // K atom 0 of H is complete: FC2's first k-block may start
  c.arrive(B_OP_READY0 + b);
  emit(vb, 64 + c.hf * 32, h_s + 16384, c.hf * 4);
  emit(vb + 16, 64 + c.hf * 32 + 16, h_s + 16384, c.hf * 4 + 2);
  fence_async_smem();
  c.arrive(B_OP_READY0B + b);
}

// ================================ PREC: fp32-equivalent arithmetic on the tensor pipe ============================
// The tile holds 64 sequence rows.  Every 16-bit operand is split x = hi + lo (two fp16 values, 22 mantissa bits)
// and the MMA row dimension carries both images: TMEM lane 32 q + j (j < 16) is the "hi" row of sequence row
// 16 q + j, lane 32 q + 16 + j its "lo" row.  One M = 128 MMA with the hi image of a weight tile followed by one
// with its lo image accumulates [A_hi; A_lo] (W_hi + W_lo)^T, so that for every sequence row
//     D[hi lane] + D[lo lane] = (A_hi + A_lo) (W_hi + W_lo)^T            (all four cross terms, fp32 accumulate).
// The two lanes of a row belong to the same warp: every drain adds them with one __shfl_xor(16) per element and
// the two threads then share the row's columns.  LayerNorm is two-pass fp32, GELU is erff, attention is fp32 FMA
// out of an fp32 staging buffer, the residual stream stays fp32 in TMEM (split over the two lanes).
// The schedule, the barriers and the shared / tensor memory maps are those of the fp16 mode; the weight tape holds
// [hi tile | lo tile] per ring group.
__device__ __forceinline__ float shx16(float v) { return __shfl_xor_sync(0xffffffffu, v, 16); }
// 8 consecutive K elements -> the 16-byte chunk of the hi row and of the lo row (16 rows = 2048 bytes below)
__device__ __forceinline__ void st_chunk_split(uint32_t addr_hi, const float* v) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split2(v[2 * i], v[2 * i + 1], h[i], l[i]);
  sts128(addr_hi, h[0], h[1], h[2], h[3]);
  sts128(addr_hi + 2048u, l[0], l[1], l[2], l[3]);
}

// A <- split(LayerNorm0(X + pend)).  Each thread reads 128 columns of its TMEM lane, hands 64 of them to its pair
// thread and receives the pair's 64 in exchange: afterwards it owns 64 complete columns of the sequence row.
template <bool DBG>
__device__ __noinline__ void ln_pass_p(const Compute c, uint32_t vec_s, float inv_d, int d_true, float* trace_row) {
  float va[32], vb[32], x[64];
  const int col0 = c.hf * 128;
  const int keep0 = col0 + c.is_lo * 64;
  tmem_ld32(c.lane_addr(kColX + col0), va);
  tmem_ld32(c.lane_addr(kColX + col0 + 64), vb);
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = (c.is_lo ? vb[i] : va[i]) + shx16(c.is_lo ? va[i] : vb[i]);
  tmem_ld32(c.lane_addr(kColX + col0 + 32), va);
  tmem_ld32(c.lane_addr(kColX + col0 + 96), vb);
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 32; ++i) x[32 + i] = (c.is_lo ? vb[i] : va[i]) + shx16(c.is_lo ? va[i] : vb[i]);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 64; i += 4) {
    const float4 pd = lds128f_ro(vec_s + (uint32_t)(keep0 + i) * 4u);
    x[i] += pd.x; x[i + 1] += pd.y; x[i + 2] += pd.z; x[i + 3] += pd.w;
    s += (x[i] + x[i + 1]) + (x[i + 2] + x[i + 3]);
  }
  if (DBG) {
    if (trace_row != nullptr) {
#pragma unroll
      for (int i = 0; i < 64; ++i) trace_row[keep0 + i] = x[i];
    }
  }
  s += shx16(s);
  float* stats = reinterpret_cast<float*>(c.sm + kSmStats);        // [sum: 2 x 64 | squares: 2 x 64]
  if (!c.is_lo) stats[c.hf * 64 + c.srow] = s;
  compute_sync();
  const float mean = (s + stats[(c.hf ^ 1) * 64 + c.srow]) * inv_d;
  // second pass over the true lanes only (padding columns are zero, not mean)
  const int nv = min(max(d_true - keep0, 0), 64);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 64; ++i) { const float t = x[i] - mean; x[i] = t; if (i < nv) q = fmaf(t, t, q); }
  q += shx16(q);
  if (!c.is_lo) stats[128 + c.hf * 64 + c.srow] = q;
  compute_sync();
  const float var = (q + stats[128 + (c.hf ^ 1) * 64 + c.srow]) * inv_d;
  const float rstd = 1.0f / sqrtf(var + 1e-5f);
  // K atom of this thread's 64 columns; hi row = lane & ~16, the lo row is 16 rows (2048 bytes) further down
  const uint32_t a_hi = c.sbase + kSmA + (uint32_t)(c.hf * 2 + c.is_lo) * 16384u + c.row_off - (uint32_t)c.is_lo * 2048u;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = x[k * 8 + i] * rstd;
    st_chunk_split(a_hi + (((uint32_t)k << 4) ^ c.rx4), y);
  }
  fence_async_smem();
  tc_fence_before();
  c.arrive(B_A_READY);
}

// G384: 192 columns per TMEM lane.  In 16-column pieces: the even pieces stay with the hi-lane thread, the odd pieces with
// the lo-lane thread, so that each of them ends up with 96 complete columns in 16-column (two-chunk) runs.
template <bool DBG>
__device__ __noinline__ void ln_pass_pw(const Compute c, uint32_t vec_s, float inv_d, int d_true, float* trace_row) {
  float x[96];
  const int col0 = c.hf * 192;
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    float va[16], vb[16];
    tmem_ld16(c.lane_addr(kColX + col0 + r * 32), va);
    tmem_ld16(c.lane_addr(kColX + col0 + r * 32 + 16), vb);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) x[r * 16 + i] = (c.is_lo ? vb[i] : va[i]) + shx16(c.is_lo ? va[i] : vb[i]);
  }
  const int cbase = col0 + c.is_lo * 16;                  // piece r of this thread = columns cbase + 32 r .. + 16
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 pd = lds128f_ro(vec_s + (uint32_t)(cbase + r * 32 + i) * 4u);
      float* e = x + r * 16 + i;
      e[0] += pd.x; e[1] += pd.y; e[2] += pd.z; e[3] += pd.w;
      s += (e[0] + e[1]) + (e[2] + e[3]);
    }
  if (DBG) {
    if (trace_row != nullptr) {
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) trace_row[cbase + r * 32 + i] = x[r * 16 + i];
    }
  }
  s += shx16(s);
  float* stats = reinterpret_cast<float*>(c.sm + G384::SmStats);  // [sum: 2 x 64 | squares: 2 x 64]
  if (!c.is_lo) stats[c.hf * 64 + c.srow] = s;
  compute_sync();
  const float mean = (s + stats[(c.hf ^ 1) * 64 + c.srow]) * inv_d;
  float q = 0.f;
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    const int nv = d_true - (cbase + r * 32);            // true lanes of this piece (padding columns are zero, not mean)
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float t = x[r * 16 + i] - mean; x[r * 16 + i] = t; if (i < nv) q = fmaf(t, t, q); }
  }
  q += shx16(q);
  if (!c.is_lo) stats[128 + c.hf * 64 + c.srow] = q;
  compute_sync();
  const float var = (q + stats[128 + (c.hf ^ 1) * 64 + c.srow]) * inv_d;
  const float rstd = 1.0f / sqrtf(var + 1e-5f);
  // hi row = lane & ~16, the lo row is 16 rows (2048 bytes) further down
  const uint32_t a_row = c.sbase + kSmA + c.row_off - (uint32_t)c.is_lo * 2048u;
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    const int col = cbase + r * 32;
    const uint32_t a_hi = a_row + (uint32_t)(col >> 6) * 16384u, k0 = (uint32_t)(col & 63) >> 3;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      float y[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = x[r * 16 + k * 8 + i] * rstd;
      st_chunk_split(a_hi + (((k0 + (uint32_t)k) << 4) ^ c.rx4), y);
    }
  }
  fence_async_smem();
  tc_fence_before();
  c.arrive(B_A_READY);
}

// [Q|K|V] accumulator of one attention pass -> fp16 staging, hi image in rows [0, 64) and lo image in rows [64, 128)
// (Q gets its bias).  Per 32-column piece the pair threads exchange halves: each ends up with 16 columns of Q, of K
// and of V of the sequence row.
template <class G, int PARTS>
__device__ __noinline__ void drain_qkv_p(const Compute c, uint32_t bq_s) {
  float v0[32], v1[32], v2[32];
  const int colb = c.hf * 32;
  if constexpr ((PARTS & 1) != 0) tmem_ld32(c.lane_addr(G::ColS0 + colb), v0);
  if constexpr ((PARTS & 2) != 0) tmem_ld32(c.lane_addr(G::ColS0 + 64 + colb), v1);
  if constexpr ((PARTS & 4) != 0) tmem_ld32(c.lane_addr(G::ColS0 + (PARTS == 7 ? 128 : 0) + colb), v2);
  tmem_wait_ld();
  tc_fence_before();
  c.arrive(B_ACC_EMPTY0);
  const int cs = colb + c.is_lo * 16;                     // first of this thread's 16 columns inside each 64-block
  const uint32_t dst = c.sbase + G::SmQkv + (uint32_t)c.srow * kQkvStride + (uint32_t)cs * 2u;
  auto emit = [&](const float (&v)[32], uint32_t d, bool bias) {
    float r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = (c.is_lo ? v[16 + i] : v[i]) + shx16(c.is_lo ? v[i] : v[16 + i]);
    if (bias) {
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 b0 = lds128f_ro(bq_s + (uint32_t)(cs + i) * 4u);
        r[i] += b0.x; r[i + 1] += b0.y; r[i + 2] += b0.z; r[i + 3] += b0.w;
      }
    }
    uint32_t h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split2(r[2 * i], r[2 * i + 1], h[i], l[i]);
    sts128(d, h[0], h[1], h[2], h[3]);
    sts128(d + 16, h[4], h[5], h[6], h[7]);
    sts128(d + 64u * kQkvStride, l[0], l[1], l[2], l[3]);
    sts128(d + 64u * kQkvStride + 16, l[4], l[5], l[6], l[7]);
  };
  if constexpr ((PARTS & 1) != 0) emit(v0, dst, true);
  if constexpr ((PARTS & 2) != 0) emit(v1, dst + 128, false);
  if constexpr ((PARTS & 4) != 0) emit(v2, dst + 256, false);
}

// Causal attention of the precise mode: the mma.sync kernel of the fp16 mode with every product split three ways
// (attention_item<..., SPLIT = true>): the operands come out of shared memory once per 16 x 8 tile, which keeps the
// attention phase off the shared-memory port the tensor pipe is streaming its operands through.
template <class G, int HSP>
__device__ __noinline__ void attention_head_p(uint32_t sbase, int awarp, int lane, int S, int T, uint32_t ybar, uint32_t ypar) {
  attention_head_t<G, HSP, true>(sbase, awarp, lane, S, T, ybar, ypar);
}

__device__ __forceinline__ float gelu_erf(float u) { return 0.5f * u * (1.0f + erff(u * 0.70710678118654752440f)); }
// The same function in 16 fp32 instructions + ONE MUFU instead of erff's ~40 with a divergent branch:
//   gelu(x) = x Phi(x),  Phi(x) = 1 - erfc(a) / 2 (x >= 0),  erfc(a) / 2 (x < 0),  a = min(|x| / sqrt 2, 5.7),
//   erfc(a) = 2^p(t),  t = a / 2.85 - 1 in [-1, 1],  p of degree 8
// p is fitted to log2 erfc with the ABSOLUTE error of erfc as the weight (tools/fit_gelu.py): far in the tail, where erfc
// is below 1e-12, the exponent is only roughly right, which nothing downstream can see.  Evaluated in fp32 the GELU is
// within 6.3e-7 of the exact one over [-10, 10]; 0.5 x (1 + erf(x / sqrt 2)) in fp32 is within 6.8e-7.  No cancellation on
// either side: the negative branch never forms 1 - (1 - small).
// (The first version used the A&S 7.1.26 form, rcp + ex2 per element; this one needs no reciprocal and is three
// instructions shorter.  Measured: same speed.  What does slow the MLP half of the 128-row precise layout is the drain as a
// whole: the two compute warps that share a scheduler with the MMA-issuing warp delay its short dependent instruction
// chains -- every MMA takes 83 cycles instead of 68 while they run this code and 68 when, as an experiment, their GELU
// math is switched off (in-kernel timeline).  Periodic nanosleep(0) in those two warps recovered only a tenth of it.)
__device__ __forceinline__ float gelu_fast32(float x) {
  const float a = fminf(fabsf(x) * 0.70710678118654752440f, 5.7f);
  const float t = fmaf(a, 2.0f / 5.7f, -1.0f);
  float p = fmaf(-0.176724888f, t, -0.795554992f);
  p = fmaf(p, t, -1.34464668f);
  p = fmaf(p, t, -1.44855133f);
  p = fmaf(p, t, -0.667739718f);
  p = fmaf(p, t, -0.570451905f);
  p = fmaf(p, t, -11.2386845f);
  p = fmaf(p, t, -24.7465583f);
  p = fmaf(p, t, -14.1333206f);
  const float half_erfc = 0.5f * ex2f(p);
  return x * (x >= 0.0f ? 1.0f - half_erfc : half_erfc);
}
// FC1 chunk accumulator (buffer b) -> + b1 -> exact erf-GELU (fp32) -> split -> H[b] (two K atoms).
// b1_s = shared address of this chunk's 128 biases (fp32).
template <class G>
__device__ __noinline__ void drain_gelu_p(const Compute c, int tb, int b, uint32_t b1_s) {
  float va[32], vb[32];
  const uint32_t s_col = (tb ? G::ColS1 : G::ColS0) + c.hf * 32;
  const uint32_t h_hi = c.sbase + (b ? G::SmH1 : G::SmH0) + c.row_off - (uint32_t)c.is_lo * 2048u;
  tmem_ld32(c.lane_addr(s_col), va);
  tmem_ld32(c.lane_addr(s_col + 64), vb);
  tmem_wait_ld();
  tc_fence_before();
  c.arrive(tb ? B_ACC_EMPTY1 : B_ACC_EMPTY0);
  const int cs = c.hf * 32 + c.is_lo * 16;                // this thread's 16 columns inside each 64-wide K atom
  auto emit = [&](const float (&v)[32], uint32_t bias_s, uint32_t atom) {
    float g[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) g[i] = (c.is_lo ? v[16 + i] : v[i]) + shx16(c.is_lo ? v[i] : v[16 + i]);
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 b0 = lds128f_ro(bias_s + (uint32_t)(cs + i) * 4u);
      g[i] = gelu_fast32(g[i] + b0.x); g[i + 1] = gelu_fast32(g[i + 1] + b0.y);
      g[i + 2] = gelu_fast32(g[i + 2] + b0.z); g[i + 3] = gelu_fast32(g[i + 3] + b0.w);
    }
    const uint32_t k0 = (uint32_t)cs >> 3;
    st_chunk_split(atom + ((k0 << 4) ^ c.rx4), g);
    st_chunk_split(atom + (((k0 + 1) << 4) ^ c.rx4), g + 8);
  };
  emit(va, b1_s, h_hi);
  fence_async_smem();                     // K atom 0 of H is complete: FC2's first k-block may start
  c.arrive(B_OP_READY0 + b);
  emit(vb, b1_s + 256u, h_hi + 16384u);
  fence_async_smem();
  c.arrive(B_OP_READY0B + b);
}

// ================================ P128: the precise mode on full 128-row tiles (geometry G256P) =================
// One thread per (row, column half) as in the fp16 mode; every operand is written as an fp16 hi image and an fp16 lo
// image (x = hi + lo).  Row -> sequence mapping: rows [0, 64) hold the first S / 2 sequences, rows [64, 128) the rest.
__device__ __forceinline__ void st_words16(uint32_t dst, const uint32_t (&w)[16]) {       // 32 fp16 = 64 contiguous bytes
  sts128(dst, w[0], w[1], w[2], w[3]); sts128(dst + 16, w[4], w[5], w[6], w[7]);
  sts128(dst + 32, w[8], w[9], w[10], w[11]); sts128(dst + 48, w[12], w[13], w[14], w[15]);
}
__device__ __forceinline__ void split32(const float (&v)[32], uint32_t (&h)[16], uint32_t (&l)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) split2(v[2 * i], v[2 * i + 1], h[i], l[i]);
}

// A <- split(LayerNorm0(X + pend)): hi image -> the shared-memory A atoms, lo image -> tensor memory (A_lo, packed
// pairs).  Exact two-pass statistics in three sweeps over X (mean | centred squares | normalise): a sweep is four
// tcgen05.ld of 32 columns, cheaper than parking 128 values per thread.
template <bool DBG>
__device__ __noinline__ void ln_pass_p128(const Compute c, uint32_t vec_s, float inv_d, int d_true, float* trace_row) {
  using G = G256P;
  float va[32], vb[32];
  const int col0 = c.hf * 128;
  float* stats = reinterpret_cast<float*>(c.sm + G::SmStats);        // [sum: 2 x 128 | squares: 2 x 128]
  auto start = [&]() {
    tmem_ld32(c.lane_addr(kColX + col0), va);
    tmem_ld32(c.lane_addr(kColX + col0 + 32), vb);
  };
  auto sweep = [&](auto&& f) {                                       // chunks 0, 1 are in flight on entry
    tmem_wait_ld();
    f(va, 0);
    tmem_ld32(c.lane_addr(kColX + col0 + 64), va);
    f(vb, 1);
    tmem_ld32(c.lane_addr(kColX + col0 + 96), vb);
    tmem_wait_ld();
    f(va, 2);
    f(vb, 3);
  };
  // (volatile loads of pend: the three sweeps must not share them and keep 128 values live)
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  start();
  sweep([&](float (&v)[32], int ch) {
    const int col = col0 + ch * 32;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 pd = lds128f_v(vec_s + (uint32_t)(col + i) * 4u);
      const float a0 = v[i] + pd.x, a1 = v[i + 1] + pd.y, a2 = v[i + 2] + pd.z, a3 = v[i + 3] + pd.w;
      s0 += a0; s1 += a1; s2 += a2; s3 += a3;
      if (DBG) {
        if (trace_row != nullptr) { float* tr = trace_row + col + i; tr[0] = a0; tr[1] = a1; tr[2] = a2; tr[3] = a3; }
      }
    }
  });
  const float sum = (s0 + s1) + (s2 + s3);
  stats[c.hf * kRows + c.row] = sum;
  start();
  compute_sync();
  const float mean = (sum + stats[(c.hf ^ 1) * kRows + c.row]) * inv_d;
  float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
  sweep([&](float (&v)[32], int ch) {
    const int col = col0 + ch * 32, nv = d_true - col;               // true lanes of this chunk (padding columns are zero, not mean)
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 pd = lds128f_v(vec_s + (uint32_t)(col + i) * 4u);
      const float t0 = (v[i] + pd.x) - mean, t1 = (v[i + 1] + pd.y) - mean, t2 = (v[i + 2] + pd.z) - mean, t3 = (v[i + 3] + pd.w) - mean;
      if (i < nv) q0 = fmaf(t0, t0, q0);
      if (i + 1 < nv) q1 = fmaf(t1, t1, q1);
      if (i + 2 < nv) q2 = fmaf(t2, t2, q2);
      if (i + 3 < nv) q3 = fmaf(t3, t3, q3);
    }
  });
  const float sq = (q0 + q1) + (q2 + q3);
  stats[256 + c.hf * kRows + c.row] = sq;
  start();
  compute_sync();
  const float var = (sq + stats[256 + (c.hf ^ 1) * kRows + c.row]) * inv_d;
  const float rstd = 1.0f / sqrtf(var + 1e-5f);
  sweep([&](float (&v)[32], int ch) {
    const int col = col0 + ch * 32;                                  // 32 columns = chunks k0 .. k0 + 3 of one K atom
    const uint32_t atom = c.sbase + kSmA + (uint32_t)(col >> 6) * 16384u, k0 = (uint32_t)(col & 63) >> 3;
    uint32_t lo[16];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 p0 = lds128f_v(vec_s + (uint32_t)(col + k * 8) * 4u), p1 = lds128f_v(vec_s + (uint32_t)(col + k * 8 + 4) * 4u);
      const float* e = v + k * 8;
      uint32_t h[4];
      split2(((e[0] + p0.x) - mean) * rstd, ((e[1] + p0.y) - mean) * rstd, h[0], lo[k * 4 + 0]);
      split2(((e[2] + p0.z) - mean) * rstd, ((e[3] + p0.w) - mean) * rstd, h[1], lo[k * 4 + 1]);
      split2(((e[4] + p1.x) - mean) * rstd, ((e[5] + p1.y) - mean) * rstd, h[2], lo[k * 4 + 2]);
      split2(((e[6] + p1.z) - mean) * rstd, ((e[7] + p1.w) - mean) * rstd, h[3], lo[k * 4 + 3]);
      sts128(c.chunk_addr(atom, k0 + (uint32_t)k), h[0], h[1], h[2], h[3]);
    }
    tmem_st16(c.lane_addr(G::ColAL + (uint32_t)(col >> 1)), lo);     // K element k -> column k / 2
  });
  tmem_wait_st();
  fence_async_smem();
  tc_fence_before();
  c.arrive(B_A_READY);
}

// Causal attention of one 64-row half of the tile out of the hi / lo staging buffer; Y rows yrow0 .. yrow0 + 63.
template <int HSP>
__device__ __noinline__ void attention_half_p128(uint32_t sbase, int slot, int n_slots, int lane, int Sh, int T, int yrow0,
                                                 uint32_t ybar, uint32_t ypar) {
  using G = G256P;
  // One round when the items fit the slots: the n2 heaviest items (later query tiles: more keys) are shared by two
  // warps each, the rest take one warp (BASELINE config 2, first half: 4 items on 6 warps -> the two 23-key items are
  // split).  More items than slots: several rounds of whole items.
  const int n_items = Sh * ((T + 15) >> 4) * (64 / HSP);
  const int n2 = n_items <= n_slots ? min(n_items, n_slots - n_items) : 0;
  if (slot < 2 * n2) attention_items<G, HSP, 2, 2>(sbase, slot >> 1, n2, slot & 1, lane, Sh, T, yrow0, ybar, ypar, 0, n2);
  else attention_items<G, HSP, 2, 1>(sbase, slot - 2 * n2, n_slots - 2 * n2, 0, lane, Sh, T, yrow0, ybar, ypar, n2);
}

// One attention pass of the compute warps: [Q|K] accumulator, then V accumulator -> registers (hi / lo packed) ->
// staging and attention, half by half.  The warps of rows 64..127 hold their 96 packed words while the first half
// is processed by the warps of rows 0..63 and the helpers.  Barriers of the 10 attention warps per pass:
//   S1 staging(half 0) complete | S2 attention(half 0) done | S3 staging(half 1) complete | S4 attention(half 1) done
// (Inlined into the compute role's loop: compute-sanitizer's synccheck wants the attn_sync() calls of all ten warps at
// the same call depth -- with this function out of line it reports the first barrier of a pass as divergent.)
template <int HSP>
__device__ __forceinline__ uint32_t attention_pass_p128(Compute c, uint32_t bq_s, int S, int T) {
  using G = G256P;
  const int colb = c.hf * 32, Sh = S >> 1;
  const bool upper = c.wq >= 2;                                      // warp-uniform
  uint32_t qh[16], ql[16], kh[16], kl[16], vh[16], vl[16];
  const uint32_t dst = c.sbase + G::SmQkv + (uint32_t)(c.row & 63) * kQkvStride + (uint32_t)colb * 2u;
  constexpr uint32_t kLoImg = 64u * kQkvStride;
  {
    float v0[32], v1[32];
    c.wait(B_ACC_FULL0);
    tc_fence_after();
    tmem_ld32(c.lane_addr(G::ColS0 + colb), v0);
    tmem_ld32(c.lane_addr(G::ColS0 + 64 + colb), v1);
    tmem_wait_ld();
    tc_fence_before();
    c.arrive(B_ACC_EMPTY0);                                          // the V job may overwrite the accumulator
#pragma unroll
    for (int i = 0; i < 32; i += 4) {                                // Q (+ bias)
      const float4 b0 = lds128f_ro(bq_s + (uint32_t)(colb + i) * 4u);
      v0[i] += b0.x; v0[i + 1] += b0.y; v0[i + 2] += b0.z; v0[i + 3] += b0.w;
    }
    split32(v0, qh, ql);
    split32(v1, kh, kl);
    if (!upper) {                                                    // (the V job is running)
      st_words16(dst, qh); st_words16(dst + kLoImg, ql);
      st_words16(dst + 128, kh); st_words16(dst + 128 + kLoImg, kl);
    }
    c.wait(B_ACC_FULL0);
    tc_fence_after();
    tmem_ld32(c.lane_addr(G::ColS0 + colb), v0);
    tmem_wait_ld();
    tc_fence_before();
    c.arrive(B_ACC_EMPTY0);
    split32(v0, vh, vl);
  }
  const int awarp = c.ctid >> 5;                                     // 0..7; rows 0..63 are warps 0, 1, 4, 5
  if (!upper) { st_words16(dst + 256, vh); st_words16(dst + 256 + kLoImg, vl); }
  // the previous pass's projection (issued behind this pass's V job) must have read Y before Y is rewritten: the
  // attention items wait for it right before their output stores
  const uint32_t ybar = c.bar(B_Y_EMPTY), ypar = (c.phases >> B_Y_EMPTY) & 1u;
  c.phases ^= 1u << B_Y_EMPTY;
  attn_sync();                                                       // S1
  if (!upper) attention_half_p128<HSP>(c.sbase, (awarp & 1) | ((awarp >> 2) << 1), 6, c.lane, Sh, T, 0, ybar, ypar);
  attn_sync();                                                       // S2
  if (upper) {
    st_words16(dst, qh); st_words16(dst + kLoImg, ql);
    st_words16(dst + 128, kh); st_words16(dst + 128 + kLoImg, kl);
    st_words16(dst + 256, vh); st_words16(dst + 256 + kLoImg, vl);
  }
  attn_sync();                                                       // S3
  attention_half_p128<HSP>(c.sbase, awarp, kAttnWarps, c.lane, Sh, T, 64, ybar, ypar);
  fence_async_smem();
  c.arrive(B_Y_READY);
  attn_sync();                                                       // S4: staging may be overwritten by the next pass
  return c.phases;
}

// FC1 chunk accumulator -> + b1 -> exact erf-GELU (fp32) -> hi / lo images of H (one buffer: FC2 of the previous
// chunk must have read it before it is overwritten).  b1_s = shared address of this chunk's 128 biases (fp32).
__device__ __noinline__ uint32_t drain_gelu_p128(Compute c, uint32_t b1_s) {
  using G = G256P;
  float va[32], vb[32];
  const uint32_t s_col = G::ColS0 + c.hf * 32;
  tmem_ld32(c.lane_addr(s_col), va);
  tmem_ld32(c.lane_addr(s_col + 64), vb);
  tmem_wait_ld();
  tc_fence_before();
  c.arrive(B_ACC_EMPTY0);
  auto gelu32 = [&](float (&v)[32], uint32_t bias_s, uint32_t (&h)[16], uint32_t (&l)[16]) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b0 = lds128f_ro(bias_s + (uint32_t)i * 4u);
      v[i] = gelu_fast32(v[i] + b0.x); v[i + 1] = gelu_fast32(v[i + 1] + b0.y);
      v[i + 2] = gelu_fast32(v[i + 2] + b0.z); v[i + 3] = gelu_fast32(v[i + 3] + b0.w);
    }
    split32(v, h, l);
  };
  // this thread's 32 columns of a K atom = its chunks hf * 4 .. hf * 4 + 3
  auto store = [&](uint32_t atom_off, const uint32_t (&h)[16], const uint32_t (&l)[16]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t a = c.chunk_addr(c.sbase + G::SmH0 + atom_off, (uint32_t)(c.hf * 4 + k));
      sts128(a, h[4 * k], h[4 * k + 1], h[4 * k + 2], h[4 * k + 3]);
      sts128(a + (G::SmHLo - G::SmH0), l[4 * k], l[4 * k + 1], l[4 * k + 2], l[4 * k + 3]);
    }
  };
  uint32_t h[16], l[16];
  gelu32(va, b1_s + (uint32_t)(c.hf * 32) * 4u, h, l);
  c.wait(B_OP_EMPTY0);                    // FC2 of the previous chunk has consumed H
  store(0, h, l);
  fence_async_smem();                     // K atom 0 of H is complete: FC2's first k-block may start
  c.arrive(B_OP_READY0);
  gelu32(vb, b1_s + (uint32_t)(64 + c.hf * 32) * 4u, h, l);
  store(16384, h, l);
  fence_async_smem();
  c.arrive(B_OP_READY0B);
  return c.phases;
}

// ---- the single-warp roles: one out-of-line step per ring group (small I-cache footprint).  All state is
// passed and returned by value: with 227 KB of shared memory there is no L1 left, so anything that lands in
// local memory (address-taken structs, spills) costs an L2 round trip per access.
// The whole warp runs the (warp-uniform) schedule so that addresses stay in uniform registers; one elected
// lane issues the copies.  CG = 2: this CTA streams only its half of the rows of each group.
template <int CG>
__device__ __forceinline__ uint32_t producer_step(uint32_t n, uint32_t pair, uint32_t kk2, uint32_t sbase, uint32_t rank,
                                               const uint8_t* src, uint32_t g) {
  const uint32_t bytes = n * 128u * (kk2 ? 2u : 1u);
  if (CG == 2) {
    const uint32_t slot = g & (kSlots - 1), par = (g >> 2) & 1u, half = bytes >> 1;
    const uint32_t full = sbase + kSmBars + (B_FULL0 + slot) * 8;
    const uint32_t dst = sbase + kSmRing + slot * kSlotBytes;
    spin_wait(sbase + kSmBars + (B_EMPTY0 + slot) * 8, par ^ 1u);
    if (elect_one()) {
      mbar_expect_tx(full, half);
      if (kk2) {          // this CTA's rows of K block 0, then of K block 1
        const uint32_t q = half >> 1;
        bulk_g2s(dst, src + rank * q, q, full);
        bulk_g2s(dst + q, src + 2u * q + rank * q, q, full);
      } else {
        bulk_g2s(dst, src + rank * half, half, full);
      }
    }
    __syncwarp();
    return g + 1;
  }
  // one 16 KB slot per fill; a pair / double-K group occupies two adjacent slots (even, odd)
  const uint32_t first = (pair || kk2) ? kSlotBytes : bytes;
#pragma unroll 1
  for (uint32_t done = 0; done < bytes;) {
    const uint32_t slot = g & (kSlots - 1), par = (g >> 2) & 1u;
    const uint32_t len = done == 0 ? first : bytes - first;
    const uint32_t full = sbase + kSmBars + (B_FULL0 + slot) * 8;
    spin_wait(sbase + kSmBars + (B_EMPTY0 + slot) * 8, par ^ 1u);
    if (elect_one()) {
      mbar_expect_tx(full, len);
      bulk_g2s(sbase + kSmRing + slot * kSlotBytes, src + done, len, full);
    }
    __syncwarp();
    done += len;
    g += 1;
  }
  return g;
}

// Warp-uniform control flow and operands, no memory loads on the issue path; one elected lane issues
// tcgen05.mma / tcgen05.commit.  sync = w0 | w1 << 4 | c0 << 8 | c1 << 12; flags = acc | pair << 1.
// Returns the updated (ring counter | barrier parities << 32).
template <int CG, bool TL>
__device__ __forceinline__ uint64_t mma_step(uint32_t a_off, uint32_t d_col, uint32_t n, uint32_t flags, uint32_t sync,
                                             uint32_t sbase, uint32_t tm, uint32_t g, uint32_t phases, long long* tl_in) {
  long long* tl = TL ? tl_in : nullptr;            // the timeline code compiles out of the production instantiation
  const uint32_t w0 = sync & 0xF, w1 = (sync >> 4) & 0xF, c0 = (sync >> 8) & 0xF, c1 = (sync >> 12) & 0xF;
  const uint32_t acc = flags & 1u;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (tl != nullptr) t0 = clock64();
  // All barrier tests of this group are issued back to back (their ~100-cycle latencies overlap); the slow
  // path spins only on those that were not complete yet.
  const uint32_t slot = g & (kSlots - 1), par = (g >> 2) & 1u;
  const bool kk2 = (flags & 4u) != 0;
  const bool two = CG == 1 && (flags & 6u);
  const uint32_t bw0 = sbase + kSmBars + (w0 != kNone ? w0 : 0u) * 8, pw0 = (phases >> (w0 & 31u)) & 1u;
  const uint32_t bw1 = sbase + kSmBars + (w1 != kNone ? w1 : 0u) * 8, pw1 = (phases >> (w1 & 31u)) & 1u;
  const uint32_t bf0 = sbase + kSmBars + (B_FULL0 + slot) * 8, bf1 = bf0 + 8;
  const uint32_t bpf = sbase + kSmBars + (B_PFULL0 + slot) * 8;
  bool r_w0 = true, r_w1 = true, r_f1 = true, r_pf = true;
  if (w0 != kNone) r_w0 = CG == 2 ? mbar_try_wait_cluster(bw0, pw0) : mbar_try_wait(bw0, pw0);
  if (w1 != kNone) r_w1 = CG == 2 ? mbar_try_wait_cluster(bw1, pw1) : mbar_try_wait(bw1, pw1);
  const bool r_f0 = mbar_try_wait(bf0, par);
  if (two) r_f1 = mbar_try_wait(bf1, par);
  if (CG == 2) r_pf = mbar_try_wait_cluster(bpf, par);
  if (!r_w0) { if (CG == 2) wait_timeout_cluster(bw0, pw0); else wait_timeout(bw0, pw0); }
  if (!r_w1) { if (CG == 2) wait_timeout_cluster(bw1, pw1); else wait_timeout(bw1, pw1); }
  if (w0 != kNone) phases ^= 1u << w0;
  if (w1 != kNone) phases ^= 1u << w1;
  if (tl != nullptr) t1 = clock64();
  if (!r_f0) wait_timeout(bf0, par);
  if (!r_f1) wait_timeout(bf1, par);
  if (!r_pf) wait_timeout_cluster(bpf, par);
  if (tl != nullptr) t2 = clock64();
  tc_fence_after();
  const uint64_t a_desc = smem_desc_sw128(sbase + a_off);
  const uint64_t b_desc = smem_desc_sw128(sbase + kSmRing + slot * kSlotBytes);
  // fp16 operands everywhere except the embedding GEMM (flag 8): its A operand carries raw, unscaled inputs
  // and bf16 hi + lo splits of the bias / position tables
  const uint32_t idesc = (flags & 8u) ? (CG == 2 ? idesc_bf16_m256(n) : idesc_bf16_m128(n))
                                      : (CG == 2 ? idesc_f16_m256(n) : idesc_f16_m128(n));
  const uint32_t d_addr = tm + d_col;
  if (elect_one()) {
    long long t3 = 0;
    if (CG == 2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) mma_bf16_cg2(d_addr, a_desc + 2u * j, b_desc + 2u * j, idesc, (acc | j) ? 1u : 0u);
      if (kk2) {          // second K block: next A atom, second half of the slot
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_bf16_cg2(d_addr, a_desc + 1024u + 2u * j, b_desc + (n * 64u >> 4) + 2u * j, idesc, 1u);
      }
      if (tl != nullptr) t3 = clock64();
      mma_commit_cg2(sbase + kSmBars + (B_EMPTY0 + slot) * 8);
      if (c0 != kNone) mma_commit_cg2(sbase + kSmBars + c0 * 8);
      if (c1 != kNone) mma_commit_cg2(sbase + kSmBars + c1 * 8);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) mma_bf16(d_addr, a_desc + 2u * j, b_desc + 2u * j, idesc, (acc | j) ? 1u : 0u);
      if (kk2) {          // second K block: next A atom, next ring slot
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_bf16(d_addr, a_desc + 1024u + 2u * j, b_desc + 1024u + 2u * j, idesc, 1u);
      }
      if (tl != nullptr) t3 = clock64();
      mma_commit(sbase + kSmBars + (B_EMPTY0 + slot) * 8);
      if (two) mma_commit(sbase + kSmBars + (B_EMPTY0 + slot + 1) * 8);
      if (c0 != kNone) mma_commit(sbase + kSmBars + c0 * 8);
      if (c1 != kNone) mma_commit(sbase + kSmBars + c1 * 8);
    }
    if (tl != nullptr) {
      const long long t4 = clock64();
      tl[0] = t0; tl[1] = t1; tl[2] = t2; tl[3] = t3; tl[4] = t4;
      if (flags & 6u) { tl[5] = tl[6] = tl[7] = tl[8] = tl[9] = t4; }
    }
  }
  __syncwarp();
  g += two ? 2 : 1;
  return (uint64_t)g | ((uint64_t)phases << 32);
}

// Per-thread description of the embedding-input task it owns for the whole tile: one (row, atom).
struct EmbedTask {
  const float* src;      // obs atom: state / goal vector of this row (nullptr = zeros)
  bool is_goal;          // src is a goal vector (rollout scaling zeroes goal dimensions)
  int row, atom, vs, tok, xoff;   // xoff >= 0: action row, offset of its act values in the tile's x buffer
  int part;              // PREC: 0 = fp16 hi image of the row, 1 = lo image, 2 = both (P128); -1 = bf16 (fp16-mode embedding GEMM)
  bool valid;
};
// Sequence row -> (virtual sequence of the tile, token).  LAY 2 (P128): rows [0, 64) hold the first S / 2 sequences,
// rows [64, 128) the others (a sequence never straddles row 64); rows beyond them get vs = S (invalid).
template <int LAY>
__device__ __forceinline__ void row_to_seq(int srow, int T, int S, int& vs, int& tok) {
  if constexpr (LAY == 2) {
    const int Sh = S >> 1, r = srow & 63, q = r / T;
    tok = r - q * T;
    vs = q < Sh ? (srow >> 6) * Sh + q : S;
  } else {
    vs = srow / T;
    tok = srow - vs * T;
  }
}
// LAY: 0 = fp16 mode (bf16 embedding operand), 1 = precise, stacked hi / lo operand rows, 2 = precise, P128 (the thread
// writes both images of its row)
template <int LAY>
__device__ EmbedTask make_embed_task(const Compute& c, const FastParams& p, int tile) {
  const bool cfg = (p.flags & BESO_FLAG_CFG) != 0;
  EmbedTask e;
  e.row = c.ctid & (kRows - 1);
  e.atom = c.ctid >> 7;
  const int srow = LAY == 1 ? ((e.row >> 5) << 4) | (e.row & 15) : e.row;   // sequence row of this operand row
  e.part = LAY == 1 ? (e.row >> 4) & 1 : (LAY == 2 ? 2 : -1);
  row_to_seq<LAY>(srow, p.T, p.S, e.vs, e.tok);
  const int ls = cfg ? (e.vs >> 1) : e.vs;
  const int seq = tile * (cfg ? p.S / 2 : p.S) + ls;
  e.valid = e.vs < p.S && seq < p.B;
  e.src = nullptr;
  e.is_goal = false;
  e.xoff = -1;
  if (e.valid) {
    const bool uncond = cfg ? ((e.vs & 1) != 0) : ((p.flags & BESO_FLAG_UNCOND) != 0);
    const int j = e.tok - 1 - p.G;
    if (e.tok >= 1 && e.tok <= p.G) { e.is_goal = true; if (!uncond) e.src = p.goal + ((size_t)seq * p.G + (e.tok - 1)) * p.obs; }
    else if (j >= 0 && (j & 1) == 0) e.src = p.state + ((size_t)seq * p.t + (j >> 1)) * p.obs;
    else if (j >= 0) e.xoff = (ls * p.t + (j >> 1)) * p.act;
  }
  return e;
}

// A <- embedding-GEMM input rows: [obs atom | misc atom] (see file header), atoms 2..3 untouched.
// One (row, atom) per thread; all global loads of a row are issued before any is used.
template <int LAY>
__device__ __noinline__ void build_embed_input(const Compute c, const EmbedTask e, int obs, int act, uint32_t flags,
                                               float sigma_data, const float* xsrc, const float* sigv,
                                               const float* in_tab, const float* goal_keep) {
  constexpr bool PREC = LAY != 0;
  constexpr uint32_t kLoAtom = 32768;              // P128: the lo image of A atom a is atom a + 2
  uint8_t* atom = c.sm + kSmA + e.atom * 16384;
  if (e.atom == 0) {
    float v[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = 0.f;
    if (e.src != nullptr) {
      if ((obs & 3) == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (i * 4 < obs) {
            const float4 f = __ldg(reinterpret_cast<const float4*>(e.src) + i);
            v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) if (i < obs) v[i] = __ldg(e.src + i);
      }
      if (in_tab != nullptr) {                       // scale_input of the rollout path, fused (uniform branch)
#pragma unroll
        for (int i = 0; i < 64; ++i) if (i < obs) v[i] = io_scale(v[i], in_tab, obs, i);
      }
      if (goal_keep != nullptr && e.is_goal) {
#pragma unroll
        for (int i = 0; i < 64; ++i) if (i < obs) v[i] = __fmul_rn(v[i], __ldg(goal_keep + i));
      }
    }
    if constexpr (LAY == 2) {
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) st_chunk_h(atom, e.row, ch, v + ch * 8);
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] -= __half2float(__float2half_rn(v[i]));
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) st_chunk_h(atom + kLoAtom, e.row, ch, v + ch * 8);
    } else if constexpr (PREC) {
#pragma unroll
      for (int i = 0; i < 64; ++i) if (e.part) v[i] -= __half2float(__float2half_rn(v[i]));
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) st_chunk_h(atom, e.row, ch, v + ch * 8);
    } else {
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) st_chunk(atom, e.row, ch, v + ch * 8);
    }
  } else {
    const bool inner = (flags & BESO_FLAG_INNER) != 0;
    const float sg = e.valid ? sigv[e.vs] : 1.0f;
    const float c_in = inner ? 1.0f : 1.0f / sqrtf(sg * sg + sigma_data * sigma_data);
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
            if (e.xoff >= 0 && k < act) x = xsrc[e.xoff + k] * c_in;
            else if (PREC) { if (e.tok == 0 && k == act) x = cn; }              // split below: [cn] . [sigma_emb.w]
            else if (e.tok == 0 && k >= act && k < act + 3) x = (k == act + 1) ? (cn - cn_hi) : cn_hi;
          } else if (k == hot || (k == hot + 1 && !PREC)) {
            x = 1.0f;
          }
        }
        if (PREC && e.part == 1) x -= __half2float(__float2half_rn(x));
        v[i] = x;
      }
      if constexpr (PREC) st_chunk_h(atom, e.row, ch, v); else st_chunk(atom, e.row, ch, v);
      if constexpr (LAY == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] -= __half2float(__float2half_rn(v[i]));
        st_chunk_h(atom + kLoAtom, e.row, ch, v);
      }
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

// ---- cold parts of the compute role, out of line: the layer loop then keeps only a handful of values live
// across its calls, which leaves the register file to the hot drains (ptxas interleaves their dependent chains
// only as far as registers allow).  Everything is re-derived from the kernel parameters.
constexpr uint32_t kSmTlCursor = kSmBars + 400;              // DBG: timeline cursor of compute thread 0
template <bool DBG>
__device__ __forceinline__ void stamp(const Compute& c) {
  if constexpr (DBG) {
    if (c.ctid == 0) {
      long long** cur = reinterpret_cast<long long**>(c.sm + kSmTlCursor);
      if (*cur != nullptr) { **cur = clock64(); ++*cur; }
    }
  }
}
struct XBufs { float *xcur, *d1, *x2, *dU, *sigv; };
template <class G>
__device__ __forceinline__ XBufs xbufs(uint8_t* sm, const FastParams& p) {
  float* sigv = reinterpret_cast<float*>(sm + G::SmSigv);
  float* xbuf;
  if constexpr (G::XGlobal) xbuf = p.xscratch + (size_t)blockIdx.x * (4 * kXFloats);   // no shared memory left for them
  else xbuf = reinterpret_cast<float*>(sm + G::SmProg + 1024);
  return {xbuf, xbuf + kXFloats, xbuf + 2 * kXFloats, xbuf + 3 * kXFloats, sigv};
}

// x of this tile's sequences -> shared memory (kept there across all sampler steps)
template <class G>
__device__ __noinline__ void tile_begin(const Compute c, const FastParams& p, int tile) {
  const bool cfg = (p.flags & BESO_FLAG_CFG) != 0;
  const int nls = cfg ? p.S / 2 : p.S, n_x = nls * p.t * p.act;
  const int seq0 = tile * nls, ns = max(0, min(nls, p.B - seq0));
  const XBufs xb = xbufs<G>(c.sm, p);
  compute_sync();                                            // previous tile's x fully written out
  for (int i = c.ctid; i < n_x; i += kComputeThreads)
    xb.xcur[i] = (i < ns * p.t * p.act) ? p.xin[(size_t)seq0 * p.t * p.act + i] : 0.f;
}
template <class G>
__device__ __noinline__ void tile_end(const Compute c, const FastParams& p, const SampleArgs& sa, int tile) {
  const bool cfg = (p.flags & BESO_FLAG_CFG) != 0;
  const int nls = cfg ? p.S / 2 : p.S;
  const int seq0 = tile * nls, ns = max(0, min(nls, p.B - seq0));
  const XBufs xb = xbufs<G>(c.sm, p);
  compute_sync();
  for (int i = c.ctid; i < ns * p.t * p.act; i += kComputeThreads) {
    float v = xb.xcur[i];
    if (sa.clip) v = io_clip(v, sa.clip, p.act, i % p.act);             // clip_action, fused
    p.out[(size_t)seq0 * p.t * p.act + i] = v;
    if (sa.unscaled) sa.unscaled[(size_t)seq0 * p.t * p.act + i] = sa.out_tab ? io_scale(v, sa.out_tab, p.act, i % p.act) : v;
  }
}

// noise levels of this evaluation + the embedding-GEMM A operand
template <class G, bool DBG, bool PREC>
__device__ __noinline__ void eval_prologue(const Compute c, const FastParams& p, const SampleArgs& sa, int tile, int step, int second) {
  const bool cfg = (p.flags & BESO_FLAG_CFG) != 0;
  const int nls = cfg ? p.S / 2 : p.S, seq0 = tile * nls;
  const XBufs xb = xbufs<G>(c.sm, p);
  const float s_hat = sa.n_steps ? sa.sig[step] : 0.f;
  const float s_next = sa.n_steps ? sa.sig[step + 1] : 0.f;
  const float s_eval = second ? (sa.sampler == BESO_SAMPLER_TWO_STAGE ? sa.sigb[step] : s_next) : s_hat;
  for (int i = c.ctid; i < p.S; i += kComputeThreads) {
    const int ls = cfg ? (i >> 1) : i;
    xb.sigv[i] = sa.n_steps ? s_eval : ((seq0 + ls < p.B) ? __ldg(p.sigma + seq0 + ls) : 1.0f);
  }
  compute_sync();
  constexpr int LAY = !PREC ? 0 : (std::is_same_v<G, G256P> ? 2 : 1);
  const EmbedTask etask = make_embed_task<LAY>(c, p, tile);
  stamp<DBG>(c);
  build_embed_input<LAY>(c, etask, p.obs, p.act, p.flags, p.sigma_data, second ? xb.x2 : xb.xcur, xb.sigv, sa.in_tab, sa.goal_keep);
  stamp<DBG>(c);
}

// ln_f + action head read-out + pre-conditioning (+ CFG mix) + sampler update of x.  Returns the barrier phases.
template <class G, bool DBG, bool PREC>
__device__ __noinline__ uint32_t eval_epilogue(Compute c, const FastParams& p, const SampleArgs& sa, int tile, int step, int second,
                                               float* trace_row) {
  const bool cfg = (p.flags & BESO_FLAG_CFG) != 0;
  const bool inner = (p.flags & BESO_FLAG_INNER) != 0;
  const int nls = cfg ? p.S / 2 : p.S, seq0 = tile * nls;
  const int ns = max(0, min(nls, p.B - seq0));
  const XBufs xb = xbufs<G>(c.sm, p);
  float* xcur = xb.xcur; float* d1 = xb.d1; float* x2 = xb.x2; float* dU = xb.dU;
  const float* xsrc = second ? xb.x2 : xb.xcur;
  const float s_hat = sa.n_steps ? sa.sig[step] : 0.f;
  const float s_next = sa.n_steps ? sa.sig[step + 1] : 0.f;
  const float* vecA = reinterpret_cast<const float*>(c.sm + G::SmVecA);
  cp_async_wait<1>();                                 // final vecA block (vecM(0) may still fly)
  compute_sync();
  c.wait(B_X_DONE);
  tc_fence_after();
  stamp<DBG>(c);
  constexpr bool P128 = std::is_same_v<G, G256P>;
  constexpr int LAY = !PREC ? 0 : (P128 ? 2 : 1);
  if constexpr (P128) {
    ln_pass_p128<DBG>(c, c.sbase + G::SmVecA, p.inv_d, p.d_true, trace_row);
  } else if constexpr (G::DP == 384) {
    if constexpr (PREC) ln_pass_pw<DBG>(c, c.sbase + G::SmVecA, p.inv_d, p.d_true, trace_row);
    else ln_pass_w<DBG>(c, c.sbase + G::SmVecA, p.inv_d, trace_row);
  } else {
    if constexpr (PREC) ln_pass_p<DBG>(c, c.sbase + G::SmVecA, p.inv_d, p.d_true, trace_row);
    else ln_pass<DBG>(c, c.sbase + G::SmVecA, p.inv_d, trace_row);
  }
  stamp<DBG>(c);
  c.wait(B_ACC_FULL0);
  tc_fence_after();
  stamp<DBG>(c);
  float pr[16];
  tmem_ld16(c.lane_addr(G::ColS0), pr);
  tmem_wait_ld();
  tc_fence_before();
  c.arrive(B_ACC_EMPTY0);
  if constexpr (LAY == 1) {                         // hi lane + lo lane of the sequence row
#pragma unroll
    for (int a = 0; a < 16; ++a) pr[a] += shx16(pr[a]);
  }
  const float* hb = vecA + G::VecBq;
  int vs, tok;
  row_to_seq<LAY>(c.srow, p.T, p.S, vs, tok);
  const int j = tok - 1 - p.G;
  const int ls = cfg ? (vs >> 1) : vs;
  const bool act_row = (c.hf == 0) && !c.is_lo && vs < p.S && tok > p.G && (j & 1) && (ls < ns);
  const int xo = (ls * p.t + (j >> 1)) * p.act;
  // D = c_out * F + c_skip * x (score_wrappers.py:81-96) of action column a of this row, straight from the
  // accumulator registers (no per-thread array: it would live in local memory)
  float c_skip = 0.f, c_out = 1.f;
  if (act_row) {
    const float sg = xb.sigv[vs];
    const float den = sg * sg + p.sigma_data * p.sigma_data;
    c_skip = p.sigma_data * p.sigma_data / den;
    c_out = sg * p.sigma_data / sqrtf(den);
  }
  auto denoised = [&](float acc, int a) -> float {
    const float f = acc + hb[a];
    return inner ? f : __fadd_rn(__fmul_rn(f, c_out), __fmul_rn(xsrc[xo + a], c_skip));
  };
  if (cfg && act_row && (vs & 1)) {                 // unconditional branch rows park their result for the mix
#pragma unroll
    for (int a = 0; a < kMaxAct; ++a) if (a < p.act) dU[xo + a] = denoised(pr[a], a);
  }
  if (cfg) compute_sync();
  if (act_row && !(cfg && (vs & 1))) {
#pragma unroll
    for (int a = 0; a < kMaxAct; ++a) {
      if (a < p.act) {
        float D = denoised(pr[a], a);
        if (cfg) D = __fadd_rn(dU[xo + a], __fmul_rn(p.lambda, __fsub_rn(D, dU[xo + a])));
        const int i = xo + a;
        if (sa.n_steps == 0) {
          p.out[(size_t)seq0 * p.t * p.act + i] = D;
        } else if (sa.sampler == BESO_SAMPLER_DDIM) {
          xcur[i] = __fsub_rn(__fmul_rn(sa.ca[step], xcur[i]), __fmul_rn(sa.ce[step], D));
        } else if (sa.sampler == BESO_SAMPLER_LMS) {                  // gc_sampling.py:454-465; history in d1, x2, dU
          const float d = __fdiv_rn(__fsub_rn(xcur[i], D), s_hat);
          float acc = __fmul_rn(sa.ca[step], d);
          if (sa.ce[step] != 0.0f) acc = __fadd_rn(acc, __fmul_rn(sa.ce[step], d1[i]));
          if (sa.c1[step] != 0.0f) acc = __fadd_rn(acc, __fmul_rn(sa.c1[step], x2[i]));
          if (sa.c2[step] != 0.0f) acc = __fadd_rn(acc, __fmul_rn(sa.c2[step], dU[i]));
          xcur[i] = __fadd_rn(xcur[i], acc);
          dU[i] = x2[i]; x2[i] = d1[i]; d1[i] = d;
        } else if (sa.sampler == BESO_SAMPLER_TWO_STAGE) {            // coefficient program (include/beso_b200.h)
          const float su = sa.su[step];
          const float nz = su != 0.0f ? __ldg(sa.noise + (size_t)step * sa.noise_stride + (size_t)seq0 * p.t * p.act + i) : 0.f;
          if (!second) {
            const float u = fmaf(sa.ca[step], xcur[i], sa.ce[step] * D);
            if (sa.sigb[step] == 0.0f) xcur[i] = fmaf(su, nz, u); else x2[i] = u;
          } else {
            xcur[i] = fmaf(su, nz, fmaf(sa.c1[step], xcur[i], fmaf(sa.c2[step], x2[i], sa.c3[step] * D)));
          }
        } else if (sa.sampler == BESO_SAMPLER_DPMPP_2M) {             // gc_sampling.py:726-735; d1 keeps old_denoised
          const float c2 = sa.c2[step];
          const float dd = (c2 != 0.0f) ? __fsub_rn(__fmul_rn(sa.c1[step], D), __fmul_rn(c2, d1[i])) : D;
          xcur[i] = __fsub_rn(__fmul_rn(sa.ca[step], xcur[i]), __fmul_rn(sa.ce[step], dd));
          d1[i] = D;
        } else if (sa.sampler == BESO_SAMPLER_EULER_ANCESTRAL) {      // gc_sampling.py:216-256
          const float s_down = sa.ca[step];
          const float dd = __fdiv_rn(__fsub_rn(xcur[i], D), s_hat);
          float xe = __fadd_rn(xcur[i], __fmul_rn(dd, __fsub_rn(s_down, s_hat)));
          if (s_down > 0.0f)
            xe = __fadd_rn(xe, __fmul_rn(__ldg(sa.noise + (size_t)step * sa.noise_stride + (size_t)seq0 * p.t * p.act + i), sa.ce[step]));
          xcur[i] = xe;
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
  // vecA is free again: first block of the next evaluation
  compute_sync();
  load_vec_async(c, G::SmVecA, p.vec, G::VecAFloats);
  return c.phases;
}

// CG = 1: every CTA issues its own M = 128 MMAs.  CG = 2: CTA pairs (cluster of 2) run cta_group::2
// M = 256 MMAs issued by the leader CTA; each CTA streams and holds only HALF of every B operand, which
// halves the L2 -> SMEM weight traffic and the shared-memory bandwidth the tensor core needs for B.
// DBG = true compiles in the diagnostics (per-phase clock stamps, LayerNorm trace dump); production is DBG = false.
// MC = 2 (with CG = 1): the two CTAs of a cluster run independent tiles but share ONE weight stream: each
// producer fetches half of every ring stage and multicasts it into both CTAs (TMA .multicast::cluster), a stage
// is free again when both CTAs' MMAs have read it (multicast commits).  Halves the L2 -> SM request traffic,
// which at full-chip scale is within a factor 1.5 of the L2 throughput cap.
// PREC = true: fp32-equivalent mode (split operands, 64 sequence rows per tile; see the PREC section above).
// GEO = 0 / 1 / 2: geometry G256 / G384 / G256P above (G256P is the precise mode on full tiles, "P128").
template <int CG, bool DBG, int MC, bool PREC, int HSP, int GEO>
__global__ void __launch_bounds__(kThreads, 1)
fast_sample_kernel(const __grid_constant__ FastParams p, const __grid_constant__ SampleArgs sa) {
  static_assert(!PREC || (CG == 1 && MC == 1), "the precise mode runs single-CTA MMAs");
  static_assert(GEO == 0 || (CG == 1 && MC == 1), "the single-accumulator geometries run single-CTA MMAs");
  static_assert(GEO != 2 || PREC, "G256P is a precise-mode geometry");
  using G = std::conditional_t<GEO == 1, G384, std::conditional_t<GEO == 2, G256P, G256>>;
  constexpr bool WIDE = GEO != 0;                    // single scratch accumulator, 16 KB ring groups
  constexpr bool P128 = GEO == 2;
  constexpr uint32_t NB = G::DP / 128;               // 128-column blocks of X
  extern __shared__ uint8_t smem_raw[];
  // dynamic shared memory is at least 16-byte aligned; the operand tiles need 1024
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kSmBars + B_COUNT * 8);
  constexpr int PAIR = CG > MC ? CG : MC;                    // CTAs per cluster
  const uint32_t rank = (PAIR == 2) ? cluster_ctarank() : 0u;

  if (threadIdx.x == 0) {
    for (int i = 0; i < B_COUNT; ++i) {
      const bool by_warps = (i == B_A_READY || i == B_ACC_EMPTY0 || i == B_ACC_EMPTY1 || i == B_OP_READY0 || i == B_OP_READY1 ||
                             i == B_OP_READY0B || i == B_OP_READY1B);
      const bool ring_empty = i >= B_EMPTY0 && i < B_EMPTY0 + 4;
      mbar_init(sbase + kSmBars + i * 8, i == B_Y_READY ? kAttnWarps * CG : (by_warps ? 8 * CG : (ring_empty ? MC : 1)));
    }
    fence_barrier_init();
  }
  if (!WIDE && threadIdx.x == kMmaWarp * 32) {
    // group table for [embedding | one layer | head]: the single-warp roles replay it with one LDS per group
    uint4* tab = reinterpret_cast<uint4*>(sm + kSmProg);
    int i = 0;
    walk_eval(1, [&](const Group& q) {
      tab[i++] = make_uint4(q.a_off, q.d_col | (q.n << 16), q.acc | (q.pair ? 2u : 0u) | (q.kk2 ? 4u : 0u) | (q.bf16 ? 8u : 0u),
                            q.w0 | (q.w1 << 4) | (q.c0 << 8) | (q.c1 << 12));
    });
  }
  if (PAIR == 2) cluster_sync_all();        // barriers of both CTAs initialised before any remote arrive / multicast
  if (warp == kMmaWarp) { if (CG == 2) tmem_alloc_cg2(smem_u32(tmem_slot), 512); else tmem_alloc(smem_u32(tmem_slot), 512); }
  for (uint32_t i = threadIdx.x; i < (G::SmVecA - G::SmU) / 16; i += kThreads)      // padding rows must stay finite
    reinterpret_cast<uint4*>(sm + G::SmU)[i] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // tiles are dealt to MMA groups (CTA or CTA pair) round-robin; the second CTA of a pair may get a dummy tile
  const int n_groups = (int)gridDim.x / PAIR, group = (int)blockIdx.x / PAIR;
  const int n_group_tiles = (p.n_tiles + PAIR - 1) / PAIR;
  const int my_tiles = (n_group_tiles - group + n_groups - 1) / n_groups;

  const uint4* gtab = reinterpret_cast<const uint4*>(sm + kSmProg);
  const int n_groups_eval = 6 + 52 * p.L;
  // group i of an evaluation -> table entry: [0,2) embedding, then 52 per layer (replayed), then 4 head
  auto table_index = [&](int i, int& li) -> int {
    if (i < 2) return i;
    if (i >= n_groups_eval - 4) return 54 + (i - (n_groups_eval - 4));
    const int r = 2 + li;
    li = li == 51 ? 0 : li + 1;
    return r;
  };
  if (warp == kProducerWarp) {
    // ======================= weight-tape producer =======================
    uint32_t g = 0;
    if constexpr (WIDE) {
      // 16 KB slots, every ring group is one slot (RingW); the tape is contiguous in consumption order
      RingW ring{0u, 0u};
      constexpr uint32_t kImages = PREC ? 2u : 1u;
      // per pass: [Q|K] NA, V NA / 2, proj NB groups; per hidden chunk: FC1 NA, FC2 2 NB
      const uint32_t per_layer_attn = (uint32_t)(G::NA + G::NA / 2 + NB) * (uint32_t)p.npass, per_layer_mlp = (uint32_t)(G::NA + 2 * NB) * (uint32_t)G::NCH;
      for (int it = 0; it < my_tiles * p.evals; ++it) {
        const uint8_t* src = p.tape;
        auto fills = [&](uint32_t n, uint32_t bytes, uint32_t n_slots) {
#pragma unroll 1
          for (uint32_t i = 0; i < n * kImages; ++i) {
            const uint32_t slot = ring.begin(n_slots), par = ring.parity(slot);
            const uint32_t full = sbase + kSmBars + (B_FULL0 + slot) * 8, dst = sbase + RingW::slot_off<G>(slot);
            spin_wait(sbase + kSmBars + (B_WEMPTY0 + slot) * 8, par ^ 1u);
            if (elect_one()) {
              mbar_expect_tx(full, bytes);
              const uint32_t h = bytes >> 1;                 // two requests in flight per group
              bulk_g2s(dst, src, h, full);
              bulk_g2s(dst + h, src + h, h, full);
            }
            __syncwarp();
            ring.end(slot);
            src += bytes;
          }
        };
        fills(2 * NB, 16384, kWSlots);                       // embedding: 2 K atoms x NB column blocks
#pragma unroll 1
        for (int l = 0; l < p.L; ++l) {
          fills(per_layer_attn, 16384, kWSlots);
          if constexpr (G::MlpSlots > kWSlots) {
            // The extra slot overlays the Y atoms: wait for this layer's attention half to be over (X_DONE completion
            // number 1 + 2 l of this evaluation; 1 + 2 L completions per evaluation).  The wait cannot be a full phase
            // early: the issuer is at most a ring of groups behind, i.e. well inside this layer's attention half, which
            // only began after the previous completion.  It costs nothing: LayerNorm 2 runs before the first FC1 group.
            const uint32_t n = (uint32_t)it * (1u + 2u * (uint32_t)p.L) + 1u + 2u * (uint32_t)l;
            spin_wait(sbase + kSmBars + B_X_DONE * 8, n & 1u);
          }
          fills(per_layer_mlp, 16384, G::MlpSlots);
        }
        fills(1, 2048 * G::NA, kWSlots);                     // action head: NA K blocks of [16 x 64]
      }
    } else if constexpr (CG == 1) {
      // Two stages of 32 KB (slots {0,1} and {2,3}), one full / empty barrier pair per stage; the tape is
      // contiguous in consumption order, so the producer only needs the byte count of each ring group.
      auto fill = [&](const uint8_t* src, uint32_t bytes) {
        const uint32_t slot = g & 2u, par = (g >> 2) & 1u;
        const uint32_t full = sbase + kSmBars + (B_FULL0 + slot) * 8, dst = sbase + kSmRing + slot * kSlotBytes;
        spin_wait(sbase + kSmBars + (B_EMPTY0 + slot) * 8, par ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(full, bytes);
          const uint32_t h = bytes >> 1;                     // two requests in flight per group
          if constexpr (MC == 2) {                           // this CTA's half, delivered to both CTAs
            const uint32_t o = rank * h, q = h >> 1;
            bulk_g2s_multicast(dst + o, src + o, q, full, 3);
            bulk_g2s_multicast(dst + o + q, src + o + q, q, full, 3);
          } else {
            bulk_g2s(dst, src, h, full);
            bulk_g2s(dst + h, src + h, h, full);
          }
        }
        __syncwarp();
        g += 2;
      };
      // PREC: every ring group is followed by the lo image of the same weights
      constexpr uint32_t kImages = PREC ? 2u : 1u;
      for (int it = 0; it < my_tiles * p.evals; ++it) {
        const uint8_t* src = p.tape;
        auto fills = [&](uint32_t n, uint32_t bytes) {
#pragma unroll 1
          for (uint32_t i = 0; i < n * kImages; ++i) { fill(src, bytes); src += bytes; }
        };
        fills(2, 32768);                                     // embedding
#pragma unroll 1
        for (int l = 0; l < p.L; ++l) {
#pragma unroll 1
          for (int h = 0; h <= p.npass; ++h) {               // Q0 Q1 P0 Q2 P1 ... P(n-1): QKV = 4 groups of 24 KB, proj = 32 KB
            if (h < p.npass) fills(4, 24576);
            if (h >= 1) fills(1, 32768);
          }
          fills(32, 32768);                                  // FC1 / FC2 groups
        }
        fills(1, 8192);                                      // action head: 4 K blocks of [16 x 64]
      }
    } else {
    for (int it = 0; it < my_tiles * p.evals; ++it) {
      uint32_t off = 0;
      int li = 0;
#pragma unroll 1
      for (int i = 0; i < n_groups_eval; ++i) {
        const uint4 e = gtab[table_index(i, li)];
        const uint32_t n = e.y >> 16, kk2 = (e.z >> 2) & 1u;
        g = producer_step<CG>(n, (e.z >> 1) & 1u, kk2, sbase, rank, p.tape + off, g);
        off += n * 128u * (kk2 ? 2u : 1u);
      }
    }
    }
  } else if (warp == kMmaWarp) {
    // ======================= MMA issuer (leader) / full-barrier forwarder (peer CTA of a pair) ==========
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    uint32_t g = 0, phases = (1u << B_ACC_EMPTY0) | (1u << B_ACC_EMPTY1);   // "empty" barriers pass the first time
    if constexpr (CG == 1) {
      // ---- straight-line schedule: shapes, operand offsets and barrier ids are compile-time constants, the
      // whole warp runs the (uniform) control flow and one elected lane issues tcgen05.mma / commit.  The issuer
      // has to stay ahead of the tensor pipe: ~50 instructions per ring group instead of a table walk.
      const uint32_t bars = sbase + kSmBars;
      const uint32_t dlo = ((sbase >> 4) & 0x3FFFu) | (1u << 16);        // descriptor low word of shared offset 0
      constexpr uint32_t dhi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
      auto desc = [&](uint32_t lo) -> uint64_t { return (uint64_t)lo | ((uint64_t)dhi << 32); };
      long long* tl = nullptr;
      long long t_bw = 0, t_rw = 0;
      auto job_begin = [&]() { if constexpr (DBG) { if (tl != nullptr) { tl[0] = clock64(); t_bw = t_rw = 0; } } };
      auto job_end = [&]() { if constexpr (DBG) { if (tl != nullptr) { tl[1] = t_bw; tl[2] = t_rw; tl[3] = clock64(); tl += 4; } } };
      auto jwait = [&](uint32_t id) {                                    // a compute -> MMA barrier
        const uint32_t par = (phases >> id) & 1u;
        phases ^= 1u << id;
        long long t = 0;
        if constexpr (DBG) t = clock64();
        spin_wait(bars + id * 8, par);
        if constexpr (DBG) t_bw += clock64() - t;
      };
      auto jcommit = [&](uint32_t id) { if (elect_one()) mma_commit(bars + id * 8); __syncwarp(); };
      // one ring group (stage of up to 32 KB): NKB K blocks of 64; A atoms 16 KB apart from a_off, B sub-tiles
      // B_STEP bytes apart in the stage
      auto group = [&](auto n_tag, auto nkb_tag, auto bstep_tag, auto bf16_tag, uint32_t a_off, uint32_t d_col, uint32_t acc) {
        constexpr uint32_t N = decltype(n_tag)::value, NKB = decltype(nkb_tag)::value, B_STEP = decltype(bstep_tag)::value;
        constexpr uint32_t idesc = decltype(bf16_tag)::value ? idesc_bf16_m128(N) : idesc_f16_m128(N);
        const uint32_t slot = g & 2u, par = (g >> 2) & 1u;
        long long t = 0;
        if constexpr (DBG) t = clock64();
        spin_wait(bars + (B_FULL0 + slot) * 8, par);
        if constexpr (DBG) t_rw += clock64() - t;
        tc_fence_after();
        const uint32_t a_lo = dlo + (a_off >> 4), b_lo = dlo + ((kSmRing + slot * kSlotBytes) >> 4);
        if (elect_one()) {
#pragma unroll
          for (uint32_t kb = 0; kb < NKB; ++kb)
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j)
              mma_bf16(tm + d_col, desc(a_lo + kb * 1024u + 2u * j), desc(b_lo + kb * (B_STEP >> 4) + 2u * j), idesc,
                       (acc | kb | j) ? 1u : 0u);
          if constexpr (MC == 2) mma_commit_multicast(bars + (B_EMPTY0 + slot) * 8, 3);
          else mma_commit(bars + (B_EMPTY0 + slot) * 8);
        }
        __syncwarp();
        g += 2;
      };
      // PREC: the hi image of the weights, then the lo image, into the same accumulator
      auto group2 = [&](auto n_tag, auto nkb_tag, auto bstep_tag, auto bf16_tag, uint32_t a_off, uint32_t d_col, uint32_t acc) {
        group(n_tag, nkb_tag, bstep_tag, bf16_tag, a_off, d_col, acc);
        if constexpr (PREC) group(n_tag, nkb_tag, bstep_tag, bf16_tag, a_off, d_col, 1);
      };
      using std::integral_constant;
      constexpr uint32_t kEmbBf16 = PREC ? 0u : 1u;          // the precise mode splits the raw inputs into fp16 hi + lo
#define BESO_IC(v) integral_constant<uint32_t, (v)>{}
      if constexpr (WIDE) {
      // ---- G384: one ring group = one 16 KB slot; every GEMM into X is three N = 128 column blocks ----
      RingW ring{0u, 0u};
      uint32_t n_slots = kWSlots;
      // One ring group: D (+)= A W^T, then (A2 != 0) D += A2 W^T with the same weights.  A2 = 1: second A operand in
      // shared memory (byte offset a2), A2 = 2: in tensor memory (column a2; one K atom = 32 columns).
      auto groupw = [&](auto n_tag, auto nkb_tag, auto bstep_tag, auto bf16_tag, auto a2_tag, uint32_t a_off, uint32_t a2, uint32_t d_col, uint32_t acc) {
        constexpr uint32_t N = decltype(n_tag)::value, NKB = decltype(nkb_tag)::value, B_STEP = decltype(bstep_tag)::value;
        constexpr uint32_t idesc = decltype(bf16_tag)::value ? idesc_bf16_m128(N) : idesc_f16_m128(N);
        constexpr uint32_t A2 = decltype(a2_tag)::value;
        const uint32_t slot = ring.begin(n_slots), par = ring.parity(slot);
        long long t = 0;
        if constexpr (DBG) t = clock64();
        spin_wait(bars + (B_FULL0 + slot) * 8, par);
        if constexpr (DBG) t_rw += clock64() - t;
        tc_fence_after();
        const uint32_t a_lo = dlo + (a_off >> 4), b_lo = dlo + (RingW::slot_off<G>(slot) >> 4);
        if (elect_one()) {
#pragma unroll
          for (uint32_t kb = 0; kb < NKB; ++kb)
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j)
              mma_bf16(tm + d_col, desc(a_lo + kb * 1024u + 2u * j), desc(b_lo + kb * (B_STEP >> 4) + 2u * j), idesc,
                       (acc | kb | j) ? 1u : 0u);
          if constexpr (A2 == 1) {
            const uint32_t a2_lo = dlo + (a2 >> 4);
#pragma unroll
            for (uint32_t kb = 0; kb < NKB; ++kb)
#pragma unroll
              for (uint32_t j = 0; j < 4; ++j)
                mma_bf16(tm + d_col, desc(a2_lo + kb * 1024u + 2u * j), desc(b_lo + kb * (B_STEP >> 4) + 2u * j), idesc, 1u);
          }
          if constexpr (A2 == 2) {
#pragma unroll
            for (uint32_t kb = 0; kb < NKB; ++kb)
#pragma unroll
              for (uint32_t j = 0; j < 4; ++j)
                mma_f16_ts(tm + d_col, tm + a2 + kb * 32u + 8u * j, desc(b_lo + kb * (B_STEP >> 4) + 2u * j), idesc, 1u);
          }
          mma_commit(bars + (B_WEMPTY0 + slot) * 8);
        }
        __syncwarp();
        ring.end(slot);
      };
      // One GEMM group in the mode's arithmetic.  fp16 mode: A W^T.  Stacked precise mode: the hi image of the weights,
      // then the lo image (A holds both images of the rows).  P128: (A_hi + A_lo) W_hi^T, then A_hi W_lo^T.
      auto group2w = [&](auto n_tag, auto nkb_tag, auto bstep_tag, auto bf16_tag, auto a2_tag, uint32_t a_off, uint32_t a2, uint32_t d_col, uint32_t acc) {
        if constexpr (P128) {
          groupw(n_tag, nkb_tag, bstep_tag, bf16_tag, a2_tag, a_off, a2, d_col, acc);
          groupw(n_tag, nkb_tag, bstep_tag, bf16_tag, BESO_IC(0), a_off, 0u, d_col, 1);
        } else {
          groupw(n_tag, nkb_tag, bstep_tag, bf16_tag, BESO_IC(0), a_off, 0u, d_col, acc);
          if constexpr (PREC) groupw(n_tag, nkb_tag, bstep_tag, bf16_tag, BESO_IC(0), a_off, 0u, d_col, 1);
        }
      };
      // X (+)= A[a_off: one K atom] W^T, the output columns as NB groups of N = 128 (a2: shared-memory lo image of A)
      auto into_x = [&](auto bf16_tag, uint32_t a_off, uint32_t a2, uint32_t acc) {
#pragma unroll
        for (uint32_t nb = 0; nb < NB; ++nb) group2w(BESO_IC(128), BESO_IC(1), BESO_IC(0), bf16_tag, BESO_IC(1), a_off, a2, kColX + nb * 128u, acc);
      };
      for (int it = 0; it < my_tiles * p.evals; ++it) {
        if constexpr (DBG) tl = (p.timeline != nullptr && blockIdx.x == 0 && it == 1) ? p.timeline : nullptr;
        // ---- embedding GEMM: X = A_emb W_emb^T, K = 128 ----
        job_begin();
        jwait(B_A_READY);
        into_x(BESO_IC(kEmbBf16), kSmA, kSmA + 32768, 0);             // (P128: lo images of atoms 0, 1 are atoms 2, 3)
        into_x(BESO_IC(kEmbBf16), kSmA + 16384, kSmA + 49152, 1);
        jcommit(B_X_DONE);
        job_end();
#pragma unroll 1
        for (int l = 0; l < p.L; ++l) {
          // ---- attention half: QK0 V0 | QK1 V1 P0 | ... | P(n-1): the projection of pass h - 1 goes behind the V job
          // of pass h (which the compute warps are waiting for), it then runs under the attention of pass h ----
#pragma unroll 1
          for (int h = 0; h <= p.npass; ++h) {
            if (h < p.npass) {                               // [Q|K] of attention pass h -> S (128 columns)
              job_begin();
              jwait(B_ACC_EMPTY0);
              if (h == 0) jwait(B_A_READY);
#pragma unroll
              for (uint32_t kb = 0; kb < G::NA; ++kb)
                group2w(BESO_IC(128), BESO_IC(1), BESO_IC(0), BESO_IC(0), BESO_IC(2), kSmA + kb * 16384, G::ColAL + kb * 32u, G::ColS0, kb);
              jcommit(B_ACC_FULL0);
              job_end();
            }
            if (h < p.npass) {                               // V of pass h -> S[0:64) once [Q|K] has been drained
              job_begin();
              jwait(B_ACC_EMPTY0);
#pragma unroll
              for (uint32_t kp = 0; kp < G::NA / 2; ++kp)
                group2w(BESO_IC(64), BESO_IC(2), BESO_IC(8192), BESO_IC(0), BESO_IC(2), kSmA + kp * 32768, G::ColAL + kp * 64u, G::ColS0, kp);
              jcommit(B_ACC_FULL0);
              job_end();
            }
            if (h >= 1) {                                    // X += Y_{h-1} Wproj[:, h-1]^T
              job_begin();
              jwait(B_Y_READY);
              into_x(BESO_IC(0), G::SmY, G::SmYLo, 1);
              jcommit(B_Y_EMPTY);
              if (h == p.npass) jcommit(B_X_DONE);
              job_end();
            }
          }
          // ---- MLP half: F1_0 | F1_c F2_c-1 | F2_11; one accumulator, hidden chunk c goes to H buffer c & 1 ----
          n_slots = G::MlpSlots;
#pragma unroll 1
          for (int c = 0; c <= G::NCH; ++c) {
            if (c < G::NCH) {
              job_begin();
              jwait(B_ACC_EMPTY0);
              if (c == 0) jwait(B_A_READY);
#pragma unroll
              for (uint32_t kb = 0; kb < G::NA; ++kb)
                group2w(BESO_IC(128), BESO_IC(1), BESO_IC(0), BESO_IC(0), BESO_IC(2), kSmA + kb * 16384, G::ColAL + kb * 32u, G::ColS0, kb);
              jcommit(B_ACC_FULL0);
              job_end();
            }
            if (c >= 1) {                                    // X += H_{c-1} W2[:, chunk c-1]^T
              const uint32_t b = G::HSingle ? 0u : (uint32_t)(c - 1) & 1u;
              job_begin();
              jwait(B_OP_READY0 + b);
              into_x(BESO_IC(0), b ? G::SmH1 : G::SmH0, G::SmHLo, 1);
              jwait(B_OP_READY0B + b);
              into_x(BESO_IC(0), (b ? G::SmH1 : G::SmH0) + 16384, G::SmHLo + 16384, 1);
              jcommit(B_OP_EMPTY0 + b);
              if (c == G::NCH) jcommit(B_X_DONE);
              job_end();
            }
          }
          n_slots = kWSlots;
        }
        // ---- action head (N = 16, all K atoms in one group) ----
        job_begin();
        jwait(B_A_READY);
        jwait(B_ACC_EMPTY0);
        group2w(BESO_IC(16), BESO_IC(G::NA), BESO_IC(2048), BESO_IC(0), BESO_IC(2), kSmA, G::ColAL, G::ColS0, 0);
        jcommit(B_ACC_FULL0);
        job_end();
      }
      } else
      for (int it = 0; it < my_tiles * p.evals; ++it) {
        if constexpr (DBG) tl = (p.timeline != nullptr && blockIdx.x == 0 && it == 1) ? p.timeline : nullptr;
        // ---- embedding GEMM (bf16): X = A_emb W_emb^T, K = 128 ----
        job_begin();
        jwait(B_A_READY);
        group2(BESO_IC(256), BESO_IC(1), BESO_IC(0), BESO_IC(kEmbBf16), kSmA, kColX, 0);
        group2(BESO_IC(256), BESO_IC(1), BESO_IC(0), BESO_IC(kEmbBf16), kSmA + 16384, kColX, 1);
        jcommit(B_X_DONE);
        job_end();
#pragma unroll 1
        for (int l = 0; l < p.L; ++l) {
          // ---- attention half: Q0 Q1 P0 Q2 P1 Q3 P2 P3 ----
#pragma unroll 1
          for (int h = 0; h <= p.npass; ++h) {
            if (h < p.npass) {                               // [Q|K|V] of attention pass h -> S0 (192 columns)
              job_begin();
              jwait(B_ACC_EMPTY0);
              if (h == 0) jwait(B_A_READY);
#pragma unroll
              for (uint32_t kb = 0; kb < 4; ++kb)
                group2(BESO_IC(192), BESO_IC(1), BESO_IC(0), BESO_IC(0), kSmA + kb * 16384, kColS0, kb);
              jcommit(B_ACC_FULL0);
              job_end();
            }
            if (h >= 1) {                                    // X += Y_{h-1} Wproj[:, h-1]^T
              job_begin();
              jwait(B_Y_READY);
              group2(BESO_IC(256), BESO_IC(1), BESO_IC(0), BESO_IC(0), kSmY, kColX, 1);
              jcommit(B_Y_EMPTY);
              if (h == p.npass) jcommit(B_X_DONE);
              job_end();
            }
          }
          // ---- MLP half: F1_0 F1_1 F2_0 (F1_c F2_c-1) F2_7; hidden chunk c lives in accumulator / H buffer c & 1 ----
#pragma unroll 1
          for (int c = 0; c <= 8; ++c) {
            if (c < 8) {
              const uint32_t b = c & 1;
              job_begin();
              jwait(B_ACC_EMPTY0 + b);
              if (c == 0) jwait(B_A_READY);
              group2(BESO_IC(128), BESO_IC(2), BESO_IC(16384), BESO_IC(0), kSmA, b ? kColS1 : kColS0, 0);
              group2(BESO_IC(128), BESO_IC(2), BESO_IC(16384), BESO_IC(0), kSmA + 32768, b ? kColS1 : kColS0, 1);
              jcommit(B_ACC_FULL0 + b);
              job_end();
            }
            if (c >= 1) {                                    // X += H_{c-1} W2[:, chunk c-1]^T
              const uint32_t b = (c - 1) & 1;
              job_begin();
              jwait(B_OP_READY0 + b);
              group2(BESO_IC(256), BESO_IC(1), BESO_IC(0), BESO_IC(0), b ? kSmH1 : kSmH0, kColX, 1);
              jwait(B_OP_READY0B + b);
              group2(BESO_IC(256), BESO_IC(1), BESO_IC(0), BESO_IC(0), (b ? kSmH1 : kSmH0) + 16384, kColX, 1);
              jcommit(B_OP_EMPTY0 + b);
              if (c == 8) jcommit(B_X_DONE);
              job_end();
            }
          }
        }
        // ---- action head (N = 16, K = 256: one 8 KB group) ----
        job_begin();
        jwait(B_A_READY);
        jwait(B_ACC_EMPTY0);
        group2(BESO_IC(16), BESO_IC(4), BESO_IC(2048), BESO_IC(0), kSmA, kColS0, 0);
        jcommit(B_ACC_FULL0);
        job_end();
      }
#undef BESO_IC
    } else
    for (int it = 0; it < my_tiles * p.evals; ++it) {
      if (CG == 2 && rank != 0) {
        // peer CTA: tell the leader when this CTA's half of each B operand has landed
#pragma unroll 1
        for (int i = 0; i < n_groups_eval; ++i) {
          const uint32_t slot = g & (kSlots - 1), par = (g >> 2) & 1u;
          spin_wait(sbase + kSmBars + (B_FULL0 + slot) * 8, par);
          if (elect_one()) mbar_arrive_cluster(sbase + kSmBars + (B_PFULL0 + slot) * 8, 0);
          __syncwarp();
          g += 1;
        }
      } else {
        long long* tl = (DBG && p.timeline != nullptr && blockIdx.x == 0 && it == 1) ? p.timeline : nullptr;
        int li = 0;
        // four groups per iteration, as straight-line code: consecutive groups then use different uniform
        // registers for their descriptors, so setting up group i+1 does not wait for group i's MMAs to issue
        auto run = [&](const uint4 c, auto tl_tag) {
          constexpr bool TLV = decltype(tl_tag)::value;
          const uint64_t r = mma_step<CG, TLV>(c.x, c.y & 0xFFFFu, c.y >> 16, c.z, c.w, sbase, tm, g, phases, tl);
          g = (uint32_t)r;
          phases = (uint32_t)(r >> 32);
          if (TLV) tl += (c.z & 6u) ? 10 : 5;
        };
        auto loop = [&](auto tl_tag) {
          int i = 0;
#pragma unroll 1
          for (; i + 4 <= n_groups_eval; i += 4) {
            const uint4 c0 = gtab[table_index(i, li)], c1 = gtab[table_index(i + 1, li)];
            const uint4 c2 = gtab[table_index(i + 2, li)], c3 = gtab[table_index(i + 3, li)];
            run(c0, tl_tag); run(c1, tl_tag); run(c2, tl_tag); run(c3, tl_tag);
          }
#pragma unroll 1
          for (; i < n_groups_eval; ++i) run(gtab[table_index(i, li)], tl_tag);
        };
        if constexpr (DBG) { if (tl != nullptr) loop(std::true_type{}); else loop(std::false_type{}); }
        else loop(std::false_type{});
      }
    }
  } else if (warp >= kHelperWarp0) {
    // ======================= attention helper warps (8, 9) =======================
    uint32_t y_phase = 1;
    for (int it = 0; it < my_tiles * p.evals * p.L * p.npass; ++it) {
      attn_sync();                                          // (P128: S1)
      // Y of the previous pass consumed by its projection MMAs: waited for here, or (single-accumulator schedules, where
      // that projection runs under this pass's attention) by the attention items right before they write Y
      uint32_t ybar = 0u;
      const uint32_t ypar = y_phase;
      if constexpr (WIDE) ybar = sbase + kSmBars + B_Y_EMPTY * 8;
      else spin_wait(sbase + kSmBars + B_Y_EMPTY * 8, y_phase);
      y_phase ^= 1u;
      if constexpr (P128) {                                 // the two 64-row halves (attention_pass_p128)
        attention_half_p128<HSP>(sbase, 4 + (warp - kHelperWarp0), 6, lane, p.S >> 1, p.T, 0, ybar, ypar);
        attn_sync();                                        // S2
        attn_sync();                                        // S3
        attention_half_p128<HSP>(sbase, 8 + (warp - kHelperWarp0), kAttnWarps, lane, p.S >> 1, p.T, 64, ybar, ypar);
      } else if constexpr (PREC) attention_head_p<G, HSP>(sbase, 8 + (warp - kHelperWarp0), lane, p.S, p.T, ybar, ypar);
      else attention_head<G, HSP>(sbase, 8 + (warp - kHelperWarp0), lane, p.S, p.T, ybar, ypar);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(sbase + kSmBars + B_Y_READY * 8, 0); else mbar_arrive(sbase + kSmBars + B_Y_READY * 8);
      }
      attn_sync();
    }
  } else {
    // ======================= compute warps =======================
    Compute c;
    c.sm = sm; c.sbase = sbase; c.tmem = tmem;
    c.ctid = threadIdx.x - kComputeWarp0 * 32;
    c.lane = lane; c.wq = warp & 3; c.hf = (warp - kComputeWarp0) >> 2;
    c.row = c.wq * 32 + lane;
    constexpr bool STACKED = PREC && !P128;                  // hi / lo images of a sequence row on two TMEM lanes
    c.is_lo = STACKED ? (lane >> 4) : 0;
    c.srow = STACKED ? c.wq * 16 + (lane & 15) : c.row;
    c.row_off = (uint32_t)c.row * 128u; c.rx4 = ((uint32_t)c.row & 7u) << 4;
    c.phases = (1u << B_OP_EMPTY0) | (1u << B_OP_EMPTY1) | (1u << B_Y_EMPTY);
    c.cg = CG;
    const uint32_t vecA_s = sbase + G::SmVecA, vecM_s = sbase + G::SmVecM;
    constexpr uint32_t layer_stride = G::VecAFloats + G::VecMFloats;
    load_vec_async(c, G::SmVecA, p.vec, G::VecAFloats);
    load_vec_async(c, G::SmVecM, p.vec + G::VecAFloats, G::VecMFloats);
    auto ln = [&](uint32_t vec_s, float* tr) {
      if constexpr (P128) {
        ln_pass_p128<DBG>(c, vec_s, p.inv_d, p.d_true, tr);
      } else if constexpr (WIDE) {
        if constexpr (PREC) ln_pass_pw<DBG>(c, vec_s, p.inv_d, p.d_true, tr);
        else ln_pass_w<DBG>(c, vec_s, p.inv_d, tr);
      } else {
        if constexpr (PREC) ln_pass_p<DBG>(c, vec_s, p.inv_d, p.d_true, tr);
        else ln_pass<DBG>(c, vec_s, p.inv_d, tr);
      }
    };
    if constexpr (DBG) { if (c.ctid == 0) *reinterpret_cast<long long**>(sm + kSmTlCursor) = nullptr; }

    for (int tj = 0; tj < my_tiles; ++tj) {
      const int tile = (group + tj * n_groups) * PAIR + (int)rank;   // >= n_tiles: dummy tile, protocol only
      tile_begin<G>(c, p, tile);
      int step = 0, second = 0;
      for (int ev = 0; ev < p.evals; ++ev) {
        auto trace_row = [&](int slot) -> float* {
          return (DBG && p.trace != nullptr && blockIdx.x == 0 && ev == 0 && tj == 0) ? p.trace + ((size_t)slot * kRows + c.srow) * G::DP : nullptr;
        };
        if constexpr (DBG) {
          if (c.ctid == 0)
            *reinterpret_cast<long long**>(sm + kSmTlCursor) =
                (p.timeline != nullptr && blockIdx.x == 0 && tile == 0 && ev == 1) ? p.timeline + 6 * p.n_fills : nullptr;
        }
        eval_prologue<G, DBG, PREC>(c, p, sa, tile, step, second);

        for (int l = 0; l < p.L; ++l) {
          // ---------------- attention half ----------------
          cp_async_wait<0>();
          compute_sync();                                   // vecA(l) (and vecM(l)) landed for everyone
          c.wait(B_X_DONE);
          tc_fence_after();
          stamp<DBG>(c);
          ln(vecA_s, trace_row(2 * l));
          stamp<DBG>(c);
          for (int h = 0; h < p.npass; ++h) {
            const uint32_t bq_s = vecA_s + (uint32_t)(G::VecBq + h * G::BqStride) * 4u;
            if constexpr (P128) {                           // both drains, both attention halves, all their barriers
              stamp<DBG>(c); stamp<DBG>(c);
              c.phases = attention_pass_p128<HSP>(c, bq_s, p.S, p.T);
              stamp<DBG>(c); stamp<DBG>(c);
            } else {
            c.wait(B_ACC_FULL0);
            tc_fence_after();
            stamp<DBG>(c);
            // arrives on ACC_EMPTY0 once its TMEM reads are done
            if constexpr (WIDE) {                           // [Q|K] job, then the V job into the same accumulator
              if constexpr (PREC) drain_qkv_p<G, 3>(c, bq_s); else drain_qkv<G, 3>(c, bq_s);
              c.wait(B_ACC_FULL0);
              tc_fence_after();
              if constexpr (PREC) drain_qkv_p<G, 4>(c, bq_s); else drain_qkv<G, 4>(c, bq_s);
            } else {
              if constexpr (PREC) drain_qkv_p<G, 7>(c, bq_s); else drain_qkv<G, 7>(c, bq_s);
            }
            attn_sync();                                    // Q|K|V of this head visible to all 10 attention warps
            uint32_t ybar = 0u;
            const uint32_t ypar = (c.phases >> B_Y_EMPTY) & 1u;
            if constexpr (WIDE) { ybar = c.bar(B_Y_EMPTY); c.phases ^= 1u << B_Y_EMPTY; }   // waited for inside the attention
            else c.wait(B_Y_EMPTY);                         // previous head's Y consumed by its proj MMAs
            stamp<DBG>(c);
            if constexpr (PREC) attention_head_p<G, HSP>(sbase, c.ctid >> 5, lane, p.S, p.T, ybar, ypar);
            else attention_head<G, HSP>(sbase, c.ctid >> 5, lane, p.S, p.T, ybar, ypar);
            fence_async_smem();
            c.arrive(B_Y_READY);
            stamp<DBG>(c);
            attn_sync();                                    // staging may be overwritten by the next drain
            stamp<DBG>(c);
            }
          }
          // vecA is free: prefetch the next layer's (or the final block)
          load_vec_async(c, G::SmVecA, p.vec + (size_t)(l + 1) * layer_stride, G::VecAFloats);
          // ---------------- MLP half ----------------
          c.wait(B_X_DONE);
          tc_fence_after();
          stamp<DBG>(c);
          ln(vecM_s, trace_row(2 * l + 1));
          stamp<DBG>(c);
          for (int ch = 0; ch < G::NCH; ++ch) {
            const int b = ch & 1, tb = WIDE ? 0 : b;
            if constexpr (P128) {                           // one H buffer: the drain itself waits for FC2(ch - 1)
              c.wait(B_ACC_FULL0);
              tc_fence_after();
              stamp<DBG>(c);
              c.phases = drain_gelu_p128(c, vecM_s + (uint32_t)G::VecB1F * 4u + (uint32_t)ch * 512u);
              stamp<DBG>(c);
            } else {
            c.wait2(tb ? B_ACC_FULL1 : B_ACC_FULL0, b ? B_OP_EMPTY1 : B_OP_EMPTY0);   // accumulator ready, H[b] consumed by FC2(ch-2)
            tc_fence_after();
            stamp<DBG>(c);
            // arrives on ACC_EMPTY and (twice) on OP_READY itself
            if constexpr (PREC) drain_gelu_p<G>(c, tb, b, vecM_s + (uint32_t)G::VecB1F * 4u + (uint32_t)ch * 512u);
            else drain_gelu<G>(c, tb, b, vecM_s + (uint32_t)G::VecB1H * 4u + (uint32_t)ch * 256u);
            stamp<DBG>(c);
            }
          }
          compute_sync();                                   // everyone done with vecM(l)
          const int nl = (l + 1 < p.L) ? l + 1 : 0;
          load_vec_async(c, G::SmVecM, p.vec + (size_t)nl * layer_stride + G::VecAFloats, G::VecMFloats);
        }
        // ---------------- ln_f + action head + pre-conditioning + sampler update ----------------
        c.phases = eval_epilogue<G, DBG, PREC>(c, p, sa, tile, step, second, trace_row(2 * p.L));
        if (sa.n_steps) {
          const bool two = (sa.sampler == BESO_SAMPLER_HEUN && sa.sig[step + 1] != 0.0f) ||
                           (sa.sampler == BESO_SAMPLER_TWO_STAGE && sa.sigb[step] != 0.0f);
          if (!second && two) second = 1;
          else { second = 0; ++step; }
        }
      }
      if (sa.n_steps) tile_end<G>(c, p, sa, tile);
    }
    cp_async_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR == 2) cluster_sync_all();        // the pair's MMAs / multicasts touch both CTAs' shared memory and TMEM
  if (warp == kMmaWarp) { if (CG == 2) tmem_dealloc_cg2(tmem, 512); else tmem_dealloc(tmem, 512); }
}

// ================================ tcgen05 issue-rate probe ==========================================
// One warp issues 64 K=16 MMAs (M=128, SS operands, SW128 K-major) per variant, as straight-line code,
// and reports SM cycles until the last one has completed: calibrates what the fused kernel can expect
// from the tensor pipe.  CE = commit to a never-waited barrier after every CE MMAs (0 = none).
template <int N, int CE, int DSTRIDE>
__device__ __forceinline__ void probe_variant(uint32_t tm, uint64_t a_desc, uint64_t b_desc, uint32_t dummy_bar,
                                              uint32_t done_bar, long long* out) {
  constexpr uint32_t idesc = idesc_bf16_m128(N);
  __syncwarp();
  const long long t0 = clock64();
  if (elect_one()) {
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      mma_bf16(tm + (uint32_t)((i >> 2) & 1) * DSTRIDE, a_desc + 2u * (i & 3) + (uint64_t)((i >> 2) & 3) * 1024u,
               b_desc + 2u * (i & 3), idesc, 1u);
      if (CE > 0 && (i % (CE > 0 ? CE : 1)) == CE - 1 && i != 63) mma_commit(dummy_bar);
    }
    mma_commit(done_bar);
  }
  __syncwarp();
  const long long t1 = clock64();
  spin_wait(done_bar, 0);
  const long long t2 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
}

// warp 0: MMA issuer; warp 1 (mode & 1): streams `src` through a 4 x 16 KB bulk-copy ring as fast as it can;
// warps 2-5 (mode & 2): hammer shared memory with 16-byte loads and stores.
__global__ void __launch_bounds__(192, 1) mma_rate_kernel(long long* out, const uint8_t* src, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(sm);
  const uint32_t a_off = 0, b_off = 65536, ring = 98304, junk = 163840, bars = 196608;   // A 4x[128x64], B [256x64]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + bars + 512);
  volatile int* stop = reinterpret_cast<volatile int*>(sm + bars + 516);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < bars / 16; i += 192) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { for (int i = 0; i < 48; ++i) mbar_init(sbase + bars + i * 8, 1); *stop = 0; fence_barrier_init(); }
  fence_async_smem();
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    const uint32_t tm = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const uint64_t a_desc = smem_desc_sw128(sbase + a_off), b_desc = smem_desc_sw128(sbase + b_off);
    const uint32_t dummy = sbase + bars + 31 * 8;
    probe_variant<256, 0, 0>(tm, a_desc, b_desc, dummy, sbase + bars + 0 * 8, out + 0);
    probe_variant<256, 4, 0>(tm, a_desc, b_desc, dummy, sbase + bars + 1 * 8, out + 2);
    probe_variant<128, 0, 0>(tm, a_desc, b_desc, dummy, sbase + bars + 2 * 8, out + 4);
    probe_variant<128, 4, 0>(tm, a_desc, b_desc, dummy, sbase + bars + 3 * 8, out + 6);
    probe_variant<256, 0, 256>(tm, a_desc, b_desc, dummy, sbase + bars + 4 * 8, out + 8);
    probe_variant<256, 1, 0>(tm, a_desc, b_desc, dummy, sbase + bars + 5 * 8, out + 10);
    probe_variant<64, 0, 0>(tm, a_desc, b_desc, dummy, sbase + bars + 6 * 8, out + 12);
    probe_variant<192, 0, 0>(tm, a_desc, b_desc, dummy, sbase + bars + 7 * 8, out + 14);
    probe_variant<256, 16, 0>(tm, a_desc, b_desc, dummy, sbase + bars + 8 * 8, out + 16);
    probe_variant<128, 16, 0>(tm, a_desc, b_desc, dummy, sbase + bars + 9 * 8, out + 18);
    *stop = 1;
    tc_fence_before();
    __syncwarp();
    tmem_dealloc(tm, 512);
  } else if (warp == 1 && (mode & 1)) {
    uint32_t g = 0, off = 0;
    long long copies = 0;
    while (!*stop) {
      const uint32_t slot = g & 3, par = (g >> 2) & 1u, full = sbase + bars + (32 + slot) * 8;
      if (g >= 4) spin_wait(full, par ^ 1u);                 // the previous copy into this slot has landed
      if (elect_one()) { mbar_expect_tx(full, 16384); bulk_g2s(sbase + ring + slot * 16384, src + off, 16384, full); }
      __syncwarp();
      off = (off + 16384) & (6291456 - 1 - 16383);
      ++g; ++copies;
    }
    for (uint32_t k = (g >= 4 ? g - 4 : 0); k < g; ++k) spin_wait(sbase + bars + (32 + (k & 3)) * 8, (k >> 2) & 1u);
    if (lane == 0) out[20] = copies;
  } else if (warp >= 2 && (mode & 2)) {
    uint4* jb = reinterpret_cast<uint4*>(sm + junk);
    uint4 acc = make_uint4(0, 0, 0, 0);
    int idx = threadIdx.x - 64;
    while (!*stop) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint4 v = jb[(idx + k * 128) & 2047];
        acc.x ^= v.x; acc.y += v.y;
        jb[(idx + k * 128 + 64) & 2047] = acc;
      }
    }
    if (acc.x == 0x12345678u) out[21] = acc.y;
  }
}

// ================================ weight packing ===================================================
// Row / column index of a packed operand -> index in the reference tensor.  Heads are packed 64 columns per
// attention pass with the head size padded to hsp (32 or 64): packed index i of a pass whose first head is head0
// is element i % hsp of head head0 + i / hsp (zero beyond the true head size hs).  hsp == 0: identity.
struct IndexMap { int hsp, hs, head0, n_heads; };
__device__ __forceinline__ int map_index(const IndexMap& m, int i, int limit) {
  if (m.hsp == 0) return i < limit ? i : -1;
  const int head = m.head0 + i / m.hsp, e = i % m.hsp;
  return (e < m.hs && head < m.n_heads) ? head * m.hs + e : -1;
}
struct PackTile {       // one [rows x 64] fp16 SW128 sub-tile of the tape from a row-major fp32 matrix
  const float* src; int ld, row0, col0, rows, n_rows, n_cols; float scale; uint32_t dst;
  const float* colscale;   // optional per-input-column factor: the preceding LayerNorm's weight
  IndexMap rmap, cmap;     // packed row / column -> source row / column (identity: index + row0 / col0 below n_rows / n_cols)
  int part;                // 0: fp16(x)   1: hi image = fp16(x)   2: lo image = fp16(x - hi)
};
__global__ void pack_tiles_kernel(const PackTile* tiles, uint8_t* tape) {
  const PackTile t = tiles[blockIdx.x];
  for (int idx = threadIdx.x; idx < t.rows * 8; idx += blockDim.x) {
    const int r = idx >> 3, chunk = idx & 7;
    const int sr = map_index(t.rmap, t.row0 + r, t.n_rows);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int sc = map_index(t.cmap, t.col0 + chunk * 8 + i, t.n_cols);
      float x = 0.f;
      if (sr >= 0 && sc >= 0) {
        x = t.src[(size_t)sr * t.ld + sc] * t.scale;
        if (t.colscale != nullptr) x *= t.colscale[sc];
      }
      if (t.part == 2) x -= __half2float(__float2half_rn(x));
      v[i] = x;
    }
    st_chunk_h(tape + t.dst, r, chunk, v);
  }
}

struct EmbSrc {
  const float *pos, *tokw, *tokb, *sigw, *sigb, *actw, *actb;
  const float* resid_bias[2 * kMaxLayers];     // effective attn.proj bias (256 padded) and mlp.2.bias (d) of every layer
  int obs, act, G, W, L, d, prec;
};
// Embedding GEMM B operand: W_emb[n][k], n < 256 (384), k < 128 (atom 0 = obs, atom 1 = misc), as fills of
// [128 rows x 64] in (k-atom, row-block) order; a ring group holds `group_blocks` consecutive fills (G256: 2 = all rows
// of a K atom, G384: 1).  fp16 mode: bf16, tables as hi + lo column pairs.
// PREC: fp16 hi image (blockIdx.y = 0) and lo image (1) of the plain values, one column per table entry, stored
// [hi group | lo group].
__global__ void pack_emb_kernel(EmbSrc s, uint8_t* tape, int nblk, int group_blocks) {
  const int fill = blockIdx.x;                 // atom = fill / nblk, rows (fill % nblk) * 128 ..
  const int atom = fill / nblk, n0 = (fill % nblk) * 128, image = blockIdx.y;
  for (int idx = threadIdx.x; idx < 128 * 8; idx += blockDim.x) {
    const int r = idx >> 3, chunk = idx & 7, n = n0 + r;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = chunk * 8 + i;
      float x = 0.f;
      if (n >= s.d) {
      } else if (atom == 0) {
        if (k < s.obs) x = s.tokw[(size_t)n * s.obs + k];
      } else if (k < s.act) {
        x = s.actw[(size_t)n * s.act + k];
      } else if (k < s.act + 3) {
        const float w = s.sigw[n];
        if (s.prec) {
          x = (k == s.act) ? w : 0.f;
        } else {
          const float hi = __bfloat162float(__float2bfloat16_rn(w));
          x = (k == s.act + 2) ? (w - hi) : hi;    // A holds [cn_hi, cn_lo, cn_hi]
        }
      } else if (k >= kOneHot0 && k < kOneHot0 + 2 * kMaxTokens) {
        const int tok = (k - kOneHot0) >> 1;
        float tbl;
        if (tok == 0) tbl = s.sigb[n];
        else if (tok <= s.G) tbl = s.tokb[n] + s.pos[(size_t)(tok - 1) * s.d + n];
        else {
          const int j = tok - 1 - s.G, step = j >> 1;
          tbl = (step < s.W) ? ((j & 1) ? s.actb[n] : s.tokb[n]) + s.pos[(size_t)(s.G + step) * s.d + n] : 0.f;
        }
        for (int i2 = 0; i2 < 2 * s.L; ++i2) tbl += s.resid_bias[i2][n];   // all residual-branch biases, up front
        if (s.prec) {
          x = ((k - kOneHot0) & 1) ? 0.f : tbl;
        } else {
          const float hi = __bfloat162float(__float2bfloat16_rn(tbl));
          x = ((k - kOneHot0) & 1) ? (tbl - hi) : hi;
        }
      }
      if (s.prec && image == 1) x -= __half2float(__float2half_rn(x));
      v[i] = x;
    }
    if (s.prec) st_chunk_h(tape + ((size_t)((fill / group_blocks) * 2 + image) * group_blocks + fill % group_blocks) * 16384, r, chunk, v);
    else st_chunk(tape + (size_t)fill * 16384, r, chunk, v);
  }
}

// pend vectors: minus the residual-branch biases that the embedding GEMM added too early.
//   LN1(l): -(sum_{l' >= l} bproj + b2)     LN2(l): LN1(l) + bproj_l      ln_f: 0
__global__ void pack_pend_kernel(EmbSrc s, float* vec, uint32_t layer_stride, uint32_t m_off, int dp) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= dp) return;
  float r = 0.f;
  vec[(size_t)s.L * layer_stride + n] = 0.f;
  for (int l = s.L - 1; l >= 0; --l) {
    const float bp = n < s.d ? s.resid_bias[2 * l][n] : 0.f, b2 = n < s.d ? s.resid_bias[2 * l + 1][n] : 0.f;
    vec[(size_t)l * layer_stride + m_off + n] = -(r + b2);
    r += bp + b2;
    vec[(size_t)l * layer_stride + n] = -r;
  }
}

// Bias of a Linear that follows a LayerNorm whose affine part is folded into it:
//   W (LN0(x) * g + beta) + b = (W diag(g)) LN0(x) + (b + W beta)        dst[i] = scale * (b[i] + W[i,:] . beta)
// half_out: the n results are stored as packed fp16 starting at float index dst.  map: packed index -> source index.
struct VecCopy { const float* src; uint32_t dst; int n, n_src; float scale; const float* W; const float* beta; int ld; int half_out; IndexMap map; };
__global__ void pack_vec_kernel(const VecCopy* cp, float* vec) {
  const VecCopy c = cp[blockIdx.x];
  for (int i = threadIdx.x; i < c.n; i += blockDim.x) {
    const int si = map_index(c.map, i, c.n_src);
    float b = 0.f;
    if (si >= 0) {
      b = c.src ? c.src[si] : 0.f;
      if (c.W != nullptr) {
        float acc = 0.f;
        for (int k = 0; k < c.ld; ++k) acc = fmaf(c.W[(size_t)si * c.ld + k], c.beta[k], acc);
        b += acc;
      }
    }
    if (c.half_out) reinterpret_cast<__half*>(vec + c.dst)[i] = __float2half_rn(b * c.scale);
    else vec[c.dst + i] = b * c.scale;
  }
}

}  // namespace

// ================================ host API =========================================================
static int padded_head(const beso_model_desc& m) { const int hs = m.d / m.n_heads; return hs <= 32 ? 32 : 64; }
bool fast_supported(const beso_model_desc& m) {
  const int G = m.goal_conditioned ? m.goal_len : 0;
  const int T = 1 + G + 2 * m.window;
  if (m.d > G384::DP || m.d % 8 != 0 || m.n_heads < 1 || m.d % m.n_heads != 0) return false;
  const int hs = m.d / m.n_heads;
  if (hs > 64) return false;
  const int per_pass = 64 / padded_head(m), npass = (m.n_heads + per_pass - 1) / per_pass;
  return npass <= kMaxPass && m.linear_output && m.n_layers <= kMaxLayers && m.obs_dim <= kMaxObs &&
         m.act_dim <= kMaxAct && T <= kMaxTokens;
}

// Precise mode on full 128-row tiles (geometry G256P): embed_dim <= 256.
bool fast_p128_supported(const beso_model_desc& m) { return fast_supported(m) && m.d <= G256::DP; }

// Which precise layout a launch uses.  One evaluation of a P128 tile (128 rows) costs about kP128Cost evaluations of a
// stacked tile (64 rows) (in-kernel timelines, K256: 669 k against 385 k cycles; profiles/): the batch goes to whichever needs less time over the SMs, so
// a batch-1 rollout keeps the short stacked tile and a batch that fills the chip takes the half as many, fuller P128
// tiles.  BESO_PREC_LAYOUT=stacked|p128 forces one (A / B measurements).
constexpr double kP128Cost = 1.7;
static int tile_waves(int B, int per_tile, int sm_count) { const int tiles = (B + per_tile - 1) / per_tile; return (tiles + sm_count - 1) / sm_count; }
static int g_prec_layout = -1;                  // -1: not set yet (environment), 0 auto, 1 stacked, 2 P128
void fast_set_prec_layout(int layout) { g_prec_layout = (layout == FAST_LAYOUT_STACKED || layout == FAST_LAYOUT_P128) ? layout : 0; }
static bool choose_p128(const beso_model_desc& m, int T, int B, bool cfg, int sm_count) {
  if (g_prec_layout < 0) {
    const char* e = getenv("BESO_PREC_LAYOUT");
    g_prec_layout = !e ? 0 : (!strcmp(e, "stacked") ? 1 : (!strcmp(e, "p128") ? 2 : 0));
  }
  const int forced = g_prec_layout;
  const int unit = cfg ? 2 : 1;
  int s_st = 64 / T, s_p = 2 * (64 / T);
  s_st -= s_st % unit; s_p -= s_p % (2 * unit);
  if (s_p < 2 * unit) return false;
  if (forced) return forced == 2;
  if (s_st < unit) return true;
  return kP128Cost * tile_waves(B, s_p / unit, sm_count) < (double)tile_waves(B, s_st / unit, sm_count);
}

int fast_seqs_per_tile(const beso_model_desc& m, int t, bool prec) {
  if (!fast_supported(m)) return 0;
  const int G = m.goal_conditioned ? m.goal_len : 0;
  return (prec ? 64 : kRows) / (1 + G + 2 * t);
}

void fast_free(FastWeights& w) {
  if (w.tape) cudaFree(w.tape);
  if (w.vec) cudaFree(w.vec);
  w.tape = nullptr; w.vec = nullptr;
}

// Parameter indices in parameters() order
static int p_layer(int l, int k) { return 3 + l * 16 + k; }   // k: 0 ln1w 1 ln1b 2 ln2w 3 ln2b 4 key.w 5 key.b 6 query.w 7 query.b
                                                               //    8 value.w 9 value.b 10 proj.w 11 proj.b 12 mlp0.w 13 mlp0.b 14 mlp2.w 15 mlp2.b

int fast_pack(FastWeights& w, const beso_model_desc& m, const float* const* prm, cudaStream_t st, int layout) {
  const bool prec = layout != FAST_LAYOUT_F16;
  if (layout == FAST_LAYOUT_P128 && !fast_p128_supported(m)) { set_error("P128 layout needs embed_dim <= 256"); return BESO_E_UNSUPPORTED; }
  const int L = m.n_layers, G = m.goal_conditioned ? m.goal_len : 0;
  const int d = m.d, H = m.n_heads, hs = d / H, hsp = padded_head(m), per_pass = 64 / hsp, npass = (H + per_pass - 1) / per_pass;
  const int ff = 4 * d;
  const size_t images = prec ? 2 : 1;
  // geometry (G256 / G384): padded width, K atoms, hidden chunks, vector block layout
  const bool wide = d > G256::DP;
  const bool oneacc = wide || layout == FAST_LAYOUT_P128;  // single-accumulator schedules: 16 KB ring groups
  const int DP = wide ? G384::DP : G256::DP, NA = DP / 64, NB = DP / 128, FFP = 4 * DP, NCH = FFP / 128;
  const uint32_t vA = wide ? G384::VecAFloats : G256::VecAFloats, vM = wide ? G384::VecMFloats : G256::VecMFloats;
  const uint32_t vBq = wide ? G384::VecBq : G256::VecBq, vBqStride = wide ? G384::BqStride : G256::BqStride;
  const uint32_t vB1H = wide ? G384::VecB1H : G256::VecB1H, vB1F = wide ? G384::VecB1F : G256::VecB1F;
  // G256 per evaluation: embedding 4 x 16 KB | per layer: per pass QKV 4 x 24 KB + proj 32 KB, FC1 32 x 16 KB, FC2 16 x 32 KB | head 4 x 2 KB
  // G384 per evaluation: embedding 6 x 16 KB | per layer: per pass ([Q|K] 6 + V 3 + proj 3) x 16 KB, (FC1 6 + FC2 6) x 16 KB per chunk | head 6 x 2 KB
  // (PREC: every ring group twice, hi image then lo image)
  const size_t tape_bytes = oneacc
      ? images * ((size_t)2 * NB * 16384 + (size_t)L * ((size_t)npass * (NA + NA / 2 + NB) + (size_t)NCH * (NA + 2 * NB)) * 16384 + (size_t)NA * 2048)
      : images * (4 * 16384 + (size_t)L * ((size_t)npass * (4 * 24576 + 32768) + 32 * 16384 + 16 * 32768) + 4 * 2048);
  const size_t vec_floats = (size_t)L * (vA + vM) + vA;
  const size_t fold_floats = (size_t)L * 2 * DP;             // per layer: effective V bias | effective proj bias (padded)
  if (!w.tape) {
    // tape | (scratch tables for the pack kernels)
    BESO_CUDA(cudaMalloc(&w.tape, tape_bytes + (4 << 20)));
    BESO_CUDA(cudaMalloc(&w.vec, (vec_floats + fold_floats) * sizeof(float)));
    BESO_CUDA(cudaMemsetAsync(w.vec, 0, (vec_floats + fold_floats) * sizeof(float), st));
    w.tape_bytes = tape_bytes; w.vec_floats = vec_floats;
  }
  uint8_t* tape = reinterpret_cast<uint8_t*>(w.tape);
  uint8_t* scratch = tape + tape_bytes;

  // ---- tape sub-tiles, in program order ----
  std::vector<PackTile> tiles;
  uint32_t off = (uint32_t)(images * 2 * NB * 16384);         // embedding fills are written by pack_emb_kernel
  size_t group_first = 0;
  uint32_t group_off = off;
  const IndexMap ident{0, 0, 0, 0};
  auto tile = [&](const float* src, int ld, int row0, int col0, int rows, int n_rows, int n_cols, float scale, const float* colscale,
                  IndexMap rmap, IndexMap cmap) {
    tiles.push_back({src, ld, row0, col0, rows, n_rows, n_cols, scale, off, colscale, rmap, cmap, prec ? 1 : 0});
    off += rows * 128;
  };
  // closes a ring group: in PREC the lo image of the same tiles follows
  auto end_group = [&]() {
    if (prec) {
      const uint32_t bytes = off - group_off;
      const size_t n = tiles.size();
      for (size_t i = group_first; i < n; ++i) { PackTile t = tiles[i]; t.part = 2; t.dst += bytes; tiles.push_back(t); }
      off += bytes;
    }
    group_first = tiles.size();
    group_off = off;
  };
  const float qscale = 1.4426950408889634f / sqrtf((float)hs);   // log2(e) / sqrt(head size): the softmax works in base 2
  const float s1 = prec ? 1.f : kGeluInScale, s2 = prec ? 1.f : kGeluOutScale;
  for (int l = 0; l < L; ++l) {
    const float *wk = prm[p_layer(l, 4)], *wq = prm[p_layer(l, 6)], *wv = prm[p_layer(l, 8)], *wp = prm[p_layer(l, 10)];
    const float *w1 = prm[p_layer(l, 12)], *w2 = prm[p_layer(l, 14)];
    const float *ln1w = prm[p_layer(l, 0)], *ln2w = prm[p_layer(l, 2)];
    auto qkv = [&](int h) {                                   // attention pass h: heads h * per_pass ..
      const IndexMap hm{hsp, hs, h * per_pass, H};
      for (int kb = 0; kb < 4; ++kb) {
        tile(wq, d, 0, kb * 64, 64, d, d, qscale, ln1w, hm, ident);
        tile(wk, d, 0, kb * 64, 64, d, d, 1.f, ln1w, hm, ident);
        tile(wv, d, 0, kb * 64, 64, d, d, 1.f, ln1w, hm, ident);
        end_group();
      }
    };
    auto qk_w = [&](int h) {                                  // G384: [Q|K] of pass h, one group per K atom
      const IndexMap hm{hsp, hs, h * per_pass, H};
      for (int kb = 0; kb < NA; ++kb) {
        tile(wq, d, 0, kb * 64, 64, d, d, qscale, ln1w, hm, ident);
        tile(wk, d, 0, kb * 64, 64, d, d, 1.f, ln1w, hm, ident);
        end_group();
      }
    };
    auto v_w = [&](int h) {                                   // G384: V of pass h, two K atoms per group
      const IndexMap hm{hsp, hs, h * per_pass, H};
      for (int kb = 0; kb < NA; ++kb) {
        tile(wv, d, 0, kb * 64, 64, d, d, 1.f, ln1w, hm, ident);
        if (kb & 1) end_group();
      }
    };
    auto proj = [&](int h) {                                  // G256: one group of 256 rows; G384: one group per 128 rows
      const IndexMap hm{hsp, hs, h * per_pass, H};
      for (int nb = 0; nb < NB; ++nb) {
        for (int s = 0; s < 2; ++s) tile(wp, d, nb * 128 + s * 64, 0, 64, d, d, 1.f, nullptr, ident, hm);
        if (oneacc) end_group();
      }
      if (!oneacc) end_group();
    };
    auto fc1 = [&](int c) {                                   // G256: two K atoms per group; single-accumulator: one
      for (int kb = 0; kb < NA; ++kb) {
        for (int s = 0; s < 2; ++s) tile(w1, d, c * 128 + s * 64, kb * 64, 64, ff, d, s1, ln2w, ident, ident);
        if (oneacc || (kb & 1)) end_group();
      }
    };
    auto fc2 = [&](int c) {
      for (int kb = 0; kb < 2; ++kb) {
        for (int nb = 0; nb < NB; ++nb) {
          for (int s = 0; s < 2; ++s) tile(w2, ff, nb * 128 + s * 64, c * 128 + kb * 64, 64, d, ff, s2, nullptr, ident, ident);
          if (oneacc) end_group();
        }
        if (!oneacc) end_group();
      }
    };
    if (oneacc) {
      qk_w(0); v_w(0);
      for (int h = 1; h < npass; ++h) { qk_w(h); v_w(h); proj(h - 1); }
    } else {
      qkv(0);
      for (int h = 1; h < npass; ++h) { qkv(h); proj(h - 1); }
    }
    proj(npass - 1);
    fc1(0); fc1(1); fc2(0);
    for (int c = 2; c < NCH; ++c) { fc1(c); fc2(c - 1); }
    fc2(NCH - 1);
  }
  const int p_tail = 3 + 16 * L;                              // ln_f.w, ln_f.b, sigma_emb.w/b, action_emb.w/b, action_pred.w/b
  for (int kb = 0; kb < NA; ++kb) tile(prm[p_tail + 6], d, 0, kb * 64, 16, m.act_dim, d, 1.f, prm[p_tail], ident, ident);
  end_group();
  if (off != tape_bytes) { set_error("internal: tape layout mismatch"); return BESO_E_INVALID; }
  if (tiles.size() * sizeof(PackTile) > (3 << 20)) { set_error("internal: pack table too large"); return BESO_E_INVALID; }
  BESO_CUDA(cudaMemcpyAsync(scratch, tiles.data(), tiles.size() * sizeof(PackTile), cudaMemcpyHostToDevice, st));
  pack_tiles_kernel<<<(unsigned)tiles.size(), 128, 0, st>>>(reinterpret_cast<const PackTile*>(scratch), tape);
  ++g_kernel_launches;
  BESO_CUDA(cudaGetLastError());
  // ---- fp32 / fp16 vectors ----
  // Linear biases with the preceding LayerNorm's beta folded in (pack_vec_kernel).  K needs no bias (softmax is
  // invariant to a per-query shift of the scores); the V bias goes through the softmax average unchanged and is
  // folded into the projection bias:  bproj_eff = bproj + Wproj (bv + Wv beta1).
  std::vector<VecCopy> vc, vc2;
  for (int l = 0; l < L; ++l) {
    const uint32_t a = (uint32_t)(l * (vA + vM)), mo = a + vA;
    const uint32_t fo = (uint32_t)(vec_floats + (size_t)l * 2 * DP);
    const float* ln1b = prm[p_layer(l, 1)];
    const float* ln2b = prm[p_layer(l, 3)];
    for (int h = 0; h < npass; ++h)
      vc.push_back({prm[p_layer(l, 7)], a + vBq + h * vBqStride, 64, d, qscale, prm[p_layer(l, 6)], ln1b, d, 0, IndexMap{hsp, hs, h * per_pass, H}});
    vc.push_back({prm[p_layer(l, 9)], fo, DP, d, 1.f, prm[p_layer(l, 8)], ln1b, d, 0, ident});
    vc2.push_back({prm[p_layer(l, 11)], fo + DP, DP, d, 1.f, prm[p_layer(l, 10)], w.vec + fo, d, 0, ident});
    if (prec) vc.push_back({prm[p_layer(l, 13)], mo + vB1F, FFP, ff, 1.f, prm[p_layer(l, 12)], ln2b, d, 0, ident});
    else vc.push_back({prm[p_layer(l, 13)], mo + vB1H, FFP, ff, kGeluInScale, prm[p_layer(l, 12)], ln2b, d, 1, ident});   // b1 / 4 as fp16
  }
  const uint32_t fa = (uint32_t)(L * (vA + vM));
  vc.push_back({prm[p_tail + 7], fa + vBq, 16, m.act_dim, 1.f, prm[p_tail + 6], prm[p_tail + 1], d, 0, ident});
  uint8_t* scratch2 = scratch + (3 << 20);
  const size_t n1 = vc.size();
  vc.insert(vc.end(), vc2.begin(), vc2.end());
  BESO_CUDA(cudaMemcpyAsync(scratch2, vc.data(), vc.size() * sizeof(VecCopy), cudaMemcpyHostToDevice, st));
  pack_vec_kernel<<<(unsigned)n1, 128, 0, st>>>(reinterpret_cast<const VecCopy*>(scratch2), w.vec);
  pack_vec_kernel<<<(unsigned)vc2.size(), 128, 0, st>>>(reinterpret_cast<const VecCopy*>(scratch2) + n1, w.vec);
  g_kernel_launches += 2;
  BESO_CUDA(cudaGetLastError());

  EmbSrc es{};
  es.pos = prm[0]; es.tokw = prm[1]; es.tokb = prm[2]; es.sigw = prm[p_tail + 2]; es.sigb = prm[p_tail + 3];
  es.actw = prm[p_tail + 4]; es.actb = prm[p_tail + 5];
  es.obs = m.obs_dim; es.act = m.act_dim; es.G = G; es.W = m.window; es.L = L; es.d = d; es.prec = prec ? 1 : 0;
  for (int l = 0; l < L; ++l) {
    es.resid_bias[2 * l] = w.vec + vec_floats + (size_t)l * 2 * DP + DP;      // effective projection bias
    es.resid_bias[2 * l + 1] = prm[p_layer(l, 15)];
  }
  pack_emb_kernel<<<dim3(2 * NB, (unsigned)images), 256, 0, st>>>(es, tape, NB, oneacc ? 1 : 2);
  pack_pend_kernel<<<1, DP, 0, st>>>(es, w.vec, vA + vM, vA, DP);
  g_kernel_launches += 2;
  BESO_CUDA(cudaGetLastError());
  BESO_CUDA(cudaStreamSynchronize(st));                       // the host tables above go out of scope
  return BESO_OK;
}

// Stream-ordered scratch for the wide geometry's x buffers: a private memory pool that keeps its pages between launches
// (the device's default pool hands freed memory back to the driver at every synchronisation, which makes each launch
// pay a real allocation -- measured ~1 ms).
static cudaMemPool_t scratch_pool(int device, cudaError_t* err) {
  static std::mutex mu;
  static cudaMemPool_t pools[64] = {};
  std::lock_guard<std::mutex> lock(mu);
  *err = cudaSuccess;
  if (device < 0 || device >= 64) { *err = cudaErrorInvalidDevice; return nullptr; }
  if (pools[device] == nullptr) {
    cudaMemPoolProps props{};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    *err = cudaMemPoolCreate(&pools[device], &props);
    if (*err != cudaSuccess) { pools[device] = nullptr; return nullptr; }
    unsigned long long keep = ~0ull;
    *err = cudaMemPoolSetAttribute(pools[device], cudaMemPoolAttrReleaseThreshold, &keep);
  }
  return pools[device];
}

static float* g_trace = nullptr;
static long long* g_timeline = nullptr;
void fast_set_trace(float* trace_dev) { g_trace = trace_dev; }
void fast_set_timeline(long long* dev) { g_timeline = dev; }
int fast_mma_rate(long long* out_dev, const void* src_dev, int mode, cudaStream_t st) {
  BESO_CUDA(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200704));
  mma_rate_kernel<<<1, 192, 200704, st>>>(out_dev, reinterpret_cast<const uint8_t*>(src_dev), mode);
  BESO_CUDA(cudaGetLastError());
  return BESO_OK;
}

int fast_launch(const FastWeights& w_in, const beso_model_desc& m, int sm_count, const SampleArgs& sa,
                const float* state, const float* goal, const float* x, const float* sigma, float* out,
                int B, int t, uint32_t flags, float lambda, cudaStream_t st, bool prec, const FastWeights* w_p128) {
  const int T_tok = 1 + (m.goal_conditioned ? m.goal_len : 0) + 2 * t;
  const bool p128 = prec && w_p128 != nullptr && w_p128->tape != nullptr &&
                    choose_p128(m, T_tok, B, (flags & BESO_FLAG_CFG) != 0, sm_count);
  const FastWeights& w = p128 ? *w_p128 : w_in;
  if (!w.tape) { set_error("fast weights not packed"); return BESO_E_NOT_PACKED; }
  FastParams p{};
  const int L = m.n_layers;
  const int hsp = padded_head(m), per_pass = 64 / hsp;
  p.hsp = hsp; p.npass = (m.n_heads + per_pass - 1) / per_pass;
  p.d_true = m.d; p.inv_d = 1.0f / (float)m.d;
  const bool wide = m.d > G256::DP;
  // (timeline buffer layout only: the issuer's per-job stamps come first, 6 * n_fills slots are reserved for them)
  const size_t n_fills = (wide || p128) ? 8 + (size_t)L * 64 : 4 + (size_t)L * 104 + 4;
  p.tape = reinterpret_cast<const uint8_t*>(w.tape);
  p.vec = w.vec;
  p.n_fills = (int)n_fills; p.L = L; p.G = m.goal_conditioned ? m.goal_len : 0; p.obs = m.obs_dim; p.act = m.act_dim;
  p.t = t; p.T = 1 + p.G + 2 * t;
  const bool cfg = flags & BESO_FLAG_CFG;
  // sequences per tile come in units of `unit`: P128 fills the two 64-row halves of a tile alike, CFG keeps the
  // cond / uncond copies of a sequence together (in the same half)
  const int unit = (p128 ? 2 : 1) * (cfg ? 2 : 1);
  p.S = p128 ? 2 * (64 / p.T) : (prec ? 64 : kRows) / p.T;
  p.S -= p.S % unit;
  if (p.S < 1) { set_error("sequence does not fit a 128-row tile"); return BESO_E_UNSUPPORTED; }
  if (cfg && p.S < 2) { set_error("a cond / uncond pair does not fit a tile"); return BESO_E_UNSUPPORTED; }
  // A tile's time hardly depends on how many of its rows are used (one thread per row, M = 128 MMAs either way) except
  // for attention, which is per sequence.  So among the packings that need the same number of waves over the SMs take
  // the one with the FEWEST sequences per tile: it spreads the batch over more SMs (BASELINE config 2: 512 sequences
  // are 103 tiles of 5 or 128 tiles of 4 on 148 SMs -- one wave either way, with 20 % less attention work per tile).
  {
    const int step = unit, s_max = p.S;
    auto waves = [&](int S) { const int per = cfg ? S / 2 : S; const int tiles = (B + per - 1) / per; return (tiles + sm_count - 1) / sm_count; };
    const int w_min = waves(s_max);
    for (int S = step; S < s_max; S += step)
      if (waves(S) == w_min) { p.S = S; break; }
  }
  const int per_tile = cfg ? p.S / 2 : p.S;
  p.n_tiles = (B + per_tile - 1) / per_tile;
  p.B = B;
  p.evals = 1;
  if (sa.n_steps) {
    p.evals = sa.n_steps;
    if (sa.sampler == BESO_SAMPLER_HEUN)
      for (int i = 0; i < sa.n_steps; ++i) if (sa.sig[i + 1] != 0.0f) ++p.evals;
    if (sa.sampler == BESO_SAMPLER_TWO_STAGE)
      for (int i = 0; i < sa.n_steps; ++i) if (sa.sigb[i] != 0.0f) ++p.evals;
  }
  p.flags = flags; p.lambda = lambda; p.sigma_data = m.sigma_data;
  p.state = state; p.goal = goal; p.xin = x; p.sigma = sigma; p.out = out;
  p.trace = g_trace;
  p.timeline = g_timeline;
  // Single-CTA MMAs (CG = 1) are the default: measured faster than CTA pairs on this workload (the pair mode
  // halves L2 -> SMEM weight traffic but pays remote-arrive latency on every compute -> MMA hand-off; see
  // profiles/).  BESO_FAST_CG=2 selects the cta_group::2 path, kept parity-tested.
  static const int forced_cg = [] { const char* e = getenv("BESO_FAST_CG"); return e ? atoi(e) : 0; }();
  const bool pairs_ok = !wide && !prec && p.npass == kH && hsp == 64 && p.n_tiles >= 2;   // the pair modes replay the fixed 4-pass group table
  const int cg = (forced_cg == 2 && pairs_ok) ? 2 : 1;
  // BESO_FAST_MC=2: independent CTAs in clusters of 2 sharing the weight stream by TMA multicast
  static const int forced_mc = [] { const char* e = getenv("BESO_FAST_MC"); return e ? atoi(e) : 0; }();
  // diagnostics (LayerNorm trace dump, clock stamps) are compiled for the 64-wide head images only (build time: every
  // image of this kernel is ~15 k instructions); with 32-wide heads the request is ignored
  const bool dbg = (p.trace != nullptr || p.timeline != nullptr) && hsp == 64;
  const int mc = (cg == 1 && !dbg && forced_mc == 2 && pairs_ok) ? 2 : 1;
  const int smem = (int)kSmemBytes + 1024;
  auto launch = [&](auto kernel, int grid, bool cluster) -> int {
    BESO_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));   // cheap, idempotent
    cudaLaunchConfig_t cfgl{};
    cfgl.gridDim = dim3(grid);
    cfgl.blockDim = dim3(kThreads);
    cfgl.dynamicSmemBytes = smem;
    cfgl.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfgl.attrs = attr; cfgl.numAttrs = cluster ? 1 : 0;
    BESO_CUDA(cudaLaunchKernelEx(&cfgl, kernel, p, sa));
    return BESO_OK;
  };
  int rc;
  if (cg == 2 || mc == 2) {
    const int pairs = (p.n_tiles + 1) / 2, max_pairs = sm_count / 2;
    const int grid = 2 * (pairs < max_pairs ? pairs : max_pairs);
    rc = cg == 2 ? launch(fast_sample_kernel<2, false, 1, false, 64, 0>, grid, true) : launch(fast_sample_kernel<1, false, 2, false, 64, 0>, grid, true);
  } else {
    const int grid = p.n_tiles < sm_count ? p.n_tiles : sm_count;
    if (wide) {
      // the sampler's x / d1 / x2 / dU buffers of every CTA: stream-ordered scratch (the wide geometry has no shared
      // memory left for them), freed behind the kernel
      int device = 0;
      BESO_CUDA(cudaGetDevice(&device));
      cudaError_t pe;
      cudaMemPool_t pool = scratch_pool(device, &pe);
      if (pe != cudaSuccess) return cuda_fail(pe, "scratch memory pool");
      BESO_CUDA(cudaMallocFromPoolAsync(reinterpret_cast<void**>(&p.xscratch), (size_t)grid * 4 * kXFloats * sizeof(float), pool, st));
    }
    const int geo = wide ? 1 : (p128 ? 2 : 0);
    const int sel = geo * 8 + ((prec ? 4 : 0) | (dbg ? 2 : 0) | (hsp == 32 ? 1 : 0));
    switch (sel) {
      case 0: rc = launch(fast_sample_kernel<1, false, 1, false, 64, 0>, grid, false); break;
      case 1: rc = launch(fast_sample_kernel<1, false, 1, false, 32, 0>, grid, false); break;
      case 2: rc = launch(fast_sample_kernel<1, true, 1, false, 64, 0>, grid, false); break;
      case 4: rc = launch(fast_sample_kernel<1, false, 1, true, 64, 0>, grid, false); break;
      case 5: rc = launch(fast_sample_kernel<1, false, 1, true, 32, 0>, grid, false); break;
      case 6: rc = launch(fast_sample_kernel<1, true, 1, true, 64, 0>, grid, false); break;
      case 8: rc = launch(fast_sample_kernel<1, false, 1, false, 64, 1>, grid, false); break;
      case 9: rc = launch(fast_sample_kernel<1, false, 1, false, 32, 1>, grid, false); break;
      case 10: rc = launch(fast_sample_kernel<1, true, 1, false, 64, 1>, grid, false); break;
      case 12: rc = launch(fast_sample_kernel<1, false, 1, true, 64, 1>, grid, false); break;
      case 13: rc = launch(fast_sample_kernel<1, false, 1, true, 32, 1>, grid, false); break;
      case 14: rc = launch(fast_sample_kernel<1, true, 1, true, 64, 1>, grid, false); break;
      case 20: rc = launch(fast_sample_kernel<1, false, 1, true, 64, 2>, grid, false); break;
      case 21: rc = launch(fast_sample_kernel<1, false, 1, true, 32, 2>, grid, false); break;
      case 22: rc = launch(fast_sample_kernel<1, true, 1, true, 64, 2>, grid, false); break;
      default: set_error("internal: no kernel image for this mode"); rc = BESO_E_INVALID; break;
    }
    if (wide) {
      const cudaError_t fe = cudaFreeAsync(p.xscratch, st);
      if (rc == BESO_OK && fe != cudaSuccess) return cuda_fail(fe, "cudaFreeAsync(xscratch)");
    }
  }
  if (rc) return rc;
  ++g_kernel_launches;
  BESO_CUDA(cudaGetLastError());
  return BESO_OK;
}

}  // namespace beso
