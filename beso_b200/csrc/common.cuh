// Shared declarations for libbeso_b200.so (host side of the C ABI + kernel argument structs).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/beso_b200.h"

namespace beso {

constexpr int kMaxLayers = 16;
constexpr int kMaxSteps = 128;   // sampler steps per launch (n_sigmas - 1)

// ---- error plumbing -------------------------------------------------------------------
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
#define BESO_CUDA(expr)                                              \
  do {                                                               \
    cudaError_t _e = (expr);                                         \
    if (_e != cudaSuccess) return ::beso::cuda_fail(_e, #expr);      \
  } while (0)

extern long long g_kernel_launches;

// ---- PRECISE mode: transposed fp32 weights, [K][Npad] row-major ---------------------------
struct SimtLayer {
  const float *ln1w, *ln1b, *ln2w, *ln2b;
  // GEMM weights: transposed [K][N] images tiled by 64-column groups, [ceil(N/64)][K][64] (pack.cu)
  const float *wqkv, *bqkv;   // K = d,  N = 3d  columns: q | k | v  (score_gpts.py:33-35)
  const float *wproj, *bproj; // K = d,  N = d
  const float *w1, *b1;       // K = d,  N = 4d                      (mlp.0)
  const float *w2, *b2;       // K = 4d, N = d                       (mlp.2)
};
struct SimtModel {
  int obs, act, W, G, d, L, H, hs, linear_out, act_pad, hid, hid_pad;
  float sigma_data;
  const float *pos;            // [G+W+1][d]
  const float *tokw, *tokb;    // [obs][d]
  const float *sigw, *sigb;    // [d]
  const float *actw, *actb;    // [act][d]
  const float *lnfw, *lnfb;
  const float *hw0, *hb0;      // head: [d][act_pad]            (linear_output)  or [d][hid_pad]
  const float *hw1, *hb1;      //       unused                                   or [hid_pad][act_pad]
  SimtLayer layer[kMaxLayers];
};

// Per-launch sampler description (kernel parameter, 1.5 KB).
struct SampleArgs {
  int n_steps;   // 0 = single model evaluation with per-sequence sigma
  int sampler;   // BESO_SAMPLER_*
  float sig[kMaxSteps + 1];
  float ca[kMaxSteps];   // DDIM: sigma_fn(t_next) / sigma_fn(t);  Euler ancestral: sigma_down
  float ce[kMaxSteps];   // DDIM: expm1(-h);                        Euler ancestral: sigma_up
  float c1[kMaxSteps];   // DPM-Solver++(2M): 1 + 1/(2r)  (0 on first-order steps)
  float c2[kMaxSteps];   // DPM-Solver++(2M): 1/(2r)
  // generic two-stage sampler (BESO_SAMPLER_TWO_STAGE): u = a1 x + b1 D1 (a1 = ca, b1 = ce);
  // x = a2 x + b2 u + c2 D2 + su noise (a2 = c1, b2 = c2)
  float sigb[kMaxSteps]; // sigma of the second evaluation, 0 = single-stage step
  float c3[kMaxSteps];   // coefficient of D2
  float su[kMaxSteps];   // noise scale (0 = no noise)
  const float* noise;    // ancestral samplers: (n_steps, B, t, act) standard-normal draws of the caller
  long long noise_stride;   // elements per step = B * t * act
  // rollout I/O scaling fused into the loop's first read and last write (beso_io_scaling of the C ABI); all NULL = off
  const float* in_tab;      // (4, obs): sub, div, mul, add -- scale_input of states and goals
  const float* goal_keep;   // (obs): goal features multiplied after scaling (0 = zeroed goal dimension)
  const double* clip;       // (2, act): lo, hi of clip_action
  const float* out_tab;     // (4, act): inverse_scale_output as ((x - sub) / div) * mul + add
  float* unscaled;          // (B, t, act): clip + inverse scale of the final x
};

// ((x - sub) / div) * mul + add with separately rounded fp32 steps: bit-identical to the reference scaler's
// element-wise torch ops (networks/scaler/scaler_class.py:69-166); tab is (4, dim) row-major
__device__ __forceinline__ float io_scale(float x, const float* __restrict__ tab, int dim, int f) {
  const float q = __fdiv_rn(__fsub_rn(x, __ldg(tab + f)), __ldg(tab + dim + f));
  return __fadd_rn(__fmul_rn(q, __ldg(tab + 2 * dim + f)), __ldg(tab + 3 * dim + f));
}
// torch.clamp(y, lo, hi) with float64 bounds (y_bounds_tensor * 1.1 is float64), result rounded to fp32
__device__ __forceinline__ float io_clip(float y, const double* __restrict__ clip, int act, int a) {
  const double lo = clip[a], hi = clip[act + a], v = (double)y;
  return (float)(v < lo ? lo : (v > hi ? hi : v));
}

struct SimtLaunch {
  int B, t, S;           // S = sequences per CTA
  uint32_t flags;
  float cond_lambda;
  size_t smem_bytes;
};

int simt_plan_launch(const SimtModel& m, int t, int max_smem, SimtLaunch* out);
int simt_launch(const SimtModel& m, const SimtLaunch& L, const SampleArgs& sa, const float* state,
                const float* goal, const float* action_or_x, const float* sigma, float* out,
                cudaStream_t stream);

// weight packing helpers (pack.cu)
int pack_transpose(const float* src, int N, int K, float* dst, int ld_dst, int col0, cudaStream_t s);
int pack_transpose_tiled(const float* src, int N, int K, float* dst, int col0, cudaStream_t s);
int pack_copy(const float* src, float* dst, int64_t n, cudaStream_t s);

}  // namespace beso
