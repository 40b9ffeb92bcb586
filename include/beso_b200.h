/*
 * beso_b200.h -- C ABI of libbeso_b200.so, the B200 (sm_100a) implementation of the BESO
 * score-based action-denoising hot path.
 *
 * The reference (intuitive-robots/beso) has no FFI: its seam is duck-typed Python
 * (Hydra `_target_` classes).  Each entry point below names the reference interface it
 * replaces (paths relative to beso/agents/diffusion_agents/); the Python classes in
 * beso_b200/ that mirror those interfaces are thin ctypes callers of this ABI
 * (INTEGRATION.md shows the binding).
 *
 * Conventions
 *  - plain C types only; every tensor argument is a raw pointer to contiguous row-major
 *    fp32 unless stated otherwise; "dev" pointers are device memory, "host" pointers host.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *  - every function returns 0 on success and a negative BESO_E_* code on failure; the
 *    message is available from beso_last_error() (thread local).  Nothing throws.
 *  - a plan owns packed weights and workspaces; the caller owns all inputs and outputs.
 *    A plan is bound to one device and is not re-entrant.
 *  - no hidden device synchronisation except in the *_host entry points, which return
 *    after their result has landed in host memory.
 */
#ifndef BESO_B200_H_
#define BESO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BESO_ABI_VERSION 1

enum {
  BESO_OK = 0,
  BESO_E_INVALID = -1,     /* bad argument / unsupported shape                 */
  BESO_E_CUDA = -2,        /* CUDA runtime error (text in beso_last_error)      */
  BESO_E_UNSUPPORTED = -3, /* valid request, but not by the selected mode       */
  BESO_E_NOT_PACKED = -4,  /* plan has no weights yet                           */
  BESO_E_NCCL = -5
};

/* Arithmetic mode of the score-GPT GEMMs. */
enum {
  BESO_MODE_PRECISE = 0, /* fp32-equivalent arithmetic; meets rtol 1e-3 / atol 1e-5 vs the fp32 reference.  On shapes
                            the tensor-core kernel supports (below) every product runs on tcgen05 with both
                            operands split into fp16 hi + lo images (fp32 accumulate in TMEM; two tile layouts,
                            chosen per launch: 128-row tiles with three MMAs per product for embed_dim <= 256,
                            or 64 sequence rows with both images stacked on the MMA row dimension), fp32
                            two-pass LayerNorm, erf GELU in fp32, split-operand attention; other shapes run the
                            fp32 CUDA-core kernel.                                                       */
  BESO_MODE_FAST = 1,    /* fp16 operands on tcgen05 tensor cores (single pass), fp32 accumulate in TMEM; fp32
                            LayerNorm statistics / softmax / residual.  Shapes: embed_dim <= 384 (multiple of 8;
                            a 256-column and a 384-column geometry of the kernel -- the reference's block-push
                            (d = 240) and kitchen (d = 360) checkpoints), head size <= 64 with n_heads * padded
                            head size (32 or 64) <= 384, <= 24 tokens, obs <= 64, act <= 13, linear action head. */
  BESO_MODE_SIMT = 2     /* force the fp32 CUDA-core kernel (any shape); same tolerance as PRECISE        */
};

/* gc_sampling.py sampler selected by BesoAgent.sample_loop (beso_agent.py:419-455). */
enum {
  BESO_SAMPLER_DDIM = 0,  /* gc_sampling.py:895-924 */
  BESO_SAMPLER_EULER = 1, /* gc_sampling.py:167-213 (s_churn = 0) */
  BESO_SAMPLER_HEUN = 2,  /* gc_sampling.py:259-314 (s_churn = 0) */
  BESO_SAMPLER_EULER_ANCESTRAL = 3, /* gc_sampling.py:216-256; needs beso_sample_loop_noise */
  BESO_SAMPLER_DPMPP_2M = 4, /* gc_sampling.py:703-736 (DPM-Solver++(2M)); needs its coefficients in coef_host */
  BESO_SAMPLER_LMS = 6,      /* gc_sampling.py:416-468 (linear multistep, order <= 4); not with BESO_FLAG_CFG in fast mode */
  BESO_SAMPLER_TWO_STAGE = 5 /* generic single-step second-order sampler as a coefficient program: sample_dpm_2
                                (:317-377), sample_dpm_2_ancestral (:380-413), sample_dpmpp_2s (:928-967),
                                sample_dpmpp_2s_ancestral (:970-1016) */
};

/* flags */
#define BESO_FLAG_UNCOND 1u /* DiffusionGPT.forward(uncond=True): goals zeroed (score_gpts.py:301-302)      */
#define BESO_FLAG_CFG 2u    /* ClassifierFreeSampleModel.forward (classifier_free_sampler.py:35-49):
                               out_u + cond_lambda * (out_c - out_u), both branches in one launch           */
#define BESO_FLAG_PRED_LAST 8u /* GCDenoiser.loss(pred_last_action_only=True): noise only on, and loss only of, the
                                  last step (score_wrappers.py:59-66, 76-77)                                     */
#define BESO_FLAG_INNER 4u  /* return DiffusionGPT.forward(state, action, goal, sigma) itself, i.e. without
                               the c_in / c_out / c_skip pre-conditioning of GCDenoiser.forward             */
#define BESO_FLAG_TRAIN_FAST 16u /* beso_loss_fwd_bwd only: one bf16 tcgen05 MMA per product in the training GEMMs
                                  (the arithmetic of bf16 mixed-precision training).  Opt-in: by default every
                                  operand is split into three bf16 images (24 mantissa bits) and every product is
                                  six MMAs (all cross terms down to 2^-24, fp32 accumulate in TMEM) -- the
                                  fp32-parity mode the gradient goldens pin (the reference multiplies in fp32).
                                  Measured: gradients within ~1.5e-6 of their tensor's scale of the reference's. */
#define BESO_FLAG_TRAIN_SPLIT2 32u /* beso_loss_fwd_bwd only: two bf16 images per operand (16 mantissa bits), three
                                  MMAs per product; measured ~1.5e-5 of the gradient scale; opt-in */
#define BESO_FLAG_TRAIN_TF32 BESO_FLAG_TRAIN_FAST /* round-1 name of the opt-in tensor-core training mode */

/* Constructor arguments of DiffusionGPT (k_diffusion/score_gpts.py:121-139) and
 * GCDenoiser.sigma_data (k_diffusion/score_wrappers.py:26-29). */
typedef struct beso_model_desc {
  int32_t obs_dim;          /* state_dim                          */
  int32_t act_dim;          /* action_dim                         */
  int32_t window;           /* obs_seq_len  W                     */
  int32_t goal_len;         /* goal_seq_len G                     */
  int32_t d;                /* embed_dim                          */
  int32_t n_layers;
  int32_t n_heads;
  int32_t linear_output;    /* 1: action_pred = Linear(d, act)    */
  int32_t goal_conditioned; /* 0: goal tokens dropped             */
  float sigma_data;
} beso_model_desc;

typedef struct beso_plan beso_plan;
typedef struct beso_comm beso_comm;

const char* beso_last_error(void);
int beso_abi_version(void);

/* Number of parameter tensors / scalars in nn.Module.parameters() order of the reference
 * (score_gpts.py:150-190).  Returns <0 on an invalid description. */
int beso_param_count(const beso_model_desc* desc);
int64_t beso_param_numel(const beso_model_desc* desc, int index);
int64_t beso_param_total(const beso_model_desc* desc);

/* Replaces: hydra.utils.instantiate(GCDenoiser(DiffusionGPT(...))).to(device)
 * (score_wrappers.py:26-29, base_agent.py:31). */
int beso_plan_create(const beso_model_desc* desc, int device, beso_plan** out);
int beso_plan_destroy(beso_plan* plan);

/* Replaces: load_state_dict / optimizer.step / EMA copy_to making new weights visible to
 * forward (beso_agent.py:343-345,380-381,462).  `params_dev` holds beso_param_count()
 * device pointers to the fp32 parameter tensors in parameters() order.  Re-callable; the
 * packed images (fp16 UMMA tapes for FAST and -- as hi / lo pairs -- for PRECISE, transposed fp32 for the CUDA-core kernel) are rebuilt on
 * `stream`.  `slot` selects one of two resident weight sets (0 = raw, 1 = EMA) so that the
 * EMA swap in predict()/evaluate() does not force a re-pack. */
int beso_plan_pack_weights(beso_plan* plan, int slot, const float* const* params_dev, int n_params,
                           void* stream);
int beso_plan_select_weights(beso_plan* plan, int slot);
/* Registers the raw fp32 parameter pointers of `slot` without re-packing (training steps change the
 * weights every iteration; the loss path reads them directly). */
int beso_plan_set_params(beso_plan* plan, int slot, const float* const* params_dev, int n_params);

/* Replaces: GCDenoiser.forward (score_wrappers.py:81-96) -> DiffusionGPT.forward
 * (score_gpts.py:272-358); with BESO_FLAG_CFG also ClassifierFreeSampleModel.forward.
 *   state (B,t,obs)  action (B,t,act)  goal (B,G,obs)  sigma (B)  ->  out (B,t,act)
 * 1 <= t <= window.  One kernel launch. */
int beso_denoise_fwd(beso_plan* plan, int mode, const float* state_dev, const float* action_dev,
                     const float* goal_dev, const float* sigma_dev, float* out_dev, int B, int t,
                     uint32_t flags, float cond_lambda, void* stream);

/* Replaces: gc_sampling.sample_ddim / sample_euler / sample_heun called from
 * BesoAgent.sample_loop (beso_agent.py:390-456) with s_churn = 0, scaler = None,
 * callback = None.  The whole loop over `n_sigmas - 1` steps is ONE persistent kernel
 * launch; x_inout (B,t,act) holds x_t on entry and x_0 on return.
 *   sigmas_host: n_sigmas fp32 noise levels (last one may be 0), as from get_sigmas_*.
 *   coef_host:   DDIM only, 2*(n_sigmas-1) fp32: for step i, [2i] = sigma_fn(t_next)/sigma_fn(t)
 *                and [2i+1] = expm1(-h) evaluated by the caller exactly as gc_sampling.py:921-923
 *                does (NULL = evaluate on the host in this library with the same fp32 formulas). */
int beso_sample_loop(beso_plan* plan, int mode, int sampler, const float* sigmas_host, int n_sigmas,
                     const float* coef_host, const float* state_dev, const float* goal_dev,
                     float* x_inout_dev, int B, int t, uint32_t flags, float cond_lambda, void* stream);

/* beso_sample_loop with per-step noise, for the ancestral samplers.
 * BESO_SAMPLER_EULER_ANCESTRAL = sample_euler_ancestral (gc_sampling.py:216-256, the kitchen evaluation default,
 * configs/evaluate_kitchen.yaml:12):  x += to_d(x, sigma_i, D) * (sigma_down - sigma_i);  if sigma_down > 0:
 * x += noise_i * sigma_up.
 *   coef_host: 2*(n_sigmas-1) fp32, [2i] = sigma_down_i, [2i+1] = sigma_up_i as returned by get_ancestral_step
 *              (gc_sampling.py:108-114), evaluated by the caller with the reference's own fp32 tensor ops.
 *   noise_dev: (n_sigmas-1, B, t, act) fp32 = the torch.randn_like(action) draws of the reference, one per step
 *              (entries of steps with sigma_down == 0 are never read).  The caller draws them, in step order,
 *              so the result is bit-comparable with the reference under the same generator state.
 * BESO_SAMPLER_DPMPP_2M = sample_dpmpp_2m (gc_sampling.py:703-736): coef_host holds 4*(n_sigmas-1) fp32, per step
 *   [sigma_fn(t_next)/sigma_fn(t), expm1(-h), 1 + 1/(2r), 1/(2r)] evaluated by the caller with the reference's fp32
 *   tensor ops; the last two are 0 for the first-order steps (the first step and a step onto sigma = 0).
 *   noise_dev is not used.
 * BESO_SAMPLER_LMS = sample_lms (gc_sampling.py:431-468): x += c0 d_i + c1 d_{i-1} + c2 d_{i-2} + c3 d_{i-3} with
 *   d = to_d(x, sigma_i, D); coef_host holds 4*(n_sigmas-1) fp32 = linear_multistep_coeff(min(i+1, order), ...)
 *   (gc_sampling.py:416-428, scipy quad on the host), 0 for the derivatives that do not exist yet.
 * BESO_SAMPLER_TWO_STAGE: every step is   D1 = model(x, sigma_i);   if sigma_b == 0:  x = a1 x + b1 D1 + su noise_i
 *   else  u = a1 x + b1 D1;  D2 = model(u, sigma_b);  x = a2 x + b2 u + c2 D2 + su noise_i.
 *   coef_host holds 8*(n_sigmas-1) fp32, per step [sigma_b, a1, b1, a2, b2, c2, su, 0], computed by the caller from
 *   the sampler's formulas (beso_b200/sampling.py: dpm_2, dpm_2_ancestral, dpmpp_2s, dpmpp_2s_ancestral);
 *   noise_dev[i] is read only where su != 0.  The kernel applies the combined coefficients, so results agree
 *   with the reference's step-by-step arithmetic to fp32 rounding, not bit for bit. */
int beso_sample_loop_noise(beso_plan* plan, int mode, int sampler, const float* sigmas_host, int n_sigmas,
                           const float* coef_host, const float* state_dev, const float* goal_dev,
                           float* x_inout_dev, const float* noise_dev, int B, int t, uint32_t flags,
                           float cond_lambda, void* stream);

/* Same two calls with HOST buffers: inputs are staged through plan-owned pinned and device
 * buffers, the kernel runs on `stream`, the result is copied back and the call returns once it
 * is in `out_host` / `x_inout_host`.  This is the end-to-end path bench.py times as "e2e". */
int beso_denoise_fwd_host(beso_plan* plan, int mode, const float* state_host, const float* action_host,
                          const float* goal_host, const float* sigma_host, float* out_host, int B, int t,
                          uint32_t flags, float cond_lambda, void* stream);
int beso_sample_loop_host(beso_plan* plan, int mode, int sampler, const float* sigmas_host, int n_sigmas,
                          const float* coef_host, const float* state_host, const float* goal_host,
                          float* x_inout_host, int B, int t, uint32_t flags, float cond_lambda,
                          void* stream);

/* The rollout path's scaler calls fused into the sample loop.  Replaces, around
 * BesoAgent.sample_loop in predict() / evaluate() (beso_agent.py:322-329,373-387; base_agent.py:111-142):
 * scaler.scale_input(state / goal) and the zeroed block-push goal dimensions (base_agent.py:119-120) on the loop's
 * first read, scaler.clip_action and scaler.inverse_scale_output (networks/scaler/scaler_class.py:69-166) on its last
 * write.  Tables are (4, dim) row-major fp32 with rows (sub, div, mul, add): y = ((x - sub) / div) * mul + add in
 * separately rounded fp32 steps (bit-identical to the scaler's element-wise ops).  Any member may be NULL = skipped.
 *   in_table_dev    (4, obs_dim)  applied to every state and goal feature
 *   goal_keep_dev   (obs_dim)     goal features are multiplied by it after scaling (0 = zeroed dimension)
 *   out_clip_dev    (2, act_dim)  float64 lo, hi: x_inout is clamped to [lo, hi] before it is written back
 *   out_table_dev   (4, act_dim)  inverse output scaling, written to unscaled_out_dev
 *   unscaled_out_dev (B, t, act)  clip + inverse scaling of the final x (x_inout itself stays in scaled units: the
 *                                 agent feeds it back as action context) */
typedef struct beso_io_scaling {
  const float* in_table_dev;
  const float* goal_keep_dev;
  const double* out_clip_dev;
  const float* out_table_dev;
  float* unscaled_out_dev;
} beso_io_scaling;
int beso_sample_loop_scaled(beso_plan* plan, int mode, int sampler, const float* sigmas_host, int n_sigmas,
                            const float* coef_host, const float* state_dev, const float* goal_dev, float* x_inout_dev,
                            const float* noise_dev, const beso_io_scaling* io, int B, int t, uint32_t flags,
                            float cond_lambda, void* stream);

/* Replaces: GCDenoiser.loss (score_wrappers.py:45-79) + loss.backward() (beso_agent.py:236-240).
 *   action = clean action a; noise n; sigma (B).  goal_keep_dev: optional (B,G,obs) {0,1} mask =
 *   1 - Bernoulli(cond_mask_prob) drawn by the caller (score_gpts.py:360-371), NULL = keep all.
 *   loss_dev: 1 fp32.  flat_grad_dev: beso_param_total() fp32 in parameters() order, overwritten
 *   (NULL = forward only).  All dense products run on tcgen05 (BESO_FLAG_TRAIN_FAST selects single-pass bf16). */
int beso_loss_fwd_bwd(beso_plan* plan, const float* state_dev, const float* action_dev,
                      const float* goal_dev, const float* noise_dev, const float* sigma_dev,
                      const float* goal_keep_dev, float* loss_dev, float* flat_grad_dev, int B,
                      uint32_t flags, void* stream);

/* Training-mode dropout of the reference (score_gpts.py:37-38,72,79,109,338): nn.Dropout draws from torch's global
 * generator in op order, so the caller draws the masks with the same torch calls in the same order (SURVEY.md H5;
 * beso_b200/training.py does) and hands them over.  Every mask holds 0 or 1 / (1 - p) (what F.dropout applies to a
 * tensor of ones).  Any pointer (or the struct itself) may be NULL = that dropout is off (p = 0).
 *   embed:       (B, T, d)      self.drop(input_seq)                       embed_pdrob
 *   attn[l]:     (B, H, T, T)   attn_drop(softmax(...)) of block l          attn_pdrop
 *   resid_attn:  (B, T, d)      resid_drop(proj(y)) of block l              resid_pdrop
 *   resid_mlp:   (B, T, d)      the Dropout that ends block l's mlp         resid_pdrop
 * T = 1 + G + 2 W tokens; attn / resid_attn / resid_mlp are arrays of n_layers device pointers. */
typedef struct beso_dropout_masks {
  const float* embed;
  const float* const* attn;
  const float* const* resid_attn;
  const float* const* resid_mlp;
} beso_dropout_masks;
int beso_loss_fwd_bwd_dropout(beso_plan* plan, const float* state_dev, const float* action_dev,
                              const float* goal_dev, const float* noise_dev, const float* sigma_dev,
                              const float* goal_keep_dev, const beso_dropout_masks* masks, float* loss_dev,
                              float* flat_grad_dev, int B, uint32_t flags, void* stream);

/* The training GEMM by itself (tests and tools): C[M][N] = A . B^T + bias, fp32 row-major device tensors.
 *   a_kmajor: A element (m, k) at A[m * lda + k], else at A[k * lda + m]; b_kmajor likewise for B (n, k).
 *   prec: 2 = three bf16 images, six MMAs per product (fp32-parity); 1 = two images, three MMAs; 0 = one bf16 MMA.
 *   bias may be NULL. */
int beso_debug_gemm(beso_plan* plan, const float* A_dev, int lda, int a_kmajor, const float* B_dev, int ldb,
                    int b_kmajor, float* C_dev, int ldc, int M, int N, int K, const float* bias_dev, int accumulate,
                    int prec, void* stream);

/* Data-parallel gradient step (BASELINE config 4): one all-reduce(sum) of the flat fp32 gradient
 * over NCCL on NVLink, scaled by 1/world.  The reference has no distributed path; this is the
 * exchange step of SURVEY.md 8e.  unique_id: 128 bytes from beso_comm_unique_id on rank 0. */
int beso_comm_unique_id(char* out128);
int beso_comm_init(int rank, int world, const char* unique_id128, int device, beso_comm** out);
int beso_comm_destroy(beso_comm* comm);
int beso_allreduce_grads(beso_comm* comm, float* flat_grad_dev, size_t n, float scale, void* stream);

/* beso_loss_fwd_bwd_dropout with the gradient exchange overlapped with the backward pass: as soon as a transformer
 * block's 16 gradient tensors are final (blocks finish last-to-first) their slice of the flat buffer is all-reduced
 * (sum, then * grad_scale: pass 1 / world) on the communicator's own stream behind an event, while `stream` goes on with
 * the next block's backward; the embedding and tail tensors follow at the end.  On return `stream` is ordered after
 * every bucket, so flat_grad_dev holds the global-batch mean gradient for whatever is launched on it next.
 * comm == NULL or world 1: identical to beso_loss_fwd_bwd_dropout. */
int beso_loss_fwd_bwd_dp(beso_plan* plan, const float* state_dev, const float* action_dev, const float* goal_dev,
                         const float* noise_dev, const float* sigma_dev, const float* goal_keep_dev,
                         const beso_dropout_masks* masks, beso_comm* comm, float grad_scale, float* loss_dev,
                         float* flat_grad_dev, int B, uint32_t flags, void* stream);

/* Fused optimiser step.  Replaces, in BesoAgent.train_step (beso_agent.py:238-247),
 * optimizer.step() of torch.optim.AdamW (configs/agents/beso_kitchen.yaml:9-12) and
 * ExponentialMovingAverage.update (beso/networks/ema_helper/ema.py:36-53): ONE launch over all parameter
 * tensors, every element read and written once, arithmetic in the order of torch's single-tensor AdamW.
 *   beso_opt_create: param_dev_ptrs / numel describe the n_tensors fp32 parameter tensors in
 *     parameters() order (the order of the flat gradient of beso_loss_fwd_bwd); they are updated in place.
 *   beso_opt_step: flat_grad_dev, exp_avg_dev, exp_avg_sq_dev, ema_dev are flat fp32 buffers of
 *     beso_opt_total() elements in the same order, owned by the caller.  step >= 1 is AdamW's step count
 *     (bias correction), lr the current learning rate (the caller applies StepLR), ema_dev == NULL skips
 *     the EMA, ema_decay is the decay of THIS update (the caller applies the warm-up min(decay,
 *     (1+n)/(10+n)) of ema.py:47-50); grad_scale multiplies the gradient first (1/world after an
 *     all-reduce(sum), 1 otherwise). */
typedef struct beso_opt beso_opt;
int beso_opt_create(int device, int n_tensors, float* const* param_dev_ptrs, const long long* numel, beso_opt** out);
int beso_opt_destroy(beso_opt* opt);
long long beso_opt_total(const beso_opt* opt);
int beso_opt_step(beso_opt* opt, const float* flat_grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev,
                  float* ema_dev, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                  float ema_decay, float grad_scale, void* stream);

/* Windowed batch gather from GPU-resident trajectories.  Replaces TrajectorySlicerDataset.__getitem__
 * (beso/envs/dataloaders/trajectory_loader.py:160-197) + DataLoader collation + the host-to-device copy of
 * BesoAgent.train_step's batch with one launch.  obs_dev (n_traj, t_max, obs_dim) and act_dev (n_traj, t_max, act_dim)
 * are the padded trajectories; for sample b: traj_dev[b], start_dev[b] (window start) and goal_start_dev[b] (start of
 * the future-observation window, < 0 = the reference's zeros placeholder).  goal_out_dev == NULL: no goal window.
 * obs_scale_dev (4, obs_dim) / act_scale_dev (4, act_dim), each NULL or rows (sub, div, mul, add): the Scaler's
 * scale_input / scale_output (beso/networks/scaler/scaler_class.py:79-112, 271-301) applied on the way as
 * ((x - sub) / div) * mul + add, each step rounded to fp32 like the reference's elementwise ops.
 * The caller guarantees start + window <= t_max and goal_start + goal_len <= t_max (beso_b200/dataset.py). */
int beso_window_gather(const float* obs_dev, const float* act_dev, int n_traj, int t_max, int obs_dim, int act_dim,
                       const int* traj_dev, const int* start_dev, const int* goal_start_dev, int window, int goal_len,
                       float* state_out_dev, float* action_out_dev, float* goal_out_dev,
                       const float* obs_scale_dev, const float* act_scale_dev, int B, void* stream);

/* Introspection used by tests and bench.py. */
int64_t beso_kernel_launches(void);                /* kernels launched by this library so far   */
int beso_plan_rows_per_cta(beso_plan* plan, int mode, int t); /* sequences handled per CTA      */
int beso_device_sm_count(int device);
/* Diagnostics / tests: tile layout of the PRECISE mode's tensor-core kernel.  0 = chosen per launch (whichever needs
 * less time for the batch), 1 = stacked (64 sequence rows per tile, any supported shape), 2 = 128-row tiles
 * (embed_dim <= 256; ignored where unsupported).  Also settable as BESO_PREC_LAYOUT=stacked|p128. */
int beso_debug_set_precise_layout(int layout);
/* Diagnostics: when trace_dev != NULL, FAST-mode launches dump the fp32 residual stream of tile 0,
 * first evaluation, as seen by every LayerNorm pass: (2 * n_layers + 1) x 128 x 256 floats. */
int beso_debug_set_trace(float* trace_dev);
/* Diagnostics: clock64 stamps of block 0 during its second model evaluation (tools/timeline_fast.py). */
int beso_debug_set_timeline(long long* timeline_dev);
/* Diagnostics: tcgen05.mma issue / completion cycles for 10 instruction mixes (20 int64; tools/mma_rate.py). */
int beso_debug_mma_rate(long long* out_dev, const void* src_6mb_dev, int mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BESO_B200_H_ */
