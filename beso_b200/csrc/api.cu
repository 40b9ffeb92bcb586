// C ABI of libbeso_b200.so (see include/beso_b200.h for the contract of every entry point).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "fast.cuh"
#include "gemm.cuh"
#include "plan.cuh"

namespace beso {

static thread_local std::string t_error;
long long g_kernel_launches = 0;

void set_error(const std::string& msg) { t_error = msg; }
int cuda_fail(cudaError_t e, const char* what) {
  t_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " at " + what;
  return BESO_E_CUDA;
}

// parameters() order of the reference DiffusionGPT (score_gpts.py:150-190)
struct ParamInfo { int64_t numel; int rows, cols; };   // Linear weight: rows = out, cols = in

static bool valid_desc(const beso_model_desc* m) {
  if (!m) return false;
  if (m->obs_dim < 1 || m->act_dim < 1 || m->window < 1 || m->goal_len < 0 || m->d < 4 || m->n_layers < 1 ||
      m->n_layers > kMaxLayers || m->n_heads < 1)
    return false;
  if (m->d % m->n_heads != 0 || m->d % 4 != 0 || (m->d / m->n_heads) % 4 != 0) return false;
  if (!(m->sigma_data > 0.f)) return false;
  return true;
}

static std::vector<ParamInfo> param_table(const beso_model_desc& m) {
  const int d = m.d, G = m.goal_conditioned ? m.goal_len : 0;
  std::vector<ParamInfo> v;
  auto mat = [&](int r, int c) { v.push_back({(int64_t)r * c, r, c}); };
  auto vec = [&](int n) { v.push_back({n, n, 1}); };
  v.push_back({(int64_t)(G + m.window + 1) * d, G + m.window + 1, d});   // pos_emb
  mat(d, m.obs_dim); vec(d);                                            // tok_emb
  for (int l = 0; l < m.n_layers; ++l) {
    vec(d); vec(d); vec(d); vec(d);                                     // ln1.w ln1.b ln2.w ln2.b
    for (int i = 0; i < 4; ++i) { mat(d, d); vec(d); }                  // key query value proj
    mat(4 * d, d); vec(4 * d); mat(d, 4 * d); vec(d);                   // mlp.0 mlp.2
  }
  vec(d); vec(d);                                                       // ln_f
  mat(d, 1); vec(d);                                                    // sigma_emb
  mat(d, m.act_dim); vec(d);                                            // action_emb
  if (m.linear_output) { mat(m.act_dim, d); vec(m.act_dim); }
  else { mat(100, d); vec(100); mat(m.act_dim, 100); vec(m.act_dim); }
  return v;
}

}  // namespace beso

using namespace beso;

namespace beso {

static size_t simt_layout(const beso_model_desc& m, float* base, SimtModel* out) {
  const int d = m.d, G = m.goal_conditioned ? m.goal_len : 0;
  size_t off = 0;
  auto take = [&](size_t n) { float* p = base ? base + off : nullptr; off += (n + 3) & ~size_t(3); return (const float*)p; };
  SimtModel s{};
  s.obs = m.obs_dim; s.act = m.act_dim; s.W = m.window; s.G = G; s.d = d; s.L = m.n_layers; s.H = m.n_heads;
  s.hs = d / m.n_heads; s.linear_out = m.linear_output; s.act_pad = (m.act_dim + 3) & ~3; s.hid = 100; s.hid_pad = 100;
  s.sigma_data = m.sigma_data;
  s.pos = take((size_t)(G + m.window + 1) * d);
  s.tokw = take((size_t)m.obs_dim * d); s.tokb = take(d);
  for (int l = 0; l < m.n_layers; ++l) {
    SimtLayer& L = s.layer[l];
    L.ln1w = take(d); L.ln1b = take(d); L.ln2w = take(d); L.ln2b = take(d);
    auto tiled = [](int K, int N) { return (size_t)((N + 63) / 64) * K * 64; };   // [ceil(N/64)][K][64]
    L.wqkv = take(tiled(d, 3 * d)); L.bqkv = take(3 * d);
    L.wproj = take(tiled(d, d)); L.bproj = take(d);
    L.w1 = take(tiled(d, 4 * d)); L.b1 = take(4 * d);
    L.w2 = take(tiled(4 * d, d)); L.b2 = take(d);
  }
  s.lnfw = take(d); s.lnfb = take(d);
  s.sigw = take(d); s.sigb = take(d);
  s.actw = take((size_t)m.act_dim * d); s.actb = take(d);
  if (m.linear_output) {
    s.hw0 = take((size_t)d * s.act_pad); s.hb0 = take(s.act_pad); s.hw1 = nullptr; s.hb1 = nullptr;
  } else {
    s.hw0 = take((size_t)d * s.hid_pad); s.hb0 = take(s.hid_pad);
    s.hw1 = take((size_t)s.hid_pad * s.act_pad); s.hb1 = take(s.act_pad);
  }
  if (out) *out = s;
  return off;
}

static int pack_simt(beso_plan* p, WeightSlot& ws, const float* const* prm, cudaStream_t st) {
  const beso_model_desc& m = p->desc;
  const int d = m.d, G = m.goal_conditioned ? m.goal_len : 0;
  const SimtModel& s = ws.simt;
  auto F = [](const float* q) { return const_cast<float*>(q); };
  int i = 0, rc;
#define TR(dst, N, K, ld, col) if ((rc = pack_transpose(prm[i++], N, K, F(dst), ld, col, st))) return rc
#define TT(dst, N, K, col) if ((rc = pack_transpose_tiled(prm[i++], N, K, F(dst), col, st))) return rc
#define CP(dst, n) if ((rc = pack_copy(prm[i++], F(dst), n, st))) return rc
  BESO_CUDA(cudaMemsetAsync(ws.simt_buf, 0, p->simt_floats * sizeof(float), st));
  CP(s.pos, (int64_t)(G + m.window + 1) * d);
  TR(s.tokw, d, m.obs_dim, d, 0); CP(s.tokb, d);
  for (int l = 0; l < m.n_layers; ++l) {
    const SimtLayer& L = s.layer[l];
    CP(L.ln1w, d); CP(L.ln1b, d); CP(L.ln2w, d); CP(L.ln2b, d);
    // reference parameter order is key, query, value, proj; packed column order is q | k | v
    TT(L.wqkv, d, d, d);            CP(L.bqkv + d, d);       // key
    TT(L.wqkv, d, d, 0);            CP(L.bqkv, d);           // query
    TT(L.wqkv, d, d, 2 * d);        CP(L.bqkv + 2 * d, d);   // value
    TT(L.wproj, d, d, 0);           CP(L.bproj, d);
    TT(L.w1, 4 * d, d, 0);          CP(L.b1, 4 * d);
    TT(L.w2, d, 4 * d, 0);          CP(L.b2, d);
  }
  CP(s.lnfw, d); CP(s.lnfb, d);
  CP(s.sigw, d); CP(s.sigb, d);                               // (d,1) weight is already a d-vector
  TR(s.actw, d, m.act_dim, d, 0); CP(s.actb, d);
  if (m.linear_output) {
    TR(s.hw0, m.act_dim, d, s.act_pad, 0); CP(s.hb0, m.act_dim);
  } else {
    TR(s.hw0, 100, d, s.hid_pad, 0); CP(s.hb0, 100);
    TR(s.hw1, m.act_dim, 100, s.act_pad, 0); CP(s.hb1, m.act_dim);
  }
#undef TR
#undef TT
#undef CP
  return BESO_OK;
}

static int make_sample_args(int sampler, const float* sig, int n_sigmas, const float* coef, SampleArgs* sa) {
  if (sampler < BESO_SAMPLER_DDIM || sampler > BESO_SAMPLER_LMS) { set_error("unknown sampler"); return BESO_E_INVALID; }
  if (sampler >= BESO_SAMPLER_EULER_ANCESTRAL && !coef) {
    set_error("euler_ancestral / dpmpp_2m / lms / two-stage samplers need their per-step coefficients (coef_host)"); return BESO_E_INVALID;
  }
  const int cstride = sampler == BESO_SAMPLER_TWO_STAGE ? 8 : ((sampler == BESO_SAMPLER_DPMPP_2M || sampler == BESO_SAMPLER_LMS) ? 4 : 2);
  if (!sig || n_sigmas < 2 || n_sigmas - 1 > kMaxSteps) { set_error("n_sigmas must be in [2, 129]"); return BESO_E_INVALID; }
  memset(sa, 0, sizeof(*sa));
  sa->n_steps = n_sigmas - 1;
  sa->sampler = sampler;
  for (int i = 0; i < n_sigmas; ++i) sa->sig[i] = sig[i];
  for (int i = 0; i < n_sigmas - 1; ++i) {
    if (!(sig[i] > 0.f)) { set_error("sigmas must be positive except the last"); return BESO_E_INVALID; }
    if (coef && cstride == 8) {
      const float* r = coef + 8 * i;
      sa->sigb[i] = r[0]; sa->ca[i] = r[1]; sa->ce[i] = r[2]; sa->c1[i] = r[3]; sa->c2[i] = r[4]; sa->c3[i] = r[5]; sa->su[i] = r[6];
      if (r[0] < 0.f) { set_error("two-stage sampler: sigma_b must be >= 0"); return BESO_E_INVALID; }
    } else if (coef) {
      sa->ca[i] = coef[cstride * i]; sa->ce[i] = coef[cstride * i + 1];
      if (cstride == 4) { sa->c1[i] = coef[4 * i + 2]; sa->c2[i] = coef[4 * i + 3]; }
    }
    else {
      // gc_sampling.py:913-923 in fp32: t = -log(sigma), h = t_next - t
      const float t = -logf(sig[i]), tn = -logf(sig[i + 1]);     // -log(0) = +inf
      const float h = tn - t;
      sa->ca[i] = expf(-tn) / expf(-t);
      sa->ce[i] = expm1f(-h);
    }
  }
  return BESO_OK;
}

static int check_call(beso_plan* p, int mode, int B, int t, uint32_t flags) {
  if (!p) { set_error("null plan"); return BESO_E_INVALID; }
  if (B < 1 || t < 1 || t > p->desc.window) { set_error("need B >= 1 and 1 <= t <= window"); return BESO_E_INVALID; }
  if (!p->slot[p->active].packed) { set_error("weights not packed (call beso_plan_pack_weights)"); return BESO_E_NOT_PACKED; }
  if (mode != BESO_MODE_PRECISE && mode != BESO_MODE_FAST && mode != BESO_MODE_SIMT) { set_error("unknown mode"); return BESO_E_INVALID; }
  if ((flags & BESO_FLAG_CFG) && (flags & BESO_FLAG_UNCOND)) { set_error("CFG and UNCOND are exclusive"); return BESO_E_INVALID; }
  if (mode == BESO_MODE_FAST && !p->fast_ok) {
    set_error("fast (tcgen05) mode needs embed_dim <= 384, head size <= 64 (<= 6 attention passes), linear_output, <= 24 tokens, "
              "obs <= 64, act <= 13; use precise mode");
    return BESO_E_UNSUPPORTED;
  }
  return BESO_OK;
}

static int run(beso_plan* p, int mode, const SampleArgs& sa, const float* state, const float* goal,
               const float* x, const float* sigma, float* out, int B, int t, uint32_t flags, float lambda,
               cudaStream_t st) {
  BESO_CUDA(cudaSetDevice(p->device));
  WeightSlot& ws = p->slot[p->active];
  if (mode == BESO_MODE_FAST)
    return fast_launch(ws.fast, p->desc, p->sm_count, sa, state, goal, x, sigma, out, B, t, flags, lambda, st, false);
  // PRECISE: split-operand tensor-core kernel where the shape allows (the CFG + LMS history needs a buffer that
  // kernel uses as CFG scratch), else the fp32 CUDA-core kernel
  if (mode == BESO_MODE_PRECISE && p->fast_ok && !p->force_simt &&
      !(sa.n_steps && sa.sampler == BESO_SAMPLER_LMS && (flags & BESO_FLAG_CFG)))
    return fast_launch(ws.fastp, p->desc, p->sm_count, sa, state, goal, x, sigma, out, B, t, flags, lambda, st, true,
                       ws.fastq.tape ? &ws.fastq : nullptr);
  SimtLaunch L{};
  int rc = simt_plan_launch(ws.simt, t, p->max_smem, &L);
  if (rc) return rc;
  L.B = B; L.flags = flags; L.cond_lambda = lambda;
  return simt_launch(ws.simt, L, sa, state, goal, x, sigma, out, st);
}

}  // namespace beso

extern "C" {

const char* beso_last_error(void) { return t_error.c_str(); }
int beso_abi_version(void) { return BESO_ABI_VERSION; }

int beso_param_count(const beso_model_desc* d) {
  if (!valid_desc(d)) { set_error("invalid model description"); return BESO_E_INVALID; }
  return (int)param_table(*d).size();
}
int64_t beso_param_numel(const beso_model_desc* d, int index) {
  if (!valid_desc(d)) { set_error("invalid model description"); return BESO_E_INVALID; }
  auto v = param_table(*d);
  if (index < 0 || index >= (int)v.size()) { set_error("parameter index out of range"); return BESO_E_INVALID; }
  return v[index].numel;
}
int64_t beso_param_total(const beso_model_desc* d) {
  if (!valid_desc(d)) { set_error("invalid model description"); return BESO_E_INVALID; }
  int64_t n = 0;
  for (auto& e : param_table(*d)) n += e.numel;
  return n;
}

int beso_device_sm_count(int device) {
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return BESO_E_CUDA;
  return n;
}

int beso_plan_create(const beso_model_desc* desc, int device, beso_plan** out) {
  if (!out) { set_error("null out"); return BESO_E_INVALID; }
  *out = nullptr;
  if (!valid_desc(desc)) {
    set_error("invalid model description (need d % n_heads == 0, d % 4 == 0, head_dim % 4 == 0, n_layers <= 16)");
    return BESO_E_INVALID;
  }
  BESO_CUDA(cudaSetDevice(device));
  beso_plan* p = new beso_plan();
  p->desc = *desc;
  p->device = device;
  BESO_CUDA(cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device));
  BESO_CUDA(cudaDeviceGetAttribute(&p->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  p->simt_floats = simt_layout(*desc, nullptr, nullptr);
  p->fast_ok = fast_supported(*desc);
  { const char* e = getenv("BESO_PRECISE_SIMT"); p->force_simt = e && atoi(e) != 0; }
  *out = p;
  return BESO_OK;
}

int beso_plan_destroy(beso_plan* p) {
  if (!p) return BESO_OK;
  cudaSetDevice(p->device);
  for (auto& s : p->slot) {
    if (s.simt_buf) cudaFree(s.simt_buf);
    fast_free(s.fast);
    fast_free(s.fastp);
    fast_free(s.fastq);
  }
  if (p->h_pin) cudaFreeHost(p->h_pin);
  if (p->d_stage) cudaFree(p->d_stage);
  train_ws_free(p->train_ws);
  delete p;
  return BESO_OK;
}

int beso_plan_pack_weights(beso_plan* p, int slot, const float* const* prm, int n_params, void* stream) {
  if (!p || !prm) { set_error("null argument"); return BESO_E_INVALID; }
  if (slot < 0 || slot > 1) { set_error("slot must be 0 or 1"); return BESO_E_INVALID; }
  if (n_params != (int)param_table(p->desc).size()) { set_error("wrong number of parameter tensors"); return BESO_E_INVALID; }
  for (int i = 0; i < n_params; ++i) if (!prm[i]) { set_error("null parameter pointer"); return BESO_E_INVALID; }
  BESO_CUDA(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  WeightSlot& ws = p->slot[slot];
  if (!ws.simt_buf) {
    BESO_CUDA(cudaMalloc(&ws.simt_buf, p->simt_floats * sizeof(float)));
    simt_layout(p->desc, ws.simt_buf, &ws.simt);
  }
  ws.params.assign(prm, prm + n_params);
  int rc = pack_simt(p, ws, prm, st);
  if (rc) return rc;
  if (p->fast_ok) {
    rc = fast_pack(ws.fast, p->desc, prm, st, FAST_LAYOUT_F16);
    if (rc) return rc;
    rc = fast_pack(ws.fastp, p->desc, prm, st, FAST_LAYOUT_STACKED);
    if (rc) return rc;
    if (fast_p128_supported(p->desc)) {
      rc = fast_pack(ws.fastq, p->desc, prm, st, FAST_LAYOUT_P128);
      if (rc) return rc;
    }
  }
  ws.packed = true;
  return BESO_OK;
}

int beso_plan_select_weights(beso_plan* p, int slot) {
  if (!p || slot < 0 || slot > 1) { set_error("bad plan or slot"); return BESO_E_INVALID; }
  if (!p->slot[slot].packed) { set_error("slot has no packed weights"); return BESO_E_NOT_PACKED; }
  p->active = slot;
  return BESO_OK;
}

int beso_denoise_fwd(beso_plan* p, int mode, const float* state, const float* action, const float* goal,
                     const float* sigma, float* out, int B, int t, uint32_t flags, float lambda, void* stream) {
  int rc = check_call(p, mode, B, t, flags);
  if (rc) return rc;
  if (!state || !action || !sigma || !out || (!goal && p->desc.goal_conditioned && p->desc.goal_len > 0)) {
    set_error("null tensor pointer"); return BESO_E_INVALID;
  }
  SampleArgs sa;
  memset(&sa, 0, sizeof(sa));
  return run(p, mode, sa, state, goal, action, sigma, out, B, t, flags, lambda, (cudaStream_t)stream);
}

// Validation shared by every sample-loop entry point (device and host buffers): a sampler that would read noise it was
// not given, or a mode / sampler combination a kernel cannot run, is an error code here, never a fault on the device.
static int check_sampler(beso_plan* p, int mode, int sampler, const SampleArgs& sa, bool have_noise, uint32_t flags) {
  if (flags & BESO_FLAG_INNER) { set_error("INNER is not meaningful for a sample loop"); return BESO_E_INVALID; }
  if (sampler == BESO_SAMPLER_EULER_ANCESTRAL || sampler == BESO_SAMPLER_TWO_STAGE) {
    bool needs_noise = false;
    for (int i = 0; i < sa.n_steps; ++i) needs_noise |= (sampler == BESO_SAMPLER_TWO_STAGE ? sa.su[i] != 0.f : sa.ca[i] > 0.f);
    if (needs_noise && !have_noise) { set_error("this sampler needs the per-step noise (beso_sample_loop_noise)"); return BESO_E_INVALID; }
  }
  if (sampler == BESO_SAMPLER_LMS && mode == BESO_MODE_FAST && (flags & BESO_FLAG_CFG)) {
    set_error("lms with classifier-free guidance is not available in fast mode (history buffers)"); return BESO_E_UNSUPPORTED;
  }
  (void)p;
  return BESO_OK;
}

int beso_sample_loop_scaled(beso_plan* p, int mode, int sampler, const float* sigmas, int n_sigmas, const float* coef,
                            const float* state, const float* goal, float* x, const float* noise, const beso_io_scaling* io,
                            int B, int t, uint32_t flags, float lambda, void* stream) {
  int rc = check_call(p, mode, B, t, flags);
  if (rc) return rc;
  if (!state || !x || (!goal && p->desc.goal_conditioned && p->desc.goal_len > 0)) {
    set_error("null tensor pointer"); return BESO_E_INVALID;
  }
  SampleArgs sa;
  rc = make_sample_args(sampler, sigmas, n_sigmas, coef, &sa);
  if (rc) return rc;
  rc = check_sampler(p, mode, sampler, sa, noise != nullptr, flags);
  if (rc) return rc;
  sa.noise = noise;
  sa.noise_stride = (long long)B * t * p->desc.act_dim;
  if (io) {
    if (io->out_table_dev && !io->unscaled_out_dev) { set_error("io scaling: out_table_dev needs unscaled_out_dev"); return BESO_E_INVALID; }
    sa.in_tab = io->in_table_dev; sa.goal_keep = io->goal_keep_dev; sa.clip = io->out_clip_dev;
    sa.out_tab = io->out_table_dev; sa.unscaled = io->unscaled_out_dev;
  }
  return run(p, mode, sa, state, goal, x, nullptr, x, B, t, flags, lambda, (cudaStream_t)stream);
}

int beso_sample_loop_noise(beso_plan* p, int mode, int sampler, const float* sigmas, int n_sigmas, const float* coef,
                           const float* state, const float* goal, float* x, const float* noise, int B, int t,
                           uint32_t flags, float lambda, void* stream) {
  return beso_sample_loop_scaled(p, mode, sampler, sigmas, n_sigmas, coef, state, goal, x, noise, nullptr, B, t, flags, lambda, stream);
}

int beso_sample_loop(beso_plan* p, int mode, int sampler, const float* sigmas, int n_sigmas, const float* coef,
                     const float* state, const float* goal, float* x, int B, int t, uint32_t flags, float lambda,
                     void* stream) {
  return beso_sample_loop_noise(p, mode, sampler, sigmas, n_sigmas, coef, state, goal, x, nullptr, B, t, flags, lambda, stream);
}

static int ensure_stage(beso_plan* p, size_t floats) {
  if (floats <= p->stage_floats) return BESO_OK;
  if (p->h_pin) cudaFreeHost(p->h_pin);
  if (p->d_stage) cudaFree(p->d_stage);
  p->h_pin = nullptr; p->d_stage = nullptr; p->stage_floats = 0;
  BESO_CUDA(cudaMallocHost(&p->h_pin, floats * sizeof(float)));
  BESO_CUDA(cudaMalloc(&p->d_stage, floats * sizeof(float)));
  p->stage_floats = floats;
  return BESO_OK;
}

static int host_call(beso_plan* p, int mode, const SampleArgs& sa, const float* state, const float* action,
                     const float* goal, const float* sigma, float* out, int B, int t, uint32_t flags, float lambda,
                     cudaStream_t st) {
  const beso_model_desc& m = p->desc;
  const int G = m.goal_conditioned ? m.goal_len : 0;
  auto r4 = [](size_t n) { return (n + 3) & ~size_t(3); };
  const size_t n_state = (size_t)B * t * m.obs_dim, n_goal = (size_t)B * G * m.obs_dim,
               n_act = (size_t)B * t * m.act_dim, n_sig = (size_t)B;
  const size_t o_state = 0, o_goal = o_state + r4(n_state), o_act = o_goal + r4(n_goal), o_sig = o_act + r4(n_act),
               o_out = o_sig + r4(n_sig), total = o_out + r4(n_act);
  BESO_CUDA(cudaSetDevice(p->device));
  int rc = ensure_stage(p, total);
  if (rc) return rc;
  // pageable -> pinned staging (a caller with pinned buffers pays only this memcpy), then one H2D
  memcpy(p->h_pin + o_state, state, n_state * sizeof(float));
  if (n_goal) memcpy(p->h_pin + o_goal, goal, n_goal * sizeof(float));
  memcpy(p->h_pin + o_act, action, n_act * sizeof(float));
  if (sigma) memcpy(p->h_pin + o_sig, sigma, n_sig * sizeof(float));
  BESO_CUDA(cudaMemcpyAsync(p->d_stage, p->h_pin, o_out * sizeof(float), cudaMemcpyHostToDevice, st));
  float* d = p->d_stage;
  rc = run(p, mode, sa, d + o_state, d + o_goal, d + o_act, d + o_sig, sa.n_steps ? d + o_act : d + o_out, B, t,
           flags, lambda, st);
  if (rc) return rc;
  BESO_CUDA(cudaMemcpyAsync(p->h_pin + o_out, sa.n_steps ? d + o_act : d + o_out, n_act * sizeof(float),
                            cudaMemcpyDeviceToHost, st));
  BESO_CUDA(cudaStreamSynchronize(st));
  memcpy(out, p->h_pin + o_out, n_act * sizeof(float));
  return BESO_OK;
}

int beso_denoise_fwd_host(beso_plan* p, int mode, const float* state, const float* action, const float* goal,
                          const float* sigma, float* out, int B, int t, uint32_t flags, float lambda, void* stream) {
  int rc = check_call(p, mode, B, t, flags);
  if (rc) return rc;
  if (!state || !action || !sigma || !out) { set_error("null tensor pointer"); return BESO_E_INVALID; }
  SampleArgs sa;
  memset(&sa, 0, sizeof(sa));
  return host_call(p, mode, sa, state, action, goal, sigma, out, B, t, flags, lambda, (cudaStream_t)stream);
}

int beso_sample_loop_host(beso_plan* p, int mode, int sampler, const float* sigmas, int n_sigmas, const float* coef,
                          const float* state, const float* goal, float* x, int B, int t, uint32_t flags,
                          float lambda, void* stream) {
  int rc = check_call(p, mode, B, t, flags);
  if (rc) return rc;
  if (!state || !x) { set_error("null tensor pointer"); return BESO_E_INVALID; }
  SampleArgs sa;
  rc = make_sample_args(sampler, sigmas, n_sigmas, coef, &sa);
  if (rc) return rc;
  rc = check_sampler(p, mode, sampler, sa, false, flags);       // no noise argument here: noisy samplers are refused
  if (rc) return rc;
  return host_call(p, mode, sa, state, x, goal, nullptr, x, B, t, flags, lambda, (cudaStream_t)stream);
}

int beso_plan_set_params(beso_plan* p, int slot, const float* const* prm, int n_params) {
  if (!p || !prm || slot < 0 || slot > 1) { set_error("bad argument"); return BESO_E_INVALID; }
  if (n_params != (int)param_table(p->desc).size()) { set_error("wrong number of parameter tensors"); return BESO_E_INVALID; }
  p->slot[slot].params.assign(prm, prm + n_params);
  return BESO_OK;
}

int beso_loss_fwd_bwd(beso_plan* p, const float* state, const float* action, const float* goal, const float* noise,
                      const float* sigma, const float* goal_keep, float* loss_dev, float* flat_grad_dev, int B,
                      uint32_t flags, void* stream) {
  return beso_loss_fwd_bwd_dropout(p, state, action, goal, noise, sigma, goal_keep, nullptr, loss_dev, flat_grad_dev, B, flags, stream);
}

int beso_debug_gemm(beso_plan* p, const float* A, int lda, int a_kmajor, const float* B, int ldb, int b_kmajor, float* Cm,
                    int ldc, int M, int N, int K, const float* bias, int accumulate, int prec, void* stream) {
  if (!p) { set_error("null plan"); return BESO_E_INVALID; }
  BESO_CUDA(cudaSetDevice(p->device));
  GemmArgs a{};
  a.A = A; a.lda = lda; a.a_kmajor = a_kmajor; a.B = B; a.ldb = ldb; a.b_kmajor = b_kmajor; a.C = Cm; a.ldc = ldc;
  a.M = M; a.N = N; a.K = K; a.bias = bias; a.accumulate = accumulate; a.prec = prec;
  return train_gemm(p->train_ws, a, (cudaStream_t)stream);
}

int beso_loss_fwd_bwd_dropout(beso_plan* p, const float* state, const float* action, const float* goal, const float* noise,
                              const float* sigma, const float* goal_keep, const beso_dropout_masks* masks, float* loss_dev,
                              float* flat_grad_dev, int B, uint32_t flags, void* stream) {
  return beso_loss_fwd_bwd_dp(p, state, action, goal, noise, sigma, goal_keep, masks, nullptr, 1.0f, loss_dev, flat_grad_dev, B,
                              flags, stream);
}

int beso_loss_fwd_bwd_dp(beso_plan* p, const float* state, const float* action, const float* goal, const float* noise,
                         const float* sigma, const float* goal_keep, const beso_dropout_masks* masks, beso_comm* comm,
                         float grad_scale, float* loss_dev, float* flat_grad_dev, int B, uint32_t flags, void* stream) {
  if (!p || !state || !action || !noise || !sigma || !loss_dev || B < 1) { set_error("null argument or B < 1"); return BESO_E_INVALID; }
  if (!goal && p->desc.goal_conditioned && p->desc.goal_len > 0) { set_error("null goal"); return BESO_E_INVALID; }
  const WeightSlot& ws = p->slot[p->active];
  if (ws.params.empty()) { set_error("no parameters registered (beso_plan_set_params / beso_plan_pack_weights)"); return BESO_E_NOT_PACKED; }
  BESO_CUDA(cudaSetDevice(p->device));
  return train_loss_fwd_bwd(p->train_ws, p->desc, ws.params.data(), state, action, goal, noise, sigma, goal_keep, masks,
                            loss_dev, flat_grad_dev, B, flags, (cudaStream_t)stream, comm, grad_scale);
}

int64_t beso_kernel_launches(void) { return g_kernel_launches; }
int beso_debug_set_precise_layout(int layout) { fast_set_prec_layout(layout); return BESO_OK; }
int beso_debug_set_trace(float* trace_dev) { fast_set_trace(trace_dev); return BESO_OK; }
int beso_debug_set_timeline(long long* dev) { fast_set_timeline(dev); return BESO_OK; }
int beso_debug_mma_rate(long long* out_dev, const void* src_dev, int mode, void* stream) { return fast_mma_rate(out_dev, src_dev, mode, (cudaStream_t)stream); }

int beso_plan_rows_per_cta(beso_plan* p, int mode, int t) {
  if (!p) return BESO_E_INVALID;
  if (mode == BESO_MODE_FAST) return p->fast_ok ? fast_seqs_per_tile(p->desc, t, false) : BESO_E_UNSUPPORTED;
  if (mode == BESO_MODE_PRECISE && p->fast_ok && !p->force_simt) return fast_seqs_per_tile(p->desc, t, true);
  SimtLaunch L{};
  SimtModel tmp{};
  simt_layout(p->desc, nullptr, &tmp);
  int rc = simt_plan_launch(tmp, t, p->max_smem, &L);
  return rc ? rc : L.S;
}

}  // extern "C"
