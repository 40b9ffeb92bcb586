// Fused AdamW + EMA step over all parameter tensors of the model in ONE launch.
//
// Replaces, for the training step of beso/agents/diffusion_agents/beso_agent.py:238-247:
//   self.optimizer.step()                      torch.optim.AdamW (configs/agents/beso_kitchen.yaml:9-12)
//   self.ema_helper.update(model.parameters()) beso/networks/ema_helper/ema.py:36-53
// The reference runs ~8 element-wise ATen launches per parameter tensor for AdamW (75 tensors at K256) plus
// a Python loop of one launch per tensor for the EMA; here every element is read and written once.
// Arithmetic follows torch's single-tensor AdamW (torch/optim/adamw.py _single_tensor_adamw) in fp32:
//   p *= 1 - lr * wd;  m += (g - m) * (1 - b1);  v = v * b2 + (1 - b2) * g * g
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps);   ema -= (1 - decay) * (ema - p)
// HBM-bound: 7 reads/writes of 4 bytes per element without EMA (p, g, m, v in; p, m, v out), 9 with.
#include <cuda_runtime.h>
#include <math.h>

#include <vector>

#include "../../include/beso_b200.h"
#include "common.cuh"

namespace beso {

struct OptChunk { float* p; long long flat_off; int n; int pad; };   // <= kChunk elements of one tensor

}  // namespace beso

struct beso_opt {
  int device = 0;
  long long total = 0;
  int n_chunks = 0;
  beso::OptChunk* chunks_dev = nullptr;
};

namespace beso {
namespace {

constexpr int kChunk = 4096, kOptThreads = 256;

struct OptArgs {
  const float* grad; float* m; float* v; float* ema;
  float decay_mul, one_minus_b1, b2, one_minus_b2, step_size, inv_bc2_sqrt, eps, ema_omd, grad_scale;
};

__device__ __forceinline__ void adamw_elem(float& p, float g, float& m, float& v, float* ema, const OptArgs& a) {
  g *= a.grad_scale;
  p = __fmul_rn(p, a.decay_mul);
  m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), a.one_minus_b1));
  v = __fadd_rn(__fmul_rn(v, a.b2), __fmul_rn(__fmul_rn(a.one_minus_b2, g), g));
  const float denom = __fadd_rn(__fmul_rn(__fsqrt_rn(v), a.inv_bc2_sqrt), a.eps);
  p = __fsub_rn(p, __fmul_rn(a.step_size, __fdiv_rn(m, denom)));
  if (ema != nullptr) *ema = __fsub_rn(*ema, __fmul_rn(a.ema_omd, __fsub_rn(*ema, p)));
}

__global__ void __launch_bounds__(kOptThreads) adamw_ema_kernel(const OptChunk* chunks, OptArgs a) {
  const OptChunk c = chunks[blockIdx.x];
  const long long o = c.flat_off;
  // tensors start at arbitrary element offsets of the flat buffers: vectorise only when everything is 16-byte aligned
  const bool vec = ((reinterpret_cast<uintptr_t>(c.p) | (uintptr_t)(o * 4)) & 15) == 0 && (c.n & 3) == 0;
  if (vec) {
    for (int i = threadIdx.x * 4; i < c.n; i += kOptThreads * 4) {
      float4 p = *reinterpret_cast<float4*>(c.p + i);
      const float4 g = *reinterpret_cast<const float4*>(a.grad + o + i);
      float4 m = *reinterpret_cast<float4*>(a.m + o + i), v = *reinterpret_cast<float4*>(a.v + o + i);
      float4 e = a.ema ? *reinterpret_cast<float4*>(a.ema + o + i) : make_float4(0, 0, 0, 0);
      adamw_elem(p.x, g.x, m.x, v.x, a.ema ? &e.x : nullptr, a);
      adamw_elem(p.y, g.y, m.y, v.y, a.ema ? &e.y : nullptr, a);
      adamw_elem(p.z, g.z, m.z, v.z, a.ema ? &e.z : nullptr, a);
      adamw_elem(p.w, g.w, m.w, v.w, a.ema ? &e.w : nullptr, a);
      *reinterpret_cast<float4*>(c.p + i) = p;
      *reinterpret_cast<float4*>(a.m + o + i) = m;
      *reinterpret_cast<float4*>(a.v + o + i) = v;
      if (a.ema) *reinterpret_cast<float4*>(a.ema + o + i) = e;
    }
  } else {
    for (int i = threadIdx.x; i < c.n; i += kOptThreads) {
      float p = c.p[i], m = a.m[o + i], v = a.v[o + i];
      float e = a.ema ? a.ema[o + i] : 0.f;
      adamw_elem(p, a.grad[o + i], m, v, a.ema ? &e : nullptr, a);
      c.p[i] = p; a.m[o + i] = m; a.v[o + i] = v;
      if (a.ema) a.ema[o + i] = e;
    }
  }
}

}  // namespace
}  // namespace beso

using namespace beso;

extern "C" {

int beso_opt_create(int device, int n_tensors, float* const* param_dev_ptrs, const long long* numel, beso_opt** out) {
  if (!out || n_tensors < 1 || !param_dev_ptrs || !numel) { set_error("beso_opt_create: null argument"); return BESO_E_INVALID; }
  BESO_CUDA(cudaSetDevice(device));
  std::vector<OptChunk> chunks;
  long long off = 0;
  for (int t = 0; t < n_tensors; ++t) {
    if (!param_dev_ptrs[t] || numel[t] < 0) { set_error("beso_opt_create: bad tensor"); return BESO_E_INVALID; }
    for (long long i = 0; i < numel[t]; i += kChunk)
      chunks.push_back({param_dev_ptrs[t] + i, off + i, (int)((numel[t] - i) < kChunk ? (numel[t] - i) : kChunk), 0});
    off += numel[t];
  }
  beso_opt* o = new beso_opt();
  o->device = device; o->total = off; o->n_chunks = (int)chunks.size();
  BESO_CUDA(cudaMalloc(&o->chunks_dev, chunks.size() * sizeof(OptChunk)));
  BESO_CUDA(cudaMemcpy(o->chunks_dev, chunks.data(), chunks.size() * sizeof(OptChunk), cudaMemcpyHostToDevice));
  *out = o;
  return BESO_OK;
}

int beso_opt_destroy(beso_opt* o) {
  if (!o) return BESO_OK;
  if (o->chunks_dev) cudaFree(o->chunks_dev);
  delete o;
  return BESO_OK;
}

long long beso_opt_total(const beso_opt* o) { return o ? o->total : 0; }

int beso_opt_step(beso_opt* o, const float* flat_grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev, float* ema_dev,
                  float lr, float beta1, float beta2, float eps, float weight_decay, int step, float ema_decay,
                  float grad_scale, void* stream) {
  if (!o || !flat_grad_dev || !exp_avg_dev || !exp_avg_sq_dev || step < 1) { set_error("beso_opt_step: null argument or step < 1"); return BESO_E_INVALID; }
  BESO_CUDA(cudaSetDevice(o->device));
  OptArgs a{};
  a.grad = flat_grad_dev; a.m = exp_avg_dev; a.v = exp_avg_sq_dev; a.ema = ema_dev;
  // scalars exactly as torch computes them (python floats = doubles, then one rounding to fp32 in the kernel call)
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  a.decay_mul = (float)(1.0 - (double)lr * (double)weight_decay);
  a.one_minus_b1 = (float)(1.0 - (double)beta1);
  a.b2 = beta2;
  a.one_minus_b2 = (float)(1.0 - (double)beta2);
  a.step_size = (float)((double)lr / bc1);
  a.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  a.eps = eps;
  a.ema_omd = (float)(1.0 - (double)ema_decay);
  a.grad_scale = grad_scale;
  adamw_ema_kernel<<<o->n_chunks, kOptThreads, 0, (cudaStream_t)stream>>>(o->chunks_dev, a);
  ++g_kernel_launches;
  BESO_CUDA(cudaGetLastError());
  return BESO_OK;
}

}  // extern "C"
