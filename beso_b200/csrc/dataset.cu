// Windowed batch gather from GPU-resident trajectories: replaces the per-sample Python slicing of
// TrajectorySlicerDataset.__getitem__ (beso/envs/dataloaders/trajectory_loader.py:160-197) plus the DataLoader's
// collation and host-to-device copy with ONE launch: sample b reads
//   observation = obs[traj[b], start[b] : start[b] + W],  action = act[traj[b], start[b] : start[b] + W],
//   goal_observation = obs[traj[b], goal_start[b] : goal_start[b] + G]   (zeros when goal_start[b] < 0: the
//   reference's "zeros placeholder", trajectory_loader.py:181-184).
// Optionally the Scaler's scale_input / scale_output (beso/networks/scaler/scaler_class.py:79-112, 271-301) is applied
// on the way: a per-feature table (sub, div, mul, add) gives ((x - sub) / div) * mul + add in round-to-nearest fp32
// steps, no FMA contraction, i.e. the exact sequence of ATen elementwise ops the reference runs.  The zeros placeholder
// is scaled too, as the reference scales it in train_step (beso_agent.py:226-229).
// HBM-bound: every output element is one read and one write.
#include <cuda_runtime.h>

#include "../../include/beso_b200.h"
#include "common.cuh"

namespace beso {
namespace {

__device__ __forceinline__ float scale_feature(float x, const float* __restrict__ tab, int dim, int f) {
  if (tab == nullptr) return x;
  const float q = __fdiv_rn(__fsub_rn(x, tab[f]), tab[dim + f]);
  return __fadd_rn(__fmul_rn(q, tab[2 * dim + f]), tab[3 * dim + f]);
}

__global__ void __launch_bounds__(128) window_gather_kernel(const float* __restrict__ obs, const float* __restrict__ act, int t_max,
                                                            int obs_dim, int act_dim, const int* __restrict__ traj,
                                                            const int* __restrict__ start, const int* __restrict__ goal_start,
                                                            int W, int G, float* __restrict__ state_out,
                                                            float* __restrict__ action_out, float* __restrict__ goal_out,
                                                            const float* __restrict__ obs_tab, const float* __restrict__ act_tab) {
  const int b = blockIdx.x;
  const size_t tr = (size_t)traj[b];
  const int s = start[b];
  const float* so = obs + (tr * t_max + s) * obs_dim;
  const float* sa = act + (tr * t_max + s) * act_dim;
  for (int i = threadIdx.x; i < W * obs_dim; i += blockDim.x) state_out[(size_t)b * W * obs_dim + i] = scale_feature(so[i], obs_tab, obs_dim, i % obs_dim);
  for (int i = threadIdx.x; i < W * act_dim; i += blockDim.x) action_out[(size_t)b * W * act_dim + i] = scale_feature(sa[i], act_tab, act_dim, i % act_dim);
  if (goal_out != nullptr) {
    const int gs = goal_start[b];
    const float* sg = obs + (tr * t_max + (gs < 0 ? 0 : gs)) * obs_dim;
    for (int i = threadIdx.x; i < G * obs_dim; i += blockDim.x) goal_out[(size_t)b * G * obs_dim + i] = scale_feature(gs < 0 ? 0.f : sg[i], obs_tab, obs_dim, i % obs_dim);
  }
}

}  // namespace
}  // namespace beso

using namespace beso;

extern "C" int beso_window_gather(const float* obs_dev, const float* act_dev, int n_traj, int t_max, int obs_dim, int act_dim,
                                  const int* traj_dev, const int* start_dev, const int* goal_start_dev, int window, int goal_len,
                                  float* state_out_dev, float* action_out_dev, float* goal_out_dev,
                                  const float* obs_scale_dev, const float* act_scale_dev, int B, void* stream) {
  if (!obs_dev || !act_dev || !traj_dev || !start_dev || !state_out_dev || !action_out_dev || B < 1 || window < 1 || n_traj < 1 ||
      t_max < window || obs_dim < 1 || act_dim < 1) {
    set_error("beso_window_gather: null pointer or bad size"); return BESO_E_INVALID;
  }
  if (goal_out_dev && (!goal_start_dev || goal_len < 1 || goal_len > t_max)) { set_error("beso_window_gather: goal output needs goal_start and 1 <= goal_len <= t_max"); return BESO_E_INVALID; }
  window_gather_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(obs_dev, act_dev, t_max, obs_dim, act_dim, traj_dev, start_dev,
                                                           goal_start_dev, window, goal_len, state_out_dev, action_out_dev, goal_out_dev,
                                                           obs_scale_dev, act_scale_dev);
  ++g_kernel_launches;
  BESO_CUDA(cudaGetLastError());
  return BESO_OK;
}
