// The plan object behind the opaque `beso_plan*` of the C ABI (shared by api.cu and train.cu).
#pragma once
#include <vector>

#include "common.cuh"
#include "fast.cuh"

namespace beso {

struct TrainWs;
void train_ws_free(TrainWs* ws);
int train_loss_fwd_bwd(TrainWs*& ws, const beso_model_desc& m, const float* const* prm, const float* state,
                       const float* action, const float* goal, const float* noise, const float* sigma,
                       const float* goal_keep, const beso_dropout_masks* drop, float* loss_out, float* grad, int B,
                       uint32_t flags, cudaStream_t st, beso_comm* comm = nullptr, float grad_scale = 1.0f);
struct GemmArgs;
int train_gemm(TrainWs*& ws, const GemmArgs& a, cudaStream_t st);   // the training GEMM by itself (tests, tools)

struct WeightSlot {
  float* simt_buf = nullptr;     // transposed fp32 images (PRECISE)
  SimtModel simt{};
  FastWeights fast{};            // fp16 UMMA tape + fp32 / fp16 vectors (FAST)
  FastWeights fastp{};           // [hi | lo] fp16 tape + fp32 vectors (PRECISE on the tensor pipe, stacked 64-row tiles)
  FastWeights fastq{};           // the same for the 128-row tile layout (P128; embed_dim <= 256)
  std::vector<const float*> params;   // raw fp32 parameter tensors, parameters() order (training path)
  bool packed = false;
};

}  // namespace beso

struct beso_plan {
  beso_model_desc desc{};
  int device = 0;
  int sm_count = 0;
  int max_smem = 0;
  int active = 0;
  beso::WeightSlot slot[2];
  size_t simt_floats = 0;
  bool fast_ok = false;
  bool force_simt = false;       // BESO_PRECISE_SIMT=1: PRECISE always runs the CUDA-core kernel
  // staging for the *_host entry points
  float *h_pin = nullptr, *d_stage = nullptr;
  size_t stage_floats = 0;
  beso::TrainWs* train_ws = nullptr;
};
