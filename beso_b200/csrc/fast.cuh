// FAST mode interface: fp16 tcgen05 fused score-GPT (fast_forward.cu).
#pragma once
#include "common.cuh"

namespace beso {

struct FastWeights {
  void* tape = nullptr;      // bf16 B-operand blocks in consumption order, pre-swizzled for UMMA
  float* vec = nullptr;      // fp32 vectors: biases, LayerNorm affine, embeddings tables
  size_t tape_bytes = 0, vec_floats = 0;
};

bool fast_supported(const beso_model_desc& m);
int fast_seqs_per_tile(const beso_model_desc& m, int t, bool prec);
// prec = false: fp16 operands (FAST).  prec = true: [hi | lo] fp16 images of every weight tile for the split-operand
// precise mode (fp32-equivalent products on the tensor pipe).
int fast_pack(FastWeights& w, const beso_model_desc& m, const float* const* params, cudaStream_t st, bool prec);
void fast_free(FastWeights& w);
void fast_set_trace(float* trace_dev);
void fast_set_timeline(long long* dev);
int fast_mma_rate(long long* out_dev, const void* src_dev, int mode, cudaStream_t st);
int fast_launch(const FastWeights& w, const beso_model_desc& m, int sm_count, const SampleArgs& sa,
                const float* state, const float* goal, const float* action_or_x, const float* sigma,
                float* out, int B, int t, uint32_t flags, float cond_lambda, cudaStream_t st, bool prec);

}  // namespace beso
