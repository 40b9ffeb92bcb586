// FAST mode interface: fp16 tcgen05 fused score-GPT (fast_forward.cu).
#pragma once
#include "common.cuh"

namespace beso {

struct FastWeights {
  void* tape = nullptr;      // bf16 B-operand blocks in consumption order, pre-swizzled for UMMA
  float* vec = nullptr;      // fp32 vectors: biases, LayerNorm affine, embeddings tables
  size_t tape_bytes = 0, vec_floats = 0;
};

// Weight tape layouts.  F16: fp16 operands (FAST mode).  The precise mode (fp32-equivalent products on the tensor pipe
// from [hi | lo] fp16 images of every weight tile) has two tile layouts with their own tapes: STACKED = 64 sequence rows
// per tile, both images of a row on the MMA row dimension (any supported shape); P128 = full 128-row tiles, three MMAs
// per product, single-accumulator schedule (embed_dim <= 256).  fast_launch takes whichever is faster for the batch.
enum { FAST_LAYOUT_F16 = 0, FAST_LAYOUT_STACKED = 1, FAST_LAYOUT_P128 = 2 };
bool fast_supported(const beso_model_desc& m);
bool fast_p128_supported(const beso_model_desc& m);
int fast_seqs_per_tile(const beso_model_desc& m, int t, bool prec);
int fast_pack(FastWeights& w, const beso_model_desc& m, const float* const* params, cudaStream_t st, int layout);
void fast_free(FastWeights& w);
void fast_set_prec_layout(int layout);     // 0 = chosen per launch, FAST_LAYOUT_STACKED / FAST_LAYOUT_P128 = forced (tests, tools)
void fast_set_trace(float* trace_dev);
void fast_set_timeline(long long* dev);
int fast_mma_rate(long long* out_dev, const void* src_dev, int mode, cudaStream_t st);
// prec: w holds the STACKED tape and w_p128 (optional) the P128 tape
int fast_launch(const FastWeights& w, const beso_model_desc& m, int sm_count, const SampleArgs& sa,
                const float* state, const float* goal, const float* action_or_x, const float* sigma,
                float* out, int B, int t, uint32_t flags, float cond_lambda, cudaStream_t st, bool prec,
                const FastWeights* w_p128 = nullptr);

}  // namespace beso
