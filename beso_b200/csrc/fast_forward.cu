#include "fast.cuh"
namespace beso {
bool fast_supported(const beso_model_desc&) { return false; }
int fast_seqs_per_tile(const beso_model_desc&, int) { return 0; }
int fast_pack(FastWeights&, const beso_model_desc&, const float* const*, cudaStream_t) { return BESO_OK; }
void fast_free(FastWeights&) {}
int fast_launch(const FastWeights&, const beso_model_desc&, int, const SampleArgs&, const float*, const float*,
                const float*, const float*, float*, int, int, uint32_t, float, cudaStream_t) {
  set_error("fast mode not built"); return BESO_E_UNSUPPORTED;
}
}  // namespace beso
