// Training path (loss forward + backward) and the data-parallel gradient exchange.
#include "common.cuh"
extern "C" {
int beso_loss_fwd_bwd(beso_plan*, const float*, const float*, const float*, const float*, const float*,
                      const float*, float*, float*, int, uint32_t, void*) {
  beso::set_error("beso_loss_fwd_bwd: not implemented in this build"); return BESO_E_UNSUPPORTED;
}
int beso_comm_unique_id(char*) { beso::set_error("comm: not implemented in this build"); return BESO_E_UNSUPPORTED; }
int beso_comm_init(int, int, const char*, int, beso_comm**) { beso::set_error("comm: not implemented in this build"); return BESO_E_UNSUPPORTED; }
int beso_comm_destroy(beso_comm*) { return BESO_OK; }
int beso_allreduce_grads(beso_comm*, float*, size_t, float, void*) { beso::set_error("comm: not implemented in this build"); return BESO_E_UNSUPPORTED; }
}
