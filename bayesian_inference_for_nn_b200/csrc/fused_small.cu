// fused_small.cu — placeholder until the fused small-width trajectory kernel lands.
#include "common.cuh"
namespace pyb {
bool fused_small_supported(pyb_handle*) { return false; }
void fused_small_eval(pyb_handle*, const float*, int64_t, float, float*, float*) {
  throw Error(PYB_ERR_UNSUPPORTED, "fused small path not built");
}
}  // namespace pyb
