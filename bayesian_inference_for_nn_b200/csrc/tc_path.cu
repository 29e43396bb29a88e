// tc_path.cu — placeholder until the tcgen05 path lands.
#include "common.cuh"
namespace pyb {
bool tc_supported(pyb_handle*, int64_t) { return false; }
void tc_eval(pyb_handle*, const float*, int64_t, float, float*, float*) {
  throw Error(PYB_ERR_UNSUPPORTED, "tensor path not built");
}
void tc_release(pyb_handle*) {}
void tc_invalidate_dataset(pyb_handle*) {}
}  // namespace pyb
