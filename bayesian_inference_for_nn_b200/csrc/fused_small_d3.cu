// kernels of the small-width path for 3 input feature(s) (see fused_small.cuh)
#include "fused_small.cuh"

namespace pyb {
FS_DEFINE_LAUNCH_D(3)
}  // namespace pyb
