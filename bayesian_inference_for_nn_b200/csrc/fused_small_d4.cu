// kernels of the small-width path for 4 input feature(s) (see fused_small.cuh)
#include "fused_small.cuh"

namespace pyb {
FS_DEFINE_LAUNCH_D(4)
}  // namespace pyb
