// rvq_fused_kernel<SL = 3, ...> (D = 384): its own translation unit so that the four D build in parallel.
#include "vq_rvq_fused.cuh"

namespace vqb {
RQ_DEFINE_LAUNCH_SL(3)
}  // namespace vqb
