// vq_fwd_tc.cu -- tcgen05 filter + exact refine forward path (placeholder until the
// kernel lands: reports "unsupported" so that VQB_PATH_AUTO takes the FMA path).
#include "vq_common.cuh"

namespace vqb {

bool tc_shape_supported(int, int) { return false; }

cudaError_t launch_fwd_tc(const FwdParams &, float *, int, int, int *, int *, cudaStream_t, cudaEvent_t, cudaEvent_t)
{
    return cudaErrorNotSupported;
}

}  // namespace vqb
