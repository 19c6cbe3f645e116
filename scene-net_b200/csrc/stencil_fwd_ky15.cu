// forward stencil, kernels with ky = 15 (see stencil_fwd_impl.cuh)
#include "stencil_fwd_impl.cuh"
namespace sn {
int stencil_fwd_ky15(const FwdParams& p, cudaStream_t s) { return stencil_fwd_ky<15>(p, s); }
}  // namespace sn
