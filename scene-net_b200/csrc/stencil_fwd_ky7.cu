// forward stencil, kernels with ky = 7 (see stencil_fwd_impl.cuh)
#include "stencil_fwd_impl.cuh"
namespace sn {
int stencil_fwd_ky7(const FwdParams& p, cudaStream_t s) { return stencil_fwd_ky<7>(p, s); }
}  // namespace sn
