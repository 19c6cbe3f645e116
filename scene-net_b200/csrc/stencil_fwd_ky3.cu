// forward stencil, kernels with ky = 3 (see stencil_fwd_impl.cuh)
#include "stencil_fwd_impl.cuh"
namespace sn {
int stencil_fwd_ky3(const FwdParams& p, cudaStream_t s) { return stencil_fwd_ky<3>(p, s); }
}  // namespace sn
