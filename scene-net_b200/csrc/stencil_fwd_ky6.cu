// forward stencil, kernels with ky = 6 (see stencil_fwd_impl.cuh)
#include "stencil_fwd_impl.cuh"
namespace sn {
int stencil_fwd_ky6(const FwdParams& p, cudaStream_t s) { return stencil_fwd_ky<6>(p, s); }
}  // namespace sn
