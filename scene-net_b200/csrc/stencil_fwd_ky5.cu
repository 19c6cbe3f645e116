// forward stencil, kernels with ky = 5 (see stencil_fwd_impl.cuh)
#include "stencil_fwd_impl.cuh"
namespace sn {
int stencil_fwd_ky5(const FwdParams& p, cudaStream_t s) { return stencil_fwd_ky<5>(p, s); }
}  // namespace sn
