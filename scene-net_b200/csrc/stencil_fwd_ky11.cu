// forward stencil, kernels with ky = 11 (see stencil_fwd_impl.cuh)
#include "stencil_fwd_impl.cuh"
namespace sn {
int stencil_fwd_ky11(const FwdParams& p, cudaStream_t s) { return stencil_fwd_ky<11>(p, s); }
}  // namespace sn
