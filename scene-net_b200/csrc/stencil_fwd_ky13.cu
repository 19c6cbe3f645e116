// forward stencil, kernels with ky = 13 (see stencil_fwd_impl.cuh)
#include "stencil_fwd_impl.cuh"
namespace sn {
int stencil_fwd_ky13(const FwdParams& p, cudaStream_t s) { return stencil_fwd_ky<13>(p, s); }
}  // namespace sn
