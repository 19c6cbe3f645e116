// forward stencil, kernels with ky = 9 (see stencil_fwd_impl.cuh)
#include "stencil_fwd_impl.cuh"
namespace sn {
int stencil_fwd_ky9(const FwdParams& p, cudaStream_t s) { return stencil_fwd_ky<9>(p, s); }
}  // namespace sn
