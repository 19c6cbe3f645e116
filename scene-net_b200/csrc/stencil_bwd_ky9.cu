// tap-gradient stencil, kernels with ky = 9 (see stencil_bwd_impl.cuh)
#include "stencil_bwd_impl.cuh"
namespace sn {
int stencil_bwd_ky9(const BwdParams& p, void* ws, int64_t wsb, int* rows, int* TP, cudaStream_t s) { return stencil_bwd_ky<9>(p, ws, wsb, rows, TP, s); }
int64_t stencil_bwd_ws_ky9(int B, int Z, int X, int Y, int kz, int kx) { return stencil_bwd_ws_ky<9>(B, Z, X, Y, kz, kx); }
}  // namespace sn
