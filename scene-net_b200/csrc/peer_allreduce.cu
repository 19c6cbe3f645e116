// All-reduce of the parameter-gradient payload (<= 96 floats) over NVLink peer memory — one tiny kernel instead
// of an NCCL collective.
//
// The multi-GPU path exchanges 11-96 floats per step: pure latency.  ncclAllReduce costs ~19 us
// per step at 2 GPUs next to a 180 us step (profiles/r1_notes.md).  The exchange itself lives in peer_exchange.cuh;
// this file is its stand-alone launch (any float32 vector); the training step uses the same exchange fused into the
// tail of the parameter-Jacobian kernel (sn_scenenet_param_grads_allreduce, synth.cu).
#include "peer_exchange.cuh"

namespace sn {

__global__ void __launch_bounds__(32 * kPeerMaxWorld)
peer_allreduce_kernel(const __grid_constant__ PeerArgs a, float* __restrict__ data, int n) {
    peer_exchange(a, data, n);
}

}  // namespace sn

extern "C" int64_t sn_peer_allreduce_buffer_bytes(int world) {
    if (world < 1 || world > sn::kPeerMaxWorld) return SN_ERR_BAD_ARG;
    return (int64_t)2 * world * sn::kPeerSlotFloats * 4;
}

extern "C" int sn_peer_allreduce(float* data, int n, int rank, int world, const uint64_t* peer_bufs_host,
                                 uint32_t* seq_counter, int32_t* status, int64_t timeout_ms, void* stream) {
    if (!data) return SN_ERR_BAD_ARG;
    if (n < 1 || n > SN_MAX_PARAM_PTRS) return SN_ERR_BAD_ARG;
    sn::PeerArgs a;
    const int rc = sn::fill_peer_args(a, rank, world, peer_bufs_host, seq_counter, status, timeout_ms);
    if (rc) return rc;
    sn::peer_allreduce_kernel<<<1, 32 * sn::kPeerMaxWorld, 0, (cudaStream_t)stream>>>(a, data, n);
    SN_LAUNCH_CHECK();
    return SN_OK;
}
