// All-reduce of the parameter-gradient payload (<= 96 floats) over NVLink peer memory — one tiny kernel instead
// of an NCCL collective.
//
// The multi-GPU path exchanges 11-96 floats per step: pure latency.  ncclAllReduce costs ~19 us
// per step at 2 GPUs next to a 180 us step (profiles/r1_notes.md).  Here every rank owns an exchange buffer in
// peer-mapped memory (allocated and rendezvoused by the host through torch's symmetric memory, which hands us
// one device pointer per rank).  Warp w of the single CTA pushes this rank's payload into rank w's buffer with
// plain stores over NVLink, fences, and raises a flag there; then it waits for rank w's flag in the LOCAL buffer
// and stages rank w's payload.  All ranks add the staged payloads in rank order 0..W-1: the result is
// bit-identical on every rank and deterministic.
//
// Flags carry a sequence number kept in local device memory (incremented by the kernel itself, so a CUDA-graph
// replay needs no new arguments); slots alternate by its parity, so a rank that runs ahead writes step k+1 into
// the other half while a slow peer still reads step k (it cannot reach step k+2 before that peer has sent its
// step-k+1 flag, i.e. has finished reading step k).  The wait is bounded: on timeout the payload is poisoned
// with NaN and *status is set instead of hanging the device.
#include "common.cuh"

namespace sn {

constexpr int kPeerSlotFloats = 128;              // 96 payload floats + flag, 512-byte slots
constexpr int kPeerFlagIdx = kPeerSlotFloats - 1;
constexpr int kPeerMaxWorld = 16;

struct PeerArgs {
    float* buf[kPeerMaxWorld];  // buf[w]: rank w's exchange buffer [2][world][kPeerSlotFloats] (peer-mapped)
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(32 * kPeerMaxWorld)
peer_allreduce_kernel(const __grid_constant__ PeerArgs a, float* __restrict__ data, int n, int rank, int world,
                      unsigned* __restrict__ seq_counter, int* __restrict__ status, long long max_polls) {
    __shared__ float s_data[kPeerMaxWorld][kPeerSlotFloats];
    __shared__ unsigned s_seq;
    __shared__ int s_bad;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        s_seq = *seq_counter + 1u;
        *seq_counter = s_seq;
        s_bad = 0;
    }
    __syncthreads();
    const unsigned seq = s_seq;
    const int par = (int)(seq & 1u);
    if (w < world) {
        // push: my payload -> rank w's buffer, slot [par][rank]
        float* dst = a.buf[w] + (size_t)(par * world + rank) * kPeerSlotFloats;
        for (int i = lane; i < n; i += 32) dst[i] = data[i];
        __threadfence_system();
        __syncwarp();
        if (lane == 0) st_release_sys(reinterpret_cast<unsigned*>(dst + kPeerFlagIdx), seq);
        // pull: wait for rank w's payload in MY buffer, slot [par][w]
        const float* src = a.buf[rank] + (size_t)(par * world + w) * kPeerSlotFloats;
        if (lane == 0) {
            long long polls = 0;
            while (ld_acquire_sys(reinterpret_cast<const unsigned*>(src + kPeerFlagIdx)) != seq) {
                if (++polls > max_polls) {
                    s_bad = 1;
                    break;
                }
                __nanosleep(20);
            }
        }
        __syncwarp();
        for (int i = lane; i < n; i += 32) s_data[w][i] = __ldcg(src + i);  // L2 (where the peer's stores land), not L1
    }
    __syncthreads();
    if (threadIdx.x < n) {
        float acc = 0.f;
        for (int r = 0; r < world; ++r) acc += s_data[r][threadIdx.x];  // rank order: identical on every rank
        data[threadIdx.x] = s_bad ? __int_as_float(0x7fc00000) : acc;
    }
    if (threadIdx.x == 0 && s_bad && status) *status = 1;
}

}  // namespace sn

extern "C" int64_t sn_peer_allreduce_buffer_bytes(int world) {
    if (world < 1 || world > sn::kPeerMaxWorld) return SN_ERR_BAD_ARG;
    return (int64_t)2 * world * sn::kPeerSlotFloats * 4;
}

extern "C" int sn_peer_allreduce(float* data, int n, int rank, int world, const uint64_t* peer_bufs_host,
                                 uint32_t* seq_counter, int32_t* status, void* stream) {
    if (!data || !peer_bufs_host || !seq_counter) return SN_ERR_BAD_ARG;
    if (world < 1 || world > sn::kPeerMaxWorld || rank < 0 || rank >= world) return SN_ERR_BAD_ARG;
    if (n < 1 || n > SN_MAX_PARAM_PTRS) return SN_ERR_BAD_ARG;
    sn::PeerArgs a;
    for (int w = 0; w < sn::kPeerMaxWorld; ++w) {
        a.buf[w] = w < world ? reinterpret_cast<float*>(peer_bufs_host[w]) : nullptr;
        if (w < world && (!a.buf[w] || (peer_bufs_host[w] & 15))) return SN_ERR_BAD_ARG;
    }
    // a system-scope poll takes ~1 us: give up after ~2-3 s
    sn::peer_allreduce_kernel<<<1, 32 * sn::kPeerMaxWorld, 0, (cudaStream_t)stream>>>(a, data, n, rank, world, seq_counter, status,
                                                                                    2000000LL);
    SN_LAUNCH_CHECK();
    return SN_OK;
}
