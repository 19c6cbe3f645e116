// The exchange step of the parameter-gradient all-reduce over NVLink peer memory, as a device function: called by the
// stand-alone kernel (peer_allreduce.cu) and by the tail of the parameter-Jacobian kernel (synth.cu: the gradients are
// exchanged by the CTA that computed them — no second launch on the step's critical path).
//
// Every rank owns an exchange buffer in peer-mapped memory (allocated and rendezvoused by the host through torch's
// symmetric memory, which hands us one device pointer per rank).  Warp w pushes this rank's payload into rank w's
// buffer over NVLink as 8-byte words {value, sequence number} — one aligned 8-byte store each, so a word arrives whole and
// carries its own flag: no fence, no separate flag store, ONE NVLink traversal on the critical path (the scheme of NCCL's
// low-latency protocol; the first version stored the payload, fenced at system scope — a round trip — and then raised a
// flag: +8 us per step at any world size).  Then the warp polls the words of rank w's slot in the LOCAL buffer until they
// carry this call's number.  All ranks add the payloads in rank order 0..W-1: the result is bit-identical on every rank
// and deterministic.
//
// The sequence number lives in local device memory (incremented by the kernel itself, so a CUDA-graph replay needs no
// new arguments); slots alternate by its parity, so a rank that runs ahead writes step k+1 into the other half while a
// slow peer still reads step k (it cannot reach step k+2 before that peer has sent its step-k+1 words, i.e. has
// finished reading step k).
//
// A peer that does not answer: ranks of a training job skew by seconds to minutes (checkpointing, validation on rank
// 0, a data-loader stall), and NCCL would simply wait.  So does this: the wait is bounded by `timeout_ns` of wall clock
// (%globaltimer; default 10 minutes, 0 = wait for ever), and running into the bound is FATAL — *status is set and the
// kernel traps, so the next CUDA call of the process fails loudly.  A gradient is never replaced by a made-up value.
#pragma once
#include "common.cuh"

namespace sn {

constexpr int kPeerSlotFloats = 256;              // 1 KB slots: up to 128 words of {value, sequence number}
constexpr int kPeerMaxPayload = SN_MAX_PARAM_PTRS;
static_assert(2 * kPeerMaxPayload <= kPeerSlotFloats, "a slot holds the largest payload");
constexpr int kPeerMaxWorld = 16;

struct PeerArgs {
    float* buf[kPeerMaxWorld];  // buf[w]: rank w's exchange buffer [2][world][kPeerSlotFloats] (peer-mapped)
    int rank, world;
    unsigned* seq_counter;      // local device memory
    int* status;                // local device memory, may be NULL
    long long timeout_ns;       // 0 = no bound
};

__device__ __forceinline__ void st_word_sys(float* p, float v, unsigned seq) {  // one 8-byte store: value + its flag
    asm volatile("st.volatile.global.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(seq) : "memory");
}
__device__ __forceinline__ void ld_word_sys(const float* p, float& v, unsigned& seq) {
    unsigned a, b;
    asm volatile("ld.volatile.global.v2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "l"(p) : "memory");
    v = __uint_as_float(a);
    seq = b;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// data[0..n) <- sum over ranks, in place.  Call with ALL threads of a CTA of at least 32 * world threads; `data` must
// be visible to the CTA (written before a __syncthreads() by its own threads, or by an earlier kernel).
__device__ __forceinline__ void peer_exchange(const PeerArgs& a, float* __restrict__ data, int n) {
    __shared__ float s_data[kPeerMaxWorld][kPeerMaxPayload];
    __shared__ unsigned s_seq;
    __shared__ int s_bad;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = a.rank, world = a.world;
    if (threadIdx.x == 0) {
        s_seq = *a.seq_counter + 1u;
        *a.seq_counter = s_seq;
        s_bad = 0;
    }
    __syncthreads();
    const unsigned seq = s_seq;
    const int par = (int)(seq & 1u);
    if (w < world) {
        // push: my payload -> rank w's buffer, slot [par][rank]
        float* dst = a.buf[w] + (size_t)(par * world + rank) * kPeerSlotFloats;
        for (int i = lane; i < n; i += 32) st_word_sys(dst + 2 * i, data[i], seq);
        // pull: rank w's payload from MY buffer, slot [par][w]: every lane polls its own words
        const float* src = a.buf[rank] + (size_t)(par * world + w) * kPeerSlotFloats;
        const unsigned long long t0 = global_timer_ns();
        for (int i = lane; i < n; i += 32) {
            float v;
            unsigned got, polls = 0;
            for (;;) {
                ld_word_sys(src + 2 * i, v, got);
                if (got == seq) break;
                if ((++polls & 1023u) == 0u && a.timeout_ns > 0 && global_timer_ns() - t0 > (unsigned long long)a.timeout_ns) {
                    s_bad = 1;
                    v = 0.f;
                    break;
                }
                if (polls > 64u) __nanosleep(100);
            }
            s_data[w][i] = v;
        }
    }
    __syncthreads();
    if (s_bad) {
        // a peer never answered within the bound: fatal, like a collective watchdog — never continue with a made-up gradient
        if (threadIdx.x == 0) {
            if (a.status) *a.status = 1;
            __threadfence_system();
            __trap();
        }
        return;
    }
    if (threadIdx.x < n) {
        float acc = 0.f;
        for (int r = 0; r < world; ++r) acc += s_data[r][threadIdx.x];  // rank order: identical on every rank
        data[threadIdx.x] = acc;
    }
}

inline int fill_peer_args(PeerArgs& a, int rank, int world, const uint64_t* peer_bufs_host, uint32_t* seq_counter, int32_t* status,
                          int64_t timeout_ms) {
    if (!peer_bufs_host || !seq_counter) return SN_ERR_BAD_ARG;
    if (world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world || timeout_ms < 0) return SN_ERR_BAD_ARG;
    for (int w = 0; w < kPeerMaxWorld; ++w) {
        a.buf[w] = w < world ? reinterpret_cast<float*>(peer_bufs_host[w]) : nullptr;
        if (w < world && (!a.buf[w] || (peer_bufs_host[w] & 15))) return SN_ERR_BAD_ARG;
    }
    a.rank = rank; a.world = world; a.seq_counter = seq_counter; a.status = status;
    a.timeout_ns = (long long)timeout_ms * 1000000LL;
    return SN_OK;
}

}  // namespace sn
