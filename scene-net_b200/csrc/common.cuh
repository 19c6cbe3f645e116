// Shared helpers for the scenenet_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <atomic>
#include "../../include/scenenet_b200.h"

namespace sn {

// bumped by every launch wrapper (sn_launch_count()); the only mutable global of the library — a relaxed atomic, so the
// entry points stay callable from any host thread
extern std::atomic<long long> g_launch_count;

// Measurement / debugging switches read from the environment exist only in -DSN_DEBUG builds: a release library never
// looks at the environment, so no variable can change which kernel runs or what it computes.
#ifdef SN_DEBUG
#define SN_ENV(name) getenv(name)
#else
#define SN_ENV(name) (static_cast<const char*>(nullptr))
#endif

inline int cuda_rc(cudaError_t e) { return e == cudaSuccess ? SN_OK : (SN_ERR_CUDA_BASE - (int)e); }

#define SN_LAUNCH_CHECK()                                   \
    do {                                                    \
        ::sn::g_launch_count.fetch_add(1, std::memory_order_relaxed); \
        cudaError_t _e = cudaGetLastError();                \
        if (_e != cudaSuccess) return ::sn::cuda_rc(_e);    \
    } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// Layout of the grid state buffer behind the SN_STATE_WORDS counters (sn_grid_state_bytes): mask words, then tile ids.
__host__ __device__ inline long long state_mask_words(long long n) { return (n + 31) / 32 + 4; }
// capacity of the dense-tile list: every shape whose forward tiles hold >= 1024 voxels on average fits (a 64^3 grid has
// 64 tiles of 4096 voxels); smaller / ragged shapes with more tiles than this keep every tile in the occupancy kernel
__host__ __device__ inline long long state_tile_cap(long long n) { return n / 1024 + 1024; }

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div64(long long a, long long b) { return (a + b - 1) / b; }

// PyTorch 'same' padding: left = (k-1)/2, right = k-1-left (asymmetric for even k)
__host__ __device__ inline int pad_left(int k) { return (k - 1) / 2; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------- mbarrier + TMA (PTX)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 4-D tiled TMA load global -> shared (coordinates innermost first; OOB elements are zero-filled,
// which implements the 'same' zero padding of the convolution for free)
__device__ __forceinline__ void tma_load_4d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

}  // namespace sn
