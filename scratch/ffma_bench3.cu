// Stencil main loop in isolation, variants approaching the real kernel.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int RZ = 8, KY = 5, C = 9, KX = 5, TYT = 16, TXT = 8, OFF = 2;
constexpr int HZ = RZ + C - 1, HX = TXT + KX - 1;
constexpr int WS_C = (TYT * 4 + OFF + KY - 1 + 3) / 4 * 4;

__device__ __forceinline__ void chunk(float (&acc)[RZ][4], const float* __restrict__ sxp, int zstride, const float* __restrict__ skp) {
    constexpr int WN = (OFF + KY + 3 + 3) / 4 * 4;
    float tap[48];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const float4 v = reinterpret_cast<const float4*>(skp)[i];
        tap[4 * i] = v.x; tap[4 * i + 1] = v.y; tap[4 * i + 2] = v.z; tap[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int zi = 0; zi < RZ + C - 1; ++zi) {
        float win[WN];
#pragma unroll
        for (int i = 0; i < WN / 4; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(sxp + zi * zstride + 4 * i);
            win[4 * i] = v.x; win[4 * i + 1] = v.y; win[4 * i + 2] = v.z; win[4 * i + 3] = v.w;
        }
#pragma unroll
        for (int dzl = 0; dzl < C; ++dzl) {
            const int zo = zi - dzl;
            if (zo >= 0 && zo < RZ) {
#pragma unroll
                for (int dy = 0; dy < KY; ++dy)
#pragma unroll
                    for (int r = 0; r < 4; ++r) acc[zo][r] = fmaf(win[OFF + r + dy], tap[dzl * KY + dy], acc[zo][r]);
            }
        }
    }
}

// MODE bit0: runtime strides; bit1: per-tile sync + accumulator reset + store; bit2: sparse binary data
template <int MODE>
__global__ void __launch_bounds__(128, 2) k(float* out, int iters, int ws_rt, int hx_rt, int kx_rt) {
    extern __shared__ __align__(128) float smem[];
    const int WS = (MODE & 1) ? ws_rt : WS_C;
    const int HXr = (MODE & 1) ? hx_rt : HX;
    const int KXr = (MODE & 1) ? kx_rt : KX;
    float* sx = smem;
    float* sk = smem + HZ * HX * WS_C;
    for (int i = threadIdx.x; i < HZ * HX * WS_C; i += 128)
        sx[i] = (MODE & 4) ? (((i * 2654435761u) >> 26) == 0 ? 1.f : 0.f) : (float)((i * 7) % 3) * 0.5f;
    for (int i = threadIdx.x; i < KX * 48; i += 128) sk[i] = 0.01f * (float)(i % 11);
    __syncthreads();
    const int tyi = threadIdx.x % TYT, txi = threadIdx.x / TYT;
    float acc[RZ][4];
    for (int i = 0; i < RZ; ++i) for (int r = 0; r < 4; ++r) acc[i][r] = 0.f;
    for (int it = 0; it < iters; ++it) {
        if (MODE & 2) {
#pragma unroll
            for (int i = 0; i < RZ; ++i)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[i][r] = 0.f;
        }
        for (int dx = 0; dx < KXr; ++dx)
            chunk(acc, sx + (txi + dx) * WS + 4 * tyi, HXr * WS, sk + dx * 48);
        if (MODE & 2) {
            __syncthreads();
            float4* o = reinterpret_cast<float4*>(out) + ((size_t)(blockIdx.x * 8 + (it & 7)) * 8) * 128 + threadIdx.x;
#pragma unroll
            for (int i = 0; i < RZ; ++i) o[i * 128] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
    }
    float s = 0; for (int i = 0; i < RZ; ++i) for (int r = 0; r < 4; ++r) s += acc[i][r];
    if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char* name, int iters) {
    float* out; cudaMalloc(&out, (size_t)296 * 8 * 8 * 128 * 16 + 1024);
    const size_t smem = 112 * 1024;  // forces 2 CTAs/SM like the shipped kernel
    auto kern = k<MODE>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, smem);
    const int blocks = 148 * occ;
    kern<<<blocks, 128, smem>>>(out, 2, WS_C, HX, KX);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        kern<<<blocks, 128, smem>>>(out, iters, WS_C, HX, KX);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    const double fl = 2.0 * RZ * 4 * C * KY * KX * (double)iters * blocks * 128;
    printf("%-52s occ %d  %8.3f ms  %7.2f TFLOP/s  (%s)\n", name, occ, best, fl / (best * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    run<0>("baseline (constexpr strides)", 400);
    run<1>("runtime strides / kx", 400);
    run<2>("per-tile reset + sync + float4 stores", 400);
    run<3>("runtime strides + per-tile reset/sync/stores", 400);
    run<4>("sparse binary data", 400);
    run<7>("all three", 400);
    run<7>("all three, 7 iterations (one launch of the real size)", 7);
    return 0;
}
