// What FFMA rate can the SM sustain for the stencil's operand pattern (registers only, no memory)?
#include <cstdio>
#include <cuda_runtime.h>

constexpr int RZ = 8, KY = 5, C = 9, WN = 12, OFF = 2;

// variant 0: probe-style chains with an immediate multiplier (the 73 TF "peak" number)
// variant 1: chains with three register operands, accumulators independent
// variant 2: the stencil body exactly as in stencil_fwd.cu (win rows regenerated arithmetically)
template <int V>
__global__ void __launch_bounds__(128) k(float* out, int iters, float seed) {
    if (V == 0) {
        float a[8];
        for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
        const float c = seed;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int u = 0; u < 180; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], 0.999f, c);
        float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
        if (s == 123.456f) out[0] = s;
    } else if (V == 1) {
        float a[32], w[8], t[5];
        for (int i = 0; i < 32; ++i) a[i] = threadIdx.x * 1e-3f + i;
        for (int i = 0; i < 8; ++i) w[i] = seed + i;
        for (int i = 0; i < 5; ++i) t[i] = seed * 0.5f + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 9; ++u)
#pragma unroll
                for (int z = 0; z < 8; ++z)
#pragma unroll
                    for (int dy = 0; dy < 5; ++dy)
#pragma unroll
                        for (int r = 0; r < 4; ++r) a[z * 4 + r] = fmaf(w[r + dy], t[dy], a[z * 4 + r]);
            for (int i = 0; i < 8; ++i) w[i] += a[i] * 1e-30f;
        }
        float s = 0; for (int i = 0; i < 32; ++i) s += a[i];
        if (s == 123.456f) out[0] = s;
    } else {
        float acc[RZ][4], tap[C * KY];
        for (int i = 0; i < RZ; ++i) for (int r = 0; r < 4; ++r) acc[i][r] = threadIdx.x * 1e-3f + i + r;
        for (int i = 0; i < C * KY; ++i) tap[i] = seed + i * 0.01f;
        float base = seed;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int zi = 0; zi < RZ + C - 1; ++zi) {
                float win[WN];
#pragma unroll
                for (int i = 0; i < WN; ++i) win[i] = base + (float)(zi * WN + i);  // stands in for the LDS
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
            base += acc[0][0] * 1e-30f;
        }
        float s = 0; for (int i = 0; i < RZ; ++i) for (int r = 0; r < 4; ++r) s += acc[i][r];
        if (s == 123.456f) out[0] = s;
    }
}

template <int V>
void run(const char* name, double fma_per_iter, int iters) {
    float* out; cudaMalloc(&out, 4);
    const int blocks = 148 * 4;
    k<V><<<blocks, 128>>>(out, 10, 1.f);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        k<V><<<blocks, 128>>>(out, iters, 1.f);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    const double fl = 2.0 * fma_per_iter * iters * blocks * 128;
    printf("%-40s %8.3f ms  %7.2f TFLOP/s  (%s)\n", name, best, fl / (best * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    run<0>("imm-form chains (8 per thread)", 8.0 * 180, 2000);
    run<1>("3-reg outer product 32 acc", 9.0 * 8 * 5 * 4, 2000);
    run<2>("stencil body, registers only", (double)RZ * 4 * C * KY, 2000);
    return 0;
}
