// Elementwise helpers, library info and the FP32-pipe peak probe.
#include "common.cuh"

namespace sn {
long long g_launch_count = 0;

// float64 -> float32, 2 doubles per thread per step (16-byte loads), grid-stride
__global__ void __launch_bounds__(256) cast_f64_f32_kernel(const double* __restrict__ in, float* __restrict__ out, long long n) {
    const long long n2 = n >> 1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        const double2 v = reinterpret_cast<const double2*>(in)[i];
        reinterpret_cast<float2*>(out)[i] = make_float2((float)v.x, (float)v.y);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) out[n - 1] = (float)in[n - 1];
}

// uint8 / bool occupancy -> float32, 16 voxels per thread per step (one 16-byte load, four 16-byte stores)
__global__ void __launch_bounds__(256) cast_u8_f32_kernel(const unsigned char* __restrict__ in, float* __restrict__ out, long long n) {
    const long long n16 = n >> 4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 v = reinterpret_cast<const uint4*>(in)[i];
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
        float4* o = reinterpret_cast<float4*>(out) + 4 * i;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            o[k] = make_float4((float)(w[k] & 0xffu), (float)((w[k] >> 8) & 0xffu), (float)((w[k] >> 16) & 0xffu), (float)(w[k] >> 24));
    }
    if (blockIdx.x == 0)
        for (long long i = (n16 << 4) + threadIdx.x; i < n; i += blockDim.x) out[i] = (float)in[i];
}


// ---- sn_grid_prepare: cast to float32 + count the non-zero voxels (the count selects the occupancy-driven
// kernels on the device).  Integer atomics only: the count is exact and order-independent.
// (one atomic per CTA: thousands of same-address atomics at the end of the kernel were a visible serial tail)
__device__ __forceinline__ void add_count(unsigned cnt, unsigned long long* nnz) {
    __shared__ unsigned s_cnt[8];
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += s_cnt[i];
        if (t) atomicAdd(nnz, (unsigned long long)t);
    }
}

__global__ void __launch_bounds__(256) prepare_f64_kernel(const double* __restrict__ in, float* __restrict__ out, long long n,
                                                          unsigned long long* nnz) {
    // 4 x 16-byte loads in flight per thread (one per step left the pass at 0.73 of the HBM copy rate)
    const long long n2 = n >> 1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned cnt = 0;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n2; i += 4 * stride) {
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = reinterpret_cast<const double2*>(in)[i + u * stride];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float2 o = make_float2((float)v[u].x, (float)v[u].y);
            reinterpret_cast<float2*>(out)[i + u * stride] = o;
            cnt += (o.x != 0.f) + (o.y != 0.f);
        }
    }
    for (; i < n2; i += stride) {
        const double2 v = reinterpret_cast<const double2*>(in)[i];
        const float2 o = make_float2((float)v.x, (float)v.y);
        reinterpret_cast<float2*>(out)[i] = o;
        cnt += (o.x != 0.f) + (o.y != 0.f);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        out[n - 1] = (float)in[n - 1];
        cnt += out[n - 1] != 0.f;
    }
    add_count(cnt, nnz);
}

__global__ void __launch_bounds__(256) prepare_u8_kernel(const unsigned char* __restrict__ in, float* __restrict__ out, long long n,
                                                         unsigned long long* nnz) {
    const long long n16 = n >> 4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned cnt = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 v = reinterpret_cast<const uint4*>(in)[i];
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
        float4* o = reinterpret_cast<float4*>(out) + 4 * i;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned b0 = w[k] & 0xffu, b1 = (w[k] >> 8) & 0xffu, b2 = (w[k] >> 16) & 0xffu, b3 = w[k] >> 24;
            o[k] = make_float4((float)b0, (float)b1, (float)b2, (float)b3);
            cnt += (b0 != 0) + (b1 != 0) + (b2 != 0) + (b3 != 0);
        }
    }
    if (blockIdx.x == 0)
        for (long long i = (n16 << 4) + threadIdx.x; i < n; i += blockDim.x) {
            out[i] = (float)in[i];
            cnt += in[i] != 0;
        }
    add_count(cnt, nnz);
}

__global__ void __launch_bounds__(256) count_f32_kernel(const float* __restrict__ in, long long n, unsigned long long* nnz) {
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned cnt = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = reinterpret_cast<const float4*>(in)[i];
        cnt += (v.x != 0.f) + (v.y != 0.f) + (v.z != 0.f) + (v.w != 0.f);
    }
    if (blockIdx.x == 0)
        for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) cnt += in[i] != 0.f;
    add_count(cnt, nnz);
}

template <typename T>
__global__ void __launch_bounds__(256) threshold_kernel(const T* __restrict__ p, T tau, long long n, T* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = p[i] >= tau ? (T)1 : (T)0;
}

// FP32 FMA-pipe probe: 8 independent dependent-chains per thread, 1024 threads/SM-slot
constexpr int kProbeUnroll = 64;
__global__ void __launch_bounds__(256) fp32_probe_kernel(float* sink, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f, c = 1e-4f + blockIdx.x * 1e-9f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kProbeUnroll; ++u) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    const float s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 12345.678f) sink[0] = s;  // never true in practice; keeps the chains alive
}

static inline int grid_for(long long n, int per_block) {
    long long b = ceil_div64(n, per_block);
    const long long cap = (long long)kNumSMs * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
}  // namespace sn

extern "C" int sn_abi_version(void) { return SN_ABI_VERSION; }
extern "C" const char* sn_build_info(void) { return "scenenet_b200 sm_100a nvcc " __DATE__ " " __TIME__; }
extern "C" int64_t sn_launch_count(void) { return sn::g_launch_count; }

extern "C" int sn_cast_f64_to_f32(const double* in, float* out, int64_t n, void* stream) {
    if (!in || !out || n < 0) return SN_ERR_BAD_ARG;
    if (n == 0) return SN_OK;
    if (((uintptr_t)in & 15) || ((uintptr_t)out & 7)) return SN_ERR_ALIGN;
    sn::cast_f64_f32_kernel<<<sn::grid_for(n / 2 + 1, 256 * 4), 256, 0, (cudaStream_t)stream>>>(in, out, n);
    SN_LAUNCH_CHECK();
    return SN_OK;
}

extern "C" int sn_cast_u8_to_f32(const unsigned char* in, float* out, int64_t n, void* stream) {
    if (!in || !out || n < 0) return SN_ERR_BAD_ARG;
    if (n == 0) return SN_OK;
    if (((uintptr_t)in & 15) || ((uintptr_t)out & 15)) return SN_ERR_ALIGN;
    sn::cast_u8_f32_kernel<<<sn::grid_for(n / 16 + 1, 256 * 2), 256, 0, (cudaStream_t)stream>>>(in, out, n);
    SN_LAUNCH_CHECK();
    return SN_OK;
}


extern "C" int sn_grid_prepare(const void* x, int dtype, int64_t n, float* x32, unsigned long long* nnz, void* stream) {
    if (!x || !nnz || n < 0) return SN_ERR_BAD_ARG;
    if (dtype != SN_F32 && dtype != SN_F64 && dtype != SN_U8) return SN_ERR_BAD_ARG;
    if (dtype != SN_F32 && !x32) return SN_ERR_BAD_ARG;
    if (dtype == SN_F32 && x32 && (const void*)x32 != x) return SN_ERR_BAD_ARG;  // float32 grids are used in place
    if (((uintptr_t)x & 15) || ((uintptr_t)x32 & 15) || ((uintptr_t)nnz & 15)) return SN_ERR_ALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(nnz, 0, 2 * sizeof(unsigned long long), s);  // [0] count, [1] ticket of the tap-gradient tail
    if (e != cudaSuccess) return sn::cuda_rc(e);
    if (n == 0) return SN_OK;
    if (dtype == SN_F64)
        sn::prepare_f64_kernel<<<sn::grid_for(n / 2 + 1, 256 * 4), 256, 0, s>>>((const double*)x, x32, n, nnz);
    else if (dtype == SN_U8)
        sn::prepare_u8_kernel<<<sn::grid_for(n / 16 + 1, 256 * 2), 256, 0, s>>>((const unsigned char*)x, x32, n, nnz);
    else
        sn::count_f32_kernel<<<sn::grid_for(n / 4 + 1, 256 * 4), 256, 0, s>>>((const float*)x, n, nnz);
    SN_LAUNCH_CHECK();
    return SN_OK;
}

extern "C" int sn_threshold(const void* p, int dtype, double tau, int64_t n, void* out, void* stream) {
    if (!p || !out || n < 0) return SN_ERR_BAD_ARG;
    if (n == 0) return SN_OK;
    const int grid = sn::grid_for(n, 256 * 4);
    if (dtype == SN_F32)
        sn::threshold_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)p, (float)tau, n, (float*)out);
    else if (dtype == SN_F64)
        sn::threshold_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)p, tau, n, (double*)out);
    else
        return SN_ERR_BAD_ARG;
    SN_LAUNCH_CHECK();
    return SN_OK;
}

extern "C" int sn_fp32_peak_probe(float* sink, int iters, double* flops_out_host, void* stream) {
    if (!sink || iters < 1) return SN_ERR_BAD_ARG;
    const int blocks = sn::kNumSMs * 8, threads = 256;
    sn::fp32_probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(sink, iters);
    SN_LAUNCH_CHECK();
    if (flops_out_host) *flops_out_host = 2.0 * 8.0 * sn::kProbeUnroll * (double)iters * blocks * threads;
    return SN_OK;
}
