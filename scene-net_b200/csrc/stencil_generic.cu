// Correct-but-slow fallbacks for kernel extents the register-blocked stencils are not
// instantiated for (ky outside {3,5,6,7,9,11,13,15}).  Same arithmetic contract.
#include "stencil_common.cuh"

namespace sn {

__global__ void __launch_bounds__(256) fwd_generic_kernel(const FwdParams p, int ky) {
    const long long V = (long long)p.B * p.Z * p.X * p.Y;
    const int plz = pad_left(p.kz), plx = pad_left(p.kx), ply = pad_left(ky);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i % p.Y), x = (int)((i / p.Y) % p.X), z = (int)((i / ((long long)p.Y * p.X)) % p.Z);
        const long long b = i / ((long long)p.Y * p.X * p.Z);
        float s = 0.f;
        for (int dz = 0; dz < p.kz; ++dz) {
            const int gz = z + dz - plz;
            if (gz < 0 || gz >= p.Z) continue;
            for (int dx = 0; dx < p.kx; ++dx) {
                const int gx = x + dx - plx;
                if (gx < 0 || gx >= p.X) continue;
                const float* row = p.x + ((b * p.Z + gz) * p.X + gx) * p.Y;
                const float* kr = p.Kstar + (dz * p.kx + dx) * ky;
                for (int dy = 0; dy < ky; ++dy) {
                    const int gy = y + dy - ply;
                    if (gy >= 0 && gy < p.Y) s = fmaf(__ldg(row + gy), __ldg(kr + dy), s);
                }
            }
        }
        const float o = s > 0.f ? tanhf(s) : 0.f;
        if (p.out_f64)
            reinterpret_cast<double*>(p.pred)[i] = (double)o;
        else
            reinterpret_cast<float*>(p.pred)[i] = o;
    }
}

// one CTA per tap (G0 precomputed)
__global__ void __launch_bounds__(256) tapgrad_generic_kernel(const BwdParams p, int ky, double* __restrict__ W) {
    __shared__ double red[8];
    const int t = blockIdx.x;
    const int dy = t % ky, dx = (t / ky) % p.kx, dz = t / (ky * p.kx);
    const int oz = dz - pad_left(p.kz), ox = dx - pad_left(p.kx), oy = dy - pad_left(ky);
    const long long V = (long long)p.B * p.Z * p.X * p.Y;
    double s = 0.0;
    for (long long i = threadIdx.x; i < V; i += blockDim.x) {
        const int y = (int)(i % p.Y), x = (int)((i / p.Y) % p.X), z = (int)((i / ((long long)p.Y * p.X)) % p.Z);
        const long long b = i / ((long long)p.Y * p.X * p.Z);
        const int gz = z + oz, gx = x + ox, gy = y + oy;
        if (gz < 0 || gz >= p.Z || gx < 0 || gx >= p.X || gy < 0 || gy >= p.Y) continue;
        const float g = __ldg(p.g0 + i);
        if (g == 0.f) continue;
        s += (double)(g * __ldg(p.x + ((b * p.Z + gz) * p.X + gx) * p.Y + gy));
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
        for (int i = 0; i < 8; ++i) a += red[i];
        W[t] = a;
    }
}

int stencil_fwd_generic(const FwdParams& p, int ky, cudaStream_t stream) {
    const long long V = (long long)p.B * p.Z * p.X * p.Y;
    const int grid = (int)min((long long)kNumSMs * 16, ceil_div64(V, 256));
    fwd_generic_kernel<<<grid, 256, 0, stream>>>(p, ky);
    SN_LAUNCH_CHECK();
    return SN_OK;
}

int stencil_tapgrad_generic(const BwdParams& p, int ky, double* W, cudaStream_t stream) {
    tapgrad_generic_kernel<<<p.kz * p.kx * ky, 256, 0, stream>>>(p, ky, W);
    SN_LAUNCH_CHECK();
    return SN_OK;
}

}  // namespace sn
