// Observer backward, data part: the tap gradient
//     W[t] = sum_{b,v} G0[b,v] * xpad[b, v + t],   G0 = dpred * (1 - pred^2) * [pred > 0]
// Replaces aten::convolution_backward (weight gradient) + the relu/tanh backward that
// autograd runs for SCENE_Net.py:325-337.  W is ONE T-vector (not G of them): the per-
// operator gradients are lambda_g * W (synth.cu::sn_scenenet_param_grads).
//
// Pass 1 (g0_kernel, HBM-bound elementwise): G0 in float64 from pred/dpred, rounded once to float32.
// Pass 2 (stencil_bwd_kernel, FP32-bound): persistent CTAs walk 8 x TX x TY voxel tiles (static
// round-robin => deterministic).  Per tile two TMA loads (x halo tile, G0 tile) land on one mbarrier.
// Warp w of a CTA owns tap group (dx, z-chunk) for the whole kernel and keeps its C*KY
// accumulators in registers; its lanes sweep the tile's 8x4 micro-tiles.  At the end every
// warp reduces its accumulators across lanes in float64 and writes one partial row; a second
// tiny kernel sums the rows in fixed order (no floating-point atomics anywhere).
#include <stdlib.h>
#include "stencil_common.cuh"
#include "tma_host.cuh"

namespace sn {

constexpr int kBwdMaxWarps = 8;

template <int KY, int CS>
__device__ __forceinline__ void bwd_chunk(float (&acc)[Geo<KY>::C * KY], const float* __restrict__ sxp, int zstride,
                                          const float* __restrict__ sgp, int gzstride) {
    constexpr int WN = Geo<KY>::WN;
    float g[kRZ][4];
    bool any = false;
#pragma unroll
    for (int z = 0; z < kRZ; ++z) {
        const float4 v = *reinterpret_cast<const float4*>(sgp + z * gzstride);
        g[z][0] = v.x; g[z][1] = v.y; g[z][2] = v.z; g[z][3] = v.w;
        any |= (v.x != 0.f) | (v.y != 0.f) | (v.z != 0.f) | (v.w != 0.f);
    }
    (void)any;
#pragma unroll
    for (int zi = 0; zi < kRZ + CS - 1; ++zi) {
        float win[WN];
#pragma unroll
        for (int i = 0; i < WN / 4; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(sxp + zi * zstride + 4 * i);
            win[4 * i] = v.x; win[4 * i + 1] = v.y; win[4 * i + 2] = v.z; win[4 * i + 3] = v.w;
        }
#pragma unroll
        for (int dzl = 0; dzl < CS; ++dzl) {
            const int zo = zi - dzl;
            if (zo >= 0 && zo < kRZ) {
#pragma unroll
                for (int dy = 0; dy < KY; ++dy) {
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        acc[dzl * KY + dy] = fmaf(g[zo][r], win[Geo<KY>::OFF + r + dy], acc[dzl * KY + dy]);
                }
            }
        }
    }
}

template <int KY, int CS>
struct BwdChunkSwitch {
    __device__ static __forceinline__ void run(int cs, float (&acc)[Geo<KY>::C * KY], const float* sxp, int zstride,
                                               const float* sgp, int gzstride) {
        if (cs == CS)
            bwd_chunk<KY, CS>(acc, sxp, zstride, sgp, gzstride);
        else
            BwdChunkSwitch<KY, CS - 1>::run(cs, acc, sxp, zstride, sgp, gzstride);
    }
};
template <int KY>
struct BwdChunkSwitch<KY, 0> {
    __device__ static __forceinline__ void run(int, float (&)[Geo<KY>::C * KY], const float*, int, const float*, int) {}
};

// G0 = dpred * (1 - pred^2) * [pred > 0]: 4 voxels per thread, all loads issued before use
template <typename TP, typename TD>
__global__ void __launch_bounds__(256) g0_kernel(const TP* __restrict__ pred, const TD* __restrict__ dpred,
                                                 float* __restrict__ g0, long long n) {
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        double pv[4], dv[4];
        if constexpr (sizeof(TP) == 8) {
            const double2 a = reinterpret_cast<const double2*>(pred)[2 * i], c = reinterpret_cast<const double2*>(pred)[2 * i + 1];
            pv[0] = a.x; pv[1] = a.y; pv[2] = c.x; pv[3] = c.y;
        } else {
            const float4 a = reinterpret_cast<const float4*>(pred)[i];
            pv[0] = a.x; pv[1] = a.y; pv[2] = a.z; pv[3] = a.w;
        }
        if constexpr (sizeof(TD) == 8) {
            const double2 a = reinterpret_cast<const double2*>(dpred)[2 * i], c = reinterpret_cast<const double2*>(dpred)[2 * i + 1];
            dv[0] = a.x; dv[1] = a.y; dv[2] = c.x; dv[3] = c.y;
        } else {
            const float4 a = reinterpret_cast<const float4*>(dpred)[i];
            dv[0] = a.x; dv[1] = a.y; dv[2] = a.z; dv[3] = a.w;
        }
        reinterpret_cast<float4*>(g0)[i] = make_float4(g0_of(pv[0], dv[0]), g0_of(pv[1], dv[1]), g0_of(pv[2], dv[2]), g0_of(pv[3], dv[3]));
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const long long i = (n4 << 2) + threadIdx.x;
        g0[i] = g0_of((double)pred[i], (double)dpred[i]);
    }
}

template <int KY, int TYT>
__global__ void __launch_bounds__(kBwdMaxWarps * 32, 2)
stencil_bwd_kernel(const BwdParams p, const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap gmap) {
    constexpr int C = Geo<KY>::C;
    constexpr int NACC = C * KY;
    constexpr int TY = TYT * 4, TX = kStencilThreads / TYT;
    constexpr int MICRO = TX * TYT;  // micro-tiles per CTA tile (z extent of a tile == kRZ)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TileGeo g = make_geo<KY, TYT>(p.B, p.Z, p.X, p.Y, p.kz, p.kx);
    const int halo_floats = g.HZ * g.HX * g.WS;
    float* sx = reinterpret_cast<float*>(smem_raw);
    float* sg = sx + ((halo_floats + 31) & ~31);  // [kRZ][TX][TY]
    uint64_t* bar = reinterpret_cast<uint64_t*>(sg + kRZ * TX * TY);

    const int tid = threadIdx.x, nthreads = blockDim.x, warp = tid >> 5, lane = tid & 31;
    const int combo = blockIdx.y * p.combos_per_cta + warp;
    const bool active = combo < p.ncombos;
    const int dx = active ? combo / g.nchunks : 0, ch = active ? combo % g.nchunks : 0;
    const int cs = min(C, p.kz - ch * C);

    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.f;

    if (p.use_tma && tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    uint32_t phase = 0;
    const int zstride = g.HX * g.WS, gzstride = TX * TY;
    // co-resident CTAs (blockIdx.x, +148, +296, ... land on the same SM) start out of phase so that their
    // TMA waits do not coincide: a single-buffered CTA is idle while its tile loads
    if (p.stagger_ns > 0) __nanosleep((unsigned)((blockIdx.x / kNumSMs) * p.stagger_ns));

    for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
        int b, z0, x0, y0;
        decode_tile(tile, g, b, z0, x0, y0);
        __syncthreads();  // everyone is done with the previous tile's shared memory
        if (p.use_tma) {
            if (tid == 0) {
                fence_proxy_async();
                mbar_arrive_expect_tx(bar, (uint32_t)(halo_floats + kRZ * TX * TY) * 4u);
                tma_load_4d(sx, &tmap, bar, y0 - g.ply, x0 - g.plx, z0 - g.plz, b);
                tma_load_4d(sg, &gmap, bar, y0, x0, z0, b);
            }
        } else {
            load_halo_plain(sx, p.x, g, p.Z, p.X, p.Y, b, z0, x0, y0, nthreads);
            for (int i = tid; i < kRZ * TX * TY; i += nthreads) {
                const int yy = i % TY, xx = (i / TY) % TX, zz = i / (TY * TX);
                const int gz = z0 + zz, gx = x0 + xx, gy = y0 + yy;
                sg[i] = (gz < p.Z && gx < p.X && gy < p.Y) ? __ldg(p.g0 + (((size_t)b * p.Z + gz) * p.X + gx) * p.Y + gy) : 0.f;
            }
        }
        __syncthreads();
        if (p.use_tma) {
            mbar_wait(bar, phase);
            phase ^= 1;
        }
        if (active) {
            for (int m = lane; m < MICRO; m += 32) {
                const int tyi = m % TYT, txi = m / TYT;
                const float* sxp = sx + (ch * C) * zstride + (txi + dx) * g.WS + 4 * tyi;
                const float* sgp = sg + txi * TY + 4 * tyi;
                if (cs == C)
                    bwd_chunk<KY, C>(acc, sxp, zstride, sgp, gzstride);
                else
                    BwdChunkSwitch<KY, C - 1>::run(cs, acc, sxp, zstride, sgp, gzstride);
            }
        }
    }

    // cross-lane reduction in float64, one partial row per CTA column (blockIdx.x)
    if (active) {
        double* row = p.partial + (size_t)blockIdx.x * p.TP;
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            const double s = warp_sum((double)acc[i]);
            const int dzl = i / KY, dy = i % KY;
            if (lane == 0 && dzl < cs) row[((ch * C + dzl) * p.kx + dx) * KY + dy] = s;
        }
    }
}

// W[t] = sum over partial rows in a fixed order (deterministic): block = 32 taps x 32 row groups
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const double* __restrict__ partial, int rows, int TP, int T,
                                                               double* __restrict__ W) {
    __shared__ double red[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int t = blockIdx.x * 32 + tx;
    double s = 0.0;
    if (t < T) {
        // independent loads first (memory-level parallelism), then a fixed-order sum
        double v[16];
        int r = ty;
        while (r < rows) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = (r + 32 * i < rows) ? partial[(size_t)(r + 32 * i) * TP + t] : 0.0;
#pragma unroll
            for (int i = 0; i < 16; ++i) s += v[i];
            r += 32 * 16;
        }
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && t < T) {
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < 32; ++i) a += red[i][tx];
        W[t] = a;
    }
}

struct BwdPlan {
    int grid_x, grid_y, combos_per_cta, ncombos, TP;
    size_t smem;
};

template <int KY, int TYT>
static BwdPlan plan_bwd(int B, int Z, int X, int Y, int kz, int kx) {
    const TileGeo g = make_geo<KY, TYT>(B, Z, X, Y, kz, kx);
    BwdPlan pl;
    pl.ncombos = kx * g.nchunks;
    pl.grid_y = ceil_div(pl.ncombos, kBwdMaxWarps);
    pl.combos_per_cta = ceil_div(pl.ncombos, pl.grid_y);
    const int halo_floats = g.HZ * g.HX * g.WS;
    pl.smem = (size_t)(((halo_floats + 31) & ~31) + kRZ * g.TX * g.TY) * 4 + 16;
    int per_sm = (int)((227 * 1024) / (pl.smem + 1024));
    const int thr = pl.combos_per_cta * 32;
    per_sm = min(per_sm, 2048 / thr);
    per_sm = min(per_sm, 65536 / (thr * 128));
    per_sm = max(1, min(per_sm, 4));
    int gx = kNumSMs * per_sm / pl.grid_y;
    gx = max(1, min(gx, g.ntiles));
    pl.grid_x = gx;
    pl.TP = (kz * kx * KY + 31) & ~31;
    return pl;
}

static inline int64_t g0_bytes(int B, int Z, int X, int Y) { return (((int64_t)B * Z * X * Y * 4) + 255) & ~(int64_t)255; }

static int launch_g0(const BwdParams& p, float* g0, cudaStream_t stream) {
    const long long n = (long long)p.B * p.Z * p.X * p.Y;
    long long blocks = ceil_div64(n / 4 + 1, 256);
    const int grid = (int)(blocks > (long long)kNumSMs * 16 ? (long long)kNumSMs * 16 : blocks);
    if (p.pred_f64 && p.dpred_f64)
        g0_kernel<double, double><<<grid, 256, 0, stream>>>((const double*)p.pred, (const double*)p.dpred, g0, n);
    else if (p.pred_f64)
        g0_kernel<double, float><<<grid, 256, 0, stream>>>((const double*)p.pred, (const float*)p.dpred, g0, n);
    else if (p.dpred_f64)
        g0_kernel<float, double><<<grid, 256, 0, stream>>>((const float*)p.pred, (const double*)p.dpred, g0, n);
    else
        g0_kernel<float, float><<<grid, 256, 0, stream>>>((const float*)p.pred, (const float*)p.dpred, g0, n);
    SN_LAUNCH_CHECK();
    return SN_OK;
}

template <int KY, int TYT>
static int launch_bwd(BwdParams p, double* W, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    const BwdPlan pl = plan_bwd<KY, TYT>(p.B, p.Z, p.X, p.Y, p.kz, p.kx);
    if (pl.smem > 227 * 1024) return SN_ERR_UNSUPPORTED;
    const int64_t gb = g0_bytes(p.B, p.Z, p.X, p.Y);
    if (gb + (int64_t)pl.grid_x * pl.TP * 8 > ws_bytes) return SN_ERR_WORKSPACE;
    if ((((uintptr_t)p.pred) | ((uintptr_t)p.dpred)) & 15) return SN_ERR_ALIGN;
    const TileGeo g = make_geo<KY, TYT>(p.B, p.Z, p.X, p.Y, p.kz, p.kx);
    float* g0 = reinterpret_cast<float*>(ws);
    p.g0 = g0;
    p.partial = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + gb);
    int rc = launch_g0(p, g0, stream);
    if (rc) return rc;
    CUtensorMap tmap, gmap;
    const bool ok_x = make_grid_tmap(&tmap, p.x, p.B, p.Z, p.X, p.Y, g.HZ, g.HX, g.WS);
    const bool ok_g = make_grid_tmap(&gmap, g0, p.B, p.Z, p.X, p.Y, kRZ, g.TX, g.TY);
    p.use_tma = (ok_x && ok_g) ? 1 : 0;
    p.ncombos = pl.ncombos;
    p.combos_per_cta = pl.combos_per_cta;
    p.TP = pl.TP;
    {
        const char* e = getenv("SN_BWD_STAGGER_NS");
        p.stagger_ns = e ? atoi(e) : 0;
    }
    auto kern = stencil_bwd_kernel<KY, TYT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return cuda_rc(e);
    kern<<<dim3(pl.grid_x, pl.grid_y), pl.combos_per_cta * 32, pl.smem, stream>>>(p, tmap, gmap);
    SN_LAUNCH_CHECK();
    const int T = p.kz * p.kx * KY;
    reduce_partials_kernel<<<ceil_div(T, 32), dim3(32, 32), 0, stream>>>(p.partial, pl.grid_x, pl.TP, T, W);
    SN_LAUNCH_CHECK();
    return SN_OK;
}

int stencil_bwd_generic(const BwdParams& p, int ky, double* W, cudaStream_t stream);  // stencil_generic.cu

template <int KY>
static int dispatch_bwd_ty(const BwdParams& p, double* W, void* ws, int64_t wsb, cudaStream_t s) {
    return p.Y > 32 ? launch_bwd<KY, 16>(p, W, ws, wsb, s) : launch_bwd<KY, 8>(p, W, ws, wsb, s);
}
template <int KY>
static int64_t ws_bytes_ty(int B, int Z, int X, int Y, int kz, int kx) {
    const BwdPlan pl = Y > 32 ? plan_bwd<KY, 16>(B, Z, X, Y, kz, kx) : plan_bwd<KY, 8>(B, Z, X, Y, kz, kx);
    return g0_bytes(B, Z, X, Y) + (int64_t)pl.grid_x * pl.TP * 8;
}

}  // namespace sn

extern "C" int64_t sn_scenenet_bwd_workspace_bytes(int B, int Z, int X, int Y, int kz, int kx, int ky) {
    if (B < 1 || Z < 1 || X < 1 || Y < 1 || kz < 1 || kx < 1 || ky < 1) return SN_ERR_BAD_ARG;
    switch (ky) {
        case 3: return sn::ws_bytes_ty<3>(B, Z, X, Y, kz, kx);
        case 5: return sn::ws_bytes_ty<5>(B, Z, X, Y, kz, kx);
        case 6: return sn::ws_bytes_ty<6>(B, Z, X, Y, kz, kx);
        case 7: return sn::ws_bytes_ty<7>(B, Z, X, Y, kz, kx);
        case 9: return sn::ws_bytes_ty<9>(B, Z, X, Y, kz, kx);
        case 11: return sn::ws_bytes_ty<11>(B, Z, X, Y, kz, kx);
        case 13: return sn::ws_bytes_ty<13>(B, Z, X, Y, kz, kx);
        case 15: return sn::ws_bytes_ty<15>(B, Z, X, Y, kz, kx);
        default: return 256;  // generic path needs no workspace
    }
}

extern "C" int sn_scenenet_bwd(const float* x, const void* pred, int pred_dtype, const void* dpred, int dpred_dtype,
                               int B, int Z, int X, int Y, int kz, int kx, int ky, double* W, void* ws, int64_t ws_bytes,
                               void* stream) {
    if (!x || !pred || !dpred || !W) return SN_ERR_BAD_ARG;
    if (B < 1 || Z < 1 || X < 1 || Y < 1 || kz < 1 || kx < 1 || ky < 1) return SN_ERR_BAD_ARG;
    if ((pred_dtype != SN_F32 && pred_dtype != SN_F64) || (dpred_dtype != SN_F32 && dpred_dtype != SN_F64))
        return SN_ERR_BAD_ARG;
    if ((long long)kz * kx * ky > SN_MAX_TAPS) return SN_ERR_UNSUPPORTED;
    if (ws && ((uintptr_t)ws & 15)) return SN_ERR_ALIGN;
    sn::BwdParams p{};
    p.x = x; p.pred = pred; p.dpred = dpred;
    p.B = B; p.Z = Z; p.X = X; p.Y = Y; p.kz = kz; p.kx = kx;
    p.pred_f64 = pred_dtype == SN_F64; p.dpred_f64 = dpred_dtype == SN_F64;
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if (!ws && ky != 0) ws_bytes = 0;
    switch (ky) {
        case 3: rc = sn::dispatch_bwd_ty<3>(p, W, ws, ws_bytes, s); break;
        case 5: rc = sn::dispatch_bwd_ty<5>(p, W, ws, ws_bytes, s); break;
        case 6: rc = sn::dispatch_bwd_ty<6>(p, W, ws, ws_bytes, s); break;
        case 7: rc = sn::dispatch_bwd_ty<7>(p, W, ws, ws_bytes, s); break;
        case 9: rc = sn::dispatch_bwd_ty<9>(p, W, ws, ws_bytes, s); break;
        case 11: rc = sn::dispatch_bwd_ty<11>(p, W, ws, ws_bytes, s); break;
        case 13: rc = sn::dispatch_bwd_ty<13>(p, W, ws, ws_bytes, s); break;
        case 15: rc = sn::dispatch_bwd_ty<15>(p, W, ws, ws_bytes, s); break;
        default: rc = SN_ERR_UNSUPPORTED; break;
    }
    if (rc == SN_ERR_UNSUPPORTED) rc = sn::stencil_bwd_generic(p, ky, W, s);
    return rc;
}
