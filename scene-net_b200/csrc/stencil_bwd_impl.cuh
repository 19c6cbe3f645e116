// Implementation of the tap-gradient (backward) stencil, included by the per-KY translation units
// (stencil_bwd_ky*.cu) so that the 8 kernel widths compile in parallel.
#pragma once
#include <stdlib.h>
#include "stencil_common.cuh"
#include "tma_host.cuh"

namespace sn {

constexpr int kBwdMaxThreads = 640;  // 20 warps = 5 per scheduler; 96 registers per thread

template <int KY, int CS>
__device__ __forceinline__ void bwd_chunk(float (&acc)[Geo<KY>::C * KY], const float* __restrict__ sxp, int zstride,
                                          const float* __restrict__ sgp, int gzstride) {
    constexpr int WN = Geo<KY>::WN;
    float g[kRZ][4];
#pragma unroll
    for (int z = 0; z < kRZ; ++z) {
        const float4 v = *reinterpret_cast<const float4*>(sgp + z * gzstride);
        g[z][0] = v.x; g[z][1] = v.y; g[z][2] = v.z; g[z][3] = v.w;
    }
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

// One persistent CTA per SM.  Warp w owns tap group (dx, z-chunk) = combo w / Q for the whole kernel and
// keeps its C*KY accumulators in registers; the Q warps of a combo split a tile's 8x4 micro-tiles
// between them (Q = 4: one micro-tile per lane per tile, and the four warps of a combo sit on the four
// schedulers).  x halo + G0 tile arrive by TMA through a two-stage pipeline: tile k+2 is requested as
// soon as every warp has finished tile k.
template <int KY, int TYT, int REM>
__global__ void __launch_bounds__(kBwdMaxThreads, 1)
stencil_bwd_kernel(const BwdParams p, const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap gmap) {
    if (p.nnz && *p.nnz <= p.nnz_max) return;  // sparse input: tapgrad_sparse_kernel does the work
    constexpr int C = Geo<KY>::C;
    constexpr int NACC = C * KY;
    constexpr int TY = TYT * 4, TX = kStencilThreads / TYT;
    constexpr int MICRO = TX * TYT;  // micro-tiles per CTA tile (z extent of a tile == kRZ)
    constexpr int G0F = kRZ * TX * TY;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TileGeo g = make_geo<KY, TYT>(p.B, p.Z, p.X, p.Y, p.kz, p.kx);
    const int halo_floats = g.HZ * g.HX * g.WS;
    const int halo_stride = (halo_floats + 31) & ~31;
    const int stage = halo_stride + G0F;
    const int nstage = p.nstage;
    float* s0 = reinterpret_cast<float*>(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(s0 + nstage * stage);  // [2]

    const int tid = threadIdx.x, nthreads = blockDim.x, warp = tid >> 5, lane = tid & 31;
    const int Q = p.Q, q = warp % Q;
    const int combo = blockIdx.y * p.combos_per_cta + warp / Q;
    const bool active = combo < p.ncombos;
    const int dx = active ? combo / g.nchunks : 0, ch = active ? combo % g.nchunks : 0;
    const int nfull = p.kz / C;
    const int G = gridDim.x;

    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.f;

    auto issue = [&](int tile, int buf) {  // thread 0 only
        int b, z0, x0, y0;
        decode_tile(tile, g, b, z0, x0, y0);
        float* sx = s0 + buf * stage;
        mbar_arrive_expect_tx(&bar[buf], (uint32_t)(halo_floats + G0F) * 4u);
        tma_load_4d(sx, &tmap, &bar[buf], y0 - g.ply, x0 - g.plx, z0 - g.plz, b);
        tma_load_4d(sx + halo_stride, &gmap, &bar[buf], y0, x0, z0, b);
    };
    if (p.use_tma && tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
        if ((int)blockIdx.x < g.ntiles) issue(blockIdx.x, 0);
        if (nstage == 2 && (int)blockIdx.x + G < g.ntiles) issue(blockIdx.x + G, 1);
    }
    __syncthreads();

    const int zstride = g.HX * g.WS, gzstride = TX * TY;
    int k = 0;
    for (int tile = blockIdx.x; tile < g.ntiles; tile += G, ++k) {
        const int buf = (nstage == 2) ? (k & 1) : 0;
        const float* sx = s0 + buf * stage;
        const float* sg = sx + halo_stride;
        if (p.use_tma) {
            mbar_wait(&bar[buf], (uint32_t)(nstage == 2 ? (k >> 1) : k) & 1u);
        } else {
            int b, z0, x0, y0;
            decode_tile(tile, g, b, z0, x0, y0);
            __syncthreads();
            load_halo_plain(s0, p.x, g, p.Z, p.X, p.Y, b, z0, x0, y0, nthreads);
            float* sgw = s0 + halo_stride;
            for (int i = tid; i < G0F; i += nthreads) {
                const int yy = i % TY, xx = (i / TY) % TX, zz = i / (TY * TX);
                const int gz = z0 + zz, gx = x0 + xx, gy = y0 + yy;
                sgw[i] = (gz < p.Z && gx < p.X && gy < p.Y) ? __ldg(p.g0 + (((size_t)b * p.Z + gz) * p.X + gx) * p.Y + gy) : 0.f;
            }
            __syncthreads();
        }
        if (active) {
            for (int m = q * 32 + lane; m < MICRO; m += 32 * Q) {
                const int tyi = m % TYT, txi = m / TYT;
                const float* sxp = sx + (ch * C) * zstride + (txi + dx) * g.WS + 4 * tyi;
                const float* sgp = sg + txi * TY + 4 * tyi;
                if (REM == 0 || ch < nfull)
                    bwd_chunk<KY, C>(acc, sxp, zstride, sgp, gzstride);
                else if constexpr (REM > 0)
                    bwd_chunk<KY, REM>(acc, sxp, zstride, sgp, gzstride);
            }
        }
        if (p.use_tma) {
            __syncthreads();  // every warp is done with this stage -> refill it
            const int next = tile + nstage * G;
            if (tid == 0 && next < g.ntiles) {
                fence_proxy_async();
                issue(next, buf);
            }
        }
    }

    // cross-lane reduction in float64, then the Q warps of a tap group are combined through shared memory in a
    // fixed order (q = 0,1,..): one partial row per CTA column (blockIdx.x)
    const int cs = (REM == 0 || ch < nfull) ? C : REM;
    double* sred = reinterpret_cast<double*>(smem_raw);  // [warps][NACC]; the tile stages are dead by now
    __syncthreads();
    if (active) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            const double s = warp_sum((double)acc[i]);
            if (lane == 0) sred[warp * NACC + i] = s;
        }
    }
    __syncthreads();
    if (active && q == 0) {
        double* row = p.partial + (size_t)blockIdx.x * p.TP;
        for (int i = lane; i < NACC; i += 32) {
            double s = sred[warp * NACC + i];
            for (int j = 1; j < Q; ++j) s += sred[(warp + j) * NACC + i];
            const int dzl = i / KY, dy = i % KY;
            if (dzl < cs) row[((ch * C + dzl) * p.kx + dx) * KY + dy] = s;
        }
    }
    if (p.ticket) {
        __syncthreads();  // sred is reused as scratch
        last_cta_row_sum(p.ticket, p.partial, (int)gridDim.x, p.TP, p.kz * p.kx * KY, p.W, sred, (int)(gridDim.x * gridDim.y));
    }
}

struct BwdPlan {
    int grid_x, grid_y, combos_per_cta, ncombos, TP, Q, nstage, threads;
    size_t smem;
};

template <int KY, int TYT>
static BwdPlan plan_bwd(int B, int Z, int X, int Y, int kz, int kx) {
    const TileGeo g = make_geo<KY, TYT>(B, Z, X, Y, kz, kx);
    constexpr int MICRO = kStencilThreads;  // TX * TYT
    BwdPlan pl;
    pl.ncombos = kx * g.nchunks;
    static const int exp_mode = SN_ENV("SN_BWD_PLAN") ? atoi(SN_ENV("SN_BWD_PLAN")) : 0;  // experiments: 1 = 2 CTAs x 10 warps, single stage
    const int max_warps = exp_mode == 1 ? 10 : kBwdMaxThreads / 32;
    pl.grid_y = ceil_div(pl.ncombos, max_warps);
    pl.combos_per_cta = ceil_div(pl.ncombos, pl.grid_y);
    pl.Q = 1;
    while (pl.Q * 2 <= MICRO / 32 && pl.combos_per_cta * pl.Q * 2 <= max_warps) pl.Q *= 2;
    pl.threads = pl.combos_per_cta * pl.Q * 32;
    const size_t stage = (size_t)(((g.HZ * g.HX * g.WS + 31) & ~31) + kRZ * g.TX * g.TY) * 4;
    pl.nstage = (exp_mode != 1 && 2 * stage + 64 <= 227 * 1024) ? 2 : 1;
    pl.smem = pl.nstage * stage + 64;
    int gx = (exp_mode == 1 ? 2 * kNumSMs : kNumSMs) / pl.grid_y;
    gx = max(1, min(gx, g.ntiles));
    pl.grid_x = gx;
    pl.TP = (kz * kx * KY + 31) & ~31;
    return pl;
}

template <int KY, int TYT, int REM>
static int launch_bwd(BwdParams p, void* ws, int64_t ws_bytes, int* rows_out, cudaStream_t stream) {
    const BwdPlan pl = plan_bwd<KY, TYT>(p.B, p.Z, p.X, p.Y, p.kz, p.kx);
    if (pl.smem > 227 * 1024) return SN_ERR_UNSUPPORTED;
    if ((int64_t)pl.grid_x * pl.TP * 8 > ws_bytes) return SN_ERR_WORKSPACE;
    const TileGeo g = make_geo<KY, TYT>(p.B, p.Z, p.X, p.Y, p.kz, p.kx);
    const float* g0 = p.g0;
    p.partial = reinterpret_cast<double*>(ws);
    CUtensorMap tmap, gmap;
    const bool ok_x = make_grid_tmap(&tmap, p.x, p.B, p.Z, p.X, p.Y, g.HZ, g.HX, g.WS);
    const bool ok_g = make_grid_tmap(&gmap, g0, p.B, p.Z, p.X, p.Y, kRZ, g.TX, g.TY);
    p.use_tma = (ok_x && ok_g) ? 1 : 0;
    p.ncombos = pl.ncombos;
    p.combos_per_cta = pl.combos_per_cta;
    p.TP = pl.TP;
    p.Q = pl.Q;
    p.nstage = p.use_tma ? pl.nstage : 1;
    auto kern = stencil_bwd_kernel<KY, TYT, REM>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return cuda_rc(e);
    kern<<<dim3(pl.grid_x, pl.grid_y), pl.threads, pl.smem, stream>>>(p, tmap, gmap);
    SN_LAUNCH_CHECK();
    *rows_out = pl.grid_x;
    return SN_OK;
}

template <int KY, int TYT, int REM>
struct BwdRemDispatch {
    static int run(const BwdParams& p, void* ws, int64_t wsb, int* rows, cudaStream_t s) {
        if (p.kz % Geo<KY>::C == REM) return launch_bwd<KY, TYT, REM>(p, ws, wsb, rows, s);
        return BwdRemDispatch<KY, TYT, REM - 1>::run(p, ws, wsb, rows, s);
    }
};
template <int KY, int TYT>
struct BwdRemDispatch<KY, TYT, -1> {
    static int run(const BwdParams&, void*, int64_t, int*, cudaStream_t) { return SN_ERR_UNSUPPORTED; }
};

// launches the persistent tap-gradient kernel; *rows = number of partial rows it wrote, *TP their pitch
template <int KY>
int stencil_bwd_ky(const BwdParams& p, void* ws, int64_t wsb, int* rows, int* TP, cudaStream_t s) {
    *TP = (p.kz * p.kx * KY + 31) & ~31;
    return p.Y > 32 ? BwdRemDispatch<KY, 16, Geo<KY>::C - 1>::run(p, ws, wsb, rows, s)
                    : BwdRemDispatch<KY, 8, Geo<KY>::C - 1>::run(p, ws, wsb, rows, s);
}

template <int KY>
int64_t stencil_bwd_ws_ky(int B, int Z, int X, int Y, int kz, int kx) {
    const BwdPlan pl = Y > 32 ? plan_bwd<KY, 16>(B, Z, X, Y, kz, kx) : plan_bwd<KY, 8>(B, Z, X, Y, kz, kx);
    return (int64_t)pl.grid_x * pl.TP * 8;
}

}  // namespace sn
