// Implementation of the forward stencil, included by the per-KY translation units
// (stencil_fwd_ky*.cu) so that the 8 kernel widths compile in parallel.
#pragma once
#include <stdlib.h>
#include "stencil_common.cuh"
#include "tma_host.cuh"

namespace sn {

template <int KY, int CS>
__device__ __forceinline__ void fwd_chunk(float (&acc)[kRZ][4], const float* __restrict__ sxp, int zstride,
                                          const float* __restrict__ skp) {
    constexpr int WN = Geo<KY>::WN;
    constexpr int NT = round4(CS * KY);
    float tap[NT];
#pragma unroll
    for (int i = 0; i < NT / 4; ++i) {
        const float4 v = reinterpret_cast<const float4*>(skp)[i];
        tap[4 * i] = v.x; tap[4 * i + 1] = v.y; tap[4 * i + 2] = v.z; tap[4 * i + 3] = v.w;
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
                    for (int r = 0; r < 4; ++r) acc[zo][r] = fmaf(win[Geo<KY>::OFF + r + dy], tap[dzl * KY + dy], acc[zo][r]);
                }
            }
        }
    }
}

// relu(tanh(s)) of one 4-voxel row segment + store.  Deliberately NOT inlined: 32 inlined tanhf bodies
// were 13 KB of straight-line code next to the 23 KB tap loop and pushed the hot path out of the 32 KB
// instruction cache level (profiles/r1_notes.md).
// tanhf, not a float64 tanh: the double version cost ~27 % of the kernel's instructions and buys nothing
// measurable (scratch/precision_probe.py: G0's float32 rounding in the backward dominates).
static __device__ __noinline__ void store_row(float a0, float a1, float a2, float a3, void* pred, size_t idx, int out_f64, int ny,
                                       bool vec, int dbg) {
    float o[4] = {a0, a1, a2, a3};
    if (!(dbg & 2)) {
#pragma unroll
        for (int r = 0; r < 4; ++r) o[r] = o[r] > 0.f ? tanhf(o[r]) : 0.f;
    }
    if ((dbg & 1) && o[0] != 123.456f) return;
    if (out_f64) {
        double* out = reinterpret_cast<double*>(pred) + idx;
        if (vec) {
            reinterpret_cast<double2*>(out)[0] = make_double2((double)o[0], (double)o[1]);
            reinterpret_cast<double2*>(out)[1] = make_double2((double)o[2], (double)o[3]);
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (r < ny) out[r] = (double)o[r];
        }
    } else {
        float* out = reinterpret_cast<float*>(pred) + idx;
        if (vec) {
            *reinterpret_cast<float4*>(out) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (r < ny) out[r] = o[r];
        }
    }
}

// One persistent CTA per SM: 8 compute warps (thread = 8 x 4 register block of outputs, tile = 8 x TX x TY
// with TX*TY/4 = 256 micro-tiles) + 1 producer warp that keeps a two-stage TMA pipeline full through
// full/empty mbarriers.  No block-wide barrier in the tile loop: a warp releases the stage as soon as its
// taps are done and applies relu(tanh) to the finished tile WHILE it computes the next one (two pending rows
// after every dx iteration), so the low-IPC epilogue never runs on its own.
// History (profiles/r1_notes.md): one CTA per tile -> all CTAs of a wave waited for their halo together;
// 2 persistent CTAs per SM with a __syncthreads per tile -> the tanhf epilogue (10-15 us of 95) ran with the
// FMA pipe idle.
constexpr int kFwdThreads = kFwdMicro + 32;

template <int KY, int TYT, int REM>
__global__ void __launch_bounds__(kFwdThreads, 1)
stencil_fwd_kernel(const FwdParams p, const __grid_constant__ CUtensorMap tmap) {
    constexpr int C = Geo<KY>::C, CKP = Geo<KY>::CKP;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TileGeo g = make_geo<KY, TYT, kFwdMicro>(p.B, p.Z, p.X, p.Y, p.kz, p.kx);
    const int halo_floats = g.HZ * g.HX * g.WS;
    const int halo_stride = (halo_floats + 31) & ~31;
    const int nstage = p.nstage;
    float* sx0 = reinterpret_cast<float*>(smem_raw);
    float* sk = sx0 + nstage * halo_stride;
    uint64_t* full = reinterpret_cast<uint64_t*>(sk + ((p.kx * g.nchunks * CKP + 31) & ~31));  // [2]
    uint64_t* empty = full + 2;                                                                 // [2]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NW = kFwdMicro / 32;
    const int G = gridDim.x;

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_init(&empty[0], NW);
        mbar_init(&empty[1], NW);
        fence_barrier_init();
    }
    // taps -> shared once per CTA, re-laid out as [dx][chunk][dzl*KY + dy] (zero padded to CKP)
    for (int i = tid; i < p.kx * g.nchunks * CKP; i += kFwdThreads) sk[i] = 0.f;
    __syncthreads();
    const int T = p.kz * p.kx * KY;
    for (int t = tid; t < T; t += kFwdThreads) {
        const int dy = t % KY, dx = (t / KY) % p.kx, dz = t / (KY * p.kx);
        sk[(dx * g.nchunks + dz / C) * CKP + (dz % C) * KY + dy] = __ldg(p.Kstar + t);
    }
    __syncthreads();

    if (p.use_tma && warp == NW) {
        // ---- producer
        if (lane == 0) {
            int k = 0;
            for (int tile = blockIdx.x; tile < g.ntiles; tile += G, ++k) {
                const int buf = k % nstage, use = k / nstage;
                if (use > 0) mbar_wait(&empty[buf], (uint32_t)(use - 1) & 1u);
                int b, z0, x0, y0;
                decode_tile(tile, g, b, z0, x0, y0);
                mbar_arrive_expect_tx(&full[buf], (uint32_t)halo_floats * 4u);
                tma_load_4d(sx0 + buf * halo_stride, &tmap, &full[buf], y0 - g.ply, x0 - g.plx, z0 - g.plz, b);
            }
        }
        return;
    }

    // ---- compute warps (the producer warp only gets here without TMA, as a loader)
    const bool compute = warp < NW;
    const int tyi = (tid % kFwdMicro) % TYT, txi = (tid % kFwdMicro) / TYT;
    const int zstride = g.HX * g.WS;
    const bool vec = ((p.Y & 3) == 0);
    const int nfull = p.kz / C;
    const int rows_per_dx = (kRZ + p.kx - 1) / p.kx;

    float pend[kRZ][4];      // finished-but-not-yet-stored outputs of the previous tile
    int npend = 0, pz = 0, pny = 0;   // rows still pending; z of pend[0]; valid y of the row segment
    size_t pidx = 0;         // element index of pend[0]
    bool pok = false;
#pragma unroll
    for (int i = 0; i < kRZ; ++i)
#pragma unroll
        for (int r = 0; r < 4; ++r) pend[i][r] = 0.f;

    auto flush_rows = [&](int n) {
        for (int j = 0; j < n && npend > 0; ++j) {
            if (pok && pz < p.Z) store_row(pend[0][0], pend[0][1], pend[0][2], pend[0][3], p.pred, pidx, p.out_f64, pny, vec, p.dbg);
#pragma unroll
            for (int i = 0; i + 1 < kRZ; ++i)
#pragma unroll
                for (int r = 0; r < 4; ++r) pend[i][r] = pend[i + 1][r];
            --npend;
            ++pz;
            pidx += (size_t)p.X * p.Y;
        }
    };

    int k = 0;
    for (int tile = blockIdx.x; tile < g.ntiles; tile += G, ++k) {
        int b, z0, x0, y0;
        decode_tile(tile, g, b, z0, x0, y0);
        const int buf = p.use_tma ? k % nstage : 0;
        const float* sx = sx0 + buf * halo_stride;
        if (p.use_tma) {
            if (!(p.dbg & 4)) mbar_wait(&full[buf], (uint32_t)(k / nstage) & 1u);
        } else {
            __syncthreads();  // previous tile fully consumed
            load_halo_plain(sx0, p.x, g, p.Z, p.X, p.Y, b, z0, x0, y0, kFwdThreads);
            __syncthreads();
        }
        if (compute) {
            float acc[kRZ][4];
#pragma unroll
            for (int i = 0; i < kRZ; ++i)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[i][r] = 0.f;

            // kz = nfull*C + REM with REM a template parameter: the hot loop holds exactly the bodies this kernel
            // size needs (a runtime switch over all remainders inflated the code and cost ~7 %)
            for (int dx = 0; dx < p.kx; ++dx) {
                const float* sxrow = sx + (txi + dx) * g.WS + 4 * tyi;
                const float* skrow = sk + dx * g.nchunks * CKP;
                for (int ch = 0; ch < nfull; ++ch) fwd_chunk<KY, C>(acc, sxrow + (ch * C) * zstride, zstride, skrow + ch * CKP);
                if constexpr (REM > 0) fwd_chunk<KY, REM>(acc, sxrow + (nfull * C) * zstride, zstride, skrow + nfull * CKP);
                flush_rows(rows_per_dx);  // a slice of the previous tile's epilogue, hidden behind this tile's FFMAs
            }
            if (p.use_tma) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[buf]);  // this warp no longer reads the stage
            }
            flush_rows(kRZ);  // whatever is left of the previous tile (kx < 8/rows_per_dx cannot happen, but be safe)
            // this tile becomes the pending one
#pragma unroll
            for (int i = 0; i < kRZ; ++i)
#pragma unroll
                for (int r = 0; r < 4; ++r) pend[i][r] = acc[i][r];
            const int gx = x0 + txi, gy = y0 + 4 * tyi;
            pok = gx < p.X && gy < p.Y;
            npend = kRZ;
            pz = z0;
            pny = p.Y - gy;
            pidx = (((size_t)b * p.Z + z0) * p.X + (pok ? gx : 0)) * p.Y + (pok ? gy : 0);
        }
    }
    if (compute) flush_rows(kRZ);
}

template <int KY, int TYT, int REM>
static int launch_fwd(const FwdParams& p0, cudaStream_t stream) {
    FwdParams p = p0;
    {
        const char* e = getenv("SN_FWD_DBG");
        p.dbg = e ? atoi(e) : 0;
    }
    const TileGeo g = make_geo<KY, TYT, kFwdMicro>(p.B, p.Z, p.X, p.Y, p.kz, p.kx);
    const int halo_stride = (g.HZ * g.HX * g.WS + 31) & ~31;
    const int tap_floats = (p.kx * g.nchunks * Geo<KY>::CKP + 31) & ~31;
    CUtensorMap tmap;
    p.use_tma = make_grid_tmap(&tmap, p.x, p.B, p.Z, p.X, p.Y, g.HZ, g.HX, g.WS) ? 1 : 0;
    p.nstage = 2;
    size_t smem = (size_t)(2 * halo_stride + tap_floats) * 4 + 64;
    if (!p.use_tma || smem > 227 * 1024) {  // huge halo or no TMA: single stage
        p.nstage = 1;
        smem = (size_t)(halo_stride + tap_floats) * 4 + 64;
    }
    if (smem > 227 * 1024) return SN_ERR_UNSUPPORTED;
    auto kern = stencil_fwd_kernel<KY, TYT, REM>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_rc(e);
    const int grid = max(1, min(g.ntiles, kNumSMs));
    kern<<<grid, kFwdThreads, smem, stream>>>(p, tmap);
    SN_LAUNCH_CHECK();
    return SN_OK;
}

// kz % C selects the instantiation
template <int KY, int TYT, int REM>
struct FwdRemDispatch {
    static int run(const FwdParams& p, cudaStream_t s) {
        if (p.kz % Geo<KY>::C == REM) return launch_fwd<KY, TYT, REM>(p, s);
        return FwdRemDispatch<KY, TYT, REM - 1>::run(p, s);
    }
};
template <int KY, int TYT>
struct FwdRemDispatch<KY, TYT, -1> {
    static int run(const FwdParams&, cudaStream_t) { return SN_ERR_UNSUPPORTED; }
};

template <int KY>
int stencil_fwd_ky(const FwdParams& p, cudaStream_t s) {
    return p.Y > 32 ? FwdRemDispatch<KY, 16, Geo<KY>::C - 1>::run(p, s) : FwdRemDispatch<KY, 8, Geo<KY>::C - 1>::run(p, s);
}

}  // namespace sn
