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
// tanhf, not a float64 tanh: inside this FFMA-bound kernel every float64 variant tried cost +25 % .. +38 % of the
// kernel (FP64 issue is scarce on B200; see tanh_pos_f64 in stencil_common.cuh, which the occupancy-driven kernel uses).
static __device__ __noinline__ void store_row(float a0, float a1, float a2, float a3, void* pred, size_t idx, int out_f64, int ny,
                                              bool vec, int pass_mode) {
    float o[4] = {a0, a1, a2, a3};
    if (pass_mode & 1) {  // z-split: add the previous passes' partial sum (stored in this row's own slots)
        if (out_f64) {
            const double* in = reinterpret_cast<const double*>(pred) + idx;
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (r < ny) o[r] = (float)in[r] + o[r];
        } else {
            const float* in = reinterpret_cast<const float*>(pred) + idx;
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (r < ny) o[r] = in[r] + o[r];
        }
    }
    if (!(pass_mode & 2)) {
#pragma unroll
        for (int r = 0; r < 4; ++r) o[r] = o[r] > 0.f ? tanhf(o[r]) : 0.f;
    }
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

// Epilogue of the COMPENSATED kernels (T > kCompTaps taps, BASELINE config 4's 11^3 .. 15^3): the sum arrives as an
// unevaluated float32 pair (hi, lo).  float64 predictions: the pair and the previous z-split passes' partial sums are
// combined in float64 (the partial sums live in pred's own float64 slots) and tanh is evaluated in float64 — next to
// >= 1000 FFMAs per voxel the float64 work is noise; float32 predictions: hi + lo rounded once, tanhf.
static __device__ __noinline__ void store_row_comp(const float (&hi)[4], const float (&lo)[4], void* pred, size_t idx, int out_f64,
                                                   int ny, bool vec, int pass_mode) {
    if (out_f64) {
        double* out = reinterpret_cast<double*>(pred) + idx;
        double o[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            o[r] = (double)hi[r] + (double)lo[r];
            if ((pass_mode & 1) && r < ny) o[r] += out[r];
            if (!(pass_mode & 2)) o[r] = tanh_pos_f64(o[r]);  // relu inside (negative sums clamp to 0)
        }
        if (vec) {
            reinterpret_cast<double2*>(out)[0] = make_double2(o[0], o[1]);
            reinterpret_cast<double2*>(out)[1] = make_double2(o[2], o[3]);
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (r < ny) out[r] = o[r];
        }
    } else {
        float* out = reinterpret_cast<float*>(pred) + idx;
        float o[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            o[r] = hi[r] + lo[r];
            if ((pass_mode & 1) && r < ny) o[r] += out[r];
            if (!(pass_mode & 2)) o[r] = o[r] > 0.f ? tanhf(o[r]) : 0.f;
        }
        if (vec) {
            *reinterpret_cast<float4*>(out) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (r < ny) out[r] = o[r];
        }
    }
}

// kernels with more taps than this accumulate compensated (COMP): float32 running sums over thousands of taps cost
// 1.2e-5 .. 1.4e-5 on the worst parameter gradient at 13^3 / 15^3 (CPU emulation scratch/precision_probe3.py: the exact
// sum rounded once to float32 gives 2e-6 .. 5e-6), above north_star's 1e-5 bar
constexpr int kCompTaps = 1000;

// Persistent CTAs (4 per SM), each walking tiles blockIdx.x, blockIdx.x + gridDim.x, ...; a CTA's halo
// load (one TMA box) is covered by the other CTAs of the SM.  (The first version launched one CTA per tile: all CTAs of
// a wave waited for their 55 KB halo at the same time — profiles/r1_notes.md — and the FMA pipe idled ~45 %; a two-stage
// pipeline with 2 CTAs per SM measured 4.5 % slower than this single stage with 4.)
// COMP: every dx step's partial sum (kz * KY taps, a handful of them non-zero) is added to an unevaluated (hi, lo)
// float32 pair by an error-free TwoSum: 6 FADDs per accumulator per dx step against kz * KY FFMAs (3.5 % at 13^3).
template <int KY, int TYT, int REM, bool COMP>
__global__ void __launch_bounds__(kStencilThreads, COMP ? 2 : 4)
stencil_fwd_kernel(const FwdParams p, const __grid_constant__ CUtensorMap tmap) {
    if (p.nnz && fwd_sparse_selected(p.nnz, p.nnz_max, p.dw_max)) return;  // sparse input: the occupancy-driven kernel does the work
    constexpr int C = Geo<KY>::C, CKP = Geo<KY>::CKP;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TileGeo g = make_geo<KY, TYT>(p.B, p.Z, p.X, p.Y, p.kz, p.kx, p.plz);
    const int halo_floats = g.HZ * g.HX * g.WS;
    const int halo_stride = (halo_floats + 31) & ~31;
    float* sx0 = reinterpret_cast<float*>(smem_raw);
    float* sk = sx0 + halo_stride;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sk + ((p.kx * g.nchunks * CKP + 31) & ~31));
    const int tid = threadIdx.x;
    const int G = gridDim.x;

    auto issue = [&](int tile) {  // thread 0 only
        int b, z0, x0, y0;
        decode_tile(tile, g, b, z0, x0, y0);
        mbar_arrive_expect_tx(&bar[0], (uint32_t)halo_floats * 4u);
        tma_load_4d(sx0, &tmap, &bar[0], y0 - g.ply, x0 - g.plx, z0 - g.plz, b);
    };

    // tile-list pass: the tiles the occupancy-driven forward left for this stencil (count written before this launch)
    const int ntl = p.tile_list ? (int)*reinterpret_cast<volatile unsigned long long*>(p.state + 3) : g.ntiles;
    if (p.use_tma && tid == 0) {
        mbar_init(&bar[0], 1);
        fence_barrier_init();
        if ((int)blockIdx.x < ntl) issue(p.tile_list ? __ldg(p.tile_list + blockIdx.x) : (int)blockIdx.x);
    }
    // taps -> shared once per CTA, re-laid out as [dx][chunk][dzl*KY + dy] (zero padded to CKP)
    for (int i = tid; i < p.kx * g.nchunks * CKP; i += kStencilThreads) sk[i] = 0.f;
    __syncthreads();
    const int T = p.kz * p.kx * KY;
    for (int t = tid; t < T; t += kStencilThreads) {
        const int dy = t % KY, dx = (t / KY) % p.kx, dz = t / (KY * p.kx);
        sk[(dx * g.nchunks + dz / C) * CKP + (dz % C) * KY + dy] = __ldg(p.Kstar + t);
    }
    __syncthreads();

    const int tyi = tid % TYT, txi = tid / TYT;
    const int zstride = g.HX * g.WS;
    const bool vec = ((p.Y & 3) == 0);
    int k = 0;
    for (int ti = blockIdx.x; ti < ntl; ti += G, ++k) {
        const int tile = p.tile_list ? __ldg(p.tile_list + ti) : ti;
        int b, z0, x0, y0;
        decode_tile(tile, g, b, z0, x0, y0);
        const float* sx = sx0;
        if (p.use_tma) {
            mbar_wait(&bar[0], (uint32_t)k & 1u);
        } else {
            __syncthreads();  // previous tile fully consumed
            load_halo_plain(sx0, p.x, g, p.Z, p.X, p.Y, b, z0, x0, y0, kStencilThreads);
            __syncthreads();
        }

        float acc[kRZ][4];
        float hi[COMP ? kRZ : 1][4], lo[COMP ? kRZ : 1][4];
#pragma unroll
        for (int i = 0; i < kRZ; ++i)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                acc[i][r] = 0.f;
                if constexpr (COMP) { hi[i][r] = 0.f; lo[i][r] = 0.f; }
            }

        // Empty halo -> the tile's sums are exactly zero: skip the tap loop.  Point-cloud grids are clustered (in the
        // reference's data-sample/sample_575.npy 41 of the 64 tiles of the 64^3 grid hold no occupied voxel), and those are
        // the grids this dense stencil is selected for (fwd_sparse_selected).  27 16-byte shared loads per thread against
        // ~2700 in the tap loop; empty grids 88 -> 26 us, rolled copies of sample_575 92 -> 75 us, uniform grids +3 %.
        // (Skipping per halo ROW inside the tap loop as well — a warp vote on per-row flags — gained 4 us more on the
        // sample_575 grids and cost 8 % on uniform ones: not kept.)
        bool any = false;
        {
            const float4* h4 = reinterpret_cast<const float4*>(sx);
            for (int i = tid; i < (halo_floats >> 2); i += kStencilThreads) {
                const float4 v = h4[i];
                any |= (v.x != 0.f) | (v.y != 0.f) | (v.z != 0.f) | (v.w != 0.f);
            }
        }
        if (__syncthreads_or(any ? 1 : 0)) {
            // kz = nfull*C + REM with REM a template parameter: the hot loop holds exactly the bodies this
            // kernel size needs (a runtime switch over all remainders inflated the code and cost ~7 %)
            const int nfull = p.kz / C;
            for (int dx = 0; dx < p.kx; ++dx) {
                const float* sxrow = sx + (txi + dx) * g.WS + 4 * tyi;
                const float* skrow = sk + dx * g.nchunks * CKP;
                for (int ch = 0; ch < nfull; ++ch) fwd_chunk<KY, C>(acc, sxrow + (ch * C) * zstride, zstride, skrow + ch * CKP);
                if constexpr (REM > 0) fwd_chunk<KY, REM>(acc, sxrow + (nfull * C) * zstride, zstride, skrow + nfull * CKP);
                if constexpr (COMP) {
#pragma unroll
                    for (int i = 0; i < kRZ; ++i)
#pragma unroll
                        for (int r = 0; r < 4; ++r) {  // (hi, lo) += acc, error-free (Knuth TwoSum)
                            const float a = hi[i][r], c = acc[i][r];
                            const float s = __fadd_rn(a, c);
                            const float bb = __fsub_rn(s, a);
                            const float e = __fadd_rn(__fsub_rn(a, __fsub_rn(s, bb)), __fsub_rn(c, bb));
                            hi[i][r] = s;
                            lo[i][r] = __fadd_rn(lo[i][r], e);
                            acc[i][r] = 0.f;
                        }
                }
            }
        }

        if (p.use_tma) {
            __syncthreads();  // every thread is done reading the buffer -> refill it with the CTA's next tile
            if (tid == 0 && ti + G < ntl) {
                fence_proxy_async();
                issue(p.tile_list ? __ldg(p.tile_list + ti + G) : ti + G);
            }
        }

        // epilogue: relu(tanh(s)) and store in the caller's dtype (overlaps the TMA just issued)
        const int gx = x0 + txi, gy = y0 + 4 * tyi;
        if (gx < p.X && gy < p.Y) {
#pragma unroll
            for (int zo = 0; zo < kRZ; ++zo) {
                const int gz = z0 + zo;
                if (gz < p.Z) {
                    const size_t idx = (((size_t)b * p.Z + gz) * p.X + gx) * p.Y + gy;
                    if constexpr (COMP)
                        store_row_comp(hi[zo], lo[zo], p.pred, idx, p.out_f64, p.Y - gy, vec, p.pass_mode);
                    else
                        store_row(acc[zo][0], acc[zo][1], acc[zo][2], acc[zo][3], p.pred, idx, p.out_f64, p.Y - gy, vec, p.pass_mode);
                }
            }
        }
    }
    if (p.tile_list && p.last_pass) {
        // the last CTA of the last pass leaves the hand-off counters at zero for the next forward on this state buffer
        // (every CTA has read state[3] before it gets here)
        __syncthreads();
        if (tid == 0 && atomicAdd(p.state + 6, 1ull) == (unsigned long long)(G - 1)) {
            atomicExch(p.state + 3, 0ull);
            atomicExch(p.state + 6, 0ull);
        }
    }
}

template <int KY, int TYT, int REM, bool COMP>
static int launch_fwd(const FwdParams& p0, cudaStream_t stream) {
    FwdParams p = p0;
    const TileGeo g = make_geo<KY, TYT>(p.B, p.Z, p.X, p.Y, p.kz, p.kx, p.plz);
    const int halo_stride = (g.HZ * g.HX * g.WS + 31) & ~31;
    const int tap_floats = (p.kx * g.nchunks * Geo<KY>::CKP + 31) & ~31;
    CUtensorMap tmap;
    p.use_tma = make_grid_tmap(&tmap, p.x, p.B, p.Z, p.X, p.Y, g.HZ, g.HX, g.WS) ? 1 : 0;
    const size_t smem = (size_t)(halo_stride + tap_floats) * 4 + 32;
    if (smem > 227 * 1024) return SN_ERR_UNSUPPORTED;
    auto kern = stencil_fwd_kernel<KY, TYT, REM, COMP>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_rc(e);
    const int per_sm = max(1, min(COMP ? 2 : 4, (int)((227 * 1024) / (smem + 1024))));
    // (tile-list pass: the number of listed tiles is only known on the device; a full grid of CTAs, most of which find
    // nothing to do when few tiles were handed over)
    const int grid = max(1, min(g.ntiles, kNumSMs * per_sm));
    kern<<<grid, kStencilThreads, smem, stream>>>(p, tmap);
    SN_LAUNCH_CHECK();
    return SN_OK;
}

// kz % C selects the instantiation
template <int KY, int TYT, int REM>
struct FwdRemDispatch {
    static int run(const FwdParams& p, cudaStream_t s) {
        if (p.kz % Geo<KY>::C == REM) {
            // compensated accumulation for kernels beyond kCompTaps taps; instantiated for the widths such kernels
            // come in (ky >= 9: config 4's 11^3 / 13^3 / 15^3 and e.g. (13,9,9))
            if constexpr (KY >= 9) {
                if (p.comp) return launch_fwd<KY, TYT, REM, true>(p, s);
            }
            return launch_fwd<KY, TYT, REM, false>(p, s);
        }
        return FwdRemDispatch<KY, TYT, REM - 1>::run(p, s);
    }
};
template <int KY, int TYT>
struct FwdRemDispatch<KY, TYT, -1> {
    static int run(const FwdParams&, cudaStream_t) { return SN_ERR_UNSUPPORTED; }
};

// shared memory of one single-stage CTA for a pass with kz z-taps
template <int KY, int TYT>
static size_t fwd_pass_smem(const FwdParams& p, int kz) {
    const TileGeo g = make_geo<KY, TYT>(p.B, p.Z, p.X, p.Y, kz, p.kx, 0);
    const int halo_stride = (g.HZ * g.HX * g.WS + 31) & ~31;
    const int tap_floats = (p.kx * g.nchunks * Geo<KY>::CKP + 31) & ~31;
    return (size_t)(halo_stride + tap_floats) * 4 + 32;
}

template <int KY, int TYT>
static int fwd_passes(const FwdParams& p0, cudaStream_t s) {
    // z-split: as few passes as leave at least two CTAs (8 warps) per SM.  Measured on 128^3 grids before the
    // split: 13^3 / 15^3 kernels ran one 4-warp CTA per SM at 26 % / 31 % of the FFMA peak, 9^3 / 11^3 (two
    // CTAs) at 58 % / 69 % (profiles/r1_notes.md)
    constexpr size_t kTwoPerSm = (227 * 1024 - 2048) / 2;
    int npass = 1;
    static const int forced = SN_ENV("SN_FWD_PASSES") ? atoi(SN_ENV("SN_FWD_PASSES")) : 0;
    if (forced > 0)
        npass = forced < p0.kz ? forced : p0.kz;
    else
        while (npass < p0.kz && fwd_pass_smem<KY, TYT>(p0, ceil_div(p0.kz, npass)) > kTwoPerSm) ++npass;
    const int kzp = ceil_div(p0.kz, npass);
    npass = ceil_div(p0.kz, kzp);
    for (int i = 0; i < npass; ++i) {
        FwdParams p = p0;
        p.comp = p0.kz * p0.kx * KY > kCompTaps ? 1 : 0;
        const int dz0 = i * kzp;
        p.kz = (p0.kz - dz0) < kzp ? (p0.kz - dz0) : kzp;
        p.Kstar = p0.Kstar + (size_t)dz0 * p0.kx * KY;
        p.plz = pad_left(p0.kz) - dz0;
        p.pass_mode = (i > 0 ? 1 : 0) | (i < npass - 1 ? 2 : 0);
        p.last_pass = (i == npass - 1 && p0.last_pass) ? 1 : 0;
        const int rc = FwdRemDispatch<KY, TYT, Geo<KY>::C - 1>::run(p, s);
        if (rc) return rc;
    }
    return SN_OK;
}

template <int KY>
int stencil_fwd_ky(const FwdParams& p, cudaStream_t s) {
    return p.Y > 32 ? fwd_passes<KY, 16>(p, s) : fwd_passes<KY, 8>(p, s);
}

}  // namespace sn
