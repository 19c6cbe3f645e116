// Shared pieces of the direct 3-D stencil kernels (forward observer + tap-gradient backward).
//
// Register blocking: every thread owns an RZ x 4 micro-tile (RZ voxels along z, 4 along the
// contiguous y axis).  The tap loop is split into runtime loops over dx and z-chunks and a
// fully unrolled inner body over (C z-taps) x (KY y-taps): one 16-byte shared-memory window
// load feeds up to C*KY*4 FFMAs, so the kernels are bound by the FP32 pipe, not by shared
// memory or HBM.
#pragma once
#include "common.cuh"

namespace sn {

constexpr int kRZ = 8;           // z extent of a micro-tile == z extent of a CTA tile
constexpr int kStencilThreads = 128;

// z-chunk (taps held in registers at once) per compile-time KY: C*KY <= 48 registers
__host__ __device__ constexpr int cmax_for(int ky) {
    return ky <= 5 ? 9 : (ky == 6 ? 8 : (ky == 7 ? 6 : (ky <= 9 ? 5 : (ky <= 11 ? 4 : 3))));
}
__host__ __device__ constexpr int round4(int v) { return (v + 3) & ~3; }

// TMA constraint measured on B200 (scratch experiments): with SWIZZLE_NONE the innermost
// start coordinate of a tiled load must be a multiple of 16 bytes (c0 = -2 floats raises an
// illegal-instruction fault, c0 = -4 is fine).  The halo box therefore starts PLA = round4(pad_left)
// columns left of the tile and the first OFF = PLA - pad_left floats of every window are dead.
template <int KY>
struct Geo {
    static constexpr int C = cmax_for(KY);
    static constexpr int CKP = round4(C * KY);  // taps of one (dx, chunk) slot, padded for 16-byte loads
    static constexpr int PL = (KY - 1) / 2;     // 'same' left pad along y
    static constexpr int PLA = round4(PL);      // left extent of the halo box (16-byte aligned start)
    static constexpr int OFF = PLA - PL;        // dead floats at the start of a thread's window
    static constexpr int WN = round4(OFF + KY + 3);  // window floats a thread reads per input row
};

struct FwdParams {
    const float* x;
    const float* Kstar;
    void* pred;
    int B, Z, X, Y, kz, kx;
    int out_f64, use_tma;
    int dbg;  // debugging/profiling switches (SN_FWD_DBG): 1 = no stores, 2 = no tanh, 4 = no TMA
    // z-split passes for kernels whose full halo would leave one 4-warp CTA per SM (13^3, 15^3): a pass applies
    // the z-taps [dz0, dz0 + kz) of the full kernel (Kstar already points at tap dz0; plz = full left pad - dz0).
    // pass_mode bit 0: add the partial sum the previous pass left in pred; bit 1: store the raw partial sum
    // (no relu(tanh)) for the next pass.  Partial sums live in pred's own element slots (float -> double is exact).
    int plz, pass_mode;
};

struct BwdParams {
    const float* x;
    const void* pred;
    const void* dpred;
    const float* g0;  // [B,1,Z,X,Y] float32 (workspace)
    double* partial;  // [gridDim.x][TP]
    int B, Z, X, Y, kz, kx;
    int pred_f64, dpred_f64, use_tma;
    int ncombos, combos_per_cta, TP;
    int Q, nstage;  // warps per tap group; TMA pipeline stages
    int dbg;        // profiling switch (SN_BWD_DBG): 4 = no TMA
    // device-side selection between this dense stencil and the occupancy-driven kernel (stencil_bwd_sparse.cu):
    // the dense kernel returns at once when nnz != NULL and *nnz <= nnz_max
    const unsigned long long* nnz;
    unsigned long long nnz_max;
};

// G0 = dL/ds = dpred * (1 - pred^2) * [pred > 0], evaluated in float64 and rounded once: the parameter
// gradients are ill-conditioned sums of G0 (zero-sum projection), so float32 rounding inside this
// product (dpred, tanh, 1-p^2) costs ~1e-5 relative on the worst gradient (scratch/precision_probe.py)
__device__ __forceinline__ float g0_of(double p, double d) { return p > 0.0 ? (float)(d * (1.0 - p * p)) : 0.f; }

// tile geometry shared by host and device
struct TileGeo {
    int TY, TX;          // CTA tile extent in y, x (z extent is kRZ)
    int HZ, HX, WS;      // halo tile extents (rows, rows, padded row length in floats)
    int tiles_z, tiles_x, tiles_y, ntiles;
    int nchunks;         // ceil(kz / C)
    int plz, plx, ply;   // left extents of the halo box (ply is rounded up to 4 floats, see Geo)
};

template <int KY, int TYT>
__host__ __device__ inline TileGeo make_geo(int B, int Z, int X, int Y, int kz, int kx, int plz = -1000) {
    TileGeo g;
    g.TY = TYT * 4;
    g.TX = kStencilThreads / TYT;
    g.HZ = kRZ + kz - 1;
    g.HX = g.TX + kx - 1;
    g.WS = round4(g.TY + Geo<KY>::OFF + KY - 1);
    g.tiles_z = ceil_div(Z, kRZ);
    g.tiles_x = ceil_div(X, g.TX);
    g.tiles_y = ceil_div(Y, g.TY);
    g.ntiles = B * g.tiles_z * g.tiles_x * g.tiles_y;
    g.nchunks = ceil_div(kz, Geo<KY>::C);
    g.plz = plz == -1000 ? pad_left(kz) : plz;
    g.plx = pad_left(kx);
    g.ply = Geo<KY>::PLA;  // box start (aligned), not the conv pad
    return g;
}

// plain-load fallback for the halo tile (odd Y, unaligned base, no driver entry point)
__device__ __forceinline__ void load_halo_plain(float* __restrict__ sx, const float* __restrict__ x, const TileGeo& g,
                                                int Z, int X, int Y, int b, int z0, int x0, int y0, int nthreads) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = nthreads >> 5;
    for (int r = warp; r < g.HZ * g.HX; r += nwarps) {
        const int gz = z0 - g.plz + r / g.HX, gx = x0 - g.plx + r % g.HX;
        const bool row_ok = gz >= 0 && gz < Z && gx >= 0 && gx < X;
        const float* src = x + (((size_t)b * Z + (row_ok ? gz : 0)) * X + (row_ok ? gx : 0)) * Y;
        for (int c = lane; c < g.WS; c += 32) {
            const int gy = y0 - g.ply + c;
            sx[r * g.WS + c] = (row_ok && gy >= 0 && gy < Y) ? __ldg(src + gy) : 0.f;
        }
    }
}

__device__ __forceinline__ void decode_tile(int tile, const TileGeo& g, int& b, int& z0, int& x0, int& y0) {
    const int ty = tile % g.tiles_y;
    tile /= g.tiles_y;
    const int tx = tile % g.tiles_x;
    tile /= g.tiles_x;
    const int tz = tile % g.tiles_z;
    b = tile / g.tiles_z;
    z0 = tz * kRZ;
    x0 = tx * g.TX;
    y0 = ty * g.TY;
}

}  // namespace sn
