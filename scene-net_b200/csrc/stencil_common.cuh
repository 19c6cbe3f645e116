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

// Device-side selection of the forward kernel from sn_grid_prepare's state buffer: the occupancy-driven kernel works iff
// the grid is sparse (st[0] = non-zero voxels) AND not clustered (st[2] = mask words with >= 8 of 32 voxels occupied).
// Its cost per tile grows with the tile's own occupancy and tiles are assigned statically, so ONE locally dense tile
// (the ground layer of a LiDAR scan) makes it slower than the dense stencil even when the grid as a whole is almost
// empty: KITTI-shaped scans at 2.8 % overall occupancy with one layer at 35 % ran 294 us against 87 us dense.
__device__ __forceinline__ bool fwd_sparse_selected(const unsigned long long* st, unsigned long long nnz_max, unsigned long long dw_max) {
    return st[0] <= nnz_max && st[2] <= dw_max;
}

struct FwdParams {
    const float* x;
    const float* Kstar;
    void* pred;
    int B, Z, X, Y, kz, kx;
    int out_f64, use_tma;
    // z-split passes for kernels whose full halo would leave one 4-warp CTA per SM (13^3, 15^3): a pass applies
    // the z-taps [dz0, dz0 + kz) of the full kernel (Kstar already points at tap dz0; plz = full left pad - dz0).
    // pass_mode bit 0: add the partial sum the previous pass left in pred; bit 1: store the raw partial sum
    // (no relu(tanh)) for the next pass.  Partial sums live in pred's own element slots (float -> double is exact).
    int plz, pass_mode;
    int comp;  // compensated accumulation (set by the launcher for kernels with more than kCompTaps taps)
    // ABI v4 — tile-list pass: the occupancy-driven forward ran first and left the tiles above its break-even occupancy
    // in tile_list (state[3] of them); this stencil computes exactly those.  The last CTA to finish the last z-split
    // pass zeroes state[3] and state[6] again.
    const int* tile_list;
    unsigned long long* state;
    int last_pass;
    // device-side selection against the occupancy-driven kernel (stencil_fwd_sparse.cu): this dense stencil returns
    // at once when nnz != NULL and *nnz <= nnz_max
    const unsigned long long* nnz;
    unsigned long long nnz_max;
    unsigned long long dw_max;  // clustering bound on st[2] (fwd_sparse_selected)
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
    // device-side selection between this dense stencil and the occupancy-driven kernel (stencil_bwd_sparse.cu):
    // the dense kernel returns at once when nnz != NULL and *nnz <= nnz_max
    const unsigned long long* nnz;
    unsigned long long nnz_max;
    double* W;                   // [T]: written by the last CTA when `ticket` is given
    unsigned long long* ticket;  // device counter zeroed by sn_grid_prepare (NULL: a separate kernel sums the rows)
};

// G0 = dL/ds = dpred * (1 - pred^2) * [pred > 0], evaluated in float64 and rounded once: the parameter
// gradients are ill-conditioned sums of G0 (zero-sum projection), so float32 rounding inside this
// product (dpred, tanh, 1-p^2) costs ~1e-5 relative on the worst gradient (scratch/precision_probe.py)
__device__ __forceinline__ float g0_of(double p, double d) { return p > 0.0 ? (float)(d * (1.0 - p * p)) : 0.f; }

// tanh(s) for s > 0 in float64 (~1e-11 relative), used by the occupancy-driven forward for float64 predictions.
// Measured on the (7,7,7) reference fixture: the worst parameter gradient is 9.6e-6 of the reference with tanhf
// in the dense stencil, 1.05e-5 with tanhf in the occupancy-driven kernel (another float32 summation order) and
// 4.4e-6 with this evaluation — tanhf's 1-2 ulp error in pred is the largest single term of the error budget.
// FP64 issue is scarce on B200: inside the FFMA-bound dense stencil this function costs +25 % (libm tanh(double):
// +27 %; a variant with fewer FP64 operations but more selects / conversions: +38 %), so the dense stencil keeps
// tanhf; the occupancy-driven kernel is not FP32-bound and pays +16 us at config 2 (profiles/r1_notes.md).
// tanh(s) = (1 - t) / (1 + t) with t = exp(-2s) = 2^k * P(r): degree-10 Taylor on |r| <= ln2/2 (2e-13), the quotient
// from the hardware reciprocal seed + two Newton steps (the first version, exp(2s) by a degree-11 polynomial and an
// IEEE division, was 66 instructions per call and 40 % of all instructions of the occupancy-driven forward at
// config 2 — ncu, profiles/r1_notes.md; this one is ~30).  Absolute error < 1e-12.
// (coefficients in constant memory: as literals every one of them cost two uniform-register moves in front of its DFMA,
// 22 of the function's 53 instructions)
__constant__ double kTanhC[14] = {
    1.4426950408889634,          // [0] log2(e)
    -6.93147180369123816490e-01, // [1] -ln2 hi
    -1.90821492927058770002e-10, // [2] -ln2 lo
    2.7557319223985893e-07,      // [3] 1/10!
    2.7557319223985888e-06,      // [4] 1/9!
    2.4801587301587302e-05,      // [5] 1/8!
    1.9841269841269841e-04,      // [6] 1/7!
    1.3888888888888889e-03,      // [7] 1/6!
    8.3333333333333332e-03,      // [8] 1/5!
    4.1666666666666664e-02,      // [9] 1/4!
    1.6666666666666666e-01,      // [10] 1/3!
    0.5, 1.0, 20.0};
// Branch-free (s is clamped to [0, 20]: tanh(0) comes out as exactly 0, tanh(20) as 1 to the last bit), so the four calls
// of a row segment interleave their dependent DFMA chains instead of running one after the other.
__device__ __forceinline__ double tanh_pos_f64(double s) {
    s = fmin(fmax(s, 0.0), 20.0);
    const double x = -2.0 * s;
    const double kf = rint(x * kTanhC[0]);
    double r = fma(kf, kTanhC[1], x);
    r = fma(kf, kTanhC[2], r);
    double q = kTanhC[3];
    q = fma(q, r, kTanhC[4]);
    q = fma(q, r, kTanhC[5]);
    q = fma(q, r, kTanhC[6]);
    q = fma(q, r, kTanhC[7]);
    q = fma(q, r, kTanhC[8]);
    q = fma(q, r, kTanhC[9]);
    q = fma(q, r, kTanhC[10]);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    q = fma(q, r, 1.0);
    const int k = (int)kf;  // in [-58, 0]
    const double t = __hiloint2double(__double2hiint(q) + k * 1048576, __double2loint(q));  // q * 2^k
    const double d = 1.0 + t;  // in (1, 2]
    double rc;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(d));  // ~20 bits
    rc = fma(rc, fma(-d, rc, 1.0), rc);
    rc = fma(rc, fma(-d, rc, 1.0), rc);
    return (1.0 - t) * rc;
}

// tanh(s) for s > 0 in float64, table-driven (the occupancy-driven forward's epilogue: one call per output voxel, so its
// instruction count is the kernel's — the polynomial version above was 38 instructions, this one ~24):
// exp(-2s) = 2^k * 2^(j/64) * e^r with n = rint(-2s * 64/ln2) = 64 k + j, |r| <= ln2/128: degree-4 Taylor (1e-13),
// 2^(j/64) from a 64-entry table (`tab`: shared memory copy of kExp2Tab — constant memory would serialise the per-lane
// index); quotient (1 - t)/(1 + t) from the reciprocal seed + two Newton steps.  Absolute error < 1e-12.
__constant__ double kExp2Tab[64] = {
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
    1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303,
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267,
    1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951};
// (takes the float32 sum: the clamp to [0, 20] — tanh(20) == 1 to the last bit, negative sums give exactly 0 — is one
// float32 instruction instead of a float64 compare + selects; branch-free, so that two calls interleave their DFMA chains)
__device__ __forceinline__ double tanh_pos_f64_tab(float sf, const double* __restrict__ tab) {
    const double x = -2.0 * (double)fminf(fmaxf(sf, 0.f), 20.f);
    const double t = fma(x, 92.33248261689366, 6755399441055744.0);  // + 1.5 * 2^52: the low word is rint(x * 64/ln2)
    const int n = __double2loint(t);                                  // in [-3694, 0]
    const double nf = t - 6755399441055744.0;
    const double r = fma(nf, -0.010830424696249145, x);               // x - n ln2/64
    double q = fma(r, 4.1666666666666664e-02, 1.6666666666666666e-01);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    q = fma(q, r, 1.0);
    const double u0 = tab[n & 63] * q;
    const double u = __hiloint2double(__double2hiint(u0) + ((n >> 6) << 20), __double2loint(u0));  // u0 * 2^k, k in [-58, 0]
    const double d = 1.0 + u;  // in (1, 2]
    double rc;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(d));
    rc = fma(rc, fma(-d, rc, 1.0), rc);
    rc = fma(rc, fma(-d, rc, 1.0), rc);
    return (1.0 - u) * rc;
}

// ---- the sign of sums near zero --------------------------------------------------------------------------------------
// pred = relu(tanh(s)) has a kink at s = 0 and its derivative jumps from 0 to 1 there: a voxel whose sum changes SIGN
// changes that voxel's whole gradient contribution.  On BASELINE config 2 at its full size (8.4 M voxels) two voxels have
// |s| < 1e-7, and one of them came out on the other side of the kink than in the reference's float64 convolution: the 11
// parameter gradients moved by 2.4e-4 relative, while with the reference's gates they agree to 2e-7
// (profiles/r2_diag_gate.log).  A float64 re-evaluation of such sums with unrounded taps was built and measured: it makes
// the dense and the occupancy-driven kernels agree with each other, but NOT with the reference — two float32 evaluations of
// the GENEO kernels (ours / the reference's CPU ops / the reference's own CUDA ops) differ by an ulp in some taps, up to
// 7e-8, which is the size of these sums: the sign of a sum below ~1e-7 is not defined by the model.  It cost 30 % of both
// forward kernels and was removed; the full-size parity tests give such voxels no upstream gradient and check every
// other gate (tests/test_gpu_full_size.py).

// Tail of the tap-gradient kernels: the CTA that draws the last ticket sums the partial rows of all CTAs in row order
// (fixed order: deterministic) into W — no separate reduction launch.  `ticket` is zeroed once per step by
// sn_grid_prepare; `total` = number of CTAs that take a ticket.  Call with all threads of the CTA.
__device__ __forceinline__ void last_cta_row_sum(unsigned long long* ticket, const double* __restrict__ partial, int rows, int TP,
                                                 int T, double* __restrict__ W, double* __restrict__ scratch, int total = -1) {
    // scratch: shared memory of this CTA, at least min(nwarps, kRowSumGroups) * 256 doubles (the tile stages are dead)
    constexpr int kRowSumGroups = 16;
    __shared__ int s_last;
    __threadfence();  // this CTA's row is visible device-wide before its ticket is drawn
    __syncthreads();
    const int nthreads = blockDim.x * blockDim.y, tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (tid == 0) s_last = atomicAdd(ticket, 1ULL) == (unsigned long long)((total < 0 ? rows : total) - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // row group g (TW threads each) sums rows g, g + G, ...; a thread owns one tap per sweep and keeps 16 loads in flight;
    // the G group sums are then added in group order: a fixed order whatever the timing
    const int TW = nthreads >= 256 ? 256 : nthreads;
    const int G = nthreads / TW < kRowSumGroups ? nthreads / TW : kRowSumGroups;
    const int g = tid / TW, tl = tid % TW;
    for (int t0 = 0; t0 < T; t0 += 256) {  // 256 taps per sweep (scratch [G][256])
        if (g < G) {
            for (int tq = tl; tq < 256; tq += TW) {
                const int tc = t0 + tq;
                double a = 0.0;
                if (tc < T) {
                    for (int r = g; r < rows; r += 16 * G) {
                        double v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = (r + i * G < rows) ? __ldcg(partial + (size_t)(r + i * G) * TP + tc) : 0.0;
#pragma unroll
                        for (int i = 0; i < 16; ++i) a += v[i];
                    }
                }
                scratch[g * 256 + tq] = a;
            }
        }
        __syncthreads();
        for (int tq = tid; tq < 256 && t0 + tq < T; tq += nthreads) {
            double a = 0.0;
            for (int q = 0; q < G; ++q) a += scratch[q * 256 + tq];
            W[t0 + tq] = a;
        }
        __syncthreads();
    }
    if (tid == 0) *ticket = 0ULL;  // ready for another backward on the same count buffer (retain_graph, re-runs)
}

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
