// Occupancy-driven observer forward: pred = relu(tanh(s)), s[v] = sum_{u : x[u] != 0} x[u] * Kstar[u - v + pad]
//
// The dense stencil (stencil_fwd_impl.cuh) multiplies every output voxel by all T taps although 98.4 % of a
// TS40K occupancy grid is zero.  This kernel scatters instead: every NON-ZERO voxel of the halo tile adds its
// kx*ky taps of one z-slice of Kstar to one output plane at a time, lanes <-> taps, accumulators in shared memory.
//
// Per 8 x IX x IY output tile (CTA = 8 warps, warp zo owns output plane zo):
//   A  warp w scans halo z-rows w, w+8, ... of x straight from global memory / L2 (coalesced 16-byte loads, all of a
//      row's loads in flight at once; out-of-bounds = 0 = the zero padding) and compacts their non-zero voxels into
//      per-row lists {value, plane offset} in shared memory (ballot + popc: position order, deterministic).
//      No halo tile is staged: a CTA needs only 44 KB of shared memory, so four CTAs (32 warps) share an SM and
//      hide the latency of the read-modify-write chains of phase B (the first version staged the halo by TMA, was
//      limited to 16 warps per SM and ran at the dense stencil's speed);
//   B  warp zo walks the lists of rows zo .. zo+kz-1 (dz = row - zo): each lane holds its tap(s) of slice dz in
//      a register and does  acc[base - off_lane] += value * tap  (LDS, FFMA, STS).  The plane pitch is chosen
//      = ky (mod 32), so the kx*ky addresses of one voxel fall into distinct banks; planes carry halo margins so
//      that no lane ever needs a bounds test (margins are scratch: never read);
//   C  warp zo applies relu(tanh(.)) to the interior of its plane and stores it in the caller's dtype.
// Rows with more non-zeros than the list capacity are handled in further rounds (rescan from the next ordinal),
// so the kernel is correct for any occupancy; it is only SELECTED (on the device, from sn_grid_prepare's count)
// when the grid is sparse.  No floating-point atomics; bit-deterministic.
#include <stdlib.h>
#include "stencil_common.cuh"

namespace sn {

constexpr int kFsWarps = 8;
constexpr int kFsThreads = kFsWarps * 32;
constexpr int kFsCap = 128;  // list entries per halo z-row and round
constexpr int kFsMaxIt = 4;  // 16-byte loads in flight per lane while scanning a halo row

struct FsParams {
    const float* x;
    const float* Kstar;
    void* pred;
    const unsigned long long* nnz;  // device; NULL = always run
    unsigned long long nnz_max;     // run iff *nnz <= nnz_max ...
    unsigned long long dw_max;      // ... and nnz[2] <= dw_max (not clustered)
    int B, Z, X, Y, kz, kx, ky;
    int out_f64, tanh64;
    int IX, IY;                 // output tile (z extent kRZ)
    int HZ, HX, WS;             // halo box
    int plz, plx, pla, off;     // box start = tile origin - (plz, plx, pla); off = pla - left pad y (dead columns)
    int AS, RP;                 // accumulator plane: row pitch (floats), rows per plane
    int tiles_z, tiles_x, tiles_y, ntiles;
    unsigned ws_magic;          // ceil(2^24 / WS): f / WS == (f * ws_magic) >> 24 for f < 2^13
    const unsigned* mask;       // occupancy bits of x by flat voxel index (sn_grid_prepare), nw words; fwd_occ_kernel only
    int nw;
    // several observers on the same grids (SCENENetQuantile): Kstar holds nq tap sets [nq][T], pred nq
    // outputs pred_qstride elements apart; the non-zero voxels of a tile are listed ONCE for all of them (fwd_occ_kernel)
    int nq;
    long long pred_qstride;
};

struct FsEntry {
    float val;
    int base;  // (ux + kx-1) * AS + (column - off) + ky-1
};

// per (scan iteration, lane) constants of phase A, built once per CTA: position of the lane's 16-byte load inside
// a halo row and what its four voxels map to
struct FsScan {
    int ux;       // halo x-row of the load
    int c;        // first column of the load inside the halo row
    int base;     // accumulator offset of the first voxel: (ux + kx-1) * AS + (c - off) + ky-1
    unsigned live;  // bit q: column c + q is a live column (not alignment padding); 0 past the end of the row
};

template <int NI2>
__global__ void __launch_bounds__(kFsThreads, 4)
fwd_sparse_kernel(const FsParams p) {
    if (p.nnz && !fwd_sparse_selected(p.nnz, p.nnz_max, p.dw_max)) return;  // dense or clustered input: stencil_fwd_kernel does the work
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int plane_floats = p.RP * p.AS;
    const int P = p.kx * p.ky, T = p.kz * P;
    const int row4 = (p.HX * p.WS) >> 2, nit = ceil_div(row4, 32);
    float* acc = reinterpret_cast<float*>(smem_raw);
    const int acc_floats = (kRZ * plane_floats + p.kx * p.AS + 31) & ~31;
    FsEntry* lists = reinterpret_cast<FsEntry*>(acc + acc_floats);  // [HZ][kFsCap]
    FsScan* scan = reinterpret_cast<FsScan*>(lists + p.HZ * kFsCap);  // [nit][32]
    float* sk = reinterpret_cast<float*>(scan + nit * 32);
    int* cnt = reinterpret_cast<int*>(sk + ((T + 31) & ~31));         // [HZ] non-zeros per halo z-row

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x;

    for (int t = tid; t < T; t += kFsThreads) sk[t] = __ldg(p.Kstar + t);
    for (int j = tid; j < nit * 32; j += kFsThreads) {
        FsScan e;
        const int f = j << 2;
        e.ux = f / p.WS;
        e.c = f - e.ux * p.WS;
        e.base = (e.ux + p.kx - 1) * p.AS + (e.c - p.off) + (p.ky - 1);
        e.live = 0u;
        if (j < row4)
            for (int q = 0; q < 4; ++q)
                if (e.c + q >= p.off && e.c + q < p.off + p.IY + p.ky - 1) e.live |= 1u << q;
        scan[j] = e;
    }
    __syncthreads();

    // lane -> plane taps t' = 32 j + lane (t' = dx * ky + dy); dead lanes point at their own scratch word
    int offp[NI2];
    bool okp[NI2];
#pragma unroll
    for (int j = 0; j < NI2; ++j) {
        const int tp = 32 * j + lane;
        okp[j] = tp < P;
        const int tq = okp[j] ? tp : 0;
        offp[j] = (tq / p.ky) * p.AS + (tq % p.ky);
    }
    float* accp = acc + warp * plane_floats;  // this warp's output plane (zo = warp)
    const unsigned lt_mask = (1u << lane) - 1u;
    const bool vec = (p.Y & 3) == 0;

    for (int tile = blockIdx.x; tile < p.ntiles; tile += G) {
        int b, z0, x0, y0;
        {
            int t = tile;
            const int ty = t % p.tiles_y; t /= p.tiles_y;
            const int tx = t % p.tiles_x; t /= p.tiles_x;
            const int tz = t % p.tiles_z;
            b = t / p.tiles_z;
            z0 = tz * kRZ; x0 = tx * p.IX; y0 = ty * p.IY;
        }
        // zero this warp's plane (16-byte stores; plane_floats is a multiple of 4 by construction)
        {
            float4* a4 = reinterpret_cast<float4*>(accp);
            for (int i = lane; i < (plane_floats >> 2); i += 32) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        int lo = 0;
        while (true) {
            // ---- A: compact the non-zero voxels [lo, lo + cap) of every halo z-row
            bool my_more = false;
            for (int zr = warp; zr < p.HZ; zr += kFsWarps) {
                FsEntry* lst = lists + zr * kFsCap;
                const int gz = z0 - p.plz + zr;
                const bool z_ok = gz >= 0 && gz < p.Z;
                const float* plane = p.x + ((size_t)b * p.Z + (z_ok ? gz : 0)) * p.X * p.Y;
                int n = 0;
                for (int it0 = 0; it0 < nit; it0 += kFsMaxIt) {
                    // up to kFsMaxIt loads in flight per lane (memory-level parallelism), then their compaction
                    float4 v[kFsMaxIt];
                    int basev[kFsMaxIt];
                    unsigned live[kFsMaxIt];
#pragma unroll
                    for (int it = 0; it < kFsMaxIt; ++it) {
                        v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                        live[it] = 0u;
                        basev[it] = 0;
                        if (it0 + it < nit) {
                            const FsScan sc = scan[(it0 + it) * 32 + lane];
                            basev[it] = sc.base;
                            const int gx = x0 - p.plx + sc.ux, gy = y0 - p.pla + sc.c;
                            if (z_ok && gx >= 0 && gx < p.X && sc.live) {
                                const float* src = plane + (size_t)gx * p.Y + gy;
                                if (vec) {  // gy and Y are multiples of 4: the four voxels are all inside or all outside
                                    if (gy >= 0 && gy < p.Y) {
                                        v[it] = __ldg(reinterpret_cast<const float4*>(src));
                                        live[it] = sc.live;
                                    }
                                } else {
                                    unsigned m = 0u;
                                    if (gy >= 0 && gy < p.Y) { v[it].x = __ldg(src); m |= 1u; }
                                    if (gy + 1 >= 0 && gy + 1 < p.Y) { v[it].y = __ldg(src + 1); m |= 2u; }
                                    if (gy + 2 >= 0 && gy + 2 < p.Y) { v[it].z = __ldg(src + 2); m |= 4u; }
                                    if (gy + 3 >= 0 && gy + 3 < p.Y) { v[it].w = __ldg(src + 3); m |= 8u; }
                                    live[it] = sc.live & m;
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int it = 0; it < kFsMaxIt; ++it) {
                        if (it0 + it < nit) {
                            const float vals[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const bool nz = vals[q] != 0.f && ((live[it] >> q) & 1u);
                                const unsigned bal = __ballot_sync(0xffffffffu, nz);
                                if (bal == 0u) continue;
                                const int pos = n + __popc(bal & lt_mask) - lo;
                                if (nz && pos >= 0 && pos < kFsCap) {
                                    FsEntry e;
                                    e.val = vals[q];
                                    e.base = basev[it] + q;
                                    lst[pos] = e;
                                }
                                n += __popc(bal);
                            }
                        }
                    }
                }
                if (lane == 0) cnt[zr] = n;
                my_more |= n > lo + kFsCap;
            }
            const int more = __syncthreads_or(my_more ? 1 : 0);
            // ---- B: warp zo adds slice dz of the taps at every listed voxel of row zo + dz
            for (int dz = 0; dz < p.kz; ++dz) {
                const int zr = warp + dz;
                int n = cnt[zr] - lo;
                n = n < 0 ? 0 : (n > kFsCap ? kFsCap : n);
                if (n == 0) continue;
                float kk[NI2];
#pragma unroll
                for (int j = 0; j < NI2; ++j) kk[j] = okp[j] ? sk[dz * P + 32 * j + lane] : 0.f;
                const FsEntry* lst = lists + zr * kFsCap;
                FsEntry en = lst[0];  // same address for all lanes: broadcast
                for (int e = 0; e < n; ++e) {
                    const FsEntry cur = en;
                    if (e + 1 < n) en = lst[e + 1];  // the next entry's load overlaps this entry's read-modify-write
                    float* a = accp + cur.base;
#pragma unroll
                    for (int j = 0; j < NI2; ++j)
                        if (okp[j]) a[-offp[j]] = fmaf(cur.val, kk[j], a[-offp[j]]);
                    __syncwarp();  // the next voxel may touch the same accumulators from other lanes
                }
            }
            if (!more) break;
            __syncthreads();  // the lists are rewritten by the next round
            lo += kFsCap;
        }
        // ---- C: epilogue of this warp's plane (lane -> 4 consecutive y)
        {
            const int gz = z0 + warp;
            const int groups_y = p.IY >> 2;
            for (int g = lane; g < p.IX * groups_y; g += 32) {
                const int xo = g / groups_y, yo = (g % groups_y) << 2;
                const int gx = x0 + xo, gy = y0 + yo;
                if (gz >= p.Z || gx >= p.X || gy >= p.Y) continue;
                const float* a = accp + (xo + p.kx - 1) * p.AS + yo + (p.ky - 1);
                float o[4];
                double od[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float s = a[r];
                    o[r] = (!p.tanh64 && s > 0.f) ? tanhf(s) : 0.f;
                    if (p.tanh64) od[r] = s > 0.f ? tanh_pos_f64((double)s) : 0.0;
                }
                const size_t idx = (((size_t)b * p.Z + gz) * p.X + gx) * p.Y + gy;
                const int ny = p.Y - gy;
                if (p.out_f64) {
                    double* out = reinterpret_cast<double*>(p.pred) + idx;
                    if (!p.tanh64) {
#pragma unroll
                        for (int r = 0; r < 4; ++r) od[r] = (double)o[r];
                    }
                    if (vec) {
                        reinterpret_cast<double2*>(out)[0] = make_double2(od[0], od[1]);
                        reinterpret_cast<double2*>(out)[1] = make_double2(od[2], od[3]);
                    } else {
#pragma unroll
                        for (int r = 0; r < 4; ++r)
                            if (r < ny) out[r] = od[r];
                    }
                } else {
                    float* out = reinterpret_cast<float*>(p.pred) + idx;
                    if (vec) {
                        *reinterpret_cast<float4*>(out) = make_float4(o[0], o[1], o[2], o[3]);
                    } else {
#pragma unroll
                        for (int r = 0; r < 4; ++r)
                            if (r < ny) out[r] = o[r];
                    }
                }
            }
        }
        __syncthreads();  // lists and counts are rewritten by the next tile
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Mask-driven variant (the one that runs when the caller hands over sn_grid_prepare's state buffer).
//
// ncu of the scanning kernel above at config 2 (profiles/r1_notes.md): 74 M warp instructions, issue slots 79 % busy —
// bound by instruction issue, and 30 M of them were phase A (every halo voxel is looked at by ~3 tiles: 16-byte
// loads, bounds tests, four ballots per load).  sn_grid_prepare now leaves one occupancy BIT per voxel next to the
// count, so phase A of this kernel reads a halo row's 32-column chunks as funnel-shifted mask words (one lane per
// (x-row, chunk) pair), gets list positions from a warp prefix sum of the popcounts and only touches x for the
// voxels that are set.  Phase B walks two list entries per 16-byte load with hand-formed shared-memory addresses;
// phase C's index arithmetic is hoisted out of the tile loop.  Same lists, fixed accumulation order, no atomics.
constexpr int kFoPairIt = 3;  // (x-row, chunk) pairs per lane: HX * ceil((IY + ky - 1) / 32) <= 96

// acc[addr] += v * k for the lanes with ok != 0 (predicated: no branch, the warp stays converged)
__device__ __forceinline__ void fo_rmw(uint32_t addr, float v, float k, int ok) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t.reg .f32 t;\n\t"
        "setp.ne.s32 q, %3, 0;\n\t"
        "@q ld.shared.f32 t, [%0];\n\t"
        "@q fma.rn.f32 t, %1, %2, t;\n\t"
        "@q st.shared.f32 [%0], t;\n\t}" ::"r"(addr),
        "f"(v), "f"(k), "r"(ok)
        : "memory");
}

template <int NI2, bool OUT64, bool MULTI>
__global__ void __launch_bounds__(kFsThreads, NI2 <= 2 ? 4 : 3)
fwd_occ_kernel(const FsParams p) {
    const int nq = MULTI ? p.nq : 1;  // compile-time 1 for the single-observer instantiation: its q loop folds away
    if (p.nnz && !fwd_sparse_selected(p.nnz, p.nnz_max, p.dw_max)) return;  // dense or clustered input: stencil_fwd_kernel does the work
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int plane_floats = p.RP * p.AS;
    const int P = p.kx * p.ky, T = p.kz * P;
    float* acc = reinterpret_cast<float*>(smem_raw);
    const int acc_floats = (kRZ * plane_floats + p.kx * p.AS + 31) & ~31;
    FsEntry* lists = reinterpret_cast<FsEntry*>(acc + acc_floats);  // [HZ][kFsCap]
    float* sk = reinterpret_cast<float*>(lists + p.HZ * kFsCap);
    int* cnt = reinterpret_cast<int*>(sk + ((nq * T + 31) & ~31));  // [HZ] non-zeros per halo z-row

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x;
    for (int t = tid; t < nq * T; t += kFsThreads) sk[t] = __ldg(p.Kstar + t);

    // phase A constants of this lane: its (x-row, 32-column chunk) pair in each iteration
    // (flat voxel indices fit 32 bits: the launcher only picks this kernel below 2^31 - 2^20 voxels)
    const int ply = p.pla - p.off;
    const int HW = p.IY + p.ky - 1;
    const int nwc = (HW + 31) >> 5;
    const int npairs = p.HX * nwc;
    int pux[kFoPairIt], pc0[kFoPairIt], poff[kFoPairIt], pbase[kFoPairIt];
#pragma unroll
    for (int i = 0; i < kFoPairIt; ++i) {
        const int pr = 32 * i + lane;
        pux[i] = pr < npairs ? pr / nwc : -1;
        pc0[i] = pr < npairs ? (pr - pux[i] * nwc) * 32 : 0;
        poff[i] = pux[i] * p.Y + pc0[i];                                     // voxel offset from the halo row's origin
        pbase[i] = ((pux[i] + p.kx - 1) * p.AS + pc0[i] + (p.ky - 1)) << 2;  // accumulator byte offset of bit 0
    }
    const int zstep = p.X * p.Y;
    // phase B: lane -> plane taps t' = 32 j + lane (t' = dx * ky + dy)
    float* accp = acc + warp * plane_floats;  // this warp's output plane (zo = warp)
    uint32_t accl[NI2];  // shared-memory byte address of this lane's accumulator for an entry with base 0
    int okp[NI2];
#pragma unroll
    for (int j = 0; j < NI2; ++j) {
        const int tp = 32 * j + lane;
        okp[j] = tp < P ? 1 : 0;
        const int tq = okp[j] ? tp : 0;
        accl[j] = smem_u32(accp) - 4u * (uint32_t)((tq / p.ky) * p.AS + (tq % p.ky));
    }
    const uint32_t lists_w = smem_u32(lists + warp * kFsCap);  // list of halo z-row `warp` (dz = 0)
    const float* skl = sk + lane;
    // phase C: lane -> four 4-voxel groups of the 8 x IX x IY tile's plane (IX * IY == 512)
    const int lg_gy = p.IY == 64 ? 4 : 3;  // 4-voxel groups per plane row: IY / 4 = 16 or 8
    const bool vec = (p.Y & 3) == 0;
    __syncthreads();

    // tile coordinates advance by the (decoded) grid stride with carries: no divisions inside the tile loop
    int tc[4], ts[4];  // ty, tx, tz, b of the current tile / of the stride G
    {
        int t = blockIdx.x, g = G;
        tc[0] = t % p.tiles_y; t /= p.tiles_y; ts[0] = g % p.tiles_y; g /= p.tiles_y;
        tc[1] = t % p.tiles_x; t /= p.tiles_x; ts[1] = g % p.tiles_x; g /= p.tiles_x;
        tc[2] = t % p.tiles_z; tc[3] = t / p.tiles_z; ts[2] = g % p.tiles_z; ts[3] = g / p.tiles_z;
    }
    for (int tile = blockIdx.x; tile < p.ntiles; tile += G) {
        const int b = tc[3], z0 = tc[2] * kRZ, x0 = tc[1] * p.IX, y0 = tc[0] * p.IY;
        {
            tc[0] += ts[0];
            int cy = tc[0] >= p.tiles_y ? 1 : 0;
            tc[0] -= cy ? p.tiles_y : 0;
            tc[1] += ts[1] + cy;
            cy = tc[1] >= p.tiles_x ? 1 : 0;
            tc[1] -= cy ? p.tiles_x : 0;
            tc[2] += ts[2] + cy;
            cy = tc[2] >= p.tiles_z ? 1 : 0;
            tc[2] -= cy ? p.tiles_z : 0;
            tc[3] += ts[3] + cy;
        }
        // live halo columns of this tile (gy = y0 - ply + c inside [0, Y)) and live x-rows, per pair of this lane
        unsigned vm[kFoPairIt];
        {
            const int c_lo = max(0, ply - y0), c_hi = min(HW, p.Y - y0 + ply);
#pragma unroll
            for (int i = 0; i < kFoPairIt; ++i) {
                const int gx = x0 - p.plx + pux[i];
                const int a = min(max(c_lo - pc0[i], 0), 32), e = min(max(c_hi - pc0[i], 0), 32);
                vm[i] = (pux[i] >= 0 && gx >= 0 && gx < p.X && e > a) ? ((0xffffffffu >> (32 - (e - a))) << a) : 0u;
            }
        }
        // flat index of the halo box origin (z-row 0, x-row 0, column 0); may be negative at the grid's first rows
        const int org = ((b * p.Z + (z0 - p.plz)) * p.X + (x0 - p.plx)) * p.Y + (y0 - ply);
        // One pass per observer q (nq == 1 outside SCENENetQuantile).  The lists of round 0 serve every q; only a tile
        // with an overflowing row (more than kFsCap non-zeros: further rounds rewrite the lists) lists again for q > 0.
        bool multi_round = false;
        for (int q = 0; q < nq; ++q) {
        if (q > 0 && multi_round) __syncthreads();  // every warp is done reading the last round's lists of q - 1
        {
            float4* a4 = reinterpret_cast<float4*>(accp);
            for (int i = lane; i < (plane_floats >> 2); i += 32) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        int lo = 0;
        while (true) {
            int more = 0;
            if (q == 0 || multi_round) {
            // ---- A: list the non-zero voxels [lo, lo + cap) of every halo z-row from the occupancy bits
            bool my_more = false;
            for (int zr = warp; zr < p.HZ; zr += kFsWarps) {
                FsEntry* lst = lists + zr * kFsCap;
                const int gz = z0 - p.plz + zr;
                const bool z_ok = gz >= 0 && gz < p.Z;
                int n = 0;
#pragma unroll
                for (int i = 0; i < kFoPairIt; ++i) {
                    if (32 * i >= npairs) break;  // warp-uniform
                    unsigned m = 0u;
                    const int bit0 = org + zr * zstep + poff[i];
                    if (z_ok && vm[i]) {
                        const int wi = bit0 >> 5;  // floor
                        const unsigned w0 = (unsigned)wi < (unsigned)p.nw ? __ldg(p.mask + wi) : 0u;
                        const unsigned w1 = (unsigned)(wi + 1) < (unsigned)p.nw ? __ldg(p.mask + wi + 1) : 0u;
                        m = __funnelshift_r(w0, w1, (unsigned)bit0 & 31u) & vm[i];
                    }
                    const int c = __popc(m);
                    int incl = c;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += t;
                    }
                    const int total = __shfl_sync(0xffffffffu, incl, 31);
                    int pos = n + incl - c - lo;
                    while (m) {
                        const int bp = __ffs(m) - 1;
                        m &= m - 1u;
                        if (pos >= 0 && pos < kFsCap) {
                            FsEntry en;
                            en.val = __ldg(p.x + (bit0 + bp));
                            en.base = pbase[i] + (bp << 2);  // byte offset inside the plane
                            lst[pos] = en;
                        }
                        ++pos;
                    }
                    n += total;
                }
                if (lane == 0) {
                    cnt[zr] = n;
                    // phase B walks the list two entries at a time: an odd list gets a no-op entry (value 0 at the
                    // accumulators of halo voxel (0, 0): inside this warp's own plane for every tap)
                    const int nl = n - lo;
                    if (nl > 0 && nl < kFsCap && (nl & 1)) {
                        FsEntry en;
                        en.val = 0.f;
                        en.base = ((p.kx - 1) * p.AS + (p.ky - 1)) << 2;
                        lst[nl] = en;
                    }
                }
                my_more |= n > lo + kFsCap;
            }
            more = __syncthreads_or(my_more ? 1 : 0);
            if (lo == 0 && more) multi_round = true;
            }
            // ---- B: warp zo adds slice dz of the taps at every listed voxel of row zo + dz
            for (int dz = 0; dz < p.kz; ++dz) {
                int n = cnt[warp + dz] - lo;
                n = n < 0 ? 0 : (n > kFsCap ? kFsCap : n);
                if (n == 0) continue;
                float kk[NI2];
#pragma unroll
                for (int j = 0; j < NI2; ++j) kk[j] = okp[j] ? skl[q * T + dz * P + 32 * j] : 0.f;
                // two list entries per 16-byte broadcast load; shared-memory addresses are formed by hand (the generic
                // C++ form cost 12 instructions per entry, this one 5) and the read-modify-write is predicated, not
                // branched, so the warp stays converged: its LDS / STS are executed in program order, which is what
                // makes an entry see the sums its predecessor stored from OTHER lanes
                __syncwarp();
                uint32_t la = lists_w + (uint32_t)dz * (kFsCap * 8u);
                const uint32_t lend = la + (((uint32_t)n + 1u) >> 1) * 16u;
#pragma unroll 1
                for (; la != lend; la += 16u) {
                    float v0, v1;
                    uint32_t b0, b1;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=f"(v0), "=r"(b0), "=f"(v1), "=r"(b1) : "r"(la) : "memory");
#pragma unroll
                    for (int j = 0; j < NI2; ++j) fo_rmw(accl[j] + b0, v0, kk[j], okp[j]);
#pragma unroll
                    for (int j = 0; j < NI2; ++j) fo_rmw(accl[j] + b1, v1, kk[j], okp[j]);
                }
            }
            if (!more) break;
            __syncthreads();  // the lists are rewritten by the next round
            lo += kFsCap;
        }
        // ---- C: epilogue of this warp's plane (lane -> 4 consecutive y)
        {
            const int gz = z0 + warp;
            const size_t idx0 = (size_t)q * (size_t)p.pred_qstride + (((size_t)b * p.Z + gz) * p.X + x0) * p.Y + y0;
#pragma unroll 1
            for (int i = 0; i < 4; ++i) {
                const int g = lane + 32 * i;
                const int xo = g >> lg_gy, yo = (g & ((1 << lg_gy) - 1)) << 2;
                const int gx = x0 + xo, gy = y0 + yo;
                if (gz >= p.Z || gx >= p.X || gy >= p.Y) continue;
                const float* a = accp + (xo + p.kx - 1) * p.AS + yo + (p.ky - 1);
                const size_t idx = idx0 + (size_t)xo * p.Y + yo;
                const int ny = p.Y - gy;
                if constexpr (OUT64) {  // float64 predictions: tanh evaluated in float64 (tanh_pos_f64)
                    double od[4] = {0.0, 0.0, 0.0, 0.0};
                    const float s0 = a[0], s1 = a[1], s2 = a[2], s3 = a[3];
                    // the four evaluations are branch-free and interleave; a warp whose 128 sums are all <= 0 (empty
                    // regions of a scene) skips them
                    if (__any_sync(__activemask(), fmaxf(fmaxf(s0, s1), fmaxf(s2, s3)) > 0.f)) {
                        od[0] = tanh_pos_f64((double)s0);  // relu inside: negative sums clamp to 0
                        od[1] = tanh_pos_f64((double)s1);
                        od[2] = tanh_pos_f64((double)s2);
                        od[3] = tanh_pos_f64((double)s3);
                    }
                    double* out = reinterpret_cast<double*>(p.pred) + idx;
                    if (vec) {
                        reinterpret_cast<double2*>(out)[0] = make_double2(od[0], od[1]);
                        reinterpret_cast<double2*>(out)[1] = make_double2(od[2], od[3]);
                    } else {
#pragma unroll
                        for (int r = 0; r < 4; ++r)
                            if (r < ny) out[r] = od[r];
                    }
                } else {
                    float o[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float s = a[r];
                        o[r] = s > 0.f ? tanhf(s) : 0.f;
                    }
                    float* out = reinterpret_cast<float*>(p.pred) + idx;
                    if (vec) {
                        *reinterpret_cast<float4*>(out) = make_float4(o[0], o[1], o[2], o[3]);
                    } else {
#pragma unroll
                        for (int r = 0; r < 4; ++r)
                            if (r < ny) out[r] = o[r];
                    }
                }
            }
        }
        }  // q
        __syncthreads();  // lists and counts are rewritten by the next tile
    }
}

// geometry; false when the kernel does not cover the shape (caller uses the dense stencil)
static bool plan_fwd_sparse(int B, int Z, int X, int Y, int kz, int kx, int ky, FsParams& p, size_t& smem, int& ni2, int nq = 1) {
    p.B = B; p.Z = Z; p.X = X; p.Y = Y; p.kz = kz; p.kx = kx; p.ky = ky;
    const int P = kx * ky;
    ni2 = ceil_div(P, 32);
    if (ni2 > 4) return false;
    p.IY = Y > 32 ? 64 : 32;
    p.IX = 512 / p.IY;
    const int ply = pad_left(ky);
    p.plz = pad_left(kz);
    p.plx = pad_left(kx);
    p.pla = round4(ply);
    p.off = p.pla - ply;
    p.HZ = kRZ + kz - 1;
    p.HX = p.IX + kx - 1;
    p.WS = round4(p.off + p.IY + ky - 1);
    p.RP = p.HX;
    const int as_min = p.IY + ky - 1;
    p.AS = as_min + (((ky - as_min) % 32) + 32) % 32;  // = ky (mod 32): conflict-free tap addresses
    // zeroing uses 16-byte stores: the plane size must be a multiple of 4 floats
    while ((p.RP * p.AS) & 3) ++p.RP;
    p.tiles_z = ceil_div(Z, kRZ);
    p.tiles_x = ceil_div(X, p.IX);
    p.tiles_y = ceil_div(Y, p.IY);
    p.ntiles = B * p.tiles_z * p.tiles_x * p.tiles_y;
    p.ws_magic = (unsigned)(((1u << 24) + p.WS - 1) / p.WS);
    if (p.HX * p.WS >= (1 << 13)) return false;
    const int T = kz * P;
    const size_t accb = (size_t)((kRZ * p.RP * p.AS + kx * p.AS + 31) & ~31) * 4;
    const size_t lst = (size_t)p.HZ * kFsCap * sizeof(FsEntry);
    const size_t taps = (size_t)((nq * T + 31) & ~31) * 4;
    const size_t cntb = (size_t)((p.HZ + 1) & ~1) * 4;
    const size_t scanb = (size_t)ceil_div(p.HX * p.WS / 4, 32) * 32 * 16;
    smem = accb + lst + scanb + taps + cntb + 16;
    return smem <= 227 * 1024;
}

bool fwd_sparse_supported(int B, int Z, int X, int Y, int kz, int kx, int ky) {
    FsParams p{};
    size_t smem;
    int ni2;
    return plan_fwd_sparse(B, Z, X, Y, kz, kx, ky, p, smem, ni2);
}

template <int NI2>
static int launch_fs(FsParams& p, size_t smem, cudaStream_t stream) {
    // the mask-driven kernel needs the occupancy bits, at most kFoPairIt * 32 (x-row, chunk) pairs per halo z-row and
    // 32-bit flat voxel indices (incl. the halo overshoot)
    const bool occ = p.mask && p.HX * ((p.IY + p.ky - 1 + 31) >> 5) <= kFoPairIt * 32 && p.IX * p.IY == 512 &&
                     (long long)p.B * p.Z * p.X * p.Y < (1ll << 31) - (1ll << 20);
    if (p.nq > 1 && !occ) return SN_ERR_UNSUPPORTED;  // only the mask-driven kernel shares its lists between observers
    auto kern = !occ ? fwd_sparse_kernel<NI2>
                : p.nq > 1 ? (p.out_f64 ? fwd_occ_kernel<NI2, true, true> : fwd_occ_kernel<NI2, false, true>)
                           : (p.out_f64 ? fwd_occ_kernel<NI2, true, false> : fwd_occ_kernel<NI2, false, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_rc(e);
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
    const int grid = max(1, min(p.ntiles, kNumSMs * per_sm));
    kern<<<grid, kFsThreads, smem, stream>>>(p);
    SN_LAUNCH_CHECK();
    return SN_OK;
}

int fwd_sparse_launch(const float* x, const float* Kstar, void* pred, int out_f64, const unsigned long long* nnz,
                      unsigned long long nnz_max, unsigned long long dw_max, const unsigned* occ_mask, int B, int Z, int X, int Y,
                      int kz, int kx, int ky, int nq, cudaStream_t stream) {
    FsParams p{};
    size_t smem;
    int ni2;
    if (nq < 1 || !plan_fwd_sparse(B, Z, X, Y, kz, kx, ky, p, smem, ni2, nq)) return SN_ERR_UNSUPPORTED;
    p.nq = nq;
    p.pred_qstride = (long long)B * Z * X * Y;
    p.x = x; p.Kstar = Kstar; p.pred = pred; p.out_f64 = out_f64; p.nnz = nnz; p.nnz_max = nnz_max; p.dw_max = dw_max;
    p.tanh64 = out_f64;  // float64 predictions: tanh evaluated in float64 (tanh_pos_f64)
    {
        static const bool no_mask = getenv("SN_FWD_NO_MASK") != nullptr;  // measurement: force the scanning kernel
        p.mask = no_mask ? nullptr : occ_mask;
        const long long nw = ((long long)B * Z * X * Y + 31) >> 5;
        p.nw = nw < (1ll << 30) ? (int)nw : 0;
    }
    if ((uintptr_t)x & 15) return SN_ERR_ALIGN;
    switch (ni2) {
        case 1: return launch_fs<1>(p, smem, stream);
        case 2: return launch_fs<2>(p, smem, stream);
        case 3: return launch_fs<3>(p, smem, stream);
        default: return launch_fs<4>(p, smem, stream);
    }
}

}  // namespace sn
