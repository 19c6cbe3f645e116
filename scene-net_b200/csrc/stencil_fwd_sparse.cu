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
    // fwd_occ_kernel (ABI v4)
    unsigned long long* state;  // the grid state buffer's counters ([3] dense tiles, [4] non-unit voxels, [5] tile counter)
    int* dense_list;            // tiles handed to the dense stencil (NULL: every tile is scattered here)
    int dense_thresh;           // ... when their halo box holds more non-zero voxels than this
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
// Mask-driven kernel (the one that runs when the caller hands over sn_grid_prepare's state buffer) — third version.
//
// History (profiles/r1_notes.md, profiles/r2_notes.md): the scanning kernel above was bound by instruction issue (74 M warp
// instructions at config 2, 30 M of them the scan).  v2 read the halo's occupancy BITS instead (41 M instructions, 62 us)
// and ncu showed where the rest went: ~20 M in the epilogue (a 38-instruction float64 tanh per voxel, scalar loads with
// 4-way bank conflicts), 6 M in the listing (one warp prefix scan per 32-column chunk, per-row lists), and the scatter
// itself at 13 scheduler cycles per (voxel, plane) pair because nothing overlapped its LDS -> FFMA -> STS chains.  v3:
//   A  one CELL = (halo z-row, x-row, 32-column chunk); every thread owns <= kFoMaxCells consecutive cells, reads their
//      mask words (two per cell, funnel-shifted), and ONE block-wide exclusive scan of the popcounts places every
//      non-zero voxel of the halo box in a single list ordered by z-row (row starts padded to even positions: phase B
//      walks 16-byte entry pairs).  Occupancy grids (state[4] == 0: every non-zero is 1) never touch x.
//   B  warp zo scatters rows zo .. zo+kz-1 into its plane as before; the next entry pair is loaded while the current one
//      is applied, and a __syncwarp() after every read-modify-write orders it against the next entry's (other lanes
//      may own the same accumulator: the PTX memory model requires the barrier; measured cost in r2_notes.md).
//   C  lanes <-> consecutive y (conflict-free scalar LDS, coalesced 8-byte stores), the accumulator is zeroed as it is
//      read (no separate zeroing pass; halo margins are scratch and never read), tanh by the table-driven float64
//      evaluation (~24 instructions).
//   Tiles are handed out by an atomic counter in the state buffer (a locally dense tile no longer makes its CTA a
//   straggler), empty tiles are zero-filled with 16-byte stores, and a tile whose halo holds more than `dense_thresh`
//   non-zeros is appended to the tile list for the dense stencil that follows on the stream (per-TILE kernel choice:
//   clustered LiDAR grids get the scatter for their sparse tiles and the FFMA stencil for the ground layer).
#ifndef SN_FO_CTAS
#define SN_FO_CTAS 4  // resident CTAs per SM for <= 32 taps per slice: 62 registers, no spills, 49 us at config 2; 5 CTAs (48 registers, 104 bytes of spills) measured 55 us (profiles/r2_notes.md)
#endif
constexpr int kFoMaxCells = 6;    // cells per thread: HZ * HX * ceil((IY + ky - 1) / 32) <= 1536
constexpr int kFoCap = 1536;      // list entries per round (incl. the <= HZ padding entries)

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
#ifdef SN_FO_NOSYNC  // measurement only: the v2 behaviour (relies on in-order LDS/STS of a converged warp)
#define FO_ORDER()
#define FO_ORDER_PTX
#else
#define FO_ORDER() __syncwarp()
#define FO_ORDER_PTX "bar.warp.sync 0xffffffff;\n\t"
#endif

template <int NI2, bool OUT64, bool MULTI, int CPT>  // CPT: cells per thread this instantiation holds (3 or kFoMaxCells)
__global__ void __launch_bounds__(kFsThreads, NI2 == 1 ? SN_FO_CTAS : (NI2 == 2 ? 4 : 3))
fwd_occ_kernel(const FsParams p) {
    const int nq = MULTI ? p.nq : 1;  // compile-time 1 for the single-observer instantiation: its q loop folds away
    if (p.nnz && !fwd_sparse_selected(p.nnz, p.nnz_max, p.dw_max)) return;  // whole-grid gate (shapes without tile hand-off)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int plane_floats = p.RP * p.AS;
    const int P = p.kx * p.ky, T = p.kz * P;
    float* acc = reinterpret_cast<float*>(smem_raw);
    const int acc_floats = (kRZ * plane_floats + p.kx * p.AS + 31) & ~31;
    FsEntry* lists = reinterpret_cast<FsEntry*>(acc + acc_floats);            // [kFoCap + 2]
    double* tab = reinterpret_cast<double*>(lists + kFoCap + 2);                // [64] 2^(j/64)
    float* sk = reinterpret_cast<float*>(tab + 64);                             // [nq * T] taps
    int* gstart = reinterpret_cast<int*>(sk + ((nq * T + 31) & ~31));           // [HZ + 1] list position of a z-row's first voxel
    int* pstart = gstart + 32;                                                  // [HZ + 1] the same with rows padded to even length
    int* wtot = pstart + 32;                                                    // [8] non-zeros per warp
    int* ctl = wtot + 8;                                                        // [2][4]: tile coordinates b, z0, x0, y0 / -1

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < acc_floats; i += kFsThreads) acc[i] = 0.f;
    for (int t = tid; t < nq * T; t += kFsThreads) sk[t] = __ldg(p.Kstar + t);
    if (tid < 64) tab[tid] = kExp2Tab[tid];

    // ---- phase A constants: this thread's cells
    const int ply = p.pla - p.off;
    const int HW = p.IY + p.ky - 1;
    const int nwc = (HW + 31) >> 5;
    const int cpr = p.HX * nwc;                       // cells per halo z-row
    const int ncells = p.HZ * cpr;
    const int cpt = (ncells + kFsThreads - 1) / kFsThreads;  // <= kFoMaxCells (plan)
    const int zstep = p.X * p.Y;
    int cvox[CPT];       // voxel offset of the cell's first column from the halo box origin
    unsigned cinf[CPT];  // accumulator byte offset of its bit 0 | chunk << 16 | x-row << 18 | z-row << 24
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        const int c = tid * cpt + i;
        const bool ok = i < cpt && c < ncells;
        const int zr = ok ? c / cpr : 0, rem = ok ? c - zr * cpr : 0;
        const int ux = rem / nwc, wc = rem - ux * nwc;
        cvox[i] = zr * zstep + ux * p.Y + wc * 32;
        cinf[i] = ok ? ((unsigned)(((ux + p.kx - 1) * p.AS + wc * 32 + (p.ky - 1)) << 2) | ((unsigned)wc << 16) | ((unsigned)ux << 18) |
                        ((unsigned)zr << 24))
                     : 0xffffffffu;
    }
    // ---- phase B constants: lane -> plane taps t' = 32 j + lane (t' = dx * ky + dy)
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
    const uint32_t lists_u = smem_u32(lists);
    const float* skl = sk + lane;
    const bool vec = (p.Y & 3) == 0;
    const bool binary = p.state ? (p.state[4] == 0ull) : false;  // occupancy grid: every listed value is 1
    // ---- tile hand-out: an atomic counter in the state buffer (static stride without one); thread 0 decodes a tile's
    // coordinates one tile ahead into ctl[slot]
    unsigned long long* tctr = p.state ? p.state + 5 : nullptr;
    const int G = gridDim.x;
    int static_next = blockIdx.x;
    auto fetch = [&](int slot) {  // thread 0 only
        int t;
        if (tctr) {
            t = (int)atomicAdd(tctr, 1ull);
            if (t >= p.ntiles) {
                if (t == p.ntiles + G - 1) atomicExch(tctr, 0ull);  // the last draw of the launch: ready for the next one
                t = -1;
            }
        } else {
            t = static_next < p.ntiles ? static_next : -1;
            static_next += G;
        }
        int* c = ctl + 4 * slot;
        if (t < 0) { c[0] = -1; c[1] = c[2] = c[3] = 0; return; }
        const int ty = t % p.tiles_y; int r = t / p.tiles_y;
        const int tx = r % p.tiles_x; r /= p.tiles_x;
        const int tz = r % p.tiles_z;
        c[0] = r / p.tiles_z; c[1] = tz * kRZ; c[2] = tx * p.IX; c[3] = ty * p.IY;
    };
    if (tid == 0) fetch(0);
    __syncthreads();

    for (int it = 0;; ++it) {
        const int* c = ctl + 4 * (it & 1);
        const int b = c[0], z0 = c[1], x0 = c[2], y0 = c[3];
        if (b < 0) break;
        if (tid == 0) fetch((it + 1) & 1);  // visible after this tile's first barrier; read after its last one

        // ---- A1: occupancy words of this thread's cells
        const int c_lo = max(0, ply - y0), c_hi = min(HW, p.Y - y0 + ply);  // live halo columns [c_lo, c_hi)
        const int org = ((b * p.Z + (z0 - p.plz)) * p.X + (x0 - p.plx)) * p.Y + (y0 - ply);  // flat index of the box origin (may be < 0)
        unsigned m[CPT];
        int mine = 0;
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
            m[i] = 0u;
            if (i < cpt && cinf[i] != 0xffffffffu) {
                const int wc = (cinf[i] >> 16) & 3, ux = (cinf[i] >> 18) & 63, zr = cinf[i] >> 24;
                const int gz = z0 - p.plz + zr, gx = x0 - p.plx + ux;
                const int a = min(max(c_lo - wc * 32, 0), 32), e = min(max(c_hi - wc * 32, 0), 32);
                if (gz >= 0 && gz < p.Z && gx >= 0 && gx < p.X && e > a) {
                    const int bit0 = org + cvox[i];
                    const int wi = bit0 >> 5;  // floor
                    const unsigned w0 = (unsigned)wi < (unsigned)p.nw ? __ldg(p.mask + wi) : 0u;
                    const unsigned w1 = (unsigned)(wi + 1) < (unsigned)p.nw ? __ldg(p.mask + wi + 1) : 0u;
                    m[i] = __funnelshift_r(w0, w1, (unsigned)bit0 & 31u) & ((0xffffffffu >> (32 - (e - a))) << a);
                }
            }
            mine += __popc(m[i]);
        }
        // ---- A2: block-wide exclusive scan of the per-thread counts
        int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();  // (1)
        int pre = incl - mine, total = 0;
#pragma unroll
        for (int w = 0; w < kFsWarps; ++w) {
            const int v = wtot[w];
            pre += w < warp ? v : 0;
            total += v;
        }
        if (total == 0) {
            // empty halo: every sum of the tile is exactly zero (41 of the 64 tiles of the reference's sample_575 grid)
            for (int q = 0; q < nq; ++q) {
                const size_t idx0 = (size_t)q * (size_t)p.pred_qstride + (((size_t)b * p.Z + z0) * p.X + x0) * p.Y + y0;
                for (int g = tid; g < kRZ * p.IX * (p.IY >> 2); g += kFsThreads) {
                    const int yo = (g % (p.IY >> 2)) << 2, r = g / (p.IY >> 2), xo = r % p.IX, zo = r / p.IX;
                    if (z0 + zo >= p.Z || x0 + xo >= p.X || y0 + yo >= p.Y) continue;
                    const size_t idx = idx0 + ((size_t)zo * p.X + xo) * p.Y + yo;
                    const int ny = p.Y - (y0 + yo);
                    if constexpr (OUT64) {
                        double* out = reinterpret_cast<double*>(p.pred) + idx;
                        if (vec) {
                            reinterpret_cast<double2*>(out)[0] = make_double2(0.0, 0.0);
                            reinterpret_cast<double2*>(out)[1] = make_double2(0.0, 0.0);
                        } else {
                            for (int r2 = 0; r2 < 4 && r2 < ny; ++r2) out[r2] = 0.0;
                        }
                    } else {
                        float* out = reinterpret_cast<float*>(p.pred) + idx;
                        if (vec) {
                            *reinterpret_cast<float4*>(out) = make_float4(0.f, 0.f, 0.f, 0.f);
                        } else {
                            for (int r2 = 0; r2 < 4 && r2 < ny; ++r2) out[r2] = 0.f;
                        }
                    }
                }
            }
            __syncthreads();  // wtot / ctl are rewritten by the next tile
            continue;
        }
        if (p.dense_list && total > p.dense_thresh) {
            // locally dense tile: the dense stencil behind us on the stream computes it (per-tile kernel choice)
            if (tid == 0) {
                const int tile = ((b * p.tiles_z + z0 / kRZ) * p.tiles_x + x0 / p.IX) * p.tiles_y + y0 / p.IY;
                const unsigned long long slot = atomicAdd(p.state + 3, 1ull);
                p.dense_list[slot] = tile;
            }
            __syncthreads();
            continue;
        }
        // ---- A3: row starts.  Cells are numbered z-row major, so the list is ordered by z-row.
        {
            int run = pre;
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                if (i < cpt && cinf[i] != 0xffffffffu && (cinf[i] & 0x00ff0000u) == 0u) gstart[cinf[i] >> 24] = run;  // chunk 0 of x-row 0
                run += __popc(m[i]);
            }
            if (tid == 0) gstart[p.HZ] = total;
        }
        __syncthreads();  // (2)
        if (warp == 0) {
            const int r = lane < p.HZ ? lane : p.HZ;
            const int g0 = gstart[r], g1 = lane < p.HZ ? gstart[r + 1] : g0;
            const unsigned odd = __ballot_sync(0xffffffffu, (g1 - g0) & 1);
            if (lane <= p.HZ) pstart[lane] = g0 + __popc(odd & ((1u << lane) - 1u));
        }
        __syncthreads();  // (3)
        const int total_p = pstart[p.HZ];  // padded length of the list

        bool multi_round = total_p > kFoCap;
        for (int q = 0; q < nq; ++q) {
            for (int lo = 0; lo < total_p; lo += kFoCap) {
                if (q == 0 || multi_round) {
                    if (lo > 0 || q > 0) __syncthreads();  // every warp is done with the previous round's list
                    // ---- A4: write the entries of the window [lo, lo + kFoCap)
                    int run = pre;
#pragma unroll
                    for (int i = 0; i < CPT; ++i) {
                        unsigned mm = m[i];
                        if (mm) {
                            const int zr = cinf[i] >> 24;
                            int pos = pstart[zr] + (run - gstart[zr]) - lo;
                            const int bit0 = org + cvox[i];
                            const int base = (int)(cinf[i] & 0xffffu);
                            run += __popc(mm);
                            while (mm) {
                                const int bp = __ffs(mm) - 1;
                                mm &= mm - 1u;
                                if (pos >= 0 && pos < kFoCap) {
                                    FsEntry en;
                                    en.val = binary ? 1.f : __ldg(p.x + (bit0 + bp));
                                    en.base = base + (bp << 2);  // byte offset inside the plane
                                    lists[pos] = en;
                                }
                                ++pos;
                            }
                        }
                    }
                    if (warp == 0 && lane < p.HZ) {
                        // odd rows end with a no-op entry (value 0 at the accumulators of halo voxel (0, 0): inside the
                        // warp's own plane for every tap)
                        const int n = gstart[lane + 1] - gstart[lane];
                        const int pos = pstart[lane] + n - lo;
                        if ((n & 1) && pos >= 0 && pos < kFoCap) {
                            FsEntry en;
                            en.val = 0.f;
                            en.base = ((p.kx - 1) * p.AS + (p.ky - 1)) << 2;
                            lists[pos] = en;
                        }
                    }
                    __syncthreads();  // (4) the list is complete
                }
                // ---- B: warp zo adds slice dz of the taps at every listed voxel of z-row zo + dz
                {
                    int s0 = pstart[warp] - lo;
                    for (int dz = 0; dz < p.kz; ++dz) {
                        int s1 = pstart[warp + dz + 1] - lo;  // even positions
                        const int e0 = s0 < 0 ? 0 : s0, e1 = s1 > kFoCap ? kFoCap : s1;
                        s0 = s1;
                        if (e0 >= e1) continue;
                        const uint32_t la = lists_u + (uint32_t)e0 * 8u, lend = lists_u + (uint32_t)e1 * 8u;
                        if constexpr (NI2 == 1) {
                            // the whole pair loop by hand (ncu: the compiler's version was 21 instructions per pair — register
                            // moves of a prefetch, uniform-datapath detours; this one is 13): two entries per 16-byte broadcast
                            // load, predicated read-modify-writes (the warp stays converged), a warp barrier after each: the next
                            // entry may own the same accumulator from another lane, and only the barrier orders the two accesses
                            // under the PTX memory model
                            const float kk = okp[0] ? skl[q * T + dz * P] : 0.f;
                            asm volatile(
                                "{\n\t.reg .pred q, more;\n\t.reg .f32 v0, v1, t, u;\n\t.reg .b32 b0, b1, a;\n\t"
                                "setp.ne.s32 q, %3, 0;\n\t"
                                "mov.u32 a, %0;\n\t"
                                "FO_PAIR:\n\t"
                                "ld.shared.v4.b32 {v0, b0, v1, b1}, [a];\n\t"
                                "add.u32 a, a, 16;\n\t"
                                "add.u32 b0, b0, %2;\n\t"
                                "add.u32 b1, b1, %2;\n\t"
                                "setp.ne.u32 more, a, %1;\n\t"
                                "@q ld.shared.f32 t, [b0];\n\t"
                                "@q fma.rn.f32 t, v0, %4, t;\n\t"
                                "@q st.shared.f32 [b0], t;\n\t"
                                FO_ORDER_PTX
                                "@q ld.shared.f32 u, [b1];\n\t"
                                "@q fma.rn.f32 u, v1, %4, u;\n\t"
                                "@q st.shared.f32 [b1], u;\n\t"
                                FO_ORDER_PTX
                                "@more bra FO_PAIR;\n\t}" ::"r"(la),
                                "r"(lend), "r"(accl[0]), "r"(okp[0]), "f"(kk)
                                : "memory");
                        } else {
                            float kk[NI2];
#pragma unroll
                            for (int j = 0; j < NI2; ++j) kk[j] = okp[j] ? skl[q * T + dz * P + 32 * j] : 0.f;
#pragma unroll 1
                            for (uint32_t a = la; a != lend; a += 16u) {
                                float v0, v1;
                                uint32_t b0, b1;
                                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=f"(v0), "=r"(b0), "=f"(v1), "=r"(b1) : "r"(a) : "memory");
#pragma unroll
                                for (int j = 0; j < NI2; ++j) fo_rmw(accl[j] + b0, v0, kk[j], okp[j]);
                                FO_ORDER();
#pragma unroll
                                for (int j = 0; j < NI2; ++j) fo_rmw(accl[j] + b1, v1, kk[j], okp[j]);
                                FO_ORDER();
                            }
                        }
                    }
                }
            }
            // ---- C: epilogue of this warp's plane: lanes <-> consecutive y (conflict-free shared loads, coalesced stores);
            // every accumulator of the interior is zeroed as it is read (rows / columns outside the grid included: they
            // collect contributions too).  ncu of the first version of this loop: 150 instructions per 32 outputs (64-bit
            // index arithmetic and parameter loads per iteration); now the pointers advance by a row per step.
            __syncwarp();
            {
                const int gz = z0 + warp;
                const int nyh = p.IY >> 5;  // 32-column chunks per output row
                float* a = accp + (p.kx - 1) * p.AS + (p.ky - 1) + lane;
                const int nrows = gz < p.Z ? min(p.IX, p.X - x0) : 0;  // rows that are stored
                const size_t idx0 = (size_t)q * (size_t)p.pred_qstride + (((size_t)b * p.Z + (gz < p.Z ? gz : 0)) * p.X + x0) * p.Y + y0 + lane;
                const bool in0 = y0 + lane < p.Y, in1 = nyh == 2 && y0 + 32 + lane < p.Y;
                if constexpr (OUT64) {
                    double* out = reinterpret_cast<double*>(p.pred) + idx0;
#pragma unroll 1
                    for (int xo = 0; xo < p.IX; ++xo, a += p.AS, out += p.Y) {
                        const float sa = a[0];
                        a[0] = 0.f;
                        float sb = 0.f;
                        if (nyh == 2) { sb = a[32]; a[32] = 0.f; }
                        if (xo < nrows) {  // warp-uniform
                            const double oa = tanh_pos_f64_tab(sa, tab), ob = tanh_pos_f64_tab(sb, tab);  // relu inside
                            if (in0) out[0] = oa;
                            if (in1) out[32] = ob;
                        }
                    }
                } else {
                    float* out = reinterpret_cast<float*>(p.pred) + idx0;
#pragma unroll 1
                    for (int xo = 0; xo < p.IX; ++xo, a += p.AS, out += p.Y) {
                        const float sa = a[0];
                        a[0] = 0.f;
                        float sb = 0.f;
                        if (nyh == 2) { sb = a[32]; a[32] = 0.f; }
                        if (xo < nrows) {
                            if (in0) out[0] = sa > 0.f ? tanhf(sa) : 0.f;
                            if (in1) out[32] = sb > 0.f ? tanhf(sb) : 0.f;
                        }
                    }
                }
            }
            __syncwarp();  // the zeroed accumulators are visible to every lane before the next observer's scatter
        }  // q
        __syncthreads();  // the list, the row tables and ctl are rewritten by the next tile
    }
}

// geometry; false when the kernel does not cover the shape (caller uses the dense stencil)
static bool plan_fwd_sparse(int B, int Z, int X, int Y, int kz, int kx, int ky, FsParams& p, size_t& smem, int& ni2, int nq = 1,
                            size_t* smem_occ = nullptr) {
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
    // (the scanning kernel zeroes planes with 16-byte stores: the plane size must be a multiple of 4 floats)
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
    if (smem_occ) *smem_occ = accb + (size_t)(kFoCap + 2) * sizeof(FsEntry) + 64 * 8 + taps + (32 + 32 + 8 + 8) * 4 + 16;
    return smem <= 227 * 1024;
}

// the mask-driven kernel: occupancy bits, at most kFoMaxCells cells per thread, field widths of the packed cell word
// (2-bit chunk, 6-bit x-row, z-rows / row tables up to 31) and 32-bit flat voxel indices (incl. the halo overshoot)
static bool occ_kernel_covers(const FsParams& p) {
    const int nwc = (p.IY + p.ky - 1 + 31) >> 5;
    return p.mask && p.HZ <= 31 && p.HX <= 63 && nwc <= 3 && p.HZ * p.HX * nwc <= kFoMaxCells * kFsThreads && p.IX * p.IY == 512 &&
           (size_t)p.RP * p.AS * 4 < 65536 && (long long)p.B * p.Z * p.X * p.Y < (1ll << 31) - (1ll << 20);
}

bool fwd_sparse_supported(int B, int Z, int X, int Y, int kz, int kx, int ky) {
    FsParams p{};
    size_t smem;
    int ni2;
    return plan_fwd_sparse(B, Z, X, Y, kz, kx, ky, p, smem, ni2);
}

// true when the mask-driven kernel can hand locally dense tiles of this shape to the dense stencil (same tiling:
// 8 x (512 / IY) x IY output tiles, and the state buffer's tile list holds every tile)
bool fwd_tile_handoff_supported(int B, int Z, int X, int Y, int kz, int kx, int ky) {
    FsParams p{};
    size_t smem;
    int ni2;
    if (!plan_fwd_sparse(B, Z, X, Y, kz, kx, ky, p, smem, ni2)) return false;
    p.mask = reinterpret_cast<const unsigned*>(&p);  // any non-null value: the geometry test only
    return occ_kernel_covers(p) && (long long)p.ntiles <= state_tile_cap((long long)B * Z * X * Y);
}

template <int NI2>
static int launch_fs(FsParams& p, size_t smem, size_t smem_occ, cudaStream_t stream) {
    const bool occ = occ_kernel_covers(p);
    if (p.nq > 1 && !occ) return SN_ERR_UNSUPPORTED;  // only the mask-driven kernel shares its lists between observers
    if (!occ) { p.dense_list = nullptr; p.state = nullptr; }
    const int nwc = (p.IY + p.ky - 1 + 31) >> 5;
    const bool small = ceil_div(p.HZ * p.HX * nwc, kFsThreads) <= 3;  // e.g. (9,5,5): 16 x 12 x 3 cells = 2.25 per thread
    auto kern = !occ ? fwd_sparse_kernel<NI2>
                : p.nq > 1 ? (p.out_f64 ? (small ? fwd_occ_kernel<NI2, true, true, 3> : fwd_occ_kernel<NI2, true, true, kFoMaxCells>)
                                        : (small ? fwd_occ_kernel<NI2, false, true, 3> : fwd_occ_kernel<NI2, false, true, kFoMaxCells>))
                           : (p.out_f64 ? (small ? fwd_occ_kernel<NI2, true, false, 3> : fwd_occ_kernel<NI2, true, false, kFoMaxCells>)
                                        : (small ? fwd_occ_kernel<NI2, false, false, 3> : fwd_occ_kernel<NI2, false, false, kFoMaxCells>));
    if (occ) smem = smem_occ;
    if (smem > 227 * 1024) return SN_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_rc(e);
    const int cap_sm = occ ? (NI2 == 1 ? SN_FO_CTAS : (NI2 == 2 ? 4 : 3)) : 4;
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > cap_sm ? cap_sm : per_sm);
    const int grid = max(1, min(p.ntiles, kNumSMs * per_sm));
    kern<<<grid, kFsThreads, smem, stream>>>(p);
    SN_LAUNCH_CHECK();
    return SN_OK;
}

// state: the grid state buffer (NULL: scanning kernel, no mask); gate: whole-grid selection against the dense stencil
// (NULL: run); handoff: append tiles above the break-even occupancy to the state buffer's tile list instead of
// scattering them (the caller enqueues the dense stencil's tile-list pass behind this launch)
int fwd_sparse_launch(const float* x, const float* Kstar, void* pred, int out_f64, unsigned long long* state,
                      const unsigned long long* gate, unsigned long long nnz_max, unsigned long long dw_max, bool handoff,
                      int B, int Z, int X, int Y, int kz, int kx, int ky, int nq, cudaStream_t stream) {
    FsParams p{};
    size_t smem, smem_occ;
    int ni2;
    if (nq < 1 || !plan_fwd_sparse(B, Z, X, Y, kz, kx, ky, p, smem, ni2, nq, &smem_occ)) return SN_ERR_UNSUPPORTED;
    p.nq = nq;
    p.pred_qstride = (long long)B * Z * X * Y;
    p.x = x; p.Kstar = Kstar; p.pred = pred; p.out_f64 = out_f64; p.nnz = gate; p.nnz_max = nnz_max; p.dw_max = dw_max;
    p.tanh64 = out_f64;  // float64 predictions: tanh evaluated in float64
    {
        static const bool no_mask = SN_ENV("SN_FWD_NO_MASK") != nullptr;  // measurement: force the scanning kernel
        const long long nvox = (long long)B * Z * X * Y;
        p.state = no_mask ? nullptr : state;
        p.mask = p.state ? reinterpret_cast<const unsigned*>(p.state + SN_STATE_WORDS) : nullptr;
        const long long nw = (nvox + 31) >> 5;
        p.nw = nw < (1ll << 30) ? (int)nw : 0;
        p.dense_list = nullptr;
        if (p.state && handoff && (long long)p.ntiles <= state_tile_cap(nvox)) {
            p.dense_list = reinterpret_cast<int*>(const_cast<unsigned*>(p.mask) + state_mask_words(nvox));
            // break-even of the scatter against the dense stencil, per tile: measured on uniform grids (CUDA-graph replay,
            // (9,5,5), float64 predictions): scatter 49 / 67 / 92 / 131 us at 1.6 / 3 / 5 / 8 % against 95.6 us for the
            // stencil at any occupancy — crossing at ~5.3 % of the halo box (profiles/r2_notes.md); wider slices move more
            // taps per listed voxel in the same instructions, so their crossing is higher
            static const double forced = SN_ENV("SN_FWD_TILE_PCT") ? atof(SN_ENV("SN_FWD_TILE_PCT")) : -1.0;
            const double pct = forced >= 0.0 ? forced : (kx * ky <= 32 ? 5.5 : 7.0);
            p.dense_thresh = (int)((double)p.HZ * p.HX * (p.IY + ky - 1) * pct / 100.0);
        }
    }
    if ((uintptr_t)x & 15) return SN_ERR_ALIGN;
    switch (ni2) {
        case 1: return launch_fs<1>(p, smem, smem_occ, stream);
        case 2: return launch_fs<2>(p, smem, smem_occ, stream);
        case 3: return launch_fs<3>(p, smem, smem_occ, stream);
        default: return launch_fs<4>(p, smem, smem_occ, stream);
    }
}

}  // namespace sn
