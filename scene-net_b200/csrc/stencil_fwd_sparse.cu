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
    unsigned long long nnz_max;     // run iff *nnz <= nnz_max
    int B, Z, X, Y, kz, kx, ky;
    int out_f64, tanh64;
    int IX, IY;                 // output tile (z extent kRZ)
    int HZ, HX, WS;             // halo box
    int plz, plx, pla, off;     // box start = tile origin - (plz, plx, pla); off = pla - left pad y (dead columns)
    int AS, RP;                 // accumulator plane: row pitch (floats), rows per plane
    int tiles_z, tiles_x, tiles_y, ntiles;
    unsigned ws_magic;          // ceil(2^24 / WS): f / WS == (f * ws_magic) >> 24 for f < 2^13
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
    if (p.nnz && *p.nnz > p.nnz_max) return;  // dense input: stencil_fwd_kernel does the work
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

// geometry; false when the kernel does not cover the shape (caller uses the dense stencil)
static bool plan_fwd_sparse(int B, int Z, int X, int Y, int kz, int kx, int ky, FsParams& p, size_t& smem, int& ni2) {
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
    const size_t taps = (size_t)((T + 31) & ~31) * 4;
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
    auto kern = fwd_sparse_kernel<NI2>;
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
                      unsigned long long nnz_max, int B, int Z, int X, int Y, int kz, int kx, int ky, cudaStream_t stream) {
    FsParams p{};
    size_t smem;
    int ni2;
    if (!plan_fwd_sparse(B, Z, X, Y, kz, kx, ky, p, smem, ni2)) return SN_ERR_UNSUPPORTED;
    p.x = x; p.Kstar = Kstar; p.pred = pred; p.out_f64 = out_f64; p.nnz = nnz; p.nnz_max = nnz_max;
    p.tanh64 = out_f64;  // float64 predictions: tanh evaluated in float64 (tanh_pos_f64)
    if ((uintptr_t)x & 15) return SN_ERR_ALIGN;
    switch (ni2) {
        case 1: return launch_fs<1>(p, smem, stream);
        case 2: return launch_fs<2>(p, smem, stream);
        case 3: return launch_fs<3>(p, smem, stream);
        default: return launch_fs<4>(p, smem, stream);
    }
}

}  // namespace sn
