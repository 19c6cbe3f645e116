// Occupancy-driven tap gradient: W[t] = sum_{b,u : x[b,u] != 0} x[b,u] * G0[b, u - t + pad]
//
// Voxel grids of point clouds are almost empty (TS40K: 1.6 % occupied), so the dense tap-gradient
// stencil (stencil_bwd_impl.cuh) spends > 98 % of its FFMAs multiplying by zero.  This kernel walks the
// NON-ZERO voxels of x instead.  The T accumulators are spread over the lanes of a warp (lane l owns taps
// l, l+32, ... of its 256-tap chunk, in registers), so one non-zero voxel costs T/32 shared-memory loads +
// FFMAs with all 32 lanes busy: 2.5 instructions per 32 multiply-adds whatever the occupancy — cheaper than the
// dense stencil up to ~40 % occupancy, 15-20x cheaper at 1.6 %.  The kernel is then bound by staging the tiles
// (L2 -> shared memory), not by arithmetic.
//
// Per tile (8 x IX x IY voxels): TMA box of x (interior) + TMA box of G0 (interior + kernel halo, out-of-
// bounds = 0 = the convolution's zero padding) through an nstage full/empty mbarrier ring fed by a producer
// warp; the 16 compute warps never meet at a block barrier.  Warp (chunk c, slice s) scans slice s of the x tile
// with 16-byte loads, ballots the non-zero lanes and handles them one after the other (warp-uniform control).
// Deterministic: fixed tile -> CTA, slice -> warp and scan orders, float32 partial sums flushed to float64 every
// kSpFlush non-zeros, fixed-order combination at the end, no floating-point atomics.
//
// Which kernel runs (this one or the dense stencil) is decided ON THE DEVICE from the non-zero count written by
// sn_grid_prepare: both are launched, the one that is not selected returns at once — no host synchronisation,
// CUDA-graph friendly.
//
// Round 2: with the grid state buffer and a BINARY grid (state[4] == 0: every non-zero voxel is 1) the voxels come from the
// occupancy bits of the state buffer — no x tile is staged or scanned, a stage is the G0 box alone (one more stage in
// flight).  The kernel is bound by the ~3 us a TMA box takes to land times the stages in flight, not by bytes.
#include <stdlib.h>
#include <map>
#include <mutex>
#include <tuple>
#include "stencil_common.cuh"
#include "tma_host.cuh"

namespace sn {

constexpr int kSpNI = 8;               // taps per lane
constexpr int kSpChunk = 32 * kSpNI;   // taps per warp
constexpr int kSpWarps = 16;           // compute warps (+ 1 producer warp)
constexpr int kSpThreads = (kSpWarps + 1) * 32;
constexpr int kSpFlush = 128;          // non-zeros between float32 -> float64 flushes
constexpr int kSpMaxStages = 4;

struct SpParams {
    const float* x;
    const float* g0;
    double* partial;                 // [gridDim.x][TP]
    double* W;                       // [T]: written by the last CTA when `ticket` is given
    unsigned long long* ticket;      // device counter zeroed by sn_grid_prepare (NULL: a separate kernel sums the rows)
    const unsigned long long* nnz;   // device; NULL = always run
    unsigned long long nnz_max;      // run iff *nnz <= nnz_max
    const unsigned long long* state; // grid state buffer of x (sn_grid_prepare), or NULL
    const unsigned* mask;            // its occupancy bits when the tiles are made of whole mask words, else NULL
    int B, Z, X, Y, kz, kx, ky;
    int IX, IY, lgIX, lgIY;          // interior tile (z extent kRZ); powers of two
    int HZ, HX, WS;                  // G0 box
    int tiles_z, tiles_x, tiles_y, ntiles;
    int prz, prx, pra;               // box start = tile origin - (prz, prx, pra); pra = round4(right pad y)
    int ybase;                       // y offset of voxel (.,.,0) + left pad inside a box row
    int nchunks, S, nactive, nstage, nstage_bits, use_tma, TP;
};

__device__ __forceinline__ void sp_decode_tile(int tile, const SpParams& p, int& b, int& z0, int& x0, int& y0) {
    const int ty = tile % p.tiles_y;
    tile /= p.tiles_y;
    const int tx = tile % p.tiles_x;
    tile /= p.tiles_x;
    const int tz = tile % p.tiles_z;
    b = tile / p.tiles_z;
    z0 = tz * kRZ;
    x0 = tx * p.IX;
    y0 = ty * p.IY;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kSpThreads, 1)
tapgrad_sparse_kernel(const SpParams p, const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap gmap) {
    if (p.nnz && *p.nnz > p.nnz_max) return;  // dense input: stencil_bwd_kernel does the work
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int xi_floats = kRZ * p.IX * p.IY;
    const int halo_floats = p.HZ * p.HX * p.WS;
    // Binary grids (every non-zero voxel is 1: state[4] == 0) with a state buffer: the non-zero voxels come from the
    // occupancy BITS — the x tile is neither staged nor scanned, and a stage is the G0 box alone (one more stage in flight)
    const bool bits = p.mask != nullptr && p.state[4] == 0ull;
    const int xoff = bits ? 0 : xi_floats;  // the G0 box inside a stage
    const int stage_floats = xoff + ((halo_floats + 31) & ~31);
    const int nstage = bits ? p.nstage_bits : p.nstage;
    float* s0 = reinterpret_cast<float*>(smem_raw);
    const int data_floats = max(max(p.nstage * (xi_floats + ((halo_floats + 31) & ~31)), p.nstage_bits * ((halo_floats + 31) & ~31)),
                                kSpWarps * kSpChunk * 2);
    uint64_t* full = reinterpret_cast<uint64_t*>(s0 + data_floats);  // [kSpMaxStages]
    uint64_t* empty = full + kSpMaxStages;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x;
    const int zs = p.HX * p.WS;
    const int T = p.kz * p.kx * p.ky;

    if (p.use_tma) {
        if (tid == 0) {
            for (int i = 0; i < kSpMaxStages; ++i) {
                mbar_init(&full[i], 1);
                mbar_init(&empty[i], p.nactive);
            }
            fence_barrier_init();
        }
        __syncthreads();
    }

    const bool compute = warp < p.nactive;
    const int c = compute ? warp / p.S : 0, s = compute ? warp % p.S : 0;
    // noff[i]: box offset of tap t_i relative to a voxel's own box position, >= 0:
    // voxel (uz,ux,uy), tap (dz,dx,dy) -> box element (uz + kz-1 - dz, ux + kx-1 - dx, uy + ybase - dy)
    uint32_t noff[kSpNI];  // in bytes
    float acc[kSpNI];
    double accd[kSpNI];
#pragma unroll
    for (int i = 0; i < kSpNI; ++i) {
        const int t = c * kSpChunk + 32 * i + lane;
        const int tt = t < T ? t : 0;  // dead lanes read a valid address; their sums are dropped at the end
        const int dy = tt % p.ky, dx = (tt / p.ky) % p.kx, dz = tt / (p.ky * p.kx);
        noff[i] = 4u * (uint32_t)((p.kz - 1 - dz) * zs + (p.kx - 1 - dx) * p.WS + p.ybase - dy);
        acc[i] = 0.f;
        accd[i] = 0.0;
    }
    int pending = 0;
    const int n4 = xi_floats >> 2;
    const int lgXY = p.lgIY + p.lgIX, mY = p.IY - 1, mX = p.IX - 1;

    // all lanes handle the non-zero voxels flagged in `bal` one after the other (warp-uniform control flow):
    // lane l's value is broadcast, every lane adds its taps' products
    // (32-bit shared-window addresses: one 3-input add per load)
    auto consume = [&](unsigned bal, float val, int e_base, uint32_t sg) {
        while (bal) {
            const int l = __ffs(bal) - 1;
            bal &= bal - 1;
            const float xu = __shfl_sync(0xffffffffu, val, l);
            const int e = e_base + (l << 2);  // tile-linear voxel index
            const uint32_t gp = sg + 4u * (uint32_t)((e >> lgXY) * zs + ((e >> p.lgIY) & mX) * p.WS + (e & mY));
#pragma unroll
            for (int i = 0; i < kSpNI; ++i) {
                float g;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(g) : "r"(gp + noff[i]));
                acc[i] = fmaf(xu, g, acc[i]);
            }
            if (++pending == kSpFlush) {
#pragma unroll
                for (int i = 0; i < kSpNI; ++i) {
                    accd[i] += (double)acc[i];
                    acc[i] = 0.f;
                }
                pending = 0;
            }
        }
    };

    // occupancy bits: slice s takes the words s, s + S, ... of the tile (32 voxels of one row each); lane l holds word
    // s + S * l, loaded one tile ahead; the tile coordinates advance by G tiles per step without divisions
    const int widx = s + p.S * lane;
    const bool wmine = compute && widx < (xi_floats >> 5);
    const int we0 = widx << 5;  // tile-linear index of the word's first voxel
    const int wy = we0 & mY, wx = (we0 >> p.lgIY) & mX, wz = we0 >> lgXY;
    int cty, ctx, ctz, cb;     // tile of the next word load (ty, tx, tz, b)
    int gty, gtx, gtz, gb;     // G tiles as (ty, tx, tz, b) steps
    {
        int t = blockIdx.x;
        cty = t % p.tiles_y; t /= p.tiles_y; ctx = t % p.tiles_x; t /= p.tiles_x; ctz = t % p.tiles_z; cb = t / p.tiles_z;
        t = G;
        gty = t % p.tiles_y; t /= p.tiles_y; gtx = t % p.tiles_x; t /= p.tiles_x; gtz = t % p.tiles_z; gb = t / p.tiles_z;
    }
    auto load_word = [&]() -> unsigned {  // the word of tile (cty, ctx, ctz, cb); then advance by G tiles
        unsigned w = 0u;
        const int gz = ctz * kRZ + wz, gx = ctx * p.IX + wx, gy = cty * p.IY + wy;
        if (wmine && cb < p.B && gz < p.Z && gx < p.X && gy < p.Y)
            w = __ldg(p.mask + (((((long long)cb * p.Z + gz) * p.X + gx) * p.Y + gy) >> 5));
        cty += gty;
        int carry = cty >= p.tiles_y ? 1 : 0;
        cty -= carry * p.tiles_y;
        ctx += gtx + carry;
        carry = ctx >= p.tiles_x ? 1 : 0;
        ctx -= carry * p.tiles_x;
        ctz += gtz + carry;
        carry = ctz >= p.tiles_z ? 1 : 0;
        ctz -= carry * p.tiles_z;
        cb += gb + carry;
        return w;
    };
    auto scan_words = [&](unsigned word, const float* sgp) {
        const uint32_t sg = smem_u32(sgp);
        unsigned have = __ballot_sync(0xffffffffu, word != 0u);
        while (have) {
            const int l = __ffs(have) - 1;
            have &= have - 1u;
            unsigned w = __shfl_sync(0xffffffffu, word, l);
            const int e0 = (s + p.S * l) << 5;
            const uint32_t gp0 = sg + 4u * (uint32_t)((e0 >> lgXY) * zs + ((e0 >> p.lgIY) & mX) * p.WS + (e0 & mY));
            while (w) {
                const uint32_t gp = gp0 + 4u * (uint32_t)(__ffs(w) - 1);
                w &= w - 1u;
#pragma unroll
                for (int i = 0; i < kSpNI; ++i) {
                    float g;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(g) : "r"(gp + noff[i]));
                    acc[i] += g;  // x = 1
                }
                if (++pending == kSpFlush) {
#pragma unroll
                    for (int i = 0; i < kSpNI; ++i) {
                        accd[i] += (double)acc[i];
                        acc[i] = 0.f;
                    }
                    pending = 0;
                }
            }
        }
    };

    auto scan_tile = [&](const float* sx, const float* sgp) {
        const uint32_t sg = smem_u32(sgp);
        const float4* sx4 = reinterpret_cast<const float4*>(sx);
        for (int j0 = s * 32; j0 < n4; j0 += p.S * 32) {
            const int j = j0 + lane;
            const float4 v = j < n4 ? sx4[j] : make_float4(0.f, 0.f, 0.f, 0.f);
            const bool any = (v.x != 0.f) | (v.y != 0.f) | (v.z != 0.f) | (v.w != 0.f);
            if (__ballot_sync(0xffffffffu, any) == 0u) continue;
            const int e0 = j0 << 2;
            consume(__ballot_sync(0xffffffffu, v.x != 0.f), v.x, e0, sg);
            consume(__ballot_sync(0xffffffffu, v.y != 0.f), v.y, e0 + 1, sg);
            consume(__ballot_sync(0xffffffffu, v.z != 0.f), v.z, e0 + 2, sg);
            consume(__ballot_sync(0xffffffffu, v.w != 0.f), v.w, e0 + 3, sg);
        }
    };

    if (p.use_tma) {
        if (warp == kSpWarps) {
            if (lane == 0) {
                int k = 0;
                for (int tile = blockIdx.x; tile < p.ntiles; tile += G, ++k) {
                    const int st = k % nstage;
                    if (k >= nstage) {
                        mbar_wait(&empty[st], (uint32_t)(k / nstage - 1) & 1u);
                        fence_proxy_async();
                    }
                    int b, z0, x0, y0;
                    sp_decode_tile(tile, p, b, z0, x0, y0);
                    float* sx = s0 + st * stage_floats;
                    mbar_arrive_expect_tx(&full[st], (uint32_t)(xoff + halo_floats) * 4u);
                    if (!bits) tma_load_4d(sx, &xmap, &full[st], y0, x0, z0, b);
                    tma_load_4d(sx + xoff, &gmap, &full[st], y0 - p.pra, x0 - p.prx, z0 - p.prz, b);
                }
            }
        } else if (compute) {
            int k = 0;
            unsigned wnext = bits ? load_word() : 0u;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += G, ++k) {
                const int st = k % nstage;
                const unsigned word = wnext;
                if (bits) wnext = load_word();
                mbar_wait(&full[st], (uint32_t)(k / nstage) & 1u);
                const float* sx = s0 + st * stage_floats;
                if (bits)
                    scan_words(word, sx);
                else
                    scan_tile(sx, sx + xi_floats);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[st]);
            }
        }
    } else {
        // plain-load path (odd Y, unaligned base, no driver entry point): block barriers, one stage
        for (int tile = blockIdx.x; tile < p.ntiles; tile += G) {
            int b, z0, x0, y0;
            sp_decode_tile(tile, p, b, z0, x0, y0);
            __syncthreads();
            float* sx = s0;
            float* sg = s0 + xi_floats;
            for (int i = tid; i < xi_floats; i += kSpThreads) {
                const int yy = i & (p.IY - 1), xx = (i >> p.lgIY) & (p.IX - 1), zz = i >> (p.lgIY + p.lgIX);
                const int gz = z0 + zz, gx = x0 + xx, gy = y0 + yy;
                sx[i] = (gz < p.Z && gx < p.X && gy < p.Y) ? __ldg(p.x + (((size_t)b * p.Z + gz) * p.X + gx) * p.Y + gy) : 0.f;
            }
            for (int i = tid; i < halo_floats; i += kSpThreads) {
                const int cc = i % p.WS, r = i / p.WS;
                const int gz = z0 - p.prz + r / p.HX, gx = x0 - p.prx + r % p.HX, gy = y0 - p.pra + cc;
                const bool ok = gz >= 0 && gz < p.Z && gx >= 0 && gx < p.X && gy >= 0 && gy < p.Y;
                sg[i] = ok ? __ldg(p.g0 + (((size_t)b * p.Z + gz) * p.X + gx) * p.Y + gy) : 0.f;
            }
            __syncthreads();
            if (compute) scan_tile(sx, sg);
        }
    }

    // fixed-order combination of the S slices of every chunk: one partial row per CTA
    __syncthreads();
    double* sred = reinterpret_cast<double*>(smem_raw);  // [kSpWarps][kSpChunk]; the stages are dead by now
    if (compute) {
#pragma unroll
        for (int i = 0; i < kSpNI; ++i) sred[warp * kSpChunk + 32 * i + lane] = accd[i] + (double)acc[i];
    }
    __syncthreads();
    double* row = p.partial + (size_t)blockIdx.x * p.TP;
    for (int t = tid; t < T; t += kSpThreads) {
        const int cc = t / kSpChunk, tl = t % kSpChunk;
        double a = 0.0;
        for (int q = 0; q < p.S; ++q) a += sred[(cc * p.S + q) * kSpChunk + tl];
        row[t] = a;
    }
    if (p.ticket) {
        __syncthreads();  // sred is reused as scratch
        last_cta_row_sum(p.ticket, p.partial, (int)gridDim.x, p.TP, T, p.W, sred);
    }
}

// Shared-memory wavefronts one non-zero voxel costs: lanes read G0 box elements at -(dz*zs + dx*WS + dy), and a
// TMA box row is a multiple of 16 bytes, so bank = f(zs mod 32, WS mod 32, dy) and taps collide.  The box is
// padded by up to 3 rows in x / 12 floats in y to the strides with the fewest wavefronts ((9,5,5): 21 -> 15;
// 8 would be conflict-free).  The answer is cached per shape (a pure function of its key).
static int sp_wavefronts(int kz, int kx, int ky, int zs, int WS) {
    const int T = kz * kx * ky;
    int total = 0;
    for (int t0 = 0; t0 < T; t0 += 32) {
        int cnt[32] = {0}, mx = 0;
        for (int t = t0; t < T && t < t0 + 32; ++t) {
            const int dy = t % ky, dx = (t / ky) % kx, dz = t / (ky * kx);
            const int b = (dz * zs + dx * WS + dy) & 31;
            if (++cnt[b] > mx) mx = cnt[b];
        }
        total += mx;
    }
    return total;
}

static void sp_pick_padding(int kz, int kx, int ky, int HX, int WS, int HZ, size_t xi_bytes, int& HXp, int& WSp) {
    static std::mutex mu;
    static std::map<std::tuple<int, int, int, int, int>, std::pair<int, int>> cache;
    const auto key = std::make_tuple(kz, kx, ky, HX, WS);
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it != cache.end()) {
            HXp = it->second.first;
            WSp = it->second.second;
            return;
        }
    }
    int best = 1 << 30;
    HXp = HX; WSp = WS;
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 16; b += 4) {
            const int hx = HX + a, ws = WS + b;
            if (ws > 256 || hx > 256) continue;
            if ((size_t)HZ * hx * ws * 4 + xi_bytes + 1024 > 227 * 1024) continue;  // must still fit one stage
            const int w = sp_wavefronts(kz, kx, ky, hx * ws, ws) * 64 + (hx * ws - HX * WS) * 64 / (HX * WS);  // ties: least padding
            if (w < best) { best = w; HXp = hx; WSp = ws; }
        }
    std::lock_guard<std::mutex> lk(mu);
    cache[key] = std::make_pair(HXp, WSp);
}

static inline int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

// geometry + launch plan; returns false when the kernel cannot handle the shape (caller uses the dense path)
static bool plan_sparse(int B, int Z, int X, int Y, int kz, int kx, int ky, SpParams& p, size_t& smem) {
    const int T = kz * kx * ky;
    p.B = B; p.Z = Z; p.X = X; p.Y = Y; p.kz = kz; p.kx = kx; p.ky = ky;
    p.nchunks = ceil_div(T, kSpChunk);
    if (p.nchunks > kSpWarps) return false;
    p.S = kSpWarps / p.nchunks;
    p.nactive = p.nchunks * p.S;
    p.IY = Y > 32 ? 64 : 32;
    p.IX = (kRZ * 64 * 8) / (kRZ * p.IY);  // 4096-voxel tiles: 8 x 8 x 64 or 8 x 16 x 32
    p.lgIX = ilog2(p.IX);
    p.lgIY = ilog2(p.IY);
    const int ply = pad_left(ky), pry = ky - 1 - ply;
    p.prz = kz - 1 - pad_left(kz);
    p.prx = kx - 1 - pad_left(kx);
    p.pra = round4(pry);
    p.ybase = p.pra + ply;  // voxel uy, tap dy -> box column uy + pra + ply - dy
    p.HZ = kRZ + kz - 1;
    p.HX = p.IX + kx - 1;
    p.WS = round4(p.IY + p.pra + ply);
    sp_pick_padding(kz, kx, ky, p.HX, p.WS, p.HZ, (size_t)kRZ * p.IX * p.IY * 4, p.HX, p.WS);
    p.tiles_z = ceil_div(Z, kRZ);
    p.tiles_x = ceil_div(X, p.IX);
    p.tiles_y = ceil_div(Y, p.IY);
    p.ntiles = B * p.tiles_z * p.tiles_x * p.tiles_y;
    p.TP = (T + 31) & ~31;
    const int xi = kRZ * p.IX * p.IY;
    const size_t stage = (size_t)(xi + ((p.HZ * p.HX * p.WS + 31) & ~31)) * 4;
    const size_t extra = 2 * kSpMaxStages * 8 + 128;
    int ns = (int)((227 * 1024 - extra) / stage);
    if (ns < 1) return false;
    p.nstage = ns > kSpMaxStages ? kSpMaxStages : ns;
    const size_t stage_bits = (size_t)((p.HZ * p.HX * p.WS + 31) & ~31) * 4;  // binary grids: the G0 box alone
    int nsb = (int)((227 * 1024 - extra) / stage_bits);
    p.nstage_bits = nsb > kSpMaxStages ? kSpMaxStages : nsb;
    const size_t red = (size_t)kSpWarps * kSpChunk * 8;
    size_t data = p.nstage * stage > p.nstage_bits * stage_bits ? p.nstage * stage : p.nstage_bits * stage_bits;
    smem = (data > red ? data : red) + extra;
    return true;
}

int64_t tapgrad_sparse_ws(int B, int Z, int X, int Y, int kz, int kx, int ky) {
    SpParams p{};
    size_t smem;
    if (!plan_sparse(B, Z, X, Y, kz, kx, ky, p, smem)) return 0;
    return (int64_t)min(p.ntiles, kNumSMs) * p.TP * 8;
}

// rows_out = partial rows written (0 when the sparse kernel is not applicable)
int tapgrad_sparse_launch(const float* x, const float* g0, const unsigned long long* nnz, unsigned long long nnz_max,
                          int B, int Z, int X, int Y, int kz, int kx, int ky, void* ws, int64_t ws_bytes, int* rows_out,
                          double* W, unsigned long long* ticket, cudaStream_t stream, const unsigned long long* state) {
    SpParams p{};
    size_t smem;
    *rows_out = 0;
    if (!plan_sparse(B, Z, X, Y, kz, kx, ky, p, smem)) return SN_OK;
    const int grid = min(p.ntiles, kNumSMs);
    if ((int64_t)grid * p.TP * 8 > ws_bytes) return SN_ERR_WORKSPACE;
    p.x = x; p.g0 = g0; p.nnz = nnz; p.nnz_max = nnz_max; p.W = W; p.ticket = ticket;
    p.partial = reinterpret_cast<double*>(ws);
    CUtensorMap xmap, gmap;
    const bool okx = make_grid_tmap(&xmap, x, B, Z, X, Y, kRZ, p.IX, p.IY);
    const bool okg = make_grid_tmap(&gmap, g0, B, Z, X, Y, p.HZ, p.HX, p.WS);
    p.use_tma = (okx && okg) ? 1 : 0;
    if (!p.use_tma) p.nstage = 1;
    // occupancy bits instead of the x tile: tiles made of whole mask words, at most 32 words per slice, TMA path
    p.state = state;
    const bool words_ok = state && p.use_tma && (Y & 31) == 0 && (p.IY & 31) == 0 && (kRZ * p.IX * p.IY >> 5) <= 32 * p.S;
    p.mask = words_ok ? reinterpret_cast<const unsigned*>(state + SN_STATE_WORDS) : nullptr;
    cudaError_t e = cudaFuncSetAttribute(tapgrad_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_rc(e);
    tapgrad_sparse_kernel<<<grid, kSpThreads, smem, stream>>>(p, xmap, gmap);
    SN_LAUNCH_CHECK();
    *rows_out = grid;
    return SN_OK;
}

}  // namespace sn
