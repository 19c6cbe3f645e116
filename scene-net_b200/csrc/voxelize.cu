// Point-cloud voxelization: bounding box -> linspace edges -> searchsorted binning with
// warp-aggregated atomics -> count / keep-count / max-label grids -> density / fraction /
// occupancy.  Replaces (paths relative to the reference root)
//   utils/pcd_processing.py:341-372 (eda.voxelize_ply -> pyntcloud==0.1.6 VoxelGrid.compute)
//   utils/voxelization.py:164-204, 207-241, 244-300 (hist_on_voxel, classes_on_voxel, reg_on_voxel)
//   utils/pcd_processing.py:305-321 (normalize_xyz = sklearn MinMaxScaler per y column)
//   core/datasets/torch_transforms.py:33-40 (ToFullDense.densify)
//
// Everything that decides a voxel index is float64 with explicitly rounded operations
// (__dmul_rn/__dadd_rn/__ddiv_rn: no FMA contraction), because TS40K coordinates are UTM
// (~5e5, ~4.6e6 m) and a 1-ulp edge difference moves points across voxels.
// HBM-bound: 24 B/point (bounding box) + 32 B/point (bin) + the grid passes.
#include <math.h>
#include "common.cuh"

namespace sn {

constexpr int kVoxThreads = 256;
constexpr int kBinUnroll = 4;  // points per lane per step of the binning kernel
// Per-CTA shared-memory aggregation table of the binning kernel: scan-coherent clouds send runs of points to the
// same few thousand voxels, so a CTA that owns a CONTIGUOUS range of points first adds into a direct-mapped table
// (native shared-memory integer atomics) and sends one global atomic per occupied slot at the end; a point whose
// slot is taken by another voxel goes to global memory directly.  Integer adds: exact, order-independent.
constexpr int kBinSlots = 4096;

__device__ __forceinline__ void atomic_min_f64(double* a, double v) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(a);
    unsigned long long old = *p;
    while (v < __longlong_as_double((long long)old)) {
        const unsigned long long assumed = old;
        old = atomicCAS(p, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}
__device__ __forceinline__ void atomic_max_f64(double* a, double v) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(a);
    unsigned long long old = *p;
    while (v > __longlong_as_double((long long)old)) {
        const unsigned long long assumed = old;
        old = atomicCAS(p, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}

// monotone map double -> int64 (an involution), so atomicMax on integers orders doubles
__device__ __forceinline__ long long f64_key(double d) {
    const long long b = __double_as_longlong(d);
    return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double key_f64(long long k) { return __longlong_as_double(k ^ ((k >> 63) & 0x7fffffffffffffffLL)); }

// ------------------------------------------------------------------ step 1: bounding boxes
__global__ void minmax_init_kernel(double* mnmx, int n_clouds) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_clouds * 6) mnmx[i] = (i % 6) < 3 ? INFINITY : -INFINITY;
}

__global__ void __launch_bounds__(kVoxThreads)
minmax_kernel(const double* __restrict__ pts, int ld, const long long* __restrict__ offsets, long long n_total,
              double* __restrict__ mnmx) {
    __shared__ double red[6][kVoxThreads / 32];
    const int c = blockIdx.y;
    const long long beg = offsets ? offsets[c] : 0, end = offsets ? offsets[c + 1] : n_total;
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    const bool vec4 = (ld == 4) && ((reinterpret_cast<uintptr_t>(pts) & 15) == 0);
#pragma unroll 4
    for (long long i = beg + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (long long)gridDim.x * blockDim.x) {
        double v[3];
        if (vec4) {
            const double2 a = __ldg(reinterpret_cast<const double2*>(pts + i * 4));
            const double2 b = __ldg(reinterpret_cast<const double2*>(pts + i * 4) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = b.x;
        } else {
            const double* q = pts + i * ld;
            v[0] = __ldg(q); v[1] = __ldg(q + 1); v[2] = __ldg(q + 2);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            mn[k] = fmin(mn[k], v[k]);
            mx[k] = fmax(mx[k], v[k]);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fmin(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmax(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
        if (lane == 0) {
            red[k][warp] = mn[k];
            red[3 + k][warp] = mx[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        const int k = threadIdx.x;
        double v = red[k][0];
        for (int w = 1; w < kVoxThreads / 32; ++w) v = k < 3 ? fmin(v, red[k][w]) : fmax(v, red[k][w]);
        if (k < 3)
            atomic_min_f64(&mnmx[c * 6 + k], v);
        else
            atomic_max_f64(&mnmx[c * 6 + k], v);
    }
}

// ------------------------------------------------------------------ step 2: edges
// pyntcloud VoxelGrid.compute (regular_bounding_box=True) + np.linspace, op by op.
// all threads of the CTA: the nx+1, ny+1, nz+1 edges of one cloud from its bounding box b[6] into e[]
__device__ __forceinline__ void compute_edges(const double* __restrict__ b, int nx, int ny, int nz, double* __restrict__ e) {
    double rng[3], lo[3], hi[3];
    for (int k = 0; k < 3; ++k) rng[k] = __dsub_rn(b[3 + k], b[k]);  // ptp
    const double rmax = fmax(rng[0], fmax(rng[1], rng[2]));
    for (int k = 0; k < 3; ++k) {
        const double margin = __dsub_rn(rmax, rng[k]);
        const double half = __ddiv_rn(margin, 2.0);
        lo[k] = __dsub_rn(b[k], half);
        hi[k] = __dadd_rn(b[3 + k], half);
    }
    const int n[3] = {nx, ny, nz};
    const int base[3] = {0, nx + 1, nx + 1 + ny + 1};
    for (int k = 0; k < 3; ++k) {
        const double delta = __dsub_rn(hi[k], lo[k]);
        const double step = __ddiv_rn(delta, (double)n[k]);
        for (int j = threadIdx.x; j <= n[k]; j += blockDim.x) {
            double v;
            if (j == n[k])
                v = hi[k];
            else if (step == 0.0)
                v = __dadd_rn(__dmul_rn(__ddiv_rn((double)j, (double)n[k]), delta), lo[k]);
            else
                v = __dadd_rn(__dmul_rn((double)j, step), lo[k]);
            e[base[k] + j] = v;
        }
    }
}

__global__ void edges_kernel(const double* __restrict__ mnmx, int nx, int ny, int nz, double* __restrict__ edges) {
    const int c = blockIdx.x;
    compute_edges(mnmx + c * 6, nx, ny, nz, edges + (size_t)c * (nx + ny + nz + 3));
}

// one launch for every initialisation of the fused entry point: bounding boxes to +-inf, grids to 0 / lowest key
__global__ void vox_init_kernel(double* __restrict__ mnmx, int n_clouds, int* __restrict__ count, int* __restrict__ keep_count,
                                long long* __restrict__ maxkey, long long n) {
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = i0; i < (long long)n_clouds * 6; i += stride) mnmx[i] = (i % 6) < 3 ? INFINITY : -INFINITY;
    for (long long i = i0; i < n; i += stride) {
        count[i] = 0;
        if (keep_count) keep_count[i] = 0;
        if (maxkey) maxkey[i] = (long long)0x8000000000000000ULL;
    }
}

// ------------------------------------------------------------------ step 3: binning
__global__ void bin_init_kernel(int* __restrict__ count, int* __restrict__ keep_count, long long* __restrict__ maxkey, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        count[i] = 0;
        if (keep_count) keep_count[i] = 0;
        if (maxkey) maxkey[i] = (long long)0x8000000000000000ULL;
    }
}

// largest j with e[j] < p, or 0 if none; == clip(searchsorted(e, p, 'left') - 1, 0, n-1) for p <= e[n].
// The multiply gives the bin up to +-1 (rounding): the common case is decided by two edge compares without a
// loop; the loops only run when the guess was off.
__device__ __forceinline__ int bin_of(const double* __restrict__ e, int n, double p, double lo, double inv_step) {
    int j = (int)((p - lo) * inv_step);
    j = j < 0 ? 0 : (j > n - 1 ? n - 1 : j);
    const double ej = e[j], ej1 = e[j + 1];  // e has n + 1 entries
    const bool down = j > 0 && !(ej < p);
    const bool up = j + 1 <= n - 1 && ej1 < p;
    if (down) {
        --j;
        while (j > 0 && !(e[j] < p)) --j;
    } else if (up) {
        ++j;
        while (j + 1 <= n - 1 && e[j + 1] < p) ++j;
    }
    return j;
}

// The common case of bin_of without a branch: the multiply's guess j, its two edges, and whether the guess has to move
// (p not in (e[j], e[j+1]] — rounding next to an edge, or a degenerate axis).  The caller runs bin_of for those lanes.
__device__ __forceinline__ int bin_guess(const double* __restrict__ e, int n, double p, double lo, double inv_step, bool& off) {
    int j = __double2int_rz((p - lo) * inv_step);  // NaN -> 0
    j = min(max(j, 0), n - 1);
    const double ej = e[j], ej1 = e[j + 1];  // e has n + 1 entries
    off = (j > 0 && !(ej < p)) || (j + 1 <= n - 1 && ej1 < p);
    return j;
}

// FAST: the layout of the data path (TS40K rows x, y, z, label: ld == 4, 16-byte aligned, the label read from the row; keep votes
// wanted; no per-point index output, no max-label grid) as a compile-time fact — the general kernel carries every one of these
// launch-uniform choices as predicated instructions in its inner loop, and the kernel is bound by instruction issue.
template <bool FAST>
__global__ void __launch_bounds__(kVoxThreads, 4)  // 4 CTAs/SM: what the 48 KB aggregation table allows; caps the registers at 64
bin_kernel(const double* __restrict__ pts, int ld, const double* __restrict__ labels_, int label_ld,
           const long long* __restrict__ offsets, const double* __restrict__ edges, int nx, int ny, int nz,
           const double* __restrict__ keep, int n_keep, int* __restrict__ count, int* __restrict__ keep_count,
           long long* __restrict__ maxkey_, int* __restrict__ lin_out_, long long n_total, const double* __restrict__ mnmx,
           double* __restrict__ edges_out) {
    const double* labels = FAST ? pts + 3 : labels_;
    long long* maxkey = FAST ? nullptr : maxkey_;
    int* lin_out = FAST ? nullptr : lin_out_;
    if (FAST) ld = 4;
    extern __shared__ double s_edges[];  // (nx+1)+(ny+1)+(nz+1) doubles, n_keep keep labels, then the aggregation table
    const int c = blockIdx.y;
    const int ne = nx + ny + nz + 3;
    if (mnmx) {
        // fused entry point: every CTA derives its cloud's edges from the bounding box itself (3 x 65 values: cheaper
        // than a launch); the first CTA of the cloud also publishes them
        compute_edges(mnmx + c * 6, nx, ny, nz, s_edges);
        __syncthreads();
        if (blockIdx.x == 0 && edges_out)
            for (int i = threadIdx.x; i < ne; i += blockDim.x) edges_out[(size_t)c * ne + i] = s_edges[i];
    } else {
        for (int i = threadIdx.x; i < ne; i += blockDim.x) s_edges[i] = edges[(size_t)c * ne + i];
    }
    double* s_keep = s_edges + ne;
    for (int i = threadIdx.x; i < n_keep; i += blockDim.x) s_keep[i] = keep[i];
    int* s_key = reinterpret_cast<int*>(s_keep + n_keep + (n_keep & 1));  // table behind the doubles
    int* s_cnt = s_key + kBinSlots;
    int* s_kcnt = s_cnt + kBinSlots;
    for (int i = threadIdx.x; i < kBinSlots; i += blockDim.x) {
        s_key[i] = -1;
        s_cnt[i] = 0;
        s_kcnt[i] = 0;
    }
    __syncthreads();
    const double* ex = s_edges;
    const double* ey = s_edges + nx + 1;
    const double* ez = ey + ny + 1;
    const double lox = ex[0], loy = ey[0], loz = ez[0];
    const double ivx = (double)nx / (ex[nx] - lox), ivy = (double)ny / (ey[ny] - loy), ivz = (double)nz / (ez[nz] - loz);
    const long long cbeg = offsets ? offsets[c] : 0, cend = offsets ? offsets[c + 1] : n_total;
    // this CTA's contiguous range of the cloud's points (multiples of a warp step)
    const long long step = 32LL * kBinUnroll;
    const long long per = ceil_div64(ceil_div64(cend - cbeg, (long long)gridDim.x), step) * step;
    const long long beg = cbeg + (long long)blockIdx.x * per;
    const long long end = beg + per < cend ? beg + per : cend;
    const long long V = (long long)nx * ny * nz;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const bool vec4 = FAST || ((ld == 4) && ((reinterpret_cast<uintptr_t>(pts) & 15) == 0));
    const bool lab_in_row = FAST || (vec4 && labels == pts + 3 && label_ld == 4);
    int* count_c = count + (size_t)c * V;
    int* keep_c = (FAST || keep_count) ? keep_count + (size_t)c * V : nullptr;

    // kBinUnroll points per lane per step, all loads issued before the first use: with one point in flight per
    // thread the kernel was bound by the HBM latency of its own loads.  The kernel is bound by instruction issue (ncu:
    // 226 instructions per 32 points in the first version): full steps run without per-point predicates, the bin comes
    // from bin_guess (two edge compares, no loop) with bin_of as the rare exact fallback, the first keep label sits in a
    // register.
    const double keep0 = n_keep > 0 ? s_keep[0] : 0.0;
    for (long long base = beg + (long long)warp * step; base < end; base += (long long)nwarps * step) {
        const bool full = base + step <= end;  // warp-uniform
        double px[kBinUnroll], py[kBinUnroll], pz[kBinUnroll], lab[kBinUnroll];
        bool valid[kBinUnroll];
        const double* q0 = pts + (base + lane) * (long long)ld;
#pragma unroll
        for (int u = 0; u < kBinUnroll; ++u) {
            valid[u] = full || base + 32 * u + lane < end;
            px[u] = py[u] = pz[u] = lab[u] = 0.0;
            if (valid[u]) {
                if (vec4) {
                    const double2 a = __ldg(reinterpret_cast<const double2*>(q0 + 32 * 4 * u));
                    const double2 b = __ldg(reinterpret_cast<const double2*>(q0 + 32 * 4 * u) + 1);
                    px[u] = a.x; py[u] = a.y; pz[u] = b.x;
                    if (lab_in_row) lab[u] = b.y;
                } else {
                    const double* q = q0 + (long long)32 * u * ld;
                    px[u] = __ldg(q); py[u] = __ldg(q + 1); pz[u] = __ldg(q + 2);
                }
                if (labels && !lab_in_row) lab[u] = __ldg(labels + (base + 32 * u + lane) * label_ld);
            }
        }
#pragma unroll
        for (int u = 0; u < kBinUnroll; ++u) {
            const unsigned mask = full ? 0xffffffffu : __ballot_sync(0xffffffffu, valid[u]);
            if (!valid[u]) continue;
            // (computing the four voxels first and counting afterwards — more independent chains — measured slower: 131 vs 111 us)
            bool ox, oy, oz;
            int vx = bin_guess(ex, nx, px[u], lox, ivx, ox);
            int vy = bin_guess(ey, ny, py[u], loy, ivy, oy);
            int vz = bin_guess(ez, nz, pz[u], loz, ivz, oz);
            if (ox | oy | oz) {  // rare: a point within rounding of an edge (or a degenerate axis)
                if (ox) vx = bin_of(ex, nx, px[u], lox, ivx);
                if (oy) vy = bin_of(ey, ny, py[u], loy, ivy);
                if (oz) vz = bin_of(ez, nz, pz[u], loz, ivz);
            }
            const int lin = (vz * nx + vx) * ny + vy;  // reference grid layout data[z, x, y]
            if (lin_out) lin_out[base + 32 * u + lane] = lin;
            bool is_keep = false;
            if (labels) {
                is_keep = n_keep > 0 && lab[u] == keep0;
                for (int k = 1; k < n_keep; ++k) is_keep |= (lab[u] == s_keep[k]);
            }
            // warp-aggregated: one table / global update per distinct voxel per warp
            const unsigned peers = __match_any_sync(mask, lin);
            const unsigned kmask = __ballot_sync(mask, is_keep);
            const bool leader = (__ffs(peers) - 1) == lane;
            if (leader) {
                const int nc = __popc(peers), kc = __popc(peers & kmask);
                const int slot = (int)(((unsigned)lin * 2654435761u) >> (32 - 12));  // kBinSlots = 2^12
                const int old = atomicCAS(&s_key[slot], -1, lin);
                if (old == -1 || old == lin) {
                    atomicAdd(&s_cnt[slot], nc);
                    if (keep_c && kc) atomicAdd(&s_kcnt[slot], kc);
                } else {
                    atomicAdd(&count_c[lin], nc);
                    if (keep_c && kc) atomicAdd(&keep_c[lin], kc);
                }
            }
            if (maxkey && labels) {
                const long long g = (long long)c * V + lin;
                const long long k = f64_key(lab[u]);
                if (k > *reinterpret_cast<volatile long long*>(&maxkey[g])) atomicMax(&maxkey[g], k);  // cheap pre-check, most points lose
            }
        }
    }
    __syncthreads();
    for (int sl = threadIdx.x; sl < kBinSlots; sl += blockDim.x) {
        const int key = s_key[sl];
        if (key >= 0) {
            atomicAdd(&count_c[key], s_cnt[sl]);
            const int kc = s_kcnt[sl];
            if (keep_c && kc) atomicAdd(&keep_c[key], kc);
        }
    }
}

// ------------------------------------------------------------------ step 4: finalize
__global__ void colminmax_init_kernel(int* __restrict__ cm, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cm[i] = (i & 1) ? 0 : 0x7fffffff;  // [.., {min,max}]
}

// per-(cloud, y) min/max of count over all (z,x) rows; cm [C][ny][2]
__global__ void __launch_bounds__(kVoxThreads)
colminmax_kernel(const int* __restrict__ count, int rows, int ny, int rows_per_block, int* __restrict__ cm) {
    const int c = blockIdx.y;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    const int* base = count + (size_t)c * rows * ny;
    // consecutive threads -> consecutive y (coalesced row reads); spare threads take other rows
    const int ly = ny < (int)blockDim.x ? ny : (int)blockDim.x;
    const int groups = blockDim.x / ly, grp = threadIdx.x / ly;
    if (grp >= groups) return;
    for (int y = threadIdx.x % ly; y < ny; y += ly) {
        int mn = 0x7fffffff, mx = 0;
        for (int r = r0 + grp; r < r1; r += groups) {
            const int v = base[(size_t)r * ny + y];
            mn = min(mn, v);
            mx = max(mx, v);
        }
        if (mn <= mx) {
            atomicMin(&cm[((size_t)c * ny + y) * 2], mn);
            atomicMax(&cm[((size_t)c * ny + y) * 2 + 1], mx);
        }
    }
}

template <typename TO>
__global__ void __launch_bounds__(kVoxThreads)
finalize_kernel(const int* __restrict__ count, const int* __restrict__ keep_count, const int* __restrict__ cm, long long n,
                int ny, long long V, double* __restrict__ density, double* __restrict__ frac, double* __restrict__ max_label,
                TO* __restrict__ occ, TO* __restrict__ occ_keep) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int cnt = count[i];
        const int kc = keep_count ? keep_count[i] : 0;
        if (density) {
            // sklearn MinMaxScaler: scale = 1/range (1 if range == 0); X*scale + (0 - min*scale)
            const int y = (int)(i % ny);
            const long long c = i / V;
            const int mn = cm[(c * ny + y) * 2], mx = cm[(c * ny + y) * 2 + 1];
            double rng = (double)mx - (double)mn;
            if (rng == 0.0) rng = 1.0;
            const double scale = __ddiv_rn(1.0, rng);
            const double off = __dsub_rn(0.0, __dmul_rn((double)mn, scale));
            density[i] = __dadd_rn(__dmul_rn((double)cnt, scale), off);
        }
        if (frac) frac[i] = cnt > 0 ? __ddiv_rn((double)kc, (double)cnt) : 0.0;
        if (max_label) {
            const long long k = reinterpret_cast<const long long*>(max_label)[i];
            max_label[i] = cnt > 0 ? key_f64(k) : 0.0;
        }
        if (occ) occ[i] = cnt > 0 ? (TO)1 : (TO)0;
        if (occ_keep) occ_keep[i] = kc > 0 ? (TO)1 : (TO)0;
    }
}

static inline int blocks_for(long long n, int cap_mult = 8) {
    long long b = ceil_div64(n, kVoxThreads);
    const long long cap = (long long)kNumSMs * cap_mult;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace sn

extern "C" int sn_vox_minmax(const double* pts, int ld, const int64_t* offsets, int n_clouds, int64_t n_points_total,
                             double* mnmx, void* stream) {
    if (!pts || !mnmx || ld < 3 || n_clouds < 1 || n_clouds > 65535 || n_points_total < 0) return SN_ERR_BAD_ARG;
    if (!offsets && n_clouds != 1) return SN_ERR_BAD_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    sn::minmax_init_kernel<<<sn::ceil_div(n_clouds * 6, 128), 128, 0, s>>>(mnmx, n_clouds);
    SN_LAUNCH_CHECK();
    // offsets live on the device: size the grid for the machine, blocks grid-stride over their cloud
    const int bx = max(1, sn::kNumSMs * 4 / n_clouds);
    sn::minmax_kernel<<<dim3(bx, n_clouds), sn::kVoxThreads, 0, s>>>(pts, ld, (const long long*)offsets, n_points_total, mnmx);
    SN_LAUNCH_CHECK();
    return SN_OK;
}

extern "C" int sn_vox_edges(const double* mnmx, int n_clouds, int nx, int ny, int nz, double* edges, void* stream) {
    if (!mnmx || !edges || n_clouds < 1 || nx < 1 || ny < 1 || nz < 1) return SN_ERR_BAD_ARG;
    sn::edges_kernel<<<n_clouds, 128, 0, (cudaStream_t)stream>>>(mnmx, nx, ny, nz, edges);
    SN_LAUNCH_CHECK();
    return SN_OK;
}

extern "C" int sn_vox_bin(const double* pts, int ld, const double* labels, int label_ld, const int64_t* offsets,
                          int n_clouds, int64_t n_points_total, const double* edges, int nx, int ny, int nz,
                          const double* keep, int n_keep, int32_t* count, int32_t* keep_count, double* max_label,
                          int32_t* lin_out, void* stream) {
    if (!pts || !edges || !count || ld < 3 || n_clouds < 1 || n_clouds > 65535) return SN_ERR_BAD_ARG;
    if (!offsets && n_clouds != 1) return SN_ERR_BAD_ARG;
    if (nx < 1 || ny < 1 || nz < 1 || n_keep < 0 || (n_keep > 0 && !keep) || n_points_total < 0) return SN_ERR_BAD_ARG;
    if ((long long)nx * ny * nz > 0x7fffffffLL) return SN_ERR_UNSUPPORTED;
    if (labels && label_ld < 1) return SN_ERR_BAD_ARG;
    if ((size_t)(nx + ny + nz + 3 + n_keep) * sizeof(double) > 48 * 1024) return SN_ERR_UNSUPPORTED;
    const size_t smem = (size_t)(nx + ny + nz + 3 + n_keep + (n_keep & 1)) * sizeof(double) + 3 * sn::kBinSlots * sizeof(int);
    cudaError_t ea = cudaFuncSetAttribute(sn::bin_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ea == cudaSuccess) ea = cudaFuncSetAttribute(sn::bin_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ea != cudaSuccess) return sn::cuda_rc(ea);
    cudaStream_t s = (cudaStream_t)stream;
    const long long nvox = (long long)n_clouds * nx * ny * nz;
    sn::bin_init_kernel<<<sn::blocks_for(nvox), sn::kVoxThreads, 0, s>>>(count, keep_count, (long long*)max_label, nvox);
    SN_LAUNCH_CHECK();
    if (n_points_total == 0) return SN_OK;
    const long long per_cloud = sn::ceil_div64(n_points_total, n_clouds);
    int bx = (int)sn::ceil_div64(per_cloud, sn::kVoxThreads * sn::kBinUnroll);
    const int cap = max(1, sn::kNumSMs * 4 / n_clouds);  // one wave: 4 CTAs per SM are resident (48 KB table each)
    bx = bx < 1 ? 1 : (bx > cap ? cap : bx);
    // the data path's layout as a compile-time fact (see bin_kernel)
    const bool fast = ld == 4 && ((uintptr_t)pts & 15) == 0 && labels == pts + 3 && label_ld == 4 && keep_count && n_keep >= 1 &&
                      !max_label && !lin_out;
    if (fast)
        sn::bin_kernel<true><<<dim3(bx, n_clouds), sn::kVoxThreads, smem, s>>>(pts, ld, labels, label_ld, (const long long*)offsets, edges,
                                                                    nx, ny, nz, keep, n_keep, count, keep_count,
                                                                    (long long*)max_label, lin_out, n_points_total, nullptr, nullptr);
    else
        sn::bin_kernel<false><<<dim3(bx, n_clouds), sn::kVoxThreads, smem, s>>>(pts, ld, labels, label_ld, (const long long*)offsets, edges,
                                                                    nx, ny, nz, keep, n_keep, count, keep_count,
                                                                    (long long*)max_label, lin_out, n_points_total, nullptr, nullptr);
    SN_LAUNCH_CHECK();
    return SN_OK;
}

extern "C" int sn_vox_voxelize(const double* pts, int ld, const double* labels, int label_ld, const int64_t* offsets,
                               int n_clouds, int64_t n_points_total, int nx, int ny, int nz, const double* keep, int n_keep,
                               double* mnmx, double* edges, int32_t* count, int32_t* keep_count, double* max_label,
                               int32_t* lin_out, void* stream) {
    if (!pts || !mnmx || !edges || !count || ld < 3 || n_clouds < 1 || n_clouds > 65535) return SN_ERR_BAD_ARG;
    if (!offsets && n_clouds != 1) return SN_ERR_BAD_ARG;
    if (nx < 1 || ny < 1 || nz < 1 || n_keep < 0 || (n_keep > 0 && !keep) || n_points_total < 0) return SN_ERR_BAD_ARG;
    if ((long long)nx * ny * nz > 0x7fffffffLL) return SN_ERR_UNSUPPORTED;
    if (labels && label_ld < 1) return SN_ERR_BAD_ARG;
    if ((size_t)(nx + ny + nz + 3 + n_keep) * sizeof(double) > 48 * 1024) return SN_ERR_UNSUPPORTED;
    const size_t smem = (size_t)(nx + ny + nz + 3 + n_keep + (n_keep & 1)) * sizeof(double) + 3 * sn::kBinSlots * sizeof(int);
    cudaError_t ea = cudaFuncSetAttribute(sn::bin_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ea == cudaSuccess) ea = cudaFuncSetAttribute(sn::bin_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ea != cudaSuccess) return sn::cuda_rc(ea);
    cudaStream_t s = (cudaStream_t)stream;
    const long long nvox = (long long)n_clouds * nx * ny * nz;
    // 1. every initialisation in one launch
    sn::vox_init_kernel<<<sn::blocks_for(nvox), sn::kVoxThreads, 0, s>>>(mnmx, n_clouds, count, keep_count, (long long*)max_label, nvox);
    SN_LAUNCH_CHECK();
    // 2. bounding boxes
    const int bxm = max(1, sn::kNumSMs * 4 / n_clouds);
    sn::minmax_kernel<<<dim3(bxm, n_clouds), sn::kVoxThreads, 0, s>>>(pts, ld, (const long long*)offsets, n_points_total, mnmx);
    SN_LAUNCH_CHECK();
    // 3. binning; the edges are derived from the boxes inside the kernel and published by the first CTA of each cloud
    const long long per_cloud = sn::ceil_div64(n_points_total > 0 ? n_points_total : 1, n_clouds);
    int bx = (int)sn::ceil_div64(per_cloud, sn::kVoxThreads * sn::kBinUnroll);
    const int cap = max(1, sn::kNumSMs * 4 / n_clouds);  // one wave: 4 CTAs per SM are resident (48 KB table each)
    bx = bx < 1 ? 1 : (bx > cap ? cap : bx);
    // the data path's layout as a compile-time fact (see bin_kernel)
    const bool fast = ld == 4 && ((uintptr_t)pts & 15) == 0 && labels == pts + 3 && label_ld == 4 && keep_count && n_keep >= 1 &&
                      !max_label && !lin_out;
    if (fast)
        sn::bin_kernel<true><<<dim3(bx, n_clouds), sn::kVoxThreads, smem, s>>>(pts, ld, labels, label_ld, (const long long*)offsets, nullptr,
                                                                    nx, ny, nz, keep, n_keep, count, keep_count,
                                                                    (long long*)max_label, lin_out, n_points_total, mnmx, edges);
    else
        sn::bin_kernel<false><<<dim3(bx, n_clouds), sn::kVoxThreads, smem, s>>>(pts, ld, labels, label_ld, (const long long*)offsets, nullptr,
                                                                    nx, ny, nz, keep, n_keep, count, keep_count,
                                                                    (long long*)max_label, lin_out, n_points_total, mnmx, edges);
    SN_LAUNCH_CHECK();
    return SN_OK;
}

extern "C" int64_t sn_vox_finalize_workspace_bytes(int n_clouds, int ny) { return (int64_t)n_clouds * ny * 2 * 4; }

extern "C" int sn_vox_finalize(const int32_t* count, const int32_t* keep_count, int n_clouds, int nx, int ny, int nz,
                               double* density, double* frac, double* max_label, void* occ, void* occ_keep, int out_dtype,
                               void* ws, void* stream) {
    if (!count || n_clouds < 1 || nx < 1 || ny < 1 || nz < 1) return SN_ERR_BAD_ARG;
    if (out_dtype != SN_F32 && out_dtype != SN_F64) return SN_ERR_BAD_ARG;
    if ((frac || occ_keep) && !keep_count) return SN_ERR_BAD_ARG;
    if (density && !ws) return SN_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    const long long V = (long long)nx * ny * nz, n = V * n_clouds;
    int* cm = reinterpret_cast<int*>(ws);
    if (density) {
        const int ncm = n_clouds * ny * 2;
        sn::colminmax_init_kernel<<<sn::ceil_div(ncm, 128), 128, 0, s>>>(cm, ncm);
        SN_LAUNCH_CHECK();
        const int rows = nz * nx;
        const int rpb = max(8, sn::ceil_div(rows, max(1, sn::kNumSMs * 4 / n_clouds)));
        sn::colminmax_kernel<<<dim3(sn::ceil_div(rows, rpb), n_clouds), sn::kVoxThreads, 0, s>>>(count, rows, ny, rpb, cm);
        SN_LAUNCH_CHECK();
    }
    const int grid = sn::blocks_for(n);
    if (out_dtype == SN_F32)
        sn::finalize_kernel<float><<<grid, sn::kVoxThreads, 0, s>>>(count, keep_count, cm, n, ny, V, density, frac, max_label,
                                                                   (float*)occ, (float*)occ_keep);
    else
        sn::finalize_kernel<double><<<grid, sn::kVoxThreads, 0, s>>>(count, keep_count, cm, n, ny, V, density, frac, max_label,
                                                                    (double*)occ, (double*)occ_keep);
    SN_LAUNCH_CHECK();
    return SN_OK;
}
