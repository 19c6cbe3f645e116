// Observer backward, data part: the tap gradient
//     W[t] = sum_{b,v} G0[b,v] * xpad[b, v + t],   G0 = dpred * (1 - pred^2) * [pred > 0]
// Replaces aten::convolution_backward (weight gradient) + the relu/tanh backward that
// autograd runs for SCENE_Net.py:325-337.  W is ONE T-vector (not G of them): the per-
// operator gradients are lambda_g * W (synth.cu::sn_scenenet_param_grads).
//
// Pass 1 (g0_kernel, HBM-bound elementwise): G0 in float64 from pred/dpred, rounded once to float32.
// Pass 2 (stencil_bwd_kernel, FP32-bound): persistent CTAs walk 8 x TX x TY voxel tiles (static
// round-robin => deterministic).  Per tile two TMA loads (x halo tile, G0 tile) land on one mbarrier.
// Warp w of a CTA owns tap group (dx, z-chunk) for the whole kernel and keeps its C*KY
// accumulators in registers; its lanes sweep the tile's 8x4 micro-tiles.  At the end every
// warp reduces its accumulators across lanes in float64 and writes one partial row; a second
// tiny kernel sums the rows in fixed order (no floating-point atomics anywhere).
#include <stdlib.h>
#include "stencil_common.cuh"

namespace sn {
int stencil_bwd_ky3(const BwdParams&, void*, int64_t, int*, int*, cudaStream_t);
int64_t stencil_bwd_ws_ky3(int, int, int, int, int, int);
int stencil_bwd_ky5(const BwdParams&, void*, int64_t, int*, int*, cudaStream_t);
int64_t stencil_bwd_ws_ky5(int, int, int, int, int, int);
int stencil_bwd_ky6(const BwdParams&, void*, int64_t, int*, int*, cudaStream_t);
int64_t stencil_bwd_ws_ky6(int, int, int, int, int, int);
int stencil_bwd_ky7(const BwdParams&, void*, int64_t, int*, int*, cudaStream_t);
int64_t stencil_bwd_ws_ky7(int, int, int, int, int, int);
int stencil_bwd_ky9(const BwdParams&, void*, int64_t, int*, int*, cudaStream_t);
int64_t stencil_bwd_ws_ky9(int, int, int, int, int, int);
int stencil_bwd_ky11(const BwdParams&, void*, int64_t, int*, int*, cudaStream_t);
int64_t stencil_bwd_ws_ky11(int, int, int, int, int, int);
int stencil_bwd_ky13(const BwdParams&, void*, int64_t, int*, int*, cudaStream_t);
int64_t stencil_bwd_ws_ky13(int, int, int, int, int, int);
int stencil_bwd_ky15(const BwdParams&, void*, int64_t, int*, int*, cudaStream_t);
int64_t stencil_bwd_ws_ky15(int, int, int, int, int, int);
int stencil_tapgrad_generic(const BwdParams& p, int ky, double* W, cudaStream_t stream);  // stencil_generic.cu
// stencil_bwd_sparse.cu
int64_t tapgrad_sparse_ws(int B, int Z, int X, int Y, int kz, int kx, int ky);
int tapgrad_sparse_launch(const float* x, const float* g0, const unsigned long long* nnz, unsigned long long nnz_max,
                          int B, int Z, int X, int Y, int kz, int kx, int ky, void* ws, int64_t ws_bytes, int* rows_out,
                          double* W, unsigned long long* ticket, cudaStream_t stream, const unsigned long long* state);

// G0 = dpred * (1 - pred^2) * [pred > 0]: 4 voxels per thread, all loads issued before use
template <typename TP, typename TD>
__global__ void __launch_bounds__(256) g0_kernel(const TP* __restrict__ pred, const TD* __restrict__ dpred,
                                                 float* __restrict__ g0, long long n) {
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        double pv[4], dv[4];
        if constexpr (sizeof(TP) == 8) {
            const double2 a = reinterpret_cast<const double2*>(pred)[2 * i], c = reinterpret_cast<const double2*>(pred)[2 * i + 1];
            pv[0] = a.x; pv[1] = a.y; pv[2] = c.x; pv[3] = c.y;
        } else {
            const float4 a = reinterpret_cast<const float4*>(pred)[i];
            pv[0] = a.x; pv[1] = a.y; pv[2] = a.z; pv[3] = a.w;
        }
        if constexpr (sizeof(TD) == 8) {
            const double2 a = reinterpret_cast<const double2*>(dpred)[2 * i], c = reinterpret_cast<const double2*>(dpred)[2 * i + 1];
            dv[0] = a.x; dv[1] = a.y; dv[2] = c.x; dv[3] = c.y;
        } else {
            const float4 a = reinterpret_cast<const float4*>(dpred)[i];
            dv[0] = a.x; dv[1] = a.y; dv[2] = a.z; dv[3] = a.w;
        }
        reinterpret_cast<float4*>(g0)[i] = make_float4(g0_of(pv[0], dv[0]), g0_of(pv[1], dv[1]), g0_of(pv[2], dv[2]), g0_of(pv[3], dv[3]));
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const long long i = (n4 << 2) + threadIdx.x;
        g0[i] = g0_of((double)pred[i], (double)dpred[i]);
    }
}

// W[t] = sum over partial rows in a fixed order (deterministic): block = 32 taps x 32 row groups
// rows: the dense kernel's row count; rows_sparse: the occupancy-driven kernel's — the same device-side test as in
// the two kernels tells which of them wrote the rows
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const double* __restrict__ partial, int rows, int rows_sparse,
                                                               const unsigned long long* __restrict__ nnz,
                                                               unsigned long long nnz_max, int TP, int T,
                                                               double* __restrict__ W) {
    __shared__ double red[32][33];
    if (nnz && *nnz <= nnz_max) rows = rows_sparse;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int t = blockIdx.x * 32 + tx;
    double s = 0.0;
    if (t < T) {
        // independent loads first (memory-level parallelism), then a fixed-order sum
        double v[16];
        int r = ty;
        while (r < rows) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = (r + 32 * i < rows) ? partial[(size_t)(r + 32 * i) * TP + t] : 0.0;
#pragma unroll
            for (int i = 0; i < 16; ++i) s += v[i];
            r += 32 * 16;
        }
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && t < T) {
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < 32; ++i) a += red[i][tx];
        W[t] = a;
    }
}

static inline int64_t g0_bytes_(int B, int Z, int X, int Y) { return (((int64_t)B * Z * X * Y * 4) + 255) & ~(int64_t)255; }

static int launch_g0(const BwdParams& p, long long n, float* g0, cudaStream_t stream) {
    long long blocks = ceil_div64(n / 4 + 1, 256);
    const int grid = (int)(blocks > (long long)kNumSMs * 16 ? (long long)kNumSMs * 16 : blocks);
    if (p.pred_f64 && p.dpred_f64)
        g0_kernel<double, double><<<grid, 256, 0, stream>>>((const double*)p.pred, (const double*)p.dpred, g0, n);
    else if (p.pred_f64)
        g0_kernel<double, float><<<grid, 256, 0, stream>>>((const double*)p.pred, (const float*)p.dpred, g0, n);
    else if (p.dpred_f64)
        g0_kernel<float, double><<<grid, 256, 0, stream>>>((const float*)p.pred, (const double*)p.dpred, g0, n);
    else
        g0_kernel<float, float><<<grid, 256, 0, stream>>>((const float*)p.pred, (const float*)p.dpred, g0, n);
    SN_LAUNCH_CHECK();
    return SN_OK;
}

}  // namespace sn

extern "C" int sn_select_fwd_path(int64_t nnz, int B, int Z, int X, int Y, int kz, int kx, int ky);  // stencil_fwd.cu

static int64_t tapgrad_ws_dense(int B, int Z, int X, int Y, int kz, int kx, int ky) {
    switch (ky) {
        case 3: return sn::stencil_bwd_ws_ky3(B, Z, X, Y, kz, kx);
        case 5: return sn::stencil_bwd_ws_ky5(B, Z, X, Y, kz, kx);
        case 6: return sn::stencil_bwd_ws_ky6(B, Z, X, Y, kz, kx);
        case 7: return sn::stencil_bwd_ws_ky7(B, Z, X, Y, kz, kx);
        case 9: return sn::stencil_bwd_ws_ky9(B, Z, X, Y, kz, kx);
        case 11: return sn::stencil_bwd_ws_ky11(B, Z, X, Y, kz, kx);
        case 13: return sn::stencil_bwd_ws_ky13(B, Z, X, Y, kz, kx);
        case 15: return sn::stencil_bwd_ws_ky15(B, Z, X, Y, kz, kx);
        default: return 256;  // generic path needs no workspace
    }
}
static int64_t tapgrad_ws(int B, int Z, int X, int Y, int kz, int kx, int ky) {
    const int64_t a = tapgrad_ws_dense(B, Z, X, Y, kz, kx, ky), b = sn::tapgrad_sparse_ws(B, Z, X, Y, kz, kx, ky);
    return a > b ? a : b;
}
static bool fast_ky(int ky) { return ky == 3 || ky == 5 || ky == 6 || ky == 7 || ky == 9 || ky == 11 || ky == 13 || ky == 15; }

extern "C" int64_t sn_scenenet_tapgrad_workspace_bytes(int B, int Z, int X, int Y, int kz, int kx, int ky) {
    if (B < 1 || Z < 1 || X < 1 || Y < 1 || kz < 1 || kx < 1 || ky < 1) return SN_ERR_BAD_ARG;
    return tapgrad_ws(B, Z, X, Y, kz, kx, ky);
}

extern "C" int64_t sn_scenenet_bwd_workspace_bytes(int B, int Z, int X, int Y, int kz, int kx, int ky) {
    if (B < 1 || Z < 1 || X < 1 || Y < 1 || kz < 1 || kx < 1 || ky < 1) return SN_ERR_BAD_ARG;
    return sn::g0_bytes_(B, Z, X, Y) + tapgrad_ws(B, Z, X, Y, kz, kx, ky);
}

extern "C" int sn_scenenet_g0(const void* pred, int pred_dtype, const void* dpred, int dpred_dtype, int64_t n, float* g0,
                              void* stream) {
    if (!pred || !dpred || !g0 || n < 1) return SN_ERR_BAD_ARG;
    if ((pred_dtype != SN_F32 && pred_dtype != SN_F64) || (dpred_dtype != SN_F32 && dpred_dtype != SN_F64))
        return SN_ERR_BAD_ARG;
    if ((((uintptr_t)pred) | ((uintptr_t)dpred) | ((uintptr_t)g0)) & 15) return SN_ERR_ALIGN;
    sn::BwdParams p{};
    p.pred = pred; p.dpred = dpred;
    p.B = 1; p.Z = 1; p.X = 1; p.Y = 1;
    p.pred_f64 = pred_dtype == SN_F64; p.dpred_f64 = dpred_dtype == SN_F64;
    return sn::launch_g0(p, n, g0, (cudaStream_t)stream);
}

// occupancy (in percent of the voxels) up to which the occupancy-driven kernel is selected: measured break-even
// at config 2 is ~11 % (dense 101 us flat; occupancy-driven 34 us + 5.5 us per percent, scratch/time_sparse.py)
static unsigned long long sparse_nnz_max(long long nvox) {
    static const int pct = SN_ENV("SN_SPARSE_PCT") ? atoi(SN_ENV("SN_SPARSE_PCT")) : 10;
    return (unsigned long long)(nvox / 100 * pct);
}

extern "C" int sn_scenenet_tapgrad(const float* x, const float* g0, const unsigned long long* nnz, int mode,
                                   int B, int Z, int X, int Y, int kz, int kx, int ky,
                                   double* W, void* ws, int64_t ws_bytes, void* stream) {
    if (!x || !g0 || !W) return SN_ERR_BAD_ARG;
    if (B < 1 || Z < 1 || X < 1 || Y < 1 || kz < 1 || kx < 1 || ky < 1) return SN_ERR_BAD_ARG;
    if (mode != SN_TAPGRAD_AUTO && mode != SN_TAPGRAD_DENSE && mode != SN_TAPGRAD_SPARSE) return SN_ERR_BAD_ARG;
    if ((long long)kz * kx * ky > SN_MAX_TAPS) return SN_ERR_UNSUPPORTED;
    if (ws && ((uintptr_t)ws & 15)) return SN_ERR_ALIGN;
    if (nnz && ((uintptr_t)nnz & 7)) return SN_ERR_ALIGN;
    sn::BwdParams p{};
    p.x = x; p.g0 = g0;
    p.B = B; p.Z = Z; p.X = X; p.Y = Y; p.kz = kz; p.kx = kx;
    cudaStream_t s = (cudaStream_t)stream;
    const int T = kz * kx * ky;
    const long long nvox = (long long)B * Z * X * Y;
    const bool sparse_ok = ws && sn::tapgrad_sparse_ws(B, Z, X, Y, kz, kx, ky) > 0;
    // which kernels are enqueued: forced modes run one of them unconditionally; AUTO with a count enqueues both and
    // the device decides; widths the dense stencil is not instantiated for always take the occupancy-driven kernel
    bool run_sparse, run_dense;
    const unsigned long long* gate = nullptr;
    if (mode == SN_TAPGRAD_SPARSE) {
        if (!sparse_ok) return ws ? SN_ERR_UNSUPPORTED : SN_ERR_WORKSPACE;
        run_sparse = true; run_dense = false;
    } else if (mode == SN_TAPGRAD_DENSE || !sparse_ok) {
        run_sparse = false; run_dense = true;
    } else if (!fast_ky(ky)) {
        run_sparse = true; run_dense = false;
    } else if (nnz) {
        run_sparse = run_dense = true; gate = nnz;
    } else {
        run_sparse = false; run_dense = true;
    }
    const unsigned long long nnz_max = sparse_nnz_max(nvox);
    if (run_dense && !fast_ky(ky)) return sn::stencil_tapgrad_generic(p, ky, W, s);
    if (!ws) return SN_ERR_WORKSPACE;
    int rows = 0, rows_sparse = 0, TP = (T + 31) & ~31, rc = SN_OK;
    // The partial rows are summed by a separate kernel.  (Round 1 let the CTA that finishes last sum them — a ticket counter
    // in the state buffer's second word, last_cta_row_sum — to save a launch: one CTA reading 148 x T doubles is slower than
    // T / 32 CTAs after a 1 us kernel boundary: config 2 step 0.1389 -> 0.1359 ms without it, 9^3 taps -10 us, 15^3 -57 us;
    // profiles/r2_notes.md.  The kernels keep the ticket parameter; the word stays reserved.)
    unsigned long long* ticket = nullptr;
    if (run_sparse) {
        rc = sn::tapgrad_sparse_launch(x, g0, gate, nnz_max, B, Z, X, Y, kz, kx, ky, ws, ws_bytes, &rows_sparse, W, ticket, s, nnz);
        if (rc) return rc;
    }
    if (run_dense) {
        p.nnz = gate; p.nnz_max = nnz_max; p.W = W; p.ticket = ticket;
        switch (ky) {
            case 3: rc = sn::stencil_bwd_ky3(p, ws, ws_bytes, &rows, &TP, s); break;
            case 5: rc = sn::stencil_bwd_ky5(p, ws, ws_bytes, &rows, &TP, s); break;
            case 6: rc = sn::stencil_bwd_ky6(p, ws, ws_bytes, &rows, &TP, s); break;
            case 7: rc = sn::stencil_bwd_ky7(p, ws, ws_bytes, &rows, &TP, s); break;
            case 9: rc = sn::stencil_bwd_ky9(p, ws, ws_bytes, &rows, &TP, s); break;
            case 11: rc = sn::stencil_bwd_ky11(p, ws, ws_bytes, &rows, &TP, s); break;
            case 13: rc = sn::stencil_bwd_ky13(p, ws, ws_bytes, &rows, &TP, s); break;
            default: rc = sn::stencil_bwd_ky15(p, ws, ws_bytes, &rows, &TP, s); break;
        }
        if (rc) return rc;
    } else {
        rows = rows_sparse;  // ungated sparse run: the reduction reads its rows
    }
    if (ticket) return SN_OK;  // W was written by the last CTA of whichever kernel ran
    // fixed-order float64 reduction of the partial rows
    sn::reduce_partials_kernel<<<sn::ceil_div(T, 32), dim3(32, 32), 0, s>>>(reinterpret_cast<const double*>(ws), rows, rows_sparse,
                                                                              gate, nnz_max, TP, T, W);
    SN_LAUNCH_CHECK();
    return SN_OK;
}

extern "C" int sn_scenenet_bwd(const float* x, const unsigned long long* nnz, int mode, const void* pred, int pred_dtype,
                               const void* dpred, int dpred_dtype, int B, int Z, int X, int Y, int kz, int kx, int ky,
                               double* W, void* ws, int64_t ws_bytes, void* stream) {
    if (!x || !pred || !dpred || !W || !ws) return SN_ERR_BAD_ARG;
    if (B < 1 || Z < 1 || X < 1 || Y < 1 || kz < 1 || kx < 1 || ky < 1) return SN_ERR_BAD_ARG;
    if ((uintptr_t)ws & 15) return SN_ERR_ALIGN;
    const int64_t gb = sn::g0_bytes_(B, Z, X, Y);
    if (gb > ws_bytes) return SN_ERR_WORKSPACE;
    float* g0 = reinterpret_cast<float*>(ws);
    int rc = sn_scenenet_g0(pred, pred_dtype, dpred, dpred_dtype, (int64_t)B * Z * X * Y, g0, stream);
    if (rc) return rc;
    return sn_scenenet_tapgrad(x, g0, nnz, mode, B, Z, X, Y, kz, kx, ky, W, reinterpret_cast<char*>(ws) + gb,
                               ws_bytes - gb, stream);
}

// the selection rule of the AUTO modes, for hosts that know the occupancy (e.g. when a step is captured for replay on
// grids of one kind): which = 0 forward, 1 tap gradient.  Returns SN_PATH_DENSE or SN_PATH_SPARSE.
extern "C" int sn_select_path(int which, int64_t nnz, int B, int Z, int X, int Y, int kz, int kx, int ky) {
    if (B < 1 || Z < 1 || X < 1 || Y < 1 || kz < 1 || kx < 1 || ky < 1 || nnz < 0) return SN_ERR_BAD_ARG;
    const long long nvox = (long long)B * Z * X * Y;
    if (which == 1) {
        const bool ok = sn::tapgrad_sparse_ws(B, Z, X, Y, kz, kx, ky) > 0;
        if (ok && !fast_ky(ky)) return SN_PATH_SPARSE;
        return (ok && (unsigned long long)nnz <= sparse_nnz_max(nvox)) ? SN_PATH_SPARSE : SN_PATH_DENSE;
    }
    if (which == 0) return sn_select_fwd_path(nnz, B, Z, X, Y, kz, kx, ky);
    return SN_ERR_BAD_ARG;
}
