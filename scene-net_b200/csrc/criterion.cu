// Fused GENEO_Tversky_Loss — SURVEY §8(f) rank 1.
//
// Reference arithmetic being replaced (paths relative to the reference root):
//   core/criterions/w_mse.py:114-151        WeightedMSE.get_dens_target / get_weight_target / forward
//   core/criterions/tversky_loss.py:81-95   FocalTverskyLoss.forward
//   core/criterions/geneo_loss.py:36-71     GENEO_Loss.cvx_loss / positive_regularizer
//   core/criterions/geneo_loss.py:145-161   GENEO_Tversky_Loss.forward
// The reference evaluates ~40 full-tensor float64 passes (argmin over 10 bins, 10 masked index_puts, weights,
// mean, three products and sums) plus ~40 scalar ops for the penalties, and autograd replays them backwards.
// Here: ONE pass over (pred, y) that accumulates per histogram bin the count and sum of (y - p)^2 together with
// TP / FP / FN, a one-thread finalisation that reproduces the reference's float32 weight arithmetic on the 10-bin
// table and emits the loss plus the coefficients of the closed-form dL/dpred, and ONE elementwise pass for the
// backward.  Deterministic: per-thread accumulators, fixed-order tree, per-CTA partial rows summed in order.
#include <math.h>
#include "common.cuh"

namespace sn {

constexpr int kCritThreads = 256;
constexpr int kCritMaxBins = SN_CRIT_MAX_BINS;
constexpr int kCritRow = 2 * kCritMaxBins + 3;  // per-CTA partial row: cnt[nb], S[nb], TP, FP, FN

struct CritTable {
    double ranges[kCritMaxBins];  // bin positions (float32 values, widened)
    float w_raw[kCritMaxBins];    // max(1 - alpha * dens_k, eps), float32 like the reference
    int nbins;
    int bin_of_zero;              // argmin_k |0 - ranges[k]| (first minimum): the bin of an empty target voxel
};

// hist bin of a target value: first index of the minimum of |y - ranges[k]| (w_mse.py:120, torch.argmin)
__device__ __forceinline__ int bin_of(double y, const CritTable& t) {
    int best = 0;
    double bd = fabs(y - t.ranges[0]);
#pragma unroll 1
    for (int k = 1; k < t.nbins; ++k) {
        const double d = fabs(y - t.ranges[k]);
        if (d < bd) { bd = d; best = k; }
    }
    return best;
}

template <typename T>
__global__ void __launch_bounds__(kCritThreads) crit_reduce_kernel(const T* __restrict__ pred, const T* __restrict__ y, long long n,
                                                                   const __grid_constant__ CritTable tab,
                                                                   double* __restrict__ partial) {
    // per-thread accumulators of the rare bins live in shared memory ([k][tid]: a thread only ever touches its own
    // column); the bin of y == 0 (almost every voxel) and TP / FP / FN stay in registers
    __shared__ double s_S[kCritMaxBins][kCritThreads];
    __shared__ unsigned s_cnt[kCritMaxBins][kCritThreads];
    __shared__ double s_red[kCritThreads / 32][kCritRow];
    const int tid = threadIdx.x, nb = tab.nbins, k0 = tab.bin_of_zero;
    for (int k = 0; k < nb; ++k) {
        s_S[k][tid] = 0.0;
        s_cnt[k][tid] = 0u;
    }
    double S0 = 0.0, tp = 0.0, fp = 0.0, fn = 0.0;
    unsigned c0 = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    constexpr int V = 16 / sizeof(T);  // elements per 16-byte load
    const long long nv = n / V;
    for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < nv; i += stride) {
        alignas(16) T pv[V];
        alignas(16) T yv[V];
        *reinterpret_cast<uint4*>(pv) = reinterpret_cast<const uint4*>(pred)[i];
        *reinterpret_cast<uint4*>(yv) = reinterpret_cast<const uint4*>(y)[i];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const double p = (double)pv[j], t = (double)yv[j];
            const double d = t - p;
            if (t == 0.0) {
                S0 += d * d;
                ++c0;
                fp += p;  // (1 - y) * p
            } else {
                const int k = bin_of(t, tab);
                s_S[k][tid] += d * d;
                s_cnt[k][tid] += 1u;
                tp += p * t;
                fp += (1.0 - t) * p;
                fn += t * (1.0 - p);
            }
        }
    }
    if (blockIdx.x == 0 && tid < (int)(n - nv * V)) {  // tail
        const long long i = nv * V + tid;
        const double p = (double)pred[i], t = (double)y[i];
        const double d = t - p;
        const int k = bin_of(t, tab);
        s_S[k][tid] += d * d;
        s_cnt[k][tid] += 1u;
        tp += p * t;
        fp += (1.0 - t) * p;
        fn += t * (1.0 - p);
    }
    s_S[k0][tid] += S0;
    s_cnt[k0][tid] += c0;
    // fixed-order reduction: lanes -> warp leader -> shared -> thread 0 -> partial row of this CTA
    const int warp = tid >> 5, lane = tid & 31;
    for (int k = 0; k < nb; ++k) {
        const double a = warp_sum((double)s_cnt[k][tid]);  // counts < 2^53: exact in double
        const double b = warp_sum(s_S[k][tid]);
        if (lane == 0) {
            s_red[warp][k] = a;
            s_red[warp][kCritMaxBins + k] = b;
        }
    }
    tp = warp_sum(tp); fp = warp_sum(fp); fn = warp_sum(fn);
    if (lane == 0) {
        s_red[warp][2 * kCritMaxBins] = tp;
        s_red[warp][2 * kCritMaxBins + 1] = fp;
        s_red[warp][2 * kCritMaxBins + 2] = fn;
    }
    __syncthreads();
    if (tid < kCritRow) {
        double a = 0.0;
#pragma unroll
        for (int w = 0; w < kCritThreads / 32; ++w) a += s_red[w][tid];
        partial[(size_t)blockIdx.x * kCritRow + tid] = a;
    }
}

// One CTA: sums the partial rows in a fixed order, then thread 0 does the scalar arithmetic.
// coef: [0, nb) a_k = d(dense)/dp = a_k * (p - y) for a voxel of bin k; [MAX] tvY; [MAX+1] tvN:
//       d(focal tversky)/dp = tvY * y + tvN * (1 - y);  [MAX+2] dense term, [MAX+3] focal tversky term (diagnostics)
__global__ void __launch_bounds__(1024) crit_finalize_kernel(const double* __restrict__ partial, int rows, long long n,
                                                             const __grid_constant__ CritTable tab, float mse_weight,
                                                             double tv_alpha, double tv_beta, double gamma, double smooth,
                                                             int terms, double* __restrict__ loss, double* __restrict__ coef) {
    // 32 row groups x 32 columns: every thread sums its rows in order, then one thread per column sums the 32 group
    // sums in order (a single warp walking all rows took 28 us: one dependent load chain per column)
    __shared__ double s_grp[32][kCritRow + 1];
    __shared__ double s_tot[kCritRow];
    const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
    for (int cc = c; cc < kCritRow; cc += 32) {
        double a = 0.0;
        // 8 independent loads in flight, then the adds in row order (one dependent load per add made this loop the kernel:
        // 19 L2 round trips per thread)
        for (int r0 = g; r0 < rows; r0 += 32 * 8) {
            double v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (r0 + 32 * i < rows) ? __ldcg(partial + (size_t)(r0 + 32 * i) * kCritRow + cc) : 0.0;
#pragma unroll
            for (int i = 0; i < 8; ++i) a += v[i];
        }
        s_grp[g][cc] = a;
    }
    __syncthreads();
    if (threadIdx.x < kCritRow) {
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < 32; ++i) a += s_grp[i][threadIdx.x];
        s_tot[threadIdx.x] = a;
    }
    __syncthreads();
    const int lane = threadIdx.x;
    if (lane != 0) return;
    const int nb = tab.nbins;
    // weights / mean(weights): float32 tensors in the reference (w_mse.py:141-145)
    double wsum = 0.0;
    for (int k = 0; k < nb; ++k) wsum += s_tot[k] * (double)tab.w_raw[k];
    const float mean_w = (float)(wsum / (double)n);
    double dense = 0.0;
    for (int k = 0; k < nb; ++k) {
        const float wk = tab.w_raw[k] / mean_w;
        const double ck = (double)(mse_weight * wk);  // self.mse_weight * weights: float32 (w_mse.py:151)
        dense += ck * s_tot[kCritMaxBins + k];
        coef[k] = (terms & 1) ? 2.0 * ck / (double)n : 0.0;
    }
    dense /= (double)n;
    const double TP = s_tot[2 * kCritMaxBins], FP = s_tot[2 * kCritMaxBins + 1], FN = s_tot[2 * kCritMaxBins + 2];
    const double N = TP + smooth, D = TP + tv_alpha * FP + tv_beta * FN + smooth;
    const double tv = N / D;
    const double ft = pow(1.0 - tv, gamma);
    const double dft = gamma == 1.0 ? -1.0 : -gamma * pow(1.0 - tv, gamma - 1.0);
    coef[kCritMaxBins] = (terms & 2) ? dft * (D - N * (1.0 - tv_beta)) / (D * D) : 0.0;
    coef[kCritMaxBins + 1] = (terms & 2) ? dft * (-N * tv_alpha) / (D * D) : 0.0;
    coef[kCritMaxBins + 2] = dense;
    coef[kCritMaxBins + 3] = ft;
    loss[0] = ((terms & 1) ? dense : 0.0) + ((terms & 2) ? ft : 0.0);
}

// dpred = grad_out * (a_bin(y) (p - y) + tvY y + tvN (1 - y)); G0 variant: * (1 - p^2) [p > 0], float32
template <typename T, bool G0>
__global__ void __launch_bounds__(256) crit_bwd_kernel(const T* __restrict__ pred, const T* __restrict__ y, long long n,
                                                       const __grid_constant__ CritTable tab, const double* __restrict__ coef,
                                                       const T* __restrict__ grad_out, void* __restrict__ out) {
    __shared__ double s_a[kCritMaxBins + 2];
    if (threadIdx.x < kCritMaxBins + 2) s_a[threadIdx.x] = coef[threadIdx.x] * (grad_out ? (double)grad_out[0] : 1.0);
    __syncthreads();
    const double a0 = s_a[tab.bin_of_zero], tvY = s_a[kCritMaxBins], tvN = s_a[kCritMaxBins + 1];
    const long long stride = (long long)gridDim.x * blockDim.x;
    constexpr int V = 16 / sizeof(T);
    const long long nv = n / V;
    auto one = [&](double p, double t) -> double {
        double g;
        if (t == 0.0)
            g = a0 * p + tvN;
        else
            g = s_a[bin_of(t, tab)] * (p - t) + tvY * t + tvN * (1.0 - t);
        if (G0) g = p > 0.0 ? g * (1.0 - p * p) : 0.0;
        return g;
    };
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        alignas(16) T pv[V];
        alignas(16) T yv[V];
        *reinterpret_cast<uint4*>(pv) = reinterpret_cast<const uint4*>(pred)[i];
        *reinterpret_cast<uint4*>(yv) = reinterpret_cast<const uint4*>(y)[i];
        if (G0) {
            float o[V];
#pragma unroll
            for (int j = 0; j < V; ++j) o[j] = (float)one((double)pv[j], (double)yv[j]);
            float* dst = reinterpret_cast<float*>(out) + i * V;
            if (V == 4)
                *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[V > 2 ? 2 : 0], o[V > 3 ? 3 : 0]);
            else
                *reinterpret_cast<float2*>(dst) = make_float2(o[0], o[1]);
        } else {
            alignas(16) T o[V];
#pragma unroll
            for (int j = 0; j < V; ++j) o[j] = (T)one((double)pv[j], (double)yv[j]);
            reinterpret_cast<uint4*>(out)[i] = *reinterpret_cast<uint4*>(o);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - nv * V)) {
        const long long i = nv * V + threadIdx.x;
        const double g = one((double)pred[i], (double)y[i]);
        if (G0)
            reinterpret_cast<float*>(out)[i] = (float)g;
        else
            reinterpret_cast<T*>(out)[i] = (T)g;
    }
}

// Penalties on the live parameters (geneo_loss.py:36-71), float32, summed left to right like python's sum():
//   role 0: relu(-v)  (positive_regularizer term / a non-last convex coefficient)
//   role 1: the last convex coefficient: contributes relu(-(1 - sum(all lambdas) + itself))
//   role 2: like 0 but it is a convex coefficient (takes part in the sum of role 1)
// out[0] = weight * cvx_loss, out[1] = weight * positive_regularizer, out[2 + i] = d(out[0] + out[1]) / d param_i.
struct PenaltyArgs {
    const float* p[SN_MAX_PARAM_PTRS];
    signed char role[SN_MAX_PARAM_PTRS];
    int n;
};
// loss_accum (nullable): the criterion's float64 loss scalar; both penalties are added to it here, in the reference's
// order ((loss + cvx) + pos, float32 terms widened), so the host adds nothing afterwards
__global__ void __launch_bounds__(128) penalty_kernel(const __grid_constant__ PenaltyArgs a, float weight, float* __restrict__ out,
                                                      double* __restrict__ loss_accum) {
    __shared__ float s_v[SN_MAX_PARAM_PTRS];
    if (threadIdx.x < a.n) s_v[threadIdx.x] = *a.p[threadIdx.x];  // all parameter loads in flight at once
    __syncthreads();
    if (threadIdx.x != 0) return;
    float pos = 0.f, cvx = 0.f, lsum = 0.f, last = 0.f;
    bool has_last = false;
    for (int i = 0; i < a.n; ++i) {
        const float v = s_v[i];
        const int r = a.role[i];
        float g = 0.f;
        if (r == 0) {
            pos = pos + fmaxf(-v, 0.f);
            g = v < 0.f ? -weight : 0.f;
        } else {
            lsum = lsum + v;
            if (r == 2) {
                cvx = cvx + fmaxf(-v, 0.f);
                g = v < 0.f ? -weight : 0.f;
            } else {
                last = v;
                has_last = true;
            }
        }
        out[2 + i] = g;
    }
    if (has_last) {
        const float e = -((1.f - lsum) + last);  // -(1 - sum(values) + last)
        cvx = cvx + fmaxf(e, 0.f);
        if (e > 0.f)
            for (int i = 0; i < a.n; ++i)
                if (a.role[i] == 2) out[2 + i] += weight;  // d e / d lambda_g = +1 for the free coefficients
    }
    out[0] = weight * cvx;
    out[1] = weight * pos;
    if (loss_accum) *loss_accum = (*loss_accum + (double)(weight * cvx)) + (double)(weight * pos);
}

static int crit_grid(long long n) {
    long long b = ceil_div64(n, (long long)kCritThreads * 8);
    const long long cap = (long long)kNumSMs * 4;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

static int fill_table(CritTable& t, const float* ranges_host, const float* w_raw_host, int nbins) {
    if (!ranges_host || !w_raw_host || nbins < 1 || nbins > kCritMaxBins) return SN_ERR_BAD_ARG;
    t.nbins = nbins;
    int best = 0;
    for (int k = 0; k < kCritMaxBins; ++k) {
        t.ranges[k] = k < nbins ? (double)ranges_host[k] : 0.0;
        t.w_raw[k] = k < nbins ? w_raw_host[k] : 0.f;
    }
    for (int k = 1; k < nbins; ++k)
        if (fabs(t.ranges[k]) < fabs(t.ranges[best])) best = k;
    t.bin_of_zero = best;
    return SN_OK;
}

}  // namespace sn

extern "C" int64_t sn_criterion_workspace_bytes(int64_t n) {
    if (n < 1) return SN_ERR_BAD_ARG;
    return (int64_t)sn::crit_grid(n) * sn::kCritRow * 8;
}

extern "C" int sn_criterion_fwd(const void* pred, const void* y, int dtype, int64_t n, const float* ranges_host,
                                const float* w_raw_host, int nbins, float mse_weight, double tversky_alpha,
                                double tversky_beta, double focal_gamma, double tversky_smooth, int terms, double* loss,
                                double* coef, void* ws, int64_t ws_bytes, void* stream) {
    if (!pred || !y || !loss || !coef || !ws || n < 1 || terms < 1 || terms > 3) return SN_ERR_BAD_ARG;
    if (dtype != SN_F32 && dtype != SN_F64) return SN_ERR_BAD_ARG;
    if ((((uintptr_t)pred) | ((uintptr_t)y)) & 15 || ((uintptr_t)ws & 7)) return SN_ERR_ALIGN;
    sn::CritTable t;
    int rc = sn::fill_table(t, ranges_host, w_raw_host, nbins);
    if (rc) return rc;
    const int grid = sn::crit_grid(n);
    if ((int64_t)grid * sn::kCritRow * 8 > ws_bytes) return SN_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    double* partial = reinterpret_cast<double*>(ws);
    if (dtype == SN_F64)
        sn::crit_reduce_kernel<double><<<grid, sn::kCritThreads, 0, s>>>((const double*)pred, (const double*)y, n, t, partial);
    else
        sn::crit_reduce_kernel<float><<<grid, sn::kCritThreads, 0, s>>>((const float*)pred, (const float*)y, n, t, partial);
    SN_LAUNCH_CHECK();
    sn::crit_finalize_kernel<<<1, 1024, 0, s>>>(partial, grid, n, t, mse_weight, tversky_alpha, tversky_beta, focal_gamma,
                                              tversky_smooth, terms, loss, coef);
    SN_LAUNCH_CHECK();
    return SN_OK;
}

extern "C" int sn_criterion_bwd(const void* pred, const void* y, int dtype, int64_t n, const float* ranges_host,
                                const float* w_raw_host, int nbins, const double* coef, const void* grad_out,
                                void* out, int out_g0, void* stream) {
    if (!pred || !y || !coef || !out || n < 1) return SN_ERR_BAD_ARG;
    if (dtype != SN_F32 && dtype != SN_F64) return SN_ERR_BAD_ARG;
    if ((((uintptr_t)pred) | ((uintptr_t)y) | ((uintptr_t)out)) & 15) return SN_ERR_ALIGN;
    sn::CritTable t;
    int rc = sn::fill_table(t, ranges_host, w_raw_host, nbins);
    if (rc) return rc;
    long long b = sn::ceil_div64(n, 256 * 4);
    const int grid = (int)(b > (long long)sn::kNumSMs * 16 ? (long long)sn::kNumSMs * 16 : b);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == SN_F64) {
        if (out_g0)
            sn::crit_bwd_kernel<double, true><<<grid, 256, 0, s>>>((const double*)pred, (const double*)y, n, t, coef, (const double*)grad_out, out);
        else
            sn::crit_bwd_kernel<double, false><<<grid, 256, 0, s>>>((const double*)pred, (const double*)y, n, t, coef, (const double*)grad_out, out);
    } else {
        if (out_g0)
            sn::crit_bwd_kernel<float, true><<<grid, 256, 0, s>>>((const float*)pred, (const float*)y, n, t, coef, (const float*)grad_out, out);
        else
            sn::crit_bwd_kernel<float, false><<<grid, 256, 0, s>>>((const float*)pred, (const float*)y, n, t, coef, (const float*)grad_out, out);
    }
    SN_LAUNCH_CHECK();
    return SN_OK;
}

extern "C" int sn_param_penalty(const float* const* param_ptrs_host, const int32_t* role_host, int n, float weight,
                                float* out, double* loss_accum, void* stream) {
    if (!param_ptrs_host || !role_host || !out || n < 0 || n > SN_MAX_PARAM_PTRS) return SN_ERR_BAD_ARG;
    sn::PenaltyArgs a;
    a.n = n;
    int n_last = 0;
    for (int i = 0; i < SN_MAX_PARAM_PTRS; ++i) {
        a.p[i] = i < n ? param_ptrs_host[i] : nullptr;
        a.role[i] = i < n ? (signed char)role_host[i] : 0;
        if (i < n) {
            if (!param_ptrs_host[i] || role_host[i] < 0 || role_host[i] > 2) return SN_ERR_BAD_ARG;
            n_last += role_host[i] == 1;
        }
    }
    if (n_last > 1) return SN_ERR_BAD_ARG;
    if ((uintptr_t)loss_accum & 7) return SN_ERR_ALIGN;
    sn::penalty_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(a, weight, out, loss_accum);
    SN_LAUNCH_CHECK();
    return SN_OK;
}
