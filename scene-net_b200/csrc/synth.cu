// GENEO kernel synthesis (forward) and its Jacobian^T (backward) — one small launch each.
//
// Reference arithmetic being replaced (paths relative to the reference root):
//   cylinder.py:72-103 / 146-176      cylinder_kernel / cylinderv2
//   arrow.py:170-205 / 208-252        cone_kernel / arrow
//   neg_sphere.py:129-158 / 160-199   neg_sphere_kernel / negSpherev2
//   SCENE_Net.py:103-106, 324-337     GENEO_Layer.compute_kernel, convex combination
//
// Forward follows the reference's float32 op sequence (this file is compiled with
// -fmad=false so no contraction changes a rounding); transcendental functions are evaluated
// in double and rounded once.  Backward evaluates the closed-form Jacobian in double.
// Latency-bound by construction (<= 16 operators x <= 4096 taps): one CTA, no roofline.
#include <math.h>
#include "common.cuh"
#include "peer_exchange.cuh"

namespace sn {

constexpr int kSynthThreads = 256;
constexpr float kEpsV2 = 1e-8f;
constexpr float kPiF = 3.14159265358979323846f;

struct SynthArgs {
    sn_model_desc d;
    const float* p[SN_MAX_PARAM_PTRS];
};

struct OpParams {
    float radius, sigma, apex, cone_inc, cone_radius, neg_factor;
    int hc, ch;
};

__device__ __forceinline__ bool is_plane_kind(int kind) { return kind <= SN_KIND_ARROW_V2; }
__device__ __forceinline__ bool is_v2(int kind) {
    return kind == SN_KIND_CYLINDER_V2 || kind == SN_KIND_ARROW_V2 || kind == SN_KIND_NEGSPHERE_V2;
}
__device__ __forceinline__ int n_params_of(int kind) {
    return kind <= SN_KIND_CYLINDER_V2 ? 2 : (kind <= SN_KIND_ARROW_V2 ? 5 : 3);
}

__device__ OpParams load_op(const SynthArgs& a, int g) {
    OpParams o;
    o.radius = o.sigma = o.apex = o.cone_inc = o.cone_radius = o.neg_factor = 0.f;
    const int kind = a.d.kind[g];
    const float* const* p = a.p + a.d.param_index[g];
    if (kind <= SN_KIND_CYLINDER_V2) {  // radius, sigma
        o.radius = *p[0];
        o.sigma = *p[1];
    } else if (kind <= SN_KIND_ARROW_V2) {  // apex, cone_inc, cone_radius, radius, sigma
        o.apex = *p[0];
        o.cone_inc = *p[1];
        o.cone_radius = *p[2];
        o.radius = *p[3];
        o.sigma = *p[4];
    } else {  // neg_factor, radius, sigma
        o.neg_factor = *p[0];
        o.radius = *p[1];
        o.sigma = *p[2];
    }
    int hc = (int)o.apex;  // apex.to(torch.int): truncation (arrow.py:235)
    hc = hc < 0 ? 0 : (hc > a.d.kz ? a.d.kz : hc);
    o.hc = hc;
    o.ch = a.d.kz - hc;
    return o;
}

// squared distance of tap t to the kernel centre, reproducing the reference's index dance
// (SURVEY §8 a-7..a-9): planes are evaluated at (i,j) = (n % kx, n / kx) with n = p*ky+q;
// volumes at (iz,ix,iy) = (t % kz, (t/kz) % kx, t/(kz*kx)).  float32: (sqrt(sum sq))^2.
__device__ __forceinline__ float tap_d2(bool plane, int t, int kz, int kx, int ky) {
    if (plane) {
        const int n = t % (kx * ky);
        const float di = (float)(n % kx) - (float)(kx - 1) / 2.f;
        const float dj = (float)(n / kx) - (float)(ky - 1) / 2.f;
        const float nrm = sqrtf(di * di + dj * dj);
        return nrm * nrm;
    }
    const float dz = (float)(t % kz) - (float)(kz - 1) / 2.f;
    const float dx = (float)((t / kz) % kx) - (float)(kx - 1) / 2.f;
    const float dy = (float)(t / (kz * kx)) - (float)(ky - 1) / 2.f;
    const float nrm = sqrtf((dz * dz + dx * dx) + dy * dy);
    return nrm * nrm;
}

__device__ __forceinline__ float exp_rn(float a) { return (float)exp((double)a); }

// per-slice shape parameter: v2 -> "rad" of slice z; v1 -> "sig" of slice z
__device__ __forceinline__ float slice_shape(int kind, const OpParams& o, int z) {
    if (kind == SN_KIND_ARROW_V2) {
        if (z >= o.ch) return o.radius;
        float inc = fminf(fmaxf(o.cone_inc, 0.f), 0.499f);
        const float tn = (float)tan((double)(inc * kPiF));
        return (o.cone_radius * (float)z) * tn;  // cone_radius*h*tan(inc*pi), h = z (arrow.py:247)
    }
    if (kind == SN_KIND_CONE_V1) {
        if (z >= o.ch) return o.sigma;
        const int h = o.ch - 1 - z;  // slices are prepended for h = 0..ch-1 (arrow.py:191-196)
        const float sn_ = (float)sin((double)((o.cone_inc * kPiF) / (float)(2 + h)));
        return o.cone_radius * sn_;
    }
    return (kind == SN_KIND_CYLINDER_V2 || kind == SN_KIND_NEGSPHERE_V2) ? o.radius : o.sigma;
}

// raw (pre zero-sum) value of tap t, float32 like the reference
__device__ __forceinline__ float raw_value(int kind, const OpParams& o, int t, int kz, int kx, int ky) {
    const bool plane = is_plane_kind(kind);
    const float d2 = tap_d2(plane, t, kz, kx, ky);
    const int z = t / (kx * ky);
    if (is_v2(kind)) {
        const float rad = plane ? slice_shape(kind, o, z) : o.radius;
        const float rp = rad + kEpsV2;
        const float c = -1.f / (2.f * (rp * rp));
        const float e = o.sigma * exp_rn((d2 * d2) * c);
        return kind == SN_KIND_NEGSPHERE_V2 ? (-o.neg_factor) * e : e;
    }
    const float sig = plane ? slice_shape(kind, o, z) : o.sigma;
    const float u = d2 - o.radius * o.radius;
    const float c = -1.f / (2.f * (sig * sig));
    return exp_rn((u * u) * c);
}

// The CTA is split into up to kSynthGroups groups of kSynthThreads threads; each group synthesises (or
// differentiates) its own operators concurrently and synchronises on its own named barrier.
constexpr int kSynthGroups = 4;
__device__ __forceinline__ void group_sync(int group) {
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(kSynthThreads) : "memory");
}
__device__ __forceinline__ double group_sum(double v, double* red, int group) {
    v = warp_sum(v);
    const int w = (threadIdx.x % kSynthThreads) >> 5, l = threadIdx.x & 31;
    group_sync(group);
    if (l == 0) red[w] = v;
    group_sync(group);
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kSynthThreads / 32; ++i) s += red[i];
    return s;
}

// N sums at once: the same per-value tree as group_sum (warp shuffles, then the warps in order) behind ONE pair of barriers
template <int N>
__device__ __forceinline__ void group_sum_vec(double (&v)[N], double* red /*[N][warps]*/, int group) {
    constexpr int NW = kSynthThreads / 32;
    const int w = (threadIdx.x % kSynthThreads) >> 5, l = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = warp_sum(v[k]);
    group_sync(group);
    if (l == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) red[k * NW + w] = v[k];
    }
    group_sync(group);
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < NW; ++i) s += red[k * NW + i];
        v[k] = s;
    }
}

__device__ float lambda_eff_of(const SynthArgs& a, int g) {
    if (a.d.lambda_index[g] < 0) return 1.f;
    if (g != a.d.last_lambda) return *a.p[a.d.lambda_index[g]];
    float s = 0.f;  // python: sum(values) starts from 0 and adds left to right in float32
    for (int i = 0; i < a.d.n_geneos; ++i) s = s + *a.p[a.d.lambda_index[a.d.lambda_sum_order[i]]];
    return (1.f - s) + *a.p[a.d.lambda_index[g]];
}

// =====================================================================================
__global__ void __launch_bounds__(kSynthThreads * kSynthGroups, 1)
synth_fwd_kernel(const __grid_constant__ SynthArgs a, float* K, float* __restrict__ lambda_eff,
                 float* __restrict__ Kstar, float* __restrict__ snapshot, int write_last) {
    extern __shared__ float s_raw_all[];  // [groups][Tp]
    __shared__ float s_off_all[kSynthGroups][64];  // per-slice mean (plane kinds) or [0] = volume offset
    __shared__ double s_red_all[kSynthGroups][kSynthThreads / 32];
    __shared__ float s_lam[SN_MAX_GENEOS];

    const int kz = a.d.kz, kx = a.d.kx, ky = a.d.ky, P = kx * ky, T = kz * P, Tp = (T + 31) & ~31;
    const int tid = threadIdx.x, group = tid / kSynthThreads, lt = tid % kSynthThreads, warp = lt >> 5, lane = tid & 31;
    const int ngroups = blockDim.x / kSynthThreads;
    const bool observer = a.d.lambda_index[0] >= 0;
    float* s_raw = s_raw_all + group * Tp;
    float* s_off = s_off_all[group];
    double* s_red = s_red_all[group];

    if (tid < a.d.n_geneos) s_lam[tid] = lambda_eff_of(a, tid);
    if (snapshot && tid < a.d.n_param_ptrs) snapshot[tid] = *a.p[tid];
    __syncthreads();

    for (int g = group; g < a.d.n_geneos; g += ngroups) {
        const int kind = a.d.kind[g];
        const OpParams o = load_op(a, g);
        for (int t = lt; t < T; t += kSynthThreads) s_raw[t] = raw_value(kind, o, t, kz, kx, ky);
        group_sync(group);
        if (is_plane_kind(kind)) {
            // zero-sum per z slice: f - sum(f)/(kx*ky)   (cylinder.py:81-82, arrow.py:167-168)
            for (int z = warp; z < kz; z += kSynthThreads / 32) {
                double s = 0.0;
                for (int n = lane; n < P; n += 32) s += (double)s_raw[z * P + n];
                s = warp_sum(s);
                if (lane == 0) s_off[z] = (float)s / (float)P;
            }
            group_sync(group);
        } else {
            double s = 0.0;
            for (int t = lt; t < T; t += kSynthThreads) s += (double)s_raw[t];
            s = group_sum(s, s_red, group);
            if (lt == 0) {
                if (kind == SN_KIND_NEGSPHERE_V2)
                    s_off[0] = ((float)s + o.neg_factor) / (float)T;  // sum_negfactor (neg_sphere.py:181-182)
                else
                    s_off[0] = (float)s / (float)T;  // sum_zero (neg_sphere.py:126-127)
            }
            group_sync(group);
        }
        for (int t = lt; t < T; t += kSynthThreads) {
            float k;
            if (is_plane_kind(kind))
                k = s_raw[t] - s_off[t / P];
            else if (kind == SN_KIND_NEGSPHERE_V2)
                k = s_raw[t] - s_off[0];
            else
                k = (s_raw[t] - s_off[0]) - o.neg_factor;  // neg_sphere.py:148
            K[(size_t)g * T + t] = k;
        }
        group_sync(group);
    }
    if (observer) {
        __threadfence_block();
        __syncthreads();  // every operator's kernel is in K now
        if (Kstar) {
            for (int t = tid; t < T; t += blockDim.x) {
                double acc = 0.0;  // Kstar = sum_g lambda_eff[g] * K_g, float64, fixed order
                for (int g = 0; g < a.d.n_geneos; ++g) acc += (double)s_lam[g] * (double)K[(size_t)g * T + t];
                Kstar[t] = (float)acc;
            }
        }
        if (lambda_eff && tid < a.d.n_geneos) lambda_eff[tid] = s_lam[tid];
        if (write_last && tid == 0 && a.d.last_lambda >= 0)
            *const_cast<float*>(a.p[a.d.lambda_index[a.d.last_lambda]]) = s_lam[a.d.last_lambda];
    }
}

// d raw_t / d parameter k of operator g (its alphabetical parameter list), before the zero-sum projection; v[k] = 0 for the
// parameters tap t does not depend on
__device__ __forceinline__ void tap_jacobian(int kind, const OpParams& o, bool plane, int t, int z, int kz, int kx, int ky,
                                             double (&v)[5]) {
#pragma unroll
    for (int k = 0; k < 5; ++k) v[k] = 0.0;
    const double d2 = (double)tap_d2(plane, t, kz, kx, ky);
    if (is_v2(kind)) {
        const float radf = plane ? slice_shape(kind, o, z) : o.radius;
        const double rp = (double)(radf + kEpsV2);
        const double d4 = d2 * d2;
        const double E = exp(-d4 / (2.0 * rp * rp));
        const double sig = (double)o.sigma;
        // raw = sigma*E (x -nf for the sphere); d raw/d rad = raw * d4 / rp^3
        double draw_drad = sig * E * d4 / (rp * rp * rp);
        double draw_dsig = E;
        if (kind == SN_KIND_CYLINDER_V2) {
            v[0] = draw_drad;
            v[1] = draw_dsig;
        } else if (kind == SN_KIND_NEGSPHERE_V2) {
            const double nf = (double)o.neg_factor;
            v[0] = (-sig * E);
            v[1] = (-nf) * draw_drad;
            v[2] = (-nf) * draw_dsig;
        } else {  // arrow: apex, cone_inc, cone_radius, radius, sigma
            v[4] = draw_dsig;
            if (z >= o.ch) {
                v[3] = draw_drad;
            } else {
                const bool open = (o.cone_inc >= 0.f) && (o.cone_inc <= 0.499f);  // clamp gate (arrow.py:244)
                const float incc = fminf(fmaxf(o.cone_inc, 0.f), 0.499f);
                const double tn = tan((double)(incc * kPiF));
                const double h = (double)z;
                v[2] = draw_drad * (h * tn);
                if (open) v[1] = draw_drad * ((double)o.cone_radius * h * (double)kPiF * (1.0 + tn * tn));
            }
        }
    } else {
        const float sigf = plane ? slice_shape(kind, o, z) : o.sigma;
        const double sg = (double)sigf, r = (double)o.radius;
        const double u = d2 - r * r;
        const double E = exp(-(u * u) / (2.0 * sg * sg));
        const double draw_drad = E * 2.0 * u * r / (sg * sg);
        const double draw_dsig = E * (u * u) / (sg * sg * sg);
        if (kind == SN_KIND_CYLINDER_V1) {
            v[0] = draw_drad;
            v[1] = draw_dsig;
        } else if (kind == SN_KIND_NEGSPHERE_V1) {
            v[1] = draw_drad;
            v[2] = draw_dsig;
        } else {  // cone v1
            v[3] = draw_drad;
            if (z >= o.ch) {
                v[4] = draw_dsig;
            } else {
                const int h = o.ch - 1 - z;
                const double ang = (double)((o.cone_inc * kPiF) / (float)(2 + h));
                v[2] = draw_dsig * sin(ang);
                v[1] = draw_dsig * ((double)o.cone_radius * cos(ang) * (double)kPiF / (double)(2 + h));
            }
        }
    }
}

// =====================================================================================
// Backward.  MODE 0: dK given [G,T] (double).  MODE 1: dK_g = lambda_eff[g] * W (observer),
// plus dlambda_g = <K_g - K_last, W>.
// =====================================================================================
// MODE 2 = MODE 1 followed by the all-reduce (sum over ranks) of dparams over NVLink peer memory, by this CTA itself
// (peer_exchange.cuh): the exchange starts the moment the gradients exist — no second launch, no launch gap on the
// critical path of a multi-GPU step.
template <int MODE>
__global__ void __launch_bounds__(kSynthThreads * kSynthGroups, 1)
synth_bwd_kernel(const __grid_constant__ SynthArgs a, const double* __restrict__ dK, const float* __restrict__ K,
                 const float* __restrict__ lambda_eff, const double* __restrict__ W, double scale,
                 float* __restrict__ dparams, const __grid_constant__ PeerArgs peer) {
    constexpr bool kObserver = MODE >= 1;
    extern __shared__ double s_W[];  // observer modes: the tap gradient, loaded ONCE (it is read by three dependent phases)
    __shared__ double s_mean_all[kSynthGroups][64];
    __shared__ double s_red_all[kSynthGroups][6 * (kSynthThreads / 32)];
    const int kz = a.d.kz, kx = a.d.kx, ky = a.d.ky, P = kx * ky, T = kz * P;
    const int group = threadIdx.x / kSynthThreads, ngroups = blockDim.x / kSynthThreads;
    const int tid = threadIdx.x % kSynthThreads, warp = tid >> 5, lane = tid & 31;  // tid: within the group
    double* s_mean = s_mean_all[group];
    double* s_red = s_red_all[group];

    for (int i = threadIdx.x; i < a.d.n_param_ptrs; i += blockDim.x) dparams[i] = 0.f;
    if (kObserver)
        for (int t = threadIdx.x; t < T; t += blockDim.x) s_W[t] = W[t];
    __syncthreads();

    for (int g = group; g < a.d.n_geneos; g += ngroups) {
        const int kind = a.d.kind[g];
        const OpParams o = load_op(a, g);
        const bool plane = is_plane_kind(kind);
        const double lam = kObserver ? (double)lambda_eff[g] : 1.0;
        auto dk = [&](int t) -> double { return kObserver ? lam * s_W[t] : dK[(size_t)g * T + t]; };

        // projection D = dK - mean(dK) over the zero-sum group (slice or whole volume)
        double vol_mean = 0.0;
        if (plane) {
            for (int z = warp; z < kz; z += kSynthThreads / 32) {
                double s = 0.0;
                for (int n = lane; n < P; n += 32) s += dk(z * P + n);
                s = warp_sum(s);
                if (lane == 0) s_mean[z] = s / (double)P;
            }
            group_sync(group);
        } else {
            double s = 0.0;
            for (int t = tid; t < T; t += kSynthThreads) s += dk(t);
            vol_mean = group_sum(s, s_red, group) / (double)T;
        }

        // acc[0..5): indexed like the operator's alphabetical parameter list; acc[5]: <K_g - K_last, W> (SCENE_Net.py:329-335)
        double acc[6] = {0, 0, 0, 0, 0, 0};
        const bool want_lambda = kObserver && a.d.lambda_index[g] >= 0 && g != a.d.last_lambda;
        const int last = a.d.last_lambda;
        for (int t = tid; t < T; t += kSynthThreads) {
            const int z = t / P;
            const double D = dk(t) - (plane ? s_mean[z] : vol_mean);
            double v[5];
            tap_jacobian(kind, o, plane, t, z, kz, kx, ky, v);
#pragma unroll
            for (int k = 0; k < 5; ++k) acc[k] += D * v[k];
            if (want_lambda) {
                double kd = (double)K[(size_t)g * T + t];
                if (last >= 0) kd -= (double)K[(size_t)last * T + t];
                acc[5] += kd * s_W[t];
            }
        }
        group_sum_vec<6>(acc, s_red, group);
        const int np = n_params_of(kind);
        if (tid == 0) {
            for (int k = 0; k < np; ++k) {
                double s = acc[k];
                if (kind == SN_KIND_NEGSPHERE_V2 && k == 0) s += -vol_mean;             // direct -nf/T term
                if (kind == SN_KIND_NEGSPHERE_V1 && k == 0) s += -vol_mean * (double)T;  // direct -nf term
                dparams[a.d.param_index[g] + k] = (float)(s * scale);
            }
            if (want_lambda) dparams[a.d.lambda_index[g]] = (float)(acc[5] * scale);
        }
        group_sync(group);
    }
    if constexpr (MODE == 2) {
        __syncthreads();  // every group's gradients are in dparams (written by this CTA: visible after the barrier)
        peer_exchange(peer, dparams, a.d.n_param_ptrs);
    }
}

static int check_desc(const sn_model_desc* d, const float* const* p) {
    if (!d || !p) return SN_ERR_BAD_ARG;
    if (d->n_geneos < 1 || d->n_geneos > SN_MAX_GENEOS) return SN_ERR_BAD_ARG;
    if (d->kz < 1 || d->kx < 1 || d->ky < 1 || d->kz > 64) return SN_ERR_BAD_ARG;
    if ((long long)d->kz * d->kx * d->ky > SN_MAX_TAPS) return SN_ERR_UNSUPPORTED;
    if (d->n_param_ptrs < 1 || d->n_param_ptrs > SN_MAX_PARAM_PTRS) return SN_ERR_BAD_ARG;
    const bool observer = d->lambda_index[0] >= 0;
    for (int g = 0; g < d->n_geneos; ++g) {
        if (d->kind[g] < 0 || d->kind[g] > SN_KIND_NEGSPHERE_V2) return SN_ERR_BAD_ARG;
        const int np = d->kind[g] <= SN_KIND_CYLINDER_V2 ? 2 : (d->kind[g] <= SN_KIND_ARROW_V2 ? 5 : 3);
        if (d->param_index[g] < 0 || d->param_index[g] + np > d->n_param_ptrs) return SN_ERR_BAD_ARG;
        if (observer) {
            if (d->lambda_index[g] < 0 || d->lambda_index[g] >= d->n_param_ptrs) return SN_ERR_BAD_ARG;
            if (d->lambda_sum_order[g] < 0 || d->lambda_sum_order[g] >= d->n_geneos) return SN_ERR_BAD_ARG;
        }
    }
    if (observer && (d->last_lambda < -1 || d->last_lambda >= d->n_geneos)) return SN_ERR_BAD_ARG;
    for (int i = 0; i < d->n_param_ptrs; ++i)
        if (!p[i]) return SN_ERR_BAD_ARG;
    return SN_OK;
}

static void fill_args(SynthArgs& a, const sn_model_desc* d, const float* const* p) {
    a.d = *d;
    for (int i = 0; i < SN_MAX_PARAM_PTRS; ++i) a.p[i] = i < d->n_param_ptrs ? p[i] : nullptr;
}

}  // namespace sn

extern "C" int sn_geneo_synth_fwd(const sn_model_desc* desc, const float* const* param_ptrs_host, float* K,
                                  float* lambda_eff, float* Kstar, float* param_snapshot, int write_last_lambda,
                                  void* stream) {
    int rc = sn::check_desc(desc, param_ptrs_host);
    if (rc) return rc;
    if (!K) return SN_ERR_BAD_ARG;
    sn::SynthArgs a;
    sn::fill_args(a, desc, param_ptrs_host);
    const int groups = desc->n_geneos < sn::kSynthGroups ? desc->n_geneos : sn::kSynthGroups;
    const int Tp = (desc->kz * desc->kx * desc->ky + 31) & ~31;
    const size_t smem = (size_t)groups * Tp * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(sn::synth_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return sn::cuda_rc(e);
    }
    sn::synth_fwd_kernel<<<1, sn::kSynthThreads * groups, smem, (cudaStream_t)stream>>>(a, K, lambda_eff, Kstar, param_snapshot, write_last_lambda);
    SN_LAUNCH_CHECK();
    return SN_OK;
}

extern "C" int sn_geneo_synth_bwd(const sn_model_desc* desc, const float* const* param_ptrs_host, const double* dK,
                                  float* dparams, void* stream) {
    int rc = sn::check_desc(desc, param_ptrs_host);
    if (rc) return rc;
    if (!dK || !dparams) return SN_ERR_BAD_ARG;
    sn::SynthArgs a;
    sn::fill_args(a, desc, param_ptrs_host);
    const int groups = desc->n_geneos < sn::kSynthGroups ? desc->n_geneos : sn::kSynthGroups;
    sn::synth_bwd_kernel<0><<<1, sn::kSynthThreads * groups, 0, (cudaStream_t)stream>>>(a, dK, nullptr, nullptr, nullptr, 1.0, dparams, sn::PeerArgs{});
    SN_LAUNCH_CHECK();
    return SN_OK;
}

extern "C" int sn_scenenet_param_grads(const sn_model_desc* desc, const float* const* param_ptrs_host, const float* K,
                                       const float* lambda_eff, const double* W, double scale, float* dparams,
                                       void* stream) {
    int rc = sn::check_desc(desc, param_ptrs_host);
    if (rc) return rc;
    if (!K || !lambda_eff || !W || !dparams || desc->lambda_index[0] < 0) return SN_ERR_BAD_ARG;
    sn::SynthArgs a;
    sn::fill_args(a, desc, param_ptrs_host);
    const int groups = desc->n_geneos < sn::kSynthGroups ? desc->n_geneos : sn::kSynthGroups;
    const size_t smem = (size_t)desc->kz * desc->kx * desc->ky * sizeof(double);  // <= SN_MAX_TAPS * 8 = 32 KB
    sn::synth_bwd_kernel<1><<<1, sn::kSynthThreads * groups, smem, (cudaStream_t)stream>>>(a, nullptr, K, lambda_eff, W, scale, dparams, sn::PeerArgs{});
    SN_LAUNCH_CHECK();
    return SN_OK;
}

extern "C" int sn_scenenet_param_grads_allreduce(const sn_model_desc* desc, const float* const* param_ptrs_host, const float* K,
                                                 const float* lambda_eff, const double* W, double scale, float* dparams,
                                                 int rank, int world, const uint64_t* peer_bufs_host, uint32_t* seq_counter,
                                                 int32_t* status, int64_t timeout_ms, void* stream) {
    int rc = sn::check_desc(desc, param_ptrs_host);
    if (rc) return rc;
    if (!K || !lambda_eff || !W || !dparams || desc->lambda_index[0] < 0) return SN_ERR_BAD_ARG;
    sn::PeerArgs peer;
    rc = sn::fill_peer_args(peer, rank, world, peer_bufs_host, seq_counter, status, timeout_ms);
    if (rc) return rc;
    sn::SynthArgs a;
    sn::fill_args(a, desc, param_ptrs_host);
    // the exchange needs one warp per rank: at least 32 * world threads
    int groups = desc->n_geneos < sn::kSynthGroups ? desc->n_geneos : sn::kSynthGroups;
    while (sn::kSynthThreads * groups < 32 * world) ++groups;
    const size_t smem = (size_t)desc->kz * desc->kx * desc->ky * sizeof(double);
    sn::synth_bwd_kernel<2><<<1, sn::kSynthThreads * groups, smem, (cudaStream_t)stream>>>(a, nullptr, K, lambda_eff, W, scale, dparams, peer);
    SN_LAUNCH_CHECK();
    return SN_OK;
}
