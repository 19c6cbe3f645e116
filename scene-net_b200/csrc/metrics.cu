// Confusion counts at a threshold — SURVEY §8(f) rank 2, the arithmetic behind the torchmetrics collection the
// reference updates every step (utils/scripts_utils.py:80-91: JaccardIndex / Precision / Recall / F1Score /
// FBetaScore at tau = 0.65; call sites core/lit_modules/lit_model_wrappers.py:170-171, 189-190, 199-200):
// each of the five metrics thresholds the flattened prediction (preds >= tau), compares with the integer target and
// keeps TP / FP / TN / FN (or the 2 x 2 confusion matrix) as its state.  Here ONE pass over (pred, y) yields the four
// counts for all of them.  Integer counts: atomics are exact and order-independent, so the result is bit-exact.
#include "common.cuh"

namespace sn {

constexpr int kMetThreads = 256;

template <typename TY, int V>
struct alignas(sizeof(TY) * V <= 16 ? sizeof(TY) * V : 16) YVec {
    TY v[V];
};

template <typename TP, typename TY>
__global__ void __launch_bounds__(kMetThreads) confusion_kernel(const TP* __restrict__ pred, const TY* __restrict__ y, long long n, TP tau,
                                                                unsigned long long* __restrict__ batch,
                                                                unsigned long long* __restrict__ total) {
    constexpr int V = 16 / sizeof(TP);  // predictions per 16-byte load
    struct alignas(16) PVec { TP v[V]; };
    const long long nv = n / V;
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned tp = 0, fp = 0, tn = 0, fn = 0;
    auto one = [&](TP p, TY t) {
        const bool pos = p >= tau, tgt = t != (TY)0;
        tp += pos && tgt;
        fp += pos && !tgt;
        fn += !pos && tgt;
        tn += !pos && !tgt;
    };
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < nv; i += 2 * stride) {  // two loads of each operand in flight
        const PVec p0 = reinterpret_cast<const PVec*>(pred)[i], p1 = reinterpret_cast<const PVec*>(pred)[i + stride];
        const YVec<TY, V> y0 = reinterpret_cast<const YVec<TY, V>*>(y)[i], y1 = reinterpret_cast<const YVec<TY, V>*>(y)[i + stride];
#pragma unroll
        for (int j = 0; j < V; ++j) one(p0.v[j], y0.v[j]);
#pragma unroll
        for (int j = 0; j < V; ++j) one(p1.v[j], y1.v[j]);
    }
    for (; i < nv; i += stride) {
        const PVec p0 = reinterpret_cast<const PVec*>(pred)[i];
        const YVec<TY, V> y0 = reinterpret_cast<const YVec<TY, V>*>(y)[i];
#pragma unroll
        for (int j = 0; j < V; ++j) one(p0.v[j], y0.v[j]);
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - nv * V)) one(pred[nv * V + threadIdx.x], y[nv * V + threadIdx.x]);

    __shared__ unsigned s_c[kMetThreads / 32][4];
    tp = __reduce_add_sync(0xffffffffu, tp);
    fp = __reduce_add_sync(0xffffffffu, fp);
    tn = __reduce_add_sync(0xffffffffu, tn);
    fn = __reduce_add_sync(0xffffffffu, fn);
    if ((threadIdx.x & 31) == 0) {
        s_c[threadIdx.x >> 5][0] = tp; s_c[threadIdx.x >> 5][1] = fp;
        s_c[threadIdx.x >> 5][2] = tn; s_c[threadIdx.x >> 5][3] = fn;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < kMetThreads / 32; ++w) t += s_c[w][threadIdx.x];
        if (t) {
            if (batch) atomicAdd(batch + threadIdx.x, t);
            if (total) atomicAdd(total + threadIdx.x, t);
        }
    }
}

template <typename TP, typename TY>
static int launch_confusion(const void* pred, const void* y, long long n, double tau, unsigned long long* batch,
                            unsigned long long* total, cudaStream_t s) {
    constexpr int V = 16 / sizeof(TP);
    constexpr size_t ya = sizeof(TY) * V <= 16 ? sizeof(TY) * V : 16;
    if (((uintptr_t)pred & 15) || ((uintptr_t)y & (ya - 1))) return SN_ERR_ALIGN;
    long long b = ceil_div64(n, (long long)kMetThreads * V * 4);
    const long long cap = (long long)kNumSMs * 8;
    const int grid = (int)(b < 1 ? 1 : (b > cap ? cap : b));
    confusion_kernel<TP, TY><<<grid, kMetThreads, 0, s>>>((const TP*)pred, (const TY*)y, n, (TP)tau, batch, total);
    SN_LAUNCH_CHECK();
    return SN_OK;
}

template <typename TP>
static int dispatch_y(const void* pred, const void* y, int y_dtype, long long n, double tau, unsigned long long* batch,
                      unsigned long long* total, cudaStream_t s) {
    switch (y_dtype) {
        case SN_F32: return launch_confusion<TP, float>(pred, y, n, tau, batch, total, s);
        case SN_F64: return launch_confusion<TP, double>(pred, y, n, tau, batch, total, s);
        case SN_U8: return launch_confusion<TP, unsigned char>(pred, y, n, tau, batch, total, s);
        case SN_I32: return launch_confusion<TP, int>(pred, y, n, tau, batch, total, s);
        case SN_I64: return launch_confusion<TP, long long>(pred, y, n, tau, batch, total, s);
        default: return SN_ERR_BAD_ARG;
    }
}

}  // namespace sn

extern "C" int sn_confusion_counts(const void* pred, int pred_dtype, const void* y, int y_dtype, int64_t n, double tau,
                                   unsigned long long* batch_counts, unsigned long long* total_counts, void* stream) {
    if (!pred || !y || n < 0 || (!batch_counts && !total_counts)) return SN_ERR_BAD_ARG;
    if (pred_dtype != SN_F32 && pred_dtype != SN_F64) return SN_ERR_BAD_ARG;
    if (((uintptr_t)batch_counts | (uintptr_t)total_counts) & 7) return SN_ERR_ALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    if (batch_counts) {
        cudaError_t e = cudaMemsetAsync(batch_counts, 0, 4 * sizeof(unsigned long long), s);
        if (e != cudaSuccess) return sn::cuda_rc(e);
    }
    if (n == 0) return SN_OK;
    // float32 predictions are compared in float32 ((float)tau), float64 in float64: what `preds >= tau` does in torch
    if (pred_dtype == SN_F32) return sn::dispatch_y<float>(pred, y, y_dtype, n, tau, batch_counts, total_counts, s);
    return sn::dispatch_y<double>(pred, y, y_dtype, n, tau, batch_counts, total_counts, s);
}
