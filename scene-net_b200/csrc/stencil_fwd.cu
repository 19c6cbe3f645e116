// Observer forward: pred = relu(tanh(conv3d_same(x, Kstar))) as a direct 3-D stencil.
// Replaces SCENE_Net.py:209-226 / 322-339 (F.conv3d with G kernels + convex combination +
// relu(tanh)): by linearity the G convolutions collapse into ONE T-tap stencil with
// Kstar = sum_g lambda_g K_g (built by synth.cu).
//
// One 8 x TX x TY output tile at a time per CTA.  The halo tile is staged in shared memory by a
// single TMA load (cp.async.bulk.tensor.4d; out-of-bounds -> 0 == 'same' zero padding), each
// thread keeps an 8 x 4 register block of outputs and walks the taps with a runtime loop over
// dx and fully unrolled (z-chunk x KY) bodies.  Bound: FP32 pipe (2*T flop/voxel vs 8 B/voxel).
#include <stdlib.h>
#include "stencil_common.cuh"

namespace sn {
int stencil_fwd_ky3(const FwdParams&, cudaStream_t);
int stencil_fwd_ky5(const FwdParams&, cudaStream_t);
int stencil_fwd_ky6(const FwdParams&, cudaStream_t);
int stencil_fwd_ky7(const FwdParams&, cudaStream_t);
int stencil_fwd_ky9(const FwdParams&, cudaStream_t);
int stencil_fwd_ky11(const FwdParams&, cudaStream_t);
int stencil_fwd_ky13(const FwdParams&, cudaStream_t);
int stencil_fwd_ky15(const FwdParams&, cudaStream_t);
int stencil_fwd_generic(const FwdParams& p, int ky, cudaStream_t stream);  // stencil_generic.cu
// stencil_fwd_sparse.cu
bool fwd_sparse_supported(int B, int Z, int X, int Y, int kz, int kx, int ky);
bool fwd_tile_handoff_supported(int B, int Z, int X, int Y, int kz, int kx, int ky);
int fwd_sparse_launch(const float* x, const float* Kstar, void* pred, int out_f64, unsigned long long* state,
                      const unsigned long long* gate, unsigned long long nnz_max, unsigned long long dw_max, bool handoff,
                      int B, int Z, int X, int Y, int kz, int kx, int ky, int nq, cudaStream_t stream);
}  // namespace sn

// occupancy (percent of the voxels) up to which the occupancy-driven forward is selected: measured break-even on B200
// with the mask-driven kernel and float64 predictions (scratch/occ_check.py, profiles/r1_notes.md): (9,5,5) on 64^3
// 66 us at 1.6 %, 85 us at 3 %, 114 us at 5 % against 92 us dense; (9,7,7) 101 us at 1.6 % against 183 us; 9^3 on 128^3
// 231 us against 571 us.  (The scanning kernel without the occupancy bits broke even at 1.3 %.)
static unsigned long long fwd_sparse_nnz_max(long long nvox, int kx, int ky) {
    static const double forced = SN_ENV("SN_SPARSE_FWD_PCT") ? atof(SN_ENV("SN_SPARSE_FWD_PCT")) : -1.0;
    const double pct = forced >= 0.0 ? forced : (kx * ky <= 64 ? 3.0 : 4.0);
    return (unsigned long long)((double)nvox * pct / 100.0);
}

// Clustering bound: mask words with >= 8 of their 32 voxels occupied that a grid may hold and still go to the
// occupancy-driven kernel.  A uniformly sparse grid has none (Bernoulli 3 %: 7e-6 of the words); a grid with a locally dense
// layer has thousands.  16 words or 1/4096 of all words, whichever is larger.
static unsigned long long fwd_dense_words_max(long long nvox) {
    static const long long forced = SN_ENV("SN_SPARSE_FWD_DW") ? atoll(SN_ENV("SN_SPARSE_FWD_DW")) : -1;
    if (forced >= 0) return (unsigned long long)forced;
    const long long nw = (nvox + 31) / 32;
    return (unsigned long long)(nw / 4096 > 16 ? nw / 4096 : 16);
}

static int dense_fwd(const sn::FwdParams& p, int ky, cudaStream_t s) {
    int rc;
    switch (ky) {
        case 3: rc = sn::stencil_fwd_ky3(p, s); break;
        case 5: rc = sn::stencil_fwd_ky5(p, s); break;
        case 6: rc = sn::stencil_fwd_ky6(p, s); break;
        case 7: rc = sn::stencil_fwd_ky7(p, s); break;
        case 9: rc = sn::stencil_fwd_ky9(p, s); break;
        case 11: rc = sn::stencil_fwd_ky11(p, s); break;
        case 13: rc = sn::stencil_fwd_ky13(p, s); break;
        case 15: rc = sn::stencil_fwd_ky15(p, s); break;
        default: rc = SN_ERR_UNSUPPORTED; break;
    }
    return rc;
}

static bool fast_fwd_ky(int ky) { return ky == 3 || ky == 5 || ky == 6 || ky == 7 || ky == 9 || ky == 11 || ky == 13 || ky == 15; }

// nq observers on the same grids: Kstar [nq][T], pred [nq][B,1,Z,X,Y].  The dense stencil runs once per observer; the
// mask-driven occupancy kernel lists the non-zero voxels of a tile once for all of them.
static int fwd_impl(const float* x, const unsigned long long* nnz, int mode, const float* Kstar, int nq,
                    int B, int Z, int X, int Y, int kz, int kx, int ky, void* pred, int pred_dtype, void* stream) {
    if (!x || !Kstar || !pred) return SN_ERR_BAD_ARG;
    if (B < 1 || Z < 1 || X < 1 || Y < 1 || kz < 1 || kx < 1 || ky < 1 || nq < 1 || nq > SN_MAX_OBSERVERS) return SN_ERR_BAD_ARG;
    if (pred_dtype != SN_F32 && pred_dtype != SN_F64) return SN_ERR_BAD_ARG;
    if (mode != SN_PATH_AUTO && mode != SN_PATH_DENSE && mode != SN_PATH_SPARSE) return SN_ERR_BAD_ARG;
    if ((long long)kz * kx * ky > SN_MAX_TAPS) return SN_ERR_UNSUPPORTED;
    if (nnz && ((uintptr_t)nnz & 7)) return SN_ERR_ALIGN;
    const long long T = (long long)kz * kx * ky, nvox = (long long)B * Z * X * Y;
    const size_t esz = pred_dtype == SN_F64 ? 8 : 4;
    cudaStream_t s = (cudaStream_t)stream;
    const bool sparse_ok = sn::fwd_sparse_supported(B, Z, X, Y, kz, kx, ky);
    // sn_grid_prepare's state buffer: counters, one occupancy bit per voxel, the tile list (sn_grid_state_bytes)
    unsigned long long* state = const_cast<unsigned long long*>(nnz);
    const unsigned long long dw_max = fwd_dense_words_max(nvox);
    const unsigned long long nnz_max = fwd_sparse_nnz_max(nvox, kx, ky);
    auto pred_q = [&](int q) { return (void*)((char*)pred + (size_t)q * (size_t)nvox * esz); };
    // gate: whole-grid selection on the device (the stencil returns at once when the occupancy-driven kernel is selected);
    // tiles: compute only the tiles the occupancy-driven kernel listed
    auto dense_all = [&](const unsigned long long* gate, bool tiles, bool allow_generic) -> int {
        for (int q = 0; q < nq; ++q) {
            sn::FwdParams p{};
            p.x = x; p.Kstar = Kstar + q * T; p.pred = pred_q(q);
            p.B = B; p.Z = Z; p.X = X; p.Y = Y; p.kz = kz; p.kx = kx;
            p.out_f64 = pred_dtype == SN_F64; p.plz = sn::pad_left(kz);
            p.nnz = gate; p.nnz_max = nnz_max; p.dw_max = dw_max;
            if (tiles) {
                p.state = state;
                p.tile_list = reinterpret_cast<const int*>(reinterpret_cast<const unsigned*>(state + SN_STATE_WORDS) + sn::state_mask_words(nvox));
            }
            p.last_pass = q == nq - 1;  // (and the last z-split pass: fwd_passes)
            int rc = dense_fwd(p, ky, s);
            if (rc == SN_ERR_UNSUPPORTED && allow_generic) rc = sn::stencil_fwd_generic(p, ky, s);
            if (rc) return rc;
        }
        return SN_OK;
    };
    auto sparse_all = [&](const unsigned long long* gate, bool handoff) -> int {
        int rc = sn::fwd_sparse_launch(x, Kstar, pred, pred_dtype == SN_F64, state, gate, nnz_max, dw_max, handoff, B, Z, X, Y,
                                       kz, kx, ky, nq, s);
        if (rc != SN_ERR_UNSUPPORTED || nq == 1) return rc;
        for (int q = 0; q < nq; ++q) {  // no shared lists (no state buffer / shape outside the mask-driven kernel): one launch each
            rc = sn::fwd_sparse_launch(x, Kstar + q * T, pred_q(q), pred_dtype == SN_F64, state, gate, nnz_max, dw_max, false,
                                       B, Z, X, Y, kz, kx, ky, 1, s);
            if (rc) return rc;
        }
        return SN_OK;
    };
    if (mode == SN_PATH_SPARSE) {
        if (!sparse_ok) return SN_ERR_UNSUPPORTED;
        return sparse_all(nullptr, false);
    }
    if (mode == SN_PATH_AUTO && sparse_ok && nnz) {
        if (fast_fwd_ky(ky) && sn::fwd_tile_handoff_supported(B, Z, X, Y, kz, kx, ky)) {
            // per-tile choice: the occupancy-driven kernel walks every tile (zero-fills the empty ones, scatters the sparse
            // ones, lists the dense ones), then the stencil computes the listed tiles
            int rc = sparse_all(nullptr, true);
            if (rc) return rc;
            return dense_all(nullptr, true, false);
        }
        // whole-grid choice: both kernels are enqueued; the grid state decides on the device which one works
        int rc = dense_all(nnz, false, false);
        if (rc == SN_ERR_UNSUPPORTED)  // no dense instantiation for this width: the occupancy-driven kernel always runs
            return sparse_all(nullptr, false);
        if (rc) return rc;
        return sparse_all(nnz, false);
    }
    int rc = dense_all(nullptr, false, false);
    if (rc == SN_ERR_UNSUPPORTED) {
        if (sparse_ok && mode == SN_PATH_AUTO) return sparse_all(nullptr, false);
        rc = dense_all(nullptr, false, true);
    }
    return rc;
}

extern "C" int sn_scenenet_fwd(const float* x, const unsigned long long* nnz, int mode, const float* Kstar,
                               int B, int Z, int X, int Y, int kz, int kx, int ky, void* pred, int pred_dtype,
                               void* stream) {
    return fwd_impl(x, nnz, mode, Kstar, 1, B, Z, X, Y, kz, kx, ky, pred, pred_dtype, stream);
}

extern "C" int sn_scenenet_fwd_multi(const float* x, const unsigned long long* nnz, int mode, const float* Kstars, int n_observers,
                                     int B, int Z, int X, int Y, int kz, int kx, int ky, void* preds, int pred_dtype,
                                     void* stream) {
    return fwd_impl(x, nnz, mode, Kstars, n_observers, B, Z, X, Y, kz, kx, ky, preds, pred_dtype, stream);
}

extern "C" int sn_select_fwd_path_state(int64_t nnz, int64_t dense_words, int B, int Z, int X, int Y, int kz, int kx, int ky) {
    if (B < 1 || Z < 1 || X < 1 || Y < 1 || kz < 1 || kx < 1 || ky < 1 || nnz < 0 || dense_words < 0) return SN_ERR_BAD_ARG;
    const bool ok = sn::fwd_sparse_supported(B, Z, X, Y, kz, kx, ky);
    const bool fast = fast_fwd_ky(ky);
    if (ok && !fast) return SN_PATH_SPARSE;
    const long long nvox = (long long)B * Z * X * Y;
    if (ok && (unsigned long long)nnz <= fwd_sparse_nnz_max(nvox, kx, ky) && (unsigned long long)dense_words <= fwd_dense_words_max(nvox))
        return SN_PATH_SPARSE;  // uniformly sparse: every tile is scattered, no stencil pass needs to be enqueued
    // clustered or moderately occupied grids: the per-tile choice (SN_PATH_AUTO) as long as a good part of the tiles can be
    // expected below the scatter's break-even; otherwise the stencil for every tile
    if (ok && sn::fwd_tile_handoff_supported(B, Z, X, Y, kz, kx, ky) && nnz <= nvox / 4) return SN_PATH_AUTO;
    return SN_PATH_DENSE;
}

extern "C" int sn_select_fwd_path(int64_t nnz, int B, int Z, int X, int Y, int kz, int kx, int ky) {
    return sn_select_fwd_path_state(nnz, 0, B, Z, X, Y, kz, kx, ky);
}
