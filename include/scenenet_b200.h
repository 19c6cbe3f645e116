/*
 * scenenet_b200 — C ABI of the B200-native (sm_100a) SCENE-Net hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / C++ types.  Every
 * buffer is DEVICE memory allocated by the caller (unless the name says `host`), every call
 * is asynchronous on the `stream` handed in (a `cudaStream_t` passed as `void*`; NULL = the
 * legacy default stream), nothing is allocated or freed by the library and there is no
 * global mutable state, so calls are re-entrant across host threads and streams.
 *
 * Return value: 0 on success; SN_ERR_* (< 0) for argument errors; -(1000 + cudaError_t)
 * for CUDA runtime errors at launch.
 *
 * The reference (dlavado/scene-net) is pure Python: there is no FFI in it to bind to.  Each
 * entry point below therefore cites the reference *Python* lines whose arithmetic it
 * replaces (paths relative to the reference root); INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add at exactly those lines.
 *
 * Layout conventions: voxel grids are [B,1,Z,X,Y] contiguous (Y fastest);
 * GENEO kernels are [G,kz,kx,ky] contiguous; T = kz*kx*ky taps; cross-correlation with
 * PyTorch 'same' zero padding: left pad (k-1)/2, right pad k-1-left.
 */
#ifndef SCENENET_B200_H
#define SCENENET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* v5: sn_vxg_to_xyz; the peer exchange moves {value, call number} words (sn_peer_allreduce_buffer_bytes doubled: every rank of a
 * job must load the same library version); state[1] reserved (v4: ticket of the tap gradient's last-CTA reduction) */
#define SN_ABI_VERSION 5

/* ---- error codes ------------------------------------------------------------------- */
#define SN_OK 0
#define SN_ERR_BAD_ARG (-1)      /* null pointer, non-positive size, unknown enum          */
#define SN_ERR_UNSUPPORTED (-2)  /* shape outside what the kernels are instantiated for    */
#define SN_ERR_ALIGN (-3)        /* pointer not aligned as documented                      */
#define SN_ERR_WORKSPACE (-4)    /* workspace smaller than sn_*_workspace_bytes() reports  */
#define SN_ERR_CUDA_BASE (-1000) /* return = SN_ERR_CUDA_BASE - cudaError_t                */

/* ---- dtypes of grid tensors crossing the boundary ------------------------------------ */
#define SN_F32 0
#define SN_F64 1
#define SN_U8 2  /* occupancy bytes (uint8 / bool); accepted by sn_grid_prepare and as a target dtype */
#define SN_I32 3 /* integer targets (sn_confusion_counts only) */
#define SN_I64 4
#define SN_BITS 5 /* occupancy, one BIT per voxel: bit i % 32 of 32-bit word i / 32 <-> flat voxel index i (sn_grid_prepare only;
                   * 64x fewer bytes than the float64 grids of the reference's ToFullDense, torch_transforms.py:33-40) */

/* ---- GENEO operator kinds (core/models/geneos/) -------------------------------------- */
#define SN_KIND_CYLINDER_V1 0 /* cylinder.py:30-140   cylinder_kernel   params: radius, sigma                */
#define SN_KIND_CYLINDER_V2 1 /* cylinder.py:146-176  cylinderv2        params: radius, sigma                */
#define SN_KIND_CONE_V1 2     /* arrow.py:30-205      cone_kernel       params: apex, cone_inc, cone_radius, radius, sigma */
#define SN_KIND_ARROW_V2 3    /* arrow.py:208-252     arrow             params: apex, cone_inc, cone_radius, radius, sigma */
#define SN_KIND_NEGSPHERE_V1 4 /* neg_sphere.py:29-158 neg_sphere_kernel params: neg_factor, radius, sigma   */
#define SN_KIND_NEGSPHERE_V2 5 /* neg_sphere.py:160-199 negSpherev2      params: neg_factor, radius, sigma   */

#define SN_MAX_GENEOS 16
#define SN_MAX_PARAM_PTRS 96 /* 16 operators x 5 params + 16 lambdas */
#define SN_MAX_TAPS 4096     /* up to 16^3 */
#define SN_CRIT_MAX_BINS 12  /* histogram bins of the weighted-MSE scheme (the reference uses 10) */
#define SN_CRIT_COEF (SN_CRIT_MAX_BINS + 4) /* doubles in the coefficient buffer of sn_criterion_fwd */

/*
 * Host-side description of one observer (SceneNet / SCENE_Net instance,
 * core/models/SCENE_Net.py:121-339).  `param_ptrs` (passed next to it) is a HOST array of
 * DEVICE pointers to the model's 0-dim float32 parameters (nn.Parameter storage, used in
 * place — no gather copy).  For operator g its parameters are
 * param_ptrs[param_index[g] .. +n) in ALPHABETICAL name order (nn.ParameterDict order,
 * SCENE_Net.py:83-89), its convex coefficient is param_ptrs[lambda_index[g]].
 */
typedef struct sn_model_desc {
    int32_t n_geneos;                      /* G, channel order cy_*, cone_*, neg_* (SCENE_Net.py:264-275) */
    int32_t kz, kx, ky;                    /* kernel_size (z, x, y)                                       */
    int32_t n_param_ptrs;                  /* entries in param_ptrs                                       */
    int32_t kind[SN_MAX_GENEOS];           /* SN_KIND_*                                                   */
    int32_t param_index[SN_MAX_GENEOS];    /* first parameter of operator g in param_ptrs                 */
    int32_t lambda_index[SN_MAX_GENEOS];   /* lambda_<g> in param_ptrs; -1 = no observer (GENEO_Layer use) */
    int32_t lambda_sum_order[SN_MAX_GENEOS]; /* operator ids in lambdas_dict iteration order: the float32
                                              left-to-right sum of SCENE_Net.py:331 is reproduced in it   */
    int32_t last_lambda;                   /* operator id whose coefficient is 1 - sum(others); -1 = none */
} sn_model_desc;

/* library / ABI version, for the loader's sanity check */
int sn_abi_version(void);
/* writes a static NUL-terminated build string (arch, flags) */
const char* sn_build_info(void);
/* number of kernel launches issued by this library since process start (bench `gpu_launches`) */
int64_t sn_launch_count(void);

/* ======================================================================================
 * GENEO kernel synthesis — replaces GENEO_Layer.compute_kernel (SCENE_Net.py:103-106) and
 * the compute_kernel methods of cylinder.py:85-103,162-176, arrow.py:170-205,228-252,
 * neg_sphere.py:129-158,185-199.  One launch for all G operators.
 *   K          [G,T] float32 out  — the G kernels (reference: geneo.kernel, float32)
 *   lambda_eff [G]   float32 out  — effective convex coefficients (SCENE_Net.py:329-335);
 *                                   the last one is (1 - sum(lambdas)) + lambda_last in float32
 *   Kstar      [T]   float32 out  — sum_g lambda_eff[g] * K[g] (accumulated in float64): by
 *                                   linearity of the convolution the observer's
 *                                   sum_g lambda_g * conv3d(x, K_g) equals conv3d(x, Kstar)
 *   write_last_lambda != 0: also stores lambda_eff[last] into the last-lambda parameter
 *                                   (the side effect of SCENE_Net.py:333)
 *   param_snapshot [n_param_ptrs] float32 out (NULL to skip) — the parameter values this
 *                                   forward used (read before the last-lambda write), so a
 *                                   later backward differentiates at the forward's point
 * lambda_eff / Kstar may be NULL when desc->lambda_index[0] < 0 (bare GENEO kernels).
 * ====================================================================================== */
int sn_geneo_synth_fwd(const sn_model_desc* desc, const float* const* param_ptrs_host,
                       float* K, float* lambda_eff, float* Kstar, float* param_snapshot,
                       int write_last_lambda, void* stream);

/* Jacobian^T of the synthesis: dparams[i] = sum_{g,t} dK[g,t] * dK_g[t]/dparam_i for every
 * entry of param_ptrs (0 for apex and for lambdas).  Replaces autograd through the
 * compute_kernel graphs cited above.  dK [G,T] float64, dparams [n_param_ptrs] float32 out. */
int sn_geneo_synth_bwd(const sn_model_desc* desc, const float* const* param_ptrs_host,
                       const double* dK, float* dparams, void* stream);

/* Observer backward tail: from the tap gradient W[t] = sum_{b,v} G0[b,v] * xpad[b,v+t]
 * (float64, produced by sn_scenenet_bwd) to the gradient of every parameter:
 *   dK_g = lambda_eff[g] * W;  dlambda_g = <K_g - K_last, W> (g != last);  dparams via the
 *   Jacobian^T above.  `scale` multiplies every output (1/world_size for DDP mean).
 * Replaces autograd through SCENE_Net.py:324-337.  dparams [n_param_ptrs] float32 out.     */
int sn_scenenet_param_grads(const sn_model_desc* desc, const float* const* param_ptrs_host,
                            const float* K, const float* lambda_eff, const double* W, double scale,
                            float* dparams, void* stream);

/* ======================================================================================
 * Observer forward — replaces F.conv3d + convex combination + relu(tanh) of
 * SceneNet.forward / SCENE_Net.forward (SCENE_Net.py:209-226, 322-339).
 *   x     [B,1,Z,X,Y] float32          (sn_grid_prepare converts the reference's float64 grids)
 *   nnz   DEVICE pointer to the grid state buffer of x written by sn_grid_prepare (counters, occupancy bits, tile list),
 *         or NULL
 *   mode  SN_PATH_AUTO: with nnz, the occupancy-driven kernel (cost proportional to the occupied voxels) walks the
 *         tiles first and decides PER TILE, on the device, from the tile's own occupancy: empty tiles are zero-filled,
 *         sparse tiles are scattered, tiles above the break-even occupancy are appended to the tile list in the state
 *         buffer and the dense stencil, enqueued right behind, computes exactly those (no host synchronisation; shapes
 *         without a dense instantiation or with more tiles than the list holds fall back to one device-side choice
 *         for the whole grid, sn_select_fwd_path_state).  Without nnz the dense stencil runs.
 *         SN_PATH_DENSE / SN_PATH_SPARSE force one kernel for every tile (measurement, tests).  Same pred either way up
 *         to float32 summation order.
 *   Kstar [T] float32                  (from sn_geneo_synth_fwd)
 *   pred  [B,1,Z,X,Y] out, dtype pred_dtype (SN_F32 / SN_F64): relu(tanh(conv3d_same(x, Kstar)))
 * x must be 16-byte aligned.
 * ====================================================================================== */
#define SN_PATH_AUTO 0
#define SN_PATH_DENSE 1
#define SN_PATH_SPARSE 2
int sn_scenenet_fwd(const float* x, const unsigned long long* nnz, int mode, const float* Kstar,
                    int B, int Z, int X, int Y, int kz, int kx, int ky, void* pred, int pred_dtype, void* stream);
/* Several observers on the SAME grids — SCENENetQuantile.forward (SCENE_Net.py:409-415: one SCENE_Net per quantile,
 * predictions concatenated).  Kstars [n_observers][T] float32, preds [n_observers][B,1,Z,X,Y] out.
 * The occupancy-driven kernel lists the non-zero voxels of a tile once and scatters them with every observer's taps
 * (x is read once); the dense stencil runs once per observer.  Selection as in sn_scenenet_fwd.  n_observers <= 8. */
#define SN_MAX_OBSERVERS 8
int sn_scenenet_fwd_multi(const float* x, const unsigned long long* nnz, int mode, const float* Kstars, int n_observers,
                          int B, int Z, int X, int Y, int kz, int kx, int ky, void* preds, int pred_dtype, void* stream);

/* Observer backward, data part — replaces aten::convolution_backward (weight gradient) and
 * the relu/tanh/convex-combination backward of SCENE_Net.py:325-337.
 *   G0 = dpred * (1 - pred^2) * [pred > 0];   W[t] = sum_{b,v} G0[b,v] * xpad[b, v + t]
 *   x [B,1,Z,X,Y] float32, pred / dpred in pred_dtype / dpred_dtype, W [T] float64 out.
 *   nnz: DEVICE pointer to the grid state buffer of x written by sn_grid_prepare (count at [0], occupancy bits behind the counters), or NULL.
 *        Voxel grids of point clouds are ~98 % empty: when nnz is given, an occupancy-driven
 *        kernel (cost proportional to the occupied voxels) and the dense stencil are both enqueued and the
 *        count selects ON THE DEVICE which of them does the work (sparse up to 10 % occupancy; no host
 *        synchronisation).  With NULL the dense stencil always runs.  mode: SN_PATH_AUTO / _DENSE / _SPARSE as for
 *        sn_scenenet_fwd (forced modes enqueue one kernel only).  Same W either way, up to float32
 *        summation order (the occupancy-driven kernel only skips terms that are exactly zero).
 *   ws: workspace of at least sn_scenenet_bwd_workspace_bytes(...) bytes, 16-byte aligned.
 * Deterministic: fixed partition, fixed-order float64 reduction, no floating-point atomics. */
int64_t sn_scenenet_bwd_workspace_bytes(int B, int Z, int X, int Y, int kz, int kx, int ky);
int sn_scenenet_bwd(const float* x, const unsigned long long* nnz, int mode, const void* pred, int pred_dtype,
                    const void* dpred, int dpred_dtype, int B, int Z, int X, int Y, int kz, int kx, int ky,
                    double* W, void* ws, int64_t ws_bytes, void* stream);
/* The selection rule of SN_PATH_AUTO for a host that already knows the occupancy (a step captured once for replay on
 * grids of one kind can then enqueue only the kernel that will work; both kernels are correct at ANY occupancy, the
 * choice only affects speed).  which: 0 = forward, 1 = tap gradient.  Returns SN_PATH_DENSE / SN_PATH_SPARSE (forward
 * also SN_PATH_AUTO, see sn_select_fwd_path_state). */
int sn_select_path(int which, int64_t nnz, int B, int Z, int X, int Y, int kz, int kx, int ky);
int sn_select_fwd_path(int64_t nnz, int B, int Z, int X, int Y, int kz, int kx, int ky);
/* The forward's rule with the clustering statistic of the state buffer (state[2] = mask words with >= 8 of 32 voxels
 * occupied): SN_PATH_SPARSE for a uniformly sparse grid (every tile is scattered; no stencil pass is needed),
 * SN_PATH_AUTO (the per-tile choice of sn_scenenet_fwd) for a sparse but CLUSTERED grid (a locally dense layer, e.g. the
 * ground of a LiDAR scan) or a moderately occupied one (<= 25 %), SN_PATH_DENSE above that or where tiles cannot be
 * handed over. */
int sn_select_fwd_path_state(int64_t nnz, int64_t dense_words, int B, int Z, int X, int Y, int kz, int kx, int ky);

/* The two passes of sn_scenenet_bwd, callable on their own (measurement, fused criterions that
 * produce G0 themselves):  G0 [n] float32 = dpred * (1 - pred^2) * [pred > 0] evaluated in float64 and
 * rounded once;  tap gradient W[t] = sum G0 * xpad from a precomputed G0.
 * sn_scenenet_bwd's workspace = G0 (n*4 bytes rounded up to 256) followed by the tap-gradient workspace.
 * mode: SN_TAPGRAD_AUTO (device-side selection from nnz; dense when nnz is NULL), SN_TAPGRAD_DENSE,
 *       SN_TAPGRAD_SPARSE (forced, nnz ignored; measurement and tests). */
#define SN_TAPGRAD_AUTO SN_PATH_AUTO
#define SN_TAPGRAD_DENSE SN_PATH_DENSE
#define SN_TAPGRAD_SPARSE SN_PATH_SPARSE
int sn_scenenet_g0(const void* pred, int pred_dtype, const void* dpred, int dpred_dtype, int64_t n, float* g0,
                   void* stream);
int64_t sn_scenenet_tapgrad_workspace_bytes(int B, int Z, int X, int Y, int kz, int kx, int ky);
int sn_scenenet_tapgrad(const float* x, const float* g0, const unsigned long long* nnz, int mode,
                        int B, int Z, int X, int Y, int kz, int kx, int ky,
                        double* W, void* ws, int64_t ws_bytes, void* stream);

/* ======================================================================================
 * Fused criterion — replaces GENEO_Tversky_Loss.forward
 * (core/criterions/geneo_loss.py:145-161) = WeightedMSE.forward (w_mse.py:114-151) +
 * FocalTverskyLoss.forward (tversky_loss.py:81-95) + cvx_loss / positive_regularizer
 * (geneo_loss.py:36-71), and their autograd backward.
 *
 *   loss = mean(mse_weight * w(y) * (y - p)^2) + (1 - Tv)^gamma,
 *   w(y) = w_raw[bin(y)] / mean(w_raw[bin(y)]) in float32 (bin = first argmin_k |y - ranges[k]|),
 *   Tv = (TP + s) / (TP + alpha FP + beta FN + s),  TP = sum p y, FP = sum (1-y) p, FN = sum y (1-p).
 * ranges_host / w_raw_host: HOST arrays of nbins floats (the 10-entry table derived from the
 * histogram pickle: w_raw[k] = max(1 - weight_alpha * dens_k, weight_epsilon), float32).
 * pred / y: DEVICE, dtype SN_F32 or SN_F64 (both the same), 16-byte aligned.
 * terms: bit 0 = weighted-MSE term, bit 1 = focal-Tversky term (WeightedMSE alone = 1, FocalTverskyLoss alone = 2).
 * sn_criterion_fwd: one pass over (pred, y) + a one-warp finalisation; loss [1] and coef [SN_CRIT_COEF]
 *   are DEVICE doubles; coef holds the closed-form backward's scalars (and, at [MAX+2], [MAX+3], the two terms).
 * sn_criterion_bwd: one elementwise pass; grad_out = DEVICE scalar of the tensors' dtype (NULL = 1):
 *   out_g0 == 0: out = dL/dpred [n] in the tensors' dtype;
 *   out_g0 != 0: out = G0 [n] float32 = dL/dpred * (1 - pred^2) * [pred > 0], ready for sn_scenenet_tapgrad.
 * Deterministic (no floating-point atomics).
 * ====================================================================================== */
int64_t sn_criterion_workspace_bytes(int64_t n);
int sn_criterion_fwd(const void* pred, const void* y, int dtype, int64_t n, const float* ranges_host,
                     const float* w_raw_host, int nbins, float mse_weight, double tversky_alpha,
                     double tversky_beta, double focal_gamma, double tversky_smooth, int terms, double* loss,
                     double* coef, void* ws, int64_t ws_bytes, void* stream);
int sn_criterion_bwd(const void* pred, const void* y, int dtype, int64_t n, const float* ranges_host,
                     const float* w_raw_host, int nbins, const double* coef, const void* grad_out,
                     void* out, int out_g0, void* stream);
/* Penalties on the live parameters (geneo_loss.py:36-71), float32, accumulated left to right like python's
 * sum().  param_ptrs_host: HOST array of n DEVICE pointers to 0-dim float32 parameters; role_host[i]:
 * 0 = relu(-v) regulariser term, 2 = free convex coefficient (relu(-v), and part of the coefficient sum),
 * 1 = the frozen last coefficient (contributes relu(-(1 - sum(all coefficients) + itself))).
 * out (DEVICE, 2 + n floats): [0] = weight * cvx_loss, [1] = weight * positive_regularizer,
 * [2 + i] = d(out[0] + out[1]) / d param_i.
 * loss_accum (nullable): DEVICE double, e.g. the loss written by sn_criterion_fwd earlier on the same stream; the two
 * penalties are added to it in the reference's order, (loss + out[0]) + out[1] (geneo_loss.py:161), so that the
 * whole GENEO_Tversky_Loss value is produced by the library's three launches. */
int sn_param_penalty(const float* const* param_ptrs_host, const int32_t* role_host, int n, float weight,
                     float* out, double* loss_accum, void* stream);

/* ======================================================================================
 * elementwise helpers
 * ====================================================================================== */
/* Grid preparation, one HBM pass: x (SN_F64 as handed over by the reference's ToTensor, torch_transforms.py:13;
 * SN_U8 occupancy bytes; SN_F32) -> float32 copy x32 for the TMA-fed stencils (SN_F32: x32 must be NULL or x,
 * nothing is copied) and the grid STATE buffer `nnz` the forward / backward take:
 *   SN_STATE_WORDS uint64 counters — [0] number of non-zero voxels, [1] reserved (a ticket counter of the tap-gradient
 *   kernels up to ABI v4.0; the partial rows are summed by a kernel of their own now), [2] number of 32-voxel mask words with >= 8 voxels occupied (how clustered the
 *   grid is), [3] number of tiles the occupancy-driven forward handed to the dense stencil, [4] number of non-zero voxels
 *   whose value is not 1 (0 <=> an occupancy grid: the forward then takes its values from the mask bits), [5] tile
 *   counter of the occupancy-driven forward (dynamic tile scheduling), [6] CTA-done counter of the dense stencil's
 *   tile-list pass, [7] reserved; the kernels that use [1], [3], [5], [6] leave them at zero again, so one call serves
 *   any number of forwards / backwards that follow each other on a stream —
 *   then from byte 8 * SN_STATE_WORDS one occupancy BIT per voxel (bit i % 32 of 32-bit word i / 32 <-> flat voxel
 *   index i) + 4 padding words: the occupancy-driven forward lists the non-zero voxels of a halo box from these words
 *   instead of scanning floats — then the tile list ([3] entries, 32-bit tile ids; ABI v4).
 * x may also be SN_BITS (packed occupancy, ceil(n / 32) words): x32 receives 0 / 1 and the words become the state's mask.
 * nnz: DEVICE buffer of sn_grid_state_bytes(n) bytes, 16-byte aligned; counters zeroed and bits written by the call.
 * x and x32 16-byte aligned. */
#define SN_STATE_WORDS 8
int64_t sn_grid_state_bytes(int64_t n);
int sn_grid_prepare(const void* x, int dtype, int64_t n, float* x32, unsigned long long* nnz, void* stream);
/* float64 -> float32 (callers hand float64 grids: torch_transforms.py:13) */
int sn_cast_f64_to_f32(const double* in, float* out, int64_t n, void* stream);
/* uint8 / bool occupancy grids (what ToFullDense produces, one byte per voxel) -> float32; 16-byte aligned */
int sn_cast_u8_to_f32(const unsigned char* in, float* out, int64_t n, void* stream);
/* prob_to_label (utils/voxelization.py:304-323) / SCENE_Net_Class.forward (SCENE_Net.py:465-466):
 * out = (p >= tau) as 0/1, same dtype as p. */
int sn_threshold(const void* p, int dtype, double tau, int64_t n, void* out, void* stream);
/* vxg_to_xyz (utils/voxelization.py:328-360): a voxel grid [d0, d1, d2] (SN_F32 / SN_F64 / SN_U8, DEVICE) as a raw point
 * cloud: out [d0*d1*d2, 4] float64 (DEVICE, 16-byte aligned), row i = (origin + (i0, i1, i2) * voxel_size, vxg[i0, i1, i2]) in
 * C order of the index — every voxel, as the reference does (callers select label == 1 themselves).  origin, voxel_size:
 * HOST double[3] (the reference's defaults are (0, 0, 0) and (1, 1, 1)). */
int sn_vxg_to_xyz(const void* vxg, int dtype, int d0, int d1, int d2, const double* origin, const double* voxel_size,
                  double* out, void* stream);

/* ======================================================================================
 * Metric state — replaces the update of the torchmetrics 0.9.0 collection the reference
 * builds in utils/scripts_utils.py:80-91 (JaccardIndex(num_classes=2) / Precision / Recall / F1Score /
 * FBetaScore(beta=0.5), all with threshold tau) and feeds every step in core/lit_modules/lit_model_wrappers.py:
 * 170-171 (train), 189-190 (validation), 199-200 (test): every one of them thresholds `preds >= tau` and counts
 * against the integer target.  One pass over (pred, y):
 *   counts = {TP, FP, TN, FN},  positive prediction <=> pred >= tau,  positive target <=> y != 0.
 * pred: SN_F32 / SN_F64, 16-byte aligned (float32 predictions are compared with (float)tau like torch does);
 * y: SN_F32 / SN_F64 / SN_U8 / SN_I32 / SN_I64, aligned to min(16, 16 / sizeof(pred) * sizeof(y)) bytes.
 * batch_counts (nullable): DEVICE uint64[4], OVERWRITTEN with this call's counts; total_counts (nullable): DEVICE
 * uint64[4], this call's counts are ADDED (the running state of an epoch).  At least one must be given. */
int sn_confusion_counts(const void* pred, int pred_dtype, const void* y, int y_dtype, int64_t n, double tau,
                        unsigned long long* batch_counts, unsigned long long* total_counts, void* stream);

/* ======================================================================================
 * Voxelization — replaces eda.voxelize_ply (utils/pcd_processing.py:341-372 -> pyntcloud
 * VoxelGrid.compute), Vox.hist_on_voxel / classes_on_voxel / reg_on_voxel
 * (utils/voxelization.py:164-204, 207-241, 244-300) and eda.normalize_xyz
 * (utils/pcd_processing.py:305-321).
 *
 * Points are float64 rows of `ld` doubles (x, y, z first; ld = 3 for an [N,3] array, ld = 4
 * for the TS40K .npy rows x,y,z,label of core/datasets/ts40k.py:207), labels float64 with
 * stride `label_ld` doubles (NULL = no labels).  Several clouds can be processed by one
 * launch: cloud c owns points [offsets[c], offsets[c+1]) and grid slice c (offsets: DEVICE int64 [C+1];
 * NULL with n_clouds == 1 = one cloud of n_points_total points).
 * ====================================================================================== */
/* Step 1: per-cloud bounding box.  mnmx [C,6] float64 out = (xmin,ymin,zmin,xmax,ymax,zmax).
 * The buffer is initialised by the call itself. */
int sn_vox_minmax(const double* pts, int ld, const int64_t* offsets, int n_clouds, int64_t n_points_total,
                  double* mnmx, void* stream);

/* Step 2: bin edges (np.linspace semantics: step=(hi-lo)/n rounded once, e[j]=fl(fl(j*step)+lo),
 * e[n]=hi) after the regular-bounding-box (cube) adjustment.  nx,ny,nz: voxelgrid_dims.
 * edges [C, (nx+1)+(ny+1)+(nz+1)] float64 out (x edges, then y, then z). */
int sn_vox_edges(const double* mnmx, int n_clouds, int nx, int ny, int nz, double* edges, void* stream);

/* Step 3: binning.  voxel = clip(searchsorted(edges, p, 'left') - 1, 0, n-1) per axis,
 * lin = (vz*nx + vx)*ny + vy.  count / keep_count [C,nz,nx,ny] int32 and max_label
 * [C,nz,nx,ny] (8 bytes per voxel; holds an order-preserving integer key until
 * sn_vox_finalize decodes it to float64) are INITIALISED BY THE CALL; keep_count, max_label
 * may be NULL.  keep [n_keep] float64 label values (device), lin_out [N] int32 (NULL to skip). */
int sn_vox_bin(const double* pts, int ld, const double* labels, int label_ld, const int64_t* offsets,
               int n_clouds, int64_t n_points_total, const double* edges, int nx, int ny, int nz,
               const double* keep, int n_keep, int32_t* count, int32_t* keep_count, double* max_label,
               int32_t* lin_out, void* stream);

/* Steps 1-3 in three launches instead of five (one initialisation kernel; the binning kernel derives the edges from
 * the bounding boxes itself and publishes them): same outputs as sn_vox_minmax + sn_vox_edges + sn_vox_bin, bit for
 * bit.  mnmx [C,6], edges [C, nx+ny+nz+3] float64 out; the other arguments as in sn_vox_bin. */
int sn_vox_voxelize(const double* pts, int ld, const double* labels, int label_ld, const int64_t* offsets,
                    int n_clouds, int64_t n_points_total, int nx, int ny, int nz, const double* keep, int n_keep,
                    double* mnmx, double* edges, int32_t* count, int32_t* keep_count, double* max_label,
                    int32_t* lin_out, void* stream);

/* Step 4: finalize.  density = MinMax-normalised count per y column (float64, sklearn
 * semantics), frac = keep/count (0 where empty), max_label keys -> float64 in place (0 where
 * empty), occ = (count>0), occ_keep = (keep_count>0) as 0/1 in out_dtype.  Any output may be
 * NULL.  ws: sn_vox_finalize_workspace_bytes() bytes (only needed when density != NULL). */
int64_t sn_vox_finalize_workspace_bytes(int n_clouds, int ny);
int sn_vox_finalize(const int32_t* count, const int32_t* keep_count, int n_clouds, int nx, int ny, int nz,
                    double* density, double* frac, double* max_label, void* occ, void* occ_keep,
                    int out_dtype, void* ws, void* stream);

/* ======================================================================================
 * Multi-GPU: all-reduce (sum) of the parameter-gradient payload over NVLink peer memory.
 * Replaces the gradient all-reduce Lightning's implicit DDP performs for the reference
 * (scripts/main.py:224-236) — 11..96 floats per step, pure latency — by ONE single-CTA kernel:
 * every rank stores its payload into every peer's exchange buffer as 8-byte words {value, call number} (a word carries its
 * own flag: no fence, one NVLink traversal), polls the words of the peers' slots in its own buffer and adds the payloads in
 * rank order (bit-identical
 * result on all ranks, deterministic).
 *   data            [n] float32 DEVICE, in place (n <= SN_MAX_PARAM_PTRS)
 *   peer_bufs_host  HOST array of `world` DEVICE pointers: entry w = rank w's exchange buffer of
 *                   sn_peer_allreduce_buffer_bytes(world) bytes, peer-mapped into this process
 *                   (e.g. torch symmetric memory), zero-initialised before the first call
 *   seq_counter     DEVICE uint32, local, zero-initialised: the call sequence number (the kernel
 *                   increments it, so a captured CUDA graph can be replayed as is)
 *   status          DEVICE int32 (may be NULL): set to 1 when a peer did not answer within the bound
 *   timeout_ms      bound on the wait for a peer in milliseconds of wall clock (0 = wait for ever).  Ranks of a
 *                   training job skew by seconds to minutes (checkpoints, validation, loader stalls) and NCCL simply
 *                   waits; use minutes (the Python wrapper's default: 10).  Running into the bound is FATAL: *status is
 *                   set and the kernel traps, so the process' next CUDA call fails — the payload is never replaced
 *                   by a made-up value (ABI v4; v3 wrote NaN after ~2 s and carried on).
 * Every rank must issue the same sequence of calls.
 * ====================================================================================== */
int64_t sn_peer_allreduce_buffer_bytes(int world);
int sn_peer_allreduce(float* data, int n, int rank, int world, const uint64_t* peer_bufs_host,
                      uint32_t* seq_counter, int32_t* status, int64_t timeout_ms, void* stream);
/* sn_scenenet_param_grads followed by that all-reduce of dparams, in ONE launch: the CTA that computes the parameter
 * gradients exchanges them itself (the collective of a data-parallel step starts the moment its payload exists; no second
 * launch on the critical path).  Arguments as for the two calls above; scale = 1 / world gives DDP's mean. */
int sn_scenenet_param_grads_allreduce(const sn_model_desc* desc, const float* const* param_ptrs_host,
                                      const float* K, const float* lambda_eff, const double* W, double scale, float* dparams,
                                      int rank, int world, const uint64_t* peer_bufs_host, uint32_t* seq_counter,
                                      int32_t* status, int64_t timeout_ms, void* stream);

/* ======================================================================================
 * measurement helper: FP32 FMA-pipe peak micro-benchmark (the roofline denominator that
 * MEASURED_PEAKS.json lacks).  Runs `iters` dependent-chain FFMA rounds on every
 * SM; the caller times it with CUDA events.  flops_out (host) receives the FLOPs issued. */
int sn_fp32_peak_probe(float* sink, int iters, double* flops_out_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCENENET_B200_H */
