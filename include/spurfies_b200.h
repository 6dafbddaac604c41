/*
 * spurfies_b200.h -- C ABI of libspurfies_b200.so (hand-written sm_100a kernels).
 *
 * This is the drop-in boundary for the reference's per-ray hot path.  Every entry point takes
 * raw DEVICE pointers, sizes, scalars and a cudaStream_t (passed as void*), returns 0 on success
 * or a negative spf_status, never allocates, never synchronises, and is re-entrant per stream.
 * The caller (the Python mirrors of torch_knnquery.VoxelGrid / PointVolSDF) owns all memory.
 *
 * Reference interfaces replaced (paths relative to kevinYitshak/spurfies @ 858a95f):
 *   torch_knnquery/src/knnquery.cu:570-577   pybind module `knnquery_cuda` (6 functions)
 *   torch_knnquery/torch_knnquery/knnquery.py:52-164, 168-285   VoxelGrid.set_pointset / query
 *   spurfies/model/utils.py:90-183, 221-281  query / get_keypoint_data / tv_regul glue
 *   spurfies/model/pointneus_disent.py:207-247, 300-346, 894-908  filter_points, compute_weights,
 *                                            get_sdf, get_gradients, get_color, volume_rendering
 *   spurfies/model/density.py:21-30          LaplaceDensity
 *   spurfies/model/ray_sampler.py:34-59, 377-588  UniformSampler / ErrorBoundSampler_pn
 * The ctypes binding a maintainer would add on the reference side is shown in INTEGRATION.md.
 */
#ifndef SPURFIES_B200_H
#define SPURFIES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  SPF_OK = 0,
  SPF_ERR_INVALID = -1,     /* bad argument (null pointer, K > 20, ...) */
  SPF_ERR_UNSUPPORTED = -2, /* shape outside what the kernels were written for */
  SPF_ERR_WORKSPACE = -3,   /* workspace too small */
  SPF_ERR_CUDA = -4         /* a CUDA runtime call / launch failed (see spf_last_cuda_error) */
} spf_status;

/* Voxel grid over the neural points in the REFERENCE geometry (knnquery.py:66-88): replaces the
 * five tensors coor_occ / coor_2_occ / occ_2_coor / occ_2_pnts / occ_numpnts (knnquery.py:92-97)
 * with a CSR of cell-sorted points plus the kernel_size-dilated occupancy.  All pointers device. */
typedef struct {
  float shift[3];           /* grid origin  (knnquery.py:88 d_coord_shift) */
  float vsize[3];           /* scaled voxel edge (knnquery.py:35) */
  int32_t dim[3];           /* scaled_vdim (knnquery.py:81) */
  int32_t ks[3];            /* kernel_size */
  int32_t n_points;
  int32_t n_cells;          /* dim[0]*dim[1]*dim[2] */
  const int32_t* cell_start;   /* [n_cells+1] */
  const float* sorted;         /* [n_in_grid][4]  x,y,z,point-id-bits, grouped by cell */
  const uint8_t* hit;          /* [n_cells] dilated occupancy (knnquery.cu:85-120) */
  /* Optional SEARCH grid (spf_grid_build_search; all zero / NULL when absent): same origin, cubic cells of edge
   * search_cell >= the query radius, so the radius ball is covered by 27 search cells holding ~3.4x fewer candidates
   * than the 27 reference voxels (0.075 vs radius 0.05).  Results are identical: with radius <= voxel edge the reference's
   * candidate set contains the whole ball.  The mask / slot rule always uses the reference geometry. */
  float search_cell;
  int32_t search_dim[3];
  const int32_t* search_cell_start;  /* [sx*sy*sz + 1] */
  const float* search_sorted;        /* [n_in_grid][4] grouped by search cell */
  /* 0: ray-slot queries use the thread-per-query kernel (K <= 8), 1: the cloud has very dense voxels (the caller saw
   * max points per voxel > 256 in the build statistics) -- a thread would scan thousands of candidates with divergent
   * loads, so the warp-per-query kernel (coalesced 32-candidate steps) is used for ray slots too.  Same results. */
  int32_t dense_cloud;
} spf_grid;

const char* spf_version(void);
const char* spf_last_cuda_error(void);

/* ---- a1: VoxelGrid.set_pointset (knnquery.py:52-164; knnquery.cu:22-168) ------------------- */
size_t spf_grid_workspace_bytes(int32_t n_points, int32_t n_cells);
/* Fills cell_start / sorted / hit of `g` (whose scalar fields the caller set).
 * stats[0]=#occupied voxels, [1]=max points in a voxel, [2]=#points inside the grid. */
int spf_grid_build(const spf_grid* g, const float* points /*[N,3]*/, int32_t* cell_start, float* sorted,
                   uint8_t* hit, int32_t* stats /*[4]*/, void* workspace, size_t workspace_bytes, void* stream);

/* Fills search_cell_start / search_sorted of `g` (caller set search_cell, search_dim; cell_start etc. must already be
 * built: only points inside the reference grid are inserted).  Workspace: spf_grid_workspace_bytes(n_points, search cells). */
int spf_grid_build_search(const spf_grid* g, const float* points, int32_t* search_cell_start, float* search_sorted,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- a2: mask + per-ray slots (knnquery.cu:171-221; knnquery.py:208-231) ------------------- */
int spf_mask_slots(const spf_grid* g, const float* raypos /*[R,D,3]*/, int32_t R, int32_t D, int32_t Smax,
                   int32_t* slot_sample /*[R,Smax]*/, float* sample_loc /*[R,Smax,3]*/, int32_t* n_slots /*[R]*/,
                   void* stream);
/* ---- a3: kNN per slot (knnquery.cu:224-308). pidx sorted by (d2, id), -1 padded.
 * ray_nvalid[r] = number of slots of ray r with >= 1 neighbour (knnquery.py:272-280). */
int spf_knn_slots(const spf_grid* g, const float* sample_loc, const int32_t* n_slots, int32_t R, int32_t Smax,
                  int32_t K, float radius2, int32_t* pidx /*[R,Smax,K]*/, int32_t* ray_nvalid /*[R]*/, void* stream);
/* Kernel choice for a3 (same results either way, tests compare them): 0 = automatic -- one THREAD per query for ray
 * slots and for point batches of >= 1 M points when K <= 8, radius2 > 0 and the cloud is not dense_cloud, one warp per
 * query otherwise; 1 = always one warp per query; 2 = one thread per query wherever K <= 8 and radius2 > 0. */
int spf_knn_set_algo(int32_t algo);
/* a2+a3 fused for point queries (D = 1, Smax = 1: sdf_importance / get_sdf_eval / pseudo_sdf / tv_regul). */
int spf_knn_points(const spf_grid* g, const float* q /*[Q,3]*/, int64_t Q, int32_t K, float radius2,
                   int32_t* pidx /*[Q,K]*/, void* stream);
/* The same under a device-side predicate: when `skip` is non-NULL and *skip != 0 at launch time, no point is searched
 * and every row of pidx is -1 (so the compaction that follows yields an empty list and the field kernels do nothing).
 * The eval sampler (ray_sampler.py:466-468 loops `while not_converge.sum() > 0`) uses it to make the iterations after
 * convergence free without reading the convergence flag back to the host. */
int spf_knn_points_pred(const spf_grid* g, const float* q, int64_t Q, int32_t K, float radius2, int32_t* pidx,
                        const int32_t* skip /*[1] device, may be NULL*/, void* stream);
/* mask only (knnquery.cu:171-196) */
int spf_mask_points(const spf_grid* g, const float* q, int64_t Q, int32_t* mask, void* stream);

/* ---- a4: compaction of valid slots (utils.py:90-113 masked_select glue), no host sync.
 * valid(i) := pidx[i*K] >= 0.  list[0..count) = valid i ascending; count stays on the device. */
size_t spf_compact_workspace_bytes(int64_t n);
int spf_compact_valid(const int32_t* pidx, int64_t n, int32_t K, int32_t* list /*[n]*/, int32_t* count /*[1]*/,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- a11 (first half): filter_points (pointneus_disent.py:207-239): t, delta, recomputed x --- */
int spf_ray_prep(const float* sample_loc /*[R,Smax,3]*/, const int32_t* pidx, const float* cam_loc /*[3]*/,
                 const float* ray_dirs /*[R,3]*/, int32_t R, int32_t Smax, int32_t K, float* t /*[R,Smax]*/,
                 float* delta /*[R,Smax]*/, float* x_new /*[R,Smax,3]*/, void* stream);

/* ---- a5-a8: gather + RBF weights + frozen geometry MLP (+ d sdf / d input) -------------------
 * Packed weights (see spurfies_b200/packing.py): fp32 mode: per layer Wt [Kin][256] then bias[256].
 * Rows are (slot, neighbour) pairs; slots come from `list`/`count` (spf_compact_valid).
 * Outputs are indexed by slot:  sdf[slot] (untouched where invalid), grad[slot][3] = d sdf / d x.
 * If jw != NULL: jw[(v*K+k)][32] = w_k/norm * d sdf_k / d latent (v = position in `list`), which is
 * everything the backward needs (frozen weights, detached RBF weights). */
typedef struct {
  const float* w1t; const float* b1;   /* [35][256], [256]  F_geometry.0 */
  const float* w2t; const float* b2;   /* [256][256]        F_geometry.2 */
  const float* w3t; const float* b3;   /*                   F_geometry.4 */
  const float* w4t; const float* b4;   /*                   F_geometry.6 */
  const float* v5;  float c5;          /* folded F_geometry.8 + T: sdf = v5 . h4 + c5 */
  const float* w1; const float* w2; const float* w3; const float* w4; /* [out][in] originals for the J pass */
} spf_geo_weights_f32;

int spf_sdf_fwd_f32(const spf_geo_weights_f32* W, const int32_t* list, const int32_t* count, int64_t n_max,
                    const float* x /*[n,3] by slot*/, const int32_t* pidx /*[n,K]*/, int32_t K,
                    const float* pts /*[N,3]*/, const float* feat_g /*[N,32]*/, float rbf,
                    float* sdf /*[n]*/, float* grad /*[n,3] or NULL*/, float* jw /*[n*K,32] or NULL*/, void* stream);
/* backward: feat_g_grad[p] += d_sdf[slot] * jw[row]  (warp-aggregated vector atomics) */
int spf_sdf_bwd(const int32_t* list, const int32_t* count, int64_t n_max, const int32_t* pidx, int32_t K,
                const float* jw, const float* d_sdf /*[n] by slot*/, float* feat_g_grad /*[N,32]*/, void* stream);

/* ---- a9: colour field (per pair) and radiance head (per sample), fp32 mode ------------------ */
typedef struct {
  const float* w1t; const float* b1;   /* [103][256] F_color.0 ; input = [PE6(x-p) (39) | feat_c (64)] */
  const float* w2t; const float* b2;   /* F_color.2 */
  const float* w3t; const float* b3;   /* F_color.4 */
  const float* w1; const float* w2; const float* w3; /* [out][in] originals for dgrad */
} spf_color_weights_f32;
/* hbar[slot][256] = sum_k w_k/norm * h3_k  (F_color.6 is linear, so it is applied per sample in the head).
 * Saved for backward (indexed by compact pair row v*K+k): in0 [.,104], h1, h2 [.,256], m3 [.,8] sign bits, wn [.] */
int spf_color_fwd_f32(const spf_color_weights_f32* W, const int32_t* list, const int32_t* count, int64_t n_max,
                      const float* x, const int32_t* pidx, int32_t K, const float* pts, const float* feat_c,
                      float rbf, float* hbar, float* in0, float* h1, float* h2, uint32_t* m3, float* wn,
                      void* stream);
int spf_color_bwd_f32(const spf_color_weights_f32* W, const int32_t* list, const int32_t* count, int64_t n_max,
                      const int32_t* pidx, int32_t K, const float* d_hbar /*[n,256] by slot*/,
                      const float* h1, const float* h2, const uint32_t* m3, const float* wn,
                      float* dz1, float* dz2, float* dz3 /*[n*K,256] compact rows*/, float* feat_c_grad /*[N,64]*/,
                      void* stream);

typedef struct {
  const float* w4t; const float* b4;   /* [256][256] F_color.6 (no activation) */
  const float* r1t; const float* rb1;  /* [277][256] R.0 ; input = [PE3(dir) (21) | f (256)] */
  const float* r2t; const float* rb2;  /* R.2 */
  const float* r3t; const float* rb3;  /* [256][3]   R.4 (then sigmoid) */
  const float* w4; const float* r1; const float* r2; const float* r3; /* [out][in] originals */
} spf_head_weights_f32;
/* rgb[slot][3]; saved (compact sample rows v): f [.,256], a1, a2 [.,256] */
int spf_head_fwd_f32(const spf_head_weights_f32* W, const int32_t* list, const int32_t* count, int64_t n_max,
                     const float* hbar, const float* ray_dirs /*[R,3]*/, int32_t Smax, float* rgb,
                     float* f, float* a1, float* a2, void* stream);
/* d_hbar[slot][256]; dz rows (compact v): dzf [.,256] (grad wrt F_color.6 output), dz1, dz2 [.,256], dz3 [.,4] */
int spf_head_bwd_f32(const spf_head_weights_f32* W, const int32_t* list, const int32_t* count, int64_t n_max,
                     const float* d_rgb /*[n,3] by slot*/, const float* rgb, const float* a1, const float* a2,
                     float* d_hbar, float* dzf, float* dz1, float* dz2, float* dz3, void* stream);

/* exact-mode weight gradient (the reference's nn.Linear backward, pointneus_disent.py:76-84, 100-107, as plain fp32 FFMA:
 * no library GEMM, no host synchronisation): dW [M,N] += dZ[:, :M]^T @ A[:, :N] and (db != NULL) db [M] += column sums of
 * dZ over the first count[0] * rows_per_unit rows of dZ (row stride ldz).  Row i of dZ pairs with row i of A (row stride
 * lda) or, with idx != NULL, with row idx[i] / idx_div.  dW / db are ACCUMULATED (zero them first). */
int spf_wgrad_f32(const float* dz, int32_t ldz, int32_t M, const float* act, int32_t lda, int32_t N,
                  const int32_t* idx /* optional */, int32_t idx_div, const int32_t* count, int32_t rows_per_unit,
                  int64_t n_max, float* dW, float* db, void* stream);

/* ---- a10 + a11: Laplace density + alpha compositing (density.py:21-30; pointneus_disent.py:894-908, 765-795)
 * Dense per-ray layout [R,Smax]; a slot is valid iff pidx[slot*K] >= 0.  ray_nvalid[r]==0 -> the ray is
 * "not hit" and gets the reference's fill values (pointneus_disent.py:817-854). */
int spf_composite_fwd(const float* sdf, const float* delta, const float* t, const float* rgb_s /*[R,Smax,3]*/,
                      const float* grad /*[R,Smax,3] or NULL*/, const int32_t* pidx, int32_t K,
                      const int32_t* ray_nvalid, const float* beta /*[1] device: |beta_p|+beta_min*/,
                      int32_t R, int32_t Smax, float* weights /*[R,Smax]*/, float* rgb /*[R,3]*/,
                      float* depth /*[R]*/, float* acc /*[R]*/, float* dist /*[R]*/, float* normal /*[R,3] or NULL*/,
                      void* stream);
int spf_composite_bwd(const float* sdf, const float* delta, const float* t, const float* rgb_s,
                      const int32_t* pidx, int32_t K, const int32_t* ray_nvalid, const float* beta,
                      int32_t R, int32_t Smax, const float* weights,
                      const float* d_weights /*[R,Smax] or NULL*/, const float* d_rgb /*[R,3] or NULL*/,
                      const float* d_depth /*[R] or NULL*/, const float* d_dist /*[R] or NULL*/,
                      float* d_sdf /*[R,Smax]*/, float* d_rgb_s /*[R,Smax,3]*/, float* d_beta /*[1], accumulated*/,
                      void* stream);

/* ---- a12: sampler (ray_sampler.py:34-59, 377-588) ------------------------------------------- */
/* coarse z (+ stratified jitter if t_rand != NULL) and sample positions */
int spf_sampler_coarse(const float* t_vals /*[M] linspace(0,1,M)*/, const float* t_rand /*[R,M] or NULL*/,
                       float near, float far, const float* cam_loc /*[3]*/, const float* ray_dirs /*[R,3]*/,
                       int32_t R, int32_t M, float* z /*[R,M]*/, float* points /*[R,M,3]*/, void* stream);
/* one iteration of Algorithm 1 given z,sdf [R,M]: d*, beta bisection, weights, inverse-CDF draw.
 * final != 0: pdf = weights+1e-5, N draws with u (u [R,N] or u_lin [N]); z_out [R, N+2+n_extra] sorted
 *             (extras: near, far, z[:, extra_idx]).
 * final == 0: pdf = error-bound opacity, N draws with u_lin; outputs new samples [R,N] + their points.
 * beta_io [R] in/out; beta0 from *beta_dev; flag_not_converged[0] |= (max beta > beta0). */
int spf_sampler_iter(const float* z, const float* sdf, int32_t R, int32_t M, const float* beta_dev, float eps,
                     int32_t beta_iters, float bound_coef, float add_tiny, int32_t first_iter, float* beta_io,
                     int32_t final, int32_t N, const float* u /*[R,N] or NULL*/, const float* u_lin /*[N]*/,
                     float near, float far, const int32_t* extra_idx /*[n_extra]*/, int32_t n_extra,
                     const float* cam_loc, const float* ray_dirs, float* out_z /*final: [R,N+2+n_extra], else [R,N]*/,
                     float* out_points /*[R,cols,3]*/, int32_t* flag_not_converged, void* stream);
/* The same launch under DEVICE-side control of Algorithm 1's outer loop (ray_sampler.py:466-474: "while not converged and
 * iterations left"), for the multi-iteration eval schedule without a host round trip per iteration.  state[0] = "not
 * converged" flag of the current iteration (cleared by the caller, set by the probe), state[1] = "converged earlier, final
 * draw made" (maintained by the caller: state[1] |= !state[0] after each non-final iteration).
 * pred 1 (final == 0): probe; if state[1] the new samples are parked outside every grid instead (no further SDF work).
 * pred 2 (final != 0): final draw only if !state[1] && !state[0] (this iteration converged).
 * pred 3 (final != 0): final draw unless state[1] (last allowed iteration). */
int spf_sampler_iter_pred(const float* z, const float* sdf, int32_t R, int32_t M, const float* beta_dev, float eps,
                          int32_t beta_iters, float bound_coef, float add_tiny, int32_t first_iter, float* beta_io,
                          int32_t final_, int32_t N, const float* u, const float* u_lin, float near_, float far_,
                          const int32_t* extra_idx, int32_t n_extra, const float* cam_loc, const float* ray_dirs,
                          float* out_z, float* out_points, int32_t* state /*[2]*/, int32_t pred, void* stream);
/* merge sorted z [R,M] (+sdf) with new sorted samples [R,N] (+sdf) -> [R,M+N] (ray_sampler.py:405-415, 533) */
int spf_sampler_merge(const float* z, const float* sdf, int32_t M, const float* zs, const float* sdf_s, int32_t N,
                      int32_t R, float* z_out, float* sdf_out, void* stream);

/* ---- a13: tv_regul (utils.py:221-281) on cached self-kNN lists ------------------------------ */
/* value[0] = mean_i( sum_j w_ij |f_j - f_i|_1 / sum_j w_ij ); grad (optional) = d value / d feat_g */
int spf_tv_fwd_bwd(const float* pts, const float* feat_g /*[N,32]*/, const int32_t* self_pidx /*[N,K]*/,
                   int32_t N, int32_t K, float* value /*[1]*/, float* grad /*[N,32] or NULL (accumulated)*/,
                   float grad_scale, void* stream);
/* the same over the points [first, first + count) only, value and gradient multiplied by `scale`: a data-parallel rank
 * computes its 1/W slice of the (ray-independent) regulariser with scale = W, so that the gradient average over the
 * ranks is the full term (the reference recomputes all of it on its single GPU every step, utils.py:221-281) */
int spf_tv_fwd_bwd_range(const float* pts, const float* feat_g, const int32_t* self_pidx, int32_t N, int32_t K,
                         int32_t first, int32_t count, float* value, float* grad, float scale, void* stream);

/* ---- a14: VolSDFLoss (spurfies/model/loss.py:51-100), forward and gradients in two launches ----------------------
 * terms[8] = {loss, rgb_loss, eikonal_loss, tv_loss, mask_loss, local_loss, pseudo_loss, #valid samples};
 * loss = w_rgb rgb + w_eik eik + w_tv tv + w_local local + w_pseudo pseudo + mask.  rgb_loss = mean |rgb - rgb_gt|;
 * eikonal = mean over valid[i] != 0 of (|grad_theta_i| - 1)^2 (grad_theta NULL -> 0); mask_loss = BCE(clip(sum_s
 * weights[r,s], 1e-3, 1 - 1e-3), mask_gt[r * mask_stride]) (weights NULL -> 0).  tv / local / pseudo are device scalars
 * (NULL -> 0).  d_rgb [R,3] and d_weights [R,S] (optional) receive d loss / d rgb and d loss / d weights. */
size_t spf_loss_workspace_bytes(void);
int spf_volsdf_loss(const float* rgb /*[R,3]*/, const float* rgb_gt /*[R,3]*/, const float* weights /*[R,S]*/,
                    const float* mask_gt, int32_t mask_stride, const float* grad_theta /*[n,3]*/,
                    const uint8_t* valid /*[n]*/, int64_t n, int32_t R, int32_t S, const float* tv, const float* local,
                    const float* pseudo, float w_rgb, float w_eik, float w_tv, float w_local, float w_pseudo,
                    float* terms /*[8]*/, float* d_rgb, float* d_weights, void* workspace, size_t workspace_bytes,
                    void* stream);

/* pseudo-point loss (pointneus_disent.py:765-780).  x[r] = cam + dist[r] dir[r] (spf_ray_points); after the kNN and the
 * geometry field at x: value[0] = mean |sdf_r| over rays with ray_nvalid[r] > 0 and pidx[r*K] >= 0 (1000 if rays hit but
 * none has a neighbour, 0 if nothing hit); u_sdf[r] = d value / d sdf_r; u_dist[r] (optional) = u_sdf[r] (grad_r . dir_r). */
int spf_ray_points(const float* cam_loc /*[3]*/, const float* ray_dirs /*[R,3]*/, const float* dist /*[R]*/, int32_t R,
                   float* x /*[R,3]*/, void* stream);
int spf_pseudo_loss(const float* sdf /*[R]*/, const float* grad /*[R,3]*/, const int32_t* pidx /*[R,K]*/, int32_t K,
                    const int32_t* ray_nvalid /*[R]*/, const float* ray_dirs /*[R,3]*/, int32_t R, float* value /*[1]*/,
                    float* u_sdf /*[R]*/, float* u_dist /*[R] or NULL*/, void* stream);

/* ---- a15: rays (rend_util.py:60-95, 143-156) ------------------------------------------------ */
int spf_camera_rays(const float* uv /*[R,2]*/, const float* pose /*[4,4]*/, const float* intrinsics /*[4,4]*/,
                    int32_t R, float* ray_dirs /*[R,3]*/, float* cam_loc /*[3]*/, float* depth_scale /*[R]*/,
                    void* stream);

/* ---- f1: the optimiser step after the hot path (spurfies/train.py:355-363 clip_grad_norm_ + on_after_backward
 * 
 * :548-564 NaN/Inf guard + Adam.step (train.py:168-189) + zero_grad) over flat fp32 buffers (16-byte aligned) -------------- */
size_t spf_optim_workspace_bytes(void);
/* norm_sq[0] = grad_scale^2 * sum(grad^2); deterministic (fixed grid, fixed summation order) */
int spf_grad_sumsq(const float* grad, int64_t n, float grad_scale, float* norm_sq /*[1]*/, void* workspace,
                   size_t workspace_bytes, void* stream);
/* One Adam step (the reference optimiser's defaults: betas as given, no amsgrad, no weight decay) on g * grad_scale * min(max_norm / (norm + 1e-6), 1)
 * (max_norm <= 0: no clipping).  state[0] = steps taken so far (float), state[1] = learning rate of this step; state[0]
 * is advanced on the device.  If norm_sq is not finite the update is skipped entirely (parameters, moments and step
 * count untouched), as the reference does by dropping the gradients.  zero_grad != 0 clears grad in the same pass.
 * info (optional) [2]: total norm, 1.0 if the step was skipped. */
int spf_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, const float* norm_sq,
                  float* state /*[2]*/, float grad_scale, float max_norm, double beta1, double beta2, float eps,
                  int32_t zero_grad, float* info, void* stream);
/* [S, 1/S] with S = 2^floor(log2(target / max|x|)) (clamped to 2^+-100): the per-step power-of-two scale of the
 * tensor-core mode's fp16 gradient chain, computed on the device in one launch.  scratch: 2 words, zero before the first
 * call, left zero by every call.  (No reference counterpart: the reference trains in fp32.) */
int spf_grad_scale(const float* x, int64_t n, float target, float* out /*[2]*/, uint32_t* scratch /*[2]*/, void* stream);

/* ---- f2: SDF-grid query feeding marching cubes (spurfies/utils/plots.py:188-287, 302-333; get_sdf_eval,
 * pointneus_disent.py:249-298) ------------------------------------------------------------------------------------- */
/* For the `count` grid points with linear index lo .. lo+count-1 in the reference's order (np.meshgrid(x, y, z) raveled:
 * index = (iy * nx + ix) * nz + iz): vol[t] = fill, and every point inside the dilated occupancy (knnquery.cu:171-196) is
 * appended to (idx_out = chunk-local index t, pts_out = xyz); *counter = how many (may exceed cap: then only the first
 * cap were stored).  Order of the compacted list is unspecified. */
int spf_grid_points_mask(const spf_grid* g, const float* xs, const float* ys, const float* zs, int32_t nx, int32_t ny,
                         int32_t nz, int64_t lo, int64_t count, float fill, float* vol /*[count]*/,
                         int32_t* idx_out /*[cap]*/, float* pts_out /*[cap,3]*/, int32_t* counter /*[1]*/, int32_t cap,
                         void* stream);
/* The same over a BLOCK-CYCLIC share of the grid (multi-GPU: contiguous slabs give the ranks whose slab crosses the object
 * all the work): the grid's linear index space is cut into blocks of `block` points, block b belongs to rank b % world, and
 * a rank numbers its own points consecutively (local index j -> global index ((j / block) * world + rank) * block +
 * j % block).  lo / count / vol / idx_out are in LOCAL indices.  world = 1 is spf_grid_points_mask. */
int spf_grid_points_mask_cyclic(const spf_grid* g, const float* xs, const float* ys, const float* zs, int32_t nx, int32_t ny,
                                int32_t nz, int64_t lo, int64_t count, int64_t block, int32_t world, int32_t rank,
                                const float* affine /*[12] device or NULL: point = M (x,y,z) + c, M row-major then c -- the
                                PCA-aligned grid of get_surface_by_grid(higher_res=True), plots.py:240-246*/,
                                float fill, float* vol /*[count]*/, int32_t* idx_out /*[cap]*/, float* pts_out /*[cap,3]*/,
                                int32_t* counter /*[1]*/, int32_t cap, void* stream);
/* out[idx[i]] = vals[i] */
int spf_scatter_f32(const int32_t* idx, const float* vals, int32_t n, float* out, void* stream);

/* ---- f4: feature-consistency ("local") loss on the hot path's per-ray SDF (spurfies/model/pointneus_disent.py:586-612
 * find_surface_points + :727-763; spurfies/feat_utils.py:377-451 get_local_loss with uncerts = None, :43-77 projection /
 * grid normalisation) -------------------------------------------------------------------------------------------------- */
/* Per ray r: cross[r] = first slot i with sdf[i] * sdf[i+1] < 0 and sdf[i+1] < sdf[i] (slots equal to 1000 = "no
 * neighbour" never cross), -1 if none; d_surface[r] = interpolated depth of the crossing (0 if none).  With m > 0 source
 * views: num[r] = sum over the m views of |1 - cos(feat_ref(x), feat_src_v(x))| where both projections are inside the image
 * and the term is < 0.5, x = cam_loc + d_surface * dir; g0 / g1 [R] = d num / d sdf at slots cross and cross + 1.  The loss
 * is sum(num) / (m * #crossing rays) (the reference's .mean() over [m, n]).  feat_ref [C,H,W] and feat_src [m][C,H,W] are
 * addressed as base + c * chan_stride + (y * W + x) * pix_stride (+ v * src_stride): NCHW = (H*W, 1), channels-last =
 * (1, C).  C must be 32 (SPF_ERR_UNSUPPORTED otherwise).  cam_ref [2,4,4], cam_src [m,2,4,4]: [0] world->camera,
 * [1][:3,:3] intrinsics at twice the feature resolution.  m == 0: surface search only (feature / camera pointers unused). */
int spf_local_loss_fwd(const float* sdf /*[R,Smax]*/, const float* t /*[R,Smax]*/, const float* cam_loc /*[3]*/,
                       const float* ray_dirs /*[R,3]*/, int32_t R, int32_t Smax, const float* feat_ref,
                       const float* feat_src, int64_t src_stride, int64_t chan_stride, int64_t pix_stride, int32_t channels,
                       const float* cam_ref, const float* cam_src, int32_t m, int32_t H, int32_t W,
                       const float* size /*[1] device*/, const float* center /*[3] device*/, float* num /*[R]*/,
                       int32_t* cross /*[R]*/, float* d_surface /*[R]*/, float* g0 /*[R]*/, float* g1 /*[R]*/, void* stream);
/* d_sdf [R,Smax] (every element written) = scale[0] * (g0 at slot cross, g1 at slot cross + 1, 0 elsewhere) */
int spf_local_loss_bwd(const int32_t* cross, const float* g0, const float* g1, const float* scale /*[1] device*/, int32_t R,
                       int32_t Smax, float* d_sdf, void* stream);

/* ---- f3: neural-point ingestion (spurfies/model/utils.py:6-37 construct_vox_points_closest) --------------------------- */
size_t spf_voxelize_workspace_bytes(int32_t cells_per_axis);
/* Voxel-downsample: voxel of a point = floor((p - space_min) / vox_size) per axis (fp32, as the reference computes it);
 * for every occupied voxel, in sorted (x, y, z) voxel order (the order of unique(dim=0)): min_idx = index of the input
 * point closest to the voxel centroid (ties: smallest index), centroid (optional) = mean of its points, grid_idx
 * (optional) = the voxel's integer coordinates.  n_out[0] = number of occupied voxels (may exceed cap: then only the
 * first cap rows were written), n_out[1] = points that fell outside the cells_per_axis^3 table (must be 0). */
int spf_voxelize_closest(const float* points /*[n,3]*/, int32_t n, float min_x, float min_y, float min_z, float vox_size,
                         int32_t cells_per_axis, int64_t* min_idx /*[cap]*/, float* centroid /*[cap,3] or NULL*/,
                         int32_t* grid_idx /*[cap,3] or NULL*/, int32_t cap, int32_t* n_out /*[2]*/, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ---- bf16 tensor-core mode (tcgen05.mma + TMEM) ------------------------------------------------
 * Packed weight images: see spurfies_b200/packing.py (k-block major, 128B-swizzled, bf16). */
typedef struct {
  const uint8_t* w1p;   /* pack([W1[:, :32] | W1[:, 32:35] | W1[:, 32:35]] -> K = 38 padded to 64): x - p enters as bf16 hi + lo */
  const uint8_t* w2p; const uint8_t* w3p; const uint8_t* w4p;     /* pack(F_geometry.{2,4,6}.weight) */
  const uint8_t* w4tp; const uint8_t* w3tp; const uint8_t* w2tp;  /* pack(weight^T) for the d sdf / d input chain */
  const uint8_t* w1tp;  /* pack(W1^T padded to [48][256]) */
  const float* b1; const float* b2; const float* b3; const float* b4; const float* v5; float c5;
} spf_geo_weights_tc;
/* same contract as spf_sdf_fwd_f32 (bf16 inputs / weights, fp32 accumulation, 2e-2 tolerance) */
int spf_sdf_fwd_tc(const spf_geo_weights_tc* W, const int32_t* list, const int32_t* count, int64_t n_max,
                   const float* x, const int32_t* pidx, int32_t K, const float* pts, const float* feat_g, float rbf,
                   float* sdf, float* grad, float* jw, void* stream);
typedef struct {
  const uint8_t* w1p;   /* pack(F_color.0.weight with input columns permuted to [c (64) | PE6 (39)], K = 103 -> 128) */
  const uint8_t* w2p; const uint8_t* w3p;                     /* pack(F_color.{2,4}.weight) */
  const uint8_t* w3tp; const uint8_t* w2tp;                   /* pack(weight^T) for dgrad */
  const uint8_t* w1ftp;                                       /* pack(F_color.0.weight[:, 39:103]^T) -> [64][256] */
  const float* b1; const float* b2; const float* b3;
} spf_color_weights_tc;
/* as spf_color_fwd_f32; saved tensors are bf16 in the TILE layout (see spf_wgrad_tc): in0 [rows,128] (permuted
 * columns, 112 used), h1, h2 [rows,256]; m3 is
 * [rows,24]: the LeakyReLU sign words of z1, z2, z3 (8 words per layer), consumed by spf_color_bwd_tc */
int spf_color_fwd_tc(const spf_color_weights_tc* W, const int32_t* list, const int32_t* count, int64_t n_max,
                     const float* x, const int32_t* pidx, int32_t K, const float* pts, const float* feat_c, float rbf,
                     float* hbar, void* in0, void* h1, void* h2, uint32_t* m3, float* wn,
                     void* hb /* optional out: bf16 copy of hbar by COMPACT slot (position in list), tile layout */,
                     void* stream);
/* as spf_color_bwd_f32; dz1..3 are fp16 [rows,256] in the TILE layout, scaled by S = gscale[0]; h1 / h2 are not read (the
 * sign words in m3 are); feat_c_grad receives the unscaled gradient */
int spf_color_bwd_tc(const spf_color_weights_tc* W, const int32_t* list, const int32_t* count, int64_t n_max,
                     const int32_t* pidx, int32_t K, const float* d_hbar, const void* h1, const void* h2,
                     const uint32_t* m3, const float* wn, void* dz1, void* dz2, void* dz3, float* feat_c_grad,
                     const void* d_hb /* optional: S * d_hbar as fp16 by COMPACT slot, tile layout (then d_hbar may be NULL) */,
                     const float* gscale /* device [S, 1/S], see spf_head_bwd_tc */, void* stream);
typedef struct {
  const uint8_t* w4p;    /* pack(F_color.6.weight) */
  const uint8_t* r1fp;   /* pack(R.0.weight[:, 21:277])  (the PE3(dir) columns enter through zpe) */
  const uint8_t* r2p;    /* pack(R.2.weight) */
  const uint8_t* r3p;    /* pack(R.4.weight padded to [32][256]) */
  const uint8_t* r3tp;   /* pack(R.4.weight^T padded to [256][16 -> 64]) */
  const uint8_t* r2tp; const uint8_t* r1ftp; const uint8_t* w4tp;  /* pack(weight^T) for dgrad */
  const float* b4; const float* rb2; const float* rb3;
} spf_head_weights_tc;
/* zpe [R,256] = PE3(dir) @ R.0.weight[:, :21]^T + R.0.bias (fp32, per ray).  Saved bf16 [rows,256] by compact sample
 * row, in the TILE layout (see spf_wgrad_tc): hb (= hbar), f, a1, a2; rows = ceil(count / 128) * 128.
 * hbar may be NULL: hb is then an INPUT (the compact copy written by spf_color_fwd_tc) and is bulk-copied as the A operand;
 * the rows of its last tile beyond count are zeroed in place. */
int spf_head_fwd_tc(const spf_head_weights_tc* W, const int32_t* list, const int32_t* count, int64_t n_max,
                    const float* hbar, const float* zpe, const float* ray_dirs, int32_t Smax, float* rgb, void* hb, void* f,
                    void* a1, void* a2, void* pe /*[rows,64] bf16 tile layout, 32 columns used: PE3(dir) per sample*/, void* stream);
/* a1, a2 in and dzf, dz1, dz2 out: bf16 [rows,256] in the TILE layout; dz3 is bf16 [rows,64] in the tile layout (3 used);
 * drb3 [3] (fp32) is accumulated in-kernel */
int spf_head_bwd_tc(const spf_head_weights_tc* W, const int32_t* list, const int32_t* count, int64_t n_max,
                    const float* d_rgb, const float* rgb, const void* a1, const void* a2, float* d_hbar, void* dzf,
                    void* dz1, void* dz2, void* dz3, float* drb3,
                    void* d_hb /* optional out: S * d_hbar as fp16 by compact sample row, tile layout, INSTEAD of d_hbar */,
                    const float* gscale /* device [S, 1/S].  The gradient chain (dZ tiles, d_hb) is fp16 scaled by a power of
                                           two S picked per step from the largest upstream gradient (16 / max|d_rgb| rounded
                                           down to a power of two) so that it sits in fp16's normal range; fp32 outputs
                                           (d_hbar, drb3) are unscaled again. */,
                    void* stream);
/* 16-bit 128B-swizzled k-block-major image of W [N][K] (row stride ld, fp32) or of its transpose (then W is [K][N]);
 * out holds ceil(K/64) * n_pad * 128 bytes.  `transpose` is a flag word: bit 0 = transpose, SPF_PACK_F16 = write fp16
 * (the FORWARD weight images: forward operands are fp16, gradient-chain operands bf16) instead of bf16. */
#define SPF_PACK_F16 2
int spf_pack_sw128(const float* W, int32_t ld, int32_t N, int32_t K, int32_t transpose, int32_t n_pad, void* out,
                   void* stream);
/* the same for up to SPF_PACK_MAX_JOBS images in one launch (the trainable weights are re-packed every step) */
#define SPF_PACK_MAX_JOBS 16
typedef struct {
  const float* W; void* out;
  int32_t ld, N, K, transpose, n_pad, reserved;
} spf_pack_job;
int spf_pack_sw128_batch(const spf_pack_job* jobs /* HOST array */, int32_t n_jobs, void* stream);
/* zpe [R,256] = PE3(ray_dirs) @ W[:, :21]^T + bias with W = R.0.weight (row stride ld >= 21), bias = R.0.bias [256] */
int spf_head_zpe(const float* ray_dirs /*[R,3]*/, const float* W, int32_t ld, const float* bias, int32_t R, float* zpe,
                 void* stream);
/* weight gradient of one linear layer: dW[256][N] += dZ^T @ A, db[256] += colsum(dZ) (both accumulated in fp32), over
 * the first ceil(count * rows_per_unit / 128) * 128 rows of dZ [.,256] / A [.,lda] (bf16): exactly the rows the dgrad
 * kernels write.  N multiple of 16, <= 256.  layout bit 0 / bit 1: dZ / A is stored in the tile layout written by the
 * colour-field kernels (128-row tiles, k-blocks of 64 columns, 128-byte rows, 16-byte chunk c of row r at position
 * c ^ (r & 7); A then has ceil(lda / 64) k-blocks per tile) instead of row-major.  No host synchronisation. */
int spf_wgrad_tc(const void* dz, const void* act, int32_t lda, int32_t N, const int32_t* count, int32_t rows_per_unit,
                 int64_t n_max, int32_t layout, float* dW, float* db, void* stream);
/* up to SPF_WGRAD_MAX_JOBS such products over the same rows in one launch; every operand in the TILE layout (dZ
 * [rows,256], A [rows,lda] with lda a multiple of 64); db may be NULL.  The CTAs are split among the jobs by bytes per row. */
#define SPF_WGRAD_MAX_JOBS 8
typedef struct {
  const void* dz; const void* act; float* dW; float* db;
  int32_t lda, N;
  int32_t fmt;      /* bit 0: dz is bf16 (else fp16); bit 1: act is bf16 (else fp16).  An fp16 operand is converted to
                     * bf16 in shared memory before the MMAs (tcgen05 kind::f16 needs A and B in the same format). */
  int32_t reserved;
} spf_wgrad_job;
int spf_wgrad_tc_multi(const spf_wgrad_job* jobs /* HOST array */, int32_t n_jobs, const int32_t* count,
                       int32_t rows_per_unit, int64_t n_max,
                       const float* gscale /* optional device [S, 1/S]: the gradient tiles carry the step's power-of-two
                                              scale S (see spf_head_bwd_tc); dW / db are multiplied by 1/S */,
                       void* stream);
/* building-block self test: out[128][N] = A[128][K] (bf16 row-major) @ W^T with W given as a packed image */
int spf_tc_gemm_test(const void* A, const void* Wpacked, int32_t N, int32_t K, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPURFIES_B200_H */
