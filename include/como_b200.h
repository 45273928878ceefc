/*
 * como_b200 -- C ABI of the B200-native (sm_100a) photometric Gauss-Newton hot path of COMO.
 *
 * Drop-in boundary: every entry point takes plain device pointers + sizes
 * + a cudaStream_t (passed as void*), returns 0 on success / a negative COMO_B200_E* code on an
 * argument or launch error, never throws, never allocates (outputs and workspaces are caller
 * owned -- torch owns the memory in the Python veneer), and never falls back to a CPU path.
 *
 * Each function cites the reference interface it replaces (paths relative to the reference repo).
 */
#ifndef COMO_B200_H
#define COMO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COMO_B200_OK 0
#define COMO_B200_EINVAL (-1)    /* bad argument (null pointer, size out of range)            */
#define COMO_B200_ELAUNCH (-2)   /* CUDA launch / runtime error (cudaGetLastError != success) */
#define COMO_B200_EWORKSPACE (-3)/* workspace too small                                       */
#define COMO_B200_EUNSUPPORTED (-4)

#define COMO_B200_MAX_LEVELS 8

/* ABI version, bumped on any signature change. */
int como_b200_abi_version(void);
/* Human readable description of the last error on this host thread ("" if none). */
const char* como_b200_last_error(void);
/* SE(3) exponential of a twist [tau (3), phi (3)] (translation first, lietorch's order; the reference swaps its
 * [omega, v] into it at como/geometry/lie_algebra.py:45-49) -> row-major 4x4, evaluated on the HOST by the same
 * function the tracking / SfM / BA update kernels call.  Test hook: no GPU needed. */
void como_b200_se3_exp(const double* tau_phi, double* T16);

/* ------------------------------------------------------------------------------------------
 * Tracking: frame-to-keyframe inverse-compositional photometric GN (fp32, gray or rgb).
 * Replaces como/odom/frontend/photo_tracking.py:10-42 (photo_tracking_pyr), :147-185
 * (photo_level_tracking), :117-143 (tracking_iter), :77-114 (robustify/solve/update) and the
 * torch ops under them (geometry/camera.py:57-68, frontend/photo_utils.py:9-31, lietorch SE3.exp).
 * One persistent cooperative launch runs the whole coarse-to-fine loop incl. termination tests.
 * ------------------------------------------------------------------------------------------ */
/* A level as the reference passes it (vals, P, J, mask) plus `pack`: the same operands re-laid out ONCE per keyframe
 * by como_b200_track_pack into 512-pixel tiles [P | I_ref | J cols 0..3 | J cols 4..5 | residual scratch] (masked and
 * padding pixels carry NaN points), so that each GN pass streams one contiguous run per tile with the TMA unit and the
 * residual travels inside the Jacobian tile instead of through a second array.  With c > 1 image channels
 * (tracking.color: rgb, como/odom/Tracking.py:57-60) the tiles of channel 0 come first, then those of channel 1, ...:
 * every (pixel, channel) pair is one entry with its own I_ref and Jacobian row and a copy of the point.
 * como_b200_track_pyr reads only pack, img, n, w, h, c, K; it WRITES the residual slots of `pack` (one launch at a
 * time per pack). */
typedef struct {
  const float* vals;   /* (n,c)   reference intensities I_i            [photo_tracking.py:24 vals_i]  */
  const float* P;      /* (n,3)   reference points in the KF frame     [Pi]                           */
  const float* J;      /* (n,c,8) dI/d[xi,a,b] as precalc_jacobians lays it out; cols 6 (rewritten every iteration by
                        *          the reference, photo_tracking.py:125) and 7 (ones) are not read     [dI_dT] */
  const uint8_t* mask; /* (n) 0/1 or NULL: which points take part      [masks]                        */
  const float* img;    /* (c,h,w) target image of this level           [img_j]                        */
  void* pack;          /* como_b200_track_pack_bytes(n, c) bytes, 128-byte aligned, filled by como_b200_track_pack */
  int32_t n, w, h;
  int32_t c;           /* image channels: 1 (gray; 0 is read as 1) or 3 (rgb)                         */
  float K[9];          /* row-major 3x3 intrinsics of this level       [intrinsics]                   */
} como_b200_track_level_t;

typedef struct {
  int32_t max_iter;    /* term_criteria["max_iter"]   (config/como.yml:12-17) */
  float delta_norm;    /* term_criteria["delta_norm"] */
  float rel_tol;       /* term_criteria["rel_tol"]    */
  float grad_norm;     /* term_criteria["grad_norm"]  */
} como_b200_track_term_t;

/* per-iteration record written to `stats` (32 floats per iteration, iteration-major) */
#define COMO_B200_TRACK_STAT_STRIDE 32
/* [0]=level [1]=mean_sq_err [2]=grad_norm [3]=delta_norm [4]=sigma_r [5]=num_valid [6]=done [7]=reserved
 * [8..23]=Tji (row-major 4x4) and [24..25]=[a,b] this iteration was linearised at (tracking_iter's inputs,
 * photo_tracking.py:117) -- lets a checker replay every iteration from identical inputs; [26..31]=reserved */

size_t como_b200_track_workspace_bytes(int32_t max_n, int32_t num_problems);
/* Keyframe-side re-layout (once per keyframe and level; the reference's update_kf_reference, Tracking.py:243-379,
 * is where vals / P / J / mask are produced).  Reads level->vals, P, J (cols 0..5), mask, n; writes level->pack. */
size_t como_b200_track_pack_bytes(int32_t n, int32_t c);
int como_b200_track_pack(const como_b200_track_level_t* level, void* stream);

/* levels: HOST array [num_problems][num_levels] (coarsest first, as the reference stores pyramids).
 * T: DEVICE (num_problems,4,4) in/out Tji; aff: DEVICE (num_problems,2) in/out [a,b].
 * stats: DEVICE (num_problems, num_levels*max_iter, 8) or NULL; num_iters: DEVICE (num_problems) int32 or NULL. */
int como_b200_track_pyr(const como_b200_track_level_t* levels, int32_t num_levels, int32_t num_problems,
                        const como_b200_track_term_t* term, float* T, float* aff, float* stats,
                        int32_t* num_iters, void* workspace, size_t workspace_bytes, void* stream);
/* Test hook: largest median-bin population that is finished through the candidate list (default 2048);
 * 0 forces the histogram-narrowing path for every iteration.  Both paths return the same exact order statistic. */
void como_b200_track_debug_candidate_cap(int32_t cap);

/* Replaces precalc_jacobians (como/odom/frontend/photo_tracking.py:46-74).
 * grads (n,c,2) [gx,gy]; P (n,3); vals (n,c); K 9 floats (host); c channels; out J (n,c,8). */
int como_b200_precalc_jacobians(const float* grads, const float* P, const float* vals, const float* K,
                                int64_t n, int32_t c, float* J, void* stream);

/* ------------------------------------------------------------------------------------------
 * Exact segmented lower median (torch.median semantics) of non-negative values; NaN entries are
 * skipped.  out[s] = scale * median(values[seg_offsets[s] : seg_offsets[s+1]]), count[s] = #valid.
 * Replaces torch.median at como/odom/backend/photo.py:124-128, como/odom/Mapping.py:757-758,
 * como/odom/Tracking.py:342-345.  Radix select, 11-bit digits (6 passes f64 / 3 passes f32).
 * ------------------------------------------------------------------------------------------ */
size_t como_b200_median_workspace_bytes(int32_t num_segments, int32_t elem_bytes);
int como_b200_median_f64(const double* values, const int64_t* seg_offsets, int32_t num_segments,
                         int64_t max_segment_len, double scale, double* out, int64_t* count, void* workspace,
                         size_t workspace_bytes, void* stream);
/* building blocks of the same selection for a median over values spread across ranks: zero hist
 * (como_b200_median_num_passes x segments x 2048 uint32), then pass(digit) [+ all-reduce of hist[digit]] ..., finish. */
int32_t como_b200_median_num_passes(int32_t elem_bytes);
int como_b200_median_pass_f64(const double* values, const int64_t* seg_offsets, int32_t num_segments,
                              int64_t max_segment_len, int32_t digit, void* hist, void* stream);
int como_b200_median_finish_f64(int32_t num_segments, const void* hist, double scale, double* out, int64_t* count,
                                void* stream);
/* Three-exchange variant for values spread over ranks: digits 0 and 1 by pass + all-reduce as above, then every rank
 * compacts its candidates of the chosen 22-bit bucket into a pack of como_b200_median_pack_words() 8-byte words per
 * segment (word 0 = count), the packs are all-gathered into (world, num_segments, words) and every rank finishes.
 * overflow[s] = 1: some rank's candidates did not fit -- fall back to the six-pass scheme for this call. */
int32_t como_b200_median_pack_words(void);
int como_b200_median_dist_compact_f64(const double* values, const int64_t* seg_offsets, int32_t num_segments,
                                      int64_t max_segment_len, const void* hist, void* pack, void* stream);
int como_b200_median_dist_finish_f64(const void* packs, int32_t world, int32_t num_segments, const void* hist,
                                     double scale, double* out, int32_t* overflow, void* stream);
int como_b200_median_f32(const float* values, const int64_t* seg_offsets, int32_t num_segments,
                         int64_t max_segment_len, float scale, float* out, int64_t* count, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Window bundle adjustment (fp64): one Gauss-Newton iteration of como/odom/Mapping.py:760-968.
 * All arrays are DEVICE pointers unless stated otherwise.  Keyframes K, one-way frames R, landmarks L,
 * anchors per keyframe M (multiple of 4, <= 64), sampled pixels per keyframe N = (H/win)*(W/win).
 * Frame index f in [0,K+R): keyframes first.  Unknown layout of H (dim x dim) / g (dim), as the
 * reference (Mapping.py:701-747): 8 per keyframe, 8 per one-way frame, then 3 per landmark.
 * ------------------------------------------------------------------------------------------ */

/* subselect_pixels (como/odom/backend/sparse_map.py:116-142): per win x win cell the first strict
 * maximum of |grad I|; coords (K,N,2) int32 [row,col] (bit-exact), vals (K,N) = I at those pixels. */
int como_b200_subselect_pixels(const double* img_and_grads /*(K,3,H,W)*/, int32_t K, int32_t H, int32_t W,
                               int32_t win, int32_t* coords, double* vals, void* stream);

/* prep_geometry_scaffold / project_landmarks (Mapping.py:603-659, sparse_map.py:18-60).
 * lm_ids (K,M) landmark id per anchor slot; fo_slots (L): flat (k*M+m) slot of the j-th first
 * observation in row-major order; scaffold out (K,M,16): logz,u,pm(2),Pc(3),zmask,dlogz/dT(6),pad(2);
 * dz_dP out (K,3).  Also re-initialises landmarks behind their camera in place (P_m). intr4=[fx,fy,cx,cy] HOST. */
int como_b200_ba_scaffold(const double* kf_poses, double* P_m, const int32_t* lm_ids, const int32_t* fo_slots,
                          const double* pm_first_obs, const double* median_depths, int32_t K, int32_t L, int32_t M,
                          const double* intr4, double* scaffold, double* dz_dP, void* stream);

/* store_vars (Mapping.py:749-758): depth (K,HW) = exp(Knm_Kmminv (K,HW,M) . logz_m); logz_m read from scaffold. */
int como_b200_predictor_apply(const double* Knm, const double* scaffold, int32_t K, int64_t HW, int32_t M,
                              double* depth, void* stream);
/* Grid size of the streaming kernel behind predictor_apply: 0 (default) = one CTA on every SM; other values are for
 * experiments with kernels on other streams (store_vars runs next to the normal-equation build in Mapping.iterate). */
void como_b200_predictor_stream_ctas(int32_t ctas);
/* column sums of one keyframe's predictor (mean_log_depth_cost, gp_priors.py:99-106); M must divide 256. */
int como_b200_predictor_colsum(const double* Knm, int64_t HW, int32_t M, double* colsum, void* stream);

/* create_photo_system (como/odom/backend/photo.py:236-353) without the materialised (b,N,3,M,1)
 * Jacobian, in three steps so that the robust scale can be made global across GPUs:
 *   1. como_b200_ba_photo_residual : frame table + residual of every (pair, pixel)  -> rbuf (P,N) [NaN = invalid],
 *                                    pairbuf (P,N,4), refbuf (K,N,8)
 *   2. exact median of |rbuf| per pair batch: como_b200_median_pass_f64 x 6 (all-reduce hist[digit] between
 *      passes when the pairs are sharded) + como_b200_median_finish_f64 with scale 1.4826  -> sigma_batch
 *   3. como_b200_ba_photo_accum    : Huber weights, J^T J / J^T r blocks, scatter into H, g (atomics)
 * Pair p: reference keyframe pair_ref[p], target frame pair_tgt[p], pair_slot[p] = position in the reference
 * keyframe's target list, pair_batch[p] = its batch (photo.py:262-300); ref_ptr (K+1) / ref_pairs (P): CSR of
 * pairs by reference keyframe.  units (num_units, como_b200_ba_unit_ints() int32 each):
 * [ref, pix_begin, pix_end, tgt_begin, tgt_end, primary, 0, 0], ordered (ref, target group, slice);
 * unit_base[ref] = first unit, unit_slices[ref] = slices per group (0: no pair on this rank).
 * photo_err += sum w (r/sigma)^2. */
size_t como_b200_ba_frames_bytes(int32_t num_frames);
size_t como_b200_ba_partial_doubles(int32_t num_units);
int32_t como_b200_ba_unit_ints(void);
int32_t como_b200_ba_target_group(void);
int como_b200_ba_photo_residual(const double* kf_poses, const double* kf_aff, const double* rec_poses,
                                const double* rec_aff, const double* kf_img, const double* rec_img, const double* Knm,
                                const int32_t* coords, const double* vals_n, const double* scaffold,
                                const int32_t* pair_tgt, const int32_t* ref_ptr, const int32_t* ref_pairs, int32_t K,
                                int32_t R, int32_t M, int32_t N, int32_t Himg, int32_t Wimg, int32_t P,
                                const double* intr4, void* frames_ws, double* refbuf, double* rbuf, double* pairbuf,
                                void* stream);
int como_b200_ba_photo_accum(const double* Knm, const int32_t* coords, const double* scaffold, const double* dz_dP,
                             const int32_t* lm_ids, const int32_t* pair_ref, const int32_t* pair_tgt,
                             const int32_t* pair_slot, const int32_t* pair_batch, const int32_t* ref_ptr,
                             const int32_t* ref_pairs, const int32_t* units, const int32_t* unit_base,
                             const int32_t* unit_slices, int32_t num_units, int32_t K, int32_t R, int32_t L, int32_t M,
                             int32_t N, int32_t Himg, int32_t Wimg, int32_t P, const double* intr4, int32_t dim,
                             const double* sigma_batch, const void* frames_ws, const double* refbuf, const double* rbuf,
                             const double* pairbuf, double* sigma_pair_ws, double* partial, double* H, double* g,
                             double* photo_err, void* stream);

/* Prior factors of Mapping.iterate (Mapping.py:809-917; como/odom/factors/).  LtL (K,M,M) = L_mm^-T L_mm^-1.
 * sigmas4 HOST = [pixel_sigma_first, pose_prior, scale_prior, mean_depth_prior]; err8 += [_, gp, logdepth, pixel,
 * pose, affine, scale, fixed]. */
int como_b200_ba_priors(const double* scaffold, const double* dz_dP, const double* LtL, const double* median_depths,
                        const uint8_t* obs_ref_mask, const double* pm_first_obs, const int32_t* lm_ids,
                        const double* kf_poses, const double* pose_anchor, const double* kf_aff,
                        const double* aff_anchor, const double* colmean, const double* P_m,
                        const double* P_m_anchors, const int32_t* fix_ids, int32_t nfix, int32_t window_full,
                        const double* sigmas4, double scale_anchor, int32_t K, int32_t R, int32_t L, int32_t M,
                        const double* intr4, int32_t dim, double* H, double* g, double* err8, void* stream);

/* update_vars (como/odom/backend/linear_system.py:115-152): T <- T Exp(delta), affine += delta, P_m += delta. */
int como_b200_ba_update(const double* delta, int32_t K, int32_t R, int32_t L, double* kf_poses, double* kf_aff,
                        double* rec_poses, double* rec_aff, double* P_m, void* stream);

/* ------------------------------------------------------------------------------------------
 * DepthCov Gaussian process.
 * ------------------------------------------------------------------------------------------ */

/* como_backends.cross_covariance (como/backend/src/cov.cpp:5-32, cov_cpu.cpp:17-64, cov_gpu.cu:17-84).
 * x1 (B,n1,2), E1 (B,n1,2,2), x2 (B,n2,2), E2 (B,n2,2,2) contiguous, elem_bytes 4 or 8 -> out (B,n1,n2). */
int como_b200_cross_covariance(const void* x1, const void* E1, const void* x2, const void* E2, double scale, int32_t B,
                               int32_t n1, int32_t n2, int32_t elem_bytes, void* out, void* stream);

/* como_backends.get_new_chol_obs_info (cov.cpp:34-65, cov_cpu.cpp:66-85, cov_gpu.cu:132-215), float32:
 * in place: L (B,n,n) gets row N, obs_info (B,n,d) gets row N, var (B,d) -= row^2. k_ni (B,N,1), k_id (B,1,d). */
int como_b200_chol_append(float* L, float* obs_info, float* var, const float* k_ni, const float* k_id, float k_ii,
                          int32_t B, int32_t n, int32_t d, int32_t N, void* stream);

/* Device-resident greedy conditional-entropy loop (como/depth_cov/core/samplers.py:211-302): selects anchors
 * m..n-1.  dom_xy (B,d,2) normalised coords, dom_E (B,d,4); sel_* / L / obs_info / var / dist_ok pre-filled
 * for the first m anchors (precalc_entropy_vars, samplers.py:115-208).  sel_idx (B,n) int64 (bit-exact
 * domain indices), count_out (B) int32 = number of anchors after the loop (early termination). */
size_t como_b200_sampler_workspace_bytes(int32_t B, int32_t d);
int como_b200_sampler_greedy(const float* dom_xy, const float* dom_E, int32_t B, int32_t d, int32_t n, int32_t m,
                             float* sel_xy, float* sel_E, int64_t* sel_idx, float* L, float* obs_info, float* var,
                             uint8_t* dist_ok, float signal_var, float fixed_var, int32_t has_fixed, float dist_thresh,
                             float max_stdev_thresh, int32_t terminate_early, int32_t* count_out, void* workspace,
                             size_t workspace_bytes, void* stream);

/* Mapping.prep_predictor (como/odom/Mapping.py:430-468) with the Python covariance formula
 * (como/depth_cov/core/covariance.py:10-39, kernels.py:22-89), fp64.
 * kmat_kmm: E_m (B,M,4) by bilinear lookup of cov_img (B,4,H,W) at coords_m (B,M,2) [row,col]; K_mm (B,M,M) + jitter I.
 * kmat_predictor: Knm_Kmminv (B,H*W,M) = K_nm(all pixels, anchors) * Kmm_inv, K_nm never materialised. */
int como_b200_kmat_kmm(const double* cov_img, int32_t B, int32_t H, int32_t W, const double* coords_m, int32_t M,
                       double scale, double jitter, double* E_m, double* K_mm, void* stream);
int como_b200_kmat_predictor(const double* cov_img, int32_t B, int32_t H, int32_t W, const double* coords_m,
                             const double* E_m, const double* Kmm_inv, int32_t M, double scale, double* Knm_Kmminv,
                             void* stream);

/* lin_sys.solve_system (como/odom/backend/linear_system.py:101-112): x = H^-1 g for the SPD normal equations,
 * H (n,n) row-major fp64 (only the lower triangle is read, H is not modified), g (n), x (n).  Tiled dataflow
 * Cholesky + fused forward substitution + backward substitution (csrc/chol.cu).  Non-PD input gives NaN. */
size_t como_b200_chol_solve_workspace_bytes(int32_t n);
/* Grid size of the factorisation (0 = one CTA per SM, the default). */
void como_b200_chol_ctas(int32_t ctas);
/* Schedule of the factorisation: 1 (default) = one CTA walks the critical path (diagonal tile + the tile below it)
 * with operands kept in shared memory while the others pre-accumulate; 0 = every tile an independent dataflow task.
 * Same arithmetic per tile; tuning / test hook. */
void como_b200_chol_schedule(int32_t mode);
int como_b200_chol_solve(const double* H, const double* g, int32_t n, double* x, void* workspace,
                         size_t workspace_bytes, void* stream);

/* Keyframe-creation path: calc_kernel_matrices + get_predictor at arbitrary test points
 * (como/depth_cov/core/distill_depth.py:8-48) and the normal equations of distill_depth /
 * distill_conditional_depth_with_scale_prior (distill_depth.py:51-82, 126-153; lstsq_chol utils/lin_alg.py:82-87).
 * kmat_rows: rows (B,n,M) = K_nm K_mm^-1 at coords_n (B,n,2) [row,col] (fractional, border-clamped bilinear
 *   covariance lookup); mask_n (B,n) optional (0 -> zero row); var_n (B,n) optional = K_nn - K_nm K_mm^-1 K_mn for
 *   valid rows; var_min (B) optional = min over valid rows (the torch.min of get_predictor).
 * weighted_gram (B=1): G (M,M) = sum_n w_n k_n k_n^T, h (M) = sum_n w_n y_n k_n, w_n = wscale / (var_n + var_add)
 *   (var != NULL) or wscale; stats (2, optional) = [sum w, sum w y^2]; masked rows skipped.
 * rows_residual (B=1): res_n = k_n . x - y_n; stats3 = [count, sum res, sum (res - mean)^2] over valid rows. */
int como_b200_kmat_rows(const double* cov_img, int32_t B, int32_t H, int32_t W, const double* coords_m,
                        const double* E_m, const double* Kmm_inv, int32_t M, double scale, const double* coords_n,
                        const uint8_t* mask_n, int64_t n, double* rows, double* var_n, double* var_min, void* stream);
int como_b200_weighted_gram(const double* rows, const double* y, const double* var, const uint8_t* mask, int64_t n,
                            int32_t M, double var_add, double wscale, double* G, double* h, double* stats,
                            void* stream);
int como_b200_rows_residual(const double* rows, const double* x, const double* y, const uint8_t* mask, int64_t n,
                            int32_t M, double* res, double* stats3, void* stream);
/* Two-frame SfM bootstrap (como/odom/frontend/two_frame_sfm.py:180-392), one Gauss-Newton iteration of
 * the (6 + M) system [T_ji | sparse log depths]:
 * sfm_linearize: Knm (N,M) predictor rows of the reference pixels, d (M) sparse log depths, coords (N,2) int64
 *   [row,col], vals (N) reference intensities, img_and_grads (3,H,W) of the other frame, T_ji12 / intr4 HOST ->
 *   rec (N,8) = [r (NaN if invalid), dI/dT (6), beta = dI/dP_i . P_i], absr (N) = rec[:,0] (input of the median),
 *   proj (N,3) = [u, v, Z_j], stats2 = [sum_n logz_n, number of valid pixels].
 * sfm_accumulate: sigma (DEVICE scalar, 1.4826 median |r|) -> G (M,M) = sum s beta^2 k k^T, St7 (7,M): rows 0..5 =
 *   sum s beta dI/dT k^T, row 6 = sum s beta r k^T, small28 = [upper triangle of sum s dI/dT^T dI/dT (21),
 *   sum s dI/dT r (6), sum w (r/sigma)^2], s = huber(r/sigma)/sigma^2. */
int como_b200_sfm_linearize(const double* Knm, const double* d, const int64_t* coords, const double* vals,
                            const double* img_and_grads, int32_t H, int32_t W, int64_t N, int32_t M,
                            const double* T_ji12, const double* intr4, double* rec, double* absr, double* proj,
                            double* stats2, void* stream);
int como_b200_sfm_accumulate(const double* Knm, const double* rec, const double* sigma, int64_t N, int32_t M, double* G,
                             double* St7, double* small28, void* stream);

/* track_and_init helpers (como/odom/frontend/corr.py:37-43, 17-29, 80-96, 116-152).
 * reproject_dense: z_img (H,W) of the last keyframe, T_ji12 HOST row-major 3x4, intr4 HOST [fx,fy,cx,cy] ->
 *   coords_j (H*W,2) [row,col], logz_j, z_j (H*W), mask (H*W) = inside [1, dim-1) and z_j > min_depth.
 * sample_depth_gradmag: zero-padded bilinear lookups at n points of z_img (coords_z -> z_out) and of
 *   |Scharr(log z_img)| (coords_g -> g_out); either pair may be NULL. */
int como_b200_reproject_dense(const double* z_img, int32_t H, int32_t W, const double* T_ji12, const double* intr4,
                              double min_depth, double* coords_j, double* logz_j, double* z_j, uint8_t* mask,
                              void* stream);
int como_b200_sample_depth_gradmag(const double* z_img, int32_t H, int32_t W, const double* coords_z,
                                   const double* coords_g, int32_t n, double* z_out, double* g_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Tracker front-end (fp32).
 * ------------------------------------------------------------------------------------------ */

/* Tracking.prep_tracking_img (como/odom/Tracking.py:88-94; utils/image_processing.py:48-87): rgb (3,H,W) ->
 * gray pyramid.  levels: HOST array of num_levels DEVICE pointers, coarsest first (reference order). */
int como_b200_gray_pyramid(const float* rgb, int32_t H, int32_t W, int32_t num_levels, float* const* levels,
                           void* stream);
/* Fused front-end: Tracking.prep_tracking_img + get_img_gradients (como/odom/Tracking.py:88-102;
 * utils/image_processing.py:8-87) in ONE launch: rgb (3,H,W) -> gray -> every pyramid level (-> Scharr/32 gradients
 * of every level when gx/gy are given).  levels / gx / gy: HOST arrays of num_levels DEVICE pointers, coarsest
 * first; gx = gy = NULL skips the gradients.  num_levels <= 4 (deeper pyramids: gray_pyramid + image_gradients). */
int como_b200_image_pyramid_fused(const float* rgb, int32_t H, int32_t W, int32_t num_levels, float* const* levels,
                                  float* const* gx, float* const* gy, void* stream);
/* ImageGradientModule (utils/image_processing.py:8-44): Scharr/32, reflect padding. */
int como_b200_image_gradients(const float* img, int32_t h, int32_t w, float* gx, float* gy, void* stream);

/* Mapping.get_img_and_grads (como/odom/Mapping.py:368-376), color "gray": rgb (3,H,W) f64 -> [I, gx, gy] (3,H,W)
 * f64 in one pass (gray = 0.2989 R + 0.587 G + 0.114 B, Scharr with reflect padding). */
int como_b200_img_and_grads_f64(const double* rgb, int32_t h, int32_t w, double* img_and_grads, void* stream);
/* One pyramid level of Tracking.update_kf_reference (Tracking.py:243-314) for one keyframe: nearest depth
 * (stride `sub` into the full-resolution depth), back-projection, rel (3x4, DEVICE) into the last keyframe,
 * +-`border` px / depth mask, precalc_jacobians.  Outputs in the reference layout: vals (n), grads (n,2),
 * P (n,3), J (n,8), mask (n) with n = h*w.  K9 HOST. */
int como_b200_kf_reference_level(const float* img, const float* gx, const float* gy, const float* depth_full,
                                 int32_t Hf, int32_t Wf, int32_t sub, int32_t h, int32_t w, const float* K9,
                                 const float* rel12_dev, float border, float depth_thresh, float* vals, float* grads,
                                 float* P, float* J, uint8_t* mask, void* stream);
/* get_reproj_last_kf (Tracking.py:169-188): depth image (h*w, NaN where empty) of the cloud P (n,3) seen
 * from T (3x4 or 4x4 row-major, DEVICE); duplicates resolved as "largest point index wins". winner_ws: h*w int32. */
int como_b200_reproj_depth(const float* P, int32_t n, const float* T_dev, const float* K9, int32_t h, int32_t w,
                           int32_t* winner_ws, float* depth_img, void* stream);

/* ------------------------------------------------------------------------------------------
 * Inter-process hand-off: replaces the per-message tensor shipping of TupleTensorQueue.push
 * (como/utils/multiprocessing.py:16-21,45-51: every tensor .to(device, dtype), then pickled through mp.Queue, i.e. a
 * fresh CUDA IPC handle per tensor per message).  One launch packs all tensors of a message into a persistent
 * device slot, converting to the consumer's dtype on the way; the slot ring is shared with the consumer process once.
 * ------------------------------------------------------------------------------------------ */
#define COMO_B200_DT_F32 0
#define COMO_B200_DT_F64 1
#define COMO_B200_DT_U8 2
#define COMO_B200_DT_I32 3
#define COMO_B200_DT_I64 4
#define COMO_B200_PACK_MAX_ITEMS 8
typedef struct {
  const void* src;           /* DEVICE pointer, contiguous, same device as the slot */
  int64_t dst_offset_bytes;  /* multiple of 16 */
  int64_t count;             /* elements */
  int32_t src_dtype;         /* COMO_B200_DT_* */
  int32_t dst_dtype;         /* COMO_B200_DT_F32 or COMO_B200_DT_F64 */
} como_b200_pack_item_t;
/* items: HOST array (passed by value to the kernel). */
int como_b200_handoff_pack(const como_b200_pack_item_t* items, int32_t num_items, void* slot, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COMO_B200_H */
