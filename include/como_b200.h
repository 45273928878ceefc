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

/* ------------------------------------------------------------------------------------------
 * Tracking: frame-to-keyframe inverse-compositional photometric GN (fp32, C = 1 channel).
 * Replaces como/odom/frontend/photo_tracking.py:10-42 (photo_tracking_pyr), :147-185
 * (photo_level_tracking), :117-143 (tracking_iter), :77-114 (robustify/solve/update) and the
 * torch ops under them (geometry/camera.py:57-68, frontend/photo_utils.py:9-31, lietorch SE3.exp).
 * One persistent cooperative launch runs the whole coarse-to-fine loop incl. termination tests.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const float* vals;   /* (n)     reference intensities I_i            [photo_tracking.py:24 vals_i]  */
  const float* P;      /* (n,3)   reference points in the KF frame     [Pi]                           */
  const float* J;      /* (n,8)   precomputed dI/d[xi,a,b]; cols 6,7 are ignored (rebuilt per iter)   */
  const uint8_t* mask; /* (n) 0/1 or NULL: which points take part      [masks]                        */
  const float* img;    /* (h,w)   target image of this level           [img_j]                        */
  int32_t n, w, h;
  float K[9];          /* row-major 3x3 intrinsics of this level       [intrinsics]                   */
} como_b200_track_level_t;

typedef struct {
  int32_t max_iter;    /* term_criteria["max_iter"]   (config/como.yml:12-17) */
  float delta_norm;    /* term_criteria["delta_norm"] */
  float rel_tol;       /* term_criteria["rel_tol"]    */
  float grad_norm;     /* term_criteria["grad_norm"]  */
} como_b200_track_term_t;

/* per-iteration record written to `stats` (8 floats per iteration, iteration-major) */
#define COMO_B200_TRACK_STAT_STRIDE 8
/* [0]=level [1]=mean_sq_err [2]=grad_norm [3]=delta_norm [4]=sigma_r [5]=num_valid [6]=done [7]=reserved */

size_t como_b200_track_workspace_bytes(int32_t max_n, int32_t num_problems);

/* levels: HOST array [num_problems][num_levels] (coarsest first, as the reference stores pyramids).
 * T: DEVICE (num_problems,4,4) in/out Tji; aff: DEVICE (num_problems,2) in/out [a,b].
 * stats: DEVICE (num_problems, num_levels*max_iter, 8) or NULL; num_iters: DEVICE (num_problems) int32 or NULL. */
int como_b200_track_pyr(const como_b200_track_level_t* levels, int32_t num_levels, int32_t num_problems,
                        const como_b200_track_term_t* term, float* T, float* aff, float* stats,
                        int32_t* num_iters, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces precalc_jacobians (como/odom/frontend/photo_tracking.py:46-74), C = 1.
 * grads (n,2) [gx,gy]; P (n,3); vals (n); K 9 floats (host); out J (n,8). */
int como_b200_precalc_jacobians(const float* grads, const float* P, const float* vals, const float* K,
                                int64_t n, float* J, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COMO_B200_H */
