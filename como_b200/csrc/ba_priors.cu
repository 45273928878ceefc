// Prior factors added to the BA normal equations every iteration (fp64), one CTA per keyframe:
//   gp_ml_cost (sigma 1)                    como/odom/factors/gp_priors.py:7-80
//   log_depth_prior (mode first_mean)       como/odom/factors/depth_prior.py:7-141
//   pixel_prior_cost (mode first, 1e-2 px)  como/odom/factors/pixel_prior.py:6-130
// and on keyframe 0: pose anchor (pose_prior_factors.py:5-19 with the reference's SE3_logmap,
// geometry/lie_algebra.py:117-176), affine anchors (scalar_prior_factors.py:4-19), then either the
// mean-log-depth scale prior (gp_priors.py:83-150) or the frozen-landmark prior (scalar_prior_factors.py:22-34).
//
// With W = L^-T L^-1 (cached per keyframe) the GP factor needs no triangular solve per iteration:
//   r = L^-1 d,  J_P = L^-1 diag(u) (x) dz/dPw,  J_T = L^-1 dlogz/dT
//   => g_P = -(W d) u d3, g_T = -dT^T (W d), H_PP = W o (u d3)(u d3)^T, H_TT = dT^T W dT, H_TP = (W dT)^T u d3.
#include "ba_common.cuh"

namespace como {

struct PriorParams {
  double info_pixel;      // float32-rounded 1/(1e-2)^2 (reference scratch tensor is float32)
  double info_pose_H;     // float32(1/sigma)^2 product as the reference forms it (J^T J in float32)
  double info_pose;       // (1/sigma)^2 in double (gradient and error)
  double info_scalar;     // 1/scale_prior^2
  double info_mean_depth; // 1/mean_depth_prior^2
  double scale_anchor;    // init_scale_anchor
  int window_full;
  int nfix;
};

__device__ void se3_log_reference(const double* T, double* xi) {
  // SO3_logmap + the reference's V^-1 t expression (elementwise (0.5 t) * (w_n x t) term kept as is)
  const double tr = T[0] + T[5] + T[10];
  const double tr3 = tr - 3.0;
  const double theta = acos(0.5 * (tr - 1.0));
  const double mag = (tr3 < -1e-6) ? theta / (2.0 * sin(theta)) : 0.5 - tr3 / 12.0 + tr3 * tr3 / 60.0;
  const double w[3] = {mag * (T[9] - T[6]), mag * (T[2] - T[8]), mag * (T[4] - T[1])};
  double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  if (th < 1e-6) th = 1e-6;
  const double wn[3] = {w[0] / th, w[1] / th, w[2] / th};
  const double tn = tan(0.5 * th);
  const double t[3] = {T[3], T[7], T[11]};
  const double c1[3] = {wn[1] * t[2] - wn[2] * t[1], wn[2] * t[0] - wn[0] * t[2], wn[0] * t[1] - wn[1] * t[0]};
  const double c2[3] = {wn[1] * c1[2] - wn[2] * c1[1], wn[2] * c1[0] - wn[0] * c1[2], wn[0] * c1[1] - wn[1] * c1[0]};
  const double k = 1.0 - th / (2.0 * tn);
  for (int q = 0; q < 3; ++q) {
    xi[q] = w[q];
    xi[3 + q] = t[q] - (0.5 * t[q]) * c1[q] + k * c2[q];
  }
}

__global__ void __launch_bounds__(256)
ba_priors_kernel(const double* __restrict__ scaf, const double* __restrict__ dz_dP, const double* __restrict__ LtL,
                 const double* __restrict__ med, const uint8_t* __restrict__ obs_ref,
                 const double* __restrict__ pm_first_obs, const int32_t* __restrict__ lm_ids,
                 const double* __restrict__ kf_poses, const double* __restrict__ pose_anchor,
                 const double* __restrict__ kf_aff, const double* __restrict__ aff_anchor,
                 const double* __restrict__ colmean, const double* __restrict__ P_m,
                 const double* __restrict__ P_anchor, const int32_t* __restrict__ fix_ids, PriorParams pp, BADims d,
                 int dim, double* __restrict__ H, double* __restrict__ g, double* __restrict__ err) {
  __shared__ double s_d[BA_MAXM], s_w[BA_MAXM], s_u[BA_MAXM], s_dT[BA_MAXM][6], s_Y[BA_MAXM][6];
  __shared__ double s_TT[36], s_gT[6], s_err[4];
  __shared__ int s_lm[BA_MAXM];
  // grid (K, NSLICE): the two big 3M x 3M scatter loops are split over blockIdx.y; everything that must happen
  // once per keyframe is done by slice 0
  const int k = blockIdx.x, tid = threadIdx.x, M = d.M;
  const int slice = blockIdx.y, nslice = gridDim.y;
  const bool lead = slice == 0;
  const double logmed = log(med[k]);
  const double d3[3] = {dz_dP[3 * k], dz_dP[3 * k + 1], dz_dP[3 * k + 2]};
  const int lm_start = 8 * (d.K + d.R);
  const double* W = LtL + (size_t)k * M * M;
  if (tid < M) {
    const double* o = scaf + ((size_t)k * M + tid) * SCAF_STRIDE;
    s_d[tid] = o[0] - logmed;
    s_u[tid] = o[1];
    for (int a = 0; a < 6; ++a) s_dT[tid][a] = o[8 + a];
    s_lm[tid] = lm_ids[k * M + tid];
  }
  if (tid < 36) s_TT[tid] = 0.0;
  if (tid < 6) s_gT[tid] = 0.0;
  if (tid < 4) s_err[tid] = 0.0;
  __syncthreads();
  // w = W d ; Y = W dT
  for (int t = tid; t < M * 7; t += 256) {
    const int m = t / 7, a = t % 7;
    double s = 0.0;
    if (a == 0) {
      for (int q = 0; q < M; ++q) s += W[m * M + q] * s_d[q];
      s_w[m] = s;
    } else {
      for (int q = 0; q < M; ++q) s += W[m * M + q] * s_dT[q][a - 1];
      s_Y[m][a - 1] = s;
    }
  }
  __syncthreads();
  const int M3 = 3 * M;
  // ---- GP marginal-likelihood prior
  for (int t = tid + 256 * slice; t < M3 * M3; t += 256 * nslice) {
    const int ra = t / M3, rb = t % M3;
    const int m = ra / 3, c = ra % 3, m2 = rb / 3, c2 = rb % 3;
    atomicAdd(H + (size_t)(lm_start + 3 * s_lm[m] + c) * dim + (lm_start + 3 * s_lm[m2] + c2),
              W[m * M + m2] * s_u[m] * d3[c] * s_u[m2] * d3[c2]);
  }
  if (!lead) {
    // slices > 0 only help with the big blocks (incl. the keyframe-0 scale prior below)
    if (k == 0 && !pp.window_full) {
      const double im = pp.info_mean_depth;
      for (int t = tid + 256 * slice; t < M3 * M3; t += 256 * nslice) {
        const int ra = t / M3, rb = t % M3;
        const int m = ra / 3, c = ra % 3, m2 = rb / 3, c2 = rb % 3;
        atomicAdd(H + (size_t)(lm_start + 3 * s_lm[m] + c) * dim + (lm_start + 3 * s_lm[m2] + c2),
                  im * colmean[m] * s_u[m] * d3[c] * colmean[m2] * s_u[m2] * d3[c2]);
      }
    }
    return;
  }
  for (int t = tid; t < M3; t += 256) {
    const int m = t / 3, c = t % 3;
    atomicAdd(g + lm_start + 3 * s_lm[m] + c, -s_w[m] * s_u[m] * d3[c]);
  }
  for (int t = tid; t < 6 * M3; t += 256) {
    const int a = t / M3, rb = t % M3;
    const int m = rb / 3, c = rb % 3;
    const double v = s_Y[m][a] * s_u[m] * d3[c];
    const size_t ri = 8 * (size_t)k + a, ci = lm_start + 3 * (size_t)s_lm[m] + c;
    atomicAdd(H + ri * dim + ci, v);
    atomicAdd(H + ci * dim + ri, v);
  }
  if (tid < 36) {
    const int a = tid / 6, b = tid % 6;
    double s = 0.0;
    for (int m = 0; m < M; ++m) s += s_dT[m][a] * s_Y[m][b];
    atomicAdd(&s_TT[tid], s);
  } else if (tid < 42) {
    const int a = tid - 36;
    double s = 0.0;
    for (int m = 0; m < M; ++m) s += s_dT[m][a] * s_w[m];
    atomicAdd(&s_gT[a], -s);
  } else if (tid == 42) {
    double s = 0.0;
    for (int m = 0; m < M; ++m) s += s_d[m] * s_w[m];
    atomicAdd(&s_err[0], s);
  }
  // ---- first-observation log-depth prior + pixel prior, one thread per anchor slot
  if (tid < M && obs_ref[k * M + tid]) {
    const int m = tid;
    const double* o = scaf + ((size_t)k * M + m) * SCAF_STRIDE;
    const double u = o[1];
    const double Pc[3] = {o[4], o[5], o[6]};
    const double* T = kf_poses + 16 * (size_t)k;
    // row Jacobians: log depth (1 row), pixel (2 rows)
    double JP[3][3], JT[3][6], r[3], info[3];
    for (int c = 0; c < 3; ++c) JP[0][c] = u * d3[c];
    for (int a = 0; a < 6; ++a) JT[0][a] = s_dT[m][a];
    r[0] = s_d[m];
    info[0] = 1.0;
    const double z = Pc[2];
    const double dpi[2][3] = {{d.fx / z, 0.0, -d.fx * Pc[0] / z / z}, {0.0, d.fy / z, -d.fy * Pc[1] / z / z}};
    // dPc/dTwc = [Pc^ | -I]
    const double sk[3][3] = {{0, -Pc[2], Pc[1]}, {Pc[2], 0, -Pc[0]}, {-Pc[1], Pc[0], 0}};
    for (int rr = 0; rr < 2; ++rr) {
      for (int c = 0; c < 3; ++c) {
        // R_cw[q][c] = T[c*4+q]
        JP[1 + rr][c] = dpi[rr][0] * T[c * 4 + 0] + dpi[rr][1] * T[c * 4 + 1] + dpi[rr][2] * T[c * 4 + 2];
        JT[1 + rr][c] = dpi[rr][0] * sk[0][c] + dpi[rr][1] * sk[1][c] + dpi[rr][2] * sk[2][c];
        JT[1 + rr][3 + c] = -dpi[rr][c];
      }
      r[1 + rr] = o[2 + rr] - pm_first_obs[2 * ((size_t)k * M + m) + rr];
      info[1 + rr] = pp.info_pixel;
    }
    const size_t l3 = lm_start + 3 * (size_t)s_lm[m];
    double eld = r[0] * r[0], epx = pp.info_pixel * (r[1] * r[1] + r[2] * r[2]);
    atomicAdd(&s_err[1], eld);
    atomicAdd(&s_err[2], epx);
    for (int c = 0; c < 3; ++c) {
      double gp = 0.0;
      for (int q = 0; q < 3; ++q) gp += info[q] * JP[q][c] * r[q];
      atomicAdd(g + l3 + c, -gp);
      for (int c2 = 0; c2 < 3; ++c2) {
        double h = 0.0;
        for (int q = 0; q < 3; ++q) h += info[q] * JP[q][c] * JP[q][c2];
        atomicAdd(H + (l3 + c) * dim + l3 + c2, h);
      }
    }
    for (int a = 0; a < 6; ++a) {
      double gt = 0.0;
      for (int q = 0; q < 3; ++q) gt += info[q] * JT[q][a] * r[q];
      atomicAdd(&s_gT[a], -gt);
      for (int b = 0; b < 6; ++b) {
        double h = 0.0;
        for (int q = 0; q < 3; ++q) h += info[q] * JT[q][a] * JT[q][b];
        atomicAdd(&s_TT[a * 6 + b], h);
      }
      for (int c = 0; c < 3; ++c) {
        double h = 0.0;
        for (int q = 0; q < 3; ++q) h += info[q] * JT[q][a] * JP[q][c];
        const size_t ri = 8 * (size_t)k + a;
        atomicAdd(H + ri * dim + l3 + c, h);
        atomicAdd(H + (l3 + c) * dim + ri, h);
      }
    }
  }
  __syncthreads();
  if (tid < 36) atomicAdd(H + (8 * (size_t)k + tid / 6) * dim + 8 * (size_t)k + tid % 6, s_TT[tid]);
  if (tid < 6) atomicAdd(g + 8 * (size_t)k + tid, s_gT[tid]);
  if (tid < 3) atomicAdd(err + 1 + tid, s_err[tid]);

  // ---- keyframe 0: anchors and the scale / frozen-landmark prior
  if (k != 0) return;
  if (tid == 0) {
    const double* T = kf_poses;
    double Ti[16], D[16], xi[6];
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) Ti[r * 4 + c] = T[c * 4 + r];
      Ti[r * 4 + 3] = -(T[0 * 4 + r] * T[3] + T[1 * 4 + r] * T[7] + T[2 * 4 + r] * T[11]);
    }
    Ti[12] = Ti[13] = Ti[14] = 0.0;
    Ti[15] = 1.0;
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) {
        double s = 0.0;
        for (int q = 0; q < 4; ++q) s += Ti[r * 4 + q] * pose_anchor[q * 4 + c];
        D[r * 4 + c] = s;
      }
    se3_log_reference(D, xi);
    double e = 0.0;
    for (int a = 0; a < 6; ++a) {
      const double x = -xi[a];
      atomicAdd(H + (size_t)a * dim + a, pp.info_pose_H);
      atomicAdd(g + a, -pp.info_pose * x);
      e += pp.info_pose * x * x;
    }
    atomicAdd(err + 4, e);
    double ea = 0.0;
    for (int c = 0; c < 2; ++c) {
      const double rv = kf_aff[c] - aff_anchor[c];
      atomicAdd(g + 6 + c, -pp.info_scalar * rv);
      atomicAdd(H + (size_t)(6 + c) * dim + 6 + c, pp.info_scalar);
      ea += pp.info_scalar * rv * rv;
    }
    atomicAdd(err + 5, ea);
  }
  if (pp.window_full) {
    double e = 0.0;
    for (int t = tid; t < 3 * pp.nfix; t += 256) {
      const int j = t / 3, c = t % 3;
      const int l = fix_ids[j];
      const double rv = P_m[3 * (size_t)l + c] - P_anchor[3 * (size_t)j + c];
      const size_t idx = lm_start + 3 * (size_t)l + c;
      atomicAdd(g + idx, -pp.info_scalar * rv);
      atomicAdd(H + idx * dim + idx, pp.info_scalar);
      e += pp.info_scalar * rv * rv;
    }
    e = warp_sum(e);
    if ((tid & 31) == 0) atomicAdd(err + 7, e);
  } else {
    // mean log depth of keyframe 0: r = colmean . logz - anchor ; J_P = colmean u d3 ; J_T = colmean^T dT
    __shared__ double s_JT[6], s_rv;
    if (tid < 6) {
      double s = 0.0;
      for (int m = 0; m < M; ++m) s += colmean[m] * s_dT[m][tid];
      s_JT[tid] = s;
    } else if (tid == 32) {
      double s = 0.0;
      for (int m = 0; m < M; ++m) s += colmean[m] * (s_d[m] + logmed);
      s_rv = s - pp.scale_anchor;
    }
    __syncthreads();
    const double rv = s_rv, im = pp.info_mean_depth;
    for (int t = tid; t < M3 * M3; t += 256 * nslice) {   // slice 0's share
      const int ra = t / M3, rb = t % M3;
      const int m = ra / 3, c = ra % 3, m2 = rb / 3, c2 = rb % 3;
      atomicAdd(H + (size_t)(lm_start + 3 * s_lm[m] + c) * dim + (lm_start + 3 * s_lm[m2] + c2),
                im * colmean[m] * s_u[m] * d3[c] * colmean[m2] * s_u[m2] * d3[c2]);
    }
    for (int t = tid; t < M3; t += 256) {
      const int m = t / 3, c = t % 3;
      const double jp = colmean[m] * s_u[m] * d3[c];
      const size_t ci = lm_start + 3 * (size_t)s_lm[m] + c;
      atomicAdd(g + ci, -im * jp * rv);
      for (int a = 0; a < 6; ++a) {
        const double v = im * s_JT[a] * jp;
        atomicAdd(H + (size_t)a * dim + ci, v);
        atomicAdd(H + ci * dim + a, v);
      }
    }
    if (tid < 36) atomicAdd(H + (size_t)(tid / 6) * dim + tid % 6, im * s_JT[tid / 6] * s_JT[tid % 6]);
    if (tid < 6) atomicAdd(g + tid, -im * s_JT[tid] * rv);
    if (tid == 0) atomicAdd(err + 6, im * rv * rv);
  }
}

}  // namespace como

using namespace como;

extern "C" int como_b200_ba_priors(const double* scaffold, const double* dz_dP, const double* LtL,
                                   const double* median_depths, const uint8_t* obs_ref_mask, const double* pm_first_obs,
                                   const int32_t* lm_ids, const double* kf_poses, const double* pose_anchor,
                                   const double* kf_aff, const double* aff_anchor, const double* colmean,
                                   const double* P_m, const double* P_m_anchors, const int32_t* fix_ids, int32_t nfix,
                                   int32_t window_full, const double* sigmas4, double scale_anchor, int32_t K, int32_t R,
                                   int32_t L, int32_t M, const double* intr4, int32_t dim, double* H, double* g,
                                   double* err8, void* stream) {
  COMO_REQUIRE(scaffold && dz_dP && LtL && median_depths && obs_ref_mask && pm_first_obs && lm_ids && kf_poses &&
                   pose_anchor && kf_aff && aff_anchor && P_m && sigmas4 && intr4 && H && g && err8,
               "ba_priors: null pointer argument");
  COMO_REQUIRE(window_full ? (nfix == 0 || (P_m_anchors && fix_ids)) : (colmean != nullptr),
               "ba_priors: missing anchors (window full) or column means (window not full)");
  COMO_REQUIRE(M >= 1 && M <= BA_MAXM, "ba_priors: M out of range");
  BADims d{};
  d.K = K; d.R = R; d.L = L; d.M = M;
  d.fx = intr4[0]; d.fy = intr4[1]; d.cx = intr4[2]; d.cy = intr4[3];
  PriorParams pp;
  // sigmas4 = [pixel_sigma_first, pose_prior, scale_prior, mean_depth_prior]
  pp.info_pixel = (double)(float)(1.0 / (sigmas4[0] * sigmas4[0]));
  const float isq = (float)(1.0 / sigmas4[1]);
  pp.info_pose_H = (double)(isq * isq);
  pp.info_pose = (1.0 / sigmas4[1]) * (1.0 / sigmas4[1]);
  pp.info_scalar = (1.0 / sigmas4[2]) * (1.0 / sigmas4[2]);
  pp.info_mean_depth = 1.0 / (sigmas4[3] * sigmas4[3]);
  pp.scale_anchor = scale_anchor;
  pp.window_full = window_full;
  pp.nfix = nfix;
  ba_priors_kernel<<<dim3(K, 8), 256, 0, (cudaStream_t)stream>>>(scaffold, dz_dP, LtL, median_depths, obs_ref_mask, pm_first_obs,
                                                        lm_ids, kf_poses, pose_anchor, kf_aff, aff_anchor, colmean, P_m,
                                                        P_m_anchors, fix_ids, pp, d, dim, H, g, err8);
  return check_launch("ba_priors");
}
