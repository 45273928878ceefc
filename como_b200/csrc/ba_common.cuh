// Shared definitions of the window bundle-adjustment kernels (fp64).
#pragma once
#include "common.cuh"

namespace como {

constexpr int BA_MAXM = 64;        // anchors per keyframe (sampling.max_num_coords), padded width
constexpr int SCAF_STRIDE = 16;    // doubles per (keyframe, anchor slot) in the scaffold buffer
// scaffold record: [0] logz [1] u=1/z [2] pm.x [3] pm.y [4..6] Pc [7] zmask(0/1) [8..13] dlogz/dTwc [14..15] pad
constexpr int REF_STRIDE = 8;      // doubles per (keyframe, pixel): z_n, q_n[6], pad
constexpr int PAIR_STRIDE = 4;     // doubles per (pair, pixel): dI/dPc[3], scaled reference value

struct BAFrame {                   // per frame f in [0, K+R): keyframes first, then one-way frames
  double Rwc[9], twc[3];
  double Rcw[9], tcw[3];
  double a, b;
  const double* img;               // (3,H,W): intensity, gx, gy
};

struct BADims {
  int K, R, L, M, N, H, W, P;      // keyframes, one-way frames, landmarks, anchors, pixels/kf, image, pairs
  double fx, fy, cx, cy;
};

__device__ __forceinline__ void mat3_vec(const double* R, const double* v, double* o) {
  o[0] = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  o[1] = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  o[2] = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
}
__device__ __forceinline__ void mat3T_vec(const double* R, const double* v, double* o) {
  o[0] = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  o[1] = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  o[2] = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
}
// o = v^T [a]_x  (row vector times skew matrix) = (a x ... ) : (v^T [a]_x)_j = sum_i v_i eps... = -(a x v)^T... use explicit form
__device__ __forceinline__ void row_times_skew(const double* v, const double* a, double* o) {
  // [a]_x = [[0,-a2,a1],[a2,0,-a0],[-a1,a0,0]];  o_j = sum_i v_i [a]_x(i,j)
  o[0] = v[1] * a[2] - v[2] * a[1];
  o[1] = -v[0] * a[2] + v[2] * a[0];
  o[2] = v[0] * a[1] - v[1] * a[0];
}

__device__ __forceinline__ double atomic_add_f64(double* p, double v) { return atomicAdd(p, v); }

}  // namespace como
