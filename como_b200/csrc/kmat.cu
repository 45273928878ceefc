// DepthCov K-matrices for the GP depth predictor.
//
// Reference: Mapping.prep_predictor (como/odom/Mapping.py:430-468) and calc_kernel_matrices / get_predictor
// (como/depth_cov/core/distill_depth.py:8-48) evaluate the Matern-3/2 probability-product kernel
// (como/depth_cov/core/kernels.py:22-89, covariance.py:10-51) between n test points and m <= 64 anchors
// with stock torch ops (an n x m matrix in HBM, ~25 temporaries of that size) and then multiply by K_mm^-1.
//
// Here one kernel produces the predictor rows  K_nm K_mm^-1  (and optionally the predictive variance
// K_nn - K_nm K_mm^-1 K_mn) without ever writing K_nm:
//   * a warp owns 16 test points at a time: lanes 0..15 fetch the per-point quantities (covariance parameters by
//     direct read on the pixel grid or border-clamped bilinear lookup at fractional coordinates, 4th root of
//     det E) once, then every lane evaluates the kernel against its two anchors for each of the 16 points
//     (anchor constants live in registers for the whole kernel) and stores the 16 x 64 K_nm tile in the warp's
//     private shared-memory slab;
//   * the (16 x 64) x (64 x 64) product with K_mm^-1 runs on the FP64 tensor path (mma.sync m8n8k4 f64, SASS
//     DMMA.8x8x4): 2 row tiles x 8 column tiles of accumulators per warp, A fragments from the slab, B
//     fragments from a CTA-wide copy of K_mm^-1; both arrays use a pitch of 68 doubles, which makes every
//     fragment load bank-conflict free;
//   * warps never synchronise with each other inside the loop (only __syncwarp), so the scalar kernel
//     evaluation of one warp overlaps the DMMA phase of another; 2 CTAs (16 warps) per SM.
// Bound: FP64 pipe (measured 37.1 TFLOP/s DFMA == DMMA on B200): 2 n m^2 GEMM flop + ~75 fp64 instructions per
// kernel evaluation; compulsory HBM write n*m*8 B.
#include "ba_common.cuh"

namespace como {

constexpr int KM = BA_MAXM;     // padded anchor count
constexpr int KPITCH = 68;      // == 4 (mod 16): conflict-free DMMA fragment loads
constexpr int KW_ROWS = 16;     // test points per warp step
constexpr int KWARPS = 8;
constexpr double SQRT3 = 1.7320508075688772;

// border-clamped bilinear lookup of the 4-channel covariance image (gaussian_kernel.py:52-79); coords are
// pixel (row, col): the reference's normalise / unnormalise round trip cancels up to rounding.
__device__ __forceinline__ void interp_cov4(const double* __restrict__ img, int H, int W, double row, double col,
                                            double* E) {
  double x = fmin(fmax(col, 0.0), (double)(W - 1));
  double y = fmin(fmax(row, 0.0), (double)(H - 1));
  const double x0f = floor(x), y0f = floor(y);
  const int x0 = (int)x0f, y0 = (int)y0f;
  const double fx = x - x0f, fy = y - y0f;
  const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const double* p = img + (size_t)c * H * W;
    E[c] = p[(size_t)y0 * W + x0] * (1 - fx) * (1 - fy) + p[(size_t)y0 * W + x1] * fx * (1 - fy) +
           p[(size_t)y1 * W + x0] * (1 - fx) * fy + p[(size_t)y1 * W + x1] * fx * fy;
  }
}

// Python-formula kernel value (kernels.py:22-68 + matern 82-89 + covariance.py:33-39).  The reference casts the
// coordinate difference to float32 and squares it in float32 (`.float()`, torch.square); the cross term and
// everything else promote to double.  r1, r2 = det(E)^(1/4) of the two points.
__device__ __forceinline__ double cov_python(double x1r, double x1c, double a00, double a01, double a11, double r1,
                                             double x2r, double x2c, double b00, double b01, double b11, double r2,
                                             double scale) {
  const float d0f = (float)(x1r - x2r), d1f = (float)(x1c - x2c);
  const double d0 = (double)d0f, d1 = (double)d1f;
  const double q0 = (double)(d0f * d0f), q1 = (double)(d1f * d1f);
  const double s00 = a00 + b00, s01 = a01 + b01, s11 = a11 + b11;
  double Q = s11 * q0;
  Q += ((-2.0 * s01) * d0) * d1;
  Q += s00 * q1;
  const double det = s00 * s11 - s01 * s01;
  Q = (Q / det) * 0.5;
  const double Cc = (2.0 * r1) * r2 / sqrt(det + 1e-8);
  const double t = SQRT3 * sqrt(Q + 1e-8);
  return (((1.0 + t) * exp(-t)) * Cc) * scale;
}

// normalised coordinate as the reference feeds it (utils/coords.py normalize_coordinates): 2*A*x + A - 1, A = 1/dim
// Every operation is rounded separately (no FMA contraction): the difference of two such coordinates is cast to
// float32 by the reference, and a 1-ulp change here can flip that rounding.
__device__ __forceinline__ double norm_coord(double x, int dim) {
  const double A = 1.0 / (double)dim;
  return __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(2.0, A), x), A), -1.0);
}

// E_m at the anchors + K_mm (+ jitter on the diagonal).  One CTA per keyframe.
__global__ void kmm_kernel(const double* __restrict__ cov_img, int H, int W, const double* __restrict__ coords_m, int M,
                           double scale, double jitter, double* __restrict__ E_m, double* __restrict__ K_mm) {
  const int b = blockIdx.x;
  const double* img = cov_img + (size_t)b * 4 * H * W;
  extern __shared__ double sE[];  // M*4 + M*2 + M
  double* sx = sE + 4 * M;
  double* sr = sx + 2 * M;
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const double r = coords_m[((size_t)b * M + m) * 2], c = coords_m[((size_t)b * M + m) * 2 + 1];
    double E[4];
    interp_cov4(img, H, W, r, c, E);
    for (int q = 0; q < 4; ++q) {
      sE[4 * m + q] = E[q];
      E_m[((size_t)b * M + m) * 4 + q] = E[q];
    }
    sr[m] = sqrt(sqrt(E[0] * E[3] - E[1] * E[2]));
    sx[2 * m] = norm_coord(r, H);
    sx[2 * m + 1] = norm_coord(c, W);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < M * M; t += blockDim.x) {
    const int i = t / M, j = t % M;
    double v = cov_python(sx[2 * i], sx[2 * i + 1], sE[4 * i], sE[4 * i + 1], sE[4 * i + 3], sr[i], sx[2 * j],
                          sx[2 * j + 1], sE[4 * j], sE[4 * j + 1], sE[4 * j + 3], sr[j], scale);
    if (i == j) v += jitter;
    K_mm[(size_t)b * M * M + t] = v;
  }
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (__longlong_as_double((long long)old) > v) {
    const unsigned long long assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}

struct KmatSmem {
  double Kinv[KM * KPITCH];
  double K[KWARPS][KW_ROWS * KPITCH];
  double px[KWARPS][KW_ROWS][8];  // xr, xc, e00, e01, e11, r1, k_nn, valid
};

// GRID: test points are the pixel grid of the covariance image (coords_n == nullptr, n == H*W, E_n read directly);
// otherwise coords_n (B, n, 2) [row, col] with optional validity mask (B, n) (0 -> row of zeros, var untouched).
template <bool GRID>
__global__ void __launch_bounds__(32 * KWARPS, 2)
kmat_rows_kernel(const double* __restrict__ cov_img, int H, int W, const double* __restrict__ coords_m,
                 const double* __restrict__ E_m, const double* __restrict__ Kinv, int M, double scale,
                 const double* __restrict__ coords_n, const unsigned char* __restrict__ mask_n, long long n,
                 double* __restrict__ out, double* __restrict__ var_out, double* __restrict__ var_min) {
  extern __shared__ __align__(16) unsigned char kraw[];
  KmatSmem& S = *reinterpret_cast<KmatSmem*>(kraw);
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* img = cov_img + (size_t)b * 4 * H * W;
  const long long HW = (long long)H * W;
  for (int t = tid; t < KM * KM; t += blockDim.x) {
    const int i = t / KM, j = t % KM;
    S.Kinv[i * KPITCH + j] = (i < M && j < M) ? Kinv[((size_t)b * M + i) * M + j] : 0.0;
  }
  // anchor constants of this lane: anchors lane and lane + 32
  double ax[2], ay[2], a00[2], a01[2], a11[2], ar[2];
  bool aok[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int m = lane + 32 * h;
    aok[h] = m < M;
    const int mm = aok[h] ? m : 0;
    const double* e = E_m + ((size_t)b * M + mm) * 4;
    a00[h] = e[0];
    a01[h] = e[1];
    a11[h] = e[3];
    ar[h] = sqrt(sqrt(e[0] * e[3] - e[1] * e[2]));
    ax[h] = norm_coord(coords_m[((size_t)b * M + mm) * 2], H);
    ay[h] = norm_coord(coords_m[((size_t)b * M + mm) * 2 + 1], W);
  }
  __syncthreads();
  double* sK = S.K[warp];
  double(*spx)[8] = S.px[warp];
  const int g4 = lane >> 2, l4 = lane & 3;
  const int ksteps = (M + 3) / 4;
  double vmin = 1e300;
  const long long ngroups = (n + KW_ROWS - 1) / KW_ROWS;
  for (long long g = (long long)blockIdx.x * KWARPS + warp; g < ngroups; g += (long long)gridDim.x * KWARPS) {
    const long long p0 = g * KW_ROWS;
    // ---- per-point quantities
    if (lane < KW_ROWS) {
      const long long p = p0 + lane;
      bool valid = p < n;
      double row = 0.0, col = 0.0, E[4] = {1.0, 0.0, 0.0, 1.0};
      if (valid) {
        if (GRID) {
          row = (double)(p / W);
          col = (double)(p % W);
#pragma unroll
          for (int q = 0; q < 4; ++q) E[q] = __ldg(img + (size_t)q * HW + p);
        } else {
          if (mask_n && !mask_n[(size_t)b * n + p]) valid = false;
          if (valid) {
            row = coords_n[((size_t)b * n + p) * 2];
            col = coords_n[((size_t)b * n + p) * 2 + 1];
            interp_cov4(img, H, W, row, col, E);
          }
        }
      }
      const double detE = E[0] * E[3] - E[1] * E[2];
      spx[lane][0] = norm_coord(row, H);
      spx[lane][1] = norm_coord(col, W);
      spx[lane][2] = E[0];
      spx[lane][3] = E[1];
      spx[lane][4] = E[3];
      spx[lane][5] = sqrt(sqrt(detE));
      // diagonal_prob_product (kernels.py:74-79): C = 2 sqrt(det E) / safe_sqrt(det 2E), Q = 0
      const double t0 = SQRT3 * sqrt(1e-8);
      const double e2det = (2.0 * E[0]) * (2.0 * E[3]) - (2.0 * E[1]) * (2.0 * E[2]);
      spx[lane][6] = ((2.0 * sqrt(detE) / sqrt(e2det + 1e-8)) * ((1.0 + t0) * exp(-t0))) * scale;
      spx[lane][7] = valid ? 1.0 : 0.0;
    }
    __syncwarp();
    // ---- K_nm tile: 16 points x (2 anchors per lane)
#pragma unroll 2
    for (int i = 0; i < KW_ROWS; ++i) {
      const double xr = spx[i][0], xc = spx[i][1], e00 = spx[i][2], e01 = spx[i][3], e11 = spx[i][4], r1 = spx[i][5];
      const bool valid = spx[i][7] != 0.0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        double v = 0.0;
        if (valid && aok[h]) v = cov_python(xr, xc, e00, e01, e11, r1, ax[h], ay[h], a00[h], a01[h], a11[h], ar[h], scale);
        sK[i * KPITCH + lane + 32 * h] = v;
      }
    }
    __syncwarp();
    // ---- (16 x 64) x (64 x 64) on the FP64 tensor path
    double acc[2][8][2];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;
    for (int ks = 0; ks < ksteps; ++ks) {
      const int k0 = 4 * ks;
      const double fa0 = sK[g4 * KPITCH + k0 + l4];
      const double fa1 = sK[(8 + g4) * KPITCH + k0 + l4];
      const double* kb = &S.Kinv[(k0 + l4) * KPITCH + g4];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const double fb = kb[8 * c];
        dmma884(acc[0][c][0], acc[0][c][1], fa0, fb);
        dmma884(acc[1][c][0], acc[1][c][1], fa1, fb);
      }
    }
    // ---- epilogue: rows g4 and 8+g4, columns 8c + 2*l4 + {0,1}
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int prow = 8 * r + g4;
      const long long p = p0 + prow;
      double dot = 0.0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int col = 8 * c + 2 * l4;
        dot += sK[prow * KPITCH + col] * acc[r][c][0] + sK[prow * KPITCH + col + 1] * acc[r][c][1];
        if (p < n) {
          double* o = out + ((size_t)b * n + p) * M + col;
          if (col < M) o[0] = acc[r][c][0];
          if (col + 1 < M) o[1] = acc[r][c][1];
        }
      }
      if (var_out) {
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        dot += __shfl_xor_sync(0xffffffffu, dot, 2);
        if (l4 == 0 && p < n && spx[prow][7] != 0.0) {
          const double v = spx[prow][6] - dot;
          var_out[(size_t)b * n + p] = v;
          vmin = fmin(vmin, v);
        }
      }
    }
    __syncwarp();
  }
  if (var_min) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    if (lane == 0 && vmin < 1e300) atomic_min_double(var_min + b, vmin);
  }
}

// Weighted Gram of predictor rows (normal equations of distill_depth, distill_depth.py:51-82 / 126-153):
//   G (M x M) += sum_n w_n k_n k_n^T,   h (M) += sum_n w_n y_n k_n,   with k_n = rows[n, :M].
// w_n = wscale / (var_n + var_add) when var != nullptr, else wscale; rows with mask == 0 are skipped.
// K-dimension = test points: DMMA m8n8k4 with A = (w k)^T fragments and B = k fragments taken from the same
// shared-memory tile (pitch 68).  Each warp accumulates the full lower-triangular set of 8x8 tiles for its
// share of points; CTA partials are combined in shared memory and flushed with fp64 atomics.
constexpr int GR_ROWS = 32;
struct GramSmem {
  double X[KWARPS][GR_ROWS * KPITCH];
  double wy[KWARPS][GR_ROWS][2];
};

__global__ void __launch_bounds__(32 * KWARPS, 1)
weighted_gram_kernel(const double* __restrict__ rows, const double* __restrict__ y, const double* __restrict__ var,
                     const unsigned char* __restrict__ mask, long long n, int M, double var_add, double wscale,
                     double* __restrict__ G, double* __restrict__ hvec, double* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char graw[];
  GramSmem& S = *reinterpret_cast<GramSmem*>(graw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g4 = lane >> 2, l4 = lane & 3;
  double* sX = S.X[warp];
  double(*swy)[2] = S.wy[warp];
  // lower-triangular tiles (i >= j), 36 of them: accumulators acc[t][2]
  double acc[36][2];
#pragma unroll
  for (int t = 0; t < 36; ++t) acc[t][0] = acc[t][1] = 0.0;
  double hacc[2] = {0.0, 0.0};  // lane owns columns lane, lane+32 of h
  double sw = 0.0, swyy = 0.0;
  const long long ngroups = (n + GR_ROWS - 1) / GR_ROWS;
  for (long long g = (long long)blockIdx.x * KWARPS + warp; g < ngroups; g += (long long)gridDim.x * KWARPS) {
    const long long p0 = g * GR_ROWS;
    // stage 32 rows (each lane: columns lane, lane+32 of every row -> coalesced 256 B segments)
    {
      const long long p = p0 + lane;
      double w = 0.0, yy = 0.0;
      if (p < n && (!mask || mask[p])) {
        w = var ? wscale / (var[p] + var_add) : wscale;
        yy = y ? y[p] : 0.0;
      }
      swy[lane][0] = w;
      swy[lane][1] = w * yy;
      sw += w;
      swyy += w * yy * yy;
    }
#pragma unroll 4
    for (int i = 0; i < GR_ROWS; ++i) {
      const long long p = p0 + i;
      double v0 = 0.0, v1 = 0.0;
      if (p < n) {
        if (lane < M) v0 = __ldcs(rows + (size_t)p * M + lane);
        if (lane + 32 < M) v1 = __ldcs(rows + (size_t)p * M + lane + 32);
      }
      sX[i * KPITCH + lane] = v0;
      sX[i * KPITCH + lane + 32] = v1;
    }
    __syncwarp();
#pragma unroll 1
    for (int ks = 0; ks < GR_ROWS / 4; ++ks) {
      const int k0 = 4 * ks;
      // B fragment of tile column c: X[k0 + l4][8c + g4];  A fragment of tile row r: w * X[k0 + l4][8r + g4]
      double fb[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) fb[c] = sX[(k0 + l4) * KPITCH + 8 * c + g4];
      const double w = swy[k0 + l4][0];
      int t = 0;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const double fa = w * fb[r];
#pragma unroll
        for (int c = 0; c <= r; ++c, ++t) dmma884(acc[t][0], acc[t][1], fa, fb[c]);
      }
    }
    // h: lane owns two columns; 32 rows
#pragma unroll 8
    for (int i = 0; i < GR_ROWS; ++i) {
      const double wyv = swy[i][1];
      hacc[0] += wyv * sX[i * KPITCH + lane];
      hacc[1] += wyv * sX[i * KPITCH + lane + 32];
    }
    __syncwarp();
  }
  // combine the 8 warps of the CTA in shared memory (the staging slabs are free now), then one global atomic
  // per non-zero entry.  C fragment: rows 8r + g4, cols 8c + 2 l4 + {0,1}; only tiles r >= c were accumulated.
  __syncthreads();
  double* sG = &S.X[0][0];         // KM x KM
  double* sh = sG + KM * KM;       // KM
  for (int t = tid; t < KM * KM + KM; t += blockDim.x) sG[t] = 0.0;
  __syncthreads();
  {
    int t = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c <= r; ++c, ++t) {
        const int row = 8 * r + g4, col = 8 * c + 2 * l4;
        atomicAdd(&sG[row * KM + col], acc[t][0]);
        atomicAdd(&sG[row * KM + col + 1], acc[t][1]);
      }
  }
  atomicAdd(&sh[lane], hacc[0]);
  atomicAdd(&sh[lane + 32], hacc[1]);
  __syncthreads();
  for (int t = tid; t < KM * KM; t += blockDim.x) {
    const int row = t / KM, col = t % KM;
    if (row < M && col <= row) {
      const double v = sG[t];
      if (v != 0.0) {
        atomicAdd(&G[row * M + col], v);
        if (col != row) atomicAdd(&G[col * M + row], v);
      }
    }
  }
  if (tid < M && sh[tid] != 0.0) atomicAdd(&hvec[tid], sh[tid]);
  if (stats) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sw += __shfl_xor_sync(0xffffffffu, sw, o);
      swyy += __shfl_xor_sync(0xffffffffu, swyy, o);
    }
    if (lane == 0) {
      atomicAdd(&stats[0], sw);
      atomicAdd(&stats[1], swyy);
    }
  }
}

// r_n = k_n . x - y_n for valid rows (logz_residuals, distill_depth.py:80), plus count / sum / sum of squares
// (for torch.std, corr.py:218).  Warp per 8 rows, same streaming shape as the BA predictor kernel.
__global__ void __launch_bounds__(256)
rows_residual_kernel(const double* __restrict__ rows, const double* __restrict__ x, const double* __restrict__ y,
                     const unsigned char* __restrict__ mask, long long n, int M, double* __restrict__ res,
                     double* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const double x0 = lane < M ? x[lane] : 0.0, x1 = lane + 32 < M ? x[lane + 32] : 0.0;
  double cnt = 0.0, s1 = 0.0;
  for (long long p = warp; p < n; p += nwarps) {
    double v = 0.0;
    if (lane < M) v = __ldcs(rows + (size_t)p * M + lane) * x0;
    if (lane + 32 < M) v += __ldcs(rows + (size_t)p * M + lane + 32) * x1;
    v = warp_sum(v);
    if (lane == 0) {
      const bool ok = !mask || mask[p];
      const double r = ok ? v - y[p] : 0.0;
      if (res) res[p] = r;
      if (ok) {
        cnt += 1.0;
        s1 += r;
      }
    }
  }
  if (stats && lane == 0 && cnt > 0.0) {
    atomicAdd(&stats[0], cnt);
    atomicAdd(&stats[1], s1);
  }
}

// second pass of the two-pass unbiased variance: sum (r - mean)^2 over valid rows
__global__ void __launch_bounds__(256)
centered_sumsq_kernel(const double* __restrict__ res, const unsigned char* __restrict__ mask, long long n,
                      const double* __restrict__ stats, double* __restrict__ out) {
  const double mean = stats[1] / stats[0];
  double s = 0.0;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x)
    if (!mask || mask[p]) {
      const double d = res[p] - mean;
      s += d * d;
    }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(out, s);
}

}  // namespace como

using namespace como;

extern "C" int como_b200_kmat_kmm(const double* cov_img, int32_t B, int32_t H, int32_t W, const double* coords_m,
                                  int32_t M, double scale, double jitter, double* E_m, double* K_mm, void* stream) {
  COMO_REQUIRE(cov_img && coords_m && E_m && K_mm, "kmat_kmm: null pointer argument");
  COMO_REQUIRE(B >= 1 && H >= 1 && W >= 1 && M >= 1 && M <= 1024, "kmat_kmm: bad shape");
  kmm_kernel<<<B, 256, (size_t)M * 7 * sizeof(double), (cudaStream_t)stream>>>(cov_img, H, W, coords_m, M, scale, jitter, E_m,
                                                                               K_mm);
  return check_launch("kmat_kmm");
}

static int kmat_rows_launch(const double* cov_img, int B, int H, int W, const double* coords_m, const double* E_m,
                            const double* Kmm_inv, int M, double scale, const double* coords_n, const unsigned char* mask_n,
                            long long n, double* out, double* var_out, double* var_min, cudaStream_t st) {
  // the attribute is per device: set on every call (cheap) rather than once per process
  cudaFuncSetAttribute(kmat_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KmatSmem));
  cudaFuncSetAttribute(kmat_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KmatSmem));
  const long long groups = (n + KW_ROWS - 1) / KW_ROWS;
  long long per = (2LL * sm_count() + B - 1) / B;
  const long long need = (groups + KWARPS - 1) / KWARPS;
  if (per > need) per = need;
  if (per < 1) per = 1;
  if (var_min) {
    const double big = 1e300;
    for (int b = 0; b < B; ++b) cudaMemcpyAsync(var_min + b, &big, sizeof(double), cudaMemcpyHostToDevice, st);
  }
  if (coords_n)
    kmat_rows_kernel<false><<<dim3((unsigned)per, B), 32 * KWARPS, sizeof(KmatSmem), st>>>(
        cov_img, H, W, coords_m, E_m, Kmm_inv, M, scale, coords_n, mask_n, n, out, var_out, var_min);
  else
    kmat_rows_kernel<true><<<dim3((unsigned)per, B), 32 * KWARPS, sizeof(KmatSmem), st>>>(
        cov_img, H, W, coords_m, E_m, Kmm_inv, M, scale, nullptr, nullptr, n, out, var_out, var_min);
  return check_launch("kmat_rows");
}

extern "C" int como_b200_kmat_predictor(const double* cov_img, int32_t B, int32_t H, int32_t W, const double* coords_m,
                                        const double* E_m, const double* Kmm_inv, int32_t M, double scale,
                                        double* Knm_Kmminv, void* stream) {
  COMO_REQUIRE(cov_img && coords_m && E_m && Kmm_inv && Knm_Kmminv, "kmat_predictor: null pointer argument");
  COMO_REQUIRE(B >= 1 && H >= 1 && W >= 1 && M >= 1 && M <= BA_MAXM, "kmat_predictor: bad shape (M <= 64)");
  return kmat_rows_launch(cov_img, B, H, W, coords_m, E_m, Kmm_inv, M, scale, nullptr, nullptr, (long long)H * W,
                          Knm_Kmminv, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int como_b200_kmat_rows(const double* cov_img, int32_t B, int32_t H, int32_t W, const double* coords_m,
                                   const double* E_m, const double* Kmm_inv, int32_t M, double scale,
                                   const double* coords_n, const uint8_t* mask_n, int64_t n, double* rows, double* var_n,
                                   double* var_min, void* stream) {
  COMO_REQUIRE(B >= 1 && H >= 1 && W >= 1 && M >= 1 && M <= BA_MAXM && n >= 0, "kmat_rows: bad shape (M <= 64)");
  if (n == 0) return COMO_B200_OK;   // empty test set: nothing to do (pointers of empty tensors may be null)
  COMO_REQUIRE(cov_img && coords_m && E_m && Kmm_inv && coords_n && rows, "kmat_rows: null pointer argument");
  COMO_REQUIRE(!var_min || var_n, "kmat_rows: var_min needs var_n");
  return kmat_rows_launch(cov_img, B, H, W, coords_m, E_m, Kmm_inv, M, scale, coords_n, mask_n, n, rows, var_n, var_min,
                          (cudaStream_t)stream);
}

extern "C" int como_b200_weighted_gram(const double* rows, const double* y, const double* var, const uint8_t* mask,
                                       int64_t n, int32_t M, double var_add, double wscale, double* G, double* h,
                                       double* stats, void* stream) {
  COMO_REQUIRE(rows && G && h, "weighted_gram: null pointer argument");
  COMO_REQUIRE(M >= 1 && M <= BA_MAXM && n >= 0, "weighted_gram: bad shape (M <= 64)");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(G, 0, sizeof(double) * M * M, st);
  cudaMemsetAsync(h, 0, sizeof(double) * M, st);
  if (stats) cudaMemsetAsync(stats, 0, sizeof(double) * 2, st);
  if (n == 0) return COMO_B200_OK;
  // the attribute is per device: set on every call (cheap) rather than once per process
  cudaFuncSetAttribute(weighted_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GramSmem));
  const long long groups = (n + GR_ROWS - 1) / GR_ROWS;
  long long grid = sm_count();
  const long long need = (groups + KWARPS - 1) / KWARPS;
  if (grid > need) grid = need;
  weighted_gram_kernel<<<(unsigned)grid, 32 * KWARPS, sizeof(GramSmem), st>>>(rows, y, var, mask, n, M, var_add, wscale, G, h,
                                                                            stats);
  return check_launch("weighted_gram");
}

extern "C" int como_b200_rows_residual(const double* rows, const double* x, const double* y, const uint8_t* mask, int64_t n,
                                       int32_t M, double* res, double* stats3, void* stream) {
  COMO_REQUIRE(rows && x && y && res && stats3, "rows_residual: null pointer argument");
  COMO_REQUIRE(M >= 1 && M <= BA_MAXM && n >= 0, "rows_residual: bad shape (M <= 64)");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(stats3, 0, sizeof(double) * 3, st);
  if (n == 0) return COMO_B200_OK;
  long long blocks = (n + 7) / 8;
  if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
  rows_residual_kernel<<<(unsigned)blocks, 256, 0, st>>>(rows, x, y, mask, n, M, res, stats3);
  long long b2 = (n + 255) / 256;
  if (b2 > 4LL * sm_count()) b2 = 4LL * sm_count();
  centered_sumsq_kernel<<<(unsigned)b2, 256, 0, st>>>(res, mask, n, stats3, stats3 + 2);
  return check_launch("rows_residual");
}
