// Two-frame SfM bootstrap: como/odom/frontend/two_frame_sfm.py:180-392.
//
// One Gauss-Newton iteration of the (6 + M)-dimensional system [relative pose | sparse log depths]:
//   sfm_linearize : per reference pixel  logz_n = k_n . d  (DMMA, 32 rows at a time, A fragments straight from HBM),
//                   back-project, transform, project, bilinear [I, gx, gy] (zero padding, the reference's float32
//                   A_norm quirk), residual, pose Jacobian J_T (6) and the depth coefficient beta_n = dI/dP_i . P_i
//                   -- the reference's (N, M) matrix dI/dd is beta_n k_n (rank one), never formed;
//   (exact median of |r| over valid pixels: select.cu)
//   sfm_accumulate: Huber weights on the fly, G = sum s beta^2 k k^T (36 lower tiles on the FP64 tensor path, pixel
//                   index as K), the 7 stack rows sum s beta J_T k^T / sum s beta r k^T, and the small pose sums.
// Priors (M x M), the 70 x 70 solve and the pose update are host-side tensor plumbing on <= 70 numbers per row.
#include "ba_common.cuh"

namespace como {

constexpr int SF_THREADS = 256;
constexpr int SREC = 8;            // per pixel: r (NaN when invalid), J_T[6], beta
constexpr double SF_HUBER = 1.345;
constexpr int SPITCH = 68;

struct SfmParams {
  double T[12];                    // row-major 3x4 of T_ji
  double fx, fy, cx, cy;
  double Ax, Ay;                   // float32(1/W), float32(1/H) widened: two_frame_sfm.py:186-189
  int H, W, M;
  long long N;
};

__device__ __forceinline__ void dmma884s(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(SF_THREADS)
sfm_linearize_kernel(const double* __restrict__ Knm, const double* __restrict__ d, const long long* __restrict__ coords,
                     const double* __restrict__ vals, const double* __restrict__ img, SfmParams q,
                     double* __restrict__ rec, double* __restrict__ absr, double* __restrict__ proj,
                     double* __restrict__ stats) {
  __shared__ double s_logz[SF_THREADS / 32][32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g4 = lane >> 2, l4 = lane & 3;
  double bfrag[16];
#pragma unroll
  for (int ks = 0; ks < 16; ++ks) bfrag[ks] = (g4 == 0 && 4 * ks + l4 < q.M) ? d[4 * ks + l4] : 0.0;
  double sum_logz = 0.0, cnt_valid = 0.0;
  const size_t HW = (size_t)q.H * q.W;
  const long long ngroups = (q.N + 31) / 32;
  for (long long g = (long long)blockIdx.x * (SF_THREADS / 32) + wid; g < ngroups; g += (long long)gridDim.x * (SF_THREADS / 32)) {
    const long long n0 = g * 32;
#pragma unroll
    for (int rt = 0; rt < 4; ++rt) {
      const long long n = n0 + 8 * rt + g4;
      const bool ok = n < q.N;
      const double* row = Knm + (size_t)(ok ? n : 0) * q.M;
      double a[16];
#pragma unroll
      for (int ks = 0; ks < 16; ++ks) a[ks] = (ok && 4 * ks + l4 < q.M) ? __ldg(row + 4 * ks + l4) : 0.0;
      double c0 = 0.0, c1 = 0.0;
#pragma unroll
      for (int ks = 0; ks < 16; ++ks) dmma884s(c0, c1, a[ks], bfrag[ks]);
      if (l4 == 0) s_logz[wid][8 * rt + g4] = c0;
    }
    __syncwarp();
    const long long n = n0 + lane;
    if (n < q.N) {
      const double logz = s_logz[wid][lane];
      sum_logz += logz;
      const double z = exp(logz);
      const double x = (double)coords[2 * n + 1], y = (double)coords[2 * n];
      const double Pi[3] = {(x - q.cx) / q.fx * z, (y - q.cy) / q.fy * z, z};
      const double Pj[3] = {q.T[0] * Pi[0] + q.T[1] * Pi[1] + q.T[2] * Pi[2] + q.T[3],
                            q.T[4] * Pi[0] + q.T[5] * Pi[1] + q.T[6] * Pi[2] + q.T[7],
                            q.T[8] * Pi[0] + q.T[9] * Pi[1] + q.T[10] * Pi[2] + q.T[11]};
      const double u = q.fx * Pj[0] / Pj[2] + q.cx, v = q.fy * Pj[1] / Pj[2] + q.cy;
      const bool valid = (u >= 1.0) && (u < (double)(q.W - 1)) && (v >= 1.0) && (v < (double)(q.H - 1)) && (Pj[2] > 0.0);
      // grid_sample position: ((2 A u + A - 1) + 1) W / 2 - 0.5
      const double us = ((2.0 * q.Ax * u + q.Ax - 1.0) + 1.0) * (double)q.W / 2.0 - 0.5;
      const double vs = ((2.0 * q.Ay * v + q.Ay - 1.0) + 1.0) * (double)q.H / 2.0 - 0.5;
      double s[3] = {0.0, 0.0, 0.0};
      if (us == us && vs == vs && fabs(us) < 1e9 && fabs(vs) < 1e9) {
        const double x0f = floor(us), y0f = floor(vs);
        const int x0 = (int)x0f, y0 = (int)y0f;
        const double fx1 = us - x0f, fy1 = vs - y0f;
#pragma unroll
        for (int ty = 0; ty < 2; ++ty)
#pragma unroll
          for (int tx = 0; tx < 2; ++tx) {
            const int xx = x0 + tx, yy = y0 + ty;
            if (xx < 0 || xx >= q.W || yy < 0 || yy >= q.H) continue;
            const double wgt = (tx ? fx1 : 1.0 - fx1) * (ty ? fy1 : 1.0 - fy1);
            const size_t o = (size_t)yy * q.W + xx;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) s[ch] += wgt * __ldg(img + ch * HW + o);
          }
      }
      const double r = s[0] - vals[n];
      const double iz = 1.0 / Pj[2];
      const double dP[3] = {s[1] * q.fx * iz, s[2] * q.fy * iz, -(s[1] * q.fx * Pj[0] + s[2] * q.fy * Pj[1]) * iz * iz};
      // dI/dPi = dI/dPj R  (row vector)
      const double a3[3] = {dP[0] * q.T[0] + dP[1] * q.T[4] + dP[2] * q.T[8], dP[0] * q.T[1] + dP[1] * q.T[5] + dP[2] * q.T[9],
                            dP[0] * q.T[2] + dP[1] * q.T[6] + dP[2] * q.T[10]};
      double* o = rec + (size_t)n * SREC;
      o[0] = valid ? r : __longlong_as_double(0x7ff8000000000000LL);
      o[1] = Pi[1] * a3[2] - Pi[2] * a3[1];     // (Pi x a)^T = -a^T [Pi]x
      o[2] = Pi[2] * a3[0] - Pi[0] * a3[2];
      o[3] = Pi[0] * a3[1] - Pi[1] * a3[0];
      o[4] = a3[0];
      o[5] = a3[1];
      o[6] = a3[2];
      o[7] = a3[0] * Pi[0] + a3[1] * Pi[1] + a3[2] * Pi[2];
      absr[n] = o[0];
      proj[3 * n] = u;
      proj[3 * n + 1] = v;
      proj[3 * n + 2] = Pj[2];
      if (valid) cnt_valid += 1.0;
    }
    __syncwarp();
  }
  sum_logz = warp_sum(sum_logz);
  cnt_valid = warp_sum(cnt_valid);
  if (lane == 0) {
    atomicAdd(&stats[0], sum_logz);
    atomicAdd(&stats[1], cnt_valid);
  }
}

struct SfmGramSmem {
  double X[8][32 * SPITCH];
  double cf[8][32][8];       // per row: s beta^2, s beta J_T[0..5], s beta r
};

// G (M x M) += sum_n s_n beta_n^2 k_n k_n^T;  St (7 x M): rows 0..5 += s beta J_T[a] k^T, row 6 += s beta r k^T,
// with s_n = huber(r / sigma) / sigma^2 (0 for invalid pixels).
__global__ void __launch_bounds__(256, 1)
sfm_gram_kernel(const double* __restrict__ Knm, const double* __restrict__ rec, const double* __restrict__ sigma_p,
                long long N, int M, double* __restrict__ G, double* __restrict__ St) {
  extern __shared__ __align__(16) unsigned char sraw[];
  SfmGramSmem& S = *reinterpret_cast<SfmGramSmem*>(sraw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g4 = lane >> 2, l4 = lane & 3;
  double* sX = S.X[warp];
  double(*cf)[8] = S.cf[warp];
  const double sigma = *sigma_p;
  const double isig = 1.0 / sigma;
  double acc[36][2];
#pragma unroll
  for (int t = 0; t < 36; ++t) acc[t][0] = acc[t][1] = 0.0;
  double st[7][2];
#pragma unroll
  for (int a = 0; a < 7; ++a) st[a][0] = st[a][1] = 0.0;
  const long long ngroups = (N + 31) / 32;
  for (long long g = (long long)blockIdx.x * 8 + warp; g < ngroups; g += (long long)gridDim.x * 8) {
    const long long p0 = g * 32;
    {
      const long long p = p0 + lane;
      double c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (p < N) {
        const double* o = rec + (size_t)p * SREC;
        const double r = o[0];
        if (r == r) {
          const double wr = fabs(r) * isig;
          const double w = (wr < SF_HUBER) ? 1.0 : SF_HUBER / wr;
          const double s = w * isig * isig;
          const double sb = s * o[7];
          c[0] = sb * o[7];
#pragma unroll
          for (int a = 0; a < 6; ++a) c[1 + a] = sb * o[1 + a];
          c[7] = sb * r;
        }
      }
#pragma unroll
      for (int a = 0; a < 8; ++a) cf[lane][a] = c[a];
    }
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
      const long long p = p0 + i;
      double v0 = 0.0, v1 = 0.0;
      if (p < N) {
        if (lane < M) v0 = __ldcs(Knm + (size_t)p * M + lane);
        if (lane + 32 < M) v1 = __ldcs(Knm + (size_t)p * M + lane + 32);
      }
      sX[i * SPITCH + lane] = v0;
      sX[i * SPITCH + lane + 32] = v1;
    }
    __syncwarp();
#pragma unroll 1
    for (int ks = 0; ks < 8; ++ks) {
      const int k0 = 4 * ks;
      double fb[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) fb[c] = sX[(k0 + l4) * SPITCH + 8 * c + g4];
      const double w = cf[k0 + l4][0];
      int t = 0;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const double fa = w * fb[r];
#pragma unroll
        for (int c = 0; c <= r; ++c, ++t) dmma884s(acc[t][0], acc[t][1], fa, fb[c]);
      }
    }
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
      const double x0 = sX[i * SPITCH + lane], x1 = sX[i * SPITCH + lane + 32];
#pragma unroll
      for (int a = 0; a < 7; ++a) {
        const double c = cf[i][1 + a];
        st[a][0] += c * x0;
        st[a][1] += c * x1;
      }
    }
    __syncwarp();
  }
  __syncthreads();
  double* sG = &S.X[0][0];          // 64 x 64 + 7 x 64
  double* sS = sG + 64 * 64;
  for (int t = tid; t < 64 * 64 + 7 * 64; t += blockDim.x) sG[t] = 0.0;
  __syncthreads();
  {
    int t = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c <= r; ++c, ++t) {
        const int row = 8 * r + g4, col = 8 * c + 2 * l4;
        atomicAdd(&sG[row * 64 + col], acc[t][0]);
        atomicAdd(&sG[row * 64 + col + 1], acc[t][1]);
      }
  }
#pragma unroll
  for (int a = 0; a < 7; ++a) {
    atomicAdd(&sS[a * 64 + lane], st[a][0]);
    atomicAdd(&sS[a * 64 + lane + 32], st[a][1]);
  }
  __syncthreads();
  for (int t = tid; t < 64 * 64; t += blockDim.x) {
    const int row = t / 64, col = t % 64;
    if (row < M && col <= row) {
      const double v = sG[t];
      if (v != 0.0) {
        atomicAdd(&G[row * M + col], v);
        if (col != row) atomicAdd(&G[col * M + row], v);
      }
    }
  }
  for (int t = tid; t < 7 * 64; t += blockDim.x) {
    const int a = t / 64, col = t % 64;
    if (col < M && sS[t] != 0.0) atomicAdd(&St[a * M + col], sS[t]);
  }
}

// small[0..20] = upper triangle of sum s J_T J_T^T (row-major a <= b), small[21..26] = sum s J_T r, small[27] = sum w wr^2
__global__ void __launch_bounds__(256)
sfm_small_kernel(const double* __restrict__ rec, const double* __restrict__ sigma_p, long long N, double* __restrict__ small) {
  const double sigma = *sigma_p, isig = 1.0 / sigma;
  double a28[28];
#pragma unroll
  for (int q = 0; q < 28; ++q) a28[q] = 0.0;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (long long)gridDim.x * blockDim.x) {
    const double* o = rec + (size_t)p * SREC;
    const double r = o[0];
    if (r == r) {
      const double wr = fabs(r) * isig;
      const double w = (wr < SF_HUBER) ? 1.0 : SF_HUBER / wr;
      const double s = w * isig * isig;
      int t = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = a; b < 6; ++b, ++t) a28[t] += s * o[1 + a] * o[1 + b];
#pragma unroll
      for (int a = 0; a < 6; ++a) a28[21 + a] += s * o[1 + a] * r;
      a28[27] += w * wr * wr;
    }
  }
#pragma unroll
  for (int q = 0; q < 28; ++q) {
    const double v = warp_sum(a28[q]);
    if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(&small[q], v);
  }
}

}  // namespace como

using namespace como;

extern "C" int como_b200_sfm_linearize(const double* Knm, const double* d, const int64_t* coords, const double* vals,
                                       const double* img_and_grads, int32_t H, int32_t W, int64_t N, int32_t M,
                                       const double* T_ji12, const double* intr4, double* rec, double* absr, double* proj,
                                       double* stats2, void* stream) {
  COMO_REQUIRE(Knm && d && coords && vals && img_and_grads && T_ji12 && intr4 && rec && absr && proj && stats2,
               "sfm_linearize: null pointer argument");
  COMO_REQUIRE(H >= 3 && W >= 3 && N >= 1 && M >= 1 && M <= BA_MAXM, "sfm_linearize: bad shape (M <= 64)");
  SfmParams q;
  for (int i = 0; i < 12; ++i) q.T[i] = T_ji12[i];
  q.fx = intr4[0]; q.fy = intr4[1]; q.cx = intr4[2]; q.cy = intr4[3];
  q.Ax = (double)(1.0f / (float)W);
  q.Ay = (double)(1.0f / (float)H);
  q.H = H; q.W = W; q.M = M; q.N = N;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(stats2, 0, sizeof(double) * 2, st);
  long long blocks = ((N + 31) / 32 + 7) / 8;
  if (blocks > 4LL * sm_count()) blocks = 4LL * sm_count();
  sfm_linearize_kernel<<<(unsigned)blocks, SF_THREADS, 0, st>>>(Knm, d, (const long long*)coords, vals, img_and_grads, q, rec, absr,
                                                                proj, stats2);
  return check_launch("sfm_linearize");
}

extern "C" int como_b200_sfm_accumulate(const double* Knm, const double* rec, const double* sigma, int64_t N, int32_t M,
                                        double* G, double* St7, double* small28, void* stream) {
  COMO_REQUIRE(Knm && rec && sigma && G && St7 && small28, "sfm_accumulate: null pointer argument");
  COMO_REQUIRE(N >= 1 && M >= 1 && M <= BA_MAXM, "sfm_accumulate: bad shape (M <= 64)");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(G, 0, sizeof(double) * M * M, st);
  cudaMemsetAsync(St7, 0, sizeof(double) * 7 * M, st);
  cudaMemsetAsync(small28, 0, sizeof(double) * 28, st);
  // the attribute is per device: set on every call (cheap) rather than once per process
  cudaFuncSetAttribute(sfm_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SfmGramSmem));
  long long grid = sm_count();
  const long long need = ((N + 31) / 32 + 7) / 8;
  if (grid > need) grid = need;
  sfm_gram_kernel<<<(unsigned)grid, 256, sizeof(SfmGramSmem), st>>>(Knm, rec, sigma, N, M, G, St7);
  long long b2 = (N + 255) / 256;
  if (b2 > 2LL * sm_count()) b2 = 2LL * sm_count();
  sfm_small_kernel<<<(unsigned)b2, 256, 0, st>>>(rec, sigma, N, small28);
  return check_launch("sfm_accumulate");
}
