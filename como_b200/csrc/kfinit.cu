// Keyframe-creation helpers (como/odom/frontend/corr.py:60-242).
//   reproject_dense: every pixel of the last keyframe's depth image is back-projected, moved by T_ji and projected
//     into the new frame (corr.py:37-43, 80-96); outputs the fractional (row, col), log depth and the
//     filter_reproj_coords mask (corr.py:17-29) in pixel order -- no compaction, downstream kernels take the mask.
//   sample_depth_gradmag: bilinear zero-padded lookups (F.grid_sample, align_corners=False) of the depth image
//     and of |Scharr(log depth)| (utils/image_processing.py:8-45, reflect padding) at a handful of points; the
//     gradient magnitude is evaluated only at the 4 taps of each point, never as an image.
#include "common.cuh"

namespace como {

struct ReprojParams {
  double T[12];  // row-major 3x4 of T_ji
  double fx, fy, cx, cy;
  double min_depth;
  int H, W;
};

__global__ void __launch_bounds__(256)
reproject_dense_kernel(const double* __restrict__ z_img, ReprojParams q, double* __restrict__ coords_j,
                       double* __restrict__ logz_j, double* __restrict__ z_j, unsigned char* __restrict__ mask) {
  const long long n = (long long)q.H * q.W;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(p / q.W), c = (int)(p % q.W);
    const double z = z_img[p];
    const double X = ((double)c - q.cx) / q.fx * z, Y = ((double)r - q.cy) / q.fy * z, Z = z;
    const double Xj = q.T[0] * X + q.T[1] * Y + q.T[2] * Z + q.T[3];
    const double Yj = q.T[4] * X + q.T[5] * Y + q.T[6] * Z + q.T[7];
    const double Zj = q.T[8] * X + q.T[9] * Y + q.T[10] * Z + q.T[11];
    const double u = q.fx * Xj / Zj + q.cx, v = q.fy * Yj / Zj + q.cy;
    const bool ok = (u >= 1.0) && (u < (double)(q.W - 1)) && (v >= 1.0) && (v < (double)(q.H - 1)) && (Zj > q.min_depth);
    coords_j[2 * p] = v;
    coords_j[2 * p + 1] = u;
    z_j[p] = Zj;
    logz_j[p] = ok ? log(Zj) : 0.0;
    mask[p] = ok ? 1 : 0;
  }
}

__device__ __forceinline__ int reflect_idx(int i, int n) {  // F.pad(mode="reflect") by one pixel
  if (i < 0) return -i;
  if (i >= n) return 2 * n - 2 - i;
  return i;
}

__device__ double scharr_mag_logz(const double* __restrict__ z, int H, int W, int y, int x) {
  double l[3][3];
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) l[dy + 1][dx + 1] = log(z[(size_t)reflect_idx(y + dy, H) * W + reflect_idx(x + dx, W)]);
  const double k3 = 3.0 / 32.0, k10 = 10.0 / 32.0;
  const double gx = -k3 * l[0][0] + k3 * l[0][2] - k10 * l[1][0] + k10 * l[1][2] - k3 * l[2][0] + k3 * l[2][2];
  const double gy = -k3 * l[0][0] - k10 * l[0][1] - k3 * l[0][2] + k3 * l[2][0] + k10 * l[2][1] + k3 * l[2][2];
  return sqrt(gx * gx + gy * gy);
}

// one thread per point; coords (n,2) [row, col] pixel units
__global__ void sample_depth_gradmag_kernel(const double* __restrict__ z_img, int H, int W, const double* __restrict__ coords_z,
                                            const double* __restrict__ coords_g, int n, double* __restrict__ z_out,
                                            double* __restrict__ g_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int which = 0; which < 2; ++which) {
    const double* cc = which == 0 ? coords_z : coords_g;
    double* out = which == 0 ? z_out : g_out;
    if (!cc || !out) continue;
    // grid_sample's normalise / unnormalise round trip in double: ((2A x + A - 1) + 1) * dim / 2 - 0.5
    const double Ay = 1.0 / (double)H, Ax = 1.0 / (double)W;
    const double yn = 2.0 * Ay * cc[2 * i] + Ay - 1.0, xn = 2.0 * Ax * cc[2 * i + 1] + Ax - 1.0;
    const double y = ((yn + 1.0) * (double)H - 1.0) / 2.0, x = ((xn + 1.0) * (double)W - 1.0) / 2.0;
    const double y0f = floor(y), x0f = floor(x);
    const int y0 = (int)y0f, x0 = (int)x0f;
    const double fy = y - y0f, fx = x - x0f;
    double acc = 0.0;
#pragma unroll
    for (int ty = 0; ty < 2; ++ty)
#pragma unroll
      for (int tx = 0; tx < 2; ++tx) {
        const int yy = y0 + ty, xx = x0 + tx;
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
        const double wgt = (ty ? fy : 1.0 - fy) * (tx ? fx : 1.0 - fx);
        const double v = which == 0 ? z_img[(size_t)yy * W + xx] : scharr_mag_logz(z_img, H, W, yy, xx);
        acc += wgt * v;
      }
    out[i] = acc;
  }
}

}  // namespace como

using namespace como;

extern "C" int como_b200_reproject_dense(const double* z_img, int32_t H, int32_t W, const double* T_ji12, const double* intr4,
                                         double min_depth, double* coords_j, double* logz_j, double* z_j, uint8_t* mask,
                                         void* stream) {
  COMO_REQUIRE(z_img && T_ji12 && intr4 && coords_j && logz_j && z_j && mask, "reproject_dense: null pointer argument");
  COMO_REQUIRE(H >= 3 && W >= 3, "reproject_dense: bad shape");
  ReprojParams q;
  for (int i = 0; i < 12; ++i) q.T[i] = T_ji12[i];
  q.fx = intr4[0]; q.fy = intr4[1]; q.cx = intr4[2]; q.cy = intr4[3];
  q.min_depth = min_depth;
  q.H = H; q.W = W;
  const long long n = (long long)H * W;
  long long blocks = (n + 255) / 256;
  if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
  reproject_dense_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(z_img, q, coords_j, logz_j, z_j, mask);
  return check_launch("reproject_dense");
}

extern "C" int como_b200_sample_depth_gradmag(const double* z_img, int32_t H, int32_t W, const double* coords_z,
                                              const double* coords_g, int32_t n, double* z_out, double* g_out, void* stream) {
  COMO_REQUIRE(z_img, "sample_depth_gradmag: null pointer argument");
  COMO_REQUIRE(H >= 2 && W >= 2 && n >= 0, "sample_depth_gradmag: bad shape");
  if (n == 0) return COMO_B200_OK;
  sample_depth_gradmag_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(z_img, H, W, coords_z, coords_g, n, z_out, g_out);
  return check_launch("sample_depth_gradmag");
}
