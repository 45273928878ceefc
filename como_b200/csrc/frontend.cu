// Tracker front-end (fp32): image pyramid + gradients, keyframe reference preparation, keyframe decisions.
//   rgb -> gray (torchvision weights), 3x3 [1 2 1]^2/16 blur + stride 2, Scharr/32 with reflect padding
//       como/utils/image_processing.py:8-87, como/odom/Tracking.py:88-102
//   keyframe reference: nearest depth pyramid, back-projection, transform into the last keyframe, border/depth
//       mask, inverse-compositional Jacobians   como/odom/Tracking.py:243-314, photo_tracking.py:46-74
//   reprojection statistics for the keyframe / one-way decisions   como/odom/Tracking.py:169-188,342-345
#include "common.cuh"

namespace como {

__device__ __forceinline__ int reflect_idx(int i, int n) {  // torch 'reflect' padding by one pixel
  if (i < 0) return -i;
  if (i >= n) return 2 * n - 2 - i;
  return i;
}

__global__ void gray_kernel(const float* __restrict__ rgb, long long hw, float* __restrict__ gray) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hw) return;
  // torchvision rgb_to_grayscale: (0.2989 r + 0.587 g + 0.114 b), left-to-right, no contraction
  const float a = __fmul_rn(0.2989f, rgb[i]);
  const float b = __fmul_rn(0.587f, rgb[hw + i]);
  const float c = __fmul_rn(0.114f, rgb[2 * hw + i]);
  gray[i] = __fadd_rn(__fadd_rn(a, b), c);
}

// out (h/2... ceil) = blur3x3(in)[::2, ::2]
__global__ void blur_down_kernel(const float* __restrict__ in, int h, int w, float* __restrict__ out, int ho, int wo) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= wo || y >= ho) return;
  const int cx = 2 * x, cy = 2 * y;
  float acc = 0.0f;
  const float wgt[3] = {1.0f / 16.0f, 2.0f / 16.0f, 1.0f / 16.0f};
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = reflect_idx(cy + dy, h);
    for (int dx = -1; dx <= 1; ++dx) {
      const int xx = reflect_idx(cx + dx, w);
      acc += (wgt[dy + 1] * wgt[dx + 1] * 16.0f) * in[(size_t)yy * w + xx];
    }
  }
  out[(size_t)y * wo + x] = acc;
}

__global__ void scharr_kernel(const float* __restrict__ in, int h, int w, float* __restrict__ gx, float* __restrict__ gy) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  float v[3][3];
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) v[dy + 1][dx + 1] = in[(size_t)reflect_idx(y + dy, h) * w + reflect_idx(x + dx, w)];
  const float k3 = 3.0f / 32.0f, k10 = 10.0f / 32.0f;
  gx[(size_t)y * w + x] = k3 * (v[0][2] - v[0][0]) + k10 * (v[1][2] - v[1][0]) + k3 * (v[2][2] - v[2][0]);
  gy[(size_t)y * w + x] = k3 * (v[2][0] - v[0][0]) + k10 * (v[2][1] - v[0][1]) + k3 * (v[2][2] - v[0][2]);
}


// ------------------------------------------------------------------------------------------------------------
// Fused front-end: RGB -> gray -> the whole [1 2 1]^2/16 stride-2 pyramid (-> optional Scharr/32
// gradients of every level) in ONE launch.  A CTA owns a 64 x 64 tile of the finest level and the matching
// 32 / 16 / 8 tiles below it; every level is built in shared memory from the level above, with a halo that doubles
// per level (1 at the coarsest, 3, 7, 15 at the finest for 4 levels), so no level is ever re-read from HBM.  Region
// entries one pixel outside an image hold torch's 'reflect' value (they are evaluated at the reflected coordinate,
// which is always inside the same region); entries further out are never read.
// Same arithmetic as gray_kernel / blur_down_kernel / scharr_kernel above (and the same parity tests).
// ------------------------------------------------------------------------------------------------------------
constexpr int PYR_TILE = 64;
constexpr int PYR_MAX_LEVELS = 4;

struct PyrArgs {
  float* img[PYR_MAX_LEVELS];   // finest first: img[0] is H x W
  float* gx[PYR_MAX_LEVELS];    // optional (nullptr: no gradients)
  float* gy[PYR_MAX_LEVELS];
  int h[PYR_MAX_LEVELS], w[PYR_MAX_LEVELS];
  int halo[PYR_MAX_LEVELS];
  int off[PYR_MAX_LEVELS];      // float offset of the level's region in shared memory
  int num_levels;
};

__global__ void __launch_bounds__(256) pyramid_fused_kernel(const float* __restrict__ rgb, PyrArgs a) {
  extern __shared__ float pyr_smem[];
  const int tid = threadIdx.x;
  const long long hw0 = (long long)a.h[0] * a.w[0];
  // ---- level 0: gray of the tile + halo (reflected coordinates outside the image)
  {
    const int ts = PYR_TILE, hl = a.halo[0], rs = ts + 2 * hl;
    const int y0 = blockIdx.y * ts - hl, x0 = blockIdx.x * ts - hl;
    float* R = pyr_smem + a.off[0];
    for (int e = tid; e < rs * rs; e += 256) {
      const int ry = e / rs, rx = e - ry * rs;
      const int y = y0 + ry, x = x0 + rx;
      float v = 0.0f;
      if (y >= -1 && y <= a.h[0] && x >= -1 && x <= a.w[0]) {
        const long long o = (long long)reflect_idx(y, a.h[0]) * a.w[0] + reflect_idx(x, a.w[0]);
        const float c0 = __fmul_rn(0.2989f, rgb[o]);
        const float c1 = __fmul_rn(0.587f, rgb[hw0 + o]);
        const float c2 = __fmul_rn(0.114f, rgb[2 * hw0 + o]);
        v = __fadd_rn(__fadd_rn(c0, c1), c2);
      }
      R[e] = v;
    }
  }
  __syncthreads();
  const float wgt[3] = {1.0f / 16.0f, 2.0f / 16.0f, 1.0f / 16.0f};
  for (int l = 0; l < a.num_levels; ++l) {
    const int ts = PYR_TILE >> l, hl = a.halo[l], rs = ts + 2 * hl;
    const int ty0 = blockIdx.y * ts, tx0 = blockIdx.x * ts;
    const float* R = pyr_smem + a.off[l];
    // ---- write this level's tile (+ gradients)
    for (int e = tid; e < ts * ts; e += 256) {
      const int ry = e / ts, rx = e - ry * ts;
      const int y = ty0 + ry, x = tx0 + rx;
      if (y < a.h[l] && x < a.w[l]) {
        const float* c = R + (ry + hl) * rs + (rx + hl);
        const size_t o = (size_t)y * a.w[l] + x;
        a.img[l][o] = c[0];
        if (a.gx[l]) {
          const float k3 = 3.0f / 32.0f, k10 = 10.0f / 32.0f;
          a.gx[l][o] = k3 * (c[-rs + 1] - c[-rs - 1]) + k10 * (c[1] - c[-1]) + k3 * (c[rs + 1] - c[rs - 1]);
          a.gy[l][o] = k3 * (c[rs - 1] - c[-rs - 1]) + k10 * (c[rs] - c[-rs]) + k3 * (c[rs + 1] - c[-rs + 1]);
        }
      }
    }
    if (l + 1 == a.num_levels) break;
    // ---- next level's region from this one: blur 3x3 at the even positions
    {
      const int ts1 = ts >> 1, hl1 = a.halo[l + 1], rs1 = ts1 + 2 * hl1;
      const int y1_0 = blockIdx.y * ts1 - hl1, x1_0 = blockIdx.x * ts1 - hl1;
      const int y0_0 = ty0 - hl, x0_0 = tx0 - hl;   // coordinates of R[0][0]
      float* R1 = pyr_smem + a.off[l + 1];
      for (int e = tid; e < rs1 * rs1; e += 256) {
        const int ry = e / rs1, rx = e - ry * rs1;
        const int y1 = y1_0 + ry, x1 = x1_0 + rx;
        float acc = 0.0f;
        if (y1 >= -1 && y1 <= a.h[l + 1] && x1 >= -1 && x1 <= a.w[l + 1]) {
          const int cy = 2 * reflect_idx(y1, a.h[l + 1]) - y0_0, cx = 2 * reflect_idx(x1, a.w[l + 1]) - x0_0;
          for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx)
              acc += (wgt[dy + 1] * wgt[dx + 1] * 16.0f) * R[(cy + dy) * rs + (cx + dx)];
        }
        R1[e] = acc;
      }
    }
    __syncthreads();
  }
}

// Mapping.get_img_and_grads (como/odom/Mapping.py:368-376) in ONE pass: rgb (3,H,W) f64 -> [I, gx, gy] (3,H,W) f64,
// the layout the BA gather kernels read.  I = 0.2989 R + 0.587 G + 0.114 B (torchvision rgb_to_grayscale), Scharr
// with reflect padding (utils/image_processing.py:8-45); the 3x3 gray neighbourhood is rebuilt from RGB on the fly.
__global__ void img_and_grads_f64_kernel(const double* __restrict__ rgb, int h, int w, double* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const size_t hw = (size_t)h * w;
  double v[3][3];
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const size_t o = (size_t)reflect_idx(y + dy, h) * w + reflect_idx(x + dx, w);
      v[dy + 1][dx + 1] = 0.2989 * rgb[o] + 0.587 * rgb[hw + o] + 0.114 * rgb[2 * hw + o];
    }
  const double k3 = 3.0 / 32.0, k10 = 10.0 / 32.0;
  const size_t o = (size_t)y * w + x;
  out[o] = v[1][1];
  out[hw + o] = k3 * (v[0][2] - v[0][0]) + k10 * (v[1][2] - v[1][0]) + k3 * (v[2][2] - v[2][0]);
  out[2 * hw + o] = k3 * (v[2][0] - v[0][0]) + k10 * (v[2][1] - v[0][1]) + k3 * (v[2][2] - v[0][2]);
}

// One level of Tracking.update_kf_reference for one keyframe b: all pixels of the level image.
// rel (3x4 row-major) maps keyframe b's camera frame into the last keyframe's frame.
__global__ void kf_reference_kernel(const float* __restrict__ img, const float* __restrict__ gx, const float* __restrict__ gy,
                                    const float* __restrict__ depth_full, int Hf, int Wf, int sub, int h, int w, float fx,
                                    float fy, float cx, float cy, const float* __restrict__ rel, float border,
                                    float depth_thresh, float* __restrict__ vals, float* __restrict__ grads,
                                    float* __restrict__ P, float* __restrict__ J, uint8_t* __restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= h * w) return;
  const int y = i / w, x = i % w;
  const float I = img[i], dx = gx[i], dy = gy[i];
  const float z = depth_full[(size_t)(y * sub) * Wf + (size_t)x * sub];  // nearest-neighbour depth pyramid
  // backprojection (camera.py:44-54): ((p - c)/f) * z
  const float rx = ((float)x - cx) / fx, ry = ((float)y - cy) / fy;
  const float Pc[3] = {z * rx, z * ry, z * 1.0f};
  float Q[3];
  for (int r = 0; r < 3; ++r) Q[r] = rel[r * 4] * Pc[0] + rel[r * 4 + 1] * Pc[1] + rel[r * 4 + 2] * Pc[2] + rel[r * 4 + 3];
  const float X = Q[0], Y = Q[1], Z = Q[2];
  const float t1 = fx * X / Z, t2 = fy * Y / Z;
  const float px = t1 + cx, py = t2 + cy;
  const bool ok = (px >= -border) && (px <= (float)(w - 1) + border) && (py >= -border) && (py <= (float)(h - 1) + border) &&
                  (Z > depth_thresh);
  vals[i] = I;
  grads[2 * i] = dx;
  grads[2 * i + 1] = dy;
  P[3 * i] = X;
  P[3 * i + 1] = Y;
  P[3 * i + 2] = Z;
  mask[i] = ok ? 1 : 0;
  const float d00 = fx / Z, d02 = -t1 / Z, d11 = fy / Z, d12 = -t2 / Z;
  const float a0 = d02 * Y, a1 = d00 * Z - d02 * X, a2 = -d00 * Y;
  const float b0 = -d11 * Z + d12 * Y, b1 = -d12 * X, b2 = d11 * X;
  float4 o0, o1;
  o0.x = dx * a0 + dy * b0;
  o0.y = dx * a1 + dy * b1;
  o0.z = dx * a2 + dy * b2;
  o0.w = dx * d00;
  o1.x = dy * d11;
  o1.y = dx * d02 + dy * d12;
  o1.z = I;
  o1.w = 1.0f;
  *reinterpret_cast<float4*>(J + 8 * (size_t)i) = o0;
  *reinterpret_cast<float4*>(J + 8 * (size_t)i + 4) = o1;
}

// Reprojection of the last keyframe's finest cloud into the current frame: per target pixel keep the
// point with the largest index (the sequential "last write wins" of coords.fill_image).
__global__ void reproj_scatter_kernel(const float* __restrict__ P, int n, const float* __restrict__ T, float fx, float fy,
                                      float cx, float cy, int h, int w, int* __restrict__ winner) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float X0 = P[3 * i], Y0 = P[3 * i + 1], Z0 = P[3 * i + 2];
  const float X = T[0] * X0 + T[1] * Y0 + T[2] * Z0 + T[3];
  const float Y = T[4] * X0 + T[5] * Y0 + T[6] * Z0 + T[7];
  const float Z = T[8] * X0 + T[9] * Y0 + T[10] * Z0 + T[11];
  const float px = fx * X / Z + cx, py = fy * Y / Z + cy;
  const bool ok = (px > 0.0f) && (px < (float)(w - 1)) && (py > 0.0f) && (py < (float)(h - 1)) && (Z > 0.0f);
  if (ok) atomicMax(&winner[(int)py * w + (int)px], i);
}

__global__ void reproj_gather_kernel(const float* __restrict__ P, const float* __restrict__ T, const int* __restrict__ winner,
                                     int hw, float* __restrict__ depth_img) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= hw) return;
  const int i = winner[t];
  float v = __int_as_float(0x7fc00000);
  if (i >= 0) v = T[8] * P[3 * i] + T[9] * P[3 * i + 1] + T[10] * P[3 * i + 2] + T[11];
  depth_img[t] = v;
}

}  // namespace como

using namespace como;

extern "C" int como_b200_gray_pyramid(const float* rgb, int32_t H, int32_t W, int32_t num_levels, float* const* levels,
                                      void* stream) {
  // levels: HOST array of num_levels DEVICE pointers, COARSEST FIRST (levels[num_levels-1] is H x W)
  COMO_REQUIRE(rgb && levels && num_levels >= 1 && num_levels <= COMO_B200_MAX_LEVELS, "gray_pyramid: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const long long hw = (long long)H * W;
  gray_kernel<<<(unsigned)((hw + 255) / 256), 256, 0, st>>>(rgb, hw, levels[num_levels - 1]);
  int h = H, w = W;
  for (int l = num_levels - 2; l >= 0; --l) {
    const int ho = (h + 1) / 2, wo = (w + 1) / 2;
    dim3 blk(32, 8), grd((wo + 31) / 32, (ho + 7) / 8);
    blur_down_kernel<<<grd, blk, 0, st>>>(levels[l + 1], h, w, levels[l], ho, wo);
    h = ho;
    w = wo;
  }
  return check_launch("gray_pyramid");
}

extern "C" int como_b200_image_pyramid_fused(const float* rgb, int32_t H, int32_t W, int32_t num_levels, float* const* levels,
                                             float* const* gx, float* const* gy, void* stream) {
  // levels / gx / gy: HOST arrays of num_levels DEVICE pointers, COARSEST FIRST (as the reference stores pyramids);
  // gx, gy may be NULL (tracking frames need the images only).  num_levels <= 4; larger pyramids use gray_pyramid.
  COMO_REQUIRE(rgb && levels && num_levels >= 1 && num_levels <= PYR_MAX_LEVELS, "image_pyramid_fused: bad arguments");
  COMO_REQUIRE((gx == nullptr) == (gy == nullptr), "image_pyramid_fused: gx and gy go together");
  COMO_REQUIRE(H >= 2 && W >= 2, "image_pyramid_fused: image too small");
  PyrArgs a;
  a.num_levels = num_levels;
  int h = H, w = W, off = 0;
  for (int l = 0; l < num_levels; ++l) {   // kernel order: finest first
    a.img[l] = levels[num_levels - 1 - l];
    a.gx[l] = gx ? gx[num_levels - 1 - l] : nullptr;
    a.gy[l] = gy ? gy[num_levels - 1 - l] : nullptr;
    COMO_REQUIRE(a.img[l] != nullptr, "image_pyramid_fused: null level pointer");
    COMO_REQUIRE(h >= 2 && w >= 2, "image_pyramid_fused: level %d degenerates (%d x %d)", l, h, w);
    a.h[l] = h;
    a.w[l] = w;
    h = (h + 1) / 2;
    w = (w + 1) / 2;
  }
  a.halo[num_levels - 1] = 1;
  for (int l = num_levels - 2; l >= 0; --l) a.halo[l] = 2 * a.halo[l + 1] + 1;
  for (int l = 0; l < num_levels; ++l) {
    const int rs = (PYR_TILE >> l) + 2 * a.halo[l];
    a.off[l] = off;
    off += rs * rs;
  }
  for (int l = num_levels; l < PYR_MAX_LEVELS; ++l) {
    a.img[l] = a.gx[l] = a.gy[l] = nullptr;
    a.h[l] = a.w[l] = a.halo[l] = a.off[l] = 0;
  }
  const size_t smem = (size_t)off * sizeof(float);
  cudaFuncSetAttribute(pyramid_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((W + PYR_TILE - 1) / PYR_TILE, (H + PYR_TILE - 1) / PYR_TILE);
  pyramid_fused_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(rgb, a);
  return check_launch("image_pyramid_fused");
}

extern "C" int como_b200_image_gradients(const float* img, int32_t h, int32_t w, float* gx, float* gy, void* stream) {
  COMO_REQUIRE(img && gx && gy && h >= 2 && w >= 2, "image_gradients: bad arguments");
  dim3 blk(32, 8), grd((w + 31) / 32, (h + 7) / 8);
  scharr_kernel<<<grd, blk, 0, (cudaStream_t)stream>>>(img, h, w, gx, gy);
  return check_launch("image_gradients");
}

extern "C" int como_b200_img_and_grads_f64(const double* rgb, int32_t h, int32_t w, double* img_and_grads, void* stream) {
  COMO_REQUIRE(rgb && img_and_grads && h >= 2 && w >= 2, "img_and_grads_f64: bad arguments");
  dim3 blk(32, 8), grd((w + 31) / 32, (h + 7) / 8);
  img_and_grads_f64_kernel<<<grd, blk, 0, (cudaStream_t)stream>>>(rgb, h, w, img_and_grads);
  return check_launch("img_and_grads_f64");
}

extern "C" int como_b200_kf_reference_level(const float* img, const float* gx, const float* gy, const float* depth_full,
                                            int32_t Hf, int32_t Wf, int32_t sub, int32_t h, int32_t w, const float* K9,
                                            const float* rel12_dev, float border, float depth_thresh, float* vals,
                                            float* grads, float* P, float* J, uint8_t* mask, void* stream) {
  COMO_REQUIRE(img && gx && gy && depth_full && K9 && rel12_dev && vals && grads && P && J && mask,
               "kf_reference_level: null pointer argument");
  COMO_REQUIRE(((uintptr_t)J & 15) == 0, "kf_reference_level: J must be 16-byte aligned");
  const int n = h * w;
  kf_reference_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(img, gx, gy, depth_full, Hf, Wf, sub, h, w, K9[0],
                                                                        K9[4], K9[2], K9[5], rel12_dev, border, depth_thresh,
                                                                        vals, grads, P, J, mask);
  return check_launch("kf_reference_level");
}

extern "C" int como_b200_reproj_depth(const float* P, int32_t n, const float* T_dev, const float* K9, int32_t h, int32_t w,
                                      int32_t* winner_ws, float* depth_img, void* stream) {
  COMO_REQUIRE(P && T_dev && K9 && winner_ws && depth_img, "reproj_depth: null pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(winner_ws, 0xff, sizeof(int) * (size_t)h * w, st);
  reproj_scatter_kernel<<<(n + 255) / 256, 256, 0, st>>>(P, n, T_dev, K9[0], K9[4], K9[2], K9[5], h, w, winner_ws);
  reproj_gather_kernel<<<(h * w + 255) / 256, 256, 0, st>>>(P, T_dev, winner_ws, h * w, depth_img);
  return check_launch("reproj_depth");
}
