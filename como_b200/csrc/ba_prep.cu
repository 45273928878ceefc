// Per-iteration preparation kernels of the window BA (fp64):
//   subselect_pixels  (como/odom/backend/sparse_map.py:116-142)     -- per KF, cached by the host
//   predictor apply   (como/odom/Mapping.py:749-758 store_vars)     -- depth = exp(Knm_Kmminv . logz_m)
//   frame table       (inverse poses, affine)                         -- feeds the pair kernels
//   scaffold          (Mapping.py:603-659 + sparse_map.py:18-60)     -- anchors -> per-KF log depths + Jacobians
//   update_vars       (como/odom/backend/linear_system.py:115-152)
#include "ba_common.cuh"

namespace como {

// ------------------------------------------------------------------------------------------------
// 4x4 (win x win) non-max selection: first strict maximum of sqrt(gx^2+gy^2) in row-major order inside
// each cell (max_pool2d(return_indices) semantics).  Bit-exact: the magnitude is formed with the same
// individually rounded operations torch uses (square, add, sqrt), no FMA contraction.
// ------------------------------------------------------------------------------------------------
__global__ void subselect_kernel(const double* __restrict__ img3, int K, int H, int W, int win,
                                 int32_t* __restrict__ coords, double* __restrict__ vals) {
  const int hc = H / win, wc = W / win;
  const int n = hc * wc;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)K * n) return;
  const int k = (int)(idx / n), cell = (int)(idx % n);
  const int cr = cell / wc, cc = cell % wc;
  const double* I = img3 + (size_t)k * 3 * H * W;
  const double* gx = I + (size_t)H * W;
  const double* gy = I + 2 * (size_t)H * W;
  double best = -1.0;
  int br = cr * win, bc = cc * win;
  for (int dr = 0; dr < win; ++dr)
    for (int dc = 0; dc < win; ++dc) {
      const int r = cr * win + dr, c = cc * win + dc;
      const double x = gx[(size_t)r * W + c], y = gy[(size_t)r * W + c];
      const double m = __dsqrt_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
      if (m > best || (m != m)) {
        best = m;
        br = r;
        bc = c;
      }
    }
  coords[2 * idx] = br;
  coords[2 * idx + 1] = bc;
  vals[idx] = I[(size_t)br * W + bc];
}

// ------------------------------------------------------------------------------------------------
// Predictor apply: out[k, p] = exp( sum_m Knm[k, p, m] * logzm[k, m] ).  Pure HBM streaming
// (K*H*W*M*8 bytes read: 5.03 GB at K=32, 640x480, M=64; measured 790 us = 6.37 TB/s on B200).
// A CTA walks 64-row chunks of the flattened (keyframe, pixel) row space; each chunk (64*M*8 bytes,
// contiguous in HBM) is fetched by ONE bulk async copy (TMA unit) into a 3-deep shared-memory ring, so
// up to 96 KB per CTA are in flight.  The modest register footprint lets the host run store_vars on a
// side stream next to the residual / median kernels of the normal-equation build.  Warps 0..7 reduce rows
// 8w..8w+7 of a chunk (lane l reads one 16-byte piece of each row, the 8 dot products are finished with a
// transposed butterfly); warp 8 is the producer: "empty" mbarriers (one arrival per consumer warp) hand a
// slot back to it without any block-wide barrier.
// ------------------------------------------------------------------------------------------------
constexpr int PA_ROWS = 8;
constexpr int PS_ROWS = 64, PS_STAGES = 3, PS_CONSUMERS = 8, PS_THREADS = 32 * (PS_CONSUMERS + 1);

struct PredStreamSmem {
  double X[PS_STAGES][PS_ROWS * BA_MAXM];
  unsigned long long full[PS_STAGES], empty[PS_STAGES];
};

__global__ void __launch_bounds__(PS_THREADS, 2)
predictor_stream_kernel(const double* __restrict__ Knm, const double* __restrict__ scaf, int K, long long HW, int M,
                        long long chunks_per_kf, double* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char ps_raw[];
  PredStreamSmem& S = *reinterpret_cast<PredStreamSmem*>(ps_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long total = chunks_per_kf * K;
  if (tid == 0) {
    for (int q = 0; q < PS_STAGES; ++q) {
      mbar_init(&S.full[q], 1);
      mbar_init(&S.empty[q], PS_CONSUMERS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == PS_CONSUMERS) {
    if (lane != 0) return;
    long long it = 0;
    for (long long c = blockIdx.x; c < total; c += gridDim.x, ++it) {
      const int slot = (int)(it % PS_STAGES);
      if (it >= PS_STAGES) mbar_wait(&S.empty[slot], (unsigned)(((it / PS_STAGES) - 1) & 1));
      const int k = (int)(c / chunks_per_kf);
      const long long p0 = (c - (long long)k * chunks_per_kf) * PS_ROWS;
      const long long rows = (HW - p0 < PS_ROWS) ? (HW - p0) : PS_ROWS;
      const unsigned bytes = (unsigned)(rows * M * sizeof(double));
      mbar_expect_tx(&S.full[slot], bytes);
      bulk_g2s(&S.X[slot][0], Knm + ((size_t)k * HW + p0) * M, bytes, &S.full[slot]);
    }
    return;
  }
  const int m0 = 2 * lane;
  const bool act = m0 < M;
  int kcur = -1;
  double l0 = 0.0, l1 = 0.0;
  long long it = 0;
  for (long long c = blockIdx.x; c < total; c += gridDim.x, ++it) {
    const int slot = (int)(it % PS_STAGES);
    const unsigned parity = (unsigned)((it / PS_STAGES) & 1);
    const int k = (int)(c / chunks_per_kf);
    const long long p0 = (c - (long long)k * chunks_per_kf) * PS_ROWS;
    if (k != kcur) {
      kcur = k;
      l0 = act ? __ldg(scaf + ((size_t)k * M + m0) * SCAF_STRIDE) : 0.0;
      l1 = act ? __ldg(scaf + ((size_t)k * M + m0 + 1) * SCAF_STRIDE) : 0.0;
    }
    mbar_wait(&S.full[slot], parity);
    const double* x = &S.X[slot][(size_t)(PA_ROWS * warp) * M];
    double acc[PA_ROWS];
#pragma unroll
    for (int j = 0; j < PA_ROWS; ++j) {
      const double2 v = act ? *reinterpret_cast<const double2*>(x + (size_t)j * M + m0) : make_double2(0.0, 0.0);
      acc[j] = v.x * l0 + v.y * l1;
    }
    stage_release(&S.empty[slot]);
    // transposed butterfly: 8 rows over 32 lanes -> lanes with (lane & 3) == 0 hold one row total each
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool hi = (lane & 16) != 0;
      const double send = hi ? acc[j] : acc[j + 4];
      const double recv = __shfl_xor_sync(0xffffffffu, send, 16);
      acc[j] = (hi ? acc[j + 4] : acc[j]) + recv;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const bool hi = (lane & 8) != 0;
      const double send = hi ? acc[j] : acc[j + 2];
      const double recv = __shfl_xor_sync(0xffffffffu, send, 8);
      acc[j] = (hi ? acc[j + 2] : acc[j]) + recv;
    }
    {
      const bool hi = (lane & 4) != 0;
      const double send = hi ? acc[0] : acc[1];
      const double recv = __shfl_xor_sync(0xffffffffu, send, 4);
      acc[0] = (hi ? acc[1] : acc[0]) + recv;
    }
    acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 2);
    acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 1);
    const int row = PA_ROWS * warp + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    if ((lane & 3) == 0) {
      const long long p = p0 + row;
      if (p < HW) out[(size_t)k * HW + p] = exp(acc[0]);
    }
  }
}

// column means of one keyframe's predictor (mean_log_depth_cost, gp_priors.py:99-106): colmean[m] = mean_p Knm[p,m]
__global__ void __launch_bounds__(256)
colmean_kernel(const double* __restrict__ Knm, long long HW, int M, double* __restrict__ colsum) {
  // grid-stride over pixels; thread t handles column t % M of pixel rows t / M + ...
  __shared__ double s[256];
  const int tid = threadIdx.x;
  const int per = 256 / M;  // rows handled per pass by the CTA (M divides 256 for M in {16,32,64})
  const int m = tid % M, r0 = tid / M;
  double acc = 0.0;
  if (r0 < per)
    for (long long p = (long long)blockIdx.x * per + r0; p < HW; p += (long long)gridDim.x * per) acc += Knm[(size_t)p * M + m];
  s[tid] = acc;
  __syncthreads();
  if (tid < M) {
    double t = 0.0;
    for (int r = 0; r < per; ++r) t += s[r * M + tid];
    atomicAdd(colsum + tid, t);
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void frames_kernel(const double* __restrict__ kf_poses, const double* __restrict__ kf_aff,
                              const double* __restrict__ rec_poses, const double* __restrict__ rec_aff,
                              const double* __restrict__ kf_img, const double* __restrict__ rec_img, int K, int R,
                              size_t img_stride, BAFrame* __restrict__ frames) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= K + R) return;
  const double* T = (f < K) ? kf_poses + 16 * (size_t)f : rec_poses + 16 * (size_t)(f - K);
  const double* af = (f < K) ? kf_aff + 2 * (size_t)f : rec_aff + 2 * (size_t)(f - K);
  BAFrame F;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) {
      F.Rwc[r * 3 + c] = T[r * 4 + c];
      F.Rcw[c * 3 + r] = T[r * 4 + c];
    }
    F.twc[r] = T[r * 4 + 3];
  }
  for (int r = 0; r < 3; ++r) F.tcw[r] = -(F.Rcw[r * 3] * F.twc[0] + F.Rcw[r * 3 + 1] * F.twc[1] + F.Rcw[r * 3 + 2] * F.twc[2]);
  F.a = af[0];
  F.b = af[1];
  F.img = (f < K) ? kf_img + img_stride * f : rec_img + img_stride * (f - K);
  frames[f] = F;
}

// ------------------------------------------------------------------------------------------------
// Scaffold: one thread per (keyframe, anchor slot).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void init_point(const double* T, double med, const double* pfo, double fx, double fy,
                                           double cx, double cy, double* Pw) {
  // backproject the first-observation pixel at the keyframe's median depth, into the world frame
  const double Pc[3] = {med * ((pfo[0] - cx) / fx), med * ((pfo[1] - cy) / fy), med};
  for (int r = 0; r < 3; ++r) Pw[r] = T[r * 4] * Pc[0] + T[r * 4 + 1] * Pc[1] + T[r * 4 + 2] * Pc[2] + T[r * 4 + 3];
}

__global__ void scaffold_kernel(const double* __restrict__ kf_poses, const double* __restrict__ P_m,
                                const int32_t* __restrict__ lm_ids, const int32_t* __restrict__ fo_slots,
                                const double* __restrict__ pm_first_obs, const double* __restrict__ med, BADims d,
                                double* __restrict__ scaf, double* __restrict__ dz_dP) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= d.K * d.M) return;
  const int k = idx / d.M;
  const double* T = kf_poses + 16 * (size_t)k;
  const int l = lm_ids[idx];
  double Pw[3] = {P_m[3 * (size_t)l], P_m[3 * (size_t)l + 1], P_m[3 * (size_t)l + 2]};
  double Pc[3];
  auto to_cam = [&](const double* w, double* c) {
    const double dx = w[0] - T[3], dy = w[1] - T[7], dzz = w[2] - T[11];
    // R^T (w - t) written as R_cw w + t_cw with t_cw = -R^T t (same association as the reference)
    double tcw[3];
    for (int r = 0; r < 3; ++r) tcw[r] = -(T[0 * 4 + r] * T[3] + T[1 * 4 + r] * T[7] + T[2 * 4 + r] * T[11]);
    for (int r = 0; r < 3; ++r) c[r] = T[0 * 4 + r] * w[0] + T[1 * 4 + r] * w[1] + T[2 * 4 + r] * w[2] + tcw[r];
    (void)dx; (void)dy; (void)dzz;
  };
  to_cam(Pw, Pc);
  const bool zmask = Pc[2] < 0.1 * med[k];
  if (zmask) {
    // reference quirk kept: the re-initialisation table is built in (keyframe, slot) order of the first
    // observations but indexed by landmark id (Mapping.py:626-648)
    const int slot = fo_slots[l];
    const int k2 = slot / d.M;
    double rPw[3];
    init_point(kf_poses + 16 * (size_t)k2, med[k2], pm_first_obs + 2 * (size_t)slot, d.fx, d.fy, d.cx, d.cy, rPw);
    to_cam(rPw, Pc);
  }
  const double z = Pc[2];
  double* o = scaf + (size_t)idx * SCAF_STRIDE;
  const double u = 1.0 / z;
  o[0] = log(z);
  o[1] = u;
  o[2] = d.fx * Pc[0] / z + d.cx;
  o[3] = d.fy * Pc[1] / z + d.cy;
  o[4] = Pc[0];
  o[5] = Pc[1];
  o[6] = Pc[2];
  o[7] = zmask ? 1.0 : 0.0;
  // dz/dTwc = row 2 of [Pc^ | -I] = [-Pc.y, Pc.x, 0, 0, 0, -1];  dlogz = u * dz
  o[8] = -u * Pc[1];
  o[9] = u * Pc[0];
  o[10] = 0.0;
  o[11] = 0.0;
  o[12] = 0.0;
  o[13] = -u;
  o[14] = o[15] = 0.0;
  if (idx % d.M == 0) {
    dz_dP[3 * k + 0] = T[0 * 4 + 2];  // row 2 of R_cw = column 2 of R_wc
    dz_dP[3 * k + 1] = T[1 * 4 + 2];
    dz_dP[3 * k + 2] = T[2 * 4 + 2];
  }
}

// P_m[j] <- init_Pm[j] where the j-th first observation (row-major) was behind the camera
__global__ void reinit_kernel(const double* __restrict__ kf_poses, const int32_t* __restrict__ fo_slots,
                              const double* __restrict__ pm_first_obs, const double* __restrict__ med,
                              const double* __restrict__ scaf, BADims d, double* __restrict__ P_m) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= d.L) return;
  const int slot = fo_slots[j];
  if (scaf[(size_t)slot * SCAF_STRIDE + 7] != 0.0) {
    const int k2 = slot / d.M;
    double Pw[3];
    init_point(kf_poses + 16 * (size_t)k2, med[k2], pm_first_obs + 2 * (size_t)slot, d.fx, d.fy, d.cx, d.cy, Pw);
    P_m[3 * (size_t)j] = Pw[0];
    P_m[3 * (size_t)j + 1] = Pw[1];
    P_m[3 * (size_t)j + 2] = Pw[2];
  }
}

// ------------------------------------------------------------------------------------------------
// update_vars: T <- T Exp(delta[0:6]) ([omega,v] -> lietorch [tau=v, phi=omega]); affine += delta[6:8];
// landmarks += delta.
// ------------------------------------------------------------------------------------------------
__global__ void update_kernel(const double* __restrict__ delta, BADims d, double* __restrict__ kf_poses,
                              double* __restrict__ kf_aff, double* __restrict__ rec_poses, double* __restrict__ rec_aff,
                              double* __restrict__ P_m) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int nf = d.K + d.R;
  if (idx < nf) {
    const double* dl = delta + 8 * (size_t)idx;
    double* T = (idx < d.K) ? kf_poses + 16 * (size_t)idx : rec_poses + 16 * (size_t)(idx - d.K);
    double* af = (idx < d.K) ? kf_aff + 2 * (size_t)idx : rec_aff + 2 * (size_t)(idx - d.K);
    const double tau[3] = {dl[3], dl[4], dl[5]}, phi[3] = {dl[0], dl[1], dl[2]};
    double E[16], Tn[16];
    se3_exp_tau_phi(tau, phi, E);
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) {
        double s = 0.0;
        for (int q = 0; q < 4; ++q) s += T[r * 4 + q] * E[q * 4 + c];
        Tn[r * 4 + c] = s;
      }
    for (int q = 0; q < 16; ++q) T[q] = Tn[q];
    af[0] += dl[6];
    af[1] += dl[7];
  } else {
    const int j = idx - nf;
    if (j < 3 * d.L) P_m[j] += delta[8 * (size_t)nf + j];
  }
}

}  // namespace como

using namespace como;

extern "C" int como_b200_subselect_pixels(const double* img_and_grads, int32_t K, int32_t H, int32_t W, int32_t win,
                                          int32_t* coords, double* vals, void* stream) {
  COMO_REQUIRE(img_and_grads && coords && vals, "subselect_pixels: null pointer argument");
  COMO_REQUIRE(K >= 1 && win >= 1 && H >= win && W >= win, "subselect_pixels: bad shape");
  const long long total = (long long)K * (H / win) * (W / win);
  subselect_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(img_and_grads, K, H, W, win, coords, vals);
  return check_launch("subselect_pixels");
}

// 0 = one CTA per SM (default).  Other grids are for experiments (a smaller grid leaves SMs to a kernel on another stream,
// but one CTA pulls only ~45 GB/s, so the stream slows down in proportion: profiles/r02_corun_probe.txt).
static int g_predictor_stream_ctas = 0;
extern "C" void como_b200_predictor_stream_ctas(int32_t ctas) { g_predictor_stream_ctas = ctas > 0 ? ctas : 0; }

extern "C" int como_b200_predictor_apply(const double* Knm, const double* scaffold, int32_t K, int64_t HW, int32_t M,
                                         double* depth, void* stream) {
  COMO_REQUIRE(Knm && scaffold && depth, "predictor_apply: null pointer argument");
  COMO_REQUIRE(K >= 1 && HW >= 1 && M >= 2 && M <= BA_MAXM && (M % 2) == 0, "predictor_apply: bad shape (M even, <= 64)");
  // the attribute is per device: set on every call (cheap) rather than once per process
  cudaFuncSetAttribute(predictor_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PredStreamSmem));
  const long long chunks_per_kf = (HW + PS_ROWS - 1) / PS_ROWS;
  // one CTA (96 KB in flight) per SM: measured 0.743 ms for the 5.03 GB of a K = 32 window against 0.783 ms with two per
  // SM (profiles/r02_corun_probe.txt)
  long long grid = sm_count();
  if (g_predictor_stream_ctas > 0) grid = g_predictor_stream_ctas;   // see como_b200_predictor_stream_ctas
  if (grid > chunks_per_kf * K) grid = chunks_per_kf * K;
  predictor_stream_kernel<<<(unsigned)grid, PS_THREADS, sizeof(PredStreamSmem), (cudaStream_t)stream>>>(
      Knm, scaffold, K, HW, M, chunks_per_kf, depth);
  return check_launch("predictor_apply");
}

extern "C" int como_b200_predictor_colsum(const double* Knm, int64_t HW, int32_t M, double* colsum, void* stream) {
  COMO_REQUIRE(Knm && colsum, "predictor_colsum: null pointer argument");
  COMO_REQUIRE(M >= 1 && M <= BA_MAXM && 256 % M == 0, "predictor_colsum: M must divide 256");
  cudaMemsetAsync(colsum, 0, sizeof(double) * M, (cudaStream_t)stream);
  colmean_kernel<<<sm_count() * 4, 256, 0, (cudaStream_t)stream>>>(Knm, HW, M, colsum);
  return check_launch("predictor_colsum");
}

extern "C" int como_b200_ba_scaffold(const double* kf_poses, double* P_m, const int32_t* lm_ids, const int32_t* fo_slots,
                                     const double* pm_first_obs, const double* median_depths, int32_t K, int32_t L,
                                     int32_t M, const double* intr4, double* scaffold, double* dz_dP, void* stream) {
  COMO_REQUIRE(kf_poses && P_m && lm_ids && fo_slots && pm_first_obs && median_depths && intr4 && scaffold && dz_dP,
               "ba_scaffold: null pointer argument");
  COMO_REQUIRE(K >= 1 && M >= 1 && M <= BA_MAXM && L >= 1, "ba_scaffold: bad shape");
  BADims d{};
  d.K = K; d.L = L; d.M = M;
  d.fx = intr4[0]; d.fy = intr4[1]; d.cx = intr4[2]; d.cy = intr4[3];
  cudaStream_t st = (cudaStream_t)stream;
  scaffold_kernel<<<(K * M + 127) / 128, 128, 0, st>>>(kf_poses, P_m, lm_ids, fo_slots, pm_first_obs, median_depths, d, scaffold, dz_dP);
  reinit_kernel<<<(L + 127) / 128, 128, 0, st>>>(kf_poses, fo_slots, pm_first_obs, median_depths, scaffold, d, P_m);
  return check_launch("ba_scaffold");
}

extern "C" int como_b200_ba_update(const double* delta, int32_t K, int32_t R, int32_t L, double* kf_poses, double* kf_aff,
                                   double* rec_poses, double* rec_aff, double* P_m, void* stream) {
  COMO_REQUIRE(delta && kf_poses && kf_aff && P_m, "ba_update: null pointer argument");
  COMO_REQUIRE(R == 0 || (rec_poses && rec_aff), "ba_update: null recent pointers");
  BADims d{};
  d.K = K; d.R = R; d.L = L;
  const int total = K + R + 3 * L;
  update_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(delta, d, kf_poses, kf_aff, rec_poses, rec_aff, P_m);
  return check_launch("ba_update");
}

namespace como {
int ba_build_frames(const double* kf_poses, const double* kf_aff, const double* rec_poses, const double* rec_aff,
                    const double* kf_img, const double* rec_img, int K, int R, size_t img_stride, BAFrame* frames,
                    cudaStream_t st) {
  frames_kernel<<<(K + R + 63) / 64, 64, 0, st>>>(kf_poses, kf_aff, rec_poses, rec_aff, kf_img, rec_img, K, R, img_stride, frames);
  return check_launch("ba_frames");
}
}  // namespace como
