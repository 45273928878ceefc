// DepthCov Gaussian-process kernels (compiled with -fmad=false: the fp32 selection path must not be
// re-associated; the GEMM part uses explicit fma()).
//
//  * cross_covariance / get_new_chol_obs_info : the two operators of the reference's pybind module
//    `como_backends` (como/backend/src/cov.cpp:5-65, cov_cpu.cpp:17-85, cov_gpu.cu:17-215), same formula:
//      K12 = scale * 2 (|E1||E2|)^(1/4) sqrt(1/|E1+E2| + 1e-8) * matern32(Q),  Q = d^T (E1+E2)^-1 d / 2
//  * sampler step: the greedy conditional-entropy loop of como/depth_cov/core/samplers.py:211-302 run on
//    the device (one small launch + one domain-wide launch per selected anchor, no host round trips).
//  * K-matrix / predictor: CovarianceModule / CrossCovarianceModule (como/depth_cov/core/covariance.py:10-39,
//    kernels.py:22-89 -- note the Python formula differs from the native one: 1/sqrt(|E1+E2| + 1e-8), and
//    the coordinate difference is rounded to float32) fused with the (HW x M) x (M x M) product of
//    Mapping.prep_predictor (como/odom/Mapping.py:430-468).
#include "ba_common.cuh"

namespace como {

// ------------------------------------------------------------------------------------------------ native formula
// fp32 path mirrors cov_cpu.cpp's mixed float/double expression types; pow(x, 0.25f) and exp are
// evaluated in double and rounded once (glibc's powf/expf are correctly rounded to within ~0.5 ulp).
__device__ __forceinline__ float cov_native_f32(float x1x, float x1y, float e1_00, float e1_01, float e1_10, float e1_11,
                                                float x2x, float x2y, float e2_00, float e2_01, float e2_10, float e2_11,
                                                float scale) {
  const float dx = x1x - x2x, dy = x1y - x2y;
  const float E00 = e1_00 + e2_00, E01 = e1_01 + e2_01, E11 = e1_11 + e2_11;
  const float det = E00 * E11 - E01 * E01;
  const float det_inv = (float)(1.0 / (double)det);
  float Q = (E11 * dx * dx) - 2.0f * (E01 * dx * dy) + (E00 * dy * dy);
  Q = (float)((double)Q * (0.5 * (double)det_inv));
  const float d1 = e1_00 * e1_11 - e1_01 * e1_10;
  const float d2 = e2_00 * e2_11 - e2_01 * e2_10;
  const float pw = (float)sqrt(sqrt((double)(d1 * d2)));
  const float ssq = (float)sqrt((double)det_inv + 1e-8);
  const float Cc = (float)(2.0 * (double)pw * (double)ssq);
  const float sq = (float)sqrt((double)Q + 1e-8);
  const float tmp = (float)(1.73205080757 * (double)sq);
  const float mat = (1.0f + tmp) * (float)exp(-(double)tmp);
  return scale * Cc * mat;
}

__device__ __forceinline__ double cov_native_f64(double x1x, double x1y, double e1_00, double e1_01, double e1_10,
                                                 double e1_11, double x2x, double x2y, double e2_00, double e2_01,
                                                 double e2_10, double e2_11, double scale) {
  const double dx = x1x - x2x, dy = x1y - x2y;
  const double E00 = e1_00 + e2_00, E01 = e1_01 + e2_01, E11 = e1_11 + e2_11;
  const double det_inv = 1.0 / (E00 * E11 - E01 * E01);
  double Q = (E11 * dx * dx) - 2.0 * (E01 * dx * dy) + (E00 * dy * dy);
  Q *= 0.5 * det_inv;
  const double d1 = e1_00 * e1_11 - e1_01 * e1_10, d2 = e2_00 * e2_11 - e2_01 * e2_10;
  // the reference's safe_sqrt/matern helpers take and return float even in the double kernel
  const double Cc = 2.0 * pow(d1 * d2, 0.25) * (double)(float)sqrt((double)(float)det_inv + 1e-8);
  const float sq = (float)sqrt((double)(float)Q + 1e-8);
  const float tmp = (float)(1.73205080757 * (double)sq);
  const float mat = (1.0f + tmp) * (float)exp(-(double)tmp);
  return scale * Cc * (double)mat;
}

template <typename T>
__global__ void cross_cov_kernel(const T* __restrict__ x1, const T* __restrict__ E1, const T* __restrict__ x2,
                                 const T* __restrict__ E2, T scale, int n1, int n2, T* __restrict__ out) {
  const int b = blockIdx.y;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (long long)n1 * n2) return;
  const int i = (int)(p / n2), j = (int)(p % n2);
  const T* a = x1 + ((size_t)b * n1 + i) * 2;
  const T* A = E1 + ((size_t)b * n1 + i) * 4;
  const T* c = x2 + ((size_t)b * n2 + j) * 2;
  const T* Cm = E2 + ((size_t)b * n2 + j) * 4;
  T v;
  if constexpr (sizeof(T) == 4)
    v = cov_native_f32(a[0], a[1], A[0], A[1], A[2], A[3], c[0], c[1], Cm[0], Cm[1], Cm[2], Cm[3], scale);
  else
    v = cov_native_f64(a[0], a[1], A[0], A[1], A[2], A[3], c[0], c[1], Cm[0], Cm[1], Cm[2], Cm[3], scale);
  out[(size_t)b * n1 * n2 + p] = v;
}

// ------------------------------------------------------------------------------------------------ Cholesky append
// L (B,n,n) row-major; appends row N:  L[N,:N] = L[:N,:N]^-1 k_ni (forward substitution),
// L[N,N] = sqrt(k_ii - sum L[N,:]^2).  One CTA per batch element, column oriented: thread r owns rows r, r + 128, ...;
// after column c is resolved every later row subtracts its L[r][c] * v_c -- each row still receives its subtractions
// in ascending column order, i.e. the same fp32 sequence as the reference's serial loop (no FMA contraction in this
// file), with N block barriers instead of N^2 / 2 dependent operations in one thread.
constexpr int CA_THREADS = 128;
constexpr int CA_ROWS = 8;   // rows per thread: N <= 1024
__global__ void __launch_bounds__(CA_THREADS)
chol_append_kernel(float* __restrict__ L, const float* __restrict__ k_ni, float k_ii, int n, int N) {
  const int b = blockIdx.x;
  float* Lb = L + (size_t)b * n * n;
  __shared__ float s_v;
  float acc[CA_ROWS];
#pragma unroll
  for (int q = 0; q < CA_ROWS; ++q) {
    const int r = threadIdx.x + q * CA_THREADS;
    acc[q] = (r < N) ? k_ni[(size_t)b * N + r] : 0.0f;
  }
  float ss = 0.0f;   // thread 0 only
  for (int c = 0; c < N; ++c) {
    if (threadIdx.x == c % CA_THREADS) {
      float a = 0.0f;
#pragma unroll
      for (int q = 0; q < CA_ROWS; ++q)
        if (q == c / CA_THREADS) a = acc[q];
      const float v = a / Lb[c * n + c];
      Lb[N * n + c] = v;
      s_v = v;
    }
    __syncthreads();
    const float v = s_v;
    if (threadIdx.x == 0) ss += v * v;
#pragma unroll
    for (int q = 0; q < CA_ROWS; ++q) {
      const int r = threadIdx.x + q * CA_THREADS;
      if (r > c && r < N) acc[q] -= Lb[r * n + c] * v;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) Lb[N * n + N] = sqrtf(k_ii - ss);
}

// obs_info (B,n,d): new row N = (k_id - sum_i L[N,i] obs_info[i,:]) / L[N,N];  var -= row^2
__global__ void obs_info_kernel(const float* __restrict__ L, float* __restrict__ obs_info, float* __restrict__ var,
                                const float* __restrict__ k_id, int n, int d, int N) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= d) return;
  const float* Lr = L + (size_t)b * n * n + (size_t)N * n;
  float* ob = obs_info + (size_t)b * n * d;
  float s = 0.0f;
  for (int i = 0; i < N; ++i) s += Lr[i] * ob[(size_t)i * d + j];
  const float v = (k_id[(size_t)b * d + j] - s) / Lr[N];
  ob[(size_t)N * d + j] = v;
  var[(size_t)b * d + j] -= v * v;
}

// ------------------------------------------------------------------------------------------------ sampler (device loop)
struct SamplerState {   // per batch element, in global memory
  int count;            // anchors selected so far
  int next;             // domain index chosen for the next step
  float next_stdev;     // its stdev (+1e-10)
  int done;             // early termination flag
};

// step A (1 CTA per batch): commit `next` as anchor number i=count: coords/E, k_ni against the previous
// anchors, Cholesky row.  Mirrors greedy_loop lines 252-274 of samplers.py.
__device__ __forceinline__ void sampler_commit_body(int b, SamplerState* st, const float* dom_xy,
                                                    const float* dom_E, int d, float* sel_xy,
                                                    float* sel_E, long long* sel_idx,
                                                    float* L, int n, float signal_var, float fixed_var,
                                                    int has_fixed, float max_stdev_thresh, int terminate_early) {
  volatile SamplerState& s = st[b];
  __shared__ float k_ni[128];
  const int i = s.count;
  if (s.done || i >= n) return;
  if (terminate_early && s.next_stdev < max_stdev_thresh) {   // batch size 1 in every reference call site
    __syncthreads();            // every thread has read the state before it changes
    if (threadIdx.x == 0) s.done = 1;
    return;
  }
  const int j = s.next;
  const float* xj = dom_xy + ((size_t)b * d + j) * 2;
  const float* Ej = dom_E + ((size_t)b * d + j) * 4;
  float* sx = sel_xy + (size_t)b * n * 2;
  float* sE = sel_E + (size_t)b * n * 4;
  for (int r = threadIdx.x; r < i; r += blockDim.x)
    k_ni[r] = cov_native_f32(sx[2 * r], sx[2 * r + 1], sE[4 * r], sE[4 * r + 1], sE[4 * r + 2], sE[4 * r + 3], xj[0], xj[1],
                             Ej[0], Ej[1], Ej[2], Ej[3], signal_var);
  __syncthreads();
  // forward substitution, column oriented: thread r owns row r; every acc[r] still receives its
  // subtractions in ascending column order (the same fp32 sequence as the row-oriented loop)
  float* Lb = L + (size_t)b * n * n;
  __shared__ float s_l[128];
  const int r = threadIdx.x;
  float acc = (r < i) ? k_ni[r] : 0.0f;
  for (int c = 0; c < i; ++c) {
    if (r == c) s_l[c] = acc / Lb[c * n + c];
    __syncthreads();
    if (r > c && r < i) acc -= Lb[r * n + c] * s_l[c];
  }
  __syncthreads();
  if (r < i) Lb[i * n + r] = s_l[r];
  if (threadIdx.x == 0) {
    float ss = 0.0f;
    for (int q = 0; q < i; ++q) ss += s_l[q] * s_l[q];
    float k_ii = signal_var;
    if (has_fixed) k_ii += fixed_var;
    Lb[i * n + i] = sqrtf(k_ii - ss);
    sx[2 * i] = xj[0];
    sx[2 * i + 1] = xj[1];
    for (int q = 0; q < 4; ++q) sE[4 * i + q] = Ej[q];
    sel_idx[(size_t)b * n + i] = j;
  }
}

__global__ void sampler_commit_kernel(SamplerState* __restrict__ st, const float* __restrict__ dom_xy,
                                      const float* __restrict__ dom_E, int d, float* __restrict__ sel_xy,
                                      float* __restrict__ sel_E, long long* __restrict__ sel_idx, float* __restrict__ L,
                                      int n, float signal_var, float fixed_var, int has_fixed, float max_stdev_thresh,
                                      int terminate_early) {
  sampler_commit_body(blockIdx.x, st, dom_xy, dom_E, d, sel_xy, sel_E, sel_idx, L, n, signal_var, fixed_var, has_fixed,
                      max_stdev_thresh, terminate_early);
}

// step B (grid over the domain): k_id against the new anchor, new obs_info row, variance downdate, distance
// mask update, block-level argmax of stdev*mask (first index wins ties) -> per-block candidates.
constexpr int SB_THREADS = 256;
__device__ __forceinline__ void sampler_domain_body(int b, int blk, int nblk, const SamplerState* st,
                                                    const float* dom_xy, const float* dom_E, int d,
                                                    const float* sel_xy, const float* sel_E,
                                                    const float* L, int n, float* obs_info,
                                                    float* var, uint8_t* dist_ok, float signal_var,
                                                    float dist_thresh_sq, float* cand_val,
                                                    int* cand_idx) {
  const volatile SamplerState* vs = st + b;   // re-read every call: the fused small-domain kernel updates it in place
  const int i = vs->count;   // row being appended
  const bool live = !(vs->done || i >= n);
  const int j = blk * SB_THREADS + threadIdx.x;
  float best = -1.0f;
  int besti = 0x7fffffff;
  if (j < d) {
    const size_t o = (size_t)b * d + j;
    float v = var[o];
    uint8_t ok = dist_ok[o];
    if (live) {
      const float* sx = sel_xy + ((size_t)b * n + i) * 2;
      const float* sE = sel_E + ((size_t)b * n + i) * 4;
      const float* xj = dom_xy + o * 2;
      const float* Ej = dom_E + o * 4;
      const float kid = cov_native_f32(sx[0], sx[1], sE[0], sE[1], sE[2], sE[3], xj[0], xj[1], Ej[0], Ej[1], Ej[2], Ej[3],
                                       signal_var);
      const float* Lr = L + (size_t)b * n * n + (size_t)i * n;
      float* ob = obs_info + (size_t)b * n * d;
      float acc = 0.0f;
      for (int r = 0; r < i; ++r) acc += Lr[r] * ob[(size_t)r * d + j];
      const float nv = (kid - acc) / Lr[i];
      ob[(size_t)i * d + j] = nv;
      v -= nv * nv;
      var[o] = v;
      const float ddx = sx[0] - xj[0], ddy = sx[1] - xj[1];
      if (!(ddx * ddx + ddy * ddy > dist_thresh_sq)) ok = 0;
      dist_ok[o] = ok;
    }
    float sd = sqrtf(v);
    if (sd != sd) sd = 0.0f;
    sd += 1e-10f;
    best = sd * (ok ? 1.0f : 0.0f);
    besti = j;
  }
  // block argmax, smallest index on ties
  __shared__ float sv[SB_THREADS];
  __shared__ int si[SB_THREADS];
  sv[threadIdx.x] = best;
  si[threadIdx.x] = besti;
  __syncthreads();
  for (int o = SB_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const float v2 = sv[threadIdx.x + o];
      const int i2 = si[threadIdx.x + o];
      if (v2 > sv[threadIdx.x] || (v2 == sv[threadIdx.x] && i2 < si[threadIdx.x])) {
        sv[threadIdx.x] = v2;
        si[threadIdx.x] = i2;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    cand_val[(size_t)b * nblk + blk] = sv[0];
    cand_idx[(size_t)b * nblk + blk] = si[0];
  }
}

__global__ void __launch_bounds__(SB_THREADS)
sampler_domain_kernel(const SamplerState* __restrict__ st, const float* __restrict__ dom_xy,
                      const float* __restrict__ dom_E, int d, const float* __restrict__ sel_xy,
                      const float* __restrict__ sel_E, const float* __restrict__ L, int n, float* __restrict__ obs_info,
                      float* __restrict__ var, uint8_t* __restrict__ dist_ok, float signal_var, float dist_thresh_sq,
                      float* __restrict__ cand_val, int* __restrict__ cand_idx) {
  sampler_domain_body(blockIdx.y, blockIdx.x, gridDim.x, st, dom_xy, dom_E, d, sel_xy, sel_E, L, n, obs_info, var, dist_ok,
                      signal_var, dist_thresh_sq, cand_val, cand_idx);
}

// step C (1 CTA per batch): final argmax over block candidates -> next index; advance the count.
__device__ __forceinline__ void sampler_pick_body(int b, SamplerState* st, const float* cand_val,
                                                  const int* cand_idx, int nblocks,
                                                  const float* var, int d, int n, int advance) {
  __shared__ float sv[256];
  __shared__ int si[256];
  float best = -1.0f;
  int besti = 0x7fffffff;
  for (int t = threadIdx.x; t < nblocks; t += blockDim.x) {
    const float v = cand_val[(size_t)b * nblocks + t];
    const int ix = cand_idx[(size_t)b * nblocks + t];
    if (v > best || (v == best && ix < besti)) {
      best = v;
      besti = ix;
    }
  }
  sv[threadIdx.x] = best;
  si[threadIdx.x] = besti;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const float v2 = sv[threadIdx.x + o];
      const int i2 = si[threadIdx.x + o];
      if (v2 > sv[threadIdx.x] || (v2 == sv[threadIdx.x] && i2 < si[threadIdx.x])) {
        sv[threadIdx.x] = v2;
        si[threadIdx.x] = i2;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    volatile SamplerState& s = st[b];
    const bool live = !(s.done || s.count >= n);
    if (advance && live) s.count = s.count + 1;
    s.next = si[0];
    // the reference reports gp_stdev at the chosen index (without the distance mask)
    float sd = sqrtf(var[(size_t)b * d + si[0]]);
    if (sd != sd) sd = 0.0f;
    s.next_stdev = sd + 1e-10f;
  }
}

__global__ void sampler_pick_kernel(SamplerState* __restrict__ st, const float* __restrict__ cand_val,
                                    const int* __restrict__ cand_idx, int nblocks, const float* __restrict__ var, int d,
                                    int n, int advance) {
  sampler_pick_body(blockIdx.x, st, cand_val, cand_idx, nblocks, var, d, n, advance);
}

// Small domains (d <= 2048, e.g. the re-selection among <= 64 tracked anchors in track_and_init): the whole greedy
// loop in ONE launch, one CTA per batch element, the same three steps separated by block barriers instead of
// kernel boundaries (3 (n - m) + 2 launches otherwise, each ~10 us of fixed cost for a few hundred flops).
constexpr int SAMPLER_SMALL_D = 8 * SB_THREADS;
__global__ void __launch_bounds__(SB_THREADS)
sampler_small_kernel(SamplerState* st, const float* dom_xy, const float* dom_E, int d, int n, int m, float* sel_xy,
                     float* sel_E, long long* sel_idx, float* L, float* obs_info, float* var, uint8_t* dist_ok,
                     float signal_var, float fixed_var, int has_fixed, float dist_thresh_sq, float max_stdev_thresh,
                     int terminate_early, float* cand_val, int* cand_idx, int* count_out) {
  const int b = blockIdx.x;
  const int nblk = (d + SB_THREADS - 1) / SB_THREADS;
  volatile SamplerState* vs = st + b;
  if (threadIdx.x == 0) {
    vs->count = n;   // "not live": the first domain pass only evaluates the arg max
    vs->next = 0;
    vs->next_stdev = 0.0f;
    vs->done = 0;
  }
  __syncthreads();
  for (int blk = 0; blk < nblk; ++blk) {
    sampler_domain_body(b, blk, nblk, st, dom_xy, dom_E, d, sel_xy, sel_E, L, n, obs_info, var, dist_ok, signal_var,
                        dist_thresh_sq, cand_val, cand_idx);
    __syncthreads();
  }
  sampler_pick_body(b, st, cand_val, cand_idx, nblk, var, d, n, 0);
  __syncthreads();
  if (threadIdx.x == 0) vs->count = m;
  __syncthreads();
  for (int i = m; i < n; ++i) {
    sampler_commit_body(b, st, dom_xy, dom_E, d, sel_xy, sel_E, sel_idx, L, n, signal_var, fixed_var, has_fixed,
                        max_stdev_thresh, terminate_early);
    __syncthreads();
    if (vs->done) break;
    for (int blk = 0; blk < nblk; ++blk) {
      sampler_domain_body(b, blk, nblk, st, dom_xy, dom_E, d, sel_xy, sel_E, L, n, obs_info, var, dist_ok, signal_var,
                          dist_thresh_sq, cand_val, cand_idx);
      __syncthreads();
    }
    sampler_pick_body(b, st, cand_val, cand_idx, nblk, var, d, n, 1);
    __syncthreads();
  }
  if (threadIdx.x == 0) count_out[b] = vs->count;
}

}  // namespace como

using namespace como;

extern "C" int como_b200_cross_covariance(const void* x1, const void* E1, const void* x2, const void* E2, double scale,
                                          int32_t B, int32_t n1, int32_t n2, int32_t elem_bytes, void* out, void* stream) {
  COMO_REQUIRE(x1 && E1 && x2 && E2 && out, "cross_covariance: null pointer argument");
  COMO_REQUIRE(elem_bytes == 4 || elem_bytes == 8, "cross_covariance: only float32/float64 are supported");
  COMO_REQUIRE(B >= 1 && n1 >= 0 && n2 >= 0, "cross_covariance: bad shape");
  if ((long long)n1 * n2 == 0) return COMO_B200_OK;
  const long long total = (long long)n1 * n2;
  dim3 grid((unsigned)((total + 255) / 256), B);
  if (elem_bytes == 4)
    cross_cov_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x1, (const float*)E1, (const float*)x2,
                                                                    (const float*)E2, (float)scale, n1, n2, (float*)out);
  else
    cross_cov_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)x1, (const double*)E1, (const double*)x2,
                                                                     (const double*)E2, scale, n1, n2, (double*)out);
  return check_launch("cross_covariance");
}

extern "C" int como_b200_chol_append(float* L, float* obs_info, float* var, const float* k_ni, const float* k_id,
                                     float k_ii, int32_t B, int32_t n, int32_t d, int32_t N, void* stream) {
  COMO_REQUIRE(L && obs_info && var && k_ni && k_id, "get_new_chol_obs_info: null pointer argument");
  COMO_REQUIRE(N >= 0 && N < n, "get_new_chol_obs_info: row %d out of range for n=%d", N, n);
  cudaStream_t st = (cudaStream_t)stream;
  COMO_REQUIRE(N <= CA_THREADS * CA_ROWS, "get_new_chol_obs_info: at most %d rows", CA_THREADS * CA_ROWS);
  chol_append_kernel<<<B, CA_THREADS, 0, st>>>(L, k_ni, k_ii, n, N);
  obs_info_kernel<<<dim3((d + 255) / 256, B), 256, 0, st>>>(L, obs_info, var, k_id, n, d, N);
  return check_launch("get_new_chol_obs_info");
}

extern "C" size_t como_b200_sampler_workspace_bytes(int32_t B, int32_t d) {
  const size_t nblk = (d + SB_THREADS - 1) / SB_THREADS;
  return 256 + (size_t)B * sizeof(SamplerState) + (size_t)B * nblk * 8 + 256;
}

// Runs greedy steps m..n-1 on the device.  Preconditions (built by the host veneer exactly as
// precalc_entropy_vars does, samplers.py:115-208): sel_xy/sel_E/sel_idx hold the first m anchors,
// L[:m,:m] their Cholesky factor, obs_info[:m] = L^-1 K_md, var = signal_var - sum obs_info^2,
// dist_ok = distance mask against the first m anchors.  count_out[b] = anchors selected in total.
extern "C" int como_b200_sampler_greedy(const float* dom_xy, const float* dom_E, int32_t B, int32_t d, int32_t n, int32_t m,
                                        float* sel_xy, float* sel_E, int64_t* sel_idx, float* L, float* obs_info,
                                        float* var, uint8_t* dist_ok, float signal_var, float fixed_var,
                                        int32_t has_fixed, float dist_thresh, float max_stdev_thresh,
                                        int32_t terminate_early, int32_t* count_out, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  COMO_REQUIRE(dom_xy && dom_E && sel_xy && sel_E && sel_idx && L && obs_info && var && dist_ok && count_out && workspace,
               "sampler_greedy: null pointer argument");
  COMO_REQUIRE(B >= 1 && d >= 1 && n >= 1 && n <= 128 && m >= 1 && m <= n, "sampler_greedy: bad sizes");
  if (workspace_bytes < como_b200_sampler_workspace_bytes(B, d)) {
    set_last_error("sampler_greedy: workspace too small");
    return COMO_B200_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = (d + SB_THREADS - 1) / SB_THREADS;
  unsigned char* ws = (unsigned char*)workspace;
  SamplerState* state = (SamplerState*)ws;
  float* cand_val = (float*)(ws + 256 + (size_t)B * sizeof(SamplerState));
  int* cand_idx = (int*)(cand_val + (size_t)B * nblk);
  if (d <= SAMPLER_SMALL_D) {
    sampler_small_kernel<<<B, SB_THREADS, 0, st>>>(state, dom_xy, dom_E, d, n, m, sel_xy, sel_E, (long long*)sel_idx, L, obs_info,
                                                   var, dist_ok, signal_var, fixed_var, has_fixed, dist_thresh * dist_thresh,
                                                   max_stdev_thresh, terminate_early, cand_val, cand_idx, count_out);
    return check_launch("sampler_greedy");
  }
  // state: count = m, not done; the first domain pass below only evaluates the arg max (row m is not live yet)
  SamplerState h;
  h.count = n;  // "not live": makes the first domain pass a pure argmax evaluation
  h.next = 0;
  h.next_stdev = 0.0f;
  h.done = 0;
  for (int b = 0; b < B; ++b) cudaMemcpyAsync(state + b, &h, sizeof(h), cudaMemcpyHostToDevice, st);
  cudaStreamSynchronize(st);
  sampler_domain_kernel<<<dim3(nblk, B), SB_THREADS, 0, st>>>(state, dom_xy, dom_E, d, sel_xy, sel_E, L, n, obs_info, var,
                                                              dist_ok, signal_var, dist_thresh * dist_thresh, cand_val, cand_idx);
  sampler_pick_kernel<<<B, 256, 0, st>>>(state, cand_val, cand_idx, nblk, var, d, n, 0);
  h.count = m;
  // set count = m without touching `next`: small kernel-free trick -> copy only the first int
  for (int b = 0; b < B; ++b) cudaMemcpyAsync(&state[b].count, &h.count, sizeof(int), cudaMemcpyHostToDevice, st);
  cudaStreamSynchronize(st);
  for (int i = m; i < n; ++i) {
    sampler_commit_kernel<<<B, 128, 0, st>>>(state, dom_xy, dom_E, d, sel_xy, sel_E, (long long*)sel_idx, L, n, signal_var,
                                            fixed_var, has_fixed, max_stdev_thresh, terminate_early);
    sampler_domain_kernel<<<dim3(nblk, B), SB_THREADS, 0, st>>>(state, dom_xy, dom_E, d, sel_xy, sel_E, L, n, obs_info, var,
                                                                dist_ok, signal_var, dist_thresh * dist_thresh, cand_val,
                                                                cand_idx);
    sampler_pick_kernel<<<B, 256, 0, st>>>(state, cand_val, cand_idx, nblk, var, d, n, 1);
  }
  for (int b = 0; b < B; ++b)
    cudaMemcpyAsync(count_out + b, &state[b].count, sizeof(int), cudaMemcpyDeviceToDevice, st);
  return check_launch("sampler_greedy");
}
