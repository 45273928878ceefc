// Exact order statistics on the device: segmented "lower median" (torch.median semantics: element
// (n-1)/2 of the sorted valid values) by most-significant-digit radix select with 11-bit digits.
//
// Used for: the robust scale sigma = 1.4826 * median|r| of a BA pair batch (como/odom/backend/photo.py:124-128),
// the per-keyframe median depth (como/odom/Mapping.py:757-758) and the tracker's reprojected-depth
// median (como/odom/Tracking.py:342-345).  A histogram/approximate median cannot meet the 1e-4 parity
// bound, so the selection is exact: non-negative IEEE values order like their bit patterns.
//
// One launch per digit; every CTA first derives (prefix, rank) of its segment from the previous
// digit's global histogram (all CTAs of a segment compute the same answer), then histograms the next
// digit of the values that match the prefix into shared memory and flushes non-empty bins with atomics.
// NaN values (used as "invalid" markers by the producers) are skipped.
#include "common.cuh"

namespace como {

constexpr int SEL_THREADS = 256;
constexpr int SEL_BINS = 2048;
constexpr int SEL_DIGIT = 11;

template <typename T> struct KeyOf;
template <> struct KeyOf<double> {
  using K = unsigned long long;
  static constexpr int BITS = 64;
  static constexpr int PASSES = 6;  // 6 x 11 = 66 >= 64
  __device__ static K key(double v) { return (K)__double_as_longlong(fabs(v)); }
  __device__ static double val(K k) { return __longlong_as_double((long long)k); }
};
template <> struct KeyOf<float> {
  using K = unsigned;
  static constexpr int BITS = 32;
  static constexpr int PASSES = 3;  // 3 x 11 = 33 >= 32
  __device__ static K key(float v) { return __float_as_uint(fabsf(v)); }
  __device__ static float val(K k) { return __uint_as_float(k); }
};

// digit d (0 = most significant) covers bits [BITS - 11(d+1), BITS - 11 d); the last digit is narrower.
template <typename T>
__device__ __forceinline__ int digit_shift(int d) {
  const int s = KeyOf<T>::BITS - SEL_DIGIT * (d + 1);
  return s < 0 ? 0 : s;
}
template <typename T>
__device__ __forceinline__ int digit_bits(int d) {
  const int s = KeyOf<T>::BITS - SEL_DIGIT * (d + 1);
  return s < 0 ? SEL_DIGIT + s : SEL_DIGIT;
}

// Walk the per-digit histograms of `seg` up to (not including) digit `upto`; returns prefix and rank.
// hist layout: [digit][seg][SEL_BINS].  Block-wide; every thread gets the result.
template <typename T>
__device__ void resolve_prefix(const unsigned* __restrict__ hist, int num_segs, int seg, int upto,
                               typename KeyOf<T>::K& prefix, unsigned long long& rank, unsigned long long& total,
                               unsigned* s_tmp /* >= SEL_THREADS/32 + 4 */) {
  using K = typename KeyOf<T>::K;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int PER = SEL_BINS / SEL_THREADS;  // 8
  prefix = 0;
  rank = 0;
  total = 0;
  for (int d = 0; d < upto; ++d) {
    const unsigned* h = hist + ((size_t)d * num_segs + seg) * SEL_BINS;
    unsigned c[PER];
    unsigned local = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      c[j] = __ldcg(h + tid * PER + j);
      local += c[j];
    }
    unsigned incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_tmp[wid] = incl;
    __syncthreads();
    unsigned base = 0, all = 0;
    for (int w = 0; w < SEL_THREADS / 32; ++w) {
      if (w < wid) base += s_tmp[w];
      all += s_tmp[w];
    }
    incl += base;
    const unsigned excl = incl - local;
    if (d == 0) {
      total = all;
      rank = (all > 0) ? (unsigned long long)((all - 1) / 2) : 0ull;
    }
    __syncthreads();
    const unsigned k = (unsigned)rank;
    if (all > 0 && k >= excl && k < incl) {
      unsigned run = excl;
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        if (k >= run && k < run + c[j]) {
          s_tmp[SEL_THREADS / 32 + 0] = tid * PER + j;
          s_tmp[SEL_THREADS / 32 + 1] = k - run;
        }
        run += c[j];
      }
    }
    __syncthreads();
    if (all > 0) {
      const unsigned bin = s_tmp[SEL_THREADS / 32 + 0];
      rank = s_tmp[SEL_THREADS / 32 + 1];
      prefix = (K)(prefix | ((K)bin << digit_shift<T>(d)));
    }
    __syncthreads();
  }
}

// values: flat array; segment s covers [seg_off[s], seg_off[s+1]).  grid = (chunks, num_segs).
template <typename T>
__global__ void __launch_bounds__(SEL_THREADS)
select_pass_kernel(const T* __restrict__ values, const long long* __restrict__ seg_off, int num_segs, int digit,
                   unsigned* __restrict__ hist) {
  using K = typename KeyOf<T>::K;
  __shared__ unsigned s_hist[SEL_BINS];
  __shared__ unsigned s_tmp[SEL_THREADS / 32 + 4];
  const int seg = blockIdx.y;
  const int tid = threadIdx.x;
  for (int b = tid; b < SEL_BINS; b += SEL_THREADS) s_hist[b] = 0;
  K prefix;
  unsigned long long rank, total;
  resolve_prefix<T>(hist, num_segs, seg, digit, prefix, rank, total, s_tmp);
  __syncthreads();
  const long long beg = seg_off[seg], end = seg_off[seg + 1];
  const int sh = digit_shift<T>(digit);
  const int nb = digit_bits<T>(digit);
  const K dmask = (K)((1u << nb) - 1u);
  // bits above the current digit must equal the prefix
  const int hi_shift = sh + nb;
  for (long long i = beg + (long long)blockIdx.x * SEL_THREADS + tid; i < end; i += (long long)gridDim.x * SEL_THREADS) {
    const T v = values[i];
    if (v == v) {
      const K key = KeyOf<T>::key(v);
      const bool match = (digit == 0) || ((hi_shift >= KeyOf<T>::BITS) ? true : ((key >> hi_shift) == (prefix >> hi_shift)));
      if (match) atomicAdd(&s_hist[(unsigned)((key >> sh) & dmask)], 1u);
    }
  }
  __syncthreads();
  unsigned* h = hist + ((size_t)digit * num_segs + seg) * SEL_BINS;
  for (int b = tid; b < SEL_BINS; b += SEL_THREADS) {
    const unsigned v = s_hist[b];
    if (v) atomicAdd(h + b, v);
  }
}

// out[seg] = scale * median (NaN for an empty segment); count[seg] = number of valid values.
template <typename T>
__global__ void __launch_bounds__(SEL_THREADS)
select_finish_kernel(int num_segs, const unsigned* __restrict__ hist, T scale, T* __restrict__ out,
                     long long* __restrict__ count) {
  using K = typename KeyOf<T>::K;
  __shared__ unsigned s_tmp[SEL_THREADS / 32 + 4];
  const int seg = blockIdx.x;
  K prefix;
  unsigned long long rank, total;
  resolve_prefix<T>(hist, num_segs, seg, KeyOf<T>::PASSES, prefix, rank, total, s_tmp);
  if (threadIdx.x == 0) {
    out[seg] = (total > 0) ? (T)(scale * KeyOf<T>::val(prefix)) : (T)NAN;
    if (count) count[seg] = (long long)total;
  }
}

template <typename T>
int median_launch(const T* values, const long long* seg_off, int num_segs, long long max_seg_len, T scale, T* out,
                  long long* count, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  const size_t need = (size_t)KeyOf<T>::PASSES * num_segs * SEL_BINS * sizeof(unsigned);
  if (workspace_bytes < need) {
    set_last_error("median: workspace %zu < required %zu", workspace_bytes, need);
    return COMO_B200_EWORKSPACE;
  }
  unsigned* hist = (unsigned*)workspace;
  cudaMemsetAsync(hist, 0, need, stream);
  long long chunks = (max_seg_len + (long long)SEL_THREADS * 8 - 1) / ((long long)SEL_THREADS * 8);
  const long long cap = (long long)sm_count() * 8 / (num_segs > 0 ? num_segs : 1);
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  dim3 grid((unsigned)chunks, (unsigned)num_segs);
  for (int d = 0; d < KeyOf<T>::PASSES; ++d)
    select_pass_kernel<T><<<grid, SEL_THREADS, 0, stream>>>(values, seg_off, num_segs, d, hist);
  select_finish_kernel<T><<<num_segs, SEL_THREADS, 0, stream>>>(num_segs, hist, scale, out, count);
  return check_launch("median");
}

template int median_launch<double>(const double*, const long long*, int, long long, double, double*, long long*,
                                   void*, size_t, cudaStream_t);
template int median_launch<float>(const float*, const long long*, int, long long, float, float*, long long*, void*,
                                  size_t, cudaStream_t);

}  // namespace como

using namespace como;

template <typename T>
static int median_pass(const T* values, const long long* seg_off, int num_segs, long long max_seg_len, int digit,
                       unsigned* hist, cudaStream_t stream) {
  long long chunks = (max_seg_len + (long long)SEL_THREADS * 8 - 1) / ((long long)SEL_THREADS * 8);
  const long long cap = (long long)sm_count() * 8 / (num_segs > 0 ? num_segs : 1);
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  select_pass_kernel<T><<<dim3((unsigned)chunks, (unsigned)num_segs), SEL_THREADS, 0, stream>>>(values, seg_off, num_segs, digit, hist);
  return check_launch("median_pass");
}

// Distributed use: zero `hist` (passes x segments x 2048 uint32), then for digit = 0..passes-1 call the pass and
// all-reduce (sum) hist[digit] across ranks before the next one; finish turns the histograms into the value.
extern "C" int32_t como_b200_median_num_passes(int32_t elem_bytes) { return elem_bytes == 8 ? 6 : 3; }
extern "C" int como_b200_median_pass_f64(const double* values, const int64_t* seg_offsets, int32_t num_segments,
                                         int64_t max_segment_len, int32_t digit, void* hist, void* stream) {
  COMO_REQUIRE(values && seg_offsets && hist, "median_pass_f64: null pointer argument");
  COMO_REQUIRE(digit >= 0 && digit < 6 && num_segments >= 1, "median_pass_f64: bad digit/segments");
  return median_pass<double>(values, (const long long*)seg_offsets, num_segments, max_segment_len, digit, (unsigned*)hist,
                             (cudaStream_t)stream);
}
extern "C" int como_b200_median_finish_f64(int32_t num_segments, const void* hist, double scale, double* out, int64_t* count,
                                           void* stream) {
  COMO_REQUIRE(hist && out && num_segments >= 1, "median_finish_f64: bad arguments");
  select_finish_kernel<double><<<num_segments, SEL_THREADS, 0, (cudaStream_t)stream>>>(num_segments, (const unsigned*)hist, scale,
                                                                                      out, (long long*)count);
  return check_launch("median_finish");
}

extern "C" size_t como_b200_median_workspace_bytes(int32_t num_segments, int32_t elem_bytes) {
  const int passes = (elem_bytes == 8) ? 6 : 3;
  return (size_t)passes * (num_segments > 0 ? num_segments : 0) * SEL_BINS * sizeof(unsigned);
}

extern "C" int como_b200_median_f64(const double* values, const int64_t* seg_offsets, int32_t num_segments,
                                    int64_t max_segment_len, double scale, double* out, int64_t* count,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  COMO_REQUIRE(values && seg_offsets && out && workspace, "median_f64: null pointer argument");
  COMO_REQUIRE(num_segments >= 1, "median_f64: num_segments must be >= 1");
  return median_launch<double>(values, (const long long*)seg_offsets, num_segments, max_segment_len, scale, out,
                               (long long*)count, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int como_b200_median_f32(const float* values, const int64_t* seg_offsets, int32_t num_segments,
                                    int64_t max_segment_len, float scale, float* out, int64_t* count,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  COMO_REQUIRE(values && seg_offsets && out && workspace, "median_f32: null pointer argument");
  COMO_REQUIRE(num_segments >= 1, "median_f32: num_segments must be >= 1");
  return median_launch<float>(values, (const long long*)seg_offsets, num_segments, max_segment_len, scale, out,
                              (long long*)count, workspace, workspace_bytes, (cudaStream_t)stream);
}
