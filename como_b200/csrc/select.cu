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

constexpr int SEL_CAP = 4096;   // candidates finished in shared memory after two global digits
struct SelSeg {
  unsigned long long prefix;    // resolved high bits (digits 0,1)
  unsigned rank;                // rank of the wanted value among the candidates
  unsigned cand;                // number of candidates (population of the chosen 22-bit bucket)
  unsigned total;               // valid values in the segment
  unsigned fallback;            // 1: not (yet) finished by the compaction path -> global digit passes continue
  unsigned list_count;          // candidates appended so far
  unsigned digits;              // number of digits resolved globally before the compaction (2 or 3)
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
                               unsigned* s_tmp /* >= SEL_THREADS/32 + 4 */, unsigned* last_count = nullptr) {
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
          s_tmp[SEL_THREADS / 32 + 2] = c[j];
        }
        run += c[j];
      }
    }
    __syncthreads();
    if (all > 0) {
      const unsigned bin = s_tmp[SEL_THREADS / 32 + 0];
      rank = s_tmp[SEL_THREADS / 32 + 1];
      prefix = (K)(prefix | ((K)bin << digit_shift<T>(d)));
      if (last_count) *last_count = s_tmp[SEL_THREADS / 32 + 2];
    } else if (last_count) {
      *last_count = 0;
    }
    __syncthreads();
  }
}

// values: flat array; segment s covers [seg_off[s], seg_off[s+1]).  grid = (chunks, num_segs).
template <typename T>
__global__ void __launch_bounds__(SEL_THREADS)
select_pass_kernel(const T* __restrict__ values, const long long* __restrict__ seg_off, int num_segs, int digit,
                   unsigned* __restrict__ hist, const SelSeg* __restrict__ info = nullptr) {
  using K = typename KeyOf<T>::K;
  __shared__ unsigned s_hist[SEL_BINS];
  __shared__ unsigned s_tmp[SEL_THREADS / 32 + 4];
  const int seg = blockIdx.y;
  const int tid = threadIdx.x;
  if (info != nullptr && info[seg].fallback == 0) return;   // this segment was finished by the compaction path
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
                     long long* __restrict__ count, const SelSeg* __restrict__ info = nullptr) {
  using K = typename KeyOf<T>::K;
  __shared__ unsigned s_tmp[SEL_THREADS / 32 + 4];
  const int seg = blockIdx.x;
  if (info != nullptr && info[seg].fallback == 0) return;
  K prefix;
  unsigned long long rank, total;
  resolve_prefix<T>(hist, num_segs, seg, KeyOf<T>::PASSES, prefix, rank, total, s_tmp);
  if (threadIdx.x == 0) {
    out[seg] = (total > 0) ? (T)(scale * KeyOf<T>::val(prefix)) : (T)NAN;
    if (count) count[seg] = (long long)total;
  }
}

// After digits 0 and 1: gather the values of the chosen 22-bit bucket of every segment (normally a few
// hundred of 300 k) so that the remaining digits never touch global memory again.
template <typename T>
__global__ void __launch_bounds__(SEL_THREADS)
select_compact_kernel(const T* __restrict__ values, const long long* __restrict__ seg_off, int num_segs,
                      const unsigned* __restrict__ hist, SelSeg* __restrict__ info, T* __restrict__ lists, int upto,
                      int first) {
  using K = typename KeyOf<T>::K;
  __shared__ unsigned s_tmp[SEL_THREADS / 32 + 4];
  const int seg = blockIdx.y, tid = threadIdx.x;
  if (!first && info[seg].fallback == 0) return;   // an earlier compaction already finished this segment
  K prefix;
  unsigned long long rank, total;
  unsigned cand = 0;
  resolve_prefix<T>(hist, num_segs, seg, upto, prefix, rank, total, s_tmp, &cand);
  const bool fb = cand > SEL_CAP;
  // every CTA of the segment reaches the same verdict; CTA 0 records it (read by LATER launches only)
  if (blockIdx.x == 0 && tid == 0) {
    info[seg].prefix = (unsigned long long)prefix;
    info[seg].rank = (unsigned)rank;
    info[seg].cand = cand;
    info[seg].total = (unsigned)total;
    info[seg].fallback = fb ? 1u : 0u;
    info[seg].digits = (unsigned)upto;
  }
  if (fb || total == 0) return;
  const int hi_shift = digit_shift<T>(upto - 1);
  const long long beg = seg_off[seg], end = seg_off[seg + 1];
  for (long long i = beg + (long long)blockIdx.x * SEL_THREADS + tid; i < end; i += (long long)gridDim.x * SEL_THREADS) {
    const T v = values[i];
    if (v == v) {
      const K key = KeyOf<T>::key(v);
      if ((key >> hi_shift) == (prefix >> hi_shift)) {
        const unsigned pos = atomicAdd(&info[seg].list_count, 1u);
        if (pos < SEL_CAP) lists[(size_t)seg * SEL_CAP + pos] = v;
      }
    }
  }
}

// One CTA per segment: exact rank selection among the gathered candidates (remaining digits in shared memory).
template <typename T>
__global__ void __launch_bounds__(SEL_THREADS)
select_small_kernel(int num_segs, const SelSeg* __restrict__ info, const T* __restrict__ lists, T scale,
                    T* __restrict__ out, long long* __restrict__ count) {
  using K = typename KeyOf<T>::K;
  __shared__ K s_keys[SEL_CAP];
  __shared__ unsigned s_hist[SEL_BINS];
  __shared__ unsigned s_warp[SEL_THREADS / 32];
  __shared__ unsigned s_pick[2];
  const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const SelSeg si = info[seg];
  if (si.fallback) return;
  if (si.total == 0) {
    if (tid == 0) {
      out[seg] = (T)NAN;
      if (count) count[seg] = 0;
    }
    return;
  }
  const int n = (int)si.cand;
  for (int i = tid; i < n; i += SEL_THREADS) s_keys[i] = KeyOf<T>::key(lists[(size_t)seg * SEL_CAP + i]);
  K prefix = (K)si.prefix;
  unsigned rank = si.rank;
  __syncthreads();
  for (int d = (int)si.digits; d < KeyOf<T>::PASSES; ++d) {
    for (int b = tid; b < SEL_BINS; b += SEL_THREADS) s_hist[b] = 0;
    __syncthreads();
    const int sh = digit_shift<T>(d), nb = digit_bits<T>(d), hi = sh + nb;
    const K dmask = (K)((1u << nb) - 1u);
    for (int i = tid; i < n; i += SEL_THREADS) {
      const K key = s_keys[i];
      if ((key >> hi) == (prefix >> hi)) atomicAdd(&s_hist[(unsigned)((key >> sh) & dmask)], 1u);
    }
    __syncthreads();
    constexpr int PER = SEL_BINS / SEL_THREADS;
    unsigned c[PER], local = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      c[j] = s_hist[tid * PER + j];
      local += c[j];
    }
    unsigned incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    unsigned base = 0;
    for (int w = 0; w < wid; ++w) base += s_warp[w];
    incl += base;
    const unsigned excl = incl - local;
    if (rank >= excl && rank < incl) {
      unsigned run = excl;
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        if (rank >= run && rank < run + c[j]) {
          s_pick[0] = tid * PER + j;
          s_pick[1] = rank - run;
        }
        run += c[j];
      }
    }
    __syncthreads();
    prefix = (K)(prefix | ((K)s_pick[0] << sh));
    rank = s_pick[1];
    __syncthreads();
  }
  if (tid == 0) {
    out[seg] = (T)(scale * KeyOf<T>::val(prefix));
    if (count) count[seg] = (long long)si.total;
  }
}


// ------------------------------------------------------------------------------------------------------------
// Median of values spread over several ranks (a BA pair batch whose pairs are sharded over GPUs) with THREE
// exchanges instead of six: digits 0 and 1 through all-reduced histograms (22 bits: sign, exponent, 10 mantissa
// bits -- a 1/1024-octave bucket, normally a few hundred of 2 M values), then every rank compacts ITS candidates of
// that bucket into a fixed-size pack, the packs are all-gathered, and every rank finishes the remaining digits on
// the gathered candidates.  pack layout per (rank, segment): word 0 = local candidate count (u64), words
// 1..SEL_CAP = candidate bit patterns.  A rank whose local count exceeds SEL_CAP sets the overflow flag of the
// result (the caller then runs the six-pass path; e.g. thousands of bit-identical residuals).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SEL_THREADS)
dist_compact_kernel(const double* __restrict__ values, const long long* __restrict__ seg_off, int num_segs,
                    const unsigned* __restrict__ hist, unsigned long long* __restrict__ pack) {
  using K = KeyOf<double>::K;
  __shared__ unsigned s_tmp[SEL_THREADS / 32 + 4];
  const int seg = blockIdx.y, tid = threadIdx.x;
  K prefix;
  unsigned long long rank, total;
  unsigned cand = 0;
  resolve_prefix<double>(hist, num_segs, seg, 2, prefix, rank, total, s_tmp, &cand);
  if (total == 0) return;
  unsigned long long* my = pack + (size_t)seg * (SEL_CAP + 1);
  const int hi_shift = digit_shift<double>(1);
  const long long beg = seg_off[seg], end = seg_off[seg + 1];
  for (long long i = beg + (long long)blockIdx.x * SEL_THREADS + tid; i < end; i += (long long)gridDim.x * SEL_THREADS) {
    const double v = values[i];
    if (v == v) {
      const K key = KeyOf<double>::key(v);
      if ((key >> hi_shift) == (prefix >> hi_shift)) {
        const unsigned long long pos = atomicAdd(my, 1ull);
        if (pos < (unsigned long long)SEL_CAP) my[1 + pos] = key;
      }
    }
  }
}

// One CTA per segment; packs: (world, num_segs, SEL_CAP + 1) words.  The remaining digits run over the gathered
// candidates straight from L2 (up to world * SEL_CAP of them).  flag[seg] = 1 if any rank overflowed its pack.
__global__ void __launch_bounds__(SEL_THREADS)
dist_finish_kernel(const unsigned long long* __restrict__ packs, int world, int num_segs, const unsigned* __restrict__ hist,
                   double scale, double* __restrict__ out, int* __restrict__ flag) {
  using K = KeyOf<double>::K;
  __shared__ unsigned s_hist[SEL_BINS];
  __shared__ unsigned s_tmp[SEL_THREADS / 32 + 4];
  __shared__ unsigned s_warp[SEL_THREADS / 32];
  __shared__ unsigned s_pick[2];
  const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  K prefix;
  unsigned long long rank64, total;
  unsigned cand = 0;
  resolve_prefix<double>(hist, num_segs, seg, 2, prefix, rank64, total, s_tmp, &cand);
  if (total == 0) {
    if (tid == 0) {
      out[seg] = NAN;
      flag[seg] = 0;
    }
    return;
  }
  bool overflow = false;
  for (int r = 0; r < world; ++r)
    overflow = overflow || (__ldcg(packs + ((size_t)r * num_segs + seg) * (SEL_CAP + 1)) > (unsigned long long)SEL_CAP);
  if (overflow) {
    if (tid == 0) {
      out[seg] = NAN;
      flag[seg] = 1;
    }
    return;
  }
  unsigned rank = (unsigned)rank64;
  for (int d = 2; d < KeyOf<double>::PASSES; ++d) {
    for (int b = tid; b < SEL_BINS; b += SEL_THREADS) s_hist[b] = 0;
    __syncthreads();
    const int sh = digit_shift<double>(d), nb = digit_bits<double>(d), hi = sh + nb;
    const K dmask = (K)((1u << nb) - 1u);
    for (int r = 0; r < world; ++r) {
      const unsigned long long* pk = packs + ((size_t)r * num_segs + seg) * (SEL_CAP + 1);
      const int n = (int)__ldcg(pk);
      for (int i = tid; i < n; i += SEL_THREADS) {
        const K key = __ldcg(pk + 1 + i);
        if ((key >> hi) == (prefix >> hi)) atomicAdd(&s_hist[(unsigned)((key >> sh) & dmask)], 1u);
      }
    }
    __syncthreads();
    constexpr int PER = SEL_BINS / SEL_THREADS;
    unsigned c[PER], local = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      c[j] = s_hist[tid * PER + j];
      local += c[j];
    }
    unsigned incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    unsigned base = 0;
    for (int w = 0; w < wid; ++w) base += s_warp[w];
    incl += base;
    const unsigned excl = incl - local;
    if (rank >= excl && rank < incl) {
      unsigned run = excl;
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        if (rank >= run && rank < run + c[j]) {
          s_pick[0] = tid * PER + j;
          s_pick[1] = rank - run;
        }
        run += c[j];
      }
    }
    __syncthreads();
    prefix = (K)(prefix | ((K)s_pick[0] << sh));
    rank = s_pick[1];
    __syncthreads();
  }
  if (tid == 0) {
    out[seg] = scale * KeyOf<double>::val(prefix);
    flag[seg] = 0;
  }
}

template <typename T>
int median_launch(const T* values, const long long* seg_off, int num_segs, long long max_seg_len, T scale, T* out,
                  long long* count, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  const size_t hist_bytes = (size_t)KeyOf<T>::PASSES * num_segs * SEL_BINS * sizeof(unsigned);
  const size_t info_bytes = ((size_t)num_segs * sizeof(SelSeg) + 255) / 256 * 256;
  const size_t need = hist_bytes + info_bytes + (size_t)num_segs * SEL_CAP * sizeof(T);
  if (workspace_bytes < need) {
    set_last_error("median: workspace %zu < required %zu", workspace_bytes, need);
    return COMO_B200_EWORKSPACE;
  }
  unsigned* hist = (unsigned*)workspace;
  SelSeg* info = (SelSeg*)((unsigned char*)workspace + hist_bytes);
  T* lists = (T*)((unsigned char*)workspace + hist_bytes + info_bytes);
  cudaMemsetAsync(workspace, 0, hist_bytes + info_bytes, stream);
  long long chunks = (max_seg_len + (long long)SEL_THREADS * 8 - 1) / ((long long)SEL_THREADS * 8);
  const long long cap = (long long)sm_count() * 8 / (num_segs > 0 ? num_segs : 1);
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  dim3 grid((unsigned)chunks, (unsigned)num_segs);
  // two global digits, then gather the surviving bucket and finish in shared memory; segments whose bucket
  // is too large (> SEL_CAP) fall back to the remaining global passes (those kernels exit at once otherwise)
  select_pass_kernel<T><<<grid, SEL_THREADS, 0, stream>>>(values, seg_off, num_segs, 0, hist);
  select_pass_kernel<T><<<grid, SEL_THREADS, 0, stream>>>(values, seg_off, num_segs, 1, hist);
  select_compact_kernel<T><<<grid, SEL_THREADS, 0, stream>>>(values, seg_off, num_segs, hist, info, lists, 2, 1);
  int next_digit = 2;
  if (KeyOf<T>::PASSES > 3) {
    // narrow value ranges (a fronto-parallel wall: all depths within 0.1 %) overflow a 22-bit bucket: one more
    // global digit for those segments only, then a second compaction attempt
    select_pass_kernel<T><<<grid, SEL_THREADS, 0, stream>>>(values, seg_off, num_segs, 2, hist, info);
    select_compact_kernel<T><<<grid, SEL_THREADS, 0, stream>>>(values, seg_off, num_segs, hist, info, lists, 3, 0);
    next_digit = 3;
  }
  select_small_kernel<T><<<num_segs, SEL_THREADS, 0, stream>>>(num_segs, info, lists, scale, out, count);
  for (int d = next_digit; d < KeyOf<T>::PASSES; ++d)
    select_pass_kernel<T><<<grid, SEL_THREADS, 0, stream>>>(values, seg_off, num_segs, d, hist, info);
  select_finish_kernel<T><<<num_segs, SEL_THREADS, 0, stream>>>(num_segs, hist, scale, out, count, info);
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

extern "C" int32_t como_b200_median_pack_words(void) { return SEL_CAP + 1; }
extern "C" int como_b200_median_dist_compact_f64(const double* values, const int64_t* seg_offsets, int32_t num_segments,
                                                 int64_t max_segment_len, const void* hist, void* pack, void* stream) {
  COMO_REQUIRE(values && seg_offsets && hist && pack && num_segments >= 1, "median_dist_compact_f64: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(pack, 0, (size_t)num_segments * (SEL_CAP + 1) * sizeof(unsigned long long), st);
  long long chunks = (max_segment_len + (long long)SEL_THREADS * 8 - 1) / ((long long)SEL_THREADS * 8);
  const long long cap = (long long)sm_count() * 8 / num_segments;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  dist_compact_kernel<<<dim3((unsigned)chunks, (unsigned)num_segments), SEL_THREADS, 0, st>>>(
      values, (const long long*)seg_offsets, num_segments, (const unsigned*)hist, (unsigned long long*)pack);
  return check_launch("median_dist_compact");
}
extern "C" int como_b200_median_dist_finish_f64(const void* packs, int32_t world, int32_t num_segments, const void* hist,
                                                double scale, double* out, int32_t* overflow, void* stream) {
  COMO_REQUIRE(packs && hist && out && overflow && world >= 1 && num_segments >= 1, "median_dist_finish_f64: bad arguments");
  dist_finish_kernel<<<num_segments, SEL_THREADS, 0, (cudaStream_t)stream>>>((const unsigned long long*)packs, world, num_segments,
                                                                            (const unsigned*)hist, scale, out, overflow);
  return check_launch("median_dist_finish");
}

extern "C" size_t como_b200_median_workspace_bytes(int32_t num_segments, int32_t elem_bytes) {
  const int passes = (elem_bytes == 8) ? 6 : 3;
  const size_t ns = (size_t)(num_segments > 0 ? num_segments : 0);
  return passes * ns * SEL_BINS * sizeof(unsigned) + (ns * sizeof(SelSeg) + 255) / 256 * 256 + ns * SEL_CAP * (size_t)elem_bytes;
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
