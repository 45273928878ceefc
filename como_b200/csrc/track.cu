// Frame-to-keyframe photometric tracking: the whole coarse-to-fine Gauss-Newton loop in ONE
// persistent cooperative launch (no host round trips for the termination tests).
//
// Reference behaviour restated (not translated) from
//   como/odom/frontend/photo_tracking.py:10-42,77-185, como/geometry/camera.py:57-68,
//   como/odom/frontend/photo_utils.py:9-31, como/odom/backend/robust_loss.py:9-16.
//
// Per GN iteration a group of G co-resident CTAs runs five phases separated by four group barriers:
//   1  warp + bilinear gather + residual r (kept in L2-resident scratch), 11-bit radix histogram of |r|
//   2  pick bucket of the lower-median rank, histogram of the next 11 bits of the candidates
//   3  same for the last 9 bits  -> exact median -> sigma_r = 1.4826 med
//   4  Huber weights, J^T W J / J^T W r / error: registers -> warp shuffle -> CTA -> per-CTA partial row
//   5  every CTA sums the partial rows in a fixed order (bitwise identical everywhere), solves the
//      8x8 system (Cholesky), applies T <- T Exp(-d), a -= d6, b -= d7 and evaluates termination.
// HBM traffic per pixel-iteration: P 12 B + I_ref 4 B + J 32 B + target 4 B (the 52 B of BASELINE.md).
#include <cooperative_groups.h>
#include <string.h>

#include "common.cuh"

namespace como {

constexpr int TRK_THREADS = 512;
constexpr int TRK_WARPS = TRK_THREADS / 32;
constexpr int NACC = 45;        // 36 (upper triangle of 8x8) + 8 (gradient) + 1 (robust error)
constexpr int NACC_PAD = 48;
constexpr int HIST_BINS = 2048;  // bits [30:20], [19:9] -> 2048 bins; [8:0] -> 512 bins
constexpr float HUBER_K = 1.345f;

struct TrackCtl {
  unsigned barrier;
  unsigned pad[31];
  unsigned hist[2][3][HIST_BINS];
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct TrackLayout {
  size_t levels_bytes, ctl_off, partials_off, resid_off, per_problem;
};

static TrackLayout track_layout(int max_n, int num_problems, int max_group) {
  TrackLayout L;
  L.levels_bytes = align_up((size_t)num_problems * COMO_B200_MAX_LEVELS * sizeof(como_b200_track_level_t), 256);
  L.ctl_off = 0;
  L.partials_off = align_up(sizeof(TrackCtl), 256);
  L.resid_off = L.partials_off + align_up((size_t)max_group * NACC_PAD * sizeof(double), 256);
  L.per_problem = L.resid_off + align_up((size_t)max_n * sizeof(float), 256);
  return L;
}

// Find the bin holding 0-based rank k in a global histogram (read through L2), block-wide.
// Returns (via shared) the bin, the rank inside the bin, and the total count.
__device__ __forceinline__ void select_bin(const unsigned* __restrict__ gh, int nbins, unsigned k,
                                           unsigned* s_warp, unsigned* s_out) {
  const int tid = threadIdx.x;
  const int per = nbins / TRK_THREADS;  // 4 or 1
  unsigned c[4] = {0, 0, 0, 0};
  unsigned local = 0;
  for (int j = 0; j < per; ++j) {
    c[j] = __ldcg(gh + tid * per + j);
    local += c[j];
  }
  // inclusive scan over threads
  unsigned incl = local;
  const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  unsigned base = 0;
  for (int w2 = 0; w2 < wid; ++w2) base += s_warp[w2];
  incl += base;
  const unsigned excl = incl - local;
  if (tid == TRK_THREADS - 1) s_out[2] = incl;  // total
  if (k >= excl && k < incl) {
    unsigned run = excl;
    for (int j = 0; j < per; ++j) {
      if (k < run + c[j]) {
        s_out[0] = tid * per + j;
        s_out[1] = k - run;
        break;
      }
      run += c[j];
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void flush_hist(unsigned* s_hist, unsigned* gh, int nbins) {
  __syncthreads();
  for (int b = threadIdx.x; b < nbins; b += TRK_THREADS) {
    const unsigned v = s_hist[b];
    if (v) atomicAdd(gh + b, v);
    s_hist[b] = 0;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(TRK_THREADS, 1)
track_pyr_kernel(const como_b200_track_level_t* __restrict__ levels_all, int num_levels,
                 como_b200_track_term_t term, float* __restrict__ T_io, float* __restrict__ aff_io,
                 float* __restrict__ stats, int* __restrict__ num_iters, uint8_t* __restrict__ ws,
                 TrackLayout lay) {
  const int G = gridDim.x;
  const int c = blockIdx.x;
  const int prob = blockIdx.y;
  const int tid = threadIdx.x;
  const int lane = tid & 31, wid = tid >> 5;

  uint8_t* base = ws + lay.levels_bytes + (size_t)prob * lay.per_problem;
  TrackCtl* ctl = reinterpret_cast<TrackCtl*>(base + lay.ctl_off);
  double* partials = reinterpret_cast<double*>(base + lay.partials_off);
  float* resid = reinterpret_cast<float*>(base + lay.resid_off);
  const como_b200_track_level_t* levels = levels_all + (size_t)prob * COMO_B200_MAX_LEVELS;

  __shared__ unsigned s_hist[HIST_BINS];
  __shared__ double s_red[TRK_WARPS][NACC_PAD];
  __shared__ double s_acc[NACC_PAD];
  __shared__ float s_T[16];
  __shared__ float s_aff[2];
  __shared__ float s_Pm[12];
  __shared__ float s_ea;
  __shared__ unsigned s_warp[TRK_WARPS];
  __shared__ unsigned s_sel[3];
  __shared__ int s_done;
  __shared__ double s_chol[64];
  __shared__ double s_delta[8];

  for (int b = tid; b < HIST_BINS; b += TRK_THREADS) s_hist[b] = 0;
  if (tid < 16) s_T[tid] = T_io[prob * 16 + tid];
  if (tid < 2) s_aff[tid] = aff_io[prob * 2 + tid];
  __syncthreads();

  unsigned epoch = 0;
  int total_iter = 0;
  const int stats_cap = num_levels * term.max_iter;

  for (int l = 0; l < num_levels; ++l) {
    const como_b200_track_level_t lv = levels[l];
    const int N = lv.n;
    const int w = lv.w, h = lv.h;
    int chunk = (N + G - 1) / G;
    chunk = (chunk + 31) & ~31;
    const int i_begin = min(N, c * chunk);
    const int i_end = min(N, i_begin + chunk);
    const float Ax = 1.0f / (float)w, Ay = 1.0f / (float)h;
    const float wf = (float)w, hf = (float)h;
    const float xmax = (float)(w - 1), ymax = (float)(h - 1);

    double mse_prev = INFINITY;
    int it = 0;
    bool level_done = (N <= 0);
    while (!level_done) {
      const int par = total_iter & 1;
      unsigned* gh0 = ctl->hist[par][0];
      unsigned* gh1 = ctl->hist[par][1];
      unsigned* gh2 = ctl->hist[par][2];

      // ---- per-iteration constants: Pmat = K * T[0:3,:], e^{-a}
      if (tid < 12) {
        const int r = tid / 4, cc = tid % 4;
        s_Pm[tid] = lv.K[r * 3 + 0] * s_T[0 * 4 + cc] + lv.K[r * 3 + 1] * s_T[1 * 4 + cc] +
                    lv.K[r * 3 + 2] * s_T[2 * 4 + cc];
      }
      if (tid == 12) s_ea = expf(-s_aff[0]);
      __syncthreads();
      const float p00 = s_Pm[0], p01 = s_Pm[1], p02 = s_Pm[2], p03 = s_Pm[3];
      const float p10 = s_Pm[4], p11 = s_Pm[5], p12 = s_Pm[6], p13 = s_Pm[7];
      const float p20 = s_Pm[8], p21 = s_Pm[9], p22 = s_Pm[10], p23 = s_Pm[11];
      const float ea = s_ea, bb = s_aff[1];

      // ---- phase 1: warp, gather, residual, first radix histogram.  Four pixels per thread and trip so that
      // the operand loads, and then the 16 bilinear taps, are all in flight together (memory-level parallelism).
      constexpr int PB = 4;
      for (int i0 = i_begin + tid; i0 < i_end; i0 += TRK_THREADS * PB) {
        float X[PB], Y[PB], Z[PB], vref[PB];
        bool use[PB];
#pragma unroll
        for (int k = 0; k < PB; ++k) {
          const int i = i0 + k * TRK_THREADS;
          use[k] = (i < i_end) && (lv.mask ? (lv.mask[i] != 0) : true);
          X[k] = Y[k] = 0.0f;
          Z[k] = 1.0f;
          vref[k] = 0.0f;
          if (use[k]) {
            X[k] = lv.P[3 * i + 0];
            Y[k] = lv.P[3 * i + 1];
            Z[k] = lv.P[3 * i + 2];
            vref[k] = lv.vals[i];
          }
        }
        float w00[PB], w01[PB], w10[PB], w11[PB];
        const float* t00[PB];
        int dxs[PB], dys[PB];
        bool valid[PB];
#pragma unroll
        for (int k = 0; k < PB; ++k) {
          const float hx = p00 * X[k] + p01 * Y[k] + p02 * Z[k] + p03;
          const float hy = p10 * X[k] + p11 * Y[k] + p12 * Z[k] + p13;
          const float hz = p20 * X[k] + p21 * Y[k] + p22 * Z[k] + p23;
          const float x = hx / hz, y = hy / hz;
          valid[k] = use[k] && (x >= 1.0f) && (x < xmax) && (y >= 1.0f) && (y < ymax) && (hz > 0.0f);
          // the reference maps pixel coords to [-1,1] and grid_sample maps them back; reproduce
          // that fp32 round trip (coords.py:18-20, grid_sample unnormalize, align_corners=False)
          const float xn = __fsub_rn(__fadd_rn(__fmul_rn(2.0f * Ax, x), Ax), 1.0f);
          const float yn = __fsub_rn(__fadd_rn(__fmul_rn(2.0f * Ay, y), Ay), 1.0f);
          const float ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(xn, 1.0f), wf), 1.0f), 0.5f);
          const float iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(yn, 1.0f), hf), 1.0f), 0.5f);
          const float x0f = floorf(ix), y0f = floorf(iy);
          const float fx0 = ix - x0f, fy0 = iy - y0f;
          const float fx1 = (x0f + 1.0f) - ix, fy1 = (y0f + 1.0f) - iy;
          w00[k] = fx1 * fy1;
          w01[k] = fx0 * fy1;
          w10[k] = fx1 * fy0;
          w11[k] = fx0 * fy0;
          int x0 = valid[k] ? (int)x0f : 0, y0 = valid[k] ? (int)y0f : 0;
          const int xa = min(max(x0, 0), w - 1), xb = min(max(x0 + 1, 0), w - 1);
          const int ya = min(max(y0, 0), h - 1), yb = min(max(y0 + 1, 0), h - 1);
          t00[k] = lv.img + (size_t)ya * w + xa;
          dxs[k] = xb - xa;
          dys[k] = (yb - ya) * w;
        }
        float v00[PB], v01[PB], v10[PB], v11[PB];
#pragma unroll
        for (int k = 0; k < PB; ++k) {
          v00[k] = v01[k] = v10[k] = v11[k] = 0.0f;
          if (valid[k]) {
            v00[k] = __ldg(t00[k]);
            v01[k] = __ldg(t00[k] + dxs[k]);
            v10[k] = __ldg(t00[k] + dys[k]);
            v11[k] = __ldg(t00[k] + dys[k] + dxs[k]);
          }
        }
#pragma unroll
        for (int k = 0; k < PB; ++k) {
          const int i = i0 + k * TRK_THREADS;
          if (i < i_end) {
            float r = __int_as_float(0x7fc00000);
            if (valid[k]) {
              float v = v00[k] * w00[k];
              v += v01[k] * w01[k];
              v += v10[k] * w10[k];
              v += v11[k] * w11[k];
              const float tmp = ea * v;
              r = (tmp + bb) - vref[k];
              const unsigned key = __float_as_uint(fabsf(r));
              atomicAdd(&s_hist[key >> 20], 1u);
            }
            resid[i] = r;
          }
        }
      }
      flush_hist(s_hist, gh0, HIST_BINS);
      group_barrier(&ctl->barrier, epoch, G);

      // ---- phase 2
      // rank of the lower median among nvalid values: (nvalid-1)/2  (torch.median semantics)
      select_bin(gh0, HIST_BINS, 0xffffffffu, s_warp, s_sel);  // first call only to get the total
      const unsigned nvalid = s_sel[2];
      __syncthreads();
      unsigned key_prefix = 0;
      float sigma = __int_as_float(0x7fc00000);
      if (nvalid > 0) {
        const unsigned k0 = (nvalid - 1) / 2;
        select_bin(gh0, HIST_BINS, k0, s_warp, s_sel);
        const unsigned b0 = s_sel[0], k1 = s_sel[1];
        __syncthreads();
        for (int i = i_begin + tid; i < i_end; i += TRK_THREADS) {
          const float r = resid[i];
          if (r == r) {
            const unsigned key = __float_as_uint(fabsf(r));
            if ((key >> 20) == b0) atomicAdd(&s_hist[(key >> 9) & 2047u], 1u);
          }
        }
        flush_hist(s_hist, gh1, HIST_BINS);
        group_barrier(&ctl->barrier, epoch, G);
        // ---- phase 3
        select_bin(gh1, HIST_BINS, k1, s_warp, s_sel);
        const unsigned b1 = s_sel[0], k2 = s_sel[1];
        __syncthreads();
        const unsigned pre = (b0 << 11) | b1;
        for (int i = i_begin + tid; i < i_end; i += TRK_THREADS) {
          const float r = resid[i];
          if (r == r) {
            const unsigned key = __float_as_uint(fabsf(r));
            if ((key >> 9) == pre) atomicAdd(&s_hist[key & 511u], 1u);
          }
        }
        flush_hist(s_hist, gh2, 512);
        group_barrier(&ctl->barrier, epoch, G);
        select_bin(gh2, 512, k2, s_warp, s_sel);
        key_prefix = (pre << 9) | s_sel[0];
        __syncthreads();
        sigma = 1.4826f * __uint_as_float(key_prefix);
      } else {
        group_barrier(&ctl->barrier, epoch, G);
        group_barrier(&ctl->barrier, epoch, G);
      }

      // ---- phase 4: robust weights + normal equations
      float acc[NACC];
#pragma unroll
      for (int k = 0; k < NACC; ++k) acc[k] = 0.0f;
      const float inv_sigma = 1.0f / sigma;
      for (int i0 = i_begin + tid; i0 < i_end; i0 += 2 * TRK_THREADS) {
        float rr[2], vv[2];
        float4 ja[2], jb[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int i = i0 + k * TRK_THREADS;
          rr[k] = __int_as_float(0x7fc00000);
          vv[k] = 0.0f;
          ja[k] = jb[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < i_end) {
            rr[k] = resid[i];
            vv[k] = lv.vals[i];
            ja[k] = *reinterpret_cast<const float4*>(lv.J + 8 * (size_t)i);
            jb[k] = *reinterpret_cast<const float4*>(lv.J + 8 * (size_t)i + 4);
          }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const float r = rr[k];
          if (r == r) {
            const float tmp = (r - bb) + vv[k];
            float j[8] = {ja[k].x, ja[k].y, ja[k].z, ja[k].w, jb[k].x, jb[k].y, -tmp, 1.0f};
            const float wr = r * inv_sigma;
            const float a = fabsf(wr);
            const float wgt = (a < HUBER_K) ? 1.0f : HUBER_K / a;
            acc[44] += wgt * wr * wr;
            int q = 0;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const float wj = wgt * j[kk];
              acc[36 + kk] += wj * r;
#pragma unroll
              for (int m = kk; m < 8; ++m) acc[q++] += wj * j[m];
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < NACC; ++k) {
        const float s = warp_sum(acc[k]);
        if (lane == 0) s_red[wid][k] = (double)s;
      }
      __syncthreads();
      if (tid < NACC) {
        double s = 0.0;
#pragma unroll
        for (int w2 = 0; w2 < TRK_WARPS; ++w2) s += s_red[w2][tid];
        __stcg(partials + (size_t)c * NACC_PAD + tid, s);
      }
      // zero the other parity's histograms for the next iteration (nobody reads them any more)
      {
        unsigned* nz = &ctl->hist[par ^ 1][0][0];
        for (int b = c * TRK_THREADS + tid; b < 3 * HIST_BINS; b += G * TRK_THREADS) nz[b] = 0u;
      }
      __threadfence();
      group_barrier(&ctl->barrier, epoch, G);

      // ---- phase 5: deterministic cross-CTA sum, solve, update, termination (identical in every CTA)
      if (tid < NACC_PAD * 8) {
        const int k = tid >> 3, s8 = tid & 7;
        double s = 0.0;
        if (k < NACC)
          for (int cc = s8; cc < G; cc += 8) s += __ldcg(partials + (size_t)cc * NACC_PAD + k);
        // fixed-order combine of the 8 strided partial sums
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (s8 == 0 && k < NACC) s_acc[k] = s;
      }
      __syncthreads();
      if (wid == 0) {
        // 8x8 solve by one warp: lane i < 8 owns row i (packed upper triangle -> full row)
        double row[8];
        const int li = lane & 7;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const int a = li < m ? li : m, b = li < m ? m : li;
          row[m] = s_acc[a * 8 - (a * (a - 1)) / 2 + (b - a)];
        }
        if (lane >= 8) {
#pragma unroll
          for (int m = 0; m < 8; ++m) row[m] = (m == li) ? 1.0 : 0.0;  // harmless identity rows
        }
        const double gi = s_acc[36 + li];
        const double di = warp_chol_solve8(row, lane < 8 ? gi : 0.0, s_chol);
        double dn2 = (lane < 8) ? di * di : 0.0, gn2 = (lane < 8) ? gi * gi : 0.0;
        dn2 = warp_sum(dn2);
        gn2 = warp_sum(gn2);
        if (lane < 8) s_delta[lane] = di;
        __syncwarp();
        if (lane == 0) {
          const double mse = s_acc[44] / (double)nvalid;
          // T <- T * Exp(-delta[0:6]); COMO tangent [omega, v] -> lietorch [tau=v, phi=omega]
          const double tau[3] = {-s_delta[3], -s_delta[4], -s_delta[5]};
          const double phi[3] = {-s_delta[0], -s_delta[1], -s_delta[2]};
          double E[16];
          se3_exp_tau_phi(tau, phi, E);
          float Tn[16];
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              double sacc = 0.0;
#pragma unroll
              for (int k = 0; k < 4; ++k) sacc += (double)s_T[r * 4 + k] * E[k * 4 + cc];
              Tn[r * 4 + cc] = (float)sacc;
            }
#pragma unroll
          for (int k = 0; k < 16; ++k) s_T[k] = Tn[k];
          s_aff[0] = (float)((double)s_aff[0] - s_delta[6]);
          s_aff[1] = (float)((double)s_aff[1] - s_delta[7]);
          const double dn = sqrt(dn2), gnorm = sqrt(gn2);
          const double rel = fabs((mse_prev - mse) / mse_prev);  // NaN on the first iteration -> false
          const bool done = (it + 1 >= term.max_iter) || (dn < (double)term.delta_norm) ||
                            (rel < (double)term.rel_tol) || (gnorm < (double)term.grad_norm) || (nvalid == 0);
          s_done = done ? 1 : 0;
          s_acc[45] = mse;
          if (c == 0 && stats != nullptr && total_iter < stats_cap) {
            float* st = stats + ((size_t)prob * stats_cap + total_iter) * COMO_B200_TRACK_STAT_STRIDE;
            st[0] = (float)l;
            st[1] = (float)mse;
            st[2] = (float)gnorm;
            st[3] = (float)dn;
            st[4] = sigma;
            st[5] = (float)nvalid;
            st[6] = done ? 1.0f : 0.0f;
            st[7] = 0.0f;
          }
        }
      }
      __syncthreads();
      mse_prev = s_acc[45];
      level_done = (s_done != 0);
      ++it;
      ++total_iter;
      __syncthreads();
    }
  }
  if (c == 0) {
    if (tid < 16) T_io[prob * 16 + tid] = s_T[tid];
    if (tid < 2) aff_io[prob * 2 + tid] = s_aff[tid];
    if (tid == 0 && num_iters != nullptr) num_iters[prob] = total_iter;
  }
}

static int track_max_group(int num_problems) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, track_pyr_kernel, TRK_THREADS, 0);
  const int total = per_sm * sm_count();
  return total / (num_problems > 0 ? num_problems : 1);
}

// ---------------------------------------------------------------------------------------------
// precalc_jacobians: dI/dxi = gradI * dpi/dP * [-P^ | I]; cols 6,7 = [I_ref, 1].
// ---------------------------------------------------------------------------------------------
__global__ void precalc_jac_kernel(const float* __restrict__ grads, const float* __restrict__ P,
                                   const float* __restrict__ vals, float fx, float fy, int64_t n,
                                   float* __restrict__ J) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gx = grads[2 * i], gy = grads[2 * i + 1];
  const float X = P[3 * i], Y = P[3 * i + 1], Z = P[3 * i + 2];
  // rows of dpi/dP: [fx, 0, -fx X/Z]/Z and [0, fy, -fy Y/Z]/Z  (camera.py:20-37)
  const float t1 = fx * X / Z, t2 = fy * Y / Z;
  const float d00 = fx / Z, d02 = -t1 / Z, d11 = fy / Z, d12 = -t2 / Z;
  // dpi/dT = dpi/dP * [-P^ | I],  -P^ = [[0,Z,-Y],[-Z,0,X],[Y,-X,0]]
  const float a0 = d02 * Y, a1 = d00 * Z - d02 * X, a2 = -d00 * Y;
  const float b0 = -d11 * Z + d12 * Y, b1 = -d12 * X, b2 = d11 * X;
  float4 o0, o1;
  o0.x = gx * a0 + gy * b0;
  o0.y = gx * a1 + gy * b1;
  o0.z = gx * a2 + gy * b2;
  o0.w = gx * d00;
  o1.x = gy * d11;
  o1.y = gx * d02 + gy * d12;
  o1.z = vals[i];
  o1.w = 1.0f;
  *reinterpret_cast<float4*>(J + 8 * i) = o0;
  *reinterpret_cast<float4*>(J + 8 * i + 4) = o1;
}

}  // namespace como

using namespace como;

extern "C" size_t como_b200_track_workspace_bytes(int32_t max_n, int32_t num_problems) {
  if (max_n < 0 || num_problems <= 0) return 0;
  const TrackLayout L = track_layout(max_n, num_problems, 1024);
  return L.levels_bytes + (size_t)num_problems * L.per_problem;
}

extern "C" int como_b200_track_pyr(const como_b200_track_level_t* levels, int32_t num_levels,
                                   int32_t num_problems, const como_b200_track_term_t* term, float* T,
                                   float* aff, float* stats, int32_t* num_iters, void* workspace,
                                   size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMO_REQUIRE(levels && term && T && aff && workspace, "track_pyr: null pointer argument");
  COMO_REQUIRE(num_levels >= 1 && num_levels <= COMO_B200_MAX_LEVELS, "track_pyr: num_levels %d out of [1,%d]",
               num_levels, COMO_B200_MAX_LEVELS);
  COMO_REQUIRE(num_problems >= 1, "track_pyr: num_problems must be >= 1");
  COMO_REQUIRE(term->max_iter >= 1, "track_pyr: max_iter must be >= 1");
  int max_n = 0;
  for (int p = 0; p < num_problems; ++p)
    for (int l = 0; l < num_levels; ++l) {
      const como_b200_track_level_t& lv = levels[p * num_levels + l];
      COMO_REQUIRE(lv.n >= 0 && lv.w >= 3 && lv.h >= 3, "track_pyr: bad level shape n=%d w=%d h=%d", lv.n, lv.w, lv.h);
      COMO_REQUIRE(lv.n == 0 || (lv.vals && lv.P && lv.J && lv.img), "track_pyr: null level pointer");
      COMO_REQUIRE(((uintptr_t)lv.J & 15) == 0, "track_pyr: J must be 16-byte aligned");
      if (lv.n > max_n) max_n = lv.n;
    }
  const int max_group = track_max_group(num_problems);
  COMO_REQUIRE(max_group >= 1, "track_pyr: %d problems exceed the co-resident CTA capacity", num_problems);
  int G = (max_n + 2 * TRK_THREADS - 1) / (2 * TRK_THREADS);
  if (G < 1) G = 1;
  if (G > max_group) G = max_group;
  const TrackLayout L = track_layout(max_n, num_problems, 1024);
  const size_t need = L.levels_bytes + (size_t)num_problems * L.per_problem;
  if (workspace_bytes < need) {
    set_last_error("track_pyr: workspace %zu < required %zu", workspace_bytes, need);
    return COMO_B200_EWORKSPACE;
  }
  uint8_t* ws = (uint8_t*)workspace;
  // level descriptors -> device (padded to MAX_LEVELS per problem) through a small ring of pinned
  // staging slots, each guarded by an event so the host never blocks on the stream
  {
    constexpr int SLOTS = 8;
    struct Slot {
      como_b200_track_level_t* buf = nullptr;
      size_t cap = 0;
      cudaEvent_t ev = nullptr;
    };
    static thread_local Slot ring[SLOTS];
    static thread_local int next = 0;
    Slot& sl = ring[next];
    next = (next + 1) % SLOTS;
    const size_t cnt = (size_t)num_problems * COMO_B200_MAX_LEVELS;
    if (sl.ev == nullptr) cudaEventCreateWithFlags(&sl.ev, cudaEventDisableTiming);
    else cudaEventSynchronize(sl.ev);
    if (sl.cap < cnt) {
      if (sl.buf) cudaFreeHost(sl.buf);
      if (cudaMallocHost((void**)&sl.buf, cnt * sizeof(como_b200_track_level_t)) != cudaSuccess) {
        sl.buf = nullptr;
        sl.cap = 0;
        set_last_error("track_pyr: pinned staging allocation failed");
        return COMO_B200_ELAUNCH;
      }
      sl.cap = cnt;
    }
    memset(sl.buf, 0, cnt * sizeof(como_b200_track_level_t));
    for (int p = 0; p < num_problems; ++p)
      for (int l = 0; l < num_levels; ++l) sl.buf[p * COMO_B200_MAX_LEVELS + l] = levels[p * num_levels + l];
    cudaMemcpyAsync(ws, sl.buf, cnt * sizeof(como_b200_track_level_t), cudaMemcpyHostToDevice, stream);
    cudaEventRecord(sl.ev, stream);
  }
  for (int p = 0; p < num_problems; ++p)
    cudaMemsetAsync(ws + L.levels_bytes + (size_t)p * L.per_problem, 0, sizeof(TrackCtl), stream);

  const como_b200_track_level_t* d_levels = (const como_b200_track_level_t*)ws;
  como_b200_track_term_t t = *term;
  TrackLayout lay = L;
  void* args[] = {(void*)&d_levels, (void*)&num_levels, (void*)&t,   (void*)&T,  (void*)&aff,
                  (void*)&stats,    (void*)&num_iters,  (void*)&ws,  (void*)&lay};
  dim3 grid(G, num_problems), block(TRK_THREADS);
  cudaError_t e = cudaLaunchCooperativeKernel((void*)track_pyr_kernel, grid, block, args, 0, stream);
  if (e != cudaSuccess) {
    set_last_error("track_pyr: cooperative launch failed: %s", cudaGetErrorString(e));
    return COMO_B200_ELAUNCH;
  }
  return COMO_B200_OK;
}

extern "C" int como_b200_precalc_jacobians(const float* grads, const float* P, const float* vals,
                                           const float* K, int64_t n, float* J, void* stream_) {
  COMO_REQUIRE(grads && P && vals && K && J, "precalc_jacobians: null pointer argument");
  COMO_REQUIRE(n >= 0, "precalc_jacobians: negative n");
  COMO_REQUIRE(((uintptr_t)J & 15) == 0, "precalc_jacobians: J must be 16-byte aligned");
  if (n == 0) return COMO_B200_OK;
  const int threads = 256;
  const int64_t blocks = (n + threads - 1) / threads;
  precalc_jac_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream_>>>(grads, P, vals, K[0], K[4], n, J);
  return check_launch("precalc_jacobians");
}
