// Shared device/host helpers for the como_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/como_b200.h"

namespace como {

void set_last_error(const char* fmt, ...);
int check_launch(const char* what);
int sm_count();

#define COMO_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      como::set_last_error(__VA_ARGS__); \
      return COMO_B200_EINVAL;           \
    }                                    \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Grid-wide barrier for cooperative (co-resident) launches.  `counter` is a monotonically
// increasing arrival count in global memory; the caller tracks `epoch` (number of barriers passed).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void group_barrier(unsigned* counter, unsigned& epoch, unsigned group_size) {
  __syncthreads();
  epoch += 1;
  if (threadIdx.x == 0) {
    red_release_add_u32(counter, 1u);
    const unsigned target = epoch * group_size;
    while (ld_acquire_u32(counter) < target) {
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// SE(3) exponential with lietorch's tangent convention [tau, phi] (translation first), restated
// from its published algorithm (the dependency is not vendored in the reference):
// R = I + A W + B W^2, t = (I + B W + C W^2) tau.  Row-major 4x4 out.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline void se3_exp_tau_phi(const double tau[3], const double phi[3], double T[16]) {
  const double th2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
  double A, B, C;
  if (th2 < 1e-12) {
    A = 1.0 - th2 / 6.0;
    B = 0.5 - th2 / 24.0;
    C = 1.0 / 6.0 - th2 / 120.0;
  } else {
    const double th = sqrt(th2);
    A = sin(th) / th;
    B = (1.0 - cos(th)) / th2;
    C = (th - sin(th)) / (th2 * th);
  }
  const double W[9] = {0, -phi[2], phi[1], phi[2], 0, -phi[0], -phi[1], phi[0], 0};
  double WW[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += W[i * 3 + k] * W[k * 3 + j];
      WW[i * 3 + j] = s;
    }
  for (int i = 0; i < 3; ++i) {
    double t = 0;
    for (int j = 0; j < 3; ++j) {
      const double I = (i == j) ? 1.0 : 0.0;
      T[i * 4 + j] = I + A * W[i * 3 + j] + B * WW[i * 3 + j];
      t += (I + B * W[i * 3 + j] + C * WW[i * 3 + j]) * tau[j];
    }
    T[i * 4 + 3] = t;
  }
  T[12] = T[13] = T[14] = 0.0;
  T[15] = 1.0;
}

// In-place lower Cholesky + solve of an n x n SPD system stored row-major in A (n<=16), rhs b -> x.
// Non-PD input yields NaN (the reference never raises: cholesky_ex(check_errors=False)).
template <int N>
__host__ __device__ inline void chol_solve_small(double* A, double* b) {
  for (int j = 0; j < N; ++j) {
    double d = A[j * N + j];
    for (int k = 0; k < j; ++k) d -= A[j * N + k] * A[j * N + k];
    d = sqrt(d);
    A[j * N + j] = d;
    for (int i = j + 1; i < N; ++i) {
      double s = A[i * N + j];
      for (int k = 0; k < j; ++k) s -= A[i * N + k] * A[j * N + k];
      A[i * N + j] = s / d;
    }
  }
  for (int i = 0; i < N; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= A[i * N + k] * b[k];
    b[i] = s / A[i * N + i];
  }
  for (int i = N - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < N; ++k) s -= A[k * N + i] * b[k];
    b[i] = s / A[i * N + i];
  }
}

// ---------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA unit) helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Consumer release of a TMA-filled stage: the stage may be overwritten by the async proxy as soon as the arrival is
// seen, so every lane's reads of the stage must have been PERFORMED first.  __syncwarp() + arrive is not enough: shared
// loads still in flight in the LSU when lane 0 arrives can lose the race against the refill when the SM is shared
// with another kernel (seen as a handful of wrong rows per launch -- always the last rows a warp reads -- in
// predictor_stream_kernel beside cuBLAS / the median kernels; scripts/stream_stress.py).  The CTA-scope fence makes each
// lane wait for its outstanding loads.
__device__ __forceinline__ void stage_release(unsigned long long* empty_bar) {
  __threadfence_block();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(empty_bar);
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// 1-D bulk async copy global -> shared (TMA unit, SASS UBLKCP), completion counted on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// the same with an L2 cache policy (createpolicy): evict_first for operands that are streamed once per pass, so that they
// do not push L2-resident scratch (e.g. the tracker's residuals) out
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, unsigned bytes, unsigned long long* bar,
                                              unsigned long long policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// Transposed butterfly: every lane holds 32 partial values v[0..31]; afterwards lane l holds the warp-wide
// total of value index l (31 double shuffles instead of 32 x 5).
__device__ __forceinline__ double warp_transpose_sum32(double v[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool hi = (lane & half) != 0;
#pragma unroll
    for (int q = 0; q < half; ++q) {
      const double send = hi ? v[q] : v[q + half];
      const double recv = __shfl_xor_sync(0xffffffffu, send, half);
      v[q] = (hi ? v[q + half] : v[q]) + recv;
    }
  }
  return v[0];
}

// Warp-cooperative 8x8 SPD solve (fp64): lane i < 8 owns row i of the symmetric matrix (r[0..7]) and its
// right-hand side; returns x_i on lane i (valid for lanes 0..7).  Right-looking Cholesky with one rsqrt per
// column instead of a sqrt + 8 divisions, substitutions through shuffles / 64 doubles of shared scratch.
// Must be called by all 32 lanes of one warp; non-PD input gives NaN (never traps).
__device__ __forceinline__ double warp_chol_solve8(double r[8], double rhs, double* s_L /* >= 64 doubles */) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  double inv_diag[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double djj = __shfl_sync(full, r[j], j);
    const double rs = rsqrt(djj);
    inv_diag[j] = rs;              // 1 / L[j][j]
    const double lij = r[j] * rs;  // L[i][j] for lane i >= j (lane j: sqrt(djj))
    r[j] = lij;
#pragma unroll
    for (int k = j + 1; k < 8; ++k) {
      const double lkj = __shfl_sync(full, lij, k);
      r[k] -= lij * lkj;           // meaningful for lanes i >= k
    }
  }
  if (lane < 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) s_L[lane * 8 + k] = r[k];
  }
  __syncwarp();
  // forward substitution L y = b
  double sacc = rhs, y = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double yj = __shfl_sync(full, sacc * inv_diag[j], j);
    if (lane == j) y = yj;
    sacc -= r[j] * yj;             // lanes i > j use L[i][j]
  }
  // back substitution L^T x = y : x_j = (y_j - sum_{k>j} L[k][j] x_k) / L[j][j]
  double t = y, x = 0.0;
#pragma unroll
  for (int j = 7; j >= 0; --j) {
    const double xj = __shfl_sync(full, t * inv_diag[j], j);
    if (lane == j) x = xj;
    if (lane < j) t -= s_L[j * 8 + lane] * xj;
  }
  return x;
}

}  // namespace como
