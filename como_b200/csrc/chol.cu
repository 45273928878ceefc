// Dense SPD solve  H x = g  (fp64) for the window-BA normal equations: replaces lin_sys.solve_system
// (como/odom/backend/linear_system.py:101-112: torch.linalg.cholesky_ex + torch.cholesky_solve, i.e. cuSOLVER
// potrf + two cuBLAS trsv: 1.21 + 0.47 ms at n = 2848 on B200).
//
// One persistent kernel factorises the matrix with a LEFT-LOOKING TILED DATAFLOW Cholesky (64 x 64 tiles):
//   * tiles of the lower triangle are numbered column-major; CTA c owns tiles c, c + grid, ... and processes them
//     in that order.  Tile (i,k) accumulates  sum_{j<k} L_ij L_kj^T  on the FP64 tensor path (DMMA.8x8x4, operands
//     staged in shared memory with cp.async, double buffered), subtracts it from A_ik and then either
//       - i == k: factorises the 64 x 64 block (blocked 8 x 8: warp Cholesky + DMMA trailing update) and inverts it, or
//       - i  > k: multiplies by L_kk^-T (the same DMMA GEMM with the inverse as operand);
//   * a tile waits only for the tiles it reads (release/acquire flags in global memory); every dependency points
//     to a smaller tile number, so with all CTAs resident (grid <= #SMs, 1 CTA per SM) the schedule cannot
//     deadlock and look-ahead happens by itself: the next diagonal block is factorised while the bulk of the
//     trailing updates of earlier columns is still running on other SMs;
//   * the right-hand side rides along as one extra block row (row 0 of tiles (nb, k)): its "panel solve" IS the
//     forward substitution  y = L^-1 g.
// A second small kernel does the backward substitution  L^T x = y  column by column (one CTA per block column,
// flags again), using the inverses of the diagonal blocks that the factorisation left behind.
// Non-PD input gives NaN (sqrt of a negative pivot), like cholesky_ex(check_errors=False) -- never a trap.
#include "common.cuh"

namespace como {

constexpr int CT = 64;        // tile size
constexpr int CPITCH = 68;    // shared-memory pitch (== 4 mod 16: conflict-free DMMA fragment loads)
constexpr int CH_THREADS = 256;

__device__ __forceinline__ void dmma884c(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// 16-byte async copy global -> shared through L2 only (.cg): tiles are produced by other SMs during this kernel
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ int ld_acquire_i32(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_i32(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct CholView {
  double* L;        // padded working copy, (64 nb) x (64 nb) row-major: lower triangle of H, identity padding
  double* aug;      // nb tiles of 64 x 64 (row 0 = right-hand side / y)
  double* linv;     // nb tiles of 64 x 64: inverses of the diagonal blocks (for the backward substitution)
  double* l8inv;    // nb x 8 blocks of 8 x 8: inverses of the 8 x 8 diagonal sub-blocks (for the panel solves)
  int* flags;       // (nb + 1) x nb
  int n, nb;
};

__device__ __forceinline__ double* tile_ptr(const CholView& v, int i, int k, int& pitch) {
  if (i == v.nb) {
    pitch = CT;
    return v.aug + (size_t)k * CT * CT;
  }
  pitch = v.nb * CT;
  return v.L + (size_t)i * CT * pitch + (size_t)k * CT;
}

// global tile -> shared (pitch CPITCH); every tile of the padded copy is complete and 16-byte aligned
__device__ __forceinline__ void load_tile_async(const CholView& v, int i, int k, double* dst) {
  int pitch;
  const double* src = tile_ptr(v, i, k, pitch);
  for (int e = threadIdx.x; e < CT * CT / 2; e += CH_THREADS) {
    const int r = e >> 5, c = 2 * (e & 31);
    cp_async16(dst + r * CPITCH + c, src + (size_t)r * pitch + c);
  }
}

// acc (rows 8 warp + g4, cols 8 c + 2 l4 + {0,1}) += A (64 x 64, [row][k]) * B^T (B [col][k]); both in shared memory
__device__ __forceinline__ void tile_gemm(const double* A, const double* B, double (&acc)[8][2]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g4 = lane >> 2, l4 = lane & 3;
  const double* ar = A + (8 * warp + g4) * CPITCH + l4;
  const double* br = B + g4 * CPITCH + l4;
#pragma unroll 4
  for (int ks = 0; ks < CT / 4; ++ks) {
    const double fa = ar[4 * ks];
#pragma unroll
    for (int c = 0; c < 8; ++c) dmma884c(acc[c][0], acc[c][1], fa, br[(8 * c) * CPITCH + 4 * ks]);
  }
}

__device__ unsigned long long* g_chol_timeline = nullptr;   // optional: 4 timestamps (ns) per tile
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct CholSmem {
  double A[2][CT * CPITCH];
  double B[2][CT * CPITCH];
  double T[CT * CPITCH];
  double X[CT * CPITCH];       // inverse of the diagonal block (built here / operand of the panel solve)
  double L8inv[8][64];         // scratch
  int2 cur;
};

// Packed lower triangle of an 8 x 8 block: element (i, k), k <= i, at i (i + 1) / 2 + k.
__device__ __forceinline__ constexpr int tri8(int i, int k) { return i * (i + 1) / 2 + k; }

// Inverse of a lower-triangular 8 x 8 block held in registers: x[i][j] = -rs_i sum_{m=j}^{i-1} l[i][m] x[m][j]
__device__ __forceinline__ void inv8_regs(const double (&l)[36], const double (&rs)[8], double (&x)[36]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    x[tri8(j, j)] = rs[j];
#pragma unroll
    for (int i = j + 1; i < 8; ++i) {
      double sacc = 0.0;
#pragma unroll
      for (int m = j; m < i; ++m) sacc += l[tri8(i, m)] * x[tri8(m, j)];
      x[tri8(i, j)] = -rs[i] * sacc;
    }
  }
}

// C (+)= sign * A * B for small matrices in shared memory (pitch CPITCH) on the tensor path.  A: M x K row-major
// ([row][k]), B: K x N row-major ([k][col]), C: M x N.  M, N multiples of 8, K multiple of 4.  Tiles are dealt to the
// 8 warps round-robin starting at `first`; all 256 threads must call it.
__device__ __forceinline__ void small_gemm(const double* A, const double* B, double* Cm, int M, int N, int K, double sign,
                                           int first) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g4 = lane >> 2, l4 = lane & 3;
  const int tn = N / 8, nt = (M / 8) * tn;
  for (int t = (warp + 8 - (first & 7)) & 7; t < nt; t += 8) {
    const int tr = t / tn, tc = t % tn;
    double c0 = 0.0, c1 = 0.0;
    const double* ar = A + (8 * tr + g4) * CPITCH + l4;
    const double* br = B + l4 * CPITCH + 8 * tc + g4;
    for (int k0 = 0; k0 < K; k0 += 4) dmma884c(c0, c1, ar[k0], br[k0 * CPITCH]);
    double* cp = Cm + (8 * tr + g4) * CPITCH + 8 * tc + 2 * l4;
    cp[0] = sign * c0;
    cp[1] = sign * c1;
  }
}

// ---- 64 x 64 Cholesky of S.T (lower triangle in place).  256 threads.
// Blocked by 8: (a, b) warps 0-1 factorise the 8 x 8 diagonal block in registers (every lane the same block: no shuffles on the
// pivot chain; ~140 cycles per pivot = one fp64 rsqrt -- MUFU seed + two Newton steps of dependent DFMAs at ~17 cycles
// each -- + the dependent multiply / update; taking pivots in pairs kept the same chain length) and solve the
// rows below it in the same sweep, one row per lane, while warp 2 inverts the PREVIOUS diagonal block (needed by the
// panel solves of other tiles, not by this loop), (c) the trailing block is updated on the tensor path.
// Measured per 8-column step (scripts/chol_probe.py): (a, b) 1.34 k cycles, (c) 0.28 k; the separate substitution
// phase this replaced cost about as much as the 0.2 k it adds to the sweep (896 -> 893 us for n = 2848).  Building
// the block inverse inside the same sweep instead of on warp 2 was tried as well: the extra dependent-free DFMAs are
// NOT hidden by the pivot latency (1.87 k cycles per step, 983 us) -- the fp64 pipe of one SM sub-partition is the
// limit, not the chain alone.  On return S.L8inv[b] holds the inverse of diagonal block b (8 x 8 row-major, zero above the
// diagonal) and s_invdiag the reciprocal pivots.
__device__ void potrf64(CholSmem& S, double* s_invdiag, long long* clk = nullptr) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g4 = lane >> 2, l4 = lane & 3;
  double* T = S.T;
  auto invert_block = [&](int blk) {   // whole warp, redundantly in registers
    const int c0 = 8 * blk;
    double l8[36], rs8[8], x8[36];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      rs8[i] = s_invdiag[c0 + i];
#pragma unroll
      for (int k = 0; k <= i; ++k) l8[tri8(i, k)] = T[(c0 + i) * CPITCH + c0 + k];
    }
    inv8_regs(l8, rs8, x8);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int k = 0; k < 8; ++k) S.L8inv[blk][8 * i + k] = (k <= i) ? x8[tri8(i, k <= i ? k : 0)] : 0.0;
  };
  for (int jb = 0; jb < 8; ++jb) {
    const int c0 = 8 * jb;
    if (clk && tid == 0) clk[4 * jb] = clock64();
    // (a)+(b) warps 0 and 1: every lane factorises the 8 x 8 diagonal block redundantly in registers AND carries one
    // row of the panel below it through the same right-looking sweep (l_rj = a_rj / L_jj, a_rk -= l_rj L_kj): the
    // panel solve hangs off the pivot chain instead of following it as a second dependent chain of the same length.
    if (warp < 2) {
      const int r = c0 + 8 + tid;          // this lane's panel row (tid < 64)
      const bool has_row = r < CT;
      double a8[36], rs8[8], pr[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int k = 0; k <= i; ++k) a8[tri8(i, k)] = T[(c0 + i) * CPITCH + c0 + k];
      const double* prow = T + (has_row ? r : c0) * CPITCH + c0;
#pragma unroll
      for (int c = 0; c < 8; ++c) pr[c] = prow[c];
      // both warps have read the block before warp 0 overwrites it with the factor
      asm volatile("bar.sync 1, 64;" ::: "memory");
      if (warp == 0 || has_row) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const double rj = rsqrt(a8[tri8(j, j)]);
          rs8[j] = rj;
#pragma unroll
          for (int i = j; i < 8; ++i) a8[tri8(i, j)] *= rj;        // (j,j): d * rsqrt(d) = sqrt(d)
          pr[j] *= rj;
#pragma unroll
          for (int i = j + 1; i < 8; ++i)
#pragma unroll
            for (int k = j + 1; k <= i; ++k) a8[tri8(i, k)] -= a8[tri8(i, j)] * a8[tri8(k, j)];
#pragma unroll
          for (int k = j + 1; k < 8; ++k) pr[k] -= pr[j] * a8[tri8(k, j)];
        }
        if (warp == 0) {
          // every lane holds the same block and stores it (same address, same value): no divergent selection
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int k = 0; k <= i; ++k) T[(c0 + i) * CPITCH + c0 + k] = a8[tri8(i, k)];
#pragma unroll
          for (int j = 0; j < 8; ++j) s_invdiag[c0 + j] = rs8[j];
        }
        if (has_row) {
#pragma unroll
          for (int c = 0; c < 8; ++c) T[r * CPITCH + c0 + c] = pr[c];
        }
      }
    } else if (warp == 2 && jb > 0) {
      invert_block(jb - 1);   // inverse of the PREVIOUS diagonal block: for the panel solves of other tiles
    }
    __syncthreads();
    if (clk && tid == 0) clk[4 * jb + 1] = clock64();
    if (clk && tid == 0) clk[4 * jb + 2] = clock64();
    // (c) trailing update T22 -= P P^T on the tensor path: warp w owns row tile jb + 1 + w (lower tiles only)
    {
      const int rt = jb + 1 + warp;
      if (rt < 8) {
        const double pa0 = -T[(8 * rt + g4) * CPITCH + c0 + l4], pa1 = -T[(8 * rt + g4) * CPITCH + c0 + 4 + l4];
        for (int ct = jb + 1; ct <= rt; ++ct) {
          double* cptr = T + (8 * rt + g4) * CPITCH + 8 * ct + 2 * l4;
          double c0v = cptr[0], c1v = cptr[1];
          const double pb0 = T[(8 * ct + g4) * CPITCH + c0 + l4], pb1 = T[(8 * ct + g4) * CPITCH + c0 + 4 + l4];
          dmma884c(c0v, c1v, pa0, pb0);
          dmma884c(c0v, c1v, pa1, pb1);
          cptr[0] = c0v;
          cptr[1] = c1v;
        }
      }
    }
    __syncthreads();
    if (clk && tid == 0) clk[4 * jb + 3] = clock64();
  }
  if (warp == 2) invert_block(7);
  __syncthreads();
  if (clk && tid == 0) clk[32] = clock64();
}

// ---- inverse of the factor in S.T into S.X from the 8 x 8 inverses, by recursive doubling on the tensor path:
// X21 = -X22 (L21 X11) for block sizes 8 -> 16 -> 32 (scratch: S.A[0]).  Off the critical path of the factorisation.
__device__ void inverse64(CholSmem& S, long long* clk = nullptr) {
  const int tid = threadIdx.x;
  double* T = S.T;
  double* X = S.X;
  for (int e = tid; e < CT * CT; e += CH_THREADS) {
    const int r = e >> 6, c = e & 63;
    X[r * CPITCH + c] = ((r >> 3) == (c >> 3)) ? S.L8inv[r >> 3][8 * (r & 7) + (c & 7)] : 0.0;
  }
  __syncthreads();
  if (clk && tid == 0) clk[33] = clock64();
  double* tmp = S.A[0];
  for (int bs = 8; bs < CT; bs *= 2) {
    const int npairs = CT / (2 * bs);
    for (int pr = 0; pr < npairs; ++pr) {
      const int base = 2 * bs * pr;
      small_gemm(T + (base + bs) * CPITCH + base, X + base * CPITCH + base, tmp + (base + bs) * CPITCH + base, bs, bs, bs, 1.0,
                 pr * (bs * bs / 64));
    }
    __syncthreads();
    for (int pr = 0; pr < npairs; ++pr) {
      const int base = 2 * bs * pr;
      small_gemm(X + (base + bs) * CPITCH + base + bs, tmp + (base + bs) * CPITCH + base, X + (base + bs) * CPITCH + base, bs, bs,
                 bs, -1.0, pr * (bs * bs / 64));
    }
    __syncthreads();
    if (clk && tid == 0) clk[34 + (bs == 8 ? 0 : bs == 16 ? 1 : 2)] = clock64();
  }
}

// ---- panel solve  X L^T = T  in place (T: S.T, L: lower factor in S.X, 8 x 8 inverses in S.L8inv), blocked
// substitution on the tensor path.  Warp w owns rows 8w..8w+7 and never needs another warp's rows: no block barrier.
//   X[:, cb] = (T[:, cb] - sum_{mb<cb} X[:, mb] L[cb, mb]^T) L8inv_cb^T
__device__ void trsm64(CholSmem& S, double* Tbuf = nullptr) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g4 = lane >> 2, l4 = lane & 3;
  double* Tw = (Tbuf ? Tbuf : S.T) + (8 * warp) * CPITCH;
  const double* L = S.X;
  for (int cb = 0; cb < 8; ++cb) {
    double c0 = 0.0, c1 = 0.0;
    for (int kk = 0; kk < 8 * cb; kk += 4)
      dmma884c(c0, c1, Tw[g4 * CPITCH + kk + l4], L[(8 * cb + g4) * CPITCH + kk + l4]);
    double* cp = Tw + g4 * CPITCH + 8 * cb + 2 * l4;
    cp[0] -= c0;
    cp[1] -= c1;
    __syncwarp();
    double x0 = 0.0, x1 = 0.0;
    const double* li = &S.L8inv[cb][0];
    dmma884c(x0, x1, Tw[g4 * CPITCH + 8 * cb + l4], li[8 * g4 + l4]);
    dmma884c(x0, x1, Tw[g4 * CPITCH + 8 * cb + 4 + l4], li[8 * g4 + 4 + l4]);
    __syncwarp();
    cp[0] = x0;
    cp[1] = x1;
    __syncwarp();
  }
}

// debug probe: factorise one 64 x 64 tile (row-major, pitch 64) and record clock64 stamps of the phases
__global__ void __launch_bounds__(CH_THREADS, 1) chol_probe_kernel(double* tile, long long* clk) {
  extern __shared__ __align__(16) unsigned char craw[];
  CholSmem& S = *reinterpret_cast<CholSmem*>(craw);
  for (int e = threadIdx.x; e < CT * CT; e += CH_THREADS) S.T[(e >> 6) * CPITCH + (e & 63)] = tile[e];
  __syncthreads();
  __shared__ double s_invdiag[CT];
  potrf64(S, s_invdiag, clk);
  inverse64(S, clk);
  for (int e = threadIdx.x; e < CT * CT; e += CH_THREADS) tile[e] = S.X[(e >> 6) * CPITCH + (e & 63)];
}

// padded working copy: lower triangle of H (tiles on or below the diagonal, complete tiles), identity on the padding
__global__ void __launch_bounds__(256)
chol_copy_kernel(CholView v, const double* __restrict__ H) {
  const int i = blockIdx.y, k = blockIdx.x;
  if (k > i) return;
  const int ld = v.nb * CT;
  for (int e = threadIdx.x; e < CT * CT; e += 256) {
    const int r = i * CT + (e >> 6), c = k * CT + (e & 63);
    double a;
    if (r < v.n && c < v.n) a = H[(size_t)r * v.n + c];
    else a = (r == c) ? 1.0 : 0.0;
    v.L[(size_t)r * ld + c] = a;
  }
}

__global__ void chol_prep_kernel(CholView v, const double* __restrict__ g, int2* __restrict__ table) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  for (int e = tid; e < (v.nb + 1) * v.nb; e += stride) v.flags[e] = 0;
  for (int e = tid; e < v.nb * CT * CT; e += stride) {
    const int k = e / (CT * CT), r = (e / CT) % CT, c = e % CT;
    const int col = k * CT + c;
    v.aug[e] = (r == 0 && col < v.n) ? g[col] : 0.0;
  }
  // column-major tile list: column k holds rows k..nb (nb = right-hand-side row), nb - k + 1 entries starting at
  // k (nb + 1) - k (k - 1) / 2; every thread locates its own entries
  const int ntiles = v.nb * (v.nb + 1) / 2 + v.nb;
  for (int t = tid; t < ntiles; t += stride) {
    int k = 0, off = 0;
    while (off + (v.nb - k + 1) <= t) {
      off += v.nb - k + 1;
      ++k;
    }
    table[t] = make_int2(k + (t - off), k);
  }
}

__global__ void __launch_bounds__(CH_THREADS, 1)
chol_factor_kernel(CholView v, const int2* __restrict__ table, int ntiles) {
  extern __shared__ __align__(16) unsigned char craw[];
  CholSmem& S = *reinterpret_cast<CholSmem*>(craw);
  __shared__ double s_invdiag[CT];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g4 = lane >> 2, l4 = lane & 3;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int2 ik = table[t];
    const int i = ik.x, k = ik.y;
    unsigned long long* tl = g_chol_timeline ? g_chol_timeline + 4 * (size_t)t : nullptr;
    if (tl && tid == 0) tl[0] = gtime();
    double acc[8][2];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c][0] = acc[c][1] = 0.0;
    // the tile's own entries (A_ik) are read up front: they do not depend on anything
    int pitch;
    double* gt = tile_ptr(v, i, k, pitch);
    double2 own[8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
      own[c] = *reinterpret_cast<const double2*>(gt + (size_t)(8 * warp + g4) * pitch + 8 * c + 2 * l4);
    auto stage = [&](int j, int buf) {
      if (tid == 0) {
        while (ld_acquire_i32(v.flags + i * v.nb + j) == 0) {
        }
        if (i != k)
          while (ld_acquire_i32(v.flags + k * v.nb + j) == 0) {
          }
      }
      __syncthreads();
      load_tile_async(v, i, j, S.A[buf]);
      if (i != k) load_tile_async(v, k, j, S.B[buf]);
      cp_async_commit();
    };
    if (k > 0) stage(0, 0);
    for (int j = 0; j < k; ++j) {
      const int buf = j & 1;
      if (j + 1 < k) {
        stage(j + 1, buf ^ 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      tile_gemm(S.A[buf], (i != k) ? S.B[buf] : S.A[buf], acc);
      __syncthreads();
    }
    if (tl && tid == 0) tl[1] = gtime();
    // val = A_ik - acc
    {
      const int r = 8 * warp + g4;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int col = 8 * c + 2 * l4;
        S.T[r * CPITCH + col] = own[c].x - acc[c][0];
        S.T[r * CPITCH + col + 1] = own[c].y - acc[c][1];
      }
    }
    __syncthreads();
    if (i == k) {
      potrf64(S, s_invdiag);
      for (int e = tid; e < CT * CT; e += CH_THREADS) {
        const int r = e >> 6, c = e & 63;
        gt[(size_t)r * pitch + c] = (c <= r) ? S.T[r * CPITCH + c] : 0.0;
      }
      for (int e = tid; e < 8 * 64; e += CH_THREADS) v.l8inv[(size_t)k * 512 + e] = S.L8inv[e >> 6][e & 63];
    } else {
      if (tid == 0)
        while (ld_acquire_i32(v.flags + k * v.nb + k) == 0) {
        }
      if (tl && tid == 0) tl[2] = gtime();
      __syncthreads();
      {
        int lp;
        const double* src = tile_ptr(v, k, k, lp);
        for (int e = tid; e < CT * CT / 2; e += CH_THREADS) {
          const int r = e >> 5, c = 2 * (e & 31);
          cp_async16(S.X + r * CPITCH + c, src + (size_t)r * lp + c);
        }
        const double* s8 = v.l8inv + (size_t)k * 512;
        for (int e = tid; e < 256; e += CH_THREADS) cp_async16(&S.L8inv[0][0] + 2 * e, s8 + 2 * e);
        cp_async_commit();
        cp_async_wait<0>();
      }
      __syncthreads();
      trsm64(S);
      __syncthreads();
      for (int e = tid; e < CT * CT / 2; e += CH_THREADS) {
        const int r = e >> 5, c = 2 * (e & 31);
        *reinterpret_cast<double2*>(gt + (size_t)r * pitch + c) = *reinterpret_cast<const double2*>(S.T + r * CPITCH + c);
      }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release_i32(v.flags + i * v.nb + k, 1);
    if (tl && tid == 0) tl[3] = gtime();
    if (i == k) {
      // off the critical path: the 64 x 64 inverse of the diagonal block, used by the backward substitution
      inverse64(S);
      for (int e = tid; e < CT * CT; e += CH_THREADS) v.linv[(size_t)k * CT * CT + e] = S.X[(e >> 6) * CPITCH + (e & 63)];
      __syncthreads();
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Second schedule of the same factorisation: ONE CTA walks the critical path.
// In the schedule above the chain  diag(k) -> L(k+1,k) -> diag(k+1)  crosses two CTAs per column: flag, fence and a
// 32 KB tile through L2 each time, ~20 us per column.  Here CTA 0 owns every diagonal tile AND the tile below it and
// keeps the operands of the chain in shared memory:
//     T = P_D(k) - L(k,k-1) L(k,k-1)^T  ->  potrf  ->  L(k+1,k) = P_S(k+1,k) L(k,k)^-T  ->  next column,
// where the helpers (all other CTAs) have pre-accumulated everything that does not depend on the previous column:
//     P_D(k)     = A(k,k)   - sum_{j <= k-2} L(k,j)   L(k,j)^T        (kind TK_PRE_DIAG, written in place)
//     P_S(k+1,k) = A(k+1,k) - sum_{j <= k-1} L(k+1,j) L(k,j)^T        (kind TK_PRE_SUB,  written in place)
// and do the rest of the panel (rows >= k+2, incl. the right-hand-side row) and the 64 x 64 inverses for the backward
// substitution (TK_INV) as before.  Helper tiles are listed column-major and dealt round-robin; every dependency of a
// helper tile is an earlier helper tile or a result of CTA 0 for a column <= its own, and CTA 0 waits only for helper
// tiles of the column it is in: no cycle, all CTAs resident (cooperative launch).
// ------------------------------------------------------------------------------------------------------------------
enum { TK_PANEL = 0, TK_PRE_DIAG = 1, TK_PRE_SUB = 2, TK_INV = 3 };

__device__ __forceinline__ int helper_tiles_in_column(int nb, int k) {
  return (k >= 2 ? 1 : 0) + (k >= 1 ? 1 : 0) + 1 + (nb - k - 1 > 0 ? nb - k - 1 : 0);
}

__global__ void chol_table2_kernel(int nb, int4* __restrict__ table) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nb) return;
  int off = 0;
  for (int c = 0; c < k; ++c) off += helper_tiles_in_column(nb, c);
  if (k >= 2) table[off++] = make_int4(k, k, TK_PRE_DIAG, 0);
  if (k >= 1) table[off++] = make_int4(k + 1, k, TK_PRE_SUB, 0);
  for (int i = k + 2; i <= nb; ++i) table[off++] = make_int4(i, k, TK_PANEL, 0);
  table[off++] = make_int4(k, k, TK_INV, 0);
}

__device__ __forceinline__ void wait_flag(const int* f) {
  while (ld_acquire_i32(f) == 0) {
  }
}

__global__ void __launch_bounds__(CH_THREADS, 1)
chol_factor2_kernel(CholView v, const int4* __restrict__ table, int ntiles, int* __restrict__ pflags) {
  extern __shared__ __align__(16) unsigned char craw[];
  CholSmem& S = *reinterpret_cast<CholSmem*>(craw);
  __shared__ double s_invdiag[CT];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g4 = lane >> 2, l4 = lane & 3;
  const int nb = v.nb;

  // ------------------------------------------------------------------------------------------ CTA 0: the chain
  if (blockIdx.x == 0) {
    double* Lprev = S.B[1];   // L(k, k-1), left here by the previous column
    double* Sub = S.A[0];     // the tile below the diagonal: P_S(k+1,k) in, L(k+1,k) out
    double* Dnext = S.A[1];   // P_D(k+1), prefetched while the solve runs
    __shared__ int s_have_next;
    if (tid == 0) s_have_next = 0;
    __syncthreads();
    for (int k = 0; k < nb; ++k) {
      unsigned long long* tl = g_chol_timeline ? g_chol_timeline + 8 * (size_t)k : nullptr;   // debug stamps
      if (tl && tid == 0) tl[0] = gtime();
      int pitch;
      double* gt = tile_ptr(v, k, k, pitch);
      // T = P_D(k) - L(k,k-1) L(k,k-1)^T, lower 8 x 8 tiles only (the factorisation never reads the others): the 36
      // tiles are dealt to the 8 warps round-robin -- 4.5 tiles per warp instead of the 8 of a full product, on an
      // SM whose fp64 rate makes a 64^3 product cost 2 us
      const bool have = (s_have_next != 0);   // uniform: written before the last barrier of the previous column
      if (!have && k >= 2) {
        if (tid == 0) wait_flag(pflags + k * nb + k);
        __syncthreads();
      }
      {
        int rbs[5], cbs[5];
        double c0[5], c1[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const int tt = min(warp + 8 * q, 35);   // the fifth slot of warps 4..7 repeats tile 35 (result discarded)
          int rb = 0;
          while ((rb + 1) * (rb + 2) / 2 <= tt) ++rb;
          rbs[q] = rb;
          cbs[q] = tt - rb * (rb + 1) / 2;
          c0[q] = c1[q] = 0.0;
        }
        if (k >= 1) {
#pragma unroll 2
          for (int ks = 0; ks < CT / 4; ++ks) {
#pragma unroll
            for (int q = 0; q < 5; ++q)   // five independent accumulator pairs in flight
              dmma884c(c0[q], c1[q], Lprev[(8 * rbs[q] + g4) * CPITCH + l4 + 4 * ks], Lprev[(8 * cbs[q] + g4) * CPITCH + l4 + 4 * ks]);
          }
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          if (warp + 8 * q < 36) {
            const int r = 8 * rbs[q] + g4, col = 8 * cbs[q] + 2 * l4;
            const double2 own = have ? *reinterpret_cast<const double2*>(Dnext + r * CPITCH + col)
                                     : __ldcg(reinterpret_cast<const double2*>(gt + (size_t)r * pitch + col));
            S.T[r * CPITCH + col] = own.x - c0[q];
            S.T[r * CPITCH + col + 1] = own.y - c1[q];
          }
        }
      }
      __syncthreads();
      if (tl && tid == 0) tl[1] = gtime();
      potrf64(S, s_invdiag);
      if (tl && tid == 0) tl[2] = gtime();
      // the tile below the diagonal (row nb = the right-hand side): start fetching it while the factor goes out.
      // (Polling for it from an idle warp during the factorisation was measured: what the earlier fetch saves, the
      // polls cost the pivot warps at the block barriers.)
      const int i = k + 1;
      int pitch2;
      double* gt2 = tile_ptr(v, i, k, pitch2);
      if (k >= 1) {
        if (tid == 0) wait_flag(pflags + i * nb + k);
        __syncthreads();
      }
      for (int e = tid; e < CT * CT / 2; e += CH_THREADS) {
        const int r = e >> 5, c = 2 * (e & 31);
        cp_async16(Sub + r * CPITCH + c, gt2 + (size_t)r * pitch2 + c);
      }
      cp_async_commit();
      if (tl && tid == 0) tl[3] = gtime();
      for (int e = tid; e < CT * CT; e += CH_THREADS) {
        const int r = e >> 6, c = e & 63;
        const double val = (c <= r) ? S.T[r * CPITCH + c] : 0.0;
        gt[(size_t)r * pitch + c] = val;
        S.X[r * CPITCH + c] = val;          // the factor, operand of the solve below
      }
      for (int e = tid; e < 8 * 64; e += CH_THREADS) v.l8inv[(size_t)k * 512 + e] = S.L8inv[e >> 6][e & 63];
      // No fence by every thread here: the block barrier orders the other threads' stores before thread 0's release,
      // and a release is cumulative -- only thread 0 waits for the stores to be performed, the rest moves on.
      __syncthreads();
      if (tid == 0) {
        st_release_i32(v.flags + k * nb + k, 1);
        // P_D(k+1) is normally long finished: fetch it behind the solve (else the next column loads it the slow way)
        s_have_next = (k + 1 < nb && (k + 1 < 2 || ld_acquire_i32(pflags + (k + 1) * nb + (k + 1)) != 0)) ? 1 : 0;
      }
      cp_async_wait<0>();
      __syncthreads();
      if (s_have_next) {
        int pn;
        const double* gn = tile_ptr(v, k + 1, k + 1, pn);
        for (int e = tid; e < CT * CT / 2; e += CH_THREADS) {
          const int r = e >> 5, c = 2 * (e & 31);
          cp_async16(Dnext + r * CPITCH + c, gn + (size_t)r * pn + c);
        }
        cp_async_commit();
      }
      if (tl && tid == 0) tl[4] = gtime();
      trsm64(S, Sub);
      __syncthreads();
      if (tl && tid == 0) tl[5] = gtime();
      for (int e = tid; e < CT * CT / 2; e += CH_THREADS) {
        const int r = e >> 5, c = 2 * (e & 31);
        const double2 val = *reinterpret_cast<const double2*>(Sub + r * CPITCH + c);
        *reinterpret_cast<double2*>(gt2 + (size_t)r * pitch2 + c) = val;
        *reinterpret_cast<double2*>(Lprev + r * CPITCH + c) = val;
      }
      cp_async_wait<0>();
      __syncthreads();
      if (tid == 0) st_release_i32(v.flags + i * nb + k, 1);
      if (tl && tid == 0) tl[6] = gtime();
    }
    return;
  }

  // ------------------------------------------------------------------------------------------ helpers
  for (int t = blockIdx.x - 1; t < ntiles; t += gridDim.x - 1) {
    const int4 ent = table[t];
    const int i = ent.x, k = ent.y, kind = ent.z;
    if (kind == TK_INV) {
      // the 64 x 64 inverse of the diagonal block, used by the backward substitution
      if (tid == 0) wait_flag(v.flags + k * nb + k);
      __syncthreads();
      int lp;
      const double* src = tile_ptr(v, k, k, lp);
      for (int e = tid; e < CT * CT / 2; e += CH_THREADS) {
        const int r = e >> 5, c = 2 * (e & 31);
        cp_async16(S.T + r * CPITCH + c, src + (size_t)r * lp + c);
      }
      const double* s8 = v.l8inv + (size_t)k * 512;
      for (int e = tid; e < 256; e += CH_THREADS) cp_async16(&S.L8inv[0][0] + 2 * e, s8 + 2 * e);
      cp_async_commit();
      cp_async_wait<0>();
      __syncthreads();
      inverse64(S);
      for (int e = tid; e < CT * CT; e += CH_THREADS) v.linv[(size_t)k * CT * CT + e] = S.X[(e >> 6) * CPITCH + (e & 63)];
      __syncthreads();
      continue;
    }
    const int jend = (kind == TK_PRE_DIAG) ? k - 1 : k;   // terms j in [0, jend)
    double acc[8][2];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c][0] = acc[c][1] = 0.0;
    int pitch;
    double* gt = tile_ptr(v, i, k, pitch);
    double2 own[8];   // A_ik: nobody has written it yet
#pragma unroll
    for (int c = 0; c < 8; ++c)
      own[c] = *reinterpret_cast<const double2*>(gt + (size_t)(8 * warp + g4) * pitch + 8 * c + 2 * l4);
    auto stage = [&](int j, int buf) {
      if (tid == 0) {
        wait_flag(v.flags + i * nb + j);
        if (i != k) wait_flag(v.flags + k * nb + j);
      }
      __syncthreads();
      load_tile_async(v, i, j, S.A[buf]);
      if (i != k) load_tile_async(v, k, j, S.B[buf]);
      cp_async_commit();
    };
    if (jend > 0) stage(0, 0);
    for (int j = 0; j < jend; ++j) {
      const int buf = j & 1;
      if (j + 1 < jend) {
        stage(j + 1, buf ^ 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      tile_gemm(S.A[buf], (i != k) ? S.B[buf] : S.A[buf], acc);
      __syncthreads();
    }
    if (kind != TK_PANEL) {
      // pre-accumulated tile for CTA 0: A_ik - acc, in place
      const int r = 8 * warp + g4;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<double2*>(gt + (size_t)r * pitch + 8 * c + 2 * l4) = make_double2(own[c].x - acc[c][0], own[c].y - acc[c][1]);
      __threadfence();
      __syncthreads();
      if (tid == 0) st_release_i32(pflags + i * nb + k, 1);
      continue;
    }
    {
      const int r = 8 * warp + g4;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int col = 8 * c + 2 * l4;
        S.T[r * CPITCH + col] = own[c].x - acc[c][0];
        S.T[r * CPITCH + col + 1] = own[c].y - acc[c][1];
      }
    }
    if (tid == 0) wait_flag(v.flags + k * nb + k);
    __syncthreads();
    {
      int lp;
      const double* src = tile_ptr(v, k, k, lp);
      for (int e = tid; e < CT * CT / 2; e += CH_THREADS) {
        const int r = e >> 5, c = 2 * (e & 31);
        cp_async16(S.X + r * CPITCH + c, src + (size_t)r * lp + c);
      }
      const double* s8 = v.l8inv + (size_t)k * 512;
      for (int e = tid; e < 256; e += CH_THREADS) cp_async16(&S.L8inv[0][0] + 2 * e, s8 + 2 * e);
      cp_async_commit();
      cp_async_wait<0>();
    }
    __syncthreads();
    trsm64(S);
    __syncthreads();
    for (int e = tid; e < CT * CT / 2; e += CH_THREADS) {
      const int r = e >> 5, c = 2 * (e & 31);
      *reinterpret_cast<double2*>(gt + (size_t)r * pitch + c) = *reinterpret_cast<const double2*>(S.T + r * CPITCH + c);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release_i32(v.flags + i * nb + k, 1);
  }
}

// Backward substitution L^T x = y.  CTA k owns block column k: s = sum_{i>k} L_ik^T x_i as the x_i arrive,
// x_k = Linv_k^T (y_k - s).
__global__ void __launch_bounds__(CH_THREADS)
chol_backsolve_kernel(CholView v, double* __restrict__ x, int* __restrict__ xflags) {
  __shared__ double s_part[4][CT];
  __shared__ double s_x[CT];
  const int k = v.nb - 1 - blockIdx.x;   // late columns first (they are needed first)
  const int tid = threadIdx.x, c = tid & 63, part = tid >> 6;
  const int ld = v.nb * CT;
  // the inverse of this column's diagonal block is needed at the very end of the chain: fetch it now
  double linv_pref[16];
  {
    const double* li = v.linv + (size_t)k * CT * CT;
#pragma unroll
    for (int q = 0; q < 16; ++q) linv_pref[q] = li[(16 * part + q) * CT + c];
  }
  double s = 0.0;
  for (int i = v.nb - 1; i > k; --i) {
    double l[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) l[q] = v.L[(size_t)(i * CT + 16 * part + q) * ld + k * CT + c];
    if (tid == 0)
      while (ld_acquire_i32(xflags + i) == 0) {
      }
    __syncthreads();
    if (tid < CT) s_x[tid] = (i * CT + tid < v.n) ? __ldcg(x + i * CT + tid) : 0.0;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 16; ++q) s += l[q] * s_x[16 * part + q];
    __syncthreads();
  }
  s_part[part][c] = s;
  __syncthreads();
  if (tid < CT) {
    const double tot = s_part[0][tid] + s_part[1][tid] + s_part[2][tid] + s_part[3][tid];
    s_x[tid] = v.aug[(size_t)k * CT * CT + tid] - tot;   // y_k - s
  }
  __syncthreads();
  double p = 0.0;
  {
#pragma unroll
    for (int q = 0; q < 16; ++q) p += linv_pref[q] * s_x[16 * part + q];
  }
  __syncthreads();
  s_part[part][c] = p;
  __syncthreads();
  if (tid < CT && k * CT + tid < v.n) x[k * CT + tid] = s_part[0][tid] + s_part[1][tid] + s_part[2][tid] + s_part[3][tid];
  __threadfence();
  __syncthreads();
  if (tid == 0) st_release_i32(xflags + k, 1);
}

}  // namespace como

using namespace como;

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// debug only (not part of the public header): timeline buffer of 4 x ntiles u64, or NULL to switch off
extern "C" int como_b200_chol_debug_timeline(unsigned long long* buf) {
  return cudaMemcpyToSymbol(g_chol_timeline, &buf, sizeof(buf)) == cudaSuccess ? 0 : -2;
}

extern "C" int como_b200_chol_debug_probe(double* tile, long long* clk) {
  cudaFuncSetAttribute(chol_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CholSmem));
  chol_probe_kernel<<<1, CH_THREADS, sizeof(CholSmem)>>>(tile, clk);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : -2;
}

// 0 = one persistent CTA per SM.  The factorisation is bound by its critical path, not by SM count: a smaller grid
// leaves SMs to a bandwidth-bound kernel on another stream.
static int g_chol_ctas = 0;
extern "C" void como_b200_chol_ctas(int32_t ctas) { g_chol_ctas = ctas > 0 ? ctas : 0; }

// 1 (default): CTA 0 walks the critical path (chol_factor2_kernel); 0: every tile is an independent dataflow task
static int g_chol_schedule = getenv("COMO_B200_CHOL_SCHEDULE") ? atoi(getenv("COMO_B200_CHOL_SCHEDULE")) : 1;
extern "C" void como_b200_chol_schedule(int32_t mode) { g_chol_schedule = mode; }

extern "C" size_t como_b200_chol_solve_workspace_bytes(int32_t n) {
  const size_t nb = (n + CT - 1) / CT;
  return align256(sizeof(int) * (2 * (nb + 1) * nb + nb)) + align256(sizeof(int4) * (nb * (nb + 3) / 2 + 3 * nb + 8)) +
         align256(sizeof(double) * nb * CT * CT) * 2 + align256(sizeof(double) * nb * 512) +
         align256(sizeof(double) * nb * CT * nb * CT) + 256;
}

extern "C" int como_b200_chol_solve(const double* H, const double* g, int32_t n, double* x, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  COMO_REQUIRE(H && g && x && workspace, "chol_solve: null pointer argument");
  COMO_REQUIRE(n >= 1, "chol_solve: bad size");
  if (workspace_bytes < como_b200_chol_solve_workspace_bytes(n)) {
    set_last_error("chol_solve: workspace too small");
    return COMO_B200_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = (n + CT - 1) / CT;
  unsigned char* w = (unsigned char*)workspace;
  CholView v;
  v.n = n;
  v.nb = nb;
  v.flags = (int*)w;
  int* xflags = v.flags + (nb + 1) * nb;
  int* pflags = xflags + nb;
  w += align256(sizeof(int) * (2 * (nb + 1) * nb + nb));
  int2* table = (int2*)w;
  w += align256(sizeof(int4) * (nb * (nb + 3) / 2 + 3 * nb + 8));
  v.aug = (double*)w;
  w += align256(sizeof(double) * nb * CT * CT);
  v.linv = (double*)w;
  w += align256(sizeof(double) * nb * CT * CT);
  v.l8inv = (double*)w;
  w += align256(sizeof(double) * nb * 512);
  v.L = (double*)w;
  COMO_REQUIRE(nb <= sm_count(), "chol_solve: n = %d too large for the resident-CTA schedule", n);
  const int ntiles = nb * (nb + 1) / 2 + nb;   // lower triangle + the right-hand-side row
  chol_copy_kernel<<<dim3(nb, nb), 256, 0, st>>>(v, H);
  chol_prep_kernel<<<64, 256, 0, st>>>(v, g, table);
  cudaMemsetAsync(xflags, 0, sizeof(int) * nb, st);
  // the attribute is per device: set on every call (cheap) rather than once per process
  cudaFuncSetAttribute(chol_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CholSmem));
  int grid = sm_count();
  if (g_chol_ctas > 0 && g_chol_ctas < grid) grid = g_chol_ctas;   // see como_b200_chol_ctas
  if (grid > ntiles) grid = ntiles;
  // Both kernels spin on flags written by other CTAs of the same grid: launched cooperatively so that the runtime
  // refuses the launch (instead of letting it hang) if the CTAs cannot all be resident.
  if (g_chol_schedule == 1 && sm_count() >= 2) {
    int nh = 0;   // helper tiles
    for (int k = 0; k < nb; ++k) nh += (k >= 2 ? 1 : 0) + (k >= 1 ? 1 : 0) + 1 + (nb - k - 1 > 0 ? nb - k - 1 : 0);
    cudaMemsetAsync(pflags, 0, sizeof(int) * (nb + 1) * nb, st);
    int4* table4 = (int4*)table;
    chol_table2_kernel<<<(nb + 63) / 64, 64, 0, st>>>(nb, table4);
    cudaFuncSetAttribute(chol_factor2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CholSmem));
    int grid2 = sm_count();
    if (g_chol_ctas > 1 && g_chol_ctas < grid2) grid2 = g_chol_ctas;
    if (grid2 > nh + 1) grid2 = nh + 1;
    const int4* tb = table4;
    int* pf = pflags;
    void* args[] = {(void*)&v, (void*)&tb, (void*)&nh, (void*)&pf};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)chol_factor2_kernel, dim3(grid2), dim3(CH_THREADS), args,
                                                sizeof(CholSmem), st);
    if (e != cudaSuccess) {
      set_last_error("chol_factor2: cooperative launch failed: %s", cudaGetErrorString(e));
      return COMO_B200_ELAUNCH;
    }
  } else {
    int nt = ntiles;
    const int2* tb = table;
    void* args[] = {(void*)&v, (void*)&tb, (void*)&nt};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)chol_factor_kernel, dim3(grid), dim3(CH_THREADS), args,
                                                sizeof(CholSmem), st);
    if (e != cudaSuccess) {
      set_last_error("chol_factor: cooperative launch failed: %s", cudaGetErrorString(e));
      return COMO_B200_ELAUNCH;
    }
  }
  {
    double* xp = x;
    int* xf = xflags;
    void* args[] = {(void*)&v, (void*)&xp, (void*)&xf};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)chol_backsolve_kernel, dim3(nb), dim3(CH_THREADS), args, 0, st);
    if (e != cudaSuccess) {
      set_last_error("chol_backsolve: cooperative launch failed: %s", cudaGetErrorString(e));
      return COMO_B200_ELAUNCH;
    }
  }
  return check_launch("chol_solve");
}
