// Window BA photometric factors (fp64): residual pass, exact robust scale, accumulation pass, scatter.
// Replaces create_photo_system / batch_photo_cost / interp_img / robustify_system_inplace
// (como/odom/backend/photo.py:24-353), backproject_cloud / setup_test_points
// (como/odom/backend/sparse_map.py:184-230) and the einsum/scatter_add_ plumbing of
// como/odom/backend/linear_system.py:6-38.
//
// The reference materialises d Pw_n / d z_m as a (b,N,3,M,1) tensor (29.5 MB per pair).  It is rank one,
//     d r / d z_m = alpha_n * Kt[n,m] * u_m,   alpha_n = (dI/dPw) . (R_wc,i Pc_n),  u_m = 1/z_m,
// so every anchor block is a product of per-pixel scalars with the predictor row Kt[n,:]:
//     H_zz = U (sum_n A_n k_n k_n^T) U,   A_n = sum_targets alpha^2          (per REFERENCE keyframe)
//     g_z  = -U sum_n B_n k_n,            B_n = sum_targets alpha r
//     H_iz = (sum_n D_n k_n^T) U,         D_n = sum_targets alpha J_i         (8-vector)
//     H_jz = (sum_n alpha J_j k_n^T) U                                        (per pair)
// and the 3M expansion by dz/dPw (constant per keyframe) happens in the scatter.
//
// Pass A (ba_residual_kernel): one warp per 32 pixels of a reference keyframe.  Stage 1 streams the 32
//   predictor rows (512 B each, one coalesced request per row) and forms logz_n and q_n = Kt dlogz/dTwc
//   with warp shuffles; stage 2 maps lanes to pixels and loops over the keyframe's targets: project,
//   bilinear gather of [I,gx,gy], residual.  Writes r (contiguous, for the median) and dI/dPc.
// Pass B (ba_accum_kernel): work unit = (reference keyframe, pixel slice, group of <= 4 targets).
//   Predictor rows of a 32-pixel tile are staged into shared memory with 1-D bulk async copies
//   (cp.async.bulk + mbarrier, 4-deep ring); 4 coefficient warps rebuild the per-(pixel,target) Jacobians
//   of tile t+1 from the stored residual data while 8 product warps run the Gram / stack / small-Gram
//   products of tile t on the FP64 tensor path (mma.sync m8n8k4 f64 = DMMA.8x8x4, pixel index as K).
#include "ba_common.cuh"

namespace como {

int ba_build_frames(const double*, const double*, const double*, const double*, const double*, const double*, int, int,
                    size_t, BAFrame*, cudaStream_t);
template <typename T>
int median_launch(const T*, const long long*, int, long long, T, T*, long long*, void*, size_t, cudaStream_t);

constexpr double HUBER_KD = 1.345;

// 8-byte read-only load with an L2 cache policy (createpolicy): for operands that are streamed once
__device__ __forceinline__ double ldg_stream(const double* p, unsigned long long policy) {
  double v;
  asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(policy));
  return v;
}

// ================================================================================================ pass A
constexpr int RA_THREADS = 256;

__global__ void __launch_bounds__(RA_THREADS)
ba_residual_kernel(const double* __restrict__ Knm, const int32_t* __restrict__ coords, const double* __restrict__ vals_n,
                   const double* __restrict__ scaf, const BAFrame* __restrict__ frames,
                   const int32_t* __restrict__ ref_ptr, const int32_t* __restrict__ ref_pairs,
                   const int32_t* __restrict__ pair_tgt, BADims d, double* __restrict__ refbuf,
                   double* __restrict__ rbuf, double* __restrict__ pairbuf) {
  __shared__ double s_vec[7][BA_MAXM];  // logzm and dlogz/dTwc columns of this keyframe, zero padded
  const int i = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int t = tid; t < 7 * BA_MAXM; t += RA_THREADS) {
    const int v = t / BA_MAXM, m = t % BA_MAXM;
    double x = 0.0;
    if (m < d.M) x = scaf[((size_t)i * d.M + m) * SCAF_STRIDE + (v == 0 ? 0 : 7 + v)];
    s_vec[v][m] = x;
  }
  __syncthreads();
  const int n0 = (blockIdx.x * (RA_THREADS / 32) + wid) * 32;
  if (n0 >= d.N) return;
  // ---- stage 1: 32 rows x 7 dot products as one (32 x 64) x (64 x 8) product on the FP64 tensor path.  The B
  // fragments (logz_m and the six dlogz/dTwc columns, column 7 = 0) are constant for the keyframe and live in 16
  // registers; A fragments come straight from HBM: for a fixed k-step the four lanes of a quad read one 32-byte
  // sector of a predictor row, and every row is consumed completely over the 16 k-steps.  (This replaced 32
  // double2 loads + 56 FMAs + a 31-step double-shuffle butterfly per 4 rows.)
  const int g4 = lane >> 2, l4 = lane & 3;
  double bfrag[16];
#pragma unroll
  for (int ks = 0; ks < 16; ++ks) bfrag[ks] = (g4 < 7) ? s_vec[g4][4 * ks + l4] : 0.0;
  __shared__ double s_dot[RA_THREADS / 32][32][8];
  const int32_t* crd = coords + 2 * ((size_t)i * d.N);
  // The 315 MB of gathered predictor rows are read once here: evict-first in L2, so that they do not push out the
  // target frames' [I, gx, gy] planes, which several pairs (and the neighbouring keyframes' blocks) sample again --
  // ncu showed every pair re-reading its whole target image from DRAM (813 MB for 416 MB algorithmic).  Measured
  // effect of the hint alone: 185 -> 181 us; the re-reads are mostly capacity misses across waves.
  const unsigned long long pol_rows = l2_policy_evict_first();
#pragma unroll
  for (int rt = 0; rt < 4; ++rt) {
    const int n = n0 + 8 * rt + g4;
    const bool rowok = n < d.N;
    const double* row = Knm;
    if (rowok) {
      const int r = crd[2 * n], c = crd[2 * n + 1];
      row = Knm + (((size_t)i * d.H + r) * d.W + c) * d.M;
    }
    double a[16];
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) a[ks] = (rowok && 4 * ks + l4 < d.M) ? ldg_stream(row + 4 * ks + l4, pol_rows) : 0.0;
    double c0 = 0.0, c1 = 0.0;
#pragma unroll
    for (int ks = 0; ks < 16; ++ks)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0), "+d"(c1)
                   : "d"(a[ks]), "d"(bfrag[ks]));
    s_dot[wid][8 * rt + g4][2 * l4] = c0;
    s_dot[wid][8 * rt + g4][2 * l4 + 1] = c1;
  }
  __syncwarp();
  double mine[7];
#pragma unroll
  for (int q = 0; q < 7; ++q) mine[q] = s_dot[wid][lane][q];
  // ---- stage 2: lane <-> pixel
  const int n = n0 + lane;
  if (n >= d.N) return;
  const BAFrame Fi = frames[i];
  const double z = exp(mine[0]);
  const int r = crd[2 * n], c = crd[2 * n + 1];
  const double ray[3] = {((double)c - d.cx) / d.fx, ((double)r - d.cy) / d.fy, 1.0};
  const double Pc[3] = {z * ray[0], z * ray[1], z};
  double Pw[3];
  mat3_vec(Fi.Rwc, Pc, Pw);
  Pw[0] += Fi.twc[0];
  Pw[1] += Fi.twc[1];
  Pw[2] += Fi.twc[2];
  double* rb = refbuf + ((size_t)i * d.N + n) * REF_STRIDE;
  rb[0] = z;
#pragma unroll
  for (int v = 1; v < 7; ++v) rb[v] = mine[v];
  rb[7] = 0.0;
  const double vi = vals_n[(size_t)i * d.N + n];
  const size_t HW = (size_t)d.H * d.W;
  for (int t = ref_ptr[i]; t < ref_ptr[i + 1]; ++t) {
    const int p = ref_pairs[t];
    const BAFrame* Fj = frames + pair_tgt[p];
    double Pj[3];
    mat3_vec(Fj->Rcw, Pw, Pj);
    Pj[0] += Fj->tcw[0];
    Pj[1] += Fj->tcw[1];
    Pj[2] += Fj->tcw[2];
    const double X = Pj[0], Y = Pj[1], Z = Pj[2];
    const double uu = d.fx * X / Z + d.cx, vv = d.fy * Y / Z + d.cy;
    const bool valid = (uu >= 1.0) && (uu < (double)(d.W - 1)) && (vv >= 1.0) && (vv < (double)(d.H - 1)) && (Z > 0.0);
    double res = __longlong_as_double(0x7ff8000000000000LL);
    double dI[3] = {0.0, 0.0, 0.0};
    const double vsc = exp(Fj->a - Fi.a) * vi;
    if (valid) {
      const double x0f = floor(uu), y0f = floor(vv);
      const int x0 = (int)x0f, y0 = (int)y0f;
      const double fx0 = uu - x0f, fy0 = vv - y0f, fx1 = 1.0 - fx0, fy1 = 1.0 - fy0;
      const double w00 = fx1 * fy1, w01 = fx0 * fy1, w10 = fx1 * fy0, w11 = fx0 * fy0;
      const double* im = Fj->img + (size_t)y0 * d.W + x0;
      double s[3];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const double* q = im + ch * HW;
        s[ch] = __ldg(q) * w00 + __ldg(q + 1) * w01 + __ldg(q + d.W) * w10 + __ldg(q + d.W + 1) * w11;
      }
      dI[0] = s[1] * d.fx / Z;
      dI[1] = s[2] * d.fy / Z;
      dI[2] = -(s[1] * d.fx * X / Z + s[2] * d.fy * Y / Z) / Z;
      res = s[0] - vsc + (Fj->b - Fi.b);
    }
    rbuf[(size_t)p * d.N + n] = res;
    double* pb = pairbuf + ((size_t)p * d.N + n) * PAIR_STRIDE;
    *reinterpret_cast<double2*>(pb) = make_double2(dI[0], dI[1]);
    *reinterpret_cast<double2*>(pb + 2) = make_double2(dI[2], vsc);
  }
}

// ================================================================================================ pass B
constexpr int AC_THREADS = 384;      // warps 0-7: product role (256 threads), warps 8-11: coefficient role (128)
constexpr int AC_ROLE = 256;
constexpr int AC_COEF = AC_THREADS - AC_ROLE;
constexpr int TP = 32;               // pixels per tile
constexpr int TG = 4;                // targets per group (one coefficient warp each)
constexpr int ZW = 17;               // [J_i(8) | J_j(8) | r]
constexpr int NSMALL = 153;          // upper triangle of the 17x17 Gram
constexpr int SMALL_STRIDE = 160;
constexpr int STACK_ROWS = 48;       // 8 (D) + 1 (B) + 8*TG (E) = 41, padded
constexpr int PART_STRIDE = BA_MAXM * BA_MAXM + STACK_ROWS * BA_MAXM + TG * SMALL_STRIDE;  // doubles per unit

constexpr int XST = 4;               // stages of the predictor-row ring
// Shared-memory pitches are == 4 (mod 16) doubles: a DMMA fragment load touches 4 consecutive rows x 8 columns
// and lands on 16 distinct 8-byte banks per half warp.
constexpr int XPITCH = 68;           // predictor rows
constexpr int CPITCH = 52;           // per-pixel coefficients: [D(8) | B | alpha J_j (8 TG) | pad(7) | A | pad(3)]
constexpr int C_B = 8, C_E = 9, C_A = 48;
constexpr int ZPITCH = 20;           // [J_i(8) | J_j(8) | r | pad(3)]

struct BAUnit {
  int ref, pix_begin, pix_end, tgt_begin, tgt_end, primary, pad0, pad1;
};

struct AccumSmem {
  double X[XST][TP][XPITCH];         // predictor rows (bulk-copied), XST-deep ring: rows are gathered from HBM
  double refz[2][TP][REF_STRIDE];    // z_n, q_n of a tile (one bulk copy), double buffered, one tile ahead of X
  double Z[2][TG][TP + 1][ZPITCH];   // [J_i | J_j | r] of the unit's own target group (+1 row: fragment overrun)
  double C[2][TP][CPITCH];           // stack coefficients and A = sum alpha^2 per pixel
  double Cp[TG][TP][11];             // per coefficient warp: partial [D(8) | B | A] of a pixel (odd pitch: no conflicts)
  unsigned long long mbarX[XST], mbarR[2];
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ int tri_index(int a, int b) {  // a <= b < 17 -> index in the packed upper triangle
  return a * ZW - (a * (a - 1)) / 2 + (b - a);
}

// One k-step (4 pixels) of the warp's 9 Gram tiles and 12 stack tiles; Q = warp quarter (static tile lists).
template <int Q>
__device__ __forceinline__ void accum_gram_step(double (&accG)[9][2], const double (&xf)[8], double Ap) {
  constexpr int RA = 7 - Q, RB = Q;
  const double fa = Ap * xf[RA], fb = Ap * xf[RB];
#pragma unroll
  for (int c = 0; c <= RA; ++c) dmma884(accG[c][0], accG[c][1], fa, xf[c]);                    // tiles (RA, c)
#pragma unroll
  for (int c = 0; c <= RB; ++c) dmma884(accG[RA + 1 + c][0], accG[RA + 1 + c][1], fb, xf[c]);  // tiles (RB, c)
}
template <int Q>
__device__ __forceinline__ void accum_stack_step(double (&accS)[12][2], const double (&xf)[8], const double* cr, int s_lo,
                                                 int s_top) {
#pragma unroll
  for (int sr = 0; sr < 6; ++sr) {
    if (sr >= s_lo && sr < s_top) {
      const double cf = cr[8 * sr];
      dmma884(accS[2 * sr][0], accS[2 * sr][1], cf, xf[2 * Q]);
      dmma884(accS[2 * sr + 1][0], accS[2 * sr + 1][1], cf, xf[2 * Q + 1]);
    }
  }
}

// Warp-specialised: the coefficient warps (8-11) rebuild the per-(target,pixel) Jacobians of tile t+1 while
// the product warps (0-7) run the tensor-path products of tile t; one CTA-wide barrier per tile.
__global__ void __launch_bounds__(AC_THREADS, 1)
ba_accum_kernel(const double* __restrict__ Knm, const int32_t* __restrict__ coords, const double* __restrict__ scaf,
                const BAFrame* __restrict__ frames, const int32_t* __restrict__ ref_ptr,
                const int32_t* __restrict__ ref_pairs, const int32_t* __restrict__ pair_tgt,
                const double* __restrict__ sigma_pair, const BAUnit* __restrict__ units, BADims d,
                const double* __restrict__ refbuf, const double* __restrict__ rbuf, const double* __restrict__ pairbuf,
                double* __restrict__ partial) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  AccumSmem& S = *reinterpret_cast<AccumSmem*>(smem_raw);
  const BAUnit un = units[blockIdx.x];
  const int i = un.ref;
  const int tid = threadIdx.x;
  const bool is_coef = tid >= AC_ROLE;
  const int rt = tid & (AC_ROLE - 1);        // role-local thread id
  const int T_all = ref_ptr[i + 1] - ref_ptr[i];
  const int ntgt = un.tgt_end - un.tgt_begin;  // <= TG targets owned by this unit
  const unsigned row_bytes = (unsigned)(d.M * sizeof(double));
  const bool primary = un.primary != 0;
  const int ntiles = (un.pix_end - un.pix_begin + TP - 1) / TP;
  const int32_t* crd = coords + 2 * ((size_t)i * d.N);

  if (tid == 0) {
    for (int q = 0; q < XST; ++q) mbar_init(&S.mbarX[q], 1);
    mbar_init(&S.mbarR[0], 1);
    mbar_init(&S.mbarR[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // zero once: X (bulk copies only write the first M columns of the rows of a tile), the padding of Z and C
  for (int t = tid; t < XST * TP * XPITCH; t += AC_THREADS) (&S.X[0][0][0])[t] = 0.0;
  for (int t = tid; t < 2 * TG * (TP + 1) * ZPITCH; t += AC_THREADS) (&S.Z[0][0][0][0])[t] = 0.0;
  for (int t = tid; t < 2 * TP * CPITCH; t += AC_THREADS) (&S.C[0][0][0])[t] = 0.0;
  __syncthreads();

  auto issue_X = [&](int tile, int buf) {      // called by warp 0 (product role): lane <-> predictor row
    const int nb0 = un.pix_begin + tile * TP;
    const int n = nb0 + tid;
    const unsigned nrows = (unsigned)min(TP, un.pix_end - nb0);
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&S.mbarX[buf], nrows * row_bytes);
    }
    __syncwarp();
    if (n < un.pix_end) {
      const int r = crd[2 * n], c = crd[2 * n + 1];
      bulk_g2s(&S.X[buf][tid][0], Knm + (((size_t)i * d.H + r) * d.W + c) * d.M, row_bytes, &S.mbarX[buf]);
    }
  };
  auto issue_R = [&](int tile, int buf) {      // called by one thread of the coefficient role
    const int nb0 = un.pix_begin + tile * TP;
    const unsigned bytes = (unsigned)min(TP, un.pix_end - nb0) * (unsigned)(REF_STRIDE * sizeof(double));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&S.mbarR[buf], bytes);
    bulk_g2s(&S.refz[buf][0][0], refbuf + ((size_t)i * d.N + nb0) * REF_STRIDE, bytes, &S.mbarR[buf]);
  };

  if (is_coef) {
    // ================================================================ coefficient role
    const int tt = rt >> 5, pl = rt & 31;      // (target slot, pixel in tile)
    const BAFrame Fi = frames[i];
    const int g_first = primary ? 0 : un.tgt_begin;
    const int g_last = primary ? T_all : un.tgt_end;
    const int t_first = g_first + tt;
    const bool has_first = t_first < g_last;
    const int pair_first = has_first ? ref_pairs[ref_ptr[i] + t_first] : 0;
    const BAFrame* Fj_first = frames + (has_first ? pair_tgt[pair_first] : 0);
    const double isig_first = has_first ? 1.0 / sigma_pair[pair_first] : 1.0;
    const double inv_fx = 1.0 / d.fx, inv_fy = 1.0 / d.fy;
    double pf_r = 0.0;
    double2 pf_a = make_double2(0.0, 0.0), pf_b = make_double2(0.0, 0.0);
    auto prefetch = [&](int tile) {
      const int n = un.pix_begin + tile * TP + pl;
      pf_r = __longlong_as_double(0x7ff8000000000000LL);
      if (has_first && n < un.pix_end) {
        pf_r = rbuf[(size_t)pair_first * d.N + n];
        const double* pb = pairbuf + ((size_t)pair_first * d.N + n) * PAIR_STRIDE;
        pf_a = *reinterpret_cast<const double2*>(pb);
        pf_b = *reinterpret_cast<const double2*>(pb + 2);
      }
    };
    // builds Z/E/dba of `tile` into buffer `buf`
    auto build = [&](int tile, int buf) {
      const int nb = un.pix_begin + tile * TP;
      const int npx = min(TP, un.pix_end - nb);
      double pa2 = 0.0, par = 0.0, pD[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int g0 = g_first; g0 < g_last; g0 += TG) {
        const int t = g0 + tt;
        const bool own_group = (g0 == un.tgt_begin);
        double Ji[8], Jj[8], rs = 0.0, alpha = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) Ji[q] = Jj[q] = 0.0;
        if (t < g_last && pl < npx) {
          const int n = nb + pl;
          double r, isig;
          double2 q0, q1;
          const BAFrame* Fj;
          if (g0 == g_first) {
            r = pf_r;
            q0 = pf_a;
            q1 = pf_b;
            Fj = Fj_first;
            isig = isig_first;
          } else {
            const int p = ref_pairs[ref_ptr[i] + t];
            r = rbuf[(size_t)p * d.N + n];
            q0 = *reinterpret_cast<const double2*>(pairbuf + ((size_t)p * d.N + n) * PAIR_STRIDE);
            q1 = *reinterpret_cast<const double2*>(pairbuf + ((size_t)p * d.N + n) * PAIR_STRIDE + 2);
            Fj = frames + pair_tgt[p];
            isig = 1.0 / sigma_pair[p];
          }
          if (r == r) {
            const double wr = fabs(r) * isig;
            const double sc = (wr < HUBER_KD) ? isig : sqrt(HUBER_KD / wr) * isig;
            rs = r * sc;
            const double dIs[3] = {q0.x * sc, q0.y * sc, q1.x * sc};
            const double vsc = q1.y;
            const double z = S.refz[buf][pl][0];
            const int rr = crd[2 * n], cc = crd[2 * n + 1];
            const double Pc[3] = {z * (((double)cc - d.cx) * inv_fx), z * (((double)rr - d.cy) * inv_fy), z};
            double RPc[3], Pw[3], Pj[3];
            mat3_vec(Fi.Rwc, Pc, RPc);
            Pw[0] = RPc[0] + Fi.twc[0];
            Pw[1] = RPc[1] + Fi.twc[1];
            Pw[2] = RPc[2] + Fi.twc[2];
            mat3_vec(Fj->Rcw, Pw, Pj);
            Pj[0] += Fj->tcw[0];
            Pj[1] += Fj->tcw[1];
            Pj[2] += Fj->tcw[2];
            double dIw[3];
            mat3T_vec(Fj->Rcw, dIs, dIw);                       // dI/dPw = dI/dPc R_cw,j (row vector)
            alpha = dIw[0] * RPc[0] + dIw[1] * RPc[1] + dIw[2] * RPc[2];
            double bvec[3], sk[3];
            mat3T_vec(Fi.Rwc, dIw, bvec);                       // reference pose: dI/dPw [-R Pc^ | R] + alpha q^T
            row_times_skew(bvec, Pc, sk);
            Ji[0] = -sk[0] + alpha * S.refz[buf][pl][1];
            Ji[1] = -sk[1] + alpha * S.refz[buf][pl][2];
            Ji[2] = -sk[2] + alpha * S.refz[buf][pl][3];
            Ji[3] = bvec[0] + alpha * S.refz[buf][pl][4];
            Ji[4] = bvec[1] + alpha * S.refz[buf][pl][5];
            Ji[5] = bvec[2] + alpha * S.refz[buf][pl][6];
            Ji[6] = vsc * sc;
            Ji[7] = -sc;
            row_times_skew(dIs, Pj, sk);                        // target pose: dI/dPc [Pc_j^ | -I]
            Jj[0] = sk[0];
            Jj[1] = sk[1];
            Jj[2] = sk[2];
            Jj[3] = -dIs[0];
            Jj[4] = -dIs[1];
            Jj[5] = -dIs[2];
            Jj[6] = -Ji[6];
            Jj[7] = -Ji[7];
          }
        }
        pa2 += alpha * alpha;
        par += alpha * rs;
#pragma unroll
        for (int q = 0; q < 8; ++q) pD[q] += alpha * Ji[q];
        if (own_group) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            S.Z[buf][tt][pl][q] = Ji[q];
            S.Z[buf][tt][pl][8 + q] = Jj[q];
            S.C[buf][pl][C_E + 8 * tt + q] = alpha * Jj[q];
          }
          S.Z[buf][tt][pl][16] = rs;
        }
      }
      // The four coefficient warps hold partial sums of the same pixel's D, B, A columns.  They used to add them
      // into S.C with shared-memory double atomics: 90 % of the kernel's excess shared wavefronts
      // (profiles/r02_ba_photo_kernels_full.txt) and an order that changed from run to run.  Now each warp leaves its
      // partials in its own slot and reduce_partials() sums the four in a fixed order.
      {
        double* dst = &S.Cp[tt][pl][0];
#pragma unroll
        for (int q = 0; q < 8; ++q) dst[q] = primary ? pD[q] : 0.0;
        dst[8] = primary ? par : 0.0;
        dst[9] = primary ? pa2 : 0.0;
      }
    };
    // thread -> (pixel, column group): D[0..3], D[4..7], B, A of C[buf] = sum over the four warps' partials
    auto reduce_partials = [&](int buf) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int px = rt & 31, part = rt >> 5;
      double* dst = &S.C[buf][px][0];
      if (part < 2) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int col = 4 * part + q;
          dst[col] = ((S.Cp[0][px][col] + S.Cp[1][px][col]) + S.Cp[2][px][col]) + S.Cp[3][px][col];
        }
      } else {
        const int col = (part == 2) ? 8 : 9;
        dst[part == 2 ? C_B : C_A] = ((S.Cp[0][px][col] + S.Cp[1][px][col]) + S.Cp[2][px][col]) + S.Cp[3][px][col];
      }
    };

    if (ntiles > 0) {
      if (rt == 0) {
        issue_R(0, 0);
        if (ntiles > 1) issue_R(1, 1);
      }
      prefetch(0);
      mbar_wait(&S.mbarR[0], 0);
      build(0, 0);
      reduce_partials(0);
      if (ntiles > 1) prefetch(1);
    }
    __syncthreads();
    for (int tile = 0; tile < ntiles; ++tile) {
      const int buf = tile & 1;
      if (tile + 1 < ntiles) {
        mbar_wait(&S.mbarR[buf ^ 1], ((tile + 1) >> 1) & 1);
        build(tile + 1, buf ^ 1);
        reduce_partials(buf ^ 1);
        // refz[buf] (tile) is free now for tile + 2; its copy overlaps the next iteration
        if (tile + 2 < ntiles) {
          if (rt == 0) issue_R(tile + 2, buf);   // refz[buf] (tile) was last read during the previous iteration
          prefetch(tile + 2);
        }
      }
      __syncthreads();
    }
    return;
  }

  // ================================================================== product role
  // FP64 tensor path (mma.sync m8n8k4, SASS DMMA.8x8x4): per 32-pixel tile the sums over pixels are three GEMMs
  // with the pixel index as K dimension,
  //   G  (64 x 64, lower tiles)  = (A X)^T X            stack (48 x 64) = C^T X           Z_t (17 x 17) = Z_t^T Z_t,
  // 36 + 48 + 24 = 108 output tiles of 8 x 8.  Warp w = (quarter q = w & 3, pixel half h = w >> 2) owns 27 tiles
  // (54 accumulator registers) and runs them over 4 of the 8 k-steps of every tile:
  //   G rows {7-q, q} (9 tiles), stack columns {2q, 2q+1} (12 tiles), small Gram of target q (6 tiles);
  // 18 conflict-free fragment loads feed 27 DMMAs per k-step.  The two pixel halves are summed at the end.
  {
    const int warp = rt >> 5, lane = rt & 31;
    const int g4 = lane >> 2, l4 = lane & 3;
    const int q = warp & 3, h = warp >> 2;
    const int ra = 7 - q, rb = q;
    const bool do_z = q < ntgt;
    const int s_lo = primary ? 0 : 1;                       // stack rows 0..7 (D) exist for the primary unit only
    const int s_top = min(6, ntgt + 2);                     // rows >= 9 + 8 ntgt are empty
    double accG[9][2], accS[12][2], accZ[6][2];
#pragma unroll
    for (int t = 0; t < 9; ++t) accG[t][0] = accG[t][1] = 0.0;
#pragma unroll
    for (int t = 0; t < 12; ++t) accS[t][0] = accS[t][1] = 0.0;
#pragma unroll
    for (int t = 0; t < 6; ++t) accZ[t][0] = accZ[t][1] = 0.0;

    if (tid < 32)
      for (int t = 0; t < XST - 1 && t < ntiles; ++t) issue_X(t, t);
    __syncthreads();   // pairs with the coefficient role's prologue barrier
    for (int tile = 0; tile < ntiles; ++tile) {
      const int buf = tile & 1;
      const int xs = tile % XST;
      // stage (tile + XST - 1) % XST held tile - 1, which every product thread finished before the last barrier
      if (tile + XST - 1 < ntiles && tid < 32) issue_X(tile + XST - 1, (tile + XST - 1) % XST);
      mbar_wait(&S.mbarX[xs], (tile / XST) & 1);
#pragma unroll 2
      for (int ks = 0; ks < 4; ++ks) {
        const int prow = 16 * h + 4 * ks + l4;              // pixel row of this lane's fragment elements
        const double* xr = &S.X[xs][prow][g4];
        const double* cr = &S.C[buf][prow][g4];
        double xf[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) xf[c] = xr[8 * c];
        if (primary) {
          const double Ap = S.C[buf][prow][C_A];
          switch (q) {   // warp-uniform
            case 0: accum_gram_step<0>(accG, xf, Ap); break;
            case 1: accum_gram_step<1>(accG, xf, Ap); break;
            case 2: accum_gram_step<2>(accG, xf, Ap); break;
            default: accum_gram_step<3>(accG, xf, Ap); break;
          }
        }
        switch (q) {
          case 0: accum_stack_step<0>(accS, xf, cr, s_lo, s_top); break;
          case 1: accum_stack_step<1>(accS, xf, cr, s_lo, s_top); break;
          case 2: accum_stack_step<2>(accS, xf, cr, s_lo, s_top); break;
          default: accum_stack_step<3>(accS, xf, cr, s_lo, s_top); break;
        }
        if (do_z) {
          const double* zr = &S.Z[buf][q][prow][g4];
          const double z0 = zr[0], z1 = zr[8], z2 = zr[16];
          dmma884(accZ[0][0], accZ[0][1], z0, z0);   // (0,0)
          dmma884(accZ[1][0], accZ[1][1], z1, z0);   // (1,0)
          dmma884(accZ[2][0], accZ[2][1], z1, z1);   // (1,1)
          dmma884(accZ[3][0], accZ[3][1], z2, z0);   // (2,0)
          dmma884(accZ[4][0], accZ[4][1], z2, z1);   // (2,1)
          dmma884(accZ[5][0], accZ[5][1], z2, z2);   // (2,2)
        }
      }
      __syncthreads();
    }

    // ---------------- combine the two pixel halves (through the idle X ring) and write the unit's partial sums;
    // the scatter kernels read full G / all stack rows / the packed upper triangle of the 17x17 Grams.
    double* xch = &S.X[0][0][0];                    // 4 quarters x 27 tiles x 64 doubles = 55 KB < sizeof(S.X)
    if (h == 1) {
      double* dst = xch + ((size_t)q * 27) * 64 + 2 * lane;
#pragma unroll
      for (int t = 0; t < 9; ++t) *reinterpret_cast<double2*>(dst + t * 64) = make_double2(accG[t][0], accG[t][1]);
#pragma unroll
      for (int t = 0; t < 12; ++t) *reinterpret_cast<double2*>(dst + (9 + t) * 64) = make_double2(accS[t][0], accS[t][1]);
#pragma unroll
      for (int t = 0; t < 6; ++t) *reinterpret_cast<double2*>(dst + (21 + t) * 64) = make_double2(accZ[t][0], accZ[t][1]);
    }
    asm volatile("bar.sync 2, 256;" ::: "memory");   // product role only (the coefficient warps have left)
    if (h == 0) {
      const double* src = xch + ((size_t)q * 27) * 64 + 2 * lane;
      double* out = partial + (size_t)blockIdx.x * PART_STRIDE;
      double* outS = out + BA_MAXM * BA_MAXM;
      double* outZ = outS + STACK_ROWS * BA_MAXM;
      // C fragment: element (8 r + g4, 8 c + 2 l4 + {0,1})
      if (primary) {
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const double2 o = *reinterpret_cast<const double2*>(src + t * 64);
          const double v0 = accG[t][0] + o.x, v1 = accG[t][1] + o.y;
          // slot t: t <= ra -> tile (ra, t); else tile (rb, t - ra - 1)
          const int tr = (t <= ra) ? ra : rb, tc = (t <= ra) ? t : (t - ra - 1);
          const int row = 8 * tr + g4, col = 8 * tc + 2 * l4;
          out[row * BA_MAXM + col] = v0;
          out[row * BA_MAXM + col + 1] = v1;
          if (tr != tc) {
            out[col * BA_MAXM + row] = v0;
            out[(col + 1) * BA_MAXM + row] = v1;
          }
        }
      }
#pragma unroll
      for (int t = 0; t < 12; ++t) {
        const double2 o = *reinterpret_cast<const double2*>(src + (9 + t) * 64);
        const int sr = t >> 1, cc = 2 * q + (t & 1);
        const int row = 8 * sr + g4, col = 8 * cc + 2 * l4;
        const bool livet = (sr >= s_lo) && (sr < s_top);
        outS[row * BA_MAXM + col] = livet ? accS[t][0] + o.x : 0.0;
        outS[row * BA_MAXM + col + 1] = livet ? accS[t][1] + o.y : 0.0;
      }
      // small Gram of target q: packed upper triangle (a <= b); tile (ta, tb) holds rows 8 ta.., cols 8 tb.. with
      // ta >= tb, i.e. element (row, col) -> packed index of (col, row) when col <= row
      {
        double* oz = outZ + q * SMALL_STRIDE;
#pragma unroll
        for (int t = 0; t < 6; ++t) {
          const double2 o = *reinterpret_cast<const double2*>(src + (21 + t) * 64);
          const double v[2] = {accZ[t][0] + o.x, accZ[t][1] + o.y};
          const int zta = (t == 0) ? 0 : (t < 3 ? 1 : 2);          // tiles (0,0) (1,0) (1,1) (2,0) (2,1) (2,2)
          const int ztb = (t == 2) ? 1 : (t < 4 ? 0 : t - 3);
          const int row = 8 * zta + g4;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = 8 * ztb + 2 * l4 + e;
            if (row < ZW && col < ZW && col <= row) oz[tri_index(col, row)] = do_z ? v[e] : 0.0;
          }
        }
        // padding slots of the packed triangle are never read
      }
    }
  }
}

// ================================================================================================ scatter
// One CTA per reference keyframe: sums the primary units, expands M -> 3M with dz/dPw, adds into H, g.
__global__ void __launch_bounds__(256)
ba_scatter_ref_kernel(const double* __restrict__ partial, const int32_t* __restrict__ unit_base,
                      const int32_t* __restrict__ unit_slices, const double* __restrict__ scaf,
                      const double* __restrict__ dz_dP, const int32_t* __restrict__ lm_ids, BADims d, int dim,
                      double* __restrict__ H, double* __restrict__ g) {
  extern __shared__ double sm[];
  double* G = sm;                          // M*M (stride BA_MAXM)
  double* St = sm + BA_MAXM * BA_MAXM;     // 9 rows x BA_MAXM
  __shared__ double s_u[BA_MAXM];
  __shared__ int s_lm[BA_MAXM];
  const int i = blockIdx.x, tid = threadIdx.x;
  const int ub = unit_base[i], ns = unit_slices[i];
  if (ns == 0) return;  // this keyframe is the reference of no pair on this rank
  for (int t = tid; t < BA_MAXM * BA_MAXM + 9 * BA_MAXM; t += 256) {
    double s = 0.0;
    for (int sl = 0; sl < ns; ++sl) s += partial[(size_t)(ub + sl) * PART_STRIDE + t];
    sm[t] = s;
  }
  if (tid < BA_MAXM) {
    s_u[tid] = (tid < d.M) ? scaf[((size_t)i * d.M + tid) * SCAF_STRIDE + 1] : 0.0;
    s_lm[tid] = (tid < d.M) ? lm_ids[i * d.M + tid] : 0;
  }
  __syncthreads();
  const double d3[3] = {dz_dP[3 * i], dz_dP[3 * i + 1], dz_dP[3 * i + 2]};
  const int lm_start = 8 * (d.K + d.R);
  const int M3 = 3 * d.M;
  // H_PP
  for (int t = tid; t < M3 * M3; t += 256) {
    const int ra = t / M3, rb = t % M3;
    const int m = ra / 3, c = ra % 3, m2 = rb / 3, c2 = rb % 3;
    const double v = d3[c] * s_u[m] * G[m * BA_MAXM + m2] * s_u[m2] * d3[c2];
    atomicAdd(H + (size_t)(lm_start + 3 * s_lm[m] + c) * dim + (lm_start + 3 * s_lm[m2] + c2), v);
  }
  // g_P = -u_m v[m] d_c   (v = stack row 8)
  for (int t = tid; t < M3; t += 256) {
    const int m = t / 3, c = t % 3;
    atomicAdd(g + lm_start + 3 * s_lm[m] + c, -s_u[m] * St[8 * BA_MAXM + m] * d3[c]);
  }
  // H_iP (reference pose x anchors), both sides
  for (int t = tid; t < 8 * M3; t += 256) {
    const int a = t / M3, rb = t % M3;
    const int m = rb / 3, c = rb % 3;
    const double v = St[a * BA_MAXM + m] * s_u[m] * d3[c];
    const size_t ri = 8 * (size_t)i + a, ci = lm_start + 3 * s_lm[m] + c;
    atomicAdd(H + ri * dim + ci, v);
    atomicAdd(H + ci * dim + ri, v);
  }
}

// One CTA per pair: pose blocks, gradient, target-pose x anchor block, robust error.
__global__ void __launch_bounds__(256)
ba_scatter_pair_kernel(const double* __restrict__ partial, const int32_t* __restrict__ unit_base,
                       const int32_t* __restrict__ unit_slices, const int32_t* __restrict__ pair_ref,
                       const int32_t* __restrict__ pair_tgt, const int32_t* __restrict__ pair_slot,
                       const double* __restrict__ scaf, const double* __restrict__ dz_dP,
                       const int32_t* __restrict__ lm_ids, BADims d, int dim, double* __restrict__ H,
                       double* __restrict__ g, double* __restrict__ err) {
  __shared__ double E[8][BA_MAXM];
  __shared__ double Zs[SMALL_STRIDE];
  const int p = blockIdx.x, tid = threadIdx.x;
  const int i = pair_ref[p], f = pair_tgt[p], slot = pair_slot[p];
  const int grp = slot / TG, tg = slot % TG;
  const int ns = unit_slices[i];
  const int ub = unit_base[i] + grp * ns;
  for (int t = tid; t < 8 * BA_MAXM + SMALL_STRIDE; t += 256) {
    double s = 0.0;
    if (t < 8 * BA_MAXM) {
      const int a = t / BA_MAXM, m = t % BA_MAXM;
      for (int sl = 0; sl < ns; ++sl)
        s += partial[(size_t)(ub + sl) * PART_STRIDE + BA_MAXM * BA_MAXM + (9 + 8 * tg + a) * BA_MAXM + m];
      E[a][m] = s;
    } else {
      const int o = t - 8 * BA_MAXM;
      for (int sl = 0; sl < ns; ++sl)
        s += partial[(size_t)(ub + sl) * PART_STRIDE + BA_MAXM * BA_MAXM + STACK_ROWS * BA_MAXM + tg * SMALL_STRIDE + o];
      Zs[o] = s;
    }
  }
  __syncthreads();
  const double d3[3] = {dz_dP[3 * i], dz_dP[3 * i + 1], dz_dP[3 * i + 2]};
  const int lm_start = 8 * (d.K + d.R);
  const size_t bi = 8 * (size_t)i, bj = 8 * (size_t)f;
  // 17x17 packed Gram: rows 0..7 J_i, 8..15 J_j, 16 r
  if (tid < 64) {
    const int a = tid / 8, b = tid % 8;
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    atomicAdd(H + (bi + a) * dim + (bi + b), Zs[tri_index(lo, hi)]);
    atomicAdd(H + (bj + a) * dim + (bj + b), Zs[tri_index(8 + lo, 8 + hi)]);
    const double hij = Zs[tri_index(a, 8 + b)];
    atomicAdd(H + (bi + a) * dim + (bj + b), hij);
    atomicAdd(H + (bj + b) * dim + (bi + a), hij);
  } else if (tid < 72) {
    const int a = tid - 64;
    atomicAdd(g + bi + a, -Zs[tri_index(a, 16)]);
    atomicAdd(g + bj + a, -Zs[tri_index(8 + a, 16)]);
  } else if (tid == 72) {
    atomicAdd(err, Zs[tri_index(16, 16)]);
  }
  const int M3 = 3 * d.M;
  for (int t = tid; t < 8 * M3; t += 256) {
    const int a = t / M3, rb = t % M3;
    const int m = rb / 3, c = rb % 3;
    const double u = scaf[((size_t)i * d.M + m) * SCAF_STRIDE + 1];
    const double v = E[a][m] * u * d3[c];
    const size_t ri = bj + a, ci = lm_start + 3 * (size_t)lm_ids[i * d.M + m] + c;
    atomicAdd(H + ri * dim + ci, v);
    atomicAdd(H + ci * dim + ri, v);
  }
}

// sigma per pair from sigma per batch
__global__ void sigma_expand_kernel(const double* __restrict__ sigma_batch, const int32_t* __restrict__ pair_batch, int P,
                                    double* __restrict__ sigma_pair) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < P) sigma_pair[p] = sigma_batch[pair_batch[p]];
}

}  // namespace como

using namespace como;

extern "C" size_t como_b200_ba_frames_bytes(int32_t num_frames) { return (size_t)num_frames * sizeof(BAFrame); }
extern "C" size_t como_b200_ba_partial_doubles(int32_t num_units) { return (size_t)num_units * PART_STRIDE; }
extern "C" int32_t como_b200_ba_unit_ints(void) { return (int32_t)(sizeof(BAUnit) / sizeof(int32_t)); }
extern "C" int32_t como_b200_ba_target_group(void) { return TG; }

static BADims make_dims(int K, int R, int L, int M, int N, int H, int W, int P, const double* intr4) {
  BADims d{};
  d.K = K; d.R = R; d.L = L; d.M = M; d.N = N; d.H = H; d.W = W; d.P = P;
  d.fx = intr4[0]; d.fy = intr4[1]; d.cx = intr4[2]; d.cy = intr4[3];
  return d;
}

// Pass A: frame table + residuals of every (pair, pixel) of this rank.
extern "C" int como_b200_ba_photo_residual(
    const double* kf_poses, const double* kf_aff, const double* rec_poses, const double* rec_aff, const double* kf_img,
    const double* rec_img, const double* Knm, const int32_t* coords, const double* vals_n, const double* scaffold,
    const int32_t* pair_tgt, const int32_t* ref_ptr, const int32_t* ref_pairs, int32_t K, int32_t R, int32_t M, int32_t N,
    int32_t Himg, int32_t Wimg, int32_t P, const double* intr4, void* frames_ws, double* refbuf, double* rbuf,
    double* pairbuf, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  COMO_REQUIRE(kf_poses && kf_aff && kf_img && Knm && coords && vals_n && scaffold && pair_tgt && ref_ptr && ref_pairs &&
                   intr4 && frames_ws && refbuf && rbuf && pairbuf,
               "ba_photo_residual: null pointer argument");
  COMO_REQUIRE(R == 0 || (rec_poses && rec_aff && rec_img), "ba_photo_residual: null one-way frame pointers");
  COMO_REQUIRE(M >= 4 && M <= BA_MAXM && (M % 4) == 0, "ba_photo: M must be a multiple of 4 and <= 64 (got %d)", M);
  COMO_REQUIRE(K >= 1 && N >= 1 && P >= 1, "ba_photo_residual: bad sizes");
  const BADims d = make_dims(K, R, 0, M, N, Himg, Wimg, P, intr4);
  BAFrame* frames = (BAFrame*)frames_ws;
  int rc = ba_build_frames(kf_poses, kf_aff, rec_poses, rec_aff, kf_img, rec_img, K, R, (size_t)3 * Himg * Wimg, frames, st);
  if (rc) return rc;
  dim3 grid((N + RA_THREADS - 1) / RA_THREADS, K);
  ba_residual_kernel<<<grid, RA_THREADS, 0, st>>>(Knm, coords, vals_n, scaffold, frames, ref_ptr, ref_pairs, pair_tgt, d, refbuf,
                                                  rbuf, pairbuf);
  return check_launch("ba_photo_residual");
}

// Pass B + scatter.  sigma_batch (num_batches) = 1.4826 * median |r| per pair batch (global across ranks);
// pair_batch (P) = batch of each local pair.
extern "C" int como_b200_ba_photo_accum(
    const double* Knm, const int32_t* coords, const double* scaffold, const double* dz_dP, const int32_t* lm_ids,
    const int32_t* pair_ref, const int32_t* pair_tgt, const int32_t* pair_slot, const int32_t* pair_batch,
    const int32_t* ref_ptr, const int32_t* ref_pairs, const int32_t* units, const int32_t* unit_base,
    const int32_t* unit_slices, int32_t num_units, int32_t K, int32_t R, int32_t L, int32_t M, int32_t N, int32_t Himg,
    int32_t Wimg, int32_t P, const double* intr4, int32_t dim, const double* sigma_batch, const void* frames_ws,
    const double* refbuf, const double* rbuf, const double* pairbuf, double* sigma_pair_ws, double* partial, double* H,
    double* g, double* photo_err, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  COMO_REQUIRE(Knm && coords && scaffold && dz_dP && lm_ids && pair_ref && pair_tgt && pair_slot && pair_batch && ref_ptr &&
                   ref_pairs && units && unit_base && unit_slices && intr4 && sigma_batch && frames_ws && refbuf && rbuf &&
                   pairbuf && sigma_pair_ws && partial && H && g && photo_err,
               "ba_photo_accum: null pointer argument");
  COMO_REQUIRE(M >= 4 && M <= BA_MAXM && (M % 4) == 0, "ba_photo: M must be a multiple of 4 and <= 64 (got %d)", M);
  COMO_REQUIRE(dim == 8 * (K + R) + 3 * L, "ba_photo_accum: dim %d != 8(K+R)+3L", dim);
  COMO_REQUIRE(num_units >= 1 && P >= 1, "ba_photo_accum: bad sizes");
  const BADims d = make_dims(K, R, L, M, N, Himg, Wimg, P, intr4);
  const BAFrame* frames = (const BAFrame*)frames_ws;
  sigma_expand_kernel<<<(P + 127) / 128, 128, 0, st>>>(sigma_batch, pair_batch, P, sigma_pair_ws);
  // the attribute is per device: set on every call (cheap) rather than once per process
  cudaFuncSetAttribute(ba_accum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AccumSmem));
  ba_accum_kernel<<<num_units, AC_THREADS, sizeof(AccumSmem), st>>>(Knm, coords, scaffold, frames, ref_ptr, ref_pairs, pair_tgt,
                                                                     sigma_pair_ws, (const BAUnit*)units, d, refbuf, rbuf,
                                                                     pairbuf, partial);
  int rc = check_launch("ba_accum");
  if (rc) return rc;
  const size_t smem = (size_t)(BA_MAXM * BA_MAXM + 9 * BA_MAXM) * sizeof(double);
  // the attribute is per device: set on every call (cheap) rather than once per process
  cudaFuncSetAttribute(ba_scatter_ref_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ba_scatter_ref_kernel<<<K, 256, smem, st>>>(partial, unit_base, unit_slices, scaffold, dz_dP, lm_ids, d, dim, H, g);
  ba_scatter_pair_kernel<<<P, 256, 0, st>>>(partial, unit_base, unit_slices, pair_ref, pair_tgt, pair_slot, scaffold, dz_dP,
                                            lm_ids, d, dim, H, g, photo_err);
  return check_launch("ba_scatter");
}
