// eigh_tridiag_hs.cuh -- K1 for the large phases (live block 96 -> 72 -> 48): Householder
// tridiagonalisation with the matrix in registers in HERMITIAN HALF STORAGE, laid out as FP64
// tensor-core accumulator fragments.
//
// Why (profiles/r2_ncu_full_summary.md): with the full matrix in registers
// (eigh_tridiag_rw.cuh) one 96 x 96 matrix fills an SM, so a Householder step is one dependency
// chain per SM: 2 400 cycles of FP64 pipe and 2 400 shared-memory wavefronts run back to back
// (FP64 43 %, LSU 48 %).  Here
//   * only the lower triangle is kept, as 24 x 24 "superblocks" of 3 x 3 tiles of 8 x 8; a tile is a
//     pair (re, im) of mma.m8n8k4 accumulator fragments: lane (g, q) = (lane >> 2, lane & 3) holds the
//     elements (row g, columns 2q, 2q + 1).  A warp owns one off-diagonal superblock (9 tiles) and, warps
//     0 .. S-1, the lower 6 tiles of a diagonal one: 120 registers of matrix per thread, 168 in all, i.e.
//     three warps per scheduler: TWO 96 x 96 matrices per SM (four at 72 x 72), whose phases overlap;
//     the 48 x 48 phase uses superblocks of 2 x 2 tiles (T = 2: 56 registers of matrix, six CTAs per SM);
//   * the rank-2 update A -= v w^H + w v^H is exactly one k = 4 DMMA per tile and per part:
//     Re -= [v_r.x v_r.y w_r.x w_r.y] . [w_c.x w_c.y v_c.x v_c.y]^T, Im likewise with the row operand
//     permuted -- 2 DMMAs per tile instead of 128 DFMAs per warp, and the operands are ONE 16-byte
//     load per row tile and ONE 8-byte load per column tile (the full-storage kernel loads v and p for
//     every row and column of the thread's tile: 4 x the shared-memory wavefronts);
//   * the symmetric product y = A x' uses every stored off-diagonal superblock twice (row direction:
//     reduce over the 4 lanes of a row; column direction: reduce over the 8 lanes of a column), the
//     partial sums of the superblocks meet in shared memory; the scalar a2 = -1/2 tau p^H v needs no
//     second pass: p^H v is a multiple of the Hermitian form x'^H A x', which every superblock
//     accumulates from its own partial products during the mat-vec (one warp sum + NW partials), so
//     after the barrier every warp finishes p, v and w = p + a2 v for its share of the rows and writes
//     them already in DMMA operand order (no selects or negations in the update).
// Three barriers per step; arithmetic identical to tools/hql_prototype.py::tridiag_lower (zhetd2,
// lower).  The kernel stops after `nsteps` steps and hands the trailing block to the next phase like
// hql_tridiag_rw_kernel (same outputs: d, e, tau, packed reflectors).
#pragma once
#include <type_traits>
#include "common.cuh"
#include "polar.cuh"  // dmma884

namespace musim {

struct HsTile {
  double re[2], im[2];
};

// NT tiles of 8 per side, superblocks of T x T tiles: (12, 3) = 96, (9, 3) = 72, (6, 2) = 48.
template <int NT, int T>
struct HsGeom {
  static constexpr int S = NT / T;              // superblock rows (4 or 3)
  static constexpr int NW = S * (S - 1) / 2;    // warps = off-diagonal superblocks (6 or 3)
  static constexpr int D = 8 * NT;
  static constexpr int TB = 8 * T;              // rows of a superblock
  static constexpr int CTAS = (T == 3) ? 12 / NW : (NW == 3 ? 5 : 3);  // T = 3: 12 warps per SM at 168 registers
};

__device__ __forceinline__ cplx hs_shfl(cplx v, int m) {
  return make_c(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}
__device__ __forceinline__ cplx hs_sel(bool c, cplx a, cplx b) { return make_c(c ? a.x : b.x, c ? a.y : b.y); }

// Reduce T per-row-tile values over the 4 lanes of a row: lane q ends up with row tile q (higher q
// duplicate the last tile) and stores it.
template <int T>
__device__ __forceinline__ void hs_reduce_rows(const cplx (&yr)[T], int lane, cplx *dst) {
  const bool b0 = (lane & 1) != 0, b1 = (lane & 2) != 0;
  const int g = lane >> 2, q = lane & 3;
  const cplx rcv = hs_shfl(hs_sel(b0, yr[0], yr[1]), 1);
  const cplx t = cadd(hs_sel(b0, yr[1], yr[0]), rcv);  // tile b0
  if (T == 4) {
    const cplx rcv3 = hs_shfl(hs_sel(b0, yr[T - 2], yr[T - 1]), 1);
    const cplx t3 = cadd(hs_sel(b0, yr[T - 1], yr[T - 2]), rcv3);  // tile 2 + b0
    const cplx rcv2 = hs_shfl(hs_sel(b1, t, t3), 2);
    const cplx f = cadd(hs_sel(b1, t3, t), rcv2);  // tile q
    dst[8 * q + g] = f;
  } else if (T == 3) {
    const cplx t2 = cadd(yr[T - 1], hs_shfl(yr[T - 1], 1));
    const cplx rcv2 = hs_shfl(hs_sel(b1, t, t2), 2);
    const cplx f = cadd(hs_sel(b1, t2, t), rcv2);
    if (q < 3) dst[8 * q + g] = f;
  } else {
    const cplx f = cadd(t, hs_shfl(t, 2));
    if (q < 2) dst[8 * q + g] = f;
  }
}
// Stage 1 of the reduction over the 8 lanes of a column: two columns -> one (column 2q + g0).
__device__ __forceinline__ cplx hs_reduce_cols1(cplx yc0, cplx yc1, int lane) {
  const bool g0 = (lane & 4) != 0;
  const cplx rcv = hs_shfl(hs_sel(g0, yc0, yc1), 4);
  return cadd(hs_sel(g0, yc1, yc0), rcv);
}
// Stages 2 and 3 for NU tile columns (u[tj] = column 2q + g0 of tile tj) and the store.
template <int NU>
__device__ __forceinline__ void hs_reduce_cols23(const cplx (&u)[NU > 0 ? NU : 1], int lane, cplx *dst) {
  const bool g1 = (lane & 8) != 0, g2 = (lane & 16) != 0;
  const int col = 2 * (lane & 3) + ((lane >> 2) & 1);
  if (NU == 3) {
    const cplx rcv = hs_shfl(hs_sel(g1, u[0], u[1]), 8);
    const cplx z = cadd(hs_sel(g1, u[1], u[0]), rcv);  // tile g1
    const cplx z2 = cadd(u[NU - 1], hs_shfl(u[NU - 1], 8));
    const cplx rcv2 = hs_shfl(hs_sel(g2, z, z2), 16);
    const cplx f = cadd(hs_sel(g2, z2, z), rcv2);  // g2 = 0: tile g1, g2 = 1: tile 2
    if (!(g1 && g2)) dst[8 * (g2 ? 2 : (g1 ? 1 : 0)) + col] = f;
  } else if (NU == 2) {
    const cplx rcv = hs_shfl(hs_sel(g1, u[0], u[1]), 8);
    cplx z = cadd(hs_sel(g1, u[1], u[0]), rcv);  // tile g1
    z = cadd(z, hs_shfl(z, 16));
    if (!g2) dst[8 * (g1 ? 1 : 0) + col] = z;
  } else if (NU == 1) {
    cplx z = cadd(u[0], hs_shfl(u[0], 8));
    z = cadd(z, hs_shfl(z, 16));
    if (!g1 && !g2) dst[col] = z;
  }
}

// y += A x' over an OFF-DIAGONAL superblock (SI > SJ): row direction y_I += A x_J (partial sums to
// yrow_dst[TB SI + ...]) and column direction y_J += A^H x_I (partial sums to ycol_dst[TB SJ + ...]).
// x' = x except x'_{k+1} = alpha - beta: only the real part differs (xp0x).  qacc += 2 Re(x_I^H A x_J).
template <int T>
__device__ __forceinline__ void hs_matvec_off(const HsTile (&a)[T][T], int SI, int SJ, int k, const cplx *x, double xp0x,
                                              int lane, cplx *yrow_dst, cplx *ycol_dst, double &qacc) {
  constexpr int TB = 8 * T;
  const int g = lane >> 2, q = lane & 3;
  const int r0 = TB * SI + g, c0 = TB * SJ + 2 * q;
  cplx xr[T];
#pragma unroll
  for (int ti = 0; ti < T; ++ti) {
    xr[ti] = x[r0 + 8 * ti];
    if (r0 + 8 * ti == k + 1) xr[ti].x = xp0x;
  }
  // two passes over the register tiles (rows, then columns) keep the live set small: 168 registers
  // hold 120 of matrix, and the one-pass version spilled
  {
    cplx yr[T];
#pragma unroll
    for (int ti = 0; ti < T; ++ti) yr[ti] = make_c(0.0, 0.0);
#pragma unroll
    for (int tj = 0; tj < T; ++tj) {
      // (dead columns <= k have x = 0: no liveness branch, the straight-line code schedules better)
      cplx xc0 = x[c0 + 8 * tj], xc1 = x[c0 + 8 * tj + 1];  // zero for columns <= k (publish_col)
      if (c0 + 8 * tj == k + 1) xc0.x = xp0x;
      if (c0 + 8 * tj + 1 == k + 1) xc1.x = xp0x;
#pragma unroll
      for (int ti = 0; ti < T; ++ti) {
        cfma(yr[ti], make_c(a[ti][tj].re[0], a[ti][tj].im[0]), xc0);
        cfma(yr[ti], make_c(a[ti][tj].re[1], a[ti][tj].im[1]), xc1);
      }
    }
    double t = 0.0;
#pragma unroll
    for (int ti = 0; ti < T; ++ti) t = fma(xr[ti].x, yr[ti].x, fma(xr[ti].y, yr[ti].y, t));
    qacc = fma(2.0, t, qacc);
    hs_reduce_rows<T>(yr, lane, yrow_dst + TB * SI);
  }
  {
    cplx u[T];
#pragma unroll
    for (int tj = 0; tj < T; ++tj) {
      cplx yc0 = make_c(0.0, 0.0), yc1 = make_c(0.0, 0.0);
#pragma unroll
      for (int ti = 0; ti < T; ++ti) {
        ccfma(yc0, make_c(a[ti][tj].re[0], a[ti][tj].im[0]), xr[ti]);
        ccfma(yc1, make_c(a[ti][tj].re[1], a[ti][tj].im[1]), xr[ti]);
      }
      u[tj] = hs_reduce_cols1(yc0, yc1, lane);
    }
    hs_reduce_cols23<T>(u, lane, ycol_dst + TB * SJ);
  }
}

// The same over the lower tiles (tj <= ti) of the DIAGONAL superblock SI: the diagonal tiles are full
// Hermitian 8 x 8 blocks (row direction only), the tiles below them work in both directions.
// qacc += x_I^H A_II x_I (row- and column-direction partial products together cover the whole block).
// LIVE: skip dead tile columns with uniform branches (pays when one warp owns the whole matrix); the CTA
// kernel runs straight-line code instead (dead columns have x = 0), which the compiler schedules better.
template <int T, bool LIVE>
__device__ __forceinline__ void hs_matvec_diag(const HsTile (&a)[T][T], int SI, int k, const cplx *x, double xp0x,
                                               int lane, cplx *yrow_dst, cplx *ycol_dst, double &qacc) {
  constexpr int TB = 8 * T;
  const int g = lane >> 2, q = lane & 3;
  const int r0 = TB * SI + g, c0 = TB * SI + 2 * q;
  cplx xr[T];
  double t = 0.0;
#pragma unroll
  for (int ti = 0; ti < T; ++ti) {
    xr[ti] = x[r0 + 8 * ti];
    if (r0 + 8 * ti == k + 1) xr[ti].x = xp0x;
  }
  {  // row direction (pass 1)
    cplx yr[T];
#pragma unroll
    for (int ti = 0; ti < T; ++ti) yr[ti] = make_c(0.0, 0.0);
#pragma unroll
    for (int tj = 0; tj < T; ++tj) {
      if (!LIVE || TB * SI + 8 * tj + 7 > k) {  // live columns; the tile rows ti >= tj are then live too
        cplx xc0 = x[c0 + 8 * tj], xc1 = x[c0 + 8 * tj + 1];
        if (c0 + 8 * tj == k + 1) xc0.x = xp0x;
        if (c0 + 8 * tj + 1 == k + 1) xc1.x = xp0x;
#pragma unroll
        for (int ti = tj; ti < T; ++ti) {
          cfma(yr[ti], make_c(a[ti][tj].re[0], a[ti][tj].im[0]), xc0);
          cfma(yr[ti], make_c(a[ti][tj].re[1], a[ti][tj].im[1]), xc1);
        }
      }
    }
#pragma unroll
    for (int ti = 0; ti < T; ++ti) t = fma(xr[ti].x, yr[ti].x, fma(xr[ti].y, yr[ti].y, t));
    hs_reduce_rows<T>(yr, lane, yrow_dst + TB * SI);
  }
  {  // column direction of the tiles below the diagonal (pass 2)
    cplx u[T - 1];
#pragma unroll
    for (int tj = 0; tj < T - 1; ++tj) {
      cplx yc0 = make_c(0.0, 0.0), yc1 = make_c(0.0, 0.0);
      if (!LIVE || TB * SI + 8 * tj + 7 > k) {
#pragma unroll
        for (int ti = tj + 1; ti < T; ++ti) {
          ccfma(yc0, make_c(a[ti][tj].re[0], a[ti][tj].im[0]), xr[ti]);
          ccfma(yc1, make_c(a[ti][tj].re[1], a[ti][tj].im[1]), xr[ti]);
        }
        cplx xc0 = x[c0 + 8 * tj], xc1 = x[c0 + 8 * tj + 1];
        if (c0 + 8 * tj == k + 1) xc0.x = xp0x;
        if (c0 + 8 * tj + 1 == k + 1) xc1.x = xp0x;
        t = fma(xc0.x, yc0.x, fma(xc0.y, yc0.y, fma(xc1.x, yc1.x, fma(xc1.y, yc1.y, t))));
      }
      u[tj] = hs_reduce_cols1(yc0, yc1, lane);
    }
    qacc += t;
    hs_reduce_cols23<T - 1>(u, lane, ycol_dst + TB * SI);
  }
}

// A -= v w^H + w v^H on one superblock: two DMMAs per live tile.  DIAG: lower tiles only.
// Operands in shared memory, already in fragment order (written by the combine step):
//   sAr[r] = -(v.x, v.y, w.x, w.y), sAi[r] = (-v.y, v.x, -w.y, w.x), sB[c] = (w.x, w.y, v.x, v.y).
template <int T, bool DIAG, bool LIVE = false>
__device__ __forceinline__ void hs_update(HsTile (&a)[T][T], int SI, int SJ, int k, const double *sAr, const double *sAi,
                                          const double *sB, int lane) {
  constexpr int TB = 8 * T;
  double are[T], aim[T], bb[T];
#pragma unroll
  for (int ti = 0; ti < T; ++ti) {
    are[ti] = sAr[4 * (TB * SI + 8 * ti) + lane];  // [row 8 I + g][q]
    aim[ti] = sAi[4 * (TB * SI + 8 * ti) + lane];
  }
#pragma unroll
  for (int tj = 0; tj < T; ++tj) bb[tj] = sB[4 * (TB * SJ + 8 * tj) + lane];  // [column 8 J + g][q]
#pragma unroll
  for (int tj = 0; tj < T; ++tj) {
    if (!LIVE || TB * SJ + 8 * tj + 7 > k) {  // (dead columns are never read again: updating them is harmless)
#pragma unroll
      for (int ti = (DIAG ? tj : 0); ti < T; ++ti) {
        dmma884(a[ti][tj].re[0], a[ti][tj].re[1], are[ti], bb[tj]);
        dmma884(a[ti][tj].im[0], a[ti][tj].im[1], aim[ti], bb[tj]);
      }
    }
  }
}

template <int NT, int T>
__global__ void __launch_bounds__(32 * HsGeom<NT, T>::NW, HsGeom<NT, T>::CTAS)
hql_tridiag_hs_kernel(int d, int dstride, int koff, int nsteps, const cplx *__restrict__ H0,
                      const cplx *__restrict__ Z, const double *__restrict__ Bf, const cplx *__restrict__ Ain,
                      double *__restrict__ dout, double *__restrict__ eout, cplx *__restrict__ Vp, size_t vcap,
                      cplx *__restrict__ tauout, cplx *__restrict__ Aout, int mirror) {
  using G = HsGeom<NT, T>;
  constexpr int S = G::S, NW = G::NW, D = G::D, TB = G::TB;
  constexpr int RPW = (D + NW - 1) / NW;  // rows per warp in the combine step (16, 24 or 11)
  static_assert(RPW <= 32, "combine step: one row per lane");
  __shared__ __align__(16) cplx sx[2][D];        // column k of the trailing matrix, by parity of k
  __shared__ __align__(16) double sxn[2][8];     // per-warp partial ||x[2:]||^2, by parity of k
  __shared__ __align__(16) double sq[8];         // per-warp partial x'^H A x'
  __shared__ __align__(16) cplx ssc[4];          // step scalars: ts = tau scale, scale, -1/2 |tau|^2 |scale|^2
  __shared__ __align__(16) double sAr[4 * D], sAi[4 * D], sB[4 * D];  // update operands (see hs_update)
  // partial products: slot c < S of super-row R comes from superblock (R, c) (row direction) or (c, R)
  // (column direction), slot S from the column direction inside the diagonal superblock (R, R)
  __shared__ __align__(16) cplx ypart[S + 1][D];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const size_t cfg = blockIdx.x;
  const size_t dd = (size_t)d * d;

  // superblocks of this warp: the diagonal block (w, w) for w < S, and one off-diagonal block
  const bool has0 = w < S;
  const int SI0 = has0 ? w : 0;
  int SI1, SJ1;
  if (S == 4) {
    // w = 0: (3,2)  1: (3,1)  2: (2,1)  3: (3,0)  4: (1,0)  5: (2,0)
    SI1 = (w == 2 || w == 5) ? 2 : ((w == 4) ? 1 : 3);
    SJ1 = (w == 0) ? 2 : ((w == 1 || w == 2) ? 1 : 0);
  } else {
    SI1 = (w == 2) ? 1 : 2;  // w = 0: (2,1)  1: (2,0)  2: (1,0)
    SJ1 = (w == 0) ? 1 : 0;
  }

  HsTile a0[T][T], a1[T][T];  // a0: only tj <= ti is used
  {
    double bx = 0, by = 0, bz = 0;
    if (!Ain) {
      bx = Bf[cfg * 3 + 0];
      by = Bf[cfg * 3 + 1];
      bz = Bf[cfg * 3 + 2];
    }
    auto load_elem = [&](int r, int c) {
      cplx v = make_c(0.0, 0.0);
      if (r < d && c < d) {
        const size_t idx = (size_t)r * d + c;
        if (Ain) {
          v = Ain[cfg * dd + idx];
        } else {
          v = H0[idx];
          const cplx z0 = Z[idx], z1 = Z[dd + idx], z2 = Z[2 * dd + idx];
          v.x += bx * z0.x + by * z1.x + bz * z2.x;
          v.y += bx * z0.y + by * z1.y + bz * z2.y;
        }
        if (r == c) v.y = 0.0;
      }
      return v;
    };
#pragma unroll
    for (int ti = 0; ti < T; ++ti)
#pragma unroll
      for (int tj = 0; tj < T; ++tj)
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (tj <= ti) {
            cplx v = make_c(0.0, 0.0);
            if (has0) v = load_elem(TB * SI0 + 8 * ti + g, TB * SI0 + 8 * tj + 2 * q + s);
            a0[ti][tj].re[s] = v.x;
            a0[ti][tj].im[s] = v.y;
          }
          const cplx v1 = load_elem(TB * SI1 + 8 * ti + g, TB * SJ1 + 8 * tj + 2 * q + s);
          a1[ti][tj].re[s] = v1.x;
          a1[ti][tj].im[s] = v1.y;
        }
  }
  for (int i = tid; i < D; i += 32 * NW) ypart[S][i] = make_c(0.0, 0.0);

  // Publish column kc of the (updated) matrix for rows > kc, the partial norms of rows > kc + 1 and
  // the diagonal element.  In half storage column kc lives in the superblocks (., kc / 24); the tile
  // column and the fragment slot are uniform, so they are branches, not selects.
  auto publish_sb = [&](const HsTile (&a)[T][T], int SI, int kc, int tjk, bool diag, double &xn) {
    const int par = kc & 1;
    cplx v[T];
#pragma unroll
    for (int ti = 0; ti < T; ++ti) v[ti] = make_c(0.0, 0.0);
#pragma unroll
    for (int tj = 0; tj < T; ++tj)
      if (tj == tjk) {
        if (kc & 1) {
#pragma unroll
          for (int ti = 0; ti < T; ++ti)
            if (!diag || tj <= ti) v[ti] = make_c(a[ti][tj].re[1], a[ti][tj].im[1]);
        } else {
#pragma unroll
          for (int ti = 0; ti < T; ++ti)
            if (!diag || tj <= ti) v[ti] = make_c(a[ti][tj].re[0], a[ti][tj].im[0]);
        }
      }
    if (!diag) {
      // off-diagonal superblock: every row is below the column (r >= kc + 1; '=' only for the first row of the block)
#pragma unroll
      for (int ti = 0; ti < T; ++ti) {
        const int r = TB * SI + 8 * ti + g;
        sx[par][r] = v[ti];
        if (ti > 0 || r > kc + 1) xn = fma(v[ti].y, v[ti].y, fma(v[ti].x, v[ti].x, xn));  // rows >= d hold zeros
      }
    } else {
#pragma unroll
      for (int ti = 0; ti < T; ++ti) {
        const int r = TB * SI + 8 * ti + g;
        if (r > kc) sx[par][r] = v[ti];  // (tiles above the diagonal only hold rows < kc)
        if (r > kc + 1) xn = fma(v[ti].y, v[ti].y, fma(v[ti].x, v[ti].x, xn));
      }
      if (g == (kc & 7)) {  // the diagonal element sits in tile (tjk, tjk), row kc & 7
        cplx dv = make_c(0.0, 0.0);
#pragma unroll
        for (int ti = 0; ti < T; ++ti)
          if (ti == tjk) dv = v[ti];
        dout[cfg * dstride + koff + kc] = dv.x;
        sx[par][kc] = make_c(0.0, 0.0);
        if (kc > 0) sx[par][kc - 1] = make_c(0.0, 0.0);
      }
    }
  };
  auto publish_col = [&](int kc) {
    const int J = kc >> 3, SJk = J / T, tjk = J - T * SJk, qk = (kc & 7) >> 1;
    double xn = 0.0;
    const bool in0 = has0 && SI0 == SJk, in1 = SJ1 == SJk;
    if (in0 || in1) {  // uniform
      if (q == qk) {
        if (in0) publish_sb(a0, SI0, kc, tjk, true, xn);
        if (in1) publish_sb(a1, SI1, kc, tjk, false, xn);
      }
      xn += __shfl_xor_sync(0xffffffffu, xn, 4);
      xn += __shfl_xor_sync(0xffffffffu, xn, 8);
      xn += __shfl_xor_sync(0xffffffffu, xn, 16);
    }
    if (lane == qk) sxn[kc & 1][w] = xn;
  };

  publish_col(0);
  const int kend = (nsteps < d - 1) ? nsteps : d - 1;
  // the step loop exists twice: warps with / without a diagonal superblock (no per-step branches on it)
#ifdef HS_TIMING
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define HS_T(i) { long long t_ = clock64(); tacc[i] += t_ - tprev; tprev = t_; }
#else
#define HS_T(i)
#endif
  auto run = [&](auto has0_tag) {
  constexpr bool HAS0 = decltype(has0_tag)::value;
#ifdef HS_TIMING
  long long tprev = clock64();
#endif
  for (int k = 0; k < kend; ++k) {
    __syncthreads();  // #1: column k and its partial norms are visible
    HS_T(0)
    const cplx *x = sx[k & 1];
    double xn = 0.0;
#pragma unroll
    for (int i = 0; i < NW; ++i) xn += sxn[k & 1][i];
    const cplx alpha = x[k + 1];
    if (xn == 0.0 && alpha.y == 0.0) {
      const int mk = d - k - 2;
      const size_t voff = (size_t)mk * (mk - 1) / 2;  // identity reflector (uniform: every thread sees the same values)
      if (tid == 0) {
        eout[cfg * dstride + koff + k] = alpha.x;
        tauout[cfg * dstride + koff + k] = make_c(0.0, 0.0);
      }
      for (int i = tid; i < mk; i += 32 * NW) Vp[cfg * vcap + voff + i] = make_c(0.0, 0.0);
      publish_col(k + 1);
      continue;
    }
    // Householder scalars: beta (-> x'_{k+1}) redundantly per thread (rsqrt, no IEEE division); the rest by one
    // thread, handed to the combine step through shared memory (keeps them out of the registers during the mat-vec)
    const double s2 = alpha.x * alpha.x + alpha.y * alpha.y + xn;
    const double ri = rsqrt(s2);
    const double sg = (alpha.x >= 0.0) ? -1.0 : 1.0;  // sign of beta
    const double beta = sg * (s2 * ri);
    const double xp0x = alpha.x - beta;  // x'_{k+1} = alpha - beta = 1/scale (imaginary part: alpha.y)
    if (tid == 32 * (NW - 1)) {
      const double ib = sg * ri;
      const cplx tau = make_c((beta - alpha.x) * ib, -alpha.y * ib);
      const double den = __drcp_rn(xp0x * xp0x + alpha.y * alpha.y);
      const cplx scale = make_c(xp0x * den, -alpha.y * den);
      const cplx ts = cmul(tau, scale);
      // p^H v = conj(ts) scale x'^H A x' and tau conj(ts) scale = |tau|^2 |scale|^2 (real), |scale|^2 = den
      ssc[0] = ts;
      ssc[1] = scale;
      ssc[2] = make_c(-0.5 * (tau.x * tau.x + tau.y * tau.y) * den, 0.0);
      eout[cfg * dstride + koff + k] = beta;
      tauout[cfg * dstride + koff + k] = tau;
    }

    HS_T(1)
    // ---- partial products y = A22 x' and the Hermitian form x'^H A22 x' ----
    {
      double qacc = 0.0;
      if (HAS0) hs_matvec_diag<T, false>(a0, SI0, k, x, xp0x, lane, ypart[SI0], ypart[S], qacc);
      hs_matvec_off<T>(a1, SI1, SJ1, k, x, xp0x, lane, ypart[SJ1], ypart[SI1], qacc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) qacc += __shfl_xor_sync(0xffffffffu, qacc, o);
      if (lane == 0) sq[w] = qacc;
    }
    HS_T(2)
    __syncthreads();  // #2: the partial products are visible
    HS_T(3)

    if (lane < RPW && RPW * w + lane < D) {  // combine, RPW rows per warp: p = tau A v, v, a2 = -1/2 tau p^H v, w = p + a2 v -> DMMA operands
      const int r = RPW * w + lane;
      double Q = 0.0;
#pragma unroll
      for (int i = 0; i < NW; ++i) Q += sq[i];
      const cplx ts = ssc[0], scale = ssc[1];
      const double a2 = ssc[2].x * Q;
      cplx y = ypart[0][r];
#pragma unroll
      for (int c = 1; c <= S; ++c) y = cadd(y, ypart[c][r]);
      cplx xr = sx[k & 1][r];  // x[r] = 0 for r <= k
      cplx v = cmul(scale, xr);
      if (r == k + 1) v = make_c(1.0, 0.0);
      const cplx pv = cmul(ts, y);
      const cplx ww = make_c(fma(a2, v.x, pv.x), fma(a2, v.y, pv.y));
      reinterpret_cast<double4 *>(sAr)[r] = make_double4(-v.x, -v.y, -ww.x, -ww.y);
      reinterpret_cast<double4 *>(sAi)[r] = make_double4(-v.y, v.x, -ww.y, ww.x);
      reinterpret_cast<double4 *>(sB)[r] = make_double4(ww.x, ww.y, v.x, v.y);
      const int iv = r - k - 2;  // reflector k for the back-transformation
      if (iv >= 0 && r < d) {
        const int mk = d - k - 2;
        Vp[cfg * vcap + (size_t)mk * (mk - 1) / 2 + iv] = v;
      }
    }
    HS_T(4)
    __syncthreads();  // #3: the update operands are visible
    HS_T(5)

    if (HAS0) hs_update<T, true>(a0, SI0, SI0, k, sAr, sAi, sB, lane);
    hs_update<T, false>(a1, SI1, SJ1, k, sAr, sAi, sB, lane);
    HS_T(6)
    publish_col(k + 1);
    HS_T(7)
  }
  };
  if (has0)
    run(std::true_type{});
  else
    run(std::false_type{});
#ifdef HS_TIMING
  if (blockIdx.x == 0 && lane == 0)
    printf("HS_T NT %d w %d steps %d: B1 %lld scal %lld mv %lld B2 %lld comb %lld B3 %lld upd %lld pub %lld\n", NT, w, kend, tacc[0] / kend,
           tacc[1] / kend, tacc[2] / kend, tacc[3] / kend, tacc[4] / kend, tacc[5] / kend, tacc[6] / kend, tacc[7] / kend);
#endif
  if (kend == d - 1) {
    if (tid == 0) eout[cfg * dstride + koff + d - 1] = 0.0;
  } else {  // hand the trailing block to the next phase (mirror: both triangles; the next half-storage phase reads the lower one only)
    const int ds = d - kend;
    cplx *Ao = Aout + cfg * (size_t)ds * ds;
    auto store_elem = [&](int r, int c, cplx v, bool mir) {
      if (r >= kend && c >= kend && r < d && c < d) {
        Ao[(size_t)(r - kend) * ds + (c - kend)] = v;
        if (mir) Ao[(size_t)(c - kend) * ds + (r - kend)] = cconj(v);
      }
    };
#pragma unroll
    for (int ti = 0; ti < T; ++ti)
#pragma unroll
      for (int tj = 0; tj < T; ++tj)
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (tj <= ti && has0)
            store_elem(TB * SI0 + 8 * ti + g, TB * SI0 + 8 * tj + 2 * q + s, make_c(a0[ti][tj].re[s], a0[ti][tj].im[s]), tj < ti && mirror);
          store_elem(TB * SI1 + 8 * ti + g, TB * SJ1 + 8 * tj + 2 * q + s, make_c(a1[ti][tj].re[s], a1[ti][tj].im[s]), mirror != 0);
        }
  }
}

// ---------------------------------------------------------------------------------------
// Warp per matrix for d <= 8 T (T = 2, 3, 4: d <= 16 / 24 / 32): the whole matrix is ONE diagonal
// superblock, i.e. T (T + 1) / 2 register tiles per warp, no block barrier at all.  Replaces the
// shared-memory warp kernel (eigh_tridiag_warp.cuh: LSU 78 % busy, every element of A loaded twice and
// stored once per step): the matrix never leaves the registers, the rank-2 update is 2 DMMAs per tile.
// Also the last phase of the large-matrix reduction (dstride / koff as in the kernels above).
// ---------------------------------------------------------------------------------------
#define HSW_WARPS 4

template <int T>
__global__ void __launch_bounds__(32 * HSW_WARPS, (T == 4 ? 3 : 4))
hql_tridiag_hsw_kernel(int d, int64_t n, const cplx *__restrict__ H0, const cplx *__restrict__ Z,
                       const double *__restrict__ Bf, const cplx *__restrict__ Ain, double *__restrict__ dout,
                       double *__restrict__ eout, cplx *__restrict__ Vp, size_t vcap, cplx *__restrict__ tauout,
                       int dstride, int koff) {
  constexpr int D = 8 * T;
  __shared__ __align__(16) cplx sx_[HSW_WARPS][D];
  __shared__ __align__(16) cplx yp_[HSW_WARPS][2][D];  // row- / column-direction partial products
  __shared__ __align__(16) double sAr_[HSW_WARPS][4 * D], sAi_[HSW_WARPS][4 * D], sB_[HSW_WARPS][4 * D];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int64_t cfg = (int64_t)blockIdx.x * HSW_WARPS + w;
  if (cfg >= n) return;  // whole warp
  cplx *sx = sx_[w];
  cplx(*yp)[D] = yp_[w];
  double *sAr = sAr_[w], *sAi = sAi_[w], *sB = sB_[w];
  const size_t dd = (size_t)d * d;

  HsTile a[T][T];  // only tj <= ti is used
  {
    double bx = 0, by = 0, bz = 0;
    if (!Ain) {
      bx = Bf[cfg * 3 + 0];
      by = Bf[cfg * 3 + 1];
      bz = Bf[cfg * 3 + 2];
    }
#pragma unroll
    for (int ti = 0; ti < T; ++ti)
#pragma unroll
      for (int tj = 0; tj <= ti; ++tj)
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int r = 8 * ti + g, c = 8 * tj + 2 * q + s;
          cplx v = make_c(0.0, 0.0);
          if (r < d && c < d) {
            const size_t idx = (size_t)r * d + c;
            if (Ain) {
              v = Ain[cfg * dd + idx];
            } else {
              v = H0[idx];
              const cplx z0 = Z[idx], z1 = Z[dd + idx], z2 = Z[2 * dd + idx];
              v.x += bx * z0.x + by * z1.x + bz * z2.x;
              v.y += bx * z0.y + by * z1.y + bz * z2.y;
            }
            if (r == c) v.y = 0.0;
          }
          a[ti][tj].re[s] = v.x;
          a[ti][tj].im[s] = v.y;
        }
  }
  if (lane < D) {
    sx[lane] = make_c(0.0, 0.0);
    yp[1][lane] = make_c(0.0, 0.0);  // the last tile has no column-direction part
  }
  __syncwarp();

  for (int k = 0; k < d; ++k) {
    // column k of the (updated) matrix: rows > k to sx, the norm of rows > k + 1, the diagonal element
    double xn = 0.0;
    {
      const int tjk = k >> 3, qk = (k & 7) >> 1;
      if (q == qk) {
        cplx v[T];
#pragma unroll
        for (int ti = 0; ti < T; ++ti) v[ti] = make_c(0.0, 0.0);
#pragma unroll
        for (int tj = 0; tj < T; ++tj)
          if (tj == tjk) {
            if (k & 1) {
#pragma unroll
              for (int ti = tj; ti < T; ++ti) v[ti] = make_c(a[ti][tj].re[1], a[ti][tj].im[1]);
            } else {
#pragma unroll
              for (int ti = tj; ti < T; ++ti) v[ti] = make_c(a[ti][tj].re[0], a[ti][tj].im[0]);
            }
          }
#pragma unroll
        for (int ti = 0; ti < T; ++ti) {
          const int r = 8 * ti + g;
          if (r > k) sx[r] = v[ti];
          if (r > k + 1) xn = fma(v[ti].y, v[ti].y, fma(v[ti].x, v[ti].x, xn));  // rows >= d hold zeros
          if (r == k) {
            dout[cfg * dstride + koff + k] = v[ti].x;
            sx[k] = make_c(0.0, 0.0);
          }
        }
      }
    }
    if (k == d - 1) break;
    xn = warp_sum(xn);
    __syncwarp();
    const cplx alpha = sx[k + 1];
    const int mk = d - k - 2;
    cplx *vrow = Vp + (cfg * vcap + (size_t)mk * (mk - 1) / 2);
    if (xn == 0.0 && alpha.y == 0.0) {  // identity reflector (uniform)
      if (lane == 0) {
        eout[cfg * dstride + koff + k] = alpha.x;
        tauout[cfg * dstride + koff + k] = make_c(0.0, 0.0);
      }
      for (int i = lane; i < mk; i += 32) vrow[i] = make_c(0.0, 0.0);
      __syncwarp();
      continue;
    }
    const double s2 = alpha.x * alpha.x + alpha.y * alpha.y + xn;
    const double ri = rsqrt(s2);
    const double sg = (alpha.x >= 0.0) ? -1.0 : 1.0;  // sign of beta
    const double beta = sg * (s2 * ri);
    const double xp0x = alpha.x - beta;  // x'_{k+1} = alpha - beta = 1/scale (imaginary part: alpha.y)
    double qacc = 0.0;
    hs_matvec_diag<T, true>(a, 0, k, sx, xp0x, lane, yp[0], yp[1], qacc);
    const double Q = warp_sum(qacc);  // x'^H A x'
    __syncwarp();
    {
      const double ib = sg * ri;
      const cplx tau = make_c((beta - alpha.x) * ib, -alpha.y * ib);
      const double den = __drcp_rn(xp0x * xp0x + alpha.y * alpha.y);
      const cplx scale = make_c(xp0x * den, -alpha.y * den);
      const cplx ts = cmul(tau, scale);
      const double a2 = -0.5 * (tau.x * tau.x + tau.y * tau.y) * den * Q;  // -1/2 tau p^H v (see the CTA kernel)
      if (lane < D) {
        const int r = lane;
        const cplx y = cadd(yp[0][r], yp[1][r]);
        cplx v = cmul(scale, sx[r]);  // x[r] = 0 for r <= k
        if (r == k + 1) v = make_c(1.0, 0.0);
        const cplx pv = cmul(ts, y);
        const cplx ww = make_c(fma(a2, v.x, pv.x), fma(a2, v.y, pv.y));
        reinterpret_cast<double4 *>(sAr)[r] = make_double4(-v.x, -v.y, -ww.x, -ww.y);
        reinterpret_cast<double4 *>(sAi)[r] = make_double4(-v.y, v.x, -ww.y, ww.x);
        reinterpret_cast<double4 *>(sB)[r] = make_double4(ww.x, ww.y, v.x, v.y);
        const int iv = r - k - 2;
        if (iv >= 0 && r < d) vrow[iv] = v;
      }
      if (lane == 0) {
        eout[cfg * dstride + koff + k] = beta;
        tauout[cfg * dstride + koff + k] = tau;
      }
    }
    __syncwarp();
    hs_update<T, true, true>(a, 0, 0, k, sAr, sAi, sB, lane);
  }
  if (lane == 0) eout[cfg * dstride + koff + d - 1] = 0.0;
}

}  // namespace musim
