// eigh_tridiag_rw.cuh -- K1 for d <= 96: Householder tridiagonalisation with the matrix in
// REGISTERS, rows owned by warps, TWO block barriers per Householder step.
//
// Why (ncu, profiles/): the shared-memory kernel (eigh_hql.cuh) spends a d = 96 step in ~6 k
// cycles: 50 % LSU, 22 % FP64, barrier + latency stalls (8 barriers per step).  The first
// register-resident attempt (eigh_tridiag_reg.cuh: a thread owns one row across 4 column
// groups) still needs 4 barriers per step and was slower.  Here:
//   * a warp owns 8 rows completely: half-warp `hf` holds 4 of them, lane l16 the
//     columns l16 + 16 jj.  The mat-vec y = A x' is then WARP-LOCAL (butterfly reduction over
//     the 16 lanes that share a row), no barrier;
//   * p = tau A v, v and the warp's share of p^H v are published, barrier #2, and every thread
//     builds w_c = p_c + a2 v_c for its own columns itself (a2 = -1/2 tau p^H v);
//   * after the update every thread that holds an element of COLUMN k+1 publishes it (and its
//     share of the norm): the column is spread over all warps, so there is no serial owner
//     publishes column k+1 = conj(row k+1) (Hermitian symmetry keeps every register index
//     static), barrier #1.
// Work still shrinks with the trailing block: column blocks of 16 and row groups drop out.
// Outputs d, e, tau and the reflectors packed for hql_reflect_kernel; Q is never formed.
// Arithmetic identical to tools/hql_prototype.py::tridiag_lower (zhetd2, lower).
#pragma once
#include "common.cuh"

namespace musim {

__device__ __forceinline__ cplx shfl_xor_c(cplx v, int m) {
  return make_c(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}

// Row r lives in row group i = r / RG (RG = 2 NW rows), half-warp hf = (r % RG) / NW, warp
// w = r % NW: the rows of a warp are spread CYCLICALLY over the matrix, so every warp keeps the
// same number of live rows while the trailing block shrinks (whole row groups and column blocks
// drop out with uniform branches) instead of whole warps going idle.
//
// PHASES.  Late steps are dominated by the fixed per-step latency (two barriers, the scalar
// chain, the butterflies) while the live block needs only a fraction of the registers, so the
// host runs the reduction as up to three launches: the kernel stops after `nsteps` steps and
// writes the trailing (d - nsteps)^2 block to Aout; the next launch (a smaller D: more CTAs per
// SM) continues on that block -- the trailing problem is self-similar, the packed reflectors
// keep their offsets, and d / e / tau are written at offset `koff` of rows of length `dstride`.
template <int D>
__global__ void __launch_bounds__(4 * D, (D <= 32 ? 4 : (D <= 48 ? 3 : (D <= 64 ? 2 : 1))))
hql_tridiag_rw_kernel(int d, int dstride, int koff, int nsteps, const cplx *__restrict__ H0,
                      const cplx *__restrict__ Z, const double *__restrict__ Bf,
                      const cplx *__restrict__ Ain, double *__restrict__ dout,
                      double *__restrict__ eout, cplx *__restrict__ Vp, size_t vcap,
                      cplx *__restrict__ tauout, cplx *__restrict__ Aout) {
  constexpr int CJ = D / 16;  // column blocks per thread
  constexpr int NW = D / 8;   // warps
  constexpr int RG = 2 * NW;  // rows per row group
  __shared__ __align__(16) cplx sx[2][D];      // column k of the trailing matrix, by parity of k
  __shared__ __align__(16) double sxn[2][NW];  // per-warp partial ||x[2:]||^2, by parity of k
  __shared__ __align__(16) cplx sv[D];         // v of the current step
  __shared__ __align__(16) cplx sp[D];         // p = tau A v
  __shared__ __align__(16) cplx sdot[NW];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int hf = lane >> 4, l16 = lane & 15;
  const int rb = w + NW * hf;  // this thread's rows rb + RG i, i = 0 .. 3
  const size_t cfg = blockIdx.x;
  const size_t dd = (size_t)d * d;

  cplx a[4][CJ];
  {
    double bx = 0, by = 0, bz = 0;
    if (!Ain) {
      bx = Bf[cfg * 3 + 0];
      by = Bf[cfg * 3 + 1];
      bz = Bf[cfg * 3 + 2];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jj = 0; jj < CJ; ++jj) {
        const int r = rb + RG * i, c = l16 + 16 * jj;
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
        a[i][jj] = v;
      }
  }

  // Publish column kc of the (updated) matrix: every row's element A[r][kc] sits in the lane
  // with l16 == kc % 16 of the half-warp that owns row r, so the column and the partial norms
  // come from ALL warps in parallel (no serial owner section).  Only live rows (r > kc) are
  // written; the owner of the diagonal zeroes the two entries that died since this parity was
  // last published, so the mat-vec below needs no masks.
  auto publish_col = [&](int kc) {
    double xn = 0.0;
    const int j1 = kc >> 4;  // column block (uniform)
    if (l16 == (kc & 15)) {
#pragma unroll
      for (int jj = 0; jj < CJ; ++jj) {
        if (jj == j1) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = rb + RG * i;
            const cplx v = a[i][jj];
            if (r > kc) sx[kc & 1][r] = v;
            if (r > kc + 1) xn = fma(v.y, v.y, fma(v.x, v.x, xn));  // rows >= d hold zeros
            if (r == kc) {
              dout[cfg * dstride + koff + kc] = v.x;
              sx[kc & 1][kc] = make_c(0.0, 0.0);
              if (kc > 0) sx[kc & 1][kc - 1] = make_c(0.0, 0.0);
            }
          }
        }
      }
    }
    xn += __shfl_xor_sync(0xffffffffu, xn, 16);
    if (lane == (kc & 15)) sxn[kc & 1][w] = xn;
  };

  for (int i = tid; i < D; i += 4 * D) {
    sv[i] = make_c(0.0, 0.0);
    sp[i] = make_c(0.0, 0.0);
  }
  publish_col(0);
  const int j_hi = (d + 15) >> 4;
  const int kend = (nsteps < d - 1) ? nsteps : d - 1;
  for (int k = 0; k < kend; ++k) {
    __syncthreads();  // #1: column k and its partial norms are visible
    const cplx *x = sx[k & 1];
    double xn;
    {
      double t[NW];  // pairwise tree: depth log2(NW) instead of a chain of NW dependent adds
#pragma unroll
      for (int q = 0; q < NW; ++q) t[q] = sxn[k & 1][q];
#pragma unroll
      for (int s = 1; s < NW; s *= 2)
#pragma unroll
        for (int q = 0; q + s < NW; q += 2 * s) t[q] += t[q + s];
      xn = t[0];
    }
    const cplx alpha = x[k + 1];
    const int mk = d - k - 2;
    const size_t voff = (size_t)mk * (mk - 1) / 2;
    if (xn == 0.0 && alpha.y == 0.0) {  // identity reflector (uniform: every thread sees the same values)
      if (tid == 0) {
        eout[cfg * dstride + koff + k] = alpha.x;
        tauout[cfg * dstride + koff + k] = make_c(0.0, 0.0);
      }
      for (int i = tid; i < mk; i += 4 * D) Vp[cfg * vcap + voff + i] = make_c(0.0, 0.0);
      publish_col(k + 1);
      continue;
    }
    // Householder scalars, redundantly per thread (rsqrt / rcp, no IEEE division); the chain
    // overlaps with the mat-vec below, which does not need them
    const double s2 = alpha.x * alpha.x + alpha.y * alpha.y + xn;
    const double ri = rsqrt(s2);
    const double sg = (alpha.x >= 0.0) ? -1.0 : 1.0;  // sign of beta
    const double beta = sg * (s2 * ri);
    const double ib = sg * ri;
    const cplx tau = make_c((beta - alpha.x) * ib, -alpha.y * ib);
    const cplx xp0 = make_c(alpha.x - beta, alpha.y);  // x'_{k+1} = alpha - beta = 1/scale
    const double den = __drcp_rn(xp0.x * xp0.x + xp0.y * xp0.y);
    const cplx scale = make_c(xp0.x * den, -xp0.y * den);
    const cplx ts = cmul(tau, scale);
    if (tid == 0) {
      eout[cfg * dstride + koff + k] = beta;
      tauout[cfg * dstride + koff + k] = tau;
    }
    const int i_lo = (k + 1) / RG;  // row groups below are dead (rows <= k)
    const int j_lo = (k + 1) >> 4;  // column blocks below are dead (columns <= k)
    // ---- y = A22 x'  (warp-local) ----
    cplx y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = make_c(0.0, 0.0);
#pragma unroll
    for (int jj = 0; jj < CJ; ++jj) {
      if (jj >= j_lo && jj < j_hi) {  // block has columns > k (uniform)
        const int c = l16 + 16 * jj;
        cplx xv = x[c];            // zero for c <= k (publish_col)
        if (c == k + 1) xv = xp0;  // the only term that needs the scalar chain
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i >= i_lo) cfma(y[i], a[i][jj], xv);
      }
    }
    // butterfly over the 16 lanes: 4 rows -> lane (b3, b2) ends up with row group 2 b3 + b2
    {
      const bool up = (l16 & 8) != 0;
      const cplx s0 = shfl_xor_c(up ? y[0] : y[2], 8);
      const cplx s1 = shfl_xor_c(up ? y[1] : y[3], 8);
      y[0] = cadd(up ? y[2] : y[0], s0);
      y[1] = cadd(up ? y[3] : y[1], s1);
    }
    {
      const bool up = (l16 & 4) != 0;
      const cplx s0 = shfl_xor_c(up ? y[0] : y[1], 4);
      y[0] = cadd(up ? y[1] : y[0], s0);
    }
    y[0] = cadd(y[0], shfl_xor_c(y[0], 2));
    y[0] = cadd(y[0], shfl_xor_c(y[0], 1));
    {
      const int r = rb + RG * (((l16 >> 3) & 1) * 2 + ((l16 >> 2) & 1));
      cplx dt = make_c(0.0, 0.0);
      if ((l16 & 3) == 0 && r > k) {
        const cplx vr = (r == k + 1) ? make_c(1.0, 0.0) : cmul(scale, x[r]);
        const cplx pr = cmul(ts, y[0]);
        sv[r] = vr;
        sp[r] = pr;
        dt = ccmul(pr, vr);  // conj(p) v
      }
#pragma unroll
      for (int o = 16; o >= 4; o >>= 1) dt = cadd(dt, shfl_xor_c(dt, o));
      if (lane == 0) sdot[w] = dt;
    }
    __syncthreads();  // #2: v, p and the partial dot products are visible
    // reflector k for hql_reflect_kernel
    for (int i = tid; i < mk; i += 4 * D) Vp[cfg * vcap + voff + i] = sv[k + 2 + i];
    double a2;
    {
      cplx t[NW];
#pragma unroll
      for (int q = 0; q < NW; ++q) t[q] = sdot[q];
#pragma unroll
      for (int s = 1; s < NW; s *= 2)
#pragma unroll
        for (int q = 0; q + s < NW; q += 2 * s) t[q] = cadd(t[q], t[q + s]);
      // a2 = -1/2 tau p^H v = -1/2 |tau|^2 v^H A v is real for Hermitian A (its imaginary part
      // is rounding noise), so w = p + a2 v costs two FMAs per entry
      a2 = -0.5 * (tau.x * t[0].x - tau.y * t[0].y);
    }
    cplx vr[4], wr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // rows / columns <= k are dead (never read again): they may take the stale, bounded
      // v / p values without masking
      vr[i] = sv[rb + RG * i];
      const cplx pi = sp[rb + RG * i];
      wr[i] = make_c(fma(a2, vr[i].x, pi.x), fma(a2, vr[i].y, pi.y));
    }
#pragma unroll
    for (int jj = 0; jj < CJ; ++jj) {
      if (jj >= j_lo && jj < j_hi) {
        const int c = l16 + 16 * jj;
        const cplx vc = sv[c];
        const cplx pc = sp[c];
        const cplx wc = make_c(fma(a2, vc.x, pc.x), fma(a2, vc.y, pc.y));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i >= i_lo) {  // A -= v w^H + w v^H, eight FMAs per element
            cplx &e = a[i][jj];
            e.x = fma(-vr[i].x, wc.x, e.x);
            e.x = fma(-vr[i].y, wc.y, e.x);
            e.x = fma(-wr[i].x, vc.x, e.x);
            e.x = fma(-wr[i].y, vc.y, e.x);
            e.y = fma(-vr[i].y, wc.x, e.y);
            e.y = fma(vr[i].x, wc.y, e.y);
            e.y = fma(-wr[i].y, vc.x, e.y);
            e.y = fma(wr[i].x, vc.y, e.y);
          }
        }
      }
    }
    publish_col(k + 1);
  }
  if (kend == d - 1) {
    if (tid == 0) eout[cfg * dstride + koff + d - 1] = 0.0;
  } else {  // hand the trailing block to the next phase
    const int ds = d - kend;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jj = 0; jj < CJ; ++jj) {
        const int r = rb + RG * i, c = l16 + 16 * jj;
        if (r >= kend && c >= kend && r < d && c < d)
          Aout[cfg * (size_t)ds * ds + (size_t)(r - kend) * ds + (c - kend)] = a[i][jj];
      }
  }
}

}  // namespace musim
