// eigh_tridiag_rw1.cuh -- K1 for d <= 96, ONE block barrier per Householder step.
//
// Same data layout as eigh_tridiag_rw.cuh (A in registers; a warp owns 8 rows, half-warp `hf` holds
// 4 of them, lane l16 the columns l16 + 16 jj; warp-local mat-vec), but the step is re-ordered so
// that the rank-2 update of step k-1 and the mat-vec of step k are ONE pass over the registers
// between two barriers.  ncu on the two-barrier kernel at d = 96: 5 400 cycles per step for
// ~2 000 cycles of FP64 work -- the step is a chain of two barrier-separated serial sections
// (publish column -> barrier -> scalars, mat-vec, butterfly -> publish v, p -> barrier -> update).
//
// What makes one barrier enough: the NEXT column does not have to be read out of the updated
// matrix.  With the update of step k-1 still pending (registers hold A_{k-1}),
//     A_k[c][k] = conj(A_{k-1}[k][c]) - v_c conj(w_k) - w_c conj(v_k),      v = v_{k-1}, w = w_{k-1},
// so if the OWNER of row k publishes that row of A_{k-1} together with v_{k-1}, p_{k-1} (before the
// barrier of step k-1), every thread can form the entries of column k it needs by itself: its own
// CJ columns for the mat-vec, and -- a half-warp holds all columns -- the norm and the Householder
// scalars of step k with 4 shuffle stages and no barrier.  Step k between its two barriers:
//   a2, w of the pending reflector;  x = column k (own columns);  |x|^2 -> shuffles -> beta, tau ...
//   (overlapped with) ONE pass: a -= v w^H + w v^H (step k-1), y += a x (step k)
//   butterfly;  publish v_k, p_k, the partial p^H v and row k+1 of A_k;  barrier.
// All shared buffers are double-buffered by the parity of k (a warp can run at most one barrier
// ahead).  Outputs and packed-reflector format are those of hql_tridiag_rw_kernel; the multi-launch
// phases (trailing block handed to a smaller instantiation) work the same way, the pending update
// is applied before the hand-off.
#pragma once
#include "common.cuh"
#include "eigh_tridiag_rw.cuh"  // shfl_xor_c

namespace musim {

template <int D>
__global__ void __launch_bounds__(4 * D, (D <= 32 ? 4 : (D <= 64 ? 2 : 1)))
hql_tridiag_rw1_kernel(int d, int dstride, int koff, int nsteps, const cplx *__restrict__ H0,
                       const cplx *__restrict__ Z, const double *__restrict__ Bf,
                       const cplx *__restrict__ Ain, double *__restrict__ dout,
                       double *__restrict__ eout, cplx *__restrict__ Vp, size_t vcap,
                       cplx *__restrict__ tauout, cplx *__restrict__ Aout) {
  constexpr int CJ = D / 16;  // column blocks per thread
  constexpr int NW = D / 8;   // warps
  constexpr int RG = 2 * NW;  // rows per row group
  __shared__ __align__(16) cplx sv[2][D];    // v of the pending reflector, by parity of its step
  __shared__ __align__(16) cplx sp[2][D];    // p = tau A v
  __shared__ __align__(16) cplx srow[2][D];  // row k of A_{k-1} (the matrix BEFORE the pending update)
  __shared__ __align__(16) cplx sdot[2][NW];
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
  // publish row kr of the matrix held in the registers into srow[q]
  auto publish_row = [&](int kr, int q) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (rb + RG * i == kr) {
#pragma unroll
        for (int jj = 0; jj < CJ; ++jj) srow[q][l16 + 16 * jj] = a[i][jj];
      }
  };
  for (int i = tid; i < D; i += 4 * D) {
    sv[0][i] = sv[1][i] = make_c(0.0, 0.0);
    sp[0][i] = sp[1][i] = make_c(0.0, 0.0);
  }
  if (tid < NW) sdot[0][tid] = sdot[1][tid] = make_c(0.0, 0.0);
  publish_row(0, 0);
  __syncthreads();

  const int j_hi = (d + 15) >> 4;
  const int kend = (nsteps < d - 1) ? nsteps : d - 1;
  cplx tau_prev = make_c(0.0, 0.0);
  // entry index k of column k of A_k from the published row and the pending reflector (sv/sp[q], a2)
  auto col_entry = [&](int q, int c, double a2, cplx vk, cplx wk) {
    const cplx r0 = srow[q][c], vc = sv[q][c], pc = sp[q][c];
    const cplx wc = make_c(fma(a2, vc.x, pc.x), fma(a2, vc.y, pc.y));
    cplx x = make_c(r0.x, -r0.y);
    // x -= vc conj(wk) + wc conj(vk)
    x.x = fma(-vc.x, wk.x, x.x);
    x.x = fma(-vc.y, wk.y, x.x);
    x.x = fma(-wc.x, vk.x, x.x);
    x.x = fma(-wc.y, vk.y, x.x);
    x.y = fma(-vc.y, wk.x, x.y);
    x.y = fma(vc.x, wk.y, x.y);
    x.y = fma(-wc.y, vk.x, x.y);
    x.y = fma(wc.x, vk.y, x.y);
    return x;
  };
  auto pending_a2 = [&](int q) {
    cplx t[NW];
#pragma unroll
    for (int u = 0; u < NW; ++u) t[u] = sdot[q][u];
#pragma unroll
    for (int s = 1; s < NW; s *= 2)
#pragma unroll
      for (int u = 0; u + s < NW; u += 2 * s) t[u] = cadd(t[u], t[u + s]);
    // a2 = -1/2 tau p^H v is real for Hermitian A
    return -0.5 * (tau_prev.x * t[0].x - tau_prev.y * t[0].y);
  };

  for (int k = 0; k < kend; ++k) {
    const int q = k & 1, q2 = q ^ 1;
    const double a2 = pending_a2(q);
    // entries at index k of the pending reflector (zero before the first step)
    const cplx vk = sv[q][k];
    const cplx wk = make_c(fma(a2, vk.x, sp[q][k].x), fma(a2, vk.y, sp[q][k].y));
    // reflector k-1 for hql_reflect / hql_backwy (entries rows k+1 ..)
    if (k > 0) {
      const int mk = d - k - 1;
      const size_t voff = (size_t)mk * (mk - 1) / 2;
      for (int i = tid; i < mk; i += 4 * D) Vp[cfg * vcap + voff + i] = sv[q][k + 1 + i];
    }
    // ---- column k: diagonal, alpha, norm of the rest ----
    const cplx xk = col_entry(q, k, a2, vk, wk);
    const cplx alpha = col_entry(q, (k + 1 < D) ? k + 1 : k, a2, vk, wk);
    if (tid == 0) dout[cfg * dstride + koff + k] = xk.x;
    double xn = 0.0;
#pragma unroll
    for (int jj = 0; jj < CJ; ++jj) {
      const int c = l16 + 16 * jj;
      if (jj < j_hi && c > k + 1) {
        const cplx x = col_entry(q, c, a2, vk, wk);
        xn = fma(x.y, x.y, fma(x.x, x.x, xn));
      }
    }
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) xn += __shfl_xor_sync(0xffffffffu, xn, o);
    // ---- Householder scalars (redundantly per thread; rsqrt / rcp, no IEEE division) ----
    const bool ident = (xn == 0.0 && alpha.y == 0.0);  // H = I (uniform: every thread sees the same values)
    const double s2 = alpha.x * alpha.x + alpha.y * alpha.y + xn;
    const double ri = ident ? 0.0 : rsqrt(s2);
    const double sg = (alpha.x >= 0.0) ? -1.0 : 1.0;  // sign of beta
    const double beta = ident ? alpha.x : sg * (s2 * ri);
    const double ib = sg * ri;
    const cplx tau = ident ? make_c(0.0, 0.0) : make_c((beta - alpha.x) * ib, -alpha.y * ib);
    const cplx xp0 = make_c(alpha.x - beta, alpha.y);  // x'_{k+1} = alpha - beta = 1 / scale
    const double den = ident ? 0.0 : __drcp_rn(xp0.x * xp0.x + xp0.y * xp0.y);
    const cplx scale = make_c(xp0.x * den, -xp0.y * den);
    const cplx ts = cmul(tau, scale);
    if (tid == 0) {
      eout[cfg * dstride + koff + k] = beta;
      tauout[cfg * dstride + koff + k] = tau;
    }
    // ---- one pass: A -= v w^H + w v^H (pending, step k-1);  y = A x (step k, x_{k+1} added below) ----
    const int i_lo = (k + 1) / RG;  // row groups below are dead (rows <= k)
    const int j_lo = (k + 1) >> 4;  // column blocks below are dead (columns <= k)
    cplx vr[4], wr[4], y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      vr[i] = sv[q][rb + RG * i];
      const cplx pi = sp[q][rb + RG * i];
      wr[i] = make_c(fma(a2, vr[i].x, pi.x), fma(a2, vr[i].y, pi.y));
      y[i] = make_c(0.0, 0.0);
    }
#pragma unroll
    for (int jj = 0; jj < CJ; ++jj) {
      if (jj >= j_lo && jj < j_hi) {
        const int c = l16 + 16 * jj;
        const cplx vc = sv[q][c], pc = sp[q][c];
        const cplx wc = make_c(fma(a2, vc.x, pc.x), fma(a2, vc.y, pc.y));
        cplx xc = col_entry(q, c, a2, vk, wk);
        if (c <= k + 1) xc = make_c(0.0, 0.0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i >= i_lo) {
            cplx &e = a[i][jj];
            e.x = fma(-vr[i].x, wc.x, e.x);
            e.x = fma(-vr[i].y, wc.y, e.x);
            e.x = fma(-wr[i].x, vc.x, e.x);
            e.x = fma(-wr[i].y, vc.y, e.x);
            e.y = fma(-vr[i].y, wc.x, e.y);
            e.y = fma(vr[i].x, wc.y, e.y);
            e.y = fma(-wr[i].y, vc.x, e.y);
            e.y = fma(wr[i].x, vc.y, e.y);
            cfma(y[i], e, xc);
          }
        }
      }
    }
    // the lane that holds column k+1 adds its term with x'_{k+1} = xp0 (the only entry that needs the scalars)
    {
      const int j1 = (k + 1) >> 4;
      if (l16 == ((k + 1) & 15)) {
#pragma unroll
        for (int jj = 0; jj < CJ; ++jj)
          if (jj == j1) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (i >= i_lo) cfma(y[i], a[i][jj], xp0);
          }
      }
    }
    // row k+1 of A_k for the next step's column
    publish_row(k + 1, q2);
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
        cplx vv = make_c(1.0, 0.0);
        if (r != k + 1) vv = cmul(scale, col_entry(q, r, a2, vk, wk));
        const cplx pr = cmul(ts, y[0]);
        sv[q2][r] = vv;
        sp[q2][r] = pr;
        dt = ccmul(pr, vv);  // conj(p) v
      }
#pragma unroll
      for (int o = 16; o >= 4; o >>= 1) dt = cadd(dt, shfl_xor_c(dt, o));
      if (lane == 0) sdot[q2][w] = dt;
    }
    tau_prev = tau;
    __syncthreads();
  }

  // ---- after the last step: reflector kend-1 is still pending in buffer kend & 1 ----
  {
    const int k = kend, q = k & 1;
    const double a2 = pending_a2(q);
    if (k > 0) {
      const int mk = d - k - 1;
      const size_t voff = (size_t)mk * (mk - 1) / 2;
      for (int i = tid; i < mk; i += 4 * D) Vp[cfg * vcap + voff + i] = sv[q][k + 1 + i];
    }
    if (kend == d - 1) {
      if (tid == 0) {
        const cplx vk = sv[q][k];
        const cplx wk = make_c(fma(a2, vk.x, sp[q][k].x), fma(a2, vk.y, sp[q][k].y));
        const cplx xk = col_entry(q, k, a2, vk, wk);
        dout[cfg * dstride + koff + d - 1] = xk.x;
        eout[cfg * dstride + koff + d - 1] = 0.0;
      }
    } else {  // apply the pending update and hand the trailing block to the next phase
      const int i_lo = k / RG, j_lo = k >> 4;
      cplx vr[4], wr[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        vr[i] = sv[q][rb + RG * i];
        const cplx pi = sp[q][rb + RG * i];
        wr[i] = make_c(fma(a2, vr[i].x, pi.x), fma(a2, vr[i].y, pi.y));
      }
#pragma unroll
      for (int jj = 0; jj < CJ; ++jj) {
        if (jj >= j_lo && jj < j_hi) {
          const int c = l16 + 16 * jj;
          const cplx vc = sv[q][c], pc = sp[q][c];
          const cplx wc = make_c(fma(a2, vc.x, pc.x), fma(a2, vc.y, pc.y));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (i >= i_lo) {
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
}

}  // namespace musim
