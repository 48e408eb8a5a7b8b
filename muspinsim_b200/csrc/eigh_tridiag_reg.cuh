// eigh_tridiag_reg.cuh -- K1, register-resident variant (d <= 96): Householder
// tridiagonalisation + explicit Q with the MATRIX IN REGISTERS.
//
// ncu on the shared-memory version (eigh_hql.cuh, hql_tridiag_kernel): 50 % LSU, 22 % FP64,
// ~870k cycles per d = 96 matrix -- every complex FMA of the mat-vec and of the rank-2 update
// moves 16-32 B through shared memory.  Here thread (r, g) of a 4R-thread CTA keeps the 24
// (d = 96) elements A[r][g + 4 jj] of its row in registers (36.8k of the SM's 64k registers), so
// the two O(m^2) loops of each Householder step run out of the register file; shared memory
// only carries the O(m) vectors (x, v, w, the partial sums).  Column k is read through the
// Hermitian symmetry as conj(row k), which keeps every register index static.  The unrolled
// column loops are entered through a fall-through switch at the first active column, so the
// work still shrinks as m^2.
//
// Q = H_0 ... H_{d-2} is then accumulated backwards in the same registers with the transposed
// distribution (thread (c, g) holds Q[g + 4 jj][c]) so that v^H Q is again thread-local.
//
// Same arithmetic as tools/hql_prototype.py (tridiag_lower / form_q_inplace).
#pragma once
#include "common.cuh"

namespace musim {

#define TRR_G 4

#define TRR_CASES(X)                                                                          \
  X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) \
      X(17) X(18) X(19) X(20) X(21) X(22) X(23)

template <int R, int CPT>
__global__ void __launch_bounds__(TRR_G *R)
hql_tridiag_reg_kernel(int d, const cplx *__restrict__ H0, const cplx *__restrict__ Z,
                       const double *__restrict__ Bf, const cplx *__restrict__ Ain,
                       double *__restrict__ dout, double *__restrict__ eout, cplx *__restrict__ Qout) {
  static_assert(CPT * TRR_G == R, "R = 4 * CPT");
  static_assert(CPT <= 24, "extend TRR_CASES");
  constexpr int G = TRR_G;
  constexpr int NWG = (R + 31) / 32;  // warps per column group
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx *sV = reinterpret_cast<cplx *>(smem_raw);  // reflectors: v_k[r] at [k*R + r]
  cplx *sx = sV + (size_t)R * R;                   // [R]
  cplx *sv = sx + R;                               // [R]
  cplx *sw = sv + R;                               // [R]
  cplx *stau = sw + R;                             // [R]
  cplx *spart = stau + R;                          // [2][G][R]
  cplx *sdot = spart + 2 * G * R;                  // [NWG]
  const int tid = threadIdx.x;
  const int r = tid % R, g = tid / R;
  const int lane = tid & 31;
  const size_t cfg = blockIdx.x;
  const size_t dd = (size_t)d * d;

  // ---- load row r, columns g + 4 jj ----
  cplx a[CPT];
  {
    double bx = 0, by = 0, bz = 0;
    if (!Ain) {
      bx = Bf[cfg * 3 + 0];
      by = Bf[cfg * 3 + 1];
      bz = Bf[cfg * 3 + 2];
    }
#pragma unroll
    for (int jj = 0; jj < CPT; ++jj) {
      const int c = g + G * jj;
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
      a[jj] = v;
    }
  }

  // ---- zhetd2 (lower) ----
  // Per step: (#1) column k published; warp 0 derives the Householder scalars (sqrt + 3
  // divisions, ~200 FP64 instructions: done ONCE, not by every thread) while all warps run
  // the mat-vec without the column-(k+1) term, which needs x'_{k+1} = alpha - beta;
  // (#2) g = 0 threads finish p, v and the dot product; (#3) w; (#4) rank-2 update.
  double *ssc = reinterpret_cast<double *>(sdot + 8);  // [12] scalars of the current step
  cplx *sfirst = reinterpret_cast<cplx *>(ssc + 12);   // [R] a[r][k+1]
  for (int k = 0; k < d - 1; ++k) {
    if (r == k) {
#pragma unroll
      for (int jj = 0; jj < CPT; ++jj) sx[g + G * jj] = cconj(a[jj]);
    }
    __syncthreads();  // #1
    if (tid < 32) {
      double xn = 0.0;
      for (int i = k + 2 + lane; i < d; i += 32) xn += cnorm2(sx[i]);
      xn = warp_sum(xn);
      if (lane == 0) {
        const cplx alpha = sx[k + 1];
        dout[cfg * d + k] = sx[k].x;
        if (xn == 0.0 && alpha.y == 0.0) {  // H_k = I
          eout[cfg * d + k] = alpha.x;
          stau[k] = make_c(0.0, 0.0);
          ssc[0] = 0.0;
        } else {
          const double beta = -copysign(sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xn), alpha.x);
          const double ib = 1.0 / beta;
          const cplx tau = make_c((beta - alpha.x) * ib, -alpha.y * ib);
          const cplx xp0 = make_c(alpha.x - beta, alpha.y);
          const double den = 1.0 / (xp0.x * xp0.x + xp0.y * xp0.y);
          const cplx scale = make_c(xp0.x * den, -xp0.y * den);
          const cplx ts = cmul(tau, scale);
          eout[cfg * d + k] = beta;
          stau[k] = tau;
          ssc[0] = 1.0;
          ssc[1] = tau.x; ssc[2] = tau.y;
          ssc[3] = xp0.x; ssc[4] = xp0.y;
          ssc[5] = scale.x; ssc[6] = scale.y;
          ssc[7] = ts.x; ssc[8] = ts.y;
        }
      }
    }
    const int jmin = (k >= g) ? (k - g) / G + 1 : 0;  // first jj with column g + 4 jj > k
    if (r > k) {
      cplx y = make_c(0.0, 0.0);
      switch (jmin) {
#define X(J)                                         \
  case J:                                            \
    if (J < CPT) {                                   \
      const int c = g + G * J;                       \
      if (c == k + 1)                                \
        sfirst[r] = a[J < CPT ? J : 0];              \
      else                                           \
        cfma(y, a[J < CPT ? J : 0], sx[c]);          \
    }
        TRR_CASES(X)
#undef X
        default:
          break;
      }
      spart[g * R + r] = y;
    }
    __syncthreads();  // #2
    if (ssc[0] == 0.0) {  // identity reflector: nothing to update
      if (g == 0) sV[(size_t)k * R + r] = make_c(0.0, 0.0);
      continue;  // (the next step's barrier #1 orders the reuse of sx / ssc)
    }
    const cplx tau = make_c(ssc[1], ssc[2]);
    cplx vr = make_c(0.0, 0.0), pr = make_c(0.0, 0.0);
    if (g == 0) {
      if (r > k) {
        const cplx xp0 = make_c(ssc[3], ssc[4]), scale = make_c(ssc[5], ssc[6]), ts = make_c(ssc[7], ssc[8]);
        cplx ys = spart[r];
#pragma unroll
        for (int gg = 1; gg < G; ++gg) ys = cadd(ys, spart[gg * R + r]);
        cfma(ys, sfirst[r], xp0);
        vr = (r == k + 1) ? make_c(1.0, 0.0) : cmul(scale, sx[r]);
        pr = cmul(ts, ys);
      }
      cplx dt = (r > k) ? ccmul(pr, vr) : make_c(0.0, 0.0);  // conj(p) v
      dt.x = warp_sum(dt.x);
      dt.y = warp_sum(dt.y);
      if (lane == 0) sdot[tid >> 5] = dt;
    }
    __syncthreads();  // #3
    if (g == 0 && r > k) {
      cplx dot = sdot[0];
#pragma unroll
      for (int q = 1; q < NWG; ++q) dot = cadd(dot, sdot[q]);
      const cplx a2 = cscale(-0.5, cmul(tau, dot));
      const cplx wr = cadd(pr, cmul(a2, vr));
      sv[r] = vr;
      sw[r] = wr;
      sV[(size_t)k * R + r] = vr;
    }
    __syncthreads();  // #4
    // A22 -= v w^H + w v^H
    if (r > k) {
      vr = sv[r];
      const cplx wr = sw[r];
      switch (jmin) {
#define X(J)                                                                         \
  case J:                                                                            \
    if (J < CPT) {                                                                   \
      const int c = g + G * J;                                                       \
      const cplx wc = sw[c], vc = sv[c];                                             \
      cplx &e = a[J < CPT ? J : 0];                                                  \
      e.x -= vr.x * wc.x + vr.y * wc.y + wr.x * vc.x + wr.y * vc.y;                  \
      e.y -= vr.y * wc.x - vr.x * wc.y + wr.y * vc.x - wr.x * vc.y;                  \
    }
        TRR_CASES(X)
#undef X
        default:
          break;
      }
    }
  }
  // last diagonal element
  if (r == d - 1) {
#pragma unroll
    for (int jj = 0; jj < CPT; ++jj) sx[g + G * jj] = cconj(a[jj]);
  }
  __syncthreads();
  if (tid == 0) {
    dout[cfg * d + d - 1] = sx[d - 1].x;
    eout[cfg * d + d - 1] = 0.0;
  }

  // ---- Q = H_0 ... H_{d-2}, backward accumulation; thread (c = r, g) holds Q[g + 4 jj][c] ----
  const int c = r;
  cplx *q = a;  // reuse the registers
#pragma unroll
  for (int jj = 0; jj < CPT; ++jj) q[jj] = make_c((g + G * jj == c) ? 1.0 : 0.0, 0.0);
  for (int k = d - 2; k >= 0; --k) {
    const cplx t = stau[k];
    const cplx *vk = sV + (size_t)k * R;
    cplx *part = spart + (k & 1) * G * R;
    const int imin = (k + 2 > g) ? (k + 2 - g + G - 1) / G : 0;  // first jj with row g + 4 jj >= k + 2
    if (c >= k + 2) {
      cplx u = make_c(0.0, 0.0);
      switch (imin) {
#define X(J)                                     \
  case J:                                        \
    if (J < CPT) ccfma(u, vk[g + G * J], q[J < CPT ? J : 0]);
        TRR_CASES(X)
#undef X
        default:
          break;
      }
      part[g * R + c] = u;
    }
    __syncthreads();
    if (c >= k + 2) {
      cplx u = part[c];
#pragma unroll
      for (int gg = 1; gg < G; ++gg) u = cadd(u, part[gg * R + c]);
      const cplx tu = cmul(t, u);
      // row k+1 of this column: -t u  (the thread that owns row k+1: g == (k+1) % 4)
      if (g == ((k + 1) & (G - 1))) {
        const int jr = (k + 1) / G;
#pragma unroll
        for (int jj = 0; jj < CPT; ++jj)
          if (jj == jr) q[jj] = make_c(-tu.x, -tu.y);
      }
      switch (imin) {
#define X(J)                                              \
  case J:                                                 \
    if (J < CPT) {                                        \
      const cplx vi = vk[g + G * J];                      \
      cplx &e = q[J < CPT ? J : 0];                       \
      e.x -= vi.x * tu.x - vi.y * tu.y;                   \
      e.y -= vi.x * tu.y + vi.y * tu.x;                   \
    }
        TRR_CASES(X)
#undef X
        default:
          break;
      }
    } else if (c == k + 1) {
      // column k+1: (1 - t) on the diagonal, -t v below
#pragma unroll
      for (int jj = 0; jj < CPT; ++jj) {
        const int i = g + G * jj;
        if (i == k + 1) {
          q[jj] = make_c(1.0 - t.x, -t.y);
        } else if (i >= k + 2 && i < d) {
          const cplx tv = cmul(t, vk[i]);
          q[jj] = make_c(-tv.x, -tv.y);
        }
      }
    }
  }
#pragma unroll
  for (int jj = 0; jj < CPT; ++jj) {
    const int i = g + G * jj;
    if (i < d && c < d) Qout[cfg * dd + (size_t)i * d + c] = q[jj];
  }
}

template <int R>
inline size_t hql_tridiag_reg_smem() {
  return ((size_t)R * R + 4 * R + 2 * TRR_G * R + 8 + 6 + R) * sizeof(cplx);
}

}  // namespace musim
