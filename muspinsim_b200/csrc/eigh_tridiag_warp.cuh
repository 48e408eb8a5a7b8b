// eigh_tridiag_warp.cuh -- K1 for small matrices (d <= 32): one WARP per matrix, lane = row,
// no block barriers at all (the 2-barriers-per-step CTA kernel of eigh_tridiag_rw.cuh spends
// most of a d = 24 step waiting: 648 ms for the 10^7 matrices of the ALC scan).  A lives in
// shared memory (column-major, odd leading dimension), v / w are exchanged through a per-warp
// shared-memory line; all reductions are warp shuffles.  Same outputs as the other K1 variants
// (d, e, tau, packed reflectors); arithmetic = tools/hql_prototype.py::tridiag_lower.
#pragma once
#include "common.cuh"

namespace musim {

#define TRW_WARPS 4

__global__ void __launch_bounds__(32 * TRW_WARPS)
hql_tridiag_warp_kernel(int d, int64_t n, const cplx *__restrict__ H0, const cplx *__restrict__ Z,
                        const double *__restrict__ Bf, const cplx *__restrict__ Ain,
                        double *__restrict__ dout, double *__restrict__ eout, cplx *__restrict__ Vp,
                        size_t vcap, cplx *__restrict__ tauout) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ld = d | 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  cplx *sA = reinterpret_cast<cplx *>(smem_raw) + (size_t)warp * ((size_t)d * ld + 64);  // (r,c) at [c*ld + r]
  cplx *sv = sA + (size_t)d * ld;  // [32]
  cplx *sw = sv + 32;              // [32]
  const int64_t cfg = (int64_t)blockIdx.x * TRW_WARPS + warp;
  if (cfg >= n) return;  // whole warp
  const size_t dd = (size_t)d * d;
  {
    double bx = 0, by = 0, bz = 0;
    if (!Ain) {
      bx = Bf[cfg * 3 + 0];
      by = Bf[cfg * 3 + 1];
      bz = Bf[cfg * 3 + 2];
    }
    for (int idx = lane; idx < d * d; idx += 32) {
      const int r = idx / d, c = idx - r * d;
      cplx v;
      if (Ain) {
        v = Ain[cfg * dd + idx];
      } else {
        v = H0[idx];
        const cplx z0 = Z[idx], z1 = Z[dd + idx], z2 = Z[2 * dd + idx];
        v.x += bx * z0.x + by * z1.x + bz * z2.x;
        v.y += bx * z0.y + by * z1.y + bz * z2.y;
      }
      if (r == c) v.y = 0.0;
      sA[c * ld + r] = v;
    }
  }
  __syncwarp();
  const int r = lane;  // my row
  for (int k = 0; k < d - 1; ++k) {
    const int mk = d - k - 2;
    const size_t voff = (size_t)mk * (mk - 1) / 2;
    // column k (rows > k); lanes <= k and >= d hold zero
    const cplx x = (r > k && r < d) ? sA[k * ld + r] : make_c(0.0, 0.0);
    double xn = (r >= k + 2) ? cnorm2(x) : 0.0;
    xn = warp_sum(xn);
    const cplx alpha = make_c(__shfl_sync(0xffffffffu, x.x, k + 1), __shfl_sync(0xffffffffu, x.y, k + 1));
    if (lane == 0) dout[cfg * d + k] = sA[k * ld + k].x;
    if (xn == 0.0 && alpha.y == 0.0) {  // H_k = I
      if (lane == 0) {
        eout[cfg * d + k] = alpha.x;
        tauout[cfg * d + k] = make_c(0.0, 0.0);
      }
      for (int i = lane; i < mk; i += 32) Vp[cfg * vcap + voff + i] = make_c(0.0, 0.0);
      continue;
    }
    const double s2 = alpha.x * alpha.x + alpha.y * alpha.y + xn;
    const double ri = rsqrt(s2);
    const double sg = (alpha.x >= 0.0) ? -1.0 : 1.0;
    const double beta = sg * (s2 * ri);
    const double ib = sg * ri;
    const cplx tau = make_c((beta - alpha.x) * ib, -alpha.y * ib);
    const double ar = alpha.x - beta, ai = alpha.y;
    const double den = __drcp_rn(ar * ar + ai * ai);
    const cplx scale = make_c(ar * den, -ai * den);
    cplx v = make_c(0.0, 0.0);
    if (r == k + 1)
      v = make_c(1.0, 0.0);
    else if (r > k + 1)
      v = cmul(scale, x);
    if (lane == 0) {
      eout[cfg * d + k] = beta;
      tauout[cfg * d + k] = tau;
    }
    sv[r] = v;
    if (r >= k + 2 && r < d) Vp[cfg * vcap + voff + (r - k - 2)] = v;
    __syncwarp();
    // p = tau A22 v
    cplx y = make_c(0.0, 0.0);
    if (r > k && r < d) {
      cplx y1 = make_c(0.0, 0.0);
      int c = k + 1;
      for (; c + 1 < d; c += 2) {
        cfma(y, sA[c * ld + r], sv[c]);
        cfma(y1, sA[(c + 1) * ld + r], sv[c + 1]);
      }
      if (c < d) cfma(y, sA[c * ld + r], sv[c]);
      y = cadd(y, y1);
    }
    const cplx p = cmul(tau, y);
    cplx dot = ccmul(p, v);  // conj(p) v  (v = 0 outside the trailing block)
    dot.x = warp_sum(dot.x);
    dot.y = warp_sum(dot.y);
    const cplx a2 = cscale(-0.5, cmul(tau, dot));
    const cplx w = cadd(p, cmul(a2, v));
    sw[r] = w;
    __syncwarp();
    // A22 -= v w^H + w v^H
    if (r > k && r < d) {
      for (int c = k + 1; c < d; ++c) {
        cplx a = sA[c * ld + r];
        const cplx wc = sw[c], vc = sv[c];
        a.x -= v.x * wc.x + v.y * wc.y + w.x * vc.x + w.y * vc.y;
        a.y -= v.y * wc.x - v.x * wc.y + w.y * vc.x - w.x * vc.y;
        sA[c * ld + r] = a;
      }
    }
    __syncwarp();
  }
  if (lane == 0) {
    dout[cfg * d + d - 1] = sA[(d - 1) * ld + d - 1].x;
    eout[cfg * d + d - 1] = 0.0;
  }
}


// ---------------------------------------------------------------------------------------
// Fused variant: the rank-2 update of step k and the matrix-vector product of step k+1 share one
// pass over the trailing block (ncu on the kernel above at d = 24: LSU 78 % busy, every element of
// A is loaded twice and stored once per step).  Column k+1 is updated first, the next Householder
// vector v' is formed from it, and the remaining columns are updated and multiplied by v'_c in the
// same loop.  Outputs identical to hql_tridiag_warp_kernel.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * TRW_WARPS)
hql_tridiag_warpf_kernel(int d, int64_t n, const cplx *__restrict__ H0, const cplx *__restrict__ Z,
                         const double *__restrict__ Bf, const cplx *__restrict__ Ain,
                         double *__restrict__ dout, double *__restrict__ eout, cplx *__restrict__ Vp,
                         size_t vcap, cplx *__restrict__ tauout, int dstride, int koff) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ld = d | 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  cplx *sA = reinterpret_cast<cplx *>(smem_raw) + (size_t)warp * ((size_t)d * ld + 96);  // (r,c) at [c*ld + r]
  cplx *sva = sA + (size_t)d * ld;  // [32] v, double-buffered
  cplx *svb = sva + 32;
  cplx *sw = svb + 32;              // [32]
  const int64_t cfg = (int64_t)blockIdx.x * TRW_WARPS + warp;
  if (cfg >= n) return;  // whole warp
  const size_t dd = (size_t)d * d;
  {
    double bx = 0, by = 0, bz = 0;
    if (!Ain) {
      bx = Bf[cfg * 3 + 0];
      by = Bf[cfg * 3 + 1];
      bz = Bf[cfg * 3 + 2];
    }
    for (int idx = lane; idx < d * d; idx += 32) {
      const int r = idx / d, c = idx - r * d;
      cplx v;
      if (Ain) {
        v = Ain[cfg * dd + idx];
      } else {
        v = H0[idx];
        const cplx z0 = Z[idx], z1 = Z[dd + idx], z2 = Z[2 * dd + idx];
        v.x += bx * z0.x + by * z1.x + bz * z2.x;
        v.y += bx * z0.y + by * z1.y + bz * z2.y;
      }
      if (r == c) v.y = 0.0;
      sA[c * ld + r] = v;
    }
  }
  __syncwarp();
  const int r = lane;  // my row

  // Householder vector of step k from x = A(r, k) (rows r > k): writes e, tau, the packed
  // reflector and d_k; returns v (1 at row k+1, 0 for r <= k and r >= d) and tau
  auto house = [&](int k, cplx x, double akk, cplx &v, cplx &tau) {
    const int mk = d - k - 2;
    const size_t voff = (size_t)mk * (mk - 1) / 2;
    double xn = (r >= k + 2 && r < d) ? cnorm2(x) : 0.0;
    xn = warp_sum(xn);
    const cplx alpha = make_c(__shfl_sync(0xffffffffu, x.x, k + 1), __shfl_sync(0xffffffffu, x.y, k + 1));
    const double dk = __shfl_sync(0xffffffffu, akk, k);
    if (lane == 0) dout[cfg * dstride + koff + k] = dk;
    if (xn == 0.0 && alpha.y == 0.0) {  // H_k = I
      if (lane == 0) {
        eout[cfg * dstride + koff + k] = alpha.x;
        tauout[cfg * dstride + koff + k] = make_c(0.0, 0.0);
      }
      for (int i = lane; i < mk; i += 32) Vp[cfg * vcap + voff + i] = make_c(0.0, 0.0);
      v = make_c(0.0, 0.0);
      tau = make_c(0.0, 0.0);
      return;
    }
    const double s2 = alpha.x * alpha.x + alpha.y * alpha.y + xn;
    const double ri = rsqrt(s2);
    const double sg = (alpha.x >= 0.0) ? -1.0 : 1.0;
    const double beta = sg * (s2 * ri);
    const double ib = sg * ri;
    tau = make_c((beta - alpha.x) * ib, -alpha.y * ib);
    const double ar = alpha.x - beta, ai = alpha.y;
    const double den = __drcp_rn(ar * ar + ai * ai);
    const cplx scale = make_c(ar * den, -ai * den);
    v = make_c(0.0, 0.0);
    if (r == k + 1)
      v = make_c(1.0, 0.0);
    else if (r > k + 1 && r < d)
      v = cmul(scale, x);
    if (lane == 0) {
      eout[cfg * dstride + koff + k] = beta;
      tauout[cfg * dstride + koff + k] = tau;
    }
    if (r >= k + 2 && r < d) Vp[cfg * vcap + voff + (r - k - 2)] = v;
  };

  cplx *sv = sva, *sv2 = svb;
  cplx v, tau;
  {  // step 0: vector from column 0, plain matrix-vector product
    const cplx x = (r > 0 && r < d) ? sA[r] : make_c(0.0, 0.0);
    const double a00 = (r == 0) ? sA[0].x : 0.0;
    house(0, x, a00, v, tau);
    sv[r] = v;
    __syncwarp();
  }
  cplx y = make_c(0.0, 0.0);
  if (r > 0 && r < d) {
    cplx y1 = make_c(0.0, 0.0);
    int c = 1;
    for (; c + 1 < d; c += 2) {
      cfma(y, sA[c * ld + r], sv[c]);
      cfma(y1, sA[(c + 1) * ld + r], sv[c + 1]);
    }
    if (c < d) cfma(y, sA[c * ld + r], sv[c]);
    y = cadd(y, y1);
  }
  for (int k = 0; k < d - 1; ++k) {
    // w of step k
    const cplx p = cmul(tau, y);
    cplx dot = ccmul(p, v);  // conj(p) v  (v = 0 outside the trailing block)
    dot.x = warp_sum(dot.x);
    dot.y = warp_sum(dot.y);
    const cplx a2 = cscale(-0.5, cmul(tau, dot));
    const cplx w = cadd(p, cmul(a2, v));
    sw[r] = w;
    __syncwarp();
    const bool live = r > k && r < d;
    // column k+1 first
    const int c1 = k + 1;
    cplx a1 = make_c(0.0, 0.0);
    if (live) {
      a1 = sA[c1 * ld + r];
      const cplx wc = sw[c1], vc = sv[c1];
      a1.x -= v.x * wc.x + v.y * wc.y + w.x * vc.x + w.y * vc.y;
      a1.y -= v.y * wc.x - v.x * wc.y + w.y * vc.x - w.x * vc.y;
    }
    if (k == d - 2) {  // last step: only the final diagonal entry remains
      const double dl = __shfl_sync(0xffffffffu, a1.x, d - 1);
      if (lane == 0) {
        dout[cfg * dstride + koff + d - 1] = dl;
        eout[cfg * dstride + koff + d - 1] = 0.0;
      }
      break;
    }
    // Householder vector of step k+1 from the updated column k+1
    cplx v2, tau2;
    house(k + 1, (r > k + 1 && r < d) ? a1 : make_c(0.0, 0.0), (r == k + 1) ? a1.x : 0.0, v2, tau2);
    sv2[r] = v2;
    __syncwarp();
    // remaining columns: update with (v, w) and multiply by v2 in the same pass
    cplx y2 = make_c(0.0, 0.0), y3 = make_c(0.0, 0.0);
    if (live) {
      int c = k + 2;
      for (; c + 1 < d; c += 2) {
        cplx a = sA[c * ld + r], b = sA[(c + 1) * ld + r];
        const cplx wc = sw[c], vc = sv[c], wd = sw[c + 1], vd = sv[c + 1];
        a.x -= v.x * wc.x + v.y * wc.y + w.x * vc.x + w.y * vc.y;
        a.y -= v.y * wc.x - v.x * wc.y + w.y * vc.x - w.x * vc.y;
        b.x -= v.x * wd.x + v.y * wd.y + w.x * vd.x + w.y * vd.y;
        b.y -= v.y * wd.x - v.x * wd.y + w.y * vd.x - w.x * vd.y;
        sA[c * ld + r] = a;
        sA[(c + 1) * ld + r] = b;
        cfma(y2, a, sv2[c]);
        cfma(y3, b, sv2[c + 1]);
      }
      if (c < d) {
        cplx a = sA[c * ld + r];
        const cplx wc = sw[c], vc = sv[c];
        a.x -= v.x * wc.x + v.y * wc.y + w.x * vc.x + w.y * vc.y;
        a.y -= v.y * wc.x - v.x * wc.y + w.y * vc.x - w.x * vc.y;
        sA[c * ld + r] = a;
        cfma(y2, a, sv2[c]);
      }
    }
    y = (r > k + 1 && r < d) ? cadd(y2, y3) : make_c(0.0, 0.0);
    v = v2;
    tau = tau2;
    cplx *t = sv;
    sv = sv2;
    sv2 = t;
    __syncwarp();
  }
  if (d == 1 && lane == 0) {
    dout[cfg * dstride + koff] = sA[0].x;
    eout[cfg * dstride + koff] = 0.0;
  }
}

inline size_t hql_tridiag_warpf_smem(int d) {
  return (size_t)TRW_WARPS * ((size_t)d * (d | 1) + 96) * sizeof(cplx);
}

inline size_t hql_tridiag_warp_smem(int d) {
  return (size_t)TRW_WARPS * ((size_t)d * (d | 1) + 64) * sizeof(cplx);
}

}  // namespace musim
