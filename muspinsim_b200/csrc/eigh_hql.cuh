// eigh_hql.cuh -- batched complex-Hermitian eigensolver: Householder tridiagonalisation +
// implicit QL.  Replaces np.linalg.eigh (LAPACK zheevd) in Hermitian.diag
// (/root/reference/muspinsim/spinop.py:51-82).  Prototype with identical loop structure:
// tools/hql_prototype.py.
//
// The batch is large (10^4 .. 10^7 matrices) and each matrix is small (d <= 120), so every
// stage uses the parallelisation that suits it, with intermediates in HBM/L2:
//
//   K1 hql_tridiag_kernel  one CTA per matrix, A in shared memory.  H = H0 + B.Z is built on
//                          chip; unblocked zhetd2 (lower): A = Q T Q^H; then Q = H_0..H_{d-2}
//                          is formed IN PLACE (zung2r backward accumulation).  Out: d, e, Q.
//   K2 hql_tql_kernel      one THREAD per matrix: the QL iteration on (d, e) is a serial
//                          chain of ~1.2 d^2 plane rotations (div + sqrt latency); with one
//                          matrix per thread thousands of chains run concurrently.  The
//                          rotations are RECORDED, not applied.  Out: eigenvalues (sorted),
//                          permutation, rotation stream.
//   K3 hql_apply_kernel    one CTA per matrix, one thread per row of the REAL eigenvector
//                          matrix Zt of T (starts as identity, lives in shared memory): replays
//                          the rotation stream with the carried-column trick (1 load + 1 store
//                          per rotation, no barriers).  Out: Zt with sorted columns.
//   K4 (rotate.cuh)        U = Q Zt  (complex x real batched GEMM).
#pragma once
#include "common.cuh"

namespace musim {

#define HQL_MAX_D 118  // A (d x (d|1) complex) must fit the 227 KB of opt-in shared memory

inline bool hql_supported(int d) { return d >= 1 && d <= HQL_MAX_D; }

struct HqlGeom {
  int R, G, nth;
};

inline HqlGeom hql_geom(int d) {
  HqlGeom g;
  if (d > 16) {
    g.R = (d + 31) & ~31;
    g.G = 4;
  } else {
    g.R = 1;
    while (g.R < d) g.R <<= 1;
    g.G = 32 / g.R;
    if (g.G > 8) g.G = 8;
    if (g.R * g.G < 32) g.G = 32 / g.R;
  }
  g.nth = g.R * g.G;
  if (g.nth < 32) g.nth = 32;
  return g;
}

// two-component block sum (re, im) -> all threads
__device__ __forceinline__ cplx block_sum2(cplx v, double *red /* >= 66 doubles */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v.x = warp_sum(v.x);
  v.y = warp_sum(v.y);
  __syncthreads();
  if (lane == 0) {
    red[wid] = v.x;
    red[32 + wid] = v.y;
  }
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  double tx = 0.0, ty = 0.0;
  for (int i = 0; i < nw; ++i) {
    tx += red[i];
    ty += red[32 + i];
  }
  return make_c(tx, ty);
}

// ---------------------------------------------------------------------------------------
// K1: tridiagonalise and form Q in place.
// ---------------------------------------------------------------------------------------
template <bool BUILD_H>
__global__ void __launch_bounds__(512)
hql_tridiag_kernel(int d, int R, int G, const cplx *__restrict__ H0, const cplx *__restrict__ Z,
                   const double *__restrict__ Bf, const cplx *__restrict__ Ain,
                   double *__restrict__ dout, double *__restrict__ eout, cplx *__restrict__ Qout) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ld = d | 1;
  cplx *sA = reinterpret_cast<cplx *>(smem_raw);  // (r,c) at [c*ld + r]
  cplx *sv = sA + (size_t)d * ld;                  // [d]
  cplx *sw = sv + d;                               // [d]
  cplx *stau = sw + d;                             // [d]
  cplx *spart = stau + d;                          // [G*R]
  double *red = reinterpret_cast<double *>(spart + (size_t)G * R);  // [66]
  const int tid = threadIdx.x, nth = blockDim.x;
  const int r = tid % R, g = tid / R;  // (row, column group); threads with g >= G idle in 2D loops
  const size_t cfg = blockIdx.x;
  const size_t dd = (size_t)d * d;

  double bx = 0, by = 0, bz = 0;
  if (BUILD_H) {
    bx = Bf[cfg * 3 + 0];
    by = Bf[cfg * 3 + 1];
    bz = Bf[cfg * 3 + 2];
  }
  for (int idx = tid; idx < d * d; idx += nth) {
    const int rr = idx / d, cc = idx - rr * d;
    cplx a;
    if (BUILD_H) {
      a = H0[idx];
      const cplx z0 = Z[idx], z1 = Z[dd + idx], z2 = Z[2 * dd + idx];
      a.x += bx * z0.x + by * z1.x + bz * z2.x;
      a.y += bx * z0.y + by * z1.y + bz * z2.y;
    } else {
      a = Ain[cfg * dd + idx];
    }
    sA[cc * ld + rr] = a;
  }
  __syncthreads();
  // enforce exact Hermiticity: average (r,c) with conj(c,r)
  for (int idx = tid; idx < d * d; idx += nth) {
    const int rr = idx / d, cc = idx - rr * d;
    if (rr < cc) {
      const cplx a = sA[cc * ld + rr], b = sA[rr * ld + cc];
      const cplx m = make_c(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
      sA[cc * ld + rr] = m;
      sA[rr * ld + cc] = cconj(m);
    } else if (rr == cc) {
      sA[cc * ld + rr].y = 0.0;
    }
  }
  __syncthreads();

  // ---- zhetd2 (lower) ----
  for (int k = 0; k < d - 1; ++k) {
    const int m = d - k - 1;
    const int o = k + 1;  // offset of the trailing block
    double xn = 0.0;
    for (int i = 1 + tid; i < m; i += nth) xn += cnorm2(sA[k * ld + o + i]);
    xn = block_sum(xn, red);
    const cplx alpha = sA[k * ld + o];
    if (tid == 0) dout[cfg * d + k] = sA[k * ld + k].x;
    if (xn == 0.0 && alpha.y == 0.0) {  // H = I
      if (tid == 0) {
        eout[cfg * d + k] = alpha.x;
        stau[k] = make_c(0.0, 0.0);
      }
      continue;
    }
    const double beta = -copysign(sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xn), alpha.x);
    const cplx tau = make_c((beta - alpha.x) / beta, -alpha.y / beta);
    // scale = 1 / (alpha - beta)
    const double ar = alpha.x - beta, ai = alpha.y;
    const double den = 1.0 / (ar * ar + ai * ai);
    const cplx scale = make_c(ar * den, -ai * den);
    if (tid == 0) {
      eout[cfg * d + k] = beta;
      stau[k] = tau;
    }
    for (int i = tid; i < m; i += nth) {
      cplx vi = make_c(1.0, 0.0);
      if (i > 0) {
        vi = cmul(sA[k * ld + o + i], scale);
        sA[k * ld + o + i] = vi;
      }
      sv[i] = vi;
    }
    __syncthreads();
    // p = tau * A22 v  (row r, columns c = g, g+G, ...)
    if (r < m && g < G) {
      cplx acc = make_c(0.0, 0.0);
      const cplx *row = sA + (size_t)o * ld + o + r;
      for (int c = g; c < m; c += G) cfma(acc, row[(size_t)c * ld], sv[c]);
      spart[g * R + r] = acc;
    }
    __syncthreads();
    cplx pr = make_c(0.0, 0.0), dot = make_c(0.0, 0.0);
    if (tid < m) {
      cplx s = spart[tid];
      for (int gg = 1; gg < G; ++gg) s = cadd(s, spart[gg * R + tid]);
      pr = cmul(tau, s);
      dot = ccmul(pr, sv[tid]);  // conj(p) * v
    }
    dot = block_sum2(dot, red);
    // alpha2 = -1/2 * tau * dot ;  w = p + alpha2 * v
    const cplx a2 = cscale(-0.5, cmul(tau, dot));
    if (tid < m) sw[tid] = cadd(pr, cmul(a2, sv[tid]));
    __syncthreads();
    // A22 -= v w^H + w v^H
    if (r < m && g < G) {
      const cplx vr = sv[r], wr = sw[r];
      cplx *row = sA + (size_t)o * ld + o + r;
      for (int c = g; c < m; c += G) {
        cplx a = row[(size_t)c * ld];
        const cplx wc = sw[c], vc = sv[c];
        // a -= vr*conj(wc) + wr*conj(vc)
        a.x -= vr.x * wc.x + vr.y * wc.y + wr.x * vc.x + wr.y * vc.y;
        a.y -= vr.y * wc.x - vr.x * wc.y + wr.y * vc.x - wr.x * vc.y;
        row[(size_t)c * ld] = a;
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    dout[cfg * d + d - 1] = sA[(d - 1) * ld + d - 1].x;
    eout[cfg * d + d - 1] = 0.0;
  }
  __syncthreads();

  // ---- zung2r, in place: Q = H_0 H_1 ... H_{d-2} ----
  for (int k = d - 2; k >= 0; --k) {
    const int m1 = d - k - 2;  // length of v[1:]
    const int o = k + 2;
    const cplx t = stau[k];
    for (int i = tid; i < m1; i += nth) sv[i] = sA[k * ld + o + i];
    __syncthreads();
    // u_c = sum_i conj(v_i) Q(o+i, o+c)   thread (c = r, g): rows i = g, g+G, ...
    if (r < m1 && g < G) {
      cplx acc = make_c(0.0, 0.0);
      const cplx *col = sA + (size_t)(o + r) * ld + o;
      for (int i = g; i < m1; i += G) ccfma(acc, sv[i], col[i]);
      spart[g * R + r] = acc;
    }
    __syncthreads();
    if (tid < m1) {
      cplx u = spart[tid];
      for (int gg = 1; gg < G; ++gg) u = cadd(u, spart[gg * R + tid]);
      sw[tid] = u;
      const cplx tu = cmul(t, u);
      sA[(size_t)(o + tid) * ld + k + 1] = make_c(-tu.x, -tu.y);  // row k+1
      const cplx tv = cmul(t, sv[tid]);
      sA[(size_t)(k + 1) * ld + o + tid] = make_c(-tv.x, -tv.y);  // column k+1
    }
    if (tid == 0) sA[(size_t)(k + 1) * ld + k + 1] = make_c(1.0 - t.x, -t.y);
    __syncthreads();
    if (r < m1 && g < G) {
      const cplx tv = cmul(t, sv[r]);
      cplx *row = sA + (size_t)o * ld + o + r;
      for (int c = g; c < m1; c += G) {
        cplx a = row[(size_t)c * ld];
        const cplx u = sw[c];
        a.x -= tv.x * u.x - tv.y * u.y;
        a.y -= tv.x * u.y + tv.y * u.x;
        row[(size_t)c * ld] = a;
      }
    }
    __syncthreads();
  }
  // row 0 / column 0 = e_0; write Q row-major
  for (int idx = tid; idx < d * d; idx += nth) {
    const int rr = idx / d, cc = idx - rr * d;
    cplx q;
    if (rr == 0 || cc == 0)
      q = make_c((rr == 0 && cc == 0) ? 1.0 : 0.0, 0.0);
    else
      q = sA[cc * ld + rr];
    Qout[cfg * dd + idx] = q;
  }
}

inline size_t hql_tridiag_smem(int d, const HqlGeom &g) {
  const int ld = d | 1;
  return ((size_t)d * ld + 3 * (size_t)d + (size_t)g.G * g.R) * sizeof(cplx) + 70 * sizeof(double);
}

// ---------------------------------------------------------------------------------------
// K2: implicit QL on (d, e), one thread per matrix; rotations are recorded.
//   rot[mat][j]  = (c, s) of the j-th rotation overall (s already negated as zlasr wants)
//   swp[mat][i]  = (l, m) of sweep i; nswp[mat] sweeps
//   lam[mat][:]  ascending eigenvalues; perm[mat][j] = column of Zt holding eigenvalue j
// ---------------------------------------------------------------------------------------
struct SweepIdx {
  unsigned short l, m;
};

__device__ __forceinline__ void dlaev2_dev(double a, double b, double c, double &rt1, double &rt2,
                                           double &cs1, double &sn1) {
  const double sm = a + c, df = a - c, adf = fabs(df), tb = b + b, ab = fabs(tb);
  double acmx, acmn;
  if (fabs(a) > fabs(c)) {
    acmx = a;
    acmn = c;
  } else {
    acmx = c;
    acmn = a;
  }
  double rt;
  if (adf > ab) {
    const double q = ab / adf;
    rt = adf * sqrt(1.0 + q * q);
  } else if (adf < ab) {
    const double q = adf / ab;
    rt = ab * sqrt(1.0 + q * q);
  } else {
    rt = ab * 1.4142135623730951;
  }
  int sgn1, sgn2;
  if (sm < 0.0) {
    rt1 = 0.5 * (sm - rt);
    sgn1 = -1;
    rt2 = (acmx / rt1) * acmn - (b / rt1) * b;
  } else if (sm > 0.0) {
    rt1 = 0.5 * (sm + rt);
    sgn1 = 1;
    rt2 = (acmx / rt1) * acmn - (b / rt1) * b;
  } else {
    rt1 = 0.5 * rt;
    rt2 = -0.5 * rt;
    sgn1 = 1;
  }
  double cs;
  if (df >= 0.0) {
    cs = df + rt;
    sgn2 = 1;
  } else {
    cs = df - rt;
    sgn2 = -1;
  }
  if (fabs(cs) > ab) {
    const double ct = -tb / cs;
    sn1 = 1.0 / sqrt(1.0 + ct * ct);
    cs1 = ct * sn1;
  } else if (ab == 0.0) {
    cs1 = 1.0;
    sn1 = 0.0;
  } else {
    const double tn = -cs / tb;
    cs1 = 1.0 / sqrt(1.0 + tn * tn);
    sn1 = tn * cs1;
  }
  if (sgn1 == sgn2) {
    const double tn = cs1;
    cs1 = -sn1;
    sn1 = tn;
  }
}

template <int MAXD>
__global__ void __launch_bounds__(128)
hql_tql_kernel(int d, int64_t n, const double *__restrict__ din, const double *__restrict__ ein,
               double *__restrict__ lam, unsigned short *__restrict__ perm, double2 *__restrict__ rot,
               size_t rot_cap, SweepIdx *__restrict__ swp, int swp_cap, int *__restrict__ nswp,
               int *__restrict__ status) {
  const int64_t mat = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (mat >= n) return;
  double dl[MAXD], el[MAXD];
  for (int i = 0; i < d; ++i) {
    dl[i] = din[mat * d + i];
    el[i] = ein[mat * d + i];
  }
  el[d - 1] = 0.0;
  double2 *myrot = rot + mat * rot_cap;
  SweepIdx *myswp = swp + mat * swp_cap;
  size_t nrot = 0;
  int ns = 0;
  bool fail = false;
  const double eps2 = 4.930380657631324e-32;  // (2^-52)^2
  const double safmin = 2.2250738585072014e-308;
  int l = 0, nit = 0;
  const int maxit = 60 * d;
  while (l < d) {
    int m = l;
    while (m < d - 1) {
      const double tst = el[m] * el[m];
      if (tst <= (eps2 * fabs(dl[m])) * fabs(dl[m + 1]) + safmin) break;
      ++m;
    }
    if (m < d - 1) el[m] = 0.0;
    if (m == l) {
      ++l;
      continue;
    }
    if (ns >= swp_cap || nrot + (size_t)(m - l) > rot_cap || nit >= maxit) {
      fail = true;
      break;
    }
    if (m == l + 1) {
      double rt1, rt2, c, s;
      dlaev2_dev(dl[l], el[l], dl[l + 1], rt1, rt2, c, s);
      myrot[nrot++] = make_double2(c, s);
      myswp[ns++] = SweepIdx{(unsigned short)l, (unsigned short)(l + 1)};
      dl[l] = rt1;
      dl[l + 1] = rt2;
      el[l] = 0.0;
      l += 2;
      continue;
    }
    ++nit;
    double p = dl[l];
    double g = (dl[l + 1] - p) / (2.0 * el[l]);
    double r = sqrt(fma(g, g, 1.0));
    g = dl[m] - p + el[l] / (g + copysign(r, g));
    double s = 1.0, c = 1.0;
    p = 0.0;
    for (int i = m - 1; i >= l; --i) {
      const double f = s * el[i];
      const double b = c * el[i];
      r = sqrt(fma(g, g, f * f));
      if (r == 0.0) {
        c = 1.0;
        s = 0.0;
      } else {
        const double ri = 1.0 / r;
        c = g * ri;
        s = f * ri;
      }
      if (i != m - 1) el[i + 1] = r;
      g = dl[i + 1] - p;
      r = (dl[i] - g) * s + 2.0 * c * b;
      p = s * r;
      dl[i + 1] = g + p;
      g = c * r - b;
      myrot[nrot++] = make_double2(c, -s);
    }
    dl[l] -= p;
    el[l] = g;
    myswp[ns++] = SweepIdx{(unsigned short)l, (unsigned short)m};
  }
  nswp[mat] = ns;
  if (fail) atomicMax(status, 1);
  // ascending order: insertion sort of indices
  unsigned short *pm = perm + mat * d;
  for (int i = 0; i < d; ++i) {
    const double v = dl[i];
    int j = i - 1;
    while (j >= 0 && dl[pm[j]] > v) {
      pm[j + 1] = pm[j];
      --j;
    }
    pm[j + 1] = (unsigned short)i;
  }
  for (int i = 0; i < d; ++i) lam[mat * d + i] = dl[pm[i]];
}

// ---------------------------------------------------------------------------------------
// K3: replay the rotations on Zt (real, starts as identity), one thread per row.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
hql_apply_kernel(int d, const double2 *__restrict__ rot, size_t rot_cap,
                 const SweepIdx *__restrict__ swp, int swp_cap, const int *__restrict__ nswp,
                 const unsigned short *__restrict__ perm, double *__restrict__ Zt) {
  extern __shared__ double sZ[];  // (r,c) at [c*ld + r]
  const int ld = d | 1;
  const int tid = threadIdx.x, nth = blockDim.x;
  const size_t mat = blockIdx.x;
  for (int idx = tid; idx < d * ld; idx += nth) sZ[idx] = 0.0;
  __syncthreads();
  for (int i = tid; i < d; i += nth) sZ[i * ld + i] = 1.0;
  __syncthreads();
  const double2 *myrot = rot + mat * rot_cap;
  const SweepIdx *myswp = swp + mat * swp_cap;
  const int ns = nswp[mat];
  for (int r = tid; r < d; r += nth) {
    size_t off = 0;
    for (int sidx = 0; sidx < ns; ++sidx) {
      const SweepIdx lm = myswp[sidx];
      const int l = lm.l, m = lm.m;
      double x = sZ[m * ld + r];
      const double2 *rs = myrot + off;
#pragma unroll 4
      for (int j = m - 1; j >= l; --j) {
        const double2 cs = rs[m - 1 - j];
        const double a = sZ[j * ld + r];
        sZ[(j + 1) * ld + r] = cs.x * x - cs.y * a;
        x = cs.y * x + cs.x * a;
      }
      sZ[l * ld + r] = x;
      off += (size_t)(m - l);
    }
  }
  __syncthreads();
  const unsigned short *pm = perm + mat * d;
  const size_t dd = (size_t)d * d;
  for (int idx = tid; idx < d * d; idx += nth) {
    const int rr = idx / d, j = idx - rr * d;
    Zt[mat * dd + idx] = sZ[pm[j] * ld + rr];
  }
}

}  // namespace musim
