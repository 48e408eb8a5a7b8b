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

#define HQL_MAX_D 112  // A (d x (d|1) complex) + vectors must fit the 227 KB of opt-in shared memory

#define HQL_LARGE_MAX_D_ 1024  // global-memory kernels of eigh_large.cuh take over above HQL_MAX_D
inline bool hql_supported(int d) { return d >= 1 && d <= HQL_LARGE_MAX_D_; }

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
                   double *__restrict__ dout, double *__restrict__ eout, cplx *__restrict__ Qout,
                   cplx *__restrict__ Vp, size_t vcap, cplx *__restrict__ tauout) {
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

  if (Vp) {
    // reflector output for hql_reflect_kernel: v_k[2:] packed in the order the backward
    // application consumes them (k = d-2 first), plus tau; Q is not formed here.
    for (int idx = tid; idx < d * d; idx += nth) {
      const int k = idx / d, i = idx - k * d;  // column k, row i
      if (k < d - 2 && i >= k + 2) {
        const int mk = d - k - 2;  // entries of v_k
        const size_t off = (size_t)mk * (mk - 1) / 2;  // sum of the entry counts of all k' > k
        Vp[cfg * vcap + off + (i - k - 2)] = sA[k * ld + i];
      }
    }
    for (int k = tid; k < d; k += nth) tauout[cfg * d + k] = (k < d - 1) ? stau[k] : make_c(0.0, 0.0);
    return;
  }
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

// (d, e) live in shared memory, laid out [i][thread] so that the threads of a warp (different
// matrices, nearly the same i) hit different banks; the next (d, e) pair is prefetched one step
// ahead so that shared-memory latency is off the serial chain.
// NT = matrices per block (one warp, NT active lanes).  The kernel is bound by the latency of
// one thread's serial chain (ncu: one warp per SM sub-partition issues every ~4.3 cycles, FP64
// pipe 4 % busy), not by lanes: NT = 8 puts four times as many warps on each sub-partition.
__device__ __forceinline__ double tql_lds(const double *p) {
  double r;
  asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(r) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
  return r;
}

template <int NT>
__global__ void __launch_bounds__(NT)
hql_tql_kernel(int d, int64_t n, const double *__restrict__ din, const double *__restrict__ ein,
               double *__restrict__ lam, unsigned short *__restrict__ perm, double2 *__restrict__ rot,
               size_t rot_cap, SweepIdx *__restrict__ swp, int swp_cap, int *__restrict__ nswp,
               int *__restrict__ status, int sorted, int dpad) {
  // dpad > 0 (register replay kernel with dpad columns): every sweep (l, m) is recorded for the
  // whole 8-column blocks it touches, columns jhi(b(m-1)) .. 8 b(l), with identity rotations
  // outside [l, m) -- the replay kernel then has no partially active blocks
  extern __shared__ double tql_smem[];
  double *dl = tql_smem + threadIdx.x;  // dl[i*NT]
  double *el = tql_smem + (size_t)d * NT + threadIdx.x;
#define DL(i) dl[(i) * NT]
#define EL(i) el[(i) * NT]
  const int64_t mat = (int64_t)blockIdx.x * NT + threadIdx.x;
  if (mat >= n) return;
  for (int i = 0; i < d; ++i) {
    DL(i) = din[mat * d + i];
    EL(i) = ein[mat * d + i];
  }
  EL(d - 1) = 0.0;
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
    // first m >= l with a negligible e_m (m = d - 1 if none), four candidates per pass with their nine loads
    // pinned in front of the tests (asm volatile: the compiler otherwise sinks each load to its test behind the
    // previous branch): the shared-memory latency is paid once per four elements (source-level ncu: this
    // scan was a third of the kernel's stall samples, all short_scoreboard).  Same tests in the same order as
    // the one-at-a-time loop, so the same m.
    int m = l;
    while (m < d - 1) {
      const int i1 = min(m + 1, d - 1), i2 = min(m + 2, d - 1), i3 = min(m + 3, d - 1), i4 = min(m + 4, d - 1);
      const double e0 = tql_lds(&EL(m)), e1 = tql_lds(&EL(i1)), e2 = tql_lds(&EL(i2)), e3 = tql_lds(&EL(i3));
      const double a0 = fabs(tql_lds(&DL(m))), a1 = fabs(tql_lds(&DL(i1))), a2 = fabs(tql_lds(&DL(i2))),
                   a3 = fabs(tql_lds(&DL(i3))), a4 = fabs(tql_lds(&DL(i4)));
      const bool s0 = e0 * e0 <= (eps2 * a0) * a1 + safmin, s1 = e1 * e1 <= (eps2 * a1) * a2 + safmin,
                 s2 = e2 * e2 <= (eps2 * a2) * a3 + safmin, s3 = e3 * e3 <= (eps2 * a3) * a4 + safmin;
      // offset of the first hit among the candidates that exist (index < d - 1), else 4
      const int lim = d - 1 - m;  // > 0
      const int hit = s0 ? 0 : ((s1 || lim <= 1) ? 1 : ((s2 || lim <= 2) ? 2 : ((s3 || lim <= 3) ? 3 : 4)));
      m += hit;
      if (hit < 4) break;
    }
    if (m < d - 1) EL(m) = 0.0;
    if (m == l) {
      ++l;
      continue;
    }
    int pad_top = 0, pad_bot = 0;
    if (dpad > 0) {
      const int jtop = min(8 * ((m - 1) >> 3) + 7, dpad - 2);
      pad_top = jtop - (m - 1);
      pad_bot = l - 8 * (l >> 3);
    }
    if (ns >= swp_cap || nrot + (size_t)(m - l + pad_top + pad_bot) > rot_cap || nit >= maxit) {
      fail = true;
      break;
    }
    for (int q = 0; q < pad_top; ++q) myrot[nrot++] = make_double2(1.0, 0.0);
    if (m == l + 1) {
      double rt1, rt2, c, s;
      dlaev2_dev(DL(l), EL(l), DL(l + 1), rt1, rt2, c, s);
      myrot[nrot++] = make_double2(c, s);
      for (int q = 0; q < pad_bot; ++q) myrot[nrot++] = make_double2(1.0, 0.0);
      myswp[ns++] = SweepIdx{(unsigned short)l, (unsigned short)(l + 1)};
      DL(l) = rt1;
      DL(l + 1) = rt2;
      EL(l) = 0.0;
      l += 2;
      continue;
    }
    ++nit;
    double p = DL(l);
    const double el_l = EL(l);
    double g = (DL(l + 1) - p) / (2.0 * el_l);
    double r = sqrt(fma(g, g, 1.0));
    g = DL(m) - p + el_l / (g + copysign(r, g));
    double s = 1.0, c = 1.0;
    p = 0.0;
    // carried / prefetched values: d_ip1 = d[i+1] (untouched so far in this sweep), e_i, d_i
    double d_ip1 = DL(m);
    double e_i = EL(m - 1), d_i = DL(m - 1);
    double2 *rp = myrot + nrot;
    for (int i = m - 1; i >= l; --i) {
      double e_nx = 0.0, d_nx = 0.0;
      if (i > l) {
        e_nx = EL(i - 1);
        d_nx = DL(i - 1);
      }
      const double f = s * e_i;
      const double b = c * e_i;
      const double q = fma(g, g, f * f);
      if (q == 0.0) {
        c = 1.0;
        s = 0.0;
        r = 0.0;
      } else {
        const double ri = rsqrt(q);
        r = q * ri;
        c = g * ri;
        s = f * ri;
      }
      if (i != m - 1) EL(i + 1) = r;
      g = d_ip1 - p;
      r = fma(d_i - g, s, 2.0 * c * b);
      p = s * r;
      DL(i + 1) = g + p;
      g = fma(c, r, -b);
      *rp++ = make_double2(c, -s);
      d_ip1 = d_i;
      e_i = e_nx;
      d_i = d_nx;
    }
    nrot += (size_t)(m - l);
    for (int q = 0; q < pad_bot; ++q) myrot[nrot++] = make_double2(1.0, 0.0);
    DL(l) = DL(l) - p;
    EL(l) = g;
    myswp[ns++] = SweepIdx{(unsigned short)l, (unsigned short)m};
  }
  nswp[mat] = ns;
  if (fail) atomicMax(status, 1);
  unsigned short *pm = perm + mat * d;
  if (!sorted) {  // pipeline use: eigenpairs stay in the order the iteration leaves them
    for (int i = 0; i < d; ++i) {
      pm[i] = (unsigned short)i;
      lam[mat * d + i] = DL(i);
    }
    return;
  }
  // ascending order: insertion sort of indices
  for (int i = 0; i < d; ++i) {
    const double v = DL(i);
    int j = i - 1;
    while (j >= 0 && DL(pm[j]) > v) {
      pm[j + 1] = pm[j];
      --j;
    }
    pm[j + 1] = (unsigned short)i;
  }
  for (int i = 0; i < d; ++i) lam[mat * d + i] = DL(pm[i]);
#undef DL
#undef EL
}

inline size_t hql_tql_smem(int d, int nt) { return 2 * (size_t)d * nt * sizeof(double); }

// ---------------------------------------------------------------------------------------
// K3: replay the rotations on Zt (real, starts as identity), one thread per row.  The
// rotation stream is staged through a shared-memory ring with cp.async (deep prefetch: the
// stream comes from HBM and is read exactly once); each thread processes rotations in groups
// of 8 with all loads issued before the dependent chain  x <- s x + c a.
// ---------------------------------------------------------------------------------------
#define HQL_TILE 256  // rotations per staging tile (4 KB)

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__global__ void __launch_bounds__(128)
hql_apply_kernel(int d, const double2 *__restrict__ rot, size_t rot_cap,
                 const SweepIdx *__restrict__ swp, int swp_cap, const int *__restrict__ nswp,
                 const unsigned short *__restrict__ perm, double *__restrict__ Zt) {
  extern __shared__ __align__(16) unsigned char apply_smem[];
  double2 *ring = reinterpret_cast<double2 *>(apply_smem);           // [2][HQL_TILE]
  double *sZ = reinterpret_cast<double *>(ring + 2 * HQL_TILE);      // (r,c) at [c*ld + r]
  const int ld = d | 1;
  SweepIdx *sswp = reinterpret_cast<SweepIdx *>(sZ + (size_t)d * ld);  // [swp_cap]
  const int tid = threadIdx.x, nth = blockDim.x;
  const size_t mat = blockIdx.x;
  const double2 *myrot = rot + mat * rot_cap;
  const int ns = nswp[mat];
  size_t total = 0;
  for (int i = tid; i < ns; i += nth) sswp[i] = swp[mat * swp_cap + i];
  for (int idx = tid; idx < d * ld; idx += nth) sZ[idx] = 0.0;
  __syncthreads();
  for (int i = 0; i < ns; ++i) total += (size_t)(sswp[i].m - sswp[i].l);
  const int ntiles = (int)((total + HQL_TILE - 1) / HQL_TILE);
  auto stage = [&](int t) {  // issue tile t into ring slot t & 1
    if (t < ntiles) {
      const size_t base = (size_t)t * HQL_TILE;
      for (int e = tid; e < HQL_TILE; e += nth)
        if (base + e < total) cp_async16(&ring[(t & 1) * HQL_TILE + e], &myrot[base + e]);
    }
    cp_async_commit();
  };
  stage(0);
  stage(1);
  for (int i = tid; i < d; i += nth) sZ[i * ld + i] = 1.0;
  cp_async_wait<1>();
  __syncthreads();

  const int r = tid;  // one row per thread (d <= 128 = blockDim bound)
  const bool active = r < d;
  size_t g = 0;  // global rotation index
  int tile = 0;
  for (int sidx = 0; sidx < ns; ++sidx) {
    const int l = sswp[sidx].l, m = sswp[sidx].m;
    double x = active ? sZ[m * ld + r] : 0.0;
    int j = m - 1;
    while (j >= l) {
      // rotations [g, g+cnt) stay inside the current tile
      const int in_tile = (int)(g - (size_t)tile * HQL_TILE);
      int cnt = min(j - l + 1, HQL_TILE - in_tile);
      const double2 *rs = ring + (tile & 1) * HQL_TILE + in_tile;
      int k = 0;
      if (active) {
        for (; k + 8 <= cnt; k += 8) {
          double a[8];
          double2 cs[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            a[u] = sZ[(j - k - u) * ld + r];
            cs[u] = rs[k + u];
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            sZ[(j - k - u + 1) * ld + r] = cs[u].x * x - cs[u].y * a[u];
            x = fma(cs[u].y, x, cs[u].x * a[u]);
          }
        }
        for (; k < cnt; ++k) {
          const double a = sZ[(j - k) * ld + r];
          const double2 cs = rs[k];
          sZ[(j - k + 1) * ld + r] = cs.x * x - cs.y * a;
          x = fma(cs.y, x, cs.x * a);
        }
      }
      g += cnt;
      j -= cnt;
      if (g == (size_t)(tile + 1) * HQL_TILE && g < total) {
        // tile consumed by everyone -> refill its slot with tile+2, wait for tile+1
        __syncthreads();
        stage(tile + 2);
        cp_async_wait<1>();
        __syncthreads();
        ++tile;
      }
    }
    if (active) sZ[l * ld + r] = x;
  }
  cp_async_wait<0>();
  __syncthreads();
  const unsigned short *pm = perm + mat * d;
  const size_t dd = (size_t)d * d;
  for (int idx = tid; idx < d * d; idx += nth) {
    const int rr = idx / d, jj = idx - rr * d;
    Zt[mat * dd + idx] = sZ[pm[jj] * ld + rr];
  }
}

// ---------------------------------------------------------------------------------------
// K3 (register-resident variant, d <= 96): thread r keeps row r of Zt in REGISTERS (D doubles).
// The column loop is fully unrolled so every register index is static; a rotation is executed
// when its column lies in the current sweep [l, m) (a uniform branch).  No shared-memory
// traffic for Zt at all -- the shared-memory version above is LSU-bound (ncu: 65 % LSU, 21 %
// FP64) -- only the broadcast read of (c, s) from the cp.async ring remains.
// ---------------------------------------------------------------------------------------
// ring geometry: 4 tiles of 512 (d > 32) or 128 (d <= 32) entries; D guard entries in front of
// the ring and D mirrored entries behind it (a sweep / reflector has < D entries)
#define HQL_RTILE_OF(D) ((D) <= 32 ? 128 : 512)
#define HQL_RTILES 4
#define HQL_ARTILE_OF(D, NTH) ((NTH) < (D) ? 256 : HQL_RTILE_OF(D))  // rotation tile of the replay kernel

// One rotation of the replay, (z_j, z_j1) <- (cy z_j1 + cx z_j, cx z_j1 - cy z_j), with both results
// written IN PLACE ("+d"): every z[j] then stays in one physical register for the whole kernel.
// Left to itself the compiler rotates register names along the dependent chain, and the merge
// points of the unrolled block structure pay for it in moves (ncu: 21 % of all instructions
// were IMAD.MOV).  The dependent chain is still one FMA per rotation.
__device__ __forceinline__ void rot_inplace(double &zj, double &zj1, const double2 cs) {
  double nu;
  asm("{\n\t.reg .f64 ncy;\n\tneg.f64 ncy, %1;\n\tmul.rn.f64 %0, ncy, %2;\n\t}" : "=d"(nu) : "d"(cs.y), "d"(zj));
  asm("mul.rn.f64 %0, %0, %1;" : "+d"(zj) : "d"(cs.x));
  asm("fma.rn.f64 %0, %1, %2, %0;" : "+d"(zj) : "d"(cs.y), "d"(zj1));
  asm("fma.rn.f64 %0, %1, %0, %2;" : "+d"(zj1) : "d"(cs.x), "d"(nu));
}

// NTH = rows (threads) per CTA.  The register file is split per SM sub-partition (16 K registers
// each): at 224 registers a sub-partition holds two warps, so 3-warp CTAs (NTH = D = 96) leave a
// quarter of the slots empty (ncu: 2 CTAs = 6 warps per SM).  NTH = 32 makes every warp its own
// CTA -- the rows of Z are independent, the warps of a matrix only share the rotation stream,
// which the second and third warp then find in L2 -- and fills all 8 slots.
template <int D, int NTH = D>
__global__ void __maxnreg__(255)
hql_apply_reg_kernel(int d, const double2 *__restrict__ rot, size_t rot_cap,
                     const SweepIdx *__restrict__ swp, int swp_cap, const int *__restrict__ nswp,
                     double *__restrict__ Zt) {
  extern __shared__ __align__(16) unsigned char apply_smem[];
  constexpr int HQL_RTILE = HQL_ARTILE_OF(D, NTH), HQL_RPAD = D;
  constexpr int RING = HQL_RTILE * HQL_RTILES;
  // [RPAD guard][RING][RPAD mirror of the first entries]: a sweep's <= 95 rotations starting
  // anywhere in the ring are contiguous, so every (c, s) load is base + immediate offset
  double2 *ring = reinterpret_cast<double2 *>(apply_smem) + HQL_RPAD;
  SweepIdx *sswp = reinterpret_cast<SweepIdx *>(ring + RING + HQL_RPAD);
  const int tid = threadIdx.x;
  const int row0 = blockIdx.y * NTH;  // first row of Z held by this CTA
  const size_t mat = blockIdx.x;
  const double2 *myrot = rot + mat * rot_cap;
  const int ns = nswp[mat];
  for (int i = tid; i < ns; i += NTH) sswp[i] = swp[mat * swp_cap + i];
  __syncthreads();
  size_t total = 0;
  for (int i = 0; i < ns; ++i) {  // padded to whole 8-column blocks by the QL kernel
    const int bh = (sswp[i].m - 1) >> 3, bl = sswp[i].l >> 3;
    total += (size_t)(((8 * bh + 7 < D - 2) ? 8 * bh + 7 : D - 2) - 8 * bl + 1);
  }
  const int ntiles = (int)((total + HQL_RTILE - 1) / HQL_RTILE);
  int t_issued = 0, t_landed = 0;
  auto issue = [&]() {  // issue tile t_issued into its ring slot (+ mirror if it is slot 0)
    const size_t base = (size_t)t_issued * HQL_RTILE;
    const int slot0 = (t_issued % HQL_RTILES) * HQL_RTILE;
    for (int e = tid; e < HQL_RTILE; e += NTH)
      if (base + e < total) {
        cp_async16(&ring[slot0 + e], &myrot[base + e]);
        if (slot0 == 0 && e < HQL_RPAD) cp_async16(&ring[RING + e], &myrot[base + e]);
      }
    cp_async_commit();
    ++t_issued;
  };
  while (t_issued < ntiles && t_issued < HQL_RTILES) issue();

  double z[D];
#pragma unroll
  for (int j = 0; j < D; ++j) z[j] = (j == row0 + tid) ? 1.0 : 0.0;

  // Sweeps are replayed in PAIRS (A = sweep s, B = sweep s + 1), B one 8-column block behind A:
  // a rotation of B at column j only needs A's rotations at columns j - 1, j, j + 1, so with A
  // eight columns ahead the two dependent chains are independent and interleave (two FMAs in
  // flight instead of one, one set of block branches for both).  The QL kernel pads every sweep
  // to whole blocks with identity rotations, so a block is either active or not: no masks.
  constexpr int BMAX = (D - 2) / 8;
  auto blk_hi = [](int b) { return (8 * b + 7 < D - 2) ? 8 * b + 7 : D - 2; };
  auto single = [&](int b, const double2 *base) {  // loads first, then the dependent chain
    const int jlo = 8 * b, jhi = blk_hi(b);
    double2 cs[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (jhi - u >= jlo) cs[u] = base[-(jhi - u)];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (jhi - u >= jlo) rot_inplace(z[jhi - u], z[jhi - u + 1], cs[u]);
  };

  size_t g = 0;
  for (int sidx = 0; sidx < ns; sidx += 2) {
    const int blA = sswp[sidx].l >> 3, bhA = (sswp[sidx].m - 1) >> 3;
    int blB = 1 << 20, bhB = -1;  // no partner: never active
    size_t lenB = 0;
    if (sidx + 1 < ns) {
      blB = sswp[sidx + 1].l >> 3;
      bhB = (sswp[sidx + 1].m - 1) >> 3;
      lenB = (size_t)(blk_hi(bhB) - 8 * blB + 1);
    }
    const size_t gB = g + (size_t)(blk_hi(bhA) - 8 * blA + 1);
    const size_t need = gB + lenB;
    while ((size_t)t_landed * HQL_RTILE < need && t_landed < ntiles) {
      cp_async_wait<0>();
      __syncthreads();
      t_landed = t_issued;
      const int consumed = (int)(g / HQL_RTILE);  // tiles < consumed are dead for every thread
      // (the mirror of slot 0 is rewritten together with slot 0; a sweep that still reads the
      //  old mirror would start in the last tile, which is then not yet consumed)
      while (t_issued < ntiles && t_issued - consumed < HQL_RTILES - 1) issue();
    }
    // rotation for column j sits at base[-j]
    const double2 *baseA = ring + (int)(g % RING) + blk_hi(bhA);
    const double2 *baseB = ring + (int)(gB % RING) + blk_hi(bhB < 0 ? 0 : bhB);
#pragma unroll
    for (int b = BMAX; b >= -1; --b) {
      // A works on block b, B on block b + 1
      const bool actA = (b >= 0) && b >= blA && b <= bhA;
      const bool actB = (b + 1 <= BMAX) && b + 1 >= blB && b + 1 <= bhB;
      if (actA && actB) {
        const int jloA = 8 * b, jhiA = blk_hi(b);
        const int jloB = 8 * (b + 1), jhiB = blk_hi(b + 1);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          double2 ca[4], cb[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (jhiA - 4 * hf - u >= jloA) ca[u] = baseA[-(jhiA - 4 * hf - u)];
            if (jhiB - 4 * hf - u >= jloB) cb[u] = baseB[-(jhiB - 4 * hf - u)];
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int ja = jhiA - 4 * hf - u, jb = jhiB - 4 * hf - u;
            if (ja >= jloA) rot_inplace(z[ja], z[ja + 1], ca[u]);
            if (jb >= jloB) rot_inplace(z[jb], z[jb + 1], cb[u]);
          }
        }
      } else if (actA) {
        single(b, baseA);
      } else if (actB) {
        single(b + 1, baseB);
      }
      // QL windows sit at the high end (l grows as eigenvalues converge, m stays near d - 1):
      // stop once the remaining blocks lie below both windows
      if (b - 1 < blA && b < blB) break;
    }
    g = need;
  }
  cp_async_wait<0>();
  __syncthreads();
  // write Zt (unpermuted) through shared memory, 32 columns at a time, for coalescing
  double *sbuf = reinterpret_cast<double *>(ring);  // [D][33]
  const size_t dd = (size_t)d * d;
  for (int c0 = 0; c0 < d; c0 += 32) {
#pragma unroll
    for (int j = 0; j < D; ++j)
      if (j >= c0 && j < c0 + 32) sbuf[tid * 33 + (j - c0)] = z[j];
    __syncthreads();
    const int w = min(32, d - c0);
    for (int idx = tid; idx < NTH * 32; idx += NTH) {
      const int rr = idx >> 5, jj = idx & 31;
      if (jj < w && row0 + rr < d) Zt[mat * dd + (size_t)(row0 + rr) * d + c0 + jj] = sbuf[rr * 33 + jj];
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// K4 (d <= 96): back-transformation U = H_0 H_1 ... H_{d-2} Zt by applying the Householder
// reflectors directly to the columns of Zt.  Thread c keeps column c (D complex numbers) in
// REGISTERS; every reflector is a dot product and an axpy on that column, so there is no
// inter-thread communication and no barrier in the main loop -- unlike forming Q inside the
// tridiagonalisation kernel (a barrier-separated chain that held the whole SM) followed by a
// GEMM.  The reflector stream (v_k packed in consumption order, 71 KB at d = 96) is staged
// through the same mirrored cp.async ring as the rotation stream; rows enter through a
// fall-through switch at row k+1 so the work shrinks with the reflector length.
// ---------------------------------------------------------------------------------------
#define HQL_ROWS(X)                                                                           \
  X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) \
      X(17) X(18) X(19) X(20) X(21) X(22) X(23)

// Thread (c, h) = (tid / TPC, tid % TPC) holds rows i = TPC jj + h of column c (D/TPC complex
// numbers); the TPC partial dot products of a column are combined with log2(TPC) shuffles.
// TPC = 8 at D = 96: 768 threads = 24 warps per SM instead of 12 -- the kernel is latency bound.
template <int D, int TPC, int CPT = 1>
__global__ void __launch_bounds__(TPC *D / CPT)
hql_reflect_kernel(int d, const double *__restrict__ Zt, const cplx *__restrict__ Vp, size_t vcap,
                   const cplx *__restrict__ tau, cplx *__restrict__ U) {
  // CPT columns per thread: every reflector entry read from shared memory is used for CPT
  // columns and is kept in registers between the dot-product pass and the update pass (ncu on
  // the CPT = 1 version at d = 96: LSU 65 %, FP64 38 %)
  constexpr int RPT = D / TPC;  // rows per thread
  static_assert(RPT % 4 == 0, "rows per thread in blocks of 4");
  constexpr int NT = TPC * D / CPT;
  constexpr int BR = 4 * TPC;  // rows covered by a block of 4 jj
  constexpr int REGS = (65536 / NT > 255) ? 255 : 65536 / NT;
  constexpr bool KEEPV = (CPT >= 2) || (4 * RPT * (CPT + 1) + 40 <= REGS);  // else v is re-read in the update pass
  extern __shared__ __align__(16) unsigned char refl_smem[];
  constexpr int HQL_RTILE = HQL_RTILE_OF(D), HQL_RPAD = D;
  constexpr int RING = HQL_RTILE * HQL_RTILES;
  cplx *ring = reinterpret_cast<cplx *>(refl_smem) + HQL_RPAD;
  cplx *stau = ring + RING + HQL_RPAD;  // [D]
  const int tid = threadIdx.x;
  const int c0 = (tid / TPC) * CPT, hh = tid % TPC;
  const size_t mat = blockIdx.x;
  const size_t dd = (size_t)d * d;
  const cplx *myv = Vp + mat * vcap;
  const size_t total = (size_t)(d - 1) * (d - 2) / 2;
  const int ntiles = (int)((total + HQL_RTILE - 1) / HQL_RTILE);
  int t_issued = 0, t_landed = 0;
  auto issue = [&]() {
    const size_t base = (size_t)t_issued * HQL_RTILE;
    const int slot0 = (t_issued % HQL_RTILES) * HQL_RTILE;
    for (int e = tid; e < HQL_RTILE; e += NT)
      if (base + e < total) {
        cp_async16(&ring[slot0 + e], &myv[base + e]);
        if (slot0 == 0 && e < HQL_RPAD) cp_async16(&ring[RING + e], &myv[base + e]);
      }
    cp_async_commit();
    ++t_issued;
  };
  while (t_issued < ntiles && t_issued < HQL_RTILES) issue();
  for (int k = tid; k < d; k += NT) stau[k] = tau[mat * d + k];

  cplx x[CPT][RPT];
#pragma unroll
  for (int cc = 0; cc < CPT; ++cc)
#pragma unroll
    for (int jj = 0; jj < RPT; ++jj) {
      const int i = TPC * jj + hh, c = c0 + cc;
      x[cc][jj] = make_c((i < d && c < d) ? Zt[mat * dd + (size_t)i * d + c] : 0.0, 0.0);
    }
  __syncthreads();

  size_t g = 0;  // stream position of v_k[2:]
  for (int k = d - 2; k >= 0; --k) {
    const int mk = d - k - 2;  // explicit entries (rows k+2 .. d-1)
    const size_t need = g + (size_t)mk;
    while ((size_t)t_landed * HQL_RTILE < need && t_landed < ntiles) {
      cp_async_wait<0>();
      __syncthreads();
      t_landed = t_issued;
      const int consumed = (int)(g / HQL_RTILE);
      while (t_issued < ntiles && t_issued - consumed < HQL_RTILES - 1) issue();
    }
    const cplx t = stau[k];
    // v for row i (i >= k+2) sits at vb0[i]; this thread's rows are i = TPC jj + hh
    const cplx *vb = ring + (int)(g % RING) - (k + 2) + hh;
    // rows in blocks of 4 jj: one uniform branch per block; blocks that lie entirely inside
    // (k+1, d) run without per-row predicates.  vv[] holds v (1 at row k+1, 0 outside the
    // reflector) for the update pass.
    cplx vv[KEEPV ? RPT : 1];
    cplx u[CPT], ub[CPT];  // two partial dot products per column (even / odd rows): half the dependent chain
#pragma unroll
    for (int cc = 0; cc < CPT; ++cc) u[cc] = ub[cc] = make_c(0.0, 0.0);
#pragma unroll
    for (int b = 0; b < RPT / 4; ++b) {
      if (BR * b + BR - 1 >= k + 1 && BR * b < d) {
        if (BR * b > k + 1 && BR * b + BR - 1 < d) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int J = 4 * b + q;
            const cplx v = vb[TPC * J];
            if (KEEPV) vv[J] = v;
#pragma unroll
            for (int cc = 0; cc < CPT; ++cc) {
              if (q & 1)
                ccfma(ub[cc], v, x[cc][J]);
              else
                ccfma(u[cc], v, x[cc][J]);
            }
          }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int J = 4 * b + q;
            const int i = TPC * J + hh;
            cplx v = make_c(0.0, 0.0);
            if (i > k + 1 && i < d)
              v = vb[TPC * J];
            else if (i == k + 1)
              v = make_c(1.0, 0.0);
            if (KEEPV) vv[J] = v;
#pragma unroll
            for (int cc = 0; cc < CPT; ++cc) {
              if (q & 1)
                ccfma(ub[cc], v, x[cc][J]);
              else
                ccfma(u[cc], v, x[cc][J]);
            }
          }
        }
      }
    }
    cplx tu[CPT];
#pragma unroll
    for (int cc = 0; cc < CPT; ++cc) {
      u[cc] = cadd(u[cc], ub[cc]);
#pragma unroll
      for (int o = 1; o < TPC; o <<= 1) {
        u[cc].x += __shfl_xor_sync(0xffffffffu, u[cc].x, o);
        u[cc].y += __shfl_xor_sync(0xffffffffu, u[cc].y, o);
      }
      tu[cc] = cmul(t, u[cc]);
    }
#pragma unroll
    for (int b = 0; b < RPT / 4; ++b) {
      if (BR * b + BR - 1 >= k + 1 && BR * b < d) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int J = 4 * b + q;
          cplx v;
          if (KEEPV) {
            v = vv[J];
          } else {
            const int i = TPC * J + hh;
            v = make_c(i == k + 1 ? 1.0 : 0.0, 0.0);
            if (i > k + 1 && i < d) v = vb[TPC * J];
          }
#pragma unroll
          for (int cc = 0; cc < CPT; ++cc) {  // x -= v (tau u)
            x[cc][J].x = fma(-v.x, tu[cc].x, x[cc][J].x);
            x[cc][J].x = fma(v.y, tu[cc].y, x[cc][J].x);
            x[cc][J].y = fma(-v.x, tu[cc].y, x[cc][J].y);
            x[cc][J].y = fma(-v.y, tu[cc].x, x[cc][J].y);
          }
        }
      }
    }
    g = need;
  }
#pragma unroll
  for (int cc = 0; cc < CPT; ++cc) {
    const int c = c0 + cc;
    if (c < d) {
#pragma unroll
      for (int jj = 0; jj < RPT; ++jj) {
        const int i = TPC * jj + hh;
        if (i < d) U[mat * dd + (size_t)i * d + c] = x[cc][jj];
      }
    }
  }
}

inline size_t hql_reflect_smem(int D) {
  return (HQL_RTILE_OF(D) * HQL_RTILES + 2 * D + D) * sizeof(cplx) + 16;
}

inline size_t hql_apply_reg_smem(int D, int NTH, int swp_cap) {
  return (HQL_ARTILE_OF(D, NTH) * HQL_RTILES + 2 * D) * sizeof(double2) + (size_t)swp_cap * sizeof(SweepIdx) + 16;
}

inline size_t hql_apply_smem(int d, int swp_cap) {
  return 2 * HQL_TILE * sizeof(double2) + (size_t)d * (d | 1) * sizeof(double) + (size_t)swp_cap * sizeof(SweepIdx) + 16;
}

}  // namespace musim
