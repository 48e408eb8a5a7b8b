// tdc_core.cuh -- numerical core of the tridiagonal DIVIDE AND CONQUER eigensolver (host + device).
//
// Second stage of np.linalg.eigh (LAPACK zheevd -> dstedc, /root/reference/muspinsim/spinop.py:69):
// eigenvalues and eigenvectors of the real symmetric tridiagonal matrix the Householder reduction
// leaves.  T is torn into NLEAF diagonal blocks (Cuppen),
//     T = blockdiag(T_0', .., T_{NLEAF-1}') + sum_tears |beta| u u^T,  u = e_last + sign(beta) e_first,
// the leaves are solved by implicit QL (short serial chains that run concurrently), and pairs of
// blocks are merged level by level: each merge is the eigenproblem of D + rho z z^T,
//   * deflation (dlaed2): components with rho |z_i| <= tol, and close eigenvalue pairs (one plane
//     rotation each), drop out;
//   * the remaining k roots of the secular equation 1 + rho sum z_i^2 / (d_i - lambda) = 0 are
//     found INDEPENDENTLY (one thread each), each in the frame of its nearest pole so that the
//     differences d_i - lambda keep full relative accuracy;
//   * z is recomputed from the computed roots (Gu / Eisenstat, dlaed3), which makes the eigenvectors
//     v_j = (zhat_i / (d_i - lambda_j))_i orthogonal to working precision;
//   * Q <- Q [V 0; 0 I] is a GEMM (DMMA on the device).
// Everything in this header is plain scalar code shared by the CUDA kernel (eigh_tdc.cuh) and by a
// host build (tdc_host.cpp -> tests/test_tdc_host.py compares it with LAPACK without a GPU).  The
// restated algorithms follow the published LAPACK routines named above (netlib LAPACK 3.x: dsteqr /
// dlaev2 for the leaves, dlaed2 / dlaed4 / dlaed3 for a merge); the root finder is our own
// (two-pole + linear-remainder model with a safeguarded bracket) instead of dlaed4's rational
// interpolation cases.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define TDC_HD __host__ __device__ __forceinline__
#else
#define TDC_HD inline
#endif

namespace musim {
namespace tdc {

constexpr double EPS = 1.1102230246251565e-16;  // dlamch('Epsilon'): relative machine precision 2^-53

#if defined(__CUDA_ARCH__)
// 1 / x to ~1 ulp without the special-case path of the IEEE division (ncu: the full sequence and its
// branch were 40 % of the instructions of the secular loop): MUFU.RCP64H seed (>= 20 bits) + 2 Newton
// steps.  Only for arguments that are neither subnormal nor huge (pole distances of a unit-norm problem).
__device__ __forceinline__ double tdc_fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}
#define TDC_RCP(x) ::musim::tdc::tdc_fast_rcp(x)
#define TDC_RSQRT(x) rsqrt(x)
#else
#define TDC_RCP(x) (1.0 / (x))
#define TDC_RSQRT(x) (1.0 / sqrt(x))
#endif

// A group of P threads (P = 1 on the host, 4 adjacent lanes on the device) shares one root / one
// entry of z: thread `part` takes the poles i = part, part + P, ... and the partial results are
// combined by the reducer (sum / product over the group, result in every member).
struct SerialGroup {
  static constexpr int P = 1;
  TDC_HD int part() const { return 0; }
  TDC_HD double sum(double v) const { return v; }
  TDC_HD double prod(double v) const { return v; }
};

#if defined(TDC_STATS) && !defined(__CUDA_ARCH__)
static long g_sec_outer = 0, g_sec_inner = 0, g_sec_roots = 0, g_sec_hist[32] = {0};  // host-only iteration counters (tests)
#define TDC_COUNT(x) (++(x))
#else
#define TDC_COUNT(x) ((void)0)
#endif

// 2 x 2 symmetric eigenproblem [[a, b], [b, c]] (LAPACK dlaev2): rt1 >= rt2 in absolute value,
// (cs1, sn1) the unit eigenvector of rt1.
TDC_HD void laev2(double a, double b, double c, double &rt1, double &rt2, double &cs1, double &sn1) {
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

// Implicit QL on the leaf (D[0..n), E[0..n), E[n-1] ignored) with the eigenvector rows updated through
// `rows` (EISPACK tql2 / LAPACK dsteqr, same arithmetic as hql_tql_kernel):
//   rows.begin(m)       f = Z[.][m];  cur = Z[.][m-1]
//   rows.load(j)        nxt = Z[.][j]                      (issued BEFORE the scalar chain of the rotation)
//   rows.rot(j, cx, cy) a = cur;  Z[.][j+1] = cx f - cy a;  f = cy f + cx a;  cur = nxt
//   rows.end(l)         Z[.][l] = f
// On the device every lane of a warp runs the scalar chain redundantly (no communication) and owns
// one row.  The values the NEXT rotation needs (e, d, the row entry) are fetched at the top of the
// current one and carried in registers, so that no shared-memory latency sits on the serial chain
// (first device version without this: 1 300 cycles per rotation, ncu).  Returns false if the
// iteration limit is hit.
template <class Rows>
TDC_HD bool leaf_ql(int n, double *D, double *E, Rows &rows) {
  const double eps2 = 4.930380657631324e-32;  // (2^-52)^2
  const double safmin = 2.2250738585072014e-308;
  if (n > 0) E[n - 1] = 0.0;
  int l = 0, nit = 0;
  const int maxit = 60 * n;
  while (l < n) {
    int m = l;
    while (m < n - 1) {
      const double em = E[m];
      if (em * em <= (eps2 * fabs(D[m])) * fabs(D[m + 1]) + safmin) break;
      ++m;
    }
    if (m < n - 1) E[m] = 0.0;
    if (m == l) {
      ++l;
      continue;
    }
    if (nit >= maxit) return false;
    if (m == l + 1) {
      double rt1, rt2, c, s;
      laev2(D[l], E[l], D[l + 1], rt1, rt2, c, s);
      rows.begin(l + 1);
      rows.rot(l, c, s);
      rows.end(l);
      D[l] = rt1;
      D[l + 1] = rt2;
      E[l] = 0.0;
      l += 2;
      continue;
    }
    ++nit;
    double p = D[l];
    const double el_l = E[l];
    double g = (D[l + 1] - p) / (2.0 * el_l);
    double r = sqrt(fma(g, g, 1.0));
    g = D[m] - p + el_l / (g + copysign(r, g));
    double s = 1.0, c = 1.0;
    p = 0.0;
    // carried / prefetched values: d_ip1 = D[i+1] (untouched so far in this sweep), e_i, d_i
    double d_ip1 = D[m];
    double e_i = E[m - 1], d_i = D[m - 1];
    rows.begin(m);
    for (int i = m - 1; i >= l; --i) {
      double e_nx = 0.0, d_nx = 0.0;
      if (i > l) {
        e_nx = E[i - 1];
        d_nx = D[i - 1];
        rows.load(i - 1);
      }
      const double f = s * e_i;
      const double b = c * e_i;
      const double q = fma(g, g, f * f);
      if (q == 0.0) {
        c = 1.0;
        s = 0.0;
        r = 0.0;
      } else {
        const double ri = TDC_RSQRT(q);
        r = q * ri;
        c = g * ri;
        s = f * ri;
      }
      if (i != m - 1) E[i + 1] = r;
      g = d_ip1 - p;
      r = fma(d_i - g, s, 2.0 * c * b);
      p = s * r;
      D[i + 1] = g + p;
      g = fma(c, r, -b);
      rows.rot(i, c, -s);
      d_ip1 = d_i;
      e_i = e_nx;
      d_i = d_nx;
    }
    rows.end(l);
    D[l] = D[l] - p;
    E[l] = g;
  }
  return true;
}

// Root j (0-based, ascending) of g(lambda) = 1 + sum_i w_i / (dk_i - lambda) = 0, w_i = rho z_i^2, for
// strictly increasing dk[0..k), non-zero z, rho > 0.  Returns the index `org` of the pole nearest to
// the root and *mu = lambda - dk[org]; dk_i - lambda must then be formed as (dk_i - dk_org) - mu.
//
// Iteration: with L, R the poles that bracket the root, psi = sum_{i <= L} and phi = sum_{i >= R} are each
// replaced by a ONE-pole model  r + s / (p - x)  and the resulting quadratic is solved in closed form for the
// step -- no inner iteration: every pass is ONE sweep over the poles (4 interleaved reciprocal chains per lane
// on the device) plus a fixed tail.
//   * third-order step: r, s AND the pole p match value, slope and curvature of psi (phi) at the current point
//     (p - x = 2 psi' / psi'').  Cubic convergence, and a cluster of poles behind the neighbouring one is
//     represented by a pole at its weighted harmonic-mean distance.
//   * "middle way" step (Li 1994 / LAPACK dlaed4; the only step until the second half of round 2): p fixed at the
//     neighbouring pole d_L (d_R), s = (d_L - mu)^2 psi'(mu).  Fallback when the fitted model's root leaves the
//     bracket (a neighbouring pole of tiny weight is invisible in the sums far away from it).  With the fixed
//     poles alone a pole of small weight next to a cluster made the iteration creep towards the root from one
//     side (C5 Hamiltonians: 4.77 sweeps per root, 8 % of the roots >= 7 and the slowest of a merge's 96 roots
//     ~9; now 3.7, 0.6 % and ~6.5: the CTA waits for its slowest root).
// A first version refined a two-pole + linear model by Newton steps: 4.4 serial inner steps of ~80 dependent
// instructions per pass were half of the merge kernel.
// Safeguards: a sign bracket [lo, hi] on g (bisection if both model steps leave it), stop when |g| is within
// its rounding error bound, when the bracket collapses, or when the step is below 1e-7 |mu| (third-order
// step) / 2e-9 |mu| (middle way): the next error is below an ulp.
template <class Group>
TDC_HD int secular_root(int k, const double *dk, const double *wk, int j, double *mu_out, const Group &grp) {
  if (k == 1) {
    *mu_out = wk[0];
    return 0;
  }
  const bool last = (j == k - 1);
  const int L = last ? k - 2 : j, R = L + 1;  // the last root lies to the right of both of its poles
  const double gap = dk[R] - dk[L];
  int org;
  double dL, dR, lo, hi, mu;
  if (!last) {
    org = L;  // first evaluation at the midpoint, in the frame of the left pole; it also picks the frame
    dL = 0.0;
    dR = gap;
    lo = 0.0;
    hi = 0.5 * gap;
    mu = hi;
  } else {
    org = R;
    dL = -gap;
    dR = 0.0;
    double s = 0.0;
    for (int i = 0; i < k; ++i) s += wk[i];
    lo = 0.0;
    hi = s;  // lambda_max <= dk[k-1] + rho |z|^2
    mu = 0.5 * hi;
  }
  double dorg = dk[org];
  TDC_COUNT(g_sec_roots);
  for (int it = 0; it < 60; ++it) {
    TDC_COUNT(g_sec_outer);
    // one sweep over the poles, left part (i <= L: psi) and right part (i >= R: phi) separately: value, slope
    // and curvature (sum w / delta^3 = psi'' / 2).  Every term of psi has the sign of the others, likewise phi,
    // so |psi| + |phi| is the sum of the absolute terms the rounding error bound needs.
    double psi = 0.0, phi = 0.0, dpsi = 0.0, dphi = 0.0, ddpsi = 0.0, ddphi = 0.0;
    {
      const int p = grp.part();
      // first index >= R with the residue of this lane
      const int iR = R + ((p - R) % Group::P + Group::P) % Group::P;
#pragma unroll 8
      for (int i = p; i <= L; i += Group::P) {
        const double rdel = TDC_RCP((dk[i] - dorg) - mu);
        const double t = wk[i] * rdel;
        psi += t;
        const double u = t * rdel;
        dpsi += u;
        ddpsi = fma(u, rdel, ddpsi);
      }
#pragma unroll 8
      for (int i = iR; i < k; i += Group::P) {
        const double rdel = TDC_RCP((dk[i] - dorg) - mu);
        const double t = wk[i] * rdel;
        phi += t;
        const double u = t * rdel;
        dphi += u;
        ddphi = fma(u, rdel, ddphi);
      }
    }
    psi = grp.sum(psi);
    phi = grp.sum(phi);
    dpsi = grp.sum(dpsi);
    dphi = grp.sum(dphi);
    ddpsi = grp.sum(ddpsi);
    ddphi = grp.sum(ddphi);
    const double g = 1.0 + (psi + phi);
    if (it == 0 && !last && g < 0.0) {
      // root in the right half: continue in the frame of the right pole, mu in [-gap/2, 0).  The
      // quantities just evaluated are frame independent (dk_i - lambda is the same point).
      org = R;
      dorg = dk[R];
      dL = -gap;
      dR = 0.0;
      mu = -0.5 * gap;
      lo = mu;
      hi = 0.0;
    }
    const double err = 8.0 * EPS * (1.0 + (fabs(psi) + fabs(phi)));
    if (fabs(g) <= err) break;
    if (g < 0.0)
      lo = mu;
    else
      hi = mu;
    if (hi - lo <= 2.0 * EPS * fmax(fabs(lo), fabs(hi))) break;
    // step eta from the model  c + s / (DL - eta) + S / (DR - eta) = 0, i.e. the quadratic
    //   c eta^2 - (c (DL + DR) + s + S) eta + DL DR g = 0;
    // the wanted root keeps x = mu + eta strictly inside the bracket
    auto model_step = [&](double DL, double DR, double sL, double sR, double c, double &x) -> bool {
      const double Bq = -(c * (DL + DR) + sL + sR), Cq = DL * DR * g;
      double disc = Bq * Bq - 4.0 * c * Cq;
      disc = disc > 0.0 ? sqrt(disc) : 0.0;
      const double q = -0.5 * (Bq + (Bq >= 0.0 ? disc : -disc));
      // reciprocals by TDC_RCP: a step only steers the iteration; an overflowing or NaN candidate fails the
      // bracket test below and the next model (or bisection) takes over
      const double e1 = (c != 0.0) ? q * TDC_RCP(c) : 0.0, e2 = (q != 0.0) ? Cq * TDC_RCP(q) : 0.0;
      const double x1 = mu + e1, x2 = mu + e2;
      const bool ok1 = (c != 0.0) && x1 > lo && x1 < hi, ok2 = (q != 0.0) && x2 > lo && x2 < hi;
      x = (ok1 && (!ok2 || fabs(e1) < fabs(e2))) ? x1 : x2;
      return ok1 || ok2;
    };
    // (1) third-order model: poles FITTED to the curvature, distance 2 psi' / psi'' = dpsi / ddpsi (a weighted
    // harmonic mean of the true distances, so never nearer than the neighbouring pole: the model stays monotone
    // between its poles like g itself)
    double x = 0.0;
    bool fitted = false;
    if (ddpsi != 0.0 && ddphi != 0.0) {
      const double DL = dpsi * TDC_RCP(ddpsi), DR = dphi * TDC_RCP(ddphi);
      fitted = model_step(DL, DR, DL * DL * dpsi, DR * DR * dphi, g - DL * dpsi - DR * dphi, x);
    }
    if (!fitted) {
      // (2) poles fixed at the two neighbours: when the nearest pole carries a tiny weight it is invisible in
      // the sums at this point and the fitted model puts the root beyond it
      TDC_COUNT(g_sec_inner);
      const double DL = dL - mu, DR = dR - mu;
      bool ok;
      if (it == 0) {
        // initial guess of dlaed4: the two neighbouring poles with their TRUE weights, the rest constant
        const double sL = wk[L], sR = wk[R];
        ok = model_step(DL, DR, sL, sR, g - sL / DL - sR / DR, x);
      } else {
        ok = model_step(DL, DR, DL * DL * dpsi, DR * DR * dphi, g - DL * dpsi - DR * dphi, x);
      }
      if (!ok) x = 0.5 * (lo + hi);
    }
    // cubic (fitted) / quadratic convergence: the error after this step is below an ulp
    const bool small_step = fabs(x - mu) <= (fitted ? 1.0e-7 : 2.0e-9) * fabs(x);
    mu = x;
    if (small_step) break;
  }
#if defined(TDC_STATS) && !defined(__CUDA_ARCH__)
  {
    static thread_local long last_outer = 0;
    long n_it = g_sec_outer - last_outer;
    last_outer = g_sec_outer;
    ++g_sec_hist[n_it < 31 ? n_it : 31];
  }
#endif
  *mu_out = mu;
  return org;
}

// Gu / Eisenstat: the z that makes the COMPUTED roots exact (dlaed3).  diff(i, j) = dk_i - lambda_j
// = (dk_i - dk[org_j]) - mu_j.  `sgn` supplies the sign (the original z_i).
template <class Group>
TDC_HD double zhat(int k, const double *dk, const double *mu, const int *org, double rho, int i, double sgn,
                   const Group &grp) {
  double prod = (grp.part() == 0) ? -((dk[i] - dk[org[i]]) - mu[i]) / rho : 1.0;  // (lambda_i - dk_i) / rho > 0
#pragma unroll 4
  for (int j = grp.part(); j < k; j += Group::P) {
    const double num = (dk[i] - dk[org[j]]) - mu[j], den = dk[i] - dk[j];
    prod *= (j == i) ? 1.0 : num * TDC_RCP(den);
  }
  prod = grp.prod(prod);
  return copysign(sqrt(fabs(prod)), sgn);
}

// 1 / || (zh_i / diff(i, j))_i ||
template <class Group>
TDC_HD double inv_colnorm(int k, const double *dk, const double *zh, double mu_j, double dorg_j, const Group &grp) {
  double s = 0.0;
#pragma unroll 4
  for (int i = grp.part(); i < k; i += Group::P) {
    const double v = zh[i] * TDC_RCP((dk[i] - dorg_j) - mu_j);
    s = fma(v, v, s);
  }
  s = grp.sum(s);
  return 1.0 / sqrt(s);
}

struct RotRec {
  int p, q;  // sorted positions: column p is deflated, q survives (LAPACK dlaed2: PJ, NJ)
  double c, s;
};

// Sequential deflation scan of dlaed2 over the merge's entries in ascending order of sD.
// flag[p]: 0 candidate, 1 deflated because rho |z| <= tol (set by the caller); the scan sets 2 for
// entries deflated by a rotation.  Returns the number of rotations recorded: columns (p, q) of Q are
// then to be rotated as  col_p' = c col_p + s col_q,  col_q' = c col_q - s col_p  (drot).
TDC_HD int deflate_scan(int n, double *sD, double *sZ, unsigned char *flag, double tol, RotRec *rots) {
  int pj = -1, nr = 0;
  for (int p = 0; p < n; ++p) {
    if (flag[p]) continue;
    if (pj < 0) {
      pj = p;
      continue;
    }
    double s = sZ[pj], c = sZ[p];
    const double tau = sqrt(c * c + s * s);
    const double t = sD[p] - sD[pj];
    c /= tau;
    s = -s / tau;
    if (fabs(t * c * s) <= tol) {
      sZ[p] = tau;
      sZ[pj] = 0.0;
      rots[nr].p = pj;
      rots[nr].q = p;
      rots[nr].c = c;
      rots[nr].s = s;
      ++nr;
      const double tt = sD[pj] * c * c + sD[p] * s * s;
      sD[p] = sD[pj] * s * s + sD[p] * c * c;
      sD[pj] = tt;
      flag[pj] = 2;
    }
    pj = p;
  }
  return nr;
}

// The closeness test of the scan for one adjacent pair of candidates with their ORIGINAL values: if
// it fails for every pair, the scan would not rotate anything and can be skipped.
TDC_HD bool close_pair(double d_prev, double z_prev, double d_cur, double z_cur, double tol) {
  const double tau = sqrt(z_cur * z_cur + z_prev * z_prev);
  const double c = z_cur / tau, s = -z_prev / tau;
  return fabs((d_cur - d_prev) * c * s) <= tol;
}

// Leaf boundaries: the d rows are cut at multiples of 8 (row tiles of the DMMA GEMMs) into 4 leaves
// of at most 24 rows (d <= 96), sizes as even as the tiles allow.  bnd[0..4].
TDC_HD void leaf_bounds(int d, int *bnd) {
  const int tiles = (d + 7) / 8;
  int t0 = 0;
  bnd[0] = 0;
  for (int q = 0; q < 4; ++q) {
    const int cnt = tiles / 4 + (q < tiles % 4 ? 1 : 0);
    t0 += cnt;
    bnd[q + 1] = (8 * t0 < d) ? 8 * t0 : d;
  }
  bnd[4] = d;
}

}  // namespace tdc
}  // namespace musim
