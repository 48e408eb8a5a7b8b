// tdc_host.cpp -- HOST build of the divide-and-conquer core (tdc_core.cuh), phase by phase the way
// the CUDA kernel of eigh_tdc.cuh runs it, with plain loops where the kernel uses threads and
// DMMAs.  TEST INFRASTRUCTURE: built into oracle/_build/libtdc_host.so by tests/test_tdc_host.py
// (g++, no GPU) and compared with LAPACK there; nothing in the product path links it.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "tdc_core.cuh"

using namespace musim::tdc;

namespace {

struct HostRows {  // all rows of one leaf at once (the device gives every lane one row)
  double *base;    // QT block: element (row r, column j) at base[j * ld + r]
  int ld, n;
  std::vector<double> f, cur, nxt;
  void begin(int m) {
    for (int r = 0; r < n; ++r) {
      f[r] = base[m * ld + r];
      cur[r] = base[(m - 1) * ld + r];
    }
  }
  void load(int j) {
    for (int r = 0; r < n; ++r) nxt[r] = base[j * ld + r];
  }
  void rot(int j, double cx, double cy) {
    for (int r = 0; r < n; ++r) {
      const double a = cur[r];
      base[(j + 1) * ld + r] = cx * f[r] - cy * a;
      f[r] = cy * f[r] + cx * a;
      cur[r] = nxt[r];
    }
  }
  void end(int l) {
    for (int r = 0; r < n; ++r) base[l * ld + r] = f[r];
  }
};

struct Stats {
  int deflated = 0, rotations = 0, scans = 0;
};

// One merge of the blocks [lo, mid) and [mid, hi): D holds their eigenvalues, QT (transposed, leading
// dimension ld) their eigenvectors as diagonal blocks.  beta = off-diagonal element of the tear.
void merge(int lo, int mid, int hi, double beta, double *D, double *QT, int ld, Stats &st) {
  const int n = hi - lo;
  const double rho = 2.0 * fabs(beta);
  if (rho == 0.0 || n == 0) return;
  const double sgn = beta < 0.0 ? -1.0 : 1.0;
  std::vector<double> z(n);
  double dmax = 0.0, zmax = 0.0;
  for (int t = lo; t < hi; ++t) {
    z[t - lo] = (t < mid ? QT[t * ld + mid - 1] : sgn * QT[t * ld + mid]) * 0.70710678118654752440;
    dmax = fmax(dmax, fabs(D[t]));
    zmax = fmax(zmax, fabs(z[t - lo]));
  }
  const double tol = 8.0 * EPS * fmax(dmax, zmax);
  if (rho * zmax <= tol) {
    st.deflated += n;
    return;
  }
  // sort by rank counting
  std::vector<double> sD(n), sZ(n);
  std::vector<int> sidx(n);
  std::vector<unsigned char> flag(n);
  for (int t = 0; t < n; ++t) {
    int rank = 0;
    for (int i = 0; i < n; ++i) rank += (D[lo + i] < D[lo + t]) || (D[lo + i] == D[lo + t] && i < t);
    sD[rank] = D[lo + t];
    sZ[rank] = z[t];
    sidx[rank] = lo + t;
    flag[rank] = (rho * fabs(z[t]) <= tol) ? 1 : 0;
  }
  // pre-check: would the sequential scan rotate anything?
  bool any_close = false;
  for (int p = 0; p < n; ++p) {
    if (flag[p]) continue;
    int q = p - 1;
    while (q >= 0 && flag[q]) --q;
    if (q >= 0 && close_pair(sD[q], sZ[q], sD[p], sZ[p], tol)) any_close = true;
  }
  if (any_close) {
    std::vector<RotRec> rots(n);
    const int nr = deflate_scan(n, sD.data(), sZ.data(), flag.data(), tol, rots.data());
    st.scans++;
    st.rotations += nr;
    for (int r = 0; r < nr; ++r) {
      double *cp = QT + (size_t)sidx[rots[r].p] * ld, *cq = QT + (size_t)sidx[rots[r].q] * ld;
      for (int row = lo; row < hi; ++row) {
        const double x = cp[row], y = cq[row];
        cp[row] = rots[r].c * x + rots[r].s * y;
        cq[row] = rots[r].c * y - rots[r].s * x;
      }
    }
  }
  // compaction
  std::vector<double> dk, zk, ddef;
  std::vector<int> kcol, dcol;
  for (int p = 0; p < n; ++p) {
    if (!flag[p]) {
      dk.push_back(sD[p]);
      zk.push_back(sZ[p]);
      kcol.push_back(sidx[p]);
    } else {
      ddef.push_back(sD[p]);
      dcol.push_back(sidx[p]);
    }
  }
  const int k = (int)dk.size();
  st.deflated += n - k;
  std::vector<double> mu(k), zh(k), sn(k), lamn(k), wk(k);
  for (int i = 0; i < k; ++i) wk[i] = rho * zk[i] * zk[i];
  std::vector<int> org(k);
  for (int j = 0; j < k; ++j) {
    org[j] = secular_root(k, dk.data(), wk.data(), j, &mu[j], SerialGroup());
    lamn[j] = dk[org[j]] + mu[j];
  }
  for (int i = 0; i < k; ++i) zh[i] = zhat(k, dk.data(), mu.data(), org.data(), rho, i, zk[i], SerialGroup());
  for (int j = 0; j < k; ++j) sn[j] = inv_colnorm(k, dk.data(), zh.data(), mu[j], dk[org[j]], SerialGroup());
  // Q_new
  std::vector<double> out((size_t)n * n);  // [output column j][row]
  for (int j = 0; j < n; ++j)
    for (int r = 0; r < n; ++r) {
      double acc = 0.0;
      if (j < k) {
        for (int i = 0; i < k; ++i)
          acc = fma(QT[(size_t)kcol[i] * ld + lo + r], zh[i] * sn[j] / ((dk[i] - dk[org[j]]) - mu[j]), acc);
      } else {
        acc = QT[(size_t)dcol[j - k] * ld + lo + r];
      }
      out[(size_t)j * n + r] = acc;
    }
  for (int j = 0; j < n; ++j) {
    for (int r = 0; r < n; ++r) QT[(size_t)(lo + j) * ld + lo + r] = out[(size_t)j * n + r];
    D[lo + j] = j < k ? lamn[j] : ddef[j - k];
  }
}

}  // namespace

// Eigen-decomposition of the symmetric tridiagonal matrix (dd[0..d), ee[0..d-1)).  lam unsorted, Z
// row-major d x d with the eigenvector of lam[j] in column j.  stats[0..3) = deflated entries,
// deflation rotations, sequential scans (may be NULL).  Returns 0, or 1 if a leaf did not converge.
extern "C" int tdc_host_eigh(int d, const double *dd, const double *ee, double *lam, double *Z, int *stats) {
  const int ld = d;
  std::vector<double> D(dd, dd + d), E(d, 0.0), QT((size_t)d * ld, 0.0);
  for (int i = 0; i + 1 < d; ++i) E[i] = ee[i];
  // scale to unit max-norm (dstedc does the same): the deflation tolerance compares |D| with the
  // entries of the unit vector z
  double orgnrm = 0.0;
  for (int i = 0; i < d; ++i) orgnrm = fmax(orgnrm, fmax(fabs(D[i]), fabs(E[i])));
  const double scl = orgnrm > 0.0 ? 1.0 / orgnrm : 1.0;
  for (int i = 0; i < d; ++i) {
    D[i] *= scl;
    E[i] *= scl;
  }
  int bnd[5];
  leaf_bounds(d, bnd);
  double beta[5] = {0, 0, 0, 0, 0};
  for (int q = 1; q <= 3; ++q) {
    const int m = bnd[q];
    if (m > 0 && m < d && bnd[q] > bnd[q - 1]) {
      beta[q] = E[m - 1];
      D[m - 1] -= fabs(beta[q]);
      D[m] -= fabs(beta[q]);
      E[m - 1] = 0.0;
    }
  }
  int rc = 0;
  for (int q = 0; q < 4; ++q) {
    const int o = bnd[q], n = bnd[q + 1] - bnd[q];
    if (n <= 0) continue;
    for (int i = 0; i < n; ++i) QT[(size_t)(o + i) * ld + o + i] = 1.0;
    HostRows rows{QT.data() + (size_t)o * ld + o, ld, n, std::vector<double>(n), std::vector<double>(n), std::vector<double>(n)};
    if (!leaf_ql(n, D.data() + o, E.data() + o, rows)) rc = 1;
  }
  Stats st;
  merge(bnd[0], bnd[1], bnd[2], beta[1], D.data(), QT.data(), ld, st);
  merge(bnd[2], bnd[3], bnd[4], beta[3], D.data(), QT.data(), ld, st);
  merge(bnd[0], bnd[2], bnd[4], beta[2], D.data(), QT.data(), ld, st);
  for (int j = 0; j < d; ++j) {
    lam[j] = D[j] * (orgnrm > 0.0 ? orgnrm : 1.0);
    for (int r = 0; r < d; ++r) Z[(size_t)r * d + j] = QT[(size_t)j * ld + r];
  }
  if (stats) {
    stats[0] = st.deflated;
    stats[1] = st.rotations;
    stats[2] = st.scans;
  }
  return rc;
}

#ifdef TDC_STATS
extern "C" void tdc_host_counters(long *out) {
  out[0] = g_sec_roots;
  out[1] = g_sec_outer;
  out[2] = g_sec_inner;
  for (int i = 0; i < 32; ++i) out[3 + i] = g_sec_hist[i];
}
#endif
