// eigh_tdc.cuh -- K2 + K3 replaced: eigenvalues AND eigenvectors of the tridiagonal matrix by
// divide and conquer (LAPACK dstedc inside np.linalg.eigh, /root/reference/muspinsim/spinop.py:69),
// 32 < d <= 96.
//
// What it replaces and why.  The QL pipeline ran (i) one THREAD per matrix through a serial chain
// of ~1.2 d^2 plane rotations (hql_tql_kernel, 3.35 ms at C5 no matter how few matrices a GPU
// holds: the chain, not the batch, sets the time -- it is what limits strong scaling at 2 500
// matrices per GPU) and (ii) replayed the recorded rotations on d rows (hql_apply_reg_kernel,
// 8.8 ms, ~7 d^3 vector flops, 9 GB of rotation-stream traffic).  Here the matrix is torn into four
// leaves of <= 24 rows:
//   leaves            tdc_prep_kernel scales and tears the matrix into a batch of 4 n independent 24 x 24
//                     problems for the batched QL kernels (hql_tql_kernel: one THREAD per leaf, chains
//                     16 x shorter than the full one; hql_apply_reg_kernel<24, 24>: row-per-thread replay);
//   tdc_merge_kernel  one CTA per matrix, two merge levels; all merges of a level run concurrently.
//                     Secular roots, Gu/Eisenstat z and the column norms use FOUR adjacent lanes per
//                     root (each sums a quarter of the poles; quad shuffles), the eigenvector update
//                     Q <- Q [V 0; 0 I] is a DMMA GEMM whose B fragments (zhat_i / (d_i - lambda_j))
//                     are formed on the fly; the row tiles of a K chunk sit in groups of three behind
//                     warp-uniform branches (a chunk supported on one block touches half or a quarter of
//                     them), the compaction counts come from warp ballots.
// The first version was ONE kernel with one thread per root and ran 31 ms at C5 (ncu: 49 % of the
// CTA lifetime in the leaf phase with 8 of 12 warps parked at the barrier, 38 % in the secular phase
// at one dependent instruction per 9 cycles); hence the split and the quads.
//
// Numerics are in tdc_core.cuh (shared with the host build that tests/test_tdc_host.py checks against
// LAPACK); this file is the parallel orchestration.  Shared memory of the merge kernel: the
// eigenvector matrix is kept TRANSPOSED, QsT[column][row] with leading dimension D + 4 (= 4 mod 16):
// the GEMM's A fragments (row = lane/4, k = lane%4 -> 4 columns) are bank-conflict free for
// consecutive columns.  Thread t < d owns global index t (an entry of z, a sorted position) of the
// merge that contains t; quad g = tid / 4 owns root g; warp w owns output columns 8w .. 8w+7.
// Output: eigenvalues in ASCENDING order, Zt row-major with eigenvectors in columns.
#pragma once
#include "common.cuh"
#include "polar.cuh"  // dmma884
#include "tdc_core.cuh"

namespace musim {

#define TDC_LEAF_MAX 24  // rows of a leaf (d <= 96 in four leaves cut at multiples of 8)

template <int D>
struct TdcGeom {
  static constexpr int LD = D + 4;
  static constexpr int NT = 4 * D;  // D / 8 warps
  static constexpr size_t smem_bytes =
      (size_t)D * LD * sizeof(double)      // QsT
      + 13 * (size_t)D * sizeof(double)    // Dv zv sD sZ nd zk wgt mu dorg zh sn lamn fin
      + 4 * (size_t)D * sizeof(int)        // sidx ncol orgi dest
      + 2 * (size_t)(D + 16) * sizeof(int) // red: one region per concurrent merge (group padding: up to 9 extra entries)
      + (size_t)D * sizeof(tdc::RotRec)    // rots
      + 2 * (size_t)D                      // flag kind
      + 64 * sizeof(double);               // per-merge scalars
};

struct QuadGroup {  // four adjacent lanes share one root (tdc_core.cuh)
  static constexpr int P = 4;
  int p;
  unsigned mask;
  __device__ __forceinline__ int part() const { return p; }
  __device__ __forceinline__ double sum(double v) const {
    v += __shfl_xor_sync(mask, v, 1);
    v += __shfl_xor_sync(mask, v, 2);
    return v;
  }
  __device__ __forceinline__ int isum(int v) const {
    v += __shfl_xor_sync(mask, v, 1);
    v += __shfl_xor_sync(mask, v, 2);
    return v;
  }
  __device__ __forceinline__ double max(double v) const {
    v = fmax(v, __shfl_xor_sync(mask, v, 1));
    v = fmax(v, __shfl_xor_sync(mask, v, 2));
    return v;
  }
  __device__ __forceinline__ double prod(double v) const {
    v *= __shfl_xor_sync(mask, v, 1);
    v *= __shfl_xor_sync(mask, v, 2);
    return v;
  }
};

// ---------------------------------------------------------------------------------------
// Leaves.  tdc_prep_kernel scales the matrix to unit max-norm (dstedc), tears it and writes the four
// leaves as a batch of 4 n independent 24 x 24 tridiagonal problems (smaller leaves are padded with
// decoupled zero rows), which the batched QL kernels solve: hql_tql_kernel (one THREAD per leaf: 32
// scalar chains per warp instruction, rotations recorded) + hql_apply_reg_kernel<24> (one thread
// per eigenvector row, rows in registers).  A first version ran the QL chain redundantly in every
// lane of one warp per leaf and updated the rows in shared memory: 126 instructions per rotation per
// WARP, 257 k warp instructions per matrix, 5.7 ms at C5 (FP64 pipe 53 %, issue bound).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
tdc_prep_kernel(int d, int64_t n, const double *__restrict__ din, const double *__restrict__ ein,
                double *__restrict__ hdr, double *__restrict__ dleaf, double *__restrict__ eleaf) {
  constexpr int LM = TDC_LEAF_MAX;
  const int lane = threadIdx.x & 31;
  const int64_t mat = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);  // one warp per matrix
  if (mat >= n) return;
  const double *dd = din + mat * d, *ee = ein + mat * d;
  double mx = 0.0;
  for (int i = lane; i < d; i += 32) mx = fmax(mx, fmax(fabs(dd[i]), i < d - 1 ? fabs(ee[i]) : 0.0));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const double scl = mx > 0.0 ? 1.0 / mx : 1.0;
  int bnd[5];
  tdc::leaf_bounds(d, bnd);
  double beta[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int q = 1; q <= 3; ++q) {
    const int m = bnd[q];
    if (m > 0 && m < d && bnd[q] > bnd[q - 1]) beta[q] = ee[m - 1] * scl;
  }
  if (lane == 0) {
    hdr[mat * 4 + 0] = mx;
    hdr[mat * 4 + 1] = beta[1];
    hdr[mat * 4 + 2] = beta[2];
    hdr[mat * 4 + 3] = beta[3];
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int o = bnd[q], nq = bnd[q + 1] - bnd[q];
    if (lane < LM) {
      double dv = 0.0, ev = 0.0;
      if (lane < nq) {
        dv = dd[o + lane] * scl;
        if (lane == 0 && q > 0) dv -= fabs(beta[q]);           // first row of the leaf: tear above
        if (lane == nq - 1 && q < 3) dv -= fabs(beta[q + 1]);  // last row: tear below
        if (lane < nq - 1) ev = ee[o + lane] * scl;
      }
      dleaf[(mat * 4 + q) * LM + lane] = dv;
      eleaf[(mat * 4 + q) * LM + lane] = ev;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Merges.
// ---------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(4 * D, 2)
tdc_merge_kernel(int d, const double *__restrict__ hdr, const double *__restrict__ lamleaf,
                 const double *__restrict__ Zleaf, double *__restrict__ lam, double *__restrict__ Zt) {
  using G = TdcGeom<D>;
  constexpr int LD = G::LD, NT = G::NT, NB = D / 8, LM = TDC_LEAF_MAX;
  extern __shared__ __align__(16) unsigned char tdc_smem[];
  double *QsT = reinterpret_cast<double *>(tdc_smem);
  double *Dv = QsT + D * LD, *zv = Dv + D, *sD = zv + D, *sZ = sD + D, *nd = sZ + D, *zk = nd + D, *wgt = zk + D, *mu = wgt + D,
         *dorg = mu + D, *zh = dorg + D, *sn = zh + D, *lamn = sn + D, *fin = lamn + D;
  double *scal = fin + D;
  int *sidx = reinterpret_cast<int *>(scal + 64), *ncol = sidx + D, *orgi = ncol + D, *dest = orgi + D, *red = dest + D;
  constexpr int RED = D + 16;  // stride of a merge slot's reduction list
  tdc::RotRec *rots = reinterpret_cast<tdc::RotRec *>(red + 2 * RED);
  unsigned char *flag = reinterpret_cast<unsigned char *>(rots + D);
  unsigned char *kind = flag + D;  // support of a column of Q: 0 top block only, 1 bottom block only, 2 both (after a deflation rotation)
  // per-merge scalars (slot m = 0, 1): rho, tol; beta[1..3]; ints: k, skip, anyclose, nrot
  double *m_rho = scal, *m_tol = scal + 2, *m_beta = scal + 4;
  int *m_k = reinterpret_cast<int *>(scal + 16), *m_skip = m_k + 2, *m_close = m_k + 4, *m_nrot = m_k + 6;
  int *bnd = m_k + 8;  // [5]
  int *m_k0 = m_k + 16, *m_k1 = m_k + 18;  // survivors supported on the top / bottom block only

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fm = lane >> 2, fj = lane & 3;
  const size_t mat = blockIdx.x;

  // ---- load the leaves: eigenvalues, eigenvector blocks (row-major [r][c]) onto the diagonal of QsT ----
  for (int i = tid; i < D * LD; i += NT) QsT[i] = 0.0;
  if (tid == 0) tdc::leaf_bounds(d, bnd);
  if (tid < 4) m_beta[tid] = hdr[mat * 4 + tid];  // [0] = orgnrm, [1..3] = beta of the tears
  __syncthreads();
  const double orgnrm = m_beta[0];
  if (tid < d) {
    int q = 0;
    while (q < 3 && tid >= bnd[q + 1]) ++q;
    Dv[tid] = lamleaf[(mat * 4 + q) * LM + (tid - bnd[q])];
  }
  for (int i = tid; i < 4 * LM * LM; i += NT) {
    const int q = i / (LM * LM), rem = i - q * LM * LM;
    const int r = rem / LM, c = rem - r * LM;
    const int o = bnd[q], nq = bnd[q + 1] - bnd[q];
    if (r < nq && c < nq) QsT[(o + c) * LD + o + r] = Zleaf[(mat * 4 + q) * LM * LM + rem];
  }

  QuadGroup grp;
  grp.p = tid & 3;
  grp.mask = 0xFu << (lane & ~3);
  const int g = tid >> 2;  // root / z entry handled by this quad (phases P6 - P8)

  for (int level = 1; level <= 2; ++level) {
    __syncthreads();
    auto describe = [&](int idx, int &ms, int &l, int &m, int &h) {
      if (level == 2) {
        ms = 0;
        l = bnd[0];
        m = bnd[2];
        h = bnd[4];
      } else if (idx < bnd[2]) {
        ms = 0;
        l = bnd[0];
        m = bnd[1];
        h = bnd[2];
      } else {
        ms = 1;
        l = bnd[2];
        m = bnd[3];
        h = bnd[4];
      }
    };
    int mslot, lo, mid, hi;
    describe(tid, mslot, lo, mid, hi);
    const bool mine = tid < d;
    const int n = hi - lo;
    // P1: z, rho
    if (mine) {
      const double beta = (level == 2) ? m_beta[2] : (mslot == 0 ? m_beta[1] : m_beta[3]);
      const double sg = beta < 0.0 ? -1.0 : 1.0;
      zv[tid] = (tid < mid ? QsT[tid * LD + mid - 1] : sg * QsT[tid * LD + mid]) * 0.70710678118654752440;
      kind[tid] = tid < mid ? 0 : 1;
      if (tid == lo) {
        m_rho[mslot] = 2.0 * fabs(beta);
        m_close[mslot] = 0;
        m_nrot[mslot] = 0;
      }
    }
    __syncthreads();
    // P2: tolerance, first deflation test, sort by rank.  Quad g / 4 lanes per element (the O(n) scans per
    // element ran on d of the 4 d threads while the other warps waited at the barrier)
    int gslot, glo, gmid, ghi;
    describe(g, gslot, glo, gmid, ghi);
    // max |D|, max |z| of the merge once per warp (the merge of a warp's eight quads is warp-uniform), the rank
    // per quad
    double dmax = 0.0, zmax = 0.0;
    if (8 * warp < d) {
      for (int i = glo + lane; i < ghi; i += 32) {
        dmax = fmax(dmax, fabs(Dv[i]));
        zmax = fmax(zmax, fabs(zv[i]));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
        zmax = fmax(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
      }
    }
    if (g < d) {
      const double rho = m_rho[gslot];
      const double my = Dv[g];
      int rank = 0;
      for (int i = glo + grp.p; i < ghi; i += 4) {
        const double di = Dv[i];
        rank += (di < my) || (di == my && i < g);
      }
      rank = grp.isum(rank);
      const double tol = 8.0 * tdc::EPS * fmax(dmax, zmax);
      const bool skip_ = (rho == 0.0) || (rho * zmax <= tol);
      if (grp.p == 0) {
        sD[glo + rank] = my;
        sZ[glo + rank] = zv[g];
        sidx[glo + rank] = g;
        flag[glo + rank] = (rho * fabs(zv[g]) <= tol) ? 1 : 0;
        if (g == glo) {
          m_tol[gslot] = tol;
          m_skip[gslot] = skip_ ? 1 : 0;
        }
      }
    }
    __syncthreads();
    const bool skip = mine ? (m_skip[mslot] != 0) : true;
    // P3: would the sequential deflation scan rotate anything?
    if (mine && !skip) {
      const int p = tid;  // sorted position (global)
      if (!flag[p]) {
        int q = p - 1;
        while (q >= lo && flag[q]) --q;
        if (q >= lo && tdc::close_pair(sD[q], sZ[q], sD[p], sZ[p], m_tol[mslot])) m_close[mslot] = 1;
      }
    }
    __syncthreads();
    // P4: the scan itself (one thread per merge; rare), then the rotations on the columns of Q
    if (mine && !skip && tid == lo && m_close[mslot]) {
      const int nr = tdc::deflate_scan(n, sD + lo, sZ + lo, flag + lo, m_tol[mslot], rots + lo);
      m_nrot[mslot] = nr;
      for (int r = 0; r < nr; ++r) {  // a rotation between the blocks makes both columns dense
        const int cp = sidx[lo + rots[lo + r].p], cq = sidx[lo + rots[lo + r].q];
        if (kind[cp] != kind[cq]) kind[cp] = kind[cq] = 2;
      }
    }
    __syncthreads();
    if (mine && !skip) {
      const int nr = m_nrot[mslot];
      for (int r = 0; r < nr; ++r) {
        const tdc::RotRec rr = rots[lo + r];
        double *cp = QsT + sidx[lo + rr.p] * LD, *cq = QsT + sidx[lo + rr.q] * LD;
        const double x = cp[tid], y = cq[tid];
        cp[tid] = rr.c * x + rr.s * y;
        cq[tid] = rr.c * y - rr.s * x;
      }
    }
    // P5: compaction into the new order [survivors (ascending), deflated].  The counts (survivors of every
    // support kind in the merge, and of those in front of position g) come from warp ballots over the sorted
    // positions -- bit masks of 32 positions each, built once per warp for its merge -- instead of a loop of
    // n / 4 iterations per lane (source-level ncu: the loop was 7 % of the kernel's instructions).  The merge of a
    // warp's eight quads is warp-uniform (leaf boundaries are multiples of 8).
    if (8 * warp < d && m_skip[gslot] == 0) {
      constexpr int NW = (D + 31) / 32;
      int k0 = 0, k1 = 0, k2 = 0, b0 = 0, b1 = 0, b2 = 0;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        const int pos = 32 * w + lane;
        int kd = 3;
        if (pos >= glo && pos < ghi && pos < D && !flag[pos]) kd = kind[sidx[pos]];
        const unsigned s0 = __ballot_sync(0xffffffffu, kd == 0), s1 = __ballot_sync(0xffffffffu, kd == 1),
                       s2 = __ballot_sync(0xffffffffu, kd == 2);
        // positions of this word in front of g
        const int rel = g - 32 * w;
        const unsigned below = rel <= 0 ? 0u : (rel >= 32 ? 0xffffffffu : ((1u << rel) - 1u));
        k0 += __popc(s0);
        k1 += __popc(s1);
        k2 += __popc(s2);
        b0 += __popc(s0 & below);
        b1 += __popc(s1 & below);
        b2 += __popc(s2 & below);
      }
      const int k = k0 + k1 + k2, before = b0 + b1 + b2;
      if (grp.p == 0 && g < d) {
        const bool surv = !flag[g];
        const int i = surv ? before : k + ((g - glo) - before);
        nd[glo + i] = sD[g];
        zk[glo + i] = sZ[g];
        wgt[glo + i] = m_rho[gslot] * sZ[g] * sZ[g];
        ncol[glo + i] = sidx[g];
        // reduction order of the GEMM: survivors grouped by support [top | bottom | both], every group
        // padded to whole k-chunks of 4 (entry -1): a chunk then touches only the row tiles of its block
        const int p0 = (k0 + 3) & ~3, p1 = (k1 + 3) & ~3, p2 = (k - k0 - k1 + 3) & ~3;
        if (surv) {
          const int kd = kind[sidx[g]];
          red[gslot * RED + (kd == 0 ? b0 : (kd == 1 ? p0 + b1 : p0 + p1 + b2))] = i;
        }
        // padding entries (at most 3 per group), written by the first elements of the merge
        {
          const int t3 = g - glo;
          if (t3 < p0 - k0) red[gslot * RED + k0 + t3] = -1;
          if (t3 < p1 - k1) red[gslot * RED + p0 + k1 + t3] = -1;
          if (t3 < p2 - (k - k0 - k1)) red[gslot * RED + p0 + p1 + (k - k0 - k1) + t3] = -1;
        }
        if (g == glo) {
          m_k[gslot] = k;
          m_k0[gslot] = k0;
          m_k1[gslot] = k1;
        }
      }
    }
    __syncthreads();
    // quad g owns root / entry g of the merge that contains g
    const bool gact = (g < d) && (m_skip[gslot] == 0);
    const int gk = gact ? m_k[gslot] : 0;
    const int gj = g - glo;
    // P6: secular roots
    if (gact && gj < gk) {
      double m_;
      const int og = tdc::secular_root(gk, nd + glo, wgt + glo, gj, &m_, grp);
      if (grp.p == 0) {
        mu[g] = m_;
        orgi[g] = og;
        dorg[g] = nd[glo + og];
        lamn[g] = nd[glo + og] + m_;
      }
    }
    __syncthreads();
    // P7: Gu / Eisenstat z
    if (gact && gj < gk) {
      const double v = tdc::zhat(gk, nd + glo, mu + glo, orgi + glo, m_rho[gslot], gj, zk[g], grp);
      if (grp.p == 0) zh[g] = v;
    }
    __syncthreads();
    // P8: column norms
    if (gact && gj < gk) {
      const double v = tdc::inv_colnorm(gk, nd + glo, zh + glo, mu[g], dorg[g], grp);
      if (grp.p == 0) sn[g] = v;
    }
    __syncthreads();

    // P9: Q_new = Q [V 0; 0 I] (new column order): warp w owns output columns 8w .. 8w+7
    int wslot, wlo, wmid, whi;
    describe(8 * warp, wslot, wlo, wmid, whi);
    const bool wactive = 8 * warp < d;
    const bool wskip = wactive ? (m_skip[wslot] != 0) : true;
    double acc[NB][2];
#pragma unroll
    for (int t = 0; t < NB; ++t) acc[t][0] = acc[t][1] = 0.0;
    if (wactive && !wskip) {
      const int wk = m_k[wslot];
      const int jb = 8 * warp + fm;  // B operand's output column
      const bool jv = (jb - wlo) < wk && jb < whi;
      const double s_j = jv ? sn[jb] : 0.0, mu_j = jv ? mu[jb] : 0.0, do_j = jv ? dorg[jb] : 0.0;
      const int tlo = wlo >> 3, tmid = wmid >> 3, thi = (whi + 7) >> 3;
      const int c0 = (m_k0[wslot] + 3) >> 2, c1 = (m_k1[wslot] + 3) >> 2;
      const int c2 = (wk - m_k0[wslot] - m_k1[wslot] + 3) >> 2;
      for (int q = 0; q < c0 + c1 + c2; ++q) {
        // chunk of the top group: row tiles [tlo, tmid); bottom group: [tmid, thi); dense columns: all
        const int ta = (q >= c0 && q < c0 + c1) ? tmid : tlo, tb = (q < c0) ? tmid : thi;
        const int i = red[wslot * RED + 4 * q + fj];
        double bval = 0.0;
        int acol = wlo;
        if (i >= 0) {
          acol = ncol[wlo + i];
          if (jv) bval = zh[wlo + i] * s_j * TDC_RCP((nd[wlo + i] - do_j) - mu_j);
        }
        const double *ap = QsT + acol * LD + fm;
        // row tiles in groups of GS behind a warp-uniform branch: a chunk of the top (bottom) group touches half
        // of the tiles at the top level and a quarter at the first level, and predicated-off LDS / DMMA pairs
        // still take issue slots (source-level ncu: this loop was 28 % of the kernel's instructions for 2.6 % DMMAs)
        constexpr int GS = (NB % 3 == 0) ? 3 : 2;
#pragma unroll
        for (int tg = 0; tg < NB; tg += GS) {
          if (tb > tg && ta < tg + GS) {
#pragma unroll
            for (int t = tg; t < tg + GS; ++t)
              if (t < NB && t >= ta && t < tb) dmma884(acc[t][0], acc[t][1], ap[8 * t], bval);
          }
        }
      }
      // deflated columns are copied
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int jj = 8 * warp + 2 * fj + e;
        if (jj < whi && (jj - wlo) >= wk) {
          const double *cp = QsT + ncol[jj] * LD + fm;
#pragma unroll
          for (int t = 0; t < NB; ++t)
            if (t >= tlo && t < thi) acc[t][e] = cp[8 * t];
        }
      }
    }
    __syncthreads();  // every read of the old Q is done
    if (level == 1) {
      if (wactive && !wskip) {
        const int wk = m_k[wslot];
        const int tlo = wlo >> 3, thi = (whi + 7) >> 3;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int jj = 8 * warp + 2 * fj + e;
          if (jj < whi) {
#pragma unroll
            for (int t = 0; t < NB; ++t)
              if (t >= tlo && t < thi) QsT[jj * LD + 8 * t + fm] = acc[t][e];
            if (fm == 0) Dv[jj] = (jj - wlo) < wk ? lamn[jj] : nd[jj];
          }
        }
      }
    } else {
      // ---- top level: final eigenvalues, ascending order, output ----
      const int k = (mine && !skip) ? m_k[mslot] : 0;
      if (mine) fin[tid] = skip ? Dv[tid] : ((tid - lo) < k ? lamn[tid] : nd[tid]);
      __syncthreads();
      if (g < d) {
        const double my = fin[g];
        int rank = 0;
        for (int i = grp.p; i < d; i += 4) {
          const double fi = fin[i];
          rank += (fi < my) || (fi == my && i < g);
        }
        rank = grp.isum(rank);
        if (grp.p == 0) {
          dest[g] = rank;
          lam[mat * d + rank] = my * (orgnrm > 0.0 ? orgnrm : 1.0);
        }
      }
      if (wactive && wskip) {  // nothing was merged at the top: Q is unchanged, column jj stays column jj
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int jj = 8 * warp + 2 * fj + e;
          if (jj < d) {
#pragma unroll
            for (int t = 0; t < NB; ++t) acc[t][e] = QsT[jj * LD + 8 * t + fm];
          }
        }
      }
      __syncthreads();
      // stage row-major in shared memory (the old Q is dead), then write coalesced
      if (wactive) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int jj = 8 * warp + 2 * fj + e;
          if (jj < d) {
            const int dj = dest[jj];
#pragma unroll
            for (int t = 0; t < NB; ++t) QsT[(8 * t + fm) * LD + dj] = acc[t][e];  // now [row][column]
          }
        }
      }
      __syncthreads();
      const size_t dd = (size_t)d * d;
      {  // (r, c) of idx = tid + j NT advanced without a division per element
        const int qn = NT / d, rn = NT - qn * d;
        int r = tid / d, c = tid - r * d;
        for (int idx = tid; idx < d * d; idx += NT) {
          Zt[mat * dd + idx] = QsT[r * LD + c];
          r += qn;
          c += rn;
          if (c >= d) {
            c -= d;
            ++r;
          }
        }
      }
    }
  }
}

}  // namespace musim
