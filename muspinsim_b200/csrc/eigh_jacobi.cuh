// eigh_jacobi.cuh -- batched complex-Hermitian eigensolver, parallel cyclic Jacobi in shared
// memory, one CTA per matrix.  Replaces np.linalg.eigh in Hermitian.diag
// (/root/reference/muspinsim/spinop.py:51-82) for small d; also the robust cross-check for
// the Householder+QL solver (eigh_hql.cuh).
//
// Layout: A and V live in shared memory column-major with an odd leading dimension
// (element (r,c) at [c*ld + r]) so that both the column phase (threads along r) and the row
// phase (threads along c) are free of bank conflicts.
// Ordering: round-robin tournament, n-1 rounds of n/2 disjoint pairs per sweep.
// Convergence: off-diagonal Frobenius mass <= (2^-53)^2 * ||A||_F^2 (quadratic convergence
// makes the last sweep cheap); exact zeros are skipped so degenerate / diagonal inputs
// terminate immediately.
#pragma once
#include "common.cuh"

namespace musim {

struct JacobiRot {
  double c;
  cplx s;  // s * exp(i*phi)
};

__device__ __forceinline__ void rr_pair(int n, int round, int k, int &p, int &q) {
  // circle method: n even, rounds 0..n-2, pairs 0..n/2-1
  if (k == 0) {
    p = n - 1;
    q = round;
  } else {
    p = (round + k) % (n - 1);
    q = (round - k + (n - 1)) % (n - 1);
  }
  if (p > q) {
    int t = p;
    p = q;
    q = t;
  }
}

// If BUILD_H: A = H0 + B[0] Z0 + B[1] Z1 + B[2] Z2 for configuration blockIdx.x
// else:       A = Ain[blockIdx.x]
template <bool BUILD_H>
__global__ void __launch_bounds__(256)
eigh_jacobi_kernel(int d, const cplx *__restrict__ H0, const cplx *__restrict__ Z,
                   const double *__restrict__ Bf, const cplx *__restrict__ Ain,
                   double *__restrict__ evals, cplx *__restrict__ U, int *__restrict__ status,
                   int max_sweeps, cplx *Vglobal) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ld = d | 1;
  cplx *sA = reinterpret_cast<cplx *>(smem_raw);
  // V in shared memory when it fits, else in a per-matrix global (L2-resident) workspace
  cplx *sV = Vglobal ? Vglobal + (size_t)blockIdx.x * d * ld : sA + (size_t)d * ld;
  JacobiRot *rot = reinterpret_cast<JacobiRot *>(sA + (size_t)d * ld * (Vglobal ? 1 : 2));
  const int n = d + (d & 1);
  const int npair = n / 2;
  int *pp = reinterpret_cast<int *>(rot + npair);
  int *qq = pp + npair;
  double *red = reinterpret_cast<double *>(qq + npair);  // 2*npair ints: 8-byte aligned
  double *lam = red + 34;
  const int tid = threadIdx.x, nth = blockDim.x;
  const size_t cfg = blockIdx.x;
  const size_t dd = (size_t)d * d;

  double bx = 0, by = 0, bz = 0;
  if (BUILD_H) {
    bx = Bf[cfg * 3 + 0];
    by = Bf[cfg * 3 + 1];
    bz = Bf[cfg * 3 + 2];
  }
  for (int idx = tid; idx < d * d; idx += nth) {
    const int r = idx / d, c = idx - r * d;
    cplx a;
    if (BUILD_H) {
      a = H0[idx];
      cplx z0 = Z[idx], z1 = Z[dd + idx], z2 = Z[2 * dd + idx];
      a.x += bx * z0.x + by * z1.x + bz * z2.x;
      a.y += bx * z0.y + by * z1.y + bz * z2.y;
    } else {
      a = Ain[cfg * dd + idx];
    }
    sA[c * ld + r] = a;
    sV[c * ld + r] = make_c(r == c ? 1.0 : 0.0, 0.0);
  }
  __syncthreads();
  // symmetrise (use the average of (r,c) and conj(c,r)) so tiny asymmetries do not matter
  for (int idx = tid; idx < d * d; idx += nth) {
    const int r = idx / d, c = idx - r * d;
    if (r < c) {
      cplx a = sA[c * ld + r], b = sA[r * ld + c];
      cplx m = make_c(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
      sA[c * ld + r] = m;
      sA[r * ld + c] = cconj(m);
    } else if (r == c) {
      sA[c * ld + r].y = 0.0;
    }
  }
  __syncthreads();

  const double eps2 = 1.2325951644078309e-32;  // (2^-53)^2
  int sweep = 0;
  bool converged = false;
  for (; sweep < max_sweeps; ++sweep) {
    double off = 0.0, tot = 0.0;
    for (int idx = tid; idx < d * d; idx += nth) {
      const int c = idx / d, r = idx - c * d;
      const double v = cnorm2(sA[c * ld + r]);
      tot += v;
      if (r != c) off += v;
    }
    off = block_sum(off, red);
    tot = block_sum(tot, red);
    if (off <= eps2 * tot) {
      converged = true;
      break;
    }
    for (int round = 0; round < n - 1; ++round) {
      // 1. rotation parameters for this round's pairs
      for (int k = tid; k < npair; k += nth) {
        int p, q;
        rr_pair(n, round, k, p, q);
        JacobiRot R;
        R.c = 1.0;
        R.s = make_c(0.0, 0.0);
        if (q < d) {
          const cplx b = sA[q * ld + p];  // A(p,q)
          const double ab2 = cnorm2(b);
          if (ab2 > 0.0) {
            const double ab = sqrt(ab2);
            const double app = sA[p * ld + p].x, aqq = sA[q * ld + q].x;
            const double tau = (aqq - app) / (2.0 * ab);
            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            const double c = 1.0 / sqrt(1.0 + t * t);
            const double s = t * c;
            R.c = c;
            R.s = make_c(s * b.x / ab, s * b.y / ab);
          }
        } else {
          p = q = -1;
        }
        rot[k] = R;
        pp[k] = p;
        qq[k] = q;
      }
      __syncthreads();
      // 2. column phase on A and V:  [X_p, X_q] <- [X_p, X_q] J
      for (int idx = tid; idx < npair * d * 2; idx += nth) {
        const int r = idx % d;
        const int k = (idx / d) % npair;
        const int which = idx / (d * npair);
        const int p = pp[k], q = qq[k];
        if (p < 0) continue;
        const JacobiRot R = rot[k];
        if (R.s.x == 0.0 && R.s.y == 0.0) continue;
        cplx *X = which ? sV : sA;
        const cplx xp = X[p * ld + r], xq = X[q * ld + r];
        // x_p' = c x_p - conj(s) x_q ; x_q' = s x_p + c x_q
        cplx np_ = cscale(R.c, xp), nq_ = cscale(R.c, xq);
        const cplx t1 = ccmul(R.s, xq), t2 = cmul(R.s, xp);
        X[p * ld + r] = csub(np_, t1);
        X[q * ld + r] = cadd(nq_, t2);
      }
      __syncthreads();
      // 3. row phase on A:  [A_p; A_q] <- J^H [A_p; A_q]
      for (int idx = tid; idx < npair * d; idx += nth) {
        const int c = idx % d;
        const int k = idx / d;
        const int p = pp[k], q = qq[k];
        if (p < 0) continue;
        const JacobiRot R = rot[k];
        if (R.s.x == 0.0 && R.s.y == 0.0) continue;
        const cplx ap = sA[c * ld + p], aq = sA[c * ld + q];
        // a_p' = c a_p - s a_q ; a_q' = conj(s) a_p + c a_q
        cplx np_ = csub(cscale(R.c, ap), cmul(R.s, aq));
        cplx nq_ = cadd(cscale(R.c, aq), ccmul(R.s, ap));
        if (c == p) np_.y = 0.0;
        if (c == q) nq_.y = 0.0;
        if (c == q) np_ = make_c(0.0, 0.0);  // A(p,q) annihilated exactly
        if (c == p) nq_ = make_c(0.0, 0.0);  // A(q,p)
        sA[c * ld + p] = np_;
        sA[c * ld + q] = nq_;
      }
      __syncthreads();
    }
  }
  if (!converged && tid == 0 && status) atomicMax(status, 1);

  // eigenvalues ascending (rank sort), eigenvectors as columns of row-major U
  for (int i = tid; i < d; i += nth) lam[i] = sA[i * ld + i].x;
  __syncthreads();
  for (int i = tid; i < d; i += nth) {
    const double li = lam[i];
    int rk = 0;
    for (int j = 0; j < d; ++j) {
      const double lj = lam[j];
      rk += (lj < li) || (lj == li && j < i);
    }
    evals[cfg * d + rk] = li;
    // stash the rank in the (now unused) imaginary part of the diagonal
    sA[i * ld + i].y = (double)rk;
  }
  __syncthreads();
  for (int idx = tid; idx < d * d; idx += nth) {
    const int k = idx / d, i = idx - k * d;  // row k, source column i (coalesced over i)
    const int rk = (int)sA[i * ld + i].y;
    U[cfg * dd + (size_t)k * d + rk] = sV[i * ld + k];
  }
}

inline size_t eigh_jacobi_smem(int d, bool v_global) {
  const int ld = d | 1;
  const int n = d + (d & 1);
  const int npair = n / 2;
  size_t b = (v_global ? 1 : 2) * (size_t)d * ld * sizeof(cplx);
  b += npair * sizeof(JacobiRot);
  b += (2 * npair) * sizeof(int);
  b += (34 + d + 2) * sizeof(double);
  return b;
}

}  // namespace musim
