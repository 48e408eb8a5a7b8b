// lindblad.cuh -- dissipative (Lindbladian) evolution, batched over configurations.
//
// Reference: /root/reference/muspinsim/lindbladian.py:18-173 (Lindbladian.from_hamiltonian,
// add_dissipative_term, evolve, integrate_decaying), spinop.py:615-715 (super-operators, row-major
// vec) and ExperimentRunner.dissipation_operators (experiment.py:271-325).
//
// The reference diagonalises the non-Hermitian d^2 x d^2 matrix L with zgeev and solves for the
// expansion coefficients, which assumes L is diagonalisable.  Here the same quantities
//     P(t)      = vec(O^T)^T exp(2 pi L t) vec(rho0)
//     integral  = vec(O^T)^T (1/tau - 2 pi L)^-1 vec(rho0) / tau
// are evaluated without an eigen-decomposition:
//   * lind_build_kernel   builds L, vec(rho0), vec(O^T) per configuration on chip;
//   * E1 = exp(2 pi L dt) by scaling-and-squaring of a degree-12 Taylor polynomial
//     (Paterson-Stockmeyer), entirely as batched complex GEMMs (rotate.cuh);
//   * for a uniform grid t_k = t0 + (a NB + b) dt:  P[a,b] = (o^T E1^b) (EB^a v0), EB = E1^NB
//     (lind_series_kernel: 2*NB-1 mat-vecs per 32x32 block of time points; for d*d > 76 the
//     matrices are read in place from global memory and only the vectors are kept on chip);
//   * the integral is one LU solve with partial pivoting (lind_solve_kernel).
#pragma once
#include <string>

#include "common.cuh"
#include "profiler.cuh"
#include "rotate.cuh"
#include "zgemm_dmma.cuh"

namespace musim {

#define LIND_MAX_OPS 32

struct LindParams {
  int d;
  SpinTable tab;
  int n_diss;
  int diss_spin[MUSIM_MAX_SPINS];
  double diss_rate[MUSIM_MAX_SPINS];
  int n_explicit;  // explicit (configuration-independent) dissipators, see musim_set_dissipators
};

__device__ inline void get_xy_dev(double zx, double zy, double zz, double *x, double *y) {
  // muspinsim/utils.py:71-93
  if (zx == 0.0 && zy == 0.0) {
    x[0] = 1.0; x[1] = 0.0; x[2] = 0.0;
    y[0] = 0.0; y[1] = zz; y[2] = 0.0;
  } else {
    const double n = sqrt(zx * zx + zy * zy);
    x[0] = zy / n; x[1] = -zx / n; x[2] = 0.0;
    y[0] = zy * x[2] - zz * x[1];
    y[1] = zz * x[0] - zx * x[2];
    y[2] = zx * x[1] - zy * x[0];
  }
}

// One CTA per configuration.  Shared: H[d*d], A[nops][d*d], AA[nops][d*d], g[nops], rho factors.
__global__ void __launch_bounds__(256)
lind_build_kernel(LindParams P, const cplx *__restrict__ H0, const cplx *__restrict__ Z,
                  const cplx *__restrict__ M, const cplx *__restrict__ rho0_explicit,
                  const cplx *__restrict__ exA, const double *__restrict__ exg,
                  const double *__restrict__ Bf, const double *__restrict__ pf,
                  const double *__restrict__ Tf, cplx *__restrict__ L, cplx *__restrict__ r0,
                  cplx *__restrict__ ov, unsigned long long *__restrict__ norm_max) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int d = P.d, dd = d * d, n = dd;
  const int nops = 2 * P.n_diss + P.n_explicit;
  cplx *sH = reinterpret_cast<cplx *>(smem_raw);
  cplx *sA = sH + dd;                 // [nops][dd]
  cplx *sAA = sA + (size_t)nops * dd;  // [nops][dd]
  double *sg = reinterpret_cast<double *>(sAA + (size_t)nops * dd);  // [nops]
  double *red = sg + LIND_MAX_OPS;                                   // [34]
  __shared__ cplx fac[MUSIM_MAX_SPINS][MUSIM_MAX_SDIM * MUSIM_MAX_SDIM];
  const int tid = threadIdx.x, nth = blockDim.x;
  const size_t cfg = blockIdx.x;
  const double bx = Bf[cfg * 3], by = Bf[cfg * 3 + 1], bz = Bf[cfg * 3 + 2];
  const double T = Tf ? Tf[cfg] : INFINITY;
  const double Bn = sqrt(bx * bx + by * by + bz * bz);
  for (int idx = tid; idx < dd; idx += nth) {
    cplx a = H0[idx];
    const cplx z0 = Z[idx], z1 = Z[dd + idx], z2 = Z[2 * dd + idx];
    a.x += bx * z0.x + by * z1.x + bz * z2.x;
    a.y += bx * z0.y + by * z1.y + bz * z2.y;
    sH[idx] = a;
  }
  // jump operators S_+/- of each dissipated spin in the frame of B (experiment.py:299-323)
  double xv[3], yv[3];
  if (Bn == 0.0) {
    xv[0] = 1; xv[1] = 0; xv[2] = 0;
    yv[0] = 0; yv[1] = 1; yv[2] = 0;
  } else {
    get_xy_dev(bx / Bn, by / Bn, bz / Bn, xv, yv);
  }
  for (int k = 0; k < P.n_diss; ++k) {
    const int s = P.diss_spin[k];
    const int ns = P.tab.dims[s];
    int stride = 1;
    for (int q = s + 1; q < P.tab.n_spins; ++q) stride *= P.tab.dims[q];
    const double I = 0.5 * (ns - 1);
    for (int idx = tid; idx < dd; idx += nth) {
      const int row = idx / d, col = idx - row * d;
      const int mr = (row / stride) % ns, mc = (col / stride) % ns;
      cplx ox = make_c(0, 0), oy = make_c(0, 0);
      if (row - mr * stride == col - mc * stride) {  // all other spins unchanged
        // local (x.S) and (y.S) elements, spinop.py:14-41 conventions
        if (mr == mc) {
          const double m = I - mr;
          ox = make_c(xv[2] * m, 0.0);
          oy = make_c(yv[2] * m, 0.0);
        } else if (mc == mr + 1 || mr == mc + 1) {
          const int lo = mr < mc ? mr : mc;  // element (lo, lo+1) of S+
          const double mp = I - (lo + 1);
          const double sp = sqrt(I * (I + 1.0) - mp * (mp + 1.0));
          const double sgn = (mc == mr + 1) ? -1.0 : 1.0;  // (a,a+1): -i ny sp/2 ; (a+1,a): +i ny sp/2
          ox = make_c(0.5 * xv[0] * sp, sgn * 0.5 * xv[1] * sp);
          oy = make_c(0.5 * yv[0] * sp, sgn * 0.5 * yv[1] * sp);
        }
      }
      // op_p = op_x + i op_y ; op_m = op_x - i op_y
      sA[(size_t)(2 * k) * dd + idx] = make_c(ox.x - oy.y, ox.y + oy.x);
      sA[(size_t)(2 * k + 1) * dd + idx] = make_c(ox.x + oy.y, ox.y - oy.x);
    }
    if (tid == 0) {
      const double a = P.diss_rate[k];
      const double kB = 1.380649e-23, hP = 6.62607015e-34;
      double fp, fm;  // Zu/(1+Zu), 1/(1+Zu)
      if (T > 0.0) {
        const double x = isinf(T) ? 0.0 : hP * P.tab.gammas[s] * Bn * 1e6 / (kB * T);
        fp = 1.0 / (1.0 + exp(x));
        fm = 1.0 / (1.0 + exp(-x));
      } else {
        fp = 0.0;
        fm = 1.0;
      }
      sg[2 * k] = a * fp / 3.14159265358979323846;
      sg[2 * k + 1] = a * fm / 3.14159265358979323846;
    }
  }
  for (int k = 0; k < P.n_explicit; ++k) {
    for (int idx = tid; idx < dd; idx += nth) sA[(size_t)(2 * P.n_diss + k) * dd + idx] = exA[(size_t)k * dd + idx];
    if (tid == 0) sg[2 * P.n_diss + k] = exg[k];
  }
  __syncthreads();
  // AA = A^H A
  for (int idx = tid; idx < nops * dd; idx += nth) {
    const int m = idx / dd, e = idx - m * dd;
    const int i = e / d, j = e - i * d;
    const cplx *A = sA + (size_t)m * dd;
    cplx acc = make_c(0, 0);
    for (int k = 0; k < d; ++k) ccfma(acc, A[k * d + i], A[k * d + j]);
    sAA[idx] = acc;
  }
  __syncthreads();
  // L[(i,j),(k,l)]  (row-major vec, spinop.py:633,655,713)
  cplx *Lc = L + cfg * (size_t)n * n;
  double rowmax = 0.0;
  for (int I_ = tid; I_ < n; I_ += nth) {
    const int i = I_ / d, j = I_ - i * d;
    double rs = 0.0;
    for (int K = 0; K < n; ++K) {
      const int k = K / d, l = K - k * d;
      cplx v = make_c(0, 0);
      // -i (H_ik d_jl - d_ik H_lj)
      if (j == l) {
        const cplx h = sH[i * d + k];
        v.x += h.y;
        v.y -= h.x;
      }
      if (i == k) {
        const cplx h = sH[l * d + j];
        v.x -= h.y;
        v.y += h.x;
      }
      for (int m = 0; m < nops; ++m) {
        const cplx *A = sA + (size_t)m * dd, *AA = sAA + (size_t)m * dd;
        const double g = sg[m];
        cplx t = cmulc(A[i * d + k], A[j * d + l]);  // A_ik conj(A_jl)
        if (j == l) {
          t.x -= 0.5 * AA[i * d + k].x;
          t.y -= 0.5 * AA[i * d + k].y;
        }
        if (i == k) {
          t.x -= 0.5 * AA[l * d + j].x;
          t.y -= 0.5 * AA[l * d + j].y;
        }
        v.x += g * t.x;
        v.y += g * t.y;
      }
      Lc[(size_t)I_ * n + K] = v;
      rs += sqrt(cnorm2(v));
    }
    rowmax = fmax(rowmax, rs);
  }
  // block max of the row sums -> global max (inf-norm of L)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rowmax = fmax(rowmax, __shfl_xor_sync(0xffffffffu, rowmax, o));
  if ((tid & 31) == 0) red[tid >> 5] = rowmax;
  __syncthreads();
  if (tid == 0) {
    double mx = 0.0;
    for (int i = 0; i < (nth + 31) / 32; ++i) mx = fmax(mx, red[i]);
    atomicMax(norm_max, (unsigned long long)__double_as_longlong(mx));
  }
  // vec(O^T) and vec(rho0)
  const double px = pf[cfg * 3], py = pf[cfg * 3 + 1], pz = pf[cfg * 3 + 2];
  for (int idx = tid; idx < dd; idx += nth) {
    const int i = idx / d, j = idx - i * d;
    const int tr = j * d + i;
    const cplx m0 = M[tr], m1 = M[dd + tr], m2 = M[2 * dd + tr];
    ov[cfg * n + idx] = make_c(px * m0.x + py * m1.x + pz * m2.x, px * m0.y + py * m1.y + pz * m2.y);
  }
  if (rho0_explicit) {
    for (int idx = tid; idx < dd; idx += nth) r0[cfg * n + idx] = rho0_explicit[idx];
  } else {
    if (tid < P.tab.n_spins) {
      const int s = tid;
      if (s == P.tab.muon_index) {
        double qx = px, qy = py, qz = pz;
        const double pn = sqrt(qx * qx + qy * qy + qz * qz);
        if (pn > 0.0) {
          qx /= pn; qy /= pn; qz /= pn;
        }
        spin_dot(2, qx, qy, qz, fac[s]);
        fac[s][0].x += 0.5;
        fac[s][3].x += 0.5;
      } else {
        thermal_factor(P.tab.dims[s], P.tab.gammas[s], bx, by, bz, T, fac[s]);
      }
    }
    __syncthreads();
    for (int idx = tid; idx < dd; idx += nth) {
      int i = idx / d, j = idx - (idx / d) * d;
      cplx v = make_c(1.0, 0.0);
      for (int s = P.tab.n_spins - 1; s >= 0; --s) {
        const int ns = P.tab.dims[s];
        v = cmul(v, fac[s][(i % ns) * ns + (j % ns)]);
        i /= ns;
        j /= ns;
      }
      r0[cfg * n + idx] = v;
    }
  }
}

// out = c0 I + c1 X + c2 X2 + c3 X3 (+ c4 X4), elementwise over a batch of n x n matrices
__global__ void lind_poly_kernel(int n, size_t total, double c0, double c1, double c2, double c3,
                                 double c4, const cplx *__restrict__ X, const cplx *__restrict__ X2,
                                 const cplx *__restrict__ X3, const cplx *__restrict__ X4,
                                 cplx *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t e = i % ((size_t)n * n);
  const int r = (int)(e / n), c = (int)(e - (size_t)r * n);
  cplx v = make_c(r == c ? c0 : 0.0, 0.0);
  const cplx x = X[i], x2 = X2[i], x3 = X3[i];
  v.x += c1 * x.x + c2 * x2.x + c3 * x3.x;
  v.y += c1 * x.y + c2 * x2.y + c3 * x3.y;
  if (X4) {
    const cplx x4 = X4[i];
    v.x += c4 * x4.x;
    v.y += c4 * x4.y;
  }
  out[i] = v;
}

__global__ void lind_scale_kernel(size_t total, double s, const cplx *__restrict__ in, cplx *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) out[i] = cscale(s, in[i]);
}

// Uniform-grid time series, one CTA (256 threads) per configuration.
//   v_a = EB^a v0 (v0 = E0 r0 or r0), o_b^T = o^T E1^b, P[a,b] = Re(o_b . v_a)
__device__ __forceinline__ cplx shfl_xor_c4(cplx v, int m) {
  return make_c(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}

// Shared-memory path (GMEM = false): ONE padded matrix buffer, holding E1 while the o_b are built and
// then EB for the v_a chain, plus AB = 16 vectors of each kind: 101 KB at n = 64, so two
// configurations share an SM (the mat-vec chains are latency bound).
template <bool GMEM>
__global__ void __launch_bounds__(256)
lind_series_kernel(int n, const cplx *__restrict__ E1, const cplx *__restrict__ EB,
                   const cplx *__restrict__ E0, const cplx *__restrict__ r0, const cplx *__restrict__ ov,
                   const double *__restrict__ wgt, const int *__restrict__ slot, int nt, int NB, int AB,
                   double *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // gmem = 0: E1 and EB are copied to shared memory (n <= 76); gmem = 1: read in place from
  // global memory (L2-resident), only the AB + AB + 1 vectors live in shared memory
  const int tid = threadIdx.x;
  const size_t cfg = blockIdx.x;
  const size_t nn = (size_t)n * n;
  if (!GMEM) AB = 16;  // compile-time constant on the shared-memory path
  cplx *sm0 = reinterpret_cast<cplx *>(smem_raw);
  const int ldB = GMEM ? n : n + 1;                                // padded in shared memory
  const size_t npad = (size_t)n * (n + 1);
  const cplx *s1 = GMEM ? E1 + cfg * nn : sm0;                     // E1 [n][ldB]
  const cplx *sB = GMEM ? EB + cfg * nn : sm0;                     // EB [n][ldB] (same buffer, loaded after the o_b)
  const int ldO = n + 1;  // padded: in the P[a,b] loop the lanes of a warp read DIFFERENT rows of OB at the same j
  cplx *sO = GMEM ? sm0 : sm0 + npad;                              // OB [AB][ldO]
  cplx *sV = sO + (size_t)AB * ldO;                                // VA [AB][n]  (column a stored as a row)
  cplx *sv = sV + (size_t)AB * n;                                  // [n] scratch
  if (!GMEM) {
    for (int idx = tid; idx < nn; idx += 256) {
      const int r = idx / n, c = idx - r * n;
      sm0[r * (n + 1) + c] = E1[cfg * nn + idx];
    }
  }
  for (int i = tid; i < n; i += 256) {
    sO[i] = ov[cfg * n + i];
    sv[i] = r0[cfg * n + i];
  }
  __syncthreads();
  if (E0) {  // v0 = E0 r0 (t0 != 0)
    for (int i = tid; i < n; i += 256) {
      cplx acc = make_c(0, 0);
      for (int k = 0; k < n; ++k) cfma(acc, E0[cfg * nn + (size_t)i * n + k], sv[k]);
      sV[i] = acc;
    }
  } else {
    for (int i = tid; i < n; i += 256) sV[i] = sv[i];
  }
  __syncthreads();
  // o_b = o_{b-1} E1.  FOUR threads per output element (inner index q, q + 4, ...; combined with
  // two shuffles): with one thread per element only n of the 256 threads work and each runs a
  // chain of 4 n dependent FMAs (ncu: this kernel was 36 % of the C4 step).
  const int tq = tid & 3, tj = tid >> 2;
  for (int b = 1; b < NB; ++b) {
    const cplx *o = sO + (size_t)(b - 1) * ldO;
    for (int j0 = 0; j0 < n; j0 += 64) {
      const int j = j0 + tj;
      cplx acc = make_c(0, 0);
      if (j < n)
        for (int i = tq; i < n; i += 4) cfma(acc, o[i], s1[(size_t)i * ldB + j]);
      acc = cadd(acc, shfl_xor_c4(acc, 1));
      acc = cadd(acc, shfl_xor_c4(acc, 2));
      if (j < n && tq == 0) sO[(size_t)b * ldO + j] = acc;
    }
    __syncthreads();
  }
  if (!GMEM) {  // E1 is done: the buffer now takes EB
    for (int idx = tid; idx < nn; idx += 256) {
      const int r = idx / n, c = idx - r * n;
      sm0[r * (n + 1) + c] = EB[cfg * nn + idx];
    }
    __syncthreads();
  }
  const int NA_total = (nt + NB - 1) / NB;
  const double wc = wgt[cfg];
  const int sl = slot[cfg];
  for (int a0 = 0; a0 < NA_total; a0 += AB) {
    const int na = min(AB, NA_total - a0);
    // v_a = EB v_{a-1}: four threads per row i
    for (int a = (a0 == 0 ? 1 : 0); a < na; ++a) {
      const cplx *vp = (a == 0) ? sv : sV + (size_t)(a - 1) * n;
      for (int i0 = 0; i0 < n; i0 += 64) {
        const int i = i0 + tj;
        cplx acc = make_c(0, 0);
        if (i < n) {
          const cplx *row = sB + (size_t)i * ldB;
          for (int k = tq; k < n; k += 4) cfma(acc, row[k], vp[k]);
        }
        acc = cadd(acc, shfl_xor_c4(acc, 1));
        acc = cadd(acc, shfl_xor_c4(acc, 2));
        if (i < n && tq == 0) sV[(size_t)a * n + i] = acc;
      }
      __syncthreads();
    }
    // P[a,b]
    for (int e = tid; e < na * NB; e += 256) {
      const int a = e / NB, b = e - a * NB;
      const long k = (long)(a0 + a) * NB + b;
      if (k < nt) {
        double acc = 0.0;
        const cplx *o = sO + (size_t)b * ldO, *v = sV + (size_t)a * n;
        for (int j = 0; j < n; ++j) acc += o[j].x * v[j].x - o[j].y * v[j].y;
        atomicAdd(&out[(size_t)sl * nt + k], wc * acc);
      }
    }
    __syncthreads();
    // carry the last vector into the next block
    for (int i = tid; i < n; i += 256) sv[i] = sV[(size_t)(na - 1) * n + i];
    __syncthreads();
  }
}

// Arbitrary time arrays: one time point per launch, P = Re( o^T E v0 ) with E = exp(2 pi L t_k)
// already formed by the batched matrix exponential.  One CTA per configuration.
__global__ void __launch_bounds__(256)
lind_point_kernel(int n, const cplx *__restrict__ E, const cplx *__restrict__ r0, const cplx *__restrict__ ov,
                  const double *__restrict__ wgt, const int *__restrict__ slot, int nt, int k,
                  double *__restrict__ out) {
  __shared__ double red[34];
  const size_t cfg = blockIdx.x;
  const cplx *Ec = E + cfg * (size_t)n * n;
  const cplx *v = r0 + cfg * n, *o = ov + cfg * n;
  double acc = 0.0;
  // sum_ij o_i E_ij v_j: thread per (i, j) element, coalesced over j
  for (size_t idx = threadIdx.x; idx < (size_t)n * n; idx += 256) {
    const int i = (int)(idx / n), j = (int)(idx - (size_t)i * n);
    const cplx ev = cmul(Ec[idx], v[j]);
    acc += o[i].x * ev.x - o[i].y * ev.y;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(&out[(size_t)slot[cfg] * nt + k], wgt[cfg] * acc);
}

inline size_t lind_series_smem(int n, int AB = 16, bool gmem = false) {
  return ((gmem ? 0 : (size_t)n * (n + 1)) + (size_t)AB * (2 * n + 1) + n) * sizeof(cplx);
}

// Integral: solve (I/tau - 2 pi L) x = r0 by LU with partial pivoting; val = Re(o . x)/tau.
template <bool GMEM>
__global__ void __launch_bounds__(256)
lind_solve_kernel(int n, const cplx *__restrict__ L, const cplx *__restrict__ r0,
                  const cplx *__restrict__ ov, const double *__restrict__ wgt,
                  const int *__restrict__ slot, double tau, double *__restrict__ out,
                  int *__restrict__ status, cplx *gwork) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[64];
  __shared__ int ipiv[4];
  const int ld = n + 2;  // augmented with the rhs column, padded
  const int tid = threadIdx.x;
  const size_t cfg = blockIdx.x;
  // [n][ld] in shared memory, or (n > 76) in a global workspace: block-scope barriers order
  // the accesses of one CTA in either memory
  cplx *sM = GMEM ? gwork + cfg * (size_t)n * ld : reinterpret_cast<cplx *>(smem_raw);
  const size_t nn = (size_t)n * n;
  const double twopi = 6.283185307179586476925286766559, it = 1.0 / tau;
  for (int idx = tid; idx < nn; idx += 256) {
    const int r = idx / n, c = idx - r * n;
    cplx v = cscale(-twopi, L[cfg * nn + idx]);
    if (r == c) v.x += it;
    sM[r * ld + c] = v;
  }
  for (int i = tid; i < n; i += 256) sM[i * ld + n] = r0[cfg * n + i];
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    // pivot search in column k
    double best = -1.0;
    int bi = k;
    for (int i = k + tid; i < n; i += 256) {
      const double v = cnorm2(sM[i * ld + k]);
      if (v > best) {
        best = v;
        bi = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    if ((tid & 31) == 0) {
      red[tid >> 5] = best;
      red[32 + (tid >> 5)] = (double)bi;
    }
    __syncthreads();
    if (tid == 0) {
      double b = red[0];
      int idx = (int)red[32];
      for (int w = 1; w < 8; ++w)
        if (red[w] > b || (red[w] == b && (int)red[32 + w] < idx)) {
          b = red[w];
          idx = (int)red[32 + w];
        }
      ipiv[0] = idx;
      if (!(b > 0.0)) atomicMax(status, 2);  // singular
    }
    __syncthreads();
    const int pr = ipiv[0];
    if (pr != k) {
      for (int c = k + tid; c <= n; c += 256) {
        const cplx t = sM[k * ld + c];
        sM[k * ld + c] = sM[pr * ld + c];
        sM[pr * ld + c] = t;
      }
    }
    __syncthreads();
    const cplx piv = sM[k * ld + k];
    const double pd = 1.0 / cnorm2(piv);
    const cplx pinv = make_c(piv.x * pd, -piv.y * pd);
    // eliminate: rows i > k, columns c > k (including the rhs)
    const int rows = n - k - 1, cols = n - k;  // columns k+1..n
    for (int e = tid; e < rows * cols; e += 256) {
      const int i = k + 1 + e / cols, c = k + 1 + e % cols;
      const cplx f = cmul(sM[i * ld + k], pinv);
      cplx v = sM[i * ld + c];
      const cplx u = sM[k * ld + c];
      v.x -= f.x * u.x - f.y * u.y;
      v.y -= f.x * u.y + f.y * u.x;
      sM[i * ld + c] = v;
    }
    __syncthreads();
  }
  // back substitution (serial over rows, parallel over the dot product is not worth it: n <= 76)
  if (tid < 32) {
    for (int i = n - 1; i >= 0; --i) {
      cplx acc = make_c(0, 0);
      for (int c = i + 1 + tid; c < n; c += 32) cfma(acc, sM[i * ld + c], sM[c * ld + n]);
      acc.x = warp_sum(acc.x);
      acc.y = warp_sum(acc.y);
      if (tid == 0) {
        const cplx rhs = csub(sM[i * ld + n], acc);
        const cplx piv = sM[i * ld + i];
        const double pd = 1.0 / cnorm2(piv);
        sM[i * ld + n] = cmul(rhs, make_c(piv.x * pd, -piv.y * pd));
      }
      __syncwarp();
    }
    double acc = 0.0;
    for (int j = tid; j < n; j += 32) {
      const cplx o = ov[cfg * n + j], x = sM[j * ld + n];
      acc += o.x * x.x - o.y * x.y;
    }
    acc = warp_sum(acc);
    if (tid == 0) atomicAdd(&out[slot[cfg]], wgt[cfg] * acc * it);
  }
}

inline size_t lind_solve_smem(int n) { return (size_t)n * (n + 2) * sizeof(cplx) + 16; }

inline size_t lind_build_smem(int d, int nops) {
  return ((size_t)d * d * (1 + 2 * nops)) * sizeof(cplx) + (LIND_MAX_OPS + 34) * sizeof(double);
}

// ---------------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------------
struct LindWs {
  int64_t cap = 0;
  int n = 0;
  cplx *L = nullptr, *X = nullptr, *X2 = nullptr, *X3 = nullptr, *X4 = nullptr, *Ra = nullptr, *Rb = nullptr,
       *E1 = nullptr, *E0 = nullptr, *EBs = nullptr, *r0 = nullptr, *ov = nullptr, *gwork = nullptr;
  unsigned long long *norm = nullptr;
  int64_t gemm_cfgs = 0;  // batched n x n complex GEMMs executed, counted per configuration (bench.py: executed-flop roofline)
  void release() {
    cplx **ps[] = {&L, &X, &X2, &X3, &X4, &Ra, &Rb, &E1, &E0, &EBs, &r0, &ov, &gwork};
    for (auto p : ps) {
      dev_free(*p);
      *p = nullptr;
    }
    dev_free(norm);
    norm = nullptr;
    cap = 0;
  }
  cudaError_t ensure(int n_, int64_t cnt, bool series) {
    const bool need_g = !series && lind_solve_smem(n_) > 200 * 1024;
    if (cap >= cnt && n == n_ && (!series || X) && (!need_g || gwork)) return cudaSuccess;
    release();
    n = n_;
    const size_t nn = (size_t)n * n;
    cudaError_t e = cudaSuccess;
    auto al = [&](cplx **p, size_t c) {
      if (e == cudaSuccess) e = dev_malloc((void **)p, c * sizeof(cplx));
    };
    al(&L, cnt * nn);
    al(&r0, (size_t)cnt * n);
    al(&ov, (size_t)cnt * n);
    if (!series && lind_solve_smem(n) > 200 * 1024) al(&gwork, (size_t)cnt * n * (n + 2));
    if (series) {
      al(&X, cnt * nn);
      al(&X2, cnt * nn);
      al(&X3, cnt * nn);
      al(&X4, cnt * nn);
      al(&Ra, cnt * nn);
      al(&Rb, cnt * nn);
      al(&E1, cnt * nn);
      al(&E0, cnt * nn);
      al(&EBs, cnt * nn);
    }
    if (e == cudaSuccess) e = dev_malloc((void **)&norm, sizeof(unsigned long long));
    if (e == cudaSuccess) cap = cnt;
    return e;
  }
};

template <int EPI>
inline void lind_gemm(int n, int64_t cnt, const cplx *A, const cplx *B, const cplx *D, cplx *C, cudaStream_t st,
                      int64_t *launches) {
  const size_t nn = (size_t)n * n;
  ++*launches;
  MuonObs none = {1, 0};
  if (launch_zgemm_dmma<false, EPI, false>(n, cnt, A, nn, B, nn, C, 1.0, D, none, nullptr, st)) return;
  dim3 grid((n + 31) / 32, (n + 31) / 32, (unsigned)cnt);
  cgemm_batched_kernel<false, EPI><<<grid, 256, 0, st>>>(n, A, nn, B, nn, C, 1.0, D);
}

// exp(scale * L) for a batch: result pointer returned (one of the workspace buffers, not X..X4).
// `norm_inf` is the largest inf-norm of scale*L over the batch.
inline cplx *lind_expm(LindWs &ws, int n, int64_t cnt, double scale, double norm_inf, cplx *dst, cudaStream_t st,
                       int64_t *launches) {
  const size_t tot = (size_t)cnt * n * n;
  const unsigned eb = (unsigned)((tot + 255) / 256);
  int s = 0;
  const double theta = 0.35;
  while (norm_inf > theta && s < 60) {
    norm_inf *= 0.5;
    ++s;
  }
  lind_scale_kernel<<<eb, 256, 0, st>>>(tot, ldexp(scale, -s), ws.L, ws.X);
  ++*launches;
  lind_gemm<0>(n, cnt, ws.X, ws.X, nullptr, ws.X2, st, launches);
  lind_gemm<0>(n, cnt, ws.X2, ws.X, nullptr, ws.X3, st, launches);
  lind_gemm<0>(n, cnt, ws.X2, ws.X2, nullptr, ws.X4, st, launches);
  double c[13];
  c[0] = 1.0;
  for (int j = 1; j <= 12; ++j) c[j] = c[j - 1] / j;
  // p(X) = B0 + X4 (B1 + X4 (B2 + c12 X4)),  B_k = c_{4k} + c_{4k+1} X + c_{4k+2} X2 + c_{4k+3} X3
  lind_poly_kernel<<<eb, 256, 0, st>>>(n, tot, c[8], c[9], c[10], c[11], c[12], ws.X, ws.X2, ws.X3, ws.X4, ws.Ra);
  lind_poly_kernel<<<eb, 256, 0, st>>>(n, tot, c[4], c[5], c[6], c[7], 0.0, ws.X, ws.X2, ws.X3, nullptr, ws.Rb);
  *launches += 2;
  lind_gemm<2>(n, cnt, ws.X4, ws.Ra, ws.Rb, ws.Rb, st, launches);  // Rb = B1 + X4 Ra   (D == C is safe: same element)
  lind_poly_kernel<<<eb, 256, 0, st>>>(n, tot, c[0], c[1], c[2], c[3], 0.0, ws.X, ws.X2, ws.X3, nullptr, ws.Ra);
  ++*launches;
  cplx *cur = (s == 0) ? dst : ws.X2;  // X2 is free after the polynomials
  lind_gemm<2>(n, cnt, ws.X4, ws.Rb, ws.Ra, cur, st, launches);  // cur = B0 + X4 Rb
  cplx *other = ws.X3;
  for (int i = 0; i < s; ++i) {
    cplx *out = (i == s - 1) ? dst : other;
    lind_gemm<0>(n, cnt, cur, cur, nullptr, out, st, launches);
    other = cur;
    cur = out;
  }
  ws.gemm_cfgs += (int64_t)(5 + s) * cnt;
  return dst;
}

struct LindCtx {
  LindParams P;
  const cplx *H0, *Z, *M, *rho0_explicit, *exA;
  const double *exg;
};

inline int lindblad_run(const LindCtx &ctx, bool integral, int64_t n_cfg, const double *B, const double *p,
                        const double *T, const double *w, const int32_t *slot, int nt, const double *times_host,
                        bool uniform, double t0, double dt, double tau, double *out, LindWs &ws, long chunk_opt, int *status,
                        cudaStream_t st, int64_t *launches, Profiler *prof, std::string &err) {
  const int d = ctx.P.d, n = d * d;
  const int nops = 2 * ctx.P.n_diss + ctx.P.n_explicit;
  if (nops > LIND_MAX_OPS) {
    err = "too many dissipation operators";
    return -5;
  }
  const size_t bsmem = lind_build_smem(d, nops);
  if (bsmem > 200 * 1024 || n > 1024) {
    err = "Lindbladian path supports d <= 32 (and fewer dissipators at d = 32: operators are built in shared memory)";
    return -5;
  }
  if (!integral && !uniform && !times_host) {
    err = "Lindbladian evolution on a non-uniform time grid needs the host time array";
    return -1;
  }
  // n <= 76: super-operator resident in shared memory; above that the series / solve kernels read
  // the matrices in place from global memory and keep only AB + AB + 1 vectors on chip
  const bool gmem = lind_series_smem(n) > 227 * 1024 || lind_solve_smem(n) > 227 * 1024;
  int AB = gmem ? 32 : 16;
  while (gmem && AB > 2 && lind_series_smem(n, AB, true) > 200 * 1024) AB >>= 1;
  const size_t nn = (size_t)n * n;
  const int nbuf = integral ? 2 : 11;
  int64_t chunk = chunk_opt > 0 ? chunk_opt : std::max<int64_t>(1, (int64_t)(2.0e9 / (nbuf * nn * sizeof(cplx))));
  chunk = std::min(chunk, n_cfg);
  cudaError_t e = ws.ensure(n, chunk, !integral);
  if (e != cudaSuccess) {
    err = std::string("Lindblad workspace: ") + cudaGetErrorString(e);
    return -2;
  }
  cudaFuncSetAttribute(lind_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem);
  const size_t ssmem = lind_series_smem(n, AB, gmem);
  const size_t vsmem = gmem ? 16 : lind_solve_smem(n);
  cudaFuncSetAttribute(lind_series_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssmem);
  cudaFuncSetAttribute(lind_series_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssmem);
  cudaFuncSetAttribute(lind_solve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vsmem);
  int NB = 1;
  while (NB * NB < nt && NB < AB) NB <<= 1;
  const double twopi = 6.283185307179586476925286766559;
  for (int64_t c0 = 0; c0 < n_cfg; c0 += chunk) {
    const int64_t cnt = std::min(chunk, n_cfg - c0);
    ProfScope ps(prof, st, PH_LINDBLAD);
    cudaMemsetAsync(ws.norm, 0, sizeof(unsigned long long), st);
    lind_build_kernel<<<(unsigned)cnt, 256, bsmem, st>>>(ctx.P, ctx.H0, ctx.Z, ctx.M, ctx.rho0_explicit, ctx.exA,
                                                        ctx.exg, B + 3 * c0, p + 3 * c0, T ? T + c0 : nullptr, ws.L,
                                                        ws.r0, ws.ov, ws.norm);
    ++*launches;
    if (integral) {
      if (gmem)
        lind_solve_kernel<true><<<(unsigned)cnt, 256, vsmem, st>>>(n, ws.L, ws.r0, ws.ov, w + c0, slot + c0, tau, out, status,
                                                                   ws.gwork);
      else
        lind_solve_kernel<false><<<(unsigned)cnt, 256, vsmem, st>>>(n, ws.L, ws.r0, ws.ov, w + c0, slot + c0, tau, out, status,
                                                                    nullptr);
      ++*launches;
    } else {
      unsigned long long bits = 0;
      e = cudaMemcpyAsync(&bits, ws.norm, sizeof bits, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) {
        err = std::string("Lindblad norm: ") + cudaGetErrorString(e);
        return -2;
      }
      double norm;
      memcpy(&norm, &bits, sizeof norm);
      if (!uniform) {
        // arbitrary time rows (the reference evaluates exp(mu_k t) for any t, lindbladian.py:103-108):
        // one matrix exponential per time point, O(nt n^3) -- correct, not fast
        for (int k = 0; k < nt; ++k) {
          const double tk = times_host[k];
          lind_expm(ws, n, cnt, twopi * tk, norm * twopi * fabs(tk), ws.E1, st, launches);
          lind_point_kernel<<<(unsigned)cnt, 256, 0, st>>>(n, ws.E1, ws.r0, ws.ov, w + c0, slot + c0, nt, k, out);
          ++*launches;
        }
        e = cudaGetLastError();
        if (e != cudaSuccess) {
          err = std::string("Lindblad launch: ") + cudaGetErrorString(e);
          return -2;
        }
        continue;
      }
      // E1 = exp(2 pi dt L); EB = E1^NB; E0 = exp(2 pi t0 L) if t0 != 0
      lind_expm(ws, n, cnt, twopi * dt, norm * twopi * fabs(dt), ws.E1, st, launches);
      cplx *cur = ws.E1, *EB = ws.E1;
      cplx *pp[2] = {ws.Ra, ws.Rb};
      int flip = 0;
      for (int q = 1; q < NB; q <<= 1) {
        lind_gemm<0>(n, cnt, cur, cur, nullptr, pp[flip], st, launches);
        ws.gemm_cfgs += cnt;
        cur = pp[flip];
        flip ^= 1;
      }
      EB = cur;
      cplx *E0 = nullptr;
      if (t0 != 0.0) {
        // lind_expm uses X..X4, Ra, Rb as scratch: move EB to its own buffer first
        if (EB != ws.E1) {
          cudaMemcpyAsync(ws.EBs, EB, (size_t)cnt * nn * sizeof(cplx), cudaMemcpyDeviceToDevice, st);
          EB = ws.EBs;
        }
        E0 = lind_expm(ws, n, cnt, twopi * t0, norm * twopi * fabs(t0), ws.E0, st, launches);
      }
      if (gmem)
        lind_series_kernel<true><<<(unsigned)cnt, 256, ssmem, st>>>(n, ws.E1, EB, E0, ws.r0, ws.ov, w + c0, slot + c0, nt, NB,
                                                                    AB, out);
      else
        lind_series_kernel<false><<<(unsigned)cnt, 256, ssmem, st>>>(n, ws.E1, EB, E0, ws.r0, ws.ov, w + c0, slot + c0, nt, NB,
                                                                     AB, out);
      ++*launches;
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) {
      err = std::string("Lindblad launch: ") + cudaGetErrorString(e);
      return -2;
    }
  }
  return 0;
}

}  // namespace musim
