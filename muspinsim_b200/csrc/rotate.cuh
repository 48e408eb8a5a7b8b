// rotate.cuh -- per-configuration operators and basis rotations.
//
//  * form_obs_kernel      O_c = sum_a p_a M_a            (MuonSpinSystem.muon_operator,
//                                                        /root/reference/muspinsim/spinsys.py:707-732)
//  * rho0_kernel          thermal product state          (ExperimentRunner.rho0, experiment.py:170-236)
//  * cgemm_batched_kernel C = op(A) B, complex FP64      (Operator.basis_change, spinop.py:330-355)
//  * weights_kernel       w_ij = rho'_ij O'_ji           (hamiltonian.py:77-103; fast path :204-217)
//  * integral_kernel      sum_ab w_ab/(1/tau+2 pi i(l_a-l_b))/tau   (hamiltonian.py:150-162,
//                                                        experiment.py:492-496)
#pragma once
#include "common.cuh"

namespace musim {

#define MUSIM_MAX_SPINS 16
#define MUSIM_MAX_SDIM 10  // largest single-spin dimension 2I+1 (I <= 9/2)

struct SpinTable {
  int n_spins;
  int muon_index;
  int dims[MUSIM_MAX_SPINS];
  double gammas[MUSIM_MAX_SPINS];
};

// O_c[idx] = px M0 + py M1 + pz M2 ; grid (ceil(d*d/256), n_cfg)
__global__ void form_obs_kernel(int d, const cplx *__restrict__ M, const double *__restrict__ pf,
                                cplx *__restrict__ O) {
  const size_t cfg = blockIdx.y;
  const size_t dd = (size_t)d * d;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= dd) return;
  const double px = pf[cfg * 3 + 0], py = pf[cfg * 3 + 1], pz = pf[cfg * 3 + 2];
  const cplx m0 = M[idx], m1 = M[dd + idx], m2 = M[2 * dd + idx];
  O[cfg * dd + idx] = make_c(px * m0.x + py * m1.x + pz * m2.x, px * m0.y + py * m1.y + pz * m2.y);
}

// Spin matrices for spin I (n = 2I+1), m from +I to -I  (spinop.py:14-41):
// h = nx Sx + ny Sy + nz Sz  as an n x n Hermitian matrix.
__device__ inline void spin_dot(int n, double nx, double ny, double nz, cplx *h /* n*n */) {
  const double I = 0.5 * (n - 1);
  for (int i = 0; i < n * n; ++i) h[i] = make_c(0.0, 0.0);
  for (int a = 0; a < n; ++a) {
    const double m = I - a;
    h[a * n + a] = make_c(nz * m, 0.0);
    if (a + 1 < n) {
      // <m|S+|m-1> = sqrt(I(I+1) - m(m-1)) ; element (a, a+1) of S+
      const double mp = I - (a + 1);
      const double sp = sqrt(I * (I + 1.0) - mp * (mp + 1.0));
      // Sx = (S+ + S-)/2, Sy = (S+ - S-)/(2i):  (a,a+1): nx*sp/2 - i*ny*sp/2
      h[a * n + a + 1] = make_c(0.5 * nx * sp, -0.5 * ny * sp);
      h[(a + 1) * n + a] = make_c(0.5 * nx * sp, 0.5 * ny * sp);
    }
  }
}

// Thermal single-spin density matrix  rho = sum_m p_m P_m(n.S),  p_m ~ exp(-x m)
// (experiment.py:204-228: eigh of B.S, E = eval*1e6*gamma, Boltzmann weights exp(-hE/kT)).
// P_m are Lagrange projectors  prod_{m' != m} (h - m')/(m - m').
__device__ inline void thermal_factor(int n, double gamma, double bx, double by, double bz,
                                      double T, cplx *rho /* n*n */) {
  const double I = 0.5 * (n - 1);
  const double Bn = sqrt(bx * bx + by * by + bz * bz);
  double pm[MUSIM_MAX_SDIM];
  bool uniform = false;
  const double kB = 1.380649e-23, hP = 6.62607015e-34;
  const double escale = gamma * Bn * 1e6;  // E_m = m * escale  (Hz)
  if (Bn == 0.0 || escale == 0.0) {
    uniform = true;
  } else if (T > 0.0) {
    if (isinf(T)) {
      uniform = true;
    } else {
      const double x = hP * escale / (kB * T);
      // subtract the maximum exponent for robustness (the reference does not; same result
      // wherever the reference does not overflow)
      double mx = -1e300;
      for (int a = 0; a < n; ++a) mx = fmax(mx, -x * (I - a));
      double s = 0.0;
      for (int a = 0; a < n; ++a) {
        pm[a] = exp(-x * (I - a) - mx);
        s += pm[a];
      }
      for (int a = 0; a < n; ++a) pm[a] /= s;
    }
  } else {  // T == 0: ground state only (experiment.py:218-219)
    for (int a = 0; a < n; ++a) pm[a] = 0.0;
    if (escale > 0.0)
      pm[n - 1] = 1.0;  // m = -I has the lowest E
    else
      pm[0] = 1.0;
  }
  if (uniform) {
    for (int i = 0; i < n * n; ++i) rho[i] = make_c(0.0, 0.0);
    for (int a = 0; a < n; ++a) rho[a * n + a] = make_c(1.0 / n, 0.0);
    return;
  }
  if (n == 2) {  // spin 1/2: P_+- = 1/2 +- n.S, so rho = 1/2 + (p_+ - p_-) n.S
    spin_dot(2, bx / Bn, by / Bn, bz / Bn, rho);
    const double dp = pm[0] - pm[1];
    for (int i = 0; i < 4; ++i) rho[i] = cscale(dp, rho[i]);
    rho[0].x += 0.5;
    rho[3].x += 0.5;
    return;
  }
  cplx h[MUSIM_MAX_SDIM * MUSIM_MAX_SDIM], P[MUSIM_MAX_SDIM * MUSIM_MAX_SDIM],
      Q[MUSIM_MAX_SDIM * MUSIM_MAX_SDIM];
  spin_dot(n, bx / Bn, by / Bn, bz / Bn, h);
  for (int i = 0; i < n * n; ++i) rho[i] = make_c(0.0, 0.0);
  for (int a = 0; a < n; ++a) {
    if (pm[a] == 0.0) continue;
    const double m = I - a;
    for (int i = 0; i < n * n; ++i) P[i] = make_c(0.0, 0.0);
    for (int i = 0; i < n; ++i) P[i * n + i] = make_c(1.0, 0.0);
    for (int b = 0; b < n; ++b) {
      if (b == a) continue;
      const double mb = I - b;
      const double inv = 1.0 / (m - mb);
      // Q = P * (h - mb) * inv
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
          cplx acc = make_c(0.0, 0.0);
          for (int k = 0; k < n; ++k) {
            cplx hk = h[k * n + j];
            if (k == j) hk.x -= mb;
            cfma(acc, P[i * n + k], hk);
          }
          Q[i * n + j] = cscale(inv, acc);
        }
      for (int i = 0; i < n * n; ++i) P[i] = Q[i];
    }
    for (int i = 0; i < n * n; ++i) {
      rho[i].x += pm[a] * P[i].x;
      rho[i].y += pm[a] * P[i].y;
    }
  }
}

// One CTA per configuration: factors in shared memory, then the Kronecker product.
__global__ void rho0_kernel(int d, SpinTable tab, const double *__restrict__ Bf,
                            const double *__restrict__ pf, const double *__restrict__ Tf,
                            cplx *__restrict__ R) {
  __shared__ cplx fac[MUSIM_MAX_SPINS][MUSIM_MAX_SDIM * MUSIM_MAX_SDIM];
  const size_t cfg = blockIdx.x;
  const double bx = Bf[cfg * 3], by = Bf[cfg * 3 + 1], bz = Bf[cfg * 3 + 2];
  const double T = Tf[cfg];
  if (threadIdx.x < tab.n_spins) {
    const int s = threadIdx.x;
    const int n = tab.dims[s];
    if (s == tab.muon_index) {
      // pure state along p:  1/2 + p_hat . S   (DensityOperator.from_vectors, spinop.py:455-530)
      double px = pf[cfg * 3], py = pf[cfg * 3 + 1], pz = pf[cfg * 3 + 2];
      const double pn = sqrt(px * px + py * py + pz * pz);
      if (pn > 0.0) {
        px /= pn;
        py /= pn;
        pz /= pn;
      }
      spin_dot(2, px, py, pz, fac[s]);
      fac[s][0].x += 0.5;
      fac[s][3].x += 0.5;
    } else {
      thermal_factor(n, tab.gammas[s], bx, by, bz, T, fac[s]);
    }
  }
  __syncthreads();
  const size_t dd = (size_t)d * d;
  for (int idx = threadIdx.x; idx < dd; idx += blockDim.x) {
    int i = idx / d, j = idx - (idx / d) * d;
    cplx v = make_c(1.0, 0.0);
    for (int s = tab.n_spins - 1; s >= 0; --s) {
      const int n = tab.dims[s];
      const int is = i % n, js = j % n;
      i /= n;
      j /= n;
      v = cmul(v, fac[s][is * n + js]);
    }
    R[cfg * dd + idx] = v;
  }
}

// X <- (1 (x) f (x) 1) X for one single-spin factor f (n x n) on a tile of 32 columns: lane = column,
// the warps stride over the d / n row groups {hi*n*stride + a*stride + lo, a = 0 .. n-1}.
// N > 0: compile-time spin dimension (factor and inputs in registers), N = 0: run-time n.
template <int N>
__device__ __forceinline__ void kron_factor_apply(int d, int stride, const cplx *f, cplx *X, int n_rt = 0) {
  const int n = N > 0 ? N : n_rt;
  const int groups = d / n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  cplx fr[N > 0 ? N * N : 1];
  if (N > 0) {
#pragma unroll
    for (int i = 0; i < N * N; ++i) fr[i] = f[i];
  }
  for (int gi = warp; gi < groups; gi += nw) {
    const int hi = gi / stride, lo = gi - hi * stride;
    cplx *x0 = X + ((size_t)hi * n * stride + lo) * 32 + lane;
    const size_t rs = (size_t)stride * 32;
    if (N > 0) {
      cplx xin[N > 0 ? N : 1];
#pragma unroll
      for (int b = 0; b < N; ++b) xin[b] = x0[b * rs];
#pragma unroll
      for (int a = 0; a < N; ++a) {
        cplx acc = make_c(0.0, 0.0);
#pragma unroll
        for (int b = 0; b < N; ++b) cfma(acc, fr[a * N + b], xin[b]);
        x0[a * rs] = acc;
      }
    } else {
      cplx xin[MUSIM_MAX_SDIM];
      for (int b = 0; b < n; ++b) xin[b] = x0[b * rs];
      for (int a = 0; a < n; ++a) {
        cplx acc = make_c(0.0, 0.0);
        for (int b = 0; b < n; ++b) cfma(acc, f[a * n + b], xin[b]);
        x0[a * rs] = acc;
      }
    }
  }
}

// T1 = rho0 U without forming rho0: rho0 = (x)_s f_s is a Kronecker product of single-spin
// density matrices (experiment.py:170-236, spinop.py:266-295), so it is applied factor by factor,
//   X <- (1 (x) f_s (x) 1) X    for s = 0 .. n_spins-1,
// at d^2 * sum_s n_s complex FMAs instead of the d^3 of the dense product (7x fewer at d = 96),
// and the d x d matrix rho0 never exists in memory.  One CTA per configuration; U is processed
// in tiles of 32 columns held in shared memory (lane = column).
// Single-spin factors of the thermal product state, one THREAD per (configuration, spin): the
// Lagrange-projector arithmetic is serial per spin and would otherwise sit in front of every
// CTA of rho0_apply_kernel.  F[cfg][off_s .. off_s + n_s^2), off_s = sum_{q<s} n_q^2  (<= d^2 entries).
__global__ void rho0_factors_kernel(int64_t n_cfg, size_t fstride, SpinTable tab, const double *__restrict__ Bf,
                                    const double *__restrict__ pf, const double *__restrict__ Tf,
                                    cplx *__restrict__ F) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_cfg * tab.n_spins) return;
  const int64_t cfg = t / tab.n_spins;
  const int s = (int)(t - cfg * tab.n_spins);
  int off = 0;
  for (int q = 0; q < s; ++q) off += tab.dims[q] * tab.dims[q];
  const int n = tab.dims[s];
  cplx fac[MUSIM_MAX_SDIM * MUSIM_MAX_SDIM];
  if (s == tab.muon_index) {
    // pure state along p:  1/2 + p_hat . S   (DensityOperator.from_vectors, spinop.py:455-530)
    double px = pf[cfg * 3], py = pf[cfg * 3 + 1], pz = pf[cfg * 3 + 2];
    const double pn = sqrt(px * px + py * py + pz * pz);
    if (pn > 0.0) {
      px /= pn;
      py /= pn;
      pz /= pn;
    }
    spin_dot(2, px, py, pz, fac);
    fac[0].x += 0.5;
    fac[3].x += 0.5;
  } else {
    thermal_factor(n, tab.gammas[s], Bf[cfg * 3], Bf[cfg * 3 + 1], Bf[cfg * 3 + 2], Tf[cfg], fac);
  }
  for (int i = 0; i < n * n; ++i) F[cfg * fstride + off + i] = fac[i];
}

__global__ void __launch_bounds__(256, 4)
rho0_apply_kernel(int d, SpinTable tab, const cplx *__restrict__ F, size_t fstride,
                  const cplx *__restrict__ U, cplx *__restrict__ T1) {
  extern __shared__ __align__(16) unsigned char rho_smem[];
  cplx *X = reinterpret_cast<cplx *>(rho_smem);  // [d][32]
  cplx *facs = X + (size_t)d * 32;                // single-spin factors, packed (sum_s n_s^2 entries)
  const size_t cfg = blockIdx.x;
  int nfac = 0;
  for (int s = 0; s < tab.n_spins; ++s) nfac += tab.dims[s] * tab.dims[s];
  for (int i = threadIdx.x; i < nfac; i += blockDim.x) facs[i] = F[cfg * fstride + i];
  __syncthreads();
  const size_t dd = (size_t)d * d;
  const cplx *Uc = U + cfg * dd;
  cplx *Tc = T1 + cfg * dd;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c0 = 0; c0 < d; c0 += 32) {
    const bool col_ok = c0 + lane < d;
    for (int i = warp; i < d; i += nw) X[(size_t)i * 32 + lane] = col_ok ? Uc[(size_t)i * d + c0 + lane] : make_c(0.0, 0.0);
    __syncthreads();
    int stride = d, foff = 0;
    for (int s = 0; s < tab.n_spins; ++s) {
      const int n = tab.dims[s];
      stride /= n;  // product of the dimensions after spin s
      const cplx *f = facs + foff;
      foff += n * n;
      // a real multiple of the identity (T = inf, B = 0) only scales
      bool ident = f[0].y == 0.0;
      for (int a = 0; a < n && ident; ++a)
        for (int b = 0; b < n; ++b)
          if (a != b ? (f[a * n + b].x != 0.0 || f[a * n + b].y != 0.0) : (f[a * n + a].x != f[0].x || f[a * n + a].y != 0.0))
            ident = false;
      if (ident) {
        const double sc = f[0].x;
        if (sc != 1.0)
          for (int i = warp; i < d; i += nw) X[(size_t)i * 32 + lane] = cscale(sc, X[(size_t)i * 32 + lane]);
      } else if (n == 2) {
        kron_factor_apply<2>(d, stride, f, X);
      } else if (n == 3) {
        kron_factor_apply<3>(d, stride, f, X);
      } else if (n == 4) {
        kron_factor_apply<4>(d, stride, f, X);
      } else {
        kron_factor_apply<0>(d, stride, f, X, n);
      }
      __syncthreads();
    }
    if (col_ok)
      for (int i = warp; i < d; i += nw) Tc[(size_t)i * d + c0 + lane] = X[(size_t)i * 32 + lane];
    __syncthreads();
  }
}

inline size_t rho0_apply_smem(int d, const SpinTable &tab) {  // one tile of 32 columns + the packed factors
  size_t nfac = 0;
  for (int s = 0; s < tab.n_spins; ++s) nfac += (size_t)tab.dims[s] * tab.dims[s];
  return ((size_t)d * 32 + nfac) * sizeof(cplx);
}

// ---------------------------------------------------------------------------------------
// Batched complex GEMM, C[c] = op(A[c]) * B[c] (all d x d, row-major).  CONJ_A: op = A^H.
// 32x32 output tile per CTA (256 threads, 2x2 complex outputs each), K staged through
// shared memory in slabs of 16.  a_stride / b_stride may be 0 (operand shared by the batch).
// EPI: 0 store C; 1 store (|C|^2 * scale, 0)   [fast-path weights |O'|^2 / d_o];
//      2 store C + D (D may alias C)           [Horner steps of the matrix exponential]
// ---------------------------------------------------------------------------------------
template <bool CONJ_A, int EPI>
__global__ void __launch_bounds__(256)
cgemm_batched_kernel(int d, const cplx *__restrict__ A, size_t a_stride,
                     const cplx *__restrict__ B, size_t b_stride, cplx *C, double scale,
                     const cplx *D) {
  constexpr int TM = 32, TN = 32, TK = 16;
  __shared__ cplx sA[TK][TM + 1];  // sA[k][m] = op(A)(m0+m, k0+k)
  __shared__ cplx sB[TK][TN + 1];  // sB[k][n] = B(k0+k, n0+n)
  const size_t cfg = blockIdx.z;
  const size_t dd = (size_t)d * d;
  const cplx *Ab = A + cfg * a_stride;
  const cplx *Bb = B + cfg * b_stride;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads
  cplx acc[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[i][j] = make_c(0.0, 0.0);

  for (int k0 = 0; k0 < d; k0 += TK) {
    // stage A slab: TK x TM elements, 512 -> 2 per thread
    for (int e = threadIdx.x; e < TK * TM; e += 256) {
      int k, m;
      cplx v = make_c(0.0, 0.0);
      if (CONJ_A) {
        // op(A)(m,k) = conj(A(k,m)): read row k0+k of A, consecutive m -> coalesced
        k = e / TM;
        m = e - k * TM;
        if (k0 + k < d && m0 + m < d) v = cconj(Ab[(size_t)(k0 + k) * d + m0 + m]);
      } else {
        // op(A)(m,k) = A(m,k): read row m0+m, consecutive k -> coalesced
        m = e / TK;
        k = e - m * TK;
        if (k0 + k < d && m0 + m < d) v = Ab[(size_t)(m0 + m) * d + k0 + k];
      }
      sA[k][m] = v;
    }
    for (int e = threadIdx.x; e < TK * TN; e += 256) {
      const int k = e / TN, n = e - k * TN;
      cplx v = make_c(0.0, 0.0);
      if (k0 + k < d && n0 + n < d) v = Bb[(size_t)(k0 + k) * d + n0 + n];
      sB[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const cplx a0 = sA[k][ty], a1 = sA[k][ty + 16];
      const cplx b0 = sB[k][tx], b1 = sB[k][tx + 16];
      cfma(acc[0][0], a0, b0);
      cfma(acc[0][1], a0, b1);
      cfma(acc[1][0], a1, b0);
      cfma(acc[1][1], a1, b1);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int m = m0 + ty + 16 * i, n = n0 + tx + 16 * j;
      if (m < d && n < d) {
        cplx v = acc[i][j];
        if (EPI == 1) v = make_c(cnorm2(v) * scale, 0.0);
        if (EPI == 2) v = cadd(v, D[cfg * dd + (size_t)m * d + n]);
        C[cfg * dd + (size_t)m * d + n] = v;
      }
    }
}

// C[c] = A[c] (complex) * B[c] (REAL), all d x d row-major: the eigenvector back-transformation
// U = Q Zt of the Householder+QL solver (2 FMA per k instead of 4).
__global__ void __launch_bounds__(256)
cgemm_realB_kernel(int d, const cplx *__restrict__ A, const double *__restrict__ B,
                   cplx *__restrict__ C) {
  constexpr int TM = 32, TN = 32, TK = 16;
  __shared__ cplx sA[TK][TM + 1];
  __shared__ double sB[TK][TN + 1];
  const size_t cfg = blockIdx.z;
  const size_t dd = (size_t)d * d;
  const cplx *Ab = A + cfg * dd;
  const double *Bb = B + cfg * dd;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  cplx acc[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[i][j] = make_c(0.0, 0.0);
  for (int k0 = 0; k0 < d; k0 += TK) {
    for (int e = threadIdx.x; e < TK * TM; e += 256) {
      const int m = e / TK, k = e - m * TK;
      cplx v = make_c(0.0, 0.0);
      if (k0 + k < d && m0 + m < d) v = Ab[(size_t)(m0 + m) * d + k0 + k];
      sA[k][m] = v;
    }
    for (int e = threadIdx.x; e < TK * TN; e += 256) {
      const int k = e / TN, n = e - k * TN;
      double v = 0.0;
      if (k0 + k < d && n0 + n < d) v = Bb[(size_t)(k0 + k) * d + n0 + n];
      sB[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const cplx a0 = sA[k][ty], a1 = sA[k][ty + 16];
      const double b0 = sB[k][tx], b1 = sB[k][tx + 16];
      acc[0][0].x = fma(a0.x, b0, acc[0][0].x);
      acc[0][0].y = fma(a0.y, b0, acc[0][0].y);
      acc[0][1].x = fma(a0.x, b1, acc[0][1].x);
      acc[0][1].y = fma(a0.y, b1, acc[0][1].y);
      acc[1][0].x = fma(a1.x, b0, acc[1][0].x);
      acc[1][0].y = fma(a1.y, b0, acc[1][0].y);
      acc[1][1].x = fma(a1.x, b1, acc[1][1].x);
      acc[1][1].y = fma(a1.y, b1, acc[1][1].y);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int m = m0 + ty + 16 * i, n = n0 + tx + 16 * j;
      if (m < d && n < d) C[cfg * dd + (size_t)m * d + n] = acc[i][j];
    }
}

// W = X .* conj(Y) elementwise (general weights  w_ij = rho'_ij * O'_ji, O' Hermitian)
__global__ void weights_kernel(size_t n, const cplx *__restrict__ X, const cplx *__restrict__ Y,
                               cplx *__restrict__ W) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) W[i] = cmulc(X[i], Y[i]);
}

// Integral of the decaying signal, one CTA per configuration:
//   val = (1/tau) Re sum_ab w_ab / (1/tau + 2 pi i (l_a - l_b));   out[slot] += weight * val
__global__ void integral_kernel(int d, const cplx *__restrict__ W, const double *__restrict__ lam,
                                const double *__restrict__ wgt, const int *__restrict__ slot,
                                double tau, double *__restrict__ out) {
  __shared__ double red[34];
  const size_t cfg = blockIdx.x;
  const size_t dd = (size_t)d * d;
  const cplx *Wc = W + cfg * dd;
  const double *lc = lam + cfg * d;
  const double it = 1.0 / tau;
  const double twopi = 6.283185307179586476925286766559;
  double acc = 0.0;
  for (int idx = threadIdx.x; idx < dd; idx += blockDim.x) {
    const int a = idx / d, b = idx - a * d;
    const cplx w = Wc[idx];
    const double om = twopi * (lc[a] - lc[b]);
    // Re[ w / (it + i om) ] = (w.x*it + w.y*om) / (it^2 + om^2)
    acc += (w.x * it + w.y * om) / (it * it + om * om);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(&out[slot[cfg]], wgt[cfg] * acc * it);
}

// ---------------------------------------------------------------------------------------
// Density-matrix output of Hamiltonian.evolve(operators=None) (hamiltonian.py:86-115):
// X[t][i][j] = R0[i][j] exp(-2 pi i (l_i - l_j) t) in the eigenbasis, and the conjugate transpose
// of the eigenvector matrix (the right factor of U X U^H for the batched GEMM).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rho_phase_kernel(int d, int nt, const cplx *__restrict__ R0, const double *__restrict__ lam,
                 const double *__restrict__ times, cplx *__restrict__ X) {
  const size_t dd = (size_t)d * d;
  const size_t total = dd * (size_t)nt;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t t = e / dd, ij = e - t * dd;
    const int i = (int)(ij / d), j = (int)(ij - (size_t)i * d);
    const cplx ph = cis_m2pi((lam[i] - lam[j]) * times[t]);
    X[e] = cmul(R0[ij], ph);
  }
}
__global__ void __launch_bounds__(256)
conj_transpose_kernel(int d, const cplx *__restrict__ U, cplx *__restrict__ Ud) {
  const int n = d * d;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int i = e / d, j = e - i * d;
    Ud[(size_t)j * d + i] = cconj(U[e]);
  }
}

}  // namespace musim
