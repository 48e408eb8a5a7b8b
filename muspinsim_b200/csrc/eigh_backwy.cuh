// eigh_backwy.cuh -- K4 on the FP64 tensor pipe: back-transformation U = Q Zt of the Hermitian
// eigensolver (second half of np.linalg.eigh, /root/reference/muspinsim/spinop.py:69; LAPACK
// zunmtr) with the Householder reflectors applied in COMPACT-WY BLOCKS of 8,
//
//     Q = Q_0 Q_1 ... Q_{NB-1},   Q_b = H_{8b} ... H_{8b+7} = I - V_b T_b V_b^H      (zlarft, forward / columnwise)
//     C <- C - V_b (T_b (V_b^H C))                                                   (zlarfb), b = NB-1 .. 0, C = Zt at the start
//
// so that every O(d^3) flop is a DMMA (mma.sync m8n8k4 f64).  The level-2 kernel it replaces
// (hql_reflect_kernel: one reflector at a time, dot product + axpy per column) ran one dependency
// chain per SM at 45 % of the FP64 pipe (profiles/r1_ncu_full_summary.md).
//
// Mapping.  One CTA per matrix, D / 8 warps; warp w owns columns 8w .. 8w+7 of C for ALL rows and
// keeps them in registers as the accumulator fragments of C^T (tile t = rows 8t .. 8t+7):
//     accumulator (m = lane/4, n = 2 (lane%4) + e)   <->   C[row 8t + n][column 8w + m].
// With C held transposed, all three products of a block are warp-local AND shuffle-free, because an
// accumulator fragment read slot by slot IS an A-operand fragment whose reduction index is
// permuted (slot e covers n = e, 2 + e, 4 + e, 6 + e):
//   1. W^T  = C^T conj(V_b)      M = column, N = reflector, K = row        A = C^T accumulators
//   2. W2^T = W^T T_b^T          M = column, N = reflector, K = reflector  A = W^T accumulators
//   3. C^T -= W2^T V_b^T         M = column, N = row,       K = reflector  A = W2^T accumulators
// Only the B operands come from shared memory.  V is staged ONCE per matrix as planar re / im
// arrays Vs[reflector][row'] with leading dimension D + 4 (= 4 mod 16) and the rows of every group
// of 8 stored in the order pos = (0, 6, 1, 7, 2, 4, 3, 5): with that permutation the fragment loads
// of step 1 (rows 2j + e for j = lane%4) and of step 3 (rows lane/4, reflectors 4e + j) are both
// bank-conflict free.  The output columns of step 2 are assigned to reflectors 4e + j for the same
// reason (T is stored in the matching order).
//
// T_b: warp b forms the Gram matrix G = V_b^H V_b with DMMAs (A and B fragment are the same loaded
// value), then lane l < 8 runs row l of the zlarft recurrence
//     T[l][l] = tau_l,   T[l][i] = -tau_i sum_{q=l}^{i-1} T[l][q] G[q][i]    (rows are independent).
#pragma once
#include "common.cuh"
#include "polar.cuh"  // dmma884

namespace musim {

template <int D>
struct BackWyGeom {
  static constexpr int NB = D / 8;    // reflector blocks = row tiles = warps
  static constexpr int LD = D + 4;    // = 4 (mod 16) for D = 32, 64, 96
  static constexpr int TLD = 12;      // leading dimension of the 8 x 8 T operand tiles
  static constexpr size_t smem_bytes =
      (size_t)2 * D * LD * sizeof(double) + (size_t)2 * NB * 8 * TLD * sizeof(double) + (size_t)NB * 64 * sizeof(cplx) +
      (size_t)D * sizeof(cplx);
};

template <int D>
__global__ void __launch_bounds__(4 * D, 1)
hql_backwy_kernel(int d, const double *__restrict__ Zt, const cplx *__restrict__ Vp, size_t vcap,
                  const cplx *__restrict__ tau, cplx *__restrict__ U) {
  using G = BackWyGeom<D>;
  constexpr int NB = G::NB, LD = G::LD, TLD = G::TLD, NT = 4 * D;
  static_assert(D % 32 == 0 && D <= 96, "D = 32, 64 or 96");
  extern __shared__ __align__(16) unsigned char bw_smem[];
  double *Vre = reinterpret_cast<double *>(bw_smem);  // [D][LD]
  double *Vim = Vre + D * LD;
  double *Tre = Vim + D * LD;                          // [NB][8][TLD]
  double *Tim = Tre + NB * 8 * TLD;
  cplx *Gs = reinterpret_cast<cplx *>(Tim + NB * 8 * TLD);  // [NB][64]
  cplx *stau = Gs + NB * 64;                                // [D]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fm = lane >> 2, fj = lane & 3;
  const size_t mat = blockIdx.x;
  const size_t dd = (size_t)d * d;
  const cplx *myv = Vp + mat * vcap;

  // ---- stage V (unit diagonal and zeros made explicit) and tau ----
  for (int i = tid; i < D; i += NT) stau[i] = (i < d - 1) ? tau[mat * d + i] : make_c(0.0, 0.0);
  for (int idx = tid; idx < D * D; idx += NT) {
    const int i = idx / D, r = idx - i * D;  // reflector i, row r
    cplx v = make_c(0.0, 0.0);
    if (i < d - 1 && r < d) {
      if (r > i + 1) {
        const int mk = d - i - 2;
        v = myv[(size_t)mk * (mk - 1) / 2 + (r - i - 2)];
      } else if (r == i + 1) {
        v = make_c(1.0, 0.0);
      }
    }
    const int p = (0x53427160u >> (4 * (r & 7))) & 7;
    Vre[i * LD + (r & ~7) + p] = v.x;
    Vim[i * LD + (r & ~7) + p] = v.y;
  }
  __syncthreads();

  // ---- T_b (warp b) ----
  {
    const int b = warp;
    double gr[2] = {0.0, 0.0}, gi[2] = {0.0, 0.0}, gr2[2] = {0.0, 0.0}, gi2[2] = {0.0, 0.0};
    const double *vrb = Vre + (8 * b + fm) * LD, *vib = Vim + (8 * b + fm) * LD;
    for (int t = b; t < NB; ++t) {
      {
        const double vr = vrb[8 * t + fj], vi = vib[8 * t + fj];
        dmma884(gr[0], gr[1], vr, vr);
        dmma884(gi[0], gi[1], vr, vi);
        dmma884(gr[0], gr[1], vi, vi);
        dmma884(gi[0], gi[1], -vi, vr);
      }
      {
        const double vr = vrb[8 * t + 4 + (fj ^ 2)], vi = vib[8 * t + 4 + (fj ^ 2)];
        dmma884(gr2[0], gr2[1], vr, vr);
        dmma884(gi2[0], gi2[1], vr, vi);
        dmma884(gr2[0], gr2[1], vi, vi);
        dmma884(gi2[0], gi2[1], -vi, vr);
      }
    }
    cplx *gs = Gs + b * 64;
    gs[fm * 8 + 2 * fj] = make_c(gr[0] + gr2[0], gi[0] + gi2[0]);
    gs[fm * 8 + 2 * fj + 1] = make_c(gr[1] + gr2[1], gi[1] + gi2[1]);
    __syncwarp();
    if (lane < 8) {
      const int l = lane;
      cplx T[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const cplx ti = stau[8 * b + i];
        cplx acc = make_c(0.0, 0.0);
#pragma unroll
        for (int q = 0; q < i; ++q)
          if (q >= l) cfma(acc, T[q], gs[q * 8 + i]);
        const cplx off = cmul(make_c(-ti.x, -ti.y), acc);
        T[i] = (i == l) ? ti : ((i > l) ? off : make_c(0.0, 0.0));
      }
      // operand order of step 2: B[k = (j, e) <-> reflector 2j + e][n = u' <-> reflector 4 (u' % 2) + u' / 2] = T[refl(u')][2j + e]
      const int up = 2 * (l & 3) + (l >> 2);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        Tre[(b * 8 + 4 * (c & 1) + (c >> 1)) * TLD + up] = T[c].x;
        Tim[(b * 8 + 4 * (c & 1) + (c >> 1)) * TLD + up] = T[c].y;
      }
    }
  }

  // ---- C^T tiles of this warp's 8 columns ----
  double cr[NB][2], ci[NB][2];
  {
    const int n = 8 * warp + fm;
#pragma unroll
    for (int t = 0; t < NB; ++t)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int r = 8 * t + 2 * fj + e;
        cr[t][e] = (r < d && n < d) ? Zt[mat * dd + (size_t)r * d + n] : 0.0;
        ci[t][e] = 0.0;
      }
  }
  __syncthreads();

  const int p3 = (0x53427160u >> (4 * fm)) & 7;  // row position of step 3's N index
#pragma unroll 1
  for (int b = NB - 1; b >= 0; --b) {
    // 1. W^T = C^T conj(V_b): two accumulator sets (even / odd tiles) halve the dependent DMMA chain
    double wr[2] = {0.0, 0.0}, wi[2] = {0.0, 0.0}, xr[2] = {0.0, 0.0}, xi[2] = {0.0, 0.0};
    {
      const double *vrb = Vre + (8 * b + fm) * LD, *vib = Vim + (8 * b + fm) * LD;
#pragma unroll
      for (int t = 0; t < NB; ++t) {
        if (t >= b) {
          const double v0r = vrb[8 * t + fj], v0i = vib[8 * t + fj];
          const double v1r = vrb[8 * t + 4 + (fj ^ 2)], v1i = vib[8 * t + 4 + (fj ^ 2)];
          dmma884(wr[0], wr[1], cr[t][0], v0r);
          dmma884(wi[0], wi[1], ci[t][0], v0r);
          dmma884(xr[0], xr[1], cr[t][1], v1r);
          dmma884(xi[0], xi[1], ci[t][1], v1r);
          dmma884(wr[0], wr[1], ci[t][0], v0i);
          dmma884(wi[0], wi[1], cr[t][0], -v0i);
          dmma884(xr[0], xr[1], ci[t][1], v1i);
          dmma884(xi[0], xi[1], cr[t][1], -v1i);
        }
      }
      wr[0] += xr[0];
      wr[1] += xr[1];
      wi[0] += xi[0];
      wi[1] += xi[1];
    }
    // 2. W2^T = W^T T_b^T
    double yr[2] = {0.0, 0.0}, yi[2] = {0.0, 0.0};
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const double tr = Tre[(b * 8 + 4 * e + fj) * TLD + fm], ti = Tim[(b * 8 + 4 * e + fj) * TLD + fm];
      dmma884(yr[0], yr[1], wr[e], tr);
      dmma884(yi[0], yi[1], wr[e], ti);
      dmma884(yr[0], yr[1], -wi[e], ti);
      dmma884(yi[0], yi[1], wi[e], tr);
    }
    // 3. C^T -= W2^T V_b^T   (accumulator column 2j + e of W2^T <-> reflector 4e + j)
    const double nyr[2] = {-yr[0], -yr[1]}, nyi[2] = {-yi[0], -yi[1]};
#pragma unroll
    for (int t = 0; t < NB; ++t) {
      if (t >= b) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const double vr = Vre[(8 * b + 4 * e + fj) * LD + 8 * t + p3], vi = Vim[(8 * b + 4 * e + fj) * LD + 8 * t + p3];
          dmma884(cr[t][0], cr[t][1], nyr[e], vr);
          dmma884(ci[t][0], ci[t][1], nyr[e], vi);
          dmma884(cr[t][0], cr[t][1], yi[e], vi);
          dmma884(ci[t][0], ci[t][1], nyi[e], vr);
        }
      }
    }
  }

  {
    const int n = 8 * warp + fm;
#pragma unroll
    for (int t = 0; t < NB; ++t)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int r = 8 * t + 2 * fj + e;
        if (r < d && n < d) U[mat * dd + (size_t)r * d + n] = make_c(cr[t][e], ci[t][e]);
      }
  }
}

}  // namespace musim
