// eigh_large.cuh -- Householder + QL stages for matrices that do not fit one SM (d > 112).
//
// Same algorithm and same intermediate formats as eigh_hql.cuh (zhetd2 / zung2r restated, QL with
// recorded rotations, U = Q Zt), replacing np.linalg.eigh in Hermitian.diag
// (/root/reference/muspinsim/spinop.py:51-82) for the system sizes the reference reaches with more
// spins (e.g. mu + e + 6 x 1H: d = 256).  The working matrix lives in GLOBAL memory (it is
// L2-resident: 4 MB at d = 512) instead of shared memory or registers:
//   * hql_tridiag_gmem_kernel: one CTA per matrix, column-major A in a workspace; thread (r, g)
//     loops over rows r, r + R, ... so consecutive threads touch consecutive addresses.
//   * hql_apply_rows_kernel: the rotation replay with a BLOCK OF ROWS of Zt per CTA in shared
//     memory (rows are independent; every CTA of a matrix streams the same rotation list).
// The QL kernel (hql_tql_kernel) and the complex x real GEMM are size-generic and shared.
// These kernels are the capability path for large d, not a tuned one: the Hamiltonian sizes
// BASELINE.json names (d <= 96) never reach them.
#pragma once
#include "eigh_hql.cuh"

namespace musim {

#define HQL_LARGE_MAX_D 1024

template <bool BUILD_H>
__global__ void __launch_bounds__(512)
hql_tridiag_gmem_kernel(int d, int R, int G, const cplx *__restrict__ H0, const cplx *__restrict__ Z,
                        const double *__restrict__ Bf, const cplx *__restrict__ Ain, cplx *__restrict__ Awork,
                        double *__restrict__ dout, double *__restrict__ eout, cplx *__restrict__ Qout) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ld = d | 1;
  cplx *sv = reinterpret_cast<cplx *>(smem_raw);  // [d]
  cplx *sw = sv + d;                               // [d]
  cplx *stau = sw + d;                             // [d]
  cplx *spart = stau + d;                          // [G][d]
  double *red = reinterpret_cast<double *>(spart + (size_t)G * d);  // [66]
  const int tid = threadIdx.x, nth = blockDim.x;
  const int r0 = tid % R, g = tid / R;
  const size_t cfg = blockIdx.x;
  const size_t dd = (size_t)d * d;
  cplx *A = Awork + cfg * (size_t)d * ld;  // (r,c) at [c*ld + r]

  double bx = 0, by = 0, bz = 0;
  if (BUILD_H) {
    bx = Bf[cfg * 3 + 0];
    by = Bf[cfg * 3 + 1];
    bz = Bf[cfg * 3 + 2];
  }
  // Hermitian part of the input, written column-major: element (rr, cc) = (in(rr,cc) + conj(in(cc,rr))) / 2
  auto elem = [&](size_t idx) {
    cplx a;
    if (BUILD_H) {
      a = H0[idx];
      const cplx z0 = Z[idx], z1 = Z[dd + idx], z2 = Z[2 * dd + idx];
      a.x += bx * z0.x + by * z1.x + bz * z2.x;
      a.y += bx * z0.y + by * z1.y + bz * z2.y;
    } else {
      a = Ain[cfg * dd + idx];
    }
    return a;
  };
  for (int idx = tid; idx < d * d; idx += nth) {
    const int cc = idx / d, rr = idx - cc * d;  // consecutive threads -> consecutive rows of a column
    const cplx a = elem((size_t)rr * d + cc), b = elem((size_t)cc * d + rr);
    A[(size_t)cc * ld + rr] = make_c(0.5 * (a.x + b.x), rr == cc ? 0.0 : 0.5 * (a.y - b.y));
  }
  __syncthreads();

  // ---- zhetd2 (lower) ----
  for (int k = 0; k < d - 1; ++k) {
    const int m = d - k - 1;
    const int o = k + 1;  // offset of the trailing block
    cplx *colk = A + (size_t)k * ld + o;
    double xn = 0.0;
    for (int i = 1 + tid; i < m; i += nth) xn += cnorm2(colk[i]);
    xn = block_sum(xn, red);
    const cplx alpha = colk[0];
    if (tid == 0) dout[cfg * d + k] = A[(size_t)k * ld + k].x;
    if (xn == 0.0 && alpha.y == 0.0) {  // H = I
      if (tid == 0) {
        eout[cfg * d + k] = alpha.x;
        stau[k] = make_c(0.0, 0.0);
      }
      __syncthreads();
      continue;
    }
    const double beta = -copysign(sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xn), alpha.x);
    const cplx tau = make_c((beta - alpha.x) / beta, -alpha.y / beta);
    const double ar = alpha.x - beta, ai = alpha.y;
    const double den = 1.0 / (ar * ar + ai * ai);
    const cplx scale = make_c(ar * den, -ai * den);
    __syncthreads();  // everyone has read alpha before colk[0..] is overwritten
    if (tid == 0) {
      eout[cfg * d + k] = beta;
      stau[k] = tau;
    }
    for (int i = tid; i < m; i += nth) {
      cplx vi = make_c(1.0, 0.0);
      if (i > 0) {
        vi = cmul(colk[i], scale);
        colk[i] = vi;
      }
      sv[i] = vi;
    }
    __syncthreads();
    // p = tau * A22 v  (row r, columns c = g, g+G, ...)
    if (g < G) {
      for (int r = r0; r < m; r += R) {
        cplx acc = make_c(0.0, 0.0);
        const cplx *row = A + (size_t)o * ld + o + r;
        for (int c = g; c < m; c += G) cfma(acc, row[(size_t)c * ld], sv[c]);
        spart[(size_t)g * d + r] = acc;
      }
    }
    __syncthreads();
    cplx dot = make_c(0.0, 0.0);
    for (int i = tid; i < m; i += nth) {
      cplx s = spart[i];
      for (int gg = 1; gg < G; ++gg) s = cadd(s, spart[(size_t)gg * d + i]);
      const cplx pr = cmul(tau, s);
      sw[i] = pr;
      const cplx t = ccmul(pr, sv[i]);  // conj(p) * v
      dot.x += t.x;
      dot.y += t.y;
    }
    dot = block_sum2(dot, red);
    // alpha2 = -1/2 * tau * dot ;  w = p + alpha2 * v
    const cplx a2 = cscale(-0.5, cmul(tau, dot));
    for (int i = tid; i < m; i += nth) sw[i] = cadd(sw[i], cmul(a2, sv[i]));
    __syncthreads();
    // A22 -= v w^H + w v^H
    if (g < G) {
      for (int r = r0; r < m; r += R) {
        const cplx vr = sv[r], wr = sw[r];
        cplx *row = A + (size_t)o * ld + o + r;
        for (int c = g; c < m; c += G) {
          cplx a = row[(size_t)c * ld];
          const cplx wc = sw[c], vc = sv[c];
          a.x -= vr.x * wc.x + vr.y * wc.y + wr.x * vc.x + wr.y * vc.y;
          a.y -= vr.y * wc.x - vr.x * wc.y + wr.y * vc.x - wr.x * vc.y;
          row[(size_t)c * ld] = a;
        }
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    dout[cfg * d + d - 1] = A[(size_t)(d - 1) * ld + d - 1].x;
    eout[cfg * d + d - 1] = 0.0;
  }
  __syncthreads();

  // ---- zung2r, in place: Q = H_0 H_1 ... H_{d-2} ----
  for (int k = d - 2; k >= 0; --k) {
    const int m1 = d - k - 2;  // length of v[1:]
    const int o = k + 2;
    const cplx t = stau[k];
    for (int i = tid; i < m1; i += nth) sv[i] = A[(size_t)k * ld + o + i];
    __syncthreads();
    // u_c = sum_i conj(v_i) Q(o+i, o+c): thread (c, g) sums rows i = g, g+G, ... of column o+c.
    // Consecutive threads read different columns here (stride ld), which is the slow direction;
    // m1 <= d and the block is L2-resident.
    if (g < G) {
      for (int c = r0; c < m1; c += R) {
        cplx acc = make_c(0.0, 0.0);
        const cplx *col = A + (size_t)(o + c) * ld + o;
        for (int i = g; i < m1; i += G) ccfma(acc, sv[i], col[i]);
        spart[(size_t)g * d + c] = acc;
      }
    }
    __syncthreads();
    for (int c = tid; c < m1; c += nth) {
      cplx u = spart[c];
      for (int gg = 1; gg < G; ++gg) u = cadd(u, spart[(size_t)gg * d + c]);
      sw[c] = u;
      const cplx tu = cmul(t, u);
      A[(size_t)(o + c) * ld + k + 1] = make_c(-tu.x, -tu.y);  // row k+1
      const cplx tv = cmul(t, sv[c]);
      A[(size_t)(k + 1) * ld + o + c] = make_c(-tv.x, -tv.y);  // column k+1
    }
    if (tid == 0) A[(size_t)(k + 1) * ld + k + 1] = make_c(1.0 - t.x, -t.y);
    __syncthreads();
    if (g < G) {
      for (int r = r0; r < m1; r += R) {
        const cplx tv = cmul(t, sv[r]);
        cplx *row = A + (size_t)o * ld + o + r;
        for (int c = g; c < m1; c += G) {
          cplx a = row[(size_t)c * ld];
          const cplx u = sw[c];
          a.x -= tv.x * u.x - tv.y * u.y;
          a.y -= tv.x * u.y + tv.y * u.x;
          row[(size_t)c * ld] = a;
        }
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
      q = A[(size_t)cc * ld + rr];
    Qout[cfg * dd + idx] = q;
  }
}

struct HqlLargeGeom {
  int R, G, nth;
};
inline HqlLargeGeom hql_large_geom(int d) {
  HqlLargeGeom g;
  g.nth = 512;
  g.R = std::min(512, (d + 31) & ~31);
  g.G = std::max(1, g.nth / g.R);
  return g;
}
inline size_t hql_tridiag_gmem_smem(int d, const HqlLargeGeom &g) {
  return (3 * (size_t)d + (size_t)g.G * d) * sizeof(cplx) + 70 * sizeof(double);
}

// Rotation replay for a block of RB rows of Zt (blockDim.x = RB, one row per thread, row block
// blockIdx.y), rows in shared memory as sZ[c * RB + tid]; same rotation stream format as
// hql_apply_kernel.  Output columns permuted by `perm` (ascending eigenvalues).
__global__ void __launch_bounds__(128)
hql_apply_rows_kernel(int d, const double2 *__restrict__ rot, size_t rot_cap, const SweepIdx *__restrict__ swp,
                      int swp_cap, const int *__restrict__ nswp, const unsigned short *__restrict__ perm,
                      double *__restrict__ Zt) {
  extern __shared__ __align__(16) unsigned char apply_smem[];
  const int RB = blockDim.x;
  double2 *ring = reinterpret_cast<double2 *>(apply_smem);       // [2][HQL_TILE]
  double *sZ = reinterpret_cast<double *>(ring + 2 * HQL_TILE);  // [d][RB]
  const int tid = threadIdx.x;
  const size_t mat = blockIdx.x;
  const int r = blockIdx.y * RB + tid;
  const bool active = r < d;
  const double2 *myrot = rot + mat * rot_cap;
  const SweepIdx *msw = swp + mat * swp_cap;
  const int ns = nswp[mat];
  size_t total = 0;
  for (int i = 0; i < ns; ++i) total += (size_t)(msw[i].m - msw[i].l);
  const int ntiles = (int)((total + HQL_TILE - 1) / HQL_TILE);
  auto stage = [&](int t) {
    if (t < ntiles) {
      const size_t base = (size_t)t * HQL_TILE;
      for (int e = tid; e < HQL_TILE; e += RB)
        if (base + e < total) cp_async16(&ring[(t & 1) * HQL_TILE + e], &myrot[base + e]);
    }
    cp_async_commit();
  };
  stage(0);
  stage(1);
  for (int c = 0; c < d; ++c) sZ[(size_t)c * RB + tid] = (c == r) ? 1.0 : 0.0;
  cp_async_wait<1>();
  __syncthreads();

  size_t gi = 0;
  int tile = 0;
  for (int sidx = 0; sidx < ns; ++sidx) {
    const int l = msw[sidx].l, m = msw[sidx].m;
    double x = sZ[(size_t)m * RB + tid];
    int j = m - 1;
    while (j >= l) {
      const int in_tile = (int)(gi - (size_t)tile * HQL_TILE);
      const int cnt = min(j - l + 1, HQL_TILE - in_tile);
      const double2 *rs = ring + (tile & 1) * HQL_TILE + in_tile;
      for (int k = 0; k < cnt; ++k) {
        const double a = sZ[(size_t)(j - k) * RB + tid];
        const double2 cs = rs[k];
        sZ[(size_t)(j - k + 1) * RB + tid] = cs.x * x - cs.y * a;
        x = fma(cs.y, x, cs.x * a);
      }
      gi += cnt;
      j -= cnt;
      if (gi == (size_t)(tile + 1) * HQL_TILE && gi < total) {
        __syncthreads();
        stage(tile + 2);
        cp_async_wait<1>();
        __syncthreads();
        ++tile;
      }
    }
    sZ[(size_t)l * RB + tid] = x;
  }
  cp_async_wait<0>();
  if (!active) return;
  const unsigned short *pm = perm + mat * d;
  double *zrow = Zt + mat * (size_t)d * d + (size_t)r * d;
  for (int jj = 0; jj < d; ++jj) zrow[jj] = sZ[(size_t)pm[jj] * RB + tid];
}

inline int hql_apply_rows_rb(int d) {
  // rows per CTA: as many as fit ~200 KB of shared memory, a multiple of 32, at most 128
  const size_t avail = 200 * 1024 - 2 * HQL_TILE * sizeof(double2);
  int rb = std::min<int>(128, (int)(avail / ((size_t)d * sizeof(double))));
  rb = rb >= 32 ? (rb & ~31) : (rb & ~7);
  return rb;
}
inline size_t hql_apply_rows_smem(int d, int rb) {
  return 2 * HQL_TILE * sizeof(double2) + (size_t)d * rb * sizeof(double) + 16;
}

}  // namespace musim
