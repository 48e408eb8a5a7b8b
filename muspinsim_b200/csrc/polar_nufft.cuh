// polar_nufft.cuh -- polarisation on a UNIFORM time grid as a type-1 non-uniform FFT.
//
//   out[slot_c, k] += weight_c * ( sum_i Re w_ii + 2 sum_{i<j} Re( w_ij exp(-2 pi i f_ij t_k) ) ),
//   t_k = t0 + k dt,  k = 0 .. N-1,   f_ij = l_i - l_j
//
// Replaces the time loop of Hamiltonian.evolve (/root/reference/muspinsim/hamiltonian.py:87-107)
// and the Cython kernel parallel_fast_time_evolve (cython/parallel.pyx:56-67) -- same sum, other
// order.  The time-factorised DMMA kernel (polar.cuh) spends 2 FMAs per (pair, time point); on a
// uniform grid the sum is the discrete Fourier transform of point masses at the non-uniform
// positions u = -f dt (cycles per step, taken mod 1 -- exact for a uniform grid), so it costs
// O(w) per PAIR plus one length-M FFT per output row instead:
//   1. spread kernel: every (configuration, pair) adds  c * phi(m - u M)  to w = 12 neighbouring
//      cells of a fine grid of M >= 2 N cells (phi = "exponential of semicircle" kernel
//      exp(beta (sqrt(1 - z^2) - 1)), beta = 2.30 w, evaluated as a degree-10 polynomial per tap);
//      c carries the weights, exp(-2 pi i f t0) and the half-band shift exp(2 pi i (N/2) u), so
//      that the wanted modes are k' = k - N/2 in [-N/2, N/2).  All configurations accumulated into
//      one output row share one grid: the powder average happens BEFORE the transform.
//   2. FFT kernel (one CTA per output row): s_k' = sum_m grid[m] exp(2 pi i k' m / M), divided
//      by the kernel's Fourier transform phi^(k').
// Aliasing error for w = 12, M/N >= 2: 7e-13 * sum|c| (tools/nufft_prototype.py), far inside the
// 1e-9 parity bar; non-uniform time arrays keep the direct kernel.
//
// Half grid: only the REAL part of the transform is wanted, and Re(c e^{2 pi i k' u}) =
// Re(conj(c) e^{2 pi i k' (-u)}), so every point is first moved to u in [0, 1/2] (strength conjugated
// if it was mirrored -- once per point, nothing per tap).  The points then occupy the cells
// -5 .. M/2 + 6 only; the private grids hold exactly that range (no wrap-around mask either), and
// when a private grid is flushed its margin cells m < 0 (m > M/2) are added, conjugated, to cell -m
// (M - m) of the global grid -- the same mirror identity applied to a cell instead of a point.  The
// global grid keeps M cells (upper half zero) and the FFT kernel is the plain complex one.
//
// Spread kernel mapping: one private grid per WARP in shared memory ((M/2 + 32) complex = 16.5 KB
// at N <= 1024: 12 warps per SM with the 156-register variant), so there are no atomics: a half-warp
// handles one point, lane = tap, plain read-modify-write of 12 consecutive cells.  The two
// half-warps' windows may overlap (then the two updates are issued one after the other).
#pragma once
#include <cstddef>
#include <vector>

#include "common.cuh"
#include "polar.cuh"

namespace musim {

#define NU_W 12
#define NU_MARG (NU_W / 2 - 1)  // cells of margin below cell 0 of the half grid (tap 0 of a point in cell 0)
#define MUSIM_NU_SMEM (227 * 1024)
#define NU_DEG 10
#define NU_BETA (2.30 * NU_W)

// coef[l][m]: phi_l(y) = sum_m coef[l][m] y^m, y = 2 x in [-1, 1], x = fractional offset of the
// point from the centre of its cell; tap l sits (l - (w-1)/2 - x) cells from the point.
__constant__ double nu_coef[NU_W][NU_DEG + 1];

struct __align__(16) NuPoint {
  double cre, cim, y;
  int base, flag;
};

__device__ __forceinline__ void lds_f64x2(unsigned addr, double &x, double &y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(x), "=d"(y) : "r"(addr) : "memory");
}
__device__ __forceinline__ void sts_f64x2(unsigned addr, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(addr), "d"(x), "d"(y) : "memory");
}

// volatile: keeps its place between the volatile shared-memory accesses around it (the compiler
// otherwise sinks the whole polynomial evaluation below the update loop it is meant to overlap)
__device__ __forceinline__ double fma_pinned(double a, double b, double c) {
  double r;
  asm volatile("fma.rn.f64 %0, %1, %2, %3;\n" : "=d"(r) : "d"(a), "d"(b), "d"(c));
  return r;
}
__device__ __forceinline__ double lds_f64(unsigned addr) {
  double r;
  asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(r) : "r"(addr) : "memory");
  return r;
}

__device__ __forceinline__ double nu_horner(const double (&cf)[NU_DEG + 1], double y) {
  double r = cf[NU_DEG];
#pragma unroll
  for (int m = NU_DEG - 1; m >= 0; --m) r = fma(r, y, cf[m]);
  return r;
}

// units: unit u = (configuration u / S, part u % S of its pair list, `plen` pairs each).  CTA b owns
// the contiguous units [b * units_per_cta, ...); its warps draw them one at a time from a shared
// counter (the warps of a CTA do not finish their units at the same time).
//
// Per warp and batch of 32 pairs:
//   prep     every lane turns one pair into a point (cell offset, strength) -> staging buffer
//   phase A  Horner evaluation of the kernel at this lane's tap for the 16 points of its half-warp
//   phase B  16 read-modify-write steps on the private grid
// Software pipeline: pair indices are loaded three batches ahead, W / lambda two batches ahead, and
// phase A of batch b+1 is interleaved with phase B of batch b (FP64 work under the LDS latency).
// CREG: the strengths of the batch's points stay in registers (230 registers: two warps per scheduler); else they
// are re-read from the staging buffer in every step (one more LDS.128 per step, 156 registers: three warps).
template <bool CREG>
__global__ void __launch_bounds__(CREG ? 256 : 384)
polar_nufft_spread_kernel(int d, int npairs, const PairIdx *__restrict__ pairs, int n_cfg, int S, int plen,
                          int units_per_cta, const cplx *__restrict__ W, const double *__restrict__ lam,
                          const double *__restrict__ wgt, const int *__restrict__ slot, int N, double t0,
                          double dt, int M, int PG, cplx *__restrict__ G, int *__restrict__ touched) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  cplx *grid_all = reinterpret_cast<cplx *>(smem_raw);
  cplx *grid = grid_all + (size_t)warp * PG;
  NuPoint *stage = reinterpret_cast<NuPoint *>(grid_all + (size_t)nwarps * PG) + warp * 32;
  int *wslot = reinterpret_cast<int *>(reinterpret_cast<NuPoint *>(grid_all + (size_t)nwarps * PG) + nwarps * 32);
  int *ctr = wslot + nwarps;
  const int tap = lane & 15, half = lane >> 4;
  const bool tap_on = tap < NU_W;
  const int Mh = M >> 1;
  const unsigned grid_s = (unsigned)__cvta_generic_to_shared(grid);
  const unsigned stage_s = (unsigned)__cvta_generic_to_shared(stage);
  double cf[NU_DEG + 1];
#pragma unroll
  for (int m = 0; m <= NU_DEG; ++m) cf[m] = tap_on ? nu_coef[tap_on ? tap : 0][m] : 0.0;

  for (int m = lane; m < PG; m += 32) grid[m] = make_c(0.0, 0.0);
  if (threadIdx.x == 0) *ctr = 0;
  __syncthreads();

  const size_t dd = (size_t)d * d;
  const long total_units = (long)n_cfg * S;
  const long U0 = (long)blockIdx.x * units_per_cta;
  const long U1 = min(total_units, U0 + units_per_cta);
  const double halfN = (double)(N / 2);
  int cur_slot = -1;

  // private cell q holds fine-grid cell m = q - NU_MARG of the HALF grid [0, M/2] plus the margins the
  // windows of points near its two ends reach into; a margin cell is the mirror image of cell -m
  // (or M - m): it is added, conjugated, to that cell (see "Half grid" above)
  auto fold_add = [&](cplx *Gs, int q, cplx v) {
    int m = q - NU_MARG;
    if (m < 0) {
      m = -m;
      v.y = -v.y;
    } else if (m > Mh) {
      m = M - m;
      v.y = -v.y;
    }
    if (v.x != 0.0) atomicAdd(&Gs[m].x, v.x);
    if (v.y != 0.0) atomicAdd(&Gs[m].y, v.y);
  };
  auto flush_warp = [&](int s) {
    cplx *Gs = G + (size_t)s * M;
    for (int m = lane; m < PG; m += 32) {
      fold_add(Gs, m, grid[m]);
      grid[m] = make_c(0.0, 0.0);
    }
    if (lane == 0) touched[s] = 1;
    __syncwarp();
  };

  struct It {
    long u;
    int c, pb, p_end;
  };
  auto it_draw = [&](It &it) {  // next unit of this CTA (warp-uniform)
    int k = 0;
    if (lane == 0) k = atomicAdd(ctr, 1);
    k = __shfl_sync(0xffffffffu, k, 0);
    it.u = min(U1, U0 + (long)k);
    it.c = 0;
    it.pb = 0;
    it.p_end = 0;
    if (it.u < U1) {
      it.c = (int)(it.u / S);
      const int part = (int)(it.u - (long)it.c * S);
      it.pb = part * plen;
      it.p_end = min(npairs, it.pb + plen);
    }
  };
  auto it_next = [&](It &it) {
    if (it.u >= U1) return;
    it.pb += 32;
    if (it.pb >= it.p_end) it_draw(it);
  };
  auto load_pair = [&](const It &it) {
    PairIdx ij;
    ij.i = 0xffff;  // marks "no point"
    ij.j = 0;
    const int p = it.pb + lane;
    if (it.u < U1 && p < it.p_end) ij = pairs[p];
    return ij;
  };
  struct Dat {
    cplx w;
    double li, lj, wc;
  };
  auto load_dat = [&](const It &it, PairIdx ij) {
    Dat q;
    q.w = make_c(0.0, 0.0);
    q.li = q.lj = q.wc = 0.0;
    if (ij.i != 0xffff) {
      q.w = W[(size_t)it.c * dd + (size_t)ij.i * d + ij.j];
      q.li = lam[(size_t)it.c * d + ij.i];
      q.lj = lam[(size_t)it.c * d + ij.j];
      q.wc = wgt[it.c];
    }
    return q;
  };
  // one pair -> one point in the staging buffer; point q of the batch goes to entry 2 (q % 16) +
  // q / 16, so the two points of a step share a 64-byte line
  auto prep = [&](PairIdx ij, const Dat &q) {
    NuPoint pt;
    pt.y = 0.0;
    pt.cre = 0.0;
    pt.cim = 0.0;
    pt.base = 0;
    pt.flag = 0;
    if (ij.i != 0xffff) {
      const double f = q.li - q.lj;
      const double x = f * dt;
      const double uu = rint(x) - x;  // -frac(f dt) in [-1/2, 1/2]: cycles per time step
      const bool mirror = uu < 0.0;   // half grid: the point at -u with the conjugate strength has the same real part
      const double pos = fabs(uu) * (double)M;  // in [0, M/2]
      const double fl = floor(pos);
      pt.y = 2.0 * (pos - fl) - 1.0;
      pt.base = (int)fl - (NU_W / 2 - 1) + NU_MARG;  // private cell of tap 0, in [0, M/2 + NU_MARG]
      // strength: weights * exp(-2 pi i f t0) * exp(2 pi i (N/2) u)
      double ph = fma(halfN, uu, -f * t0);
      ph -= rint(ph);
      double sn, cs;
      sincospi(2.0 * ph, &sn, &cs);
      const double sc = (ij.i == ij.j) ? q.wc : 2.0 * q.wc;
      pt.cre = sc * fma(q.w.x, cs, -q.w.y * sn);
      pt.cim = sc * fma(q.w.x, sn, q.w.y * cs);
      if (mirror) pt.cim = -pt.cim;
    }
    // points q and q + 16 are updated in the same step by the two half-warps: do their windows
    // (16 cells: lanes 12..15 add zeros) overlap?  Symmetric, so both halves see the same flag.
    const int ob = __shfl_xor_sync(0xffffffffu, pt.base, 16);
    const int diff = pt.base - ob;
    pt.flag = (diff < 16 && diff > -16) ? 1 : 0;
    stage[2 * tap + half] = pt;
  };

  double yy[16], phi[16], cre[CREG ? 16 : 1], cim[CREG ? 16 : 1];
  cre[0] = cim[0] = 0.0;
  unsigned off[16];
  unsigned ovl = 0;
  auto load_y = [&]() {
#pragma unroll
    for (int st = 0; st < 16; ++st) yy[st] = lds_f64(stage_s + (unsigned)((2 * st + half) * sizeof(NuPoint) + offsetof(NuPoint, y)));
  };
  auto load_points = [&]() {
    ovl = 0;
#pragma unroll
    for (int st = 0; st < 16; ++st) {
      const NuPoint *q = stage + 2 * st + half;
      double2 a = make_double2(0.0, 0.0);
      if (CREG) a = *reinterpret_cast<const double2 *>(&q->cre);
      const int2 bf = *reinterpret_cast<const int2 *>(&q->base);
      if (CREG) {
        cre[st] = a.x;
        cim[st] = a.y;
      }
      off[st] = grid_s + (((unsigned)bf.x + tap) << 4);
      ovl |= (unsigned)bf.y << st;
    }
  };

  It i0, i1, i2;
  it_draw(i0);
  i1 = i0;
  it_next(i1);
  PairIdx ij0 = load_pair(i0), ij1 = load_pair(i1);
  Dat d0 = load_dat(i0, ij0);
  while (i0.u < U1) {
    // global loads for the next batches
    i2 = i1;
    it_next(i2);
    const PairIdx ij2 = load_pair(i2);
    const Dat d1 = load_dat(i1, ij1);
    {
      const int s = slot[i0.c];
      if (s != cur_slot) {
        if (cur_slot >= 0) flush_warp(cur_slot);
        cur_slot = s;
      }
    }
    prep(ij0, d0);
    __syncwarp();
    load_y();
    load_points();
    // Phases A and B as one systolic schedule: the Horner chain of point q advances by one
    // iteration in each of the ten steps before step q (true dependencies, so neither compiler
    // stage can move the polynomial work away from the loads it is meant to overlap); each step
    // therefore has up to 10 INDEPENDENT DFMAs between the grid load and its use.
    // Shared-memory accesses of one converged warp are performed in program order, so the
    // update of step s is visible to the load of step s+1.  Where the two half-warps' windows
    // overlap, the upper half skips its store and repeats the step afterwards.
#pragma unroll
    for (int q = 0; q < 16; ++q) phi[q] = cf[NU_DEG];
#pragma unroll
    for (int it = 0; it < NU_DEG; ++it)  // iterations that fall before step 0
#pragma unroll
      for (int q = 0; q < NU_DEG - it; ++q) phi[q] = fma_pinned(phi[q], yy[q], cf[NU_DEG - 1 - it]);
    const unsigned skip = half ? ovl : 0u;
#pragma unroll
    for (int st = 0; st < 16; ++st) {
      double vx, vy;
      lds_f64x2(off[st], vx, vy);
#pragma unroll
      for (int q = st + 1; q < 16 && q <= st + NU_DEG; ++q)
        phi[q] = fma_pinned(phi[q], yy[q], cf[NU_DEG - 1 - (st - q + NU_DEG)]);
      double c_re = cre[CREG ? st : 0], c_im = cim[CREG ? st : 0];
      if (!CREG) lds_f64x2(stage_s + (unsigned)((2 * st + half) * sizeof(NuPoint)), c_re, c_im);
      vx = fma(c_re, phi[st], vx);
      vy = fma(c_im, phi[st], vy);
      if (!((skip >> st) & 1u)) sts_f64x2(off[st], vx, vy);
    }
    if (ovl) {  // warp-uniform, rare
      __syncwarp();
#pragma unroll
      for (int st = 0; st < 16; ++st) {
        if ((skip >> st) & 1u) {
          double vx, vy;
          lds_f64x2(off[st], vx, vy);
          double c_re = cre[CREG ? st : 0], c_im = cim[CREG ? st : 0];
          if (!CREG) lds_f64x2(stage_s + (unsigned)((2 * st + half) * sizeof(NuPoint)), c_re, c_im);
          vx = fma(c_re, phi[st], vx);
          vy = fma(c_im, phi[st], vy);
          sts_f64x2(off[st], vx, vy);
        }
        __syncwarp();
      }
    }
    __syncwarp();
    i0 = i1;
    i1 = i2;
    ij0 = ij1;
    ij1 = ij2;
    d0 = d1;
  }
  // ---- end: if every warp of the CTA ended on the same output row, reduce the private grids
  // inside the CTA first (one atomic per cell per CTA instead of per warp) ----
  if (lane == 0) wslot[warp] = cur_slot;
  __syncthreads();
  bool same = true;
  const int s0 = wslot[0];
  for (int q = 1; q < nwarps; ++q) same = same && (wslot[q] == s0);
  if (same) {
    if (s0 >= 0) {
      cplx *Gs = G + (size_t)s0 * M;
      for (int m = threadIdx.x; m < PG; m += blockDim.x) {
        cplx v = grid_all[m];
        for (int q = 1; q < nwarps; ++q) {
          const cplx t = grid_all[(size_t)q * PG + m];
          v.x += t.x;
          v.y += t.y;
        }
        fold_add(Gs, m, v);
      }
      if (threadIdx.x == 0) touched[s0] = 1;
    }
  } else if (cur_slot >= 0) {
    flush_warp(cur_slot);
  }
}

// One CTA per output row: in-place radix-2 FFT (positive exponent) of the row's fine grid in
// shared memory, deconvolution, accumulation into out.  Leaves G[row] zeroed and the flag reset,
// so the workspace keeps its all-zero invariant between launches.
__global__ void __launch_bounds__(256)
polar_nufft_fft_kernel(int M, int logM, int N, cplx *__restrict__ G, int *__restrict__ touched,
                       const double *__restrict__ deconv, double *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx *a = reinterpret_cast<cplx *>(smem_raw);
  const int row = blockIdx.x;
  if (!touched[row]) return;
  cplx *Gs = G + (size_t)row * M;
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const int r = (int)(__brev((unsigned)m) >> (32 - logM));
    a[r] = Gs[m];
    Gs[m] = make_c(0.0, 0.0);
  }
  __syncthreads();
  if (threadIdx.x == 0) touched[row] = 0;
  for (int s = 0; s < logM; ++s) {
    const int hs = 1 << s;
    for (int t = threadIdx.x; t < M / 2; t += blockDim.x) {
      const int k = t & (hs - 1);
      const int i0 = ((t >> s) << (s + 1)) + k, i1 = i0 + hs;
      double sn, cs;
      sincospi((double)k / (double)hs, &sn, &cs);  // exp(+2 pi i k / (2 hs))
      const cplx x0 = a[i0], x1 = a[i1];
      const cplx tw = make_c(fma(x1.x, cs, -x1.y * sn), fma(x1.x, sn, x1.y * cs));
      a[i0] = make_c(x0.x + tw.x, x0.y + tw.y);
      a[i1] = make_c(x0.x - tw.x, x0.y - tw.y);
    }
    __syncthreads();
  }
  const int hN = N / 2;
  for (int k = threadIdx.x; k < N; k += blockDim.x) {
    const int idx = (k - hN) & (M - 1);
    atomicAdd(&out[(size_t)row * N + k], a[idx].x * deconv[k]);
  }
}

// ---------------------------------------------------------------------------------------
// host side: kernel tables, workspace, launcher
// ---------------------------------------------------------------------------------------
inline double nu_es(double z) {
  if (!(fabs(z) < 1.0)) return 0.0;
  return exp(NU_BETA * (sqrt(1.0 - z * z) - 1.0));
}

// per-tap interpolation polynomials at Chebyshev nodes (monomials in y = 2x), long double solve
inline void nu_build_coef(double coef[NU_W][NU_DEG + 1]) {
  const int n = NU_DEG + 1;
  const long double PI = 3.14159265358979323846264338327950288L;
  for (int l = 0; l < NU_W; ++l) {
    long double A[NU_DEG + 1][NU_DEG + 2];
    for (int r = 0; r < n; ++r) {
      const long double y = cosl(PI * (r + 0.5L) / n);
      const double dist = (double)l - (NU_W - 1) / 2.0 - (double)(y / 2);
      long double pw = 1.0L;
      for (int m = 0; m < n; ++m) {
        A[r][m] = pw;
        pw *= y;
      }
      A[r][n] = (long double)nu_es(2.0 * dist / NU_W);
    }
    for (int c = 0; c < n; ++c) {
      int piv = c;
      for (int r = c + 1; r < n; ++r)
        if (fabsl(A[r][c]) > fabsl(A[piv][c])) piv = r;
      for (int m = 0; m <= n; ++m) std::swap(A[c][m], A[piv][m]);
      for (int r = 0; r < n; ++r) {
        if (r == c) continue;
        const long double f = A[r][c] / A[c][c];
        for (int m = c; m <= n; ++m) A[r][m] -= f * A[c][m];
      }
    }
    for (int m = 0; m < n; ++m) coef[l][m] = (double)(A[m][n] / A[m][m]);
  }
}

// deconv[k] = 1 / phi^(k - N/2),  phi^(k') = (w/2) int_{-1}^{1} phi(z) cos(pi w k' z / M) dz
// (Gauss-Legendre, 128 nodes)
inline void nu_build_deconv(int N, int M, std::vector<double> &dec) {
  const int nq = 128;
  std::vector<double> xq(nq), wq(nq);
  const double PI = 3.14159265358979323846;
  for (int i = 0; i < (nq + 1) / 2; ++i) {
    double x = cos(PI * (i + 0.75) / (nq + 0.5));
    double dp = 1.0;
    for (int it = 0; it < 100; ++it) {
      double p0 = 1.0, p1 = x;
      for (int k = 2; k <= nq; ++k) {
        const double p2 = ((2.0 * k - 1.0) * x * p1 - (k - 1.0) * p0) / k;
        p0 = p1;
        p1 = p2;
      }
      dp = nq * (x * p1 - p0) / (x * x - 1.0);
      const double dx = p1 / dp;
      x -= dx;
      if (fabs(dx) < 1e-16) break;
    }
    {  // derivative at the converged node
      double p0 = 1.0, p1 = x;
      for (int k = 2; k <= nq; ++k) {
        const double p2 = ((2.0 * k - 1.0) * x * p1 - (k - 1.0) * p0) / k;
        p0 = p1;
        p1 = p2;
      }
      dp = nq * (x * p1 - p0) / (x * x - 1.0);
    }
    xq[i] = -x;
    xq[nq - 1 - i] = x;
    wq[i] = wq[nq - 1 - i] = 2.0 / ((1.0 - x * x) * dp * dp);
  }
  std::vector<double> ph(nq);
  for (int i = 0; i < nq; ++i) ph[i] = wq[i] * nu_es(xq[i]);
  dec.resize(N);
  for (int k = 0; k < N; ++k) {
    const double kk = (double)(k - N / 2);
    double s = 0.0;
    for (int i = 0; i < nq; ++i) s += ph[i] * cos(PI * NU_W * kk * xq[i] / M);
    dec[k] = 1.0 / (0.5 * NU_W * s);
  }
}

struct NufftWs {
  cplx *G = nullptr;
  int *touched = nullptr;
  double *deconv = nullptr;
  int64_t rows = 0;
  int M = 0, N = 0, dN = 0, dM = 0;
  bool coef_ready = false;

  void release() {
    dev_free(G);
    dev_free(touched);
    dev_free(deconv);
    G = nullptr;
    touched = nullptr;
    deconv = nullptr;
    rows = 0;
    M = N = dN = dM = 0;
  }
};

inline int nu_grid_size(int N) {
  int M = 64;
  while (M < 2 * N) M <<= 1;
  return M;
}

// Cells of a warp's private HALF grid: cells -NU_MARG .. M/2 + 15 - NU_MARG (lane 15 of a point in cell M/2:
// the lanes 12..15 of a half-warp add zeros to the four cells behind the window), padded.
inline int nu_private_cells(int M) { return ((M / 2 + NU_MARG + 16) + 15) & ~15; }

// Whether the NUFFT path applies: uniform grid handled by the caller; here only sizes.
inline bool nu_supported(int N, int64_t n_slots) {
  if (N < 2) return false;
  const int M = nu_grid_size(N);
  if (M > 8192) return false;                                    // one private grid per warp must fit an SM
  if ((double)n_slots * M * sizeof(cplx) > 1.0e9) return false;  // fine-grid workspace
  return true;
}

// Returns cudaSuccess or the failing error; *launches += 2.
inline cudaError_t launch_polar_nufft(NufftWs &ws, int d, int npairs, const PairIdx *pairs, int64_t n,
                                      const cplx *W, const double *lam, const double *wgt, const int *slot,
                                      int N, double t0, double dt, int n_slots, double *out, cudaStream_t st,
                                      int64_t *launches) {
  cudaError_t e;
  const int M = nu_grid_size(N);
  int logM = 0;
  while ((1 << logM) < M) ++logM;
  if (!ws.coef_ready) {
    double coef[NU_W][NU_DEG + 1];
    nu_build_coef(coef);
    e = cudaMemcpyToSymbol(nu_coef, coef, sizeof coef);
    if (e != cudaSuccess) return e;
    ws.coef_ready = true;
  }
  if (ws.M != M || ws.rows < n_slots) {
    dev_free(ws.G);
    dev_free(ws.touched);
    ws.G = nullptr;
    ws.touched = nullptr;
    ws.rows = 0;
    e = dev_malloc((void **)&ws.G, (size_t)n_slots * M * sizeof(cplx));
    if (e != cudaSuccess) return e;
    e = dev_malloc((void **)&ws.touched, (size_t)n_slots * sizeof(int));
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(ws.G, 0, (size_t)n_slots * M * sizeof(cplx), st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(ws.touched, 0, (size_t)n_slots * sizeof(int), st);
    if (e != cudaSuccess) return e;
    ws.M = M;
    ws.rows = n_slots;
  }
  if (ws.dN != N || ws.dM != M) {
    std::vector<double> dec;
    nu_build_deconv(N, M, dec);
    dev_free(ws.deconv);
    ws.deconv = nullptr;
    e = dev_malloc((void **)&ws.deconv, (size_t)N * sizeof(double));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(ws.deconv, dec.data(), (size_t)N * sizeof(double), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(st);  // `dec` is pageable host memory going out of scope
    if (e != cudaSuccess) return e;
    ws.dN = N;
    ws.dM = M;
  }
  // geometry: as many warps (private grids) per CTA as shared memory allows, one CTA per SM
  const int PG = nu_private_cells(M);
  const size_t per_warp = (size_t)PG * sizeof(cplx) + 32 * sizeof(NuPoint);
  int nwarps = (int)std::min<size_t>(8, (size_t)(MUSIM_NU_SMEM - 256) / per_warp);
  // 12 private grids fit (M <= 2048): the 156-register variant that re-reads the strengths from the staging
  // buffer, three warps per scheduler (3.43 vs 3.91 ms at C5); else the 230-register variant with up to 8 warps
  const bool creg = (size_t)(MUSIM_NU_SMEM - 256) / per_warp < 12;
  if (!creg) nwarps = 12;
  if (nwarps < 1) return cudaErrorInvalidConfiguration;
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const long total_warps = (long)n_sm * nwarps;
  int S = 1;
  if (n < 8 * total_warps) {
    S = (int)std::min<long>((8 * total_warps + n - 1) / n, std::max(1, npairs / 64));
    if (S < 1) S = 1;
  }
  int plen = (npairs + S - 1) / S;
  plen = (plen + 31) & ~31;
  S = (npairs + plen - 1) / plen;
  const long units = (long)n * S;
  const long upc = (units + n_sm - 1) / n_sm;  // contiguous units per CTA, drawn dynamically by its warps
  const int ctas = (int)((units + upc - 1) / upc);
  const size_t smem = nwarps * per_warp + (nwarps + 1) * sizeof(int) + 16;
  e = cudaFuncSetAttribute(polar_nufft_spread_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(polar_nufft_spread_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (creg)
    polar_nufft_spread_kernel<true><<<ctas, 32 * nwarps, smem, st>>>(d, npairs, pairs, (int)n, S, plen, (int)upc, W, lam, wgt,
                                                                  slot, N, t0, dt, M, PG, ws.G, ws.touched);
  else
    polar_nufft_spread_kernel<false><<<ctas, 32 * nwarps, smem, st>>>(d, npairs, pairs, (int)n, S, plen, (int)upc, W, lam, wgt,
                                                                   slot, N, t0, dt, M, PG, ws.G, ws.touched);
  const size_t fsmem = (size_t)M * sizeof(cplx);
  e = cudaFuncSetAttribute(polar_nufft_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
  if (e != cudaSuccess) return e;
  polar_nufft_fft_kernel<<<n_slots, 256, fsmem, st>>>(M, logM, N, ws.G, ws.touched, ws.deconv, out);
  *launches += 2;
  return cudaGetLastError();
}

}  // namespace musim
