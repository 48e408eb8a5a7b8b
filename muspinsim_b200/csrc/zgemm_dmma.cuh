// zgemm_dmma.cuh -- batched complex FP64 GEMM on the FP64 tensor pipe (mma.sync m8n8k4 f64).
//
//   C[c] = op(A[c]) * B[c]      all d x d row-major complex128, op = identity or conj-transpose
//
// Replaces the dense `@` products of Operator.basis_change (/root/reference/muspinsim/
// spinop.py:349-350) and of Hamiltonian.fast_evolve (hamiltonian.py:207).  ncu on the
// vector-FMA version (rotate.cuh, cgemm_batched_kernel): 81 % LSU, 50 % FP64 -- a 2x2 register
// tile moves 4 B of shared memory per FMA.  Here a warp owns a 32 x 16 complex output tile as
// 8 DMMA accumulator fragments x (re, im) = 64 registers; one k4-step loads 8 A- and 4
// B-fragments (12 LDS.64) for 32 DMMAs = 8192 FMAs.  Operands are staged as separate re / im planes with a leading
// dimension = 4 (mod 16) so that every fragment load is bank-conflict free.
//
// B_MUON: B is not read from memory but formed on the fly as O U with the muon observable
// O = sum_a p_a S_mu^a (x) 1 (two non-zeros per row: spinsys.py:707-732), which removes the
// T = O U product and its HBM round trip from the fast path.
// EPI: 0 store C; 1 store (|C|^2 scale, 0); 2 store C + D; 3 store C .* conj(D);
//      4 / 5: the weights of 1 / 3 are not stored but summed into the decaying integral
//      out[slot] += weight/tau * Re sum_ab w_ab / (1/tau + 2 pi i (l_a - l_b))
//      (Hamiltonian.integrate_decaying, hamiltonian.py:150-162; experiment.py:492-496) -- the
//      ALC modes then never write or re-read the d x d weight matrix (9 KB per configuration at
//      d = 24, the HBM traffic that bounded integral_kernel).
// upper: only the tiles that touch i <= j are computed and stored (the polarisation kernels read
// the upper triangle of the Hermitian weight matrix only).
#pragma once
#include "common.cuh"
#include "polar.cuh"  // dmma884

namespace musim {

#define ZG_KS 16

struct IntEpi {  // arguments of the integral epilogues (EPI 4 / 5)
  const double *lam = nullptr, *wgt = nullptr;
  const int *slot = nullptr;
  double it = 0.0;  // 1 / tau
  double *out = nullptr;
};

struct MuonObs {   // O = [[pz/2, (px - i py)/2], [(px + i py)/2, -pz/2]] on the muon index
  int stride;      // product of the dimensions of the spins after the muon
  int enabled;
};

// PIPE (T = 3, upper only): 12 warps, one per LIVE tile (the 6 tiles strictly below the diagonal
// are not assigned at all), which lifts the register cap from 112 to 168 and leaves room to
// prefetch the next K slab into registers while the DMMAs of the current one run.
template <int T, bool CONJ_A, int EPI, bool B_MUON, bool PIPE = false>
__global__ void __launch_bounds__(PIPE ? 384 : 64 * T * T, (T == 3 ? 1 : (T == 2 ? 2 : 8)))
zgemm_dmma_kernel(int d, const cplx *__restrict__ A, size_t a_stride, const cplx *__restrict__ B,
                  size_t b_stride, cplx *C, double scale, const cplx *D, MuonObs mu,
                  const double *__restrict__ pvec, int upper, IntEpi ie = IntEpi()) {
  constexpr int DP = 32 * T;     // padded dimension
  constexpr int LD = DP + 4;     // = 4 (mod 16)
  constexpr int LDK = ZG_KS + 4; // for the non-transposed A slab [m][k]
  constexpr int NT = PIPE ? 384 : 64 * T * T;
  static_assert(!PIPE || T == 3, "PIPE is the T = 3 upper-triangle variant");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sAr = reinterpret_cast<double *>(smem_raw);
  double *sAi = sAr + (CONJ_A ? ZG_KS * LD : DP * LDK);
  double *sBr = sAi + (CONJ_A ? ZG_KS * LD : DP * LDK);
  double *sBi = sBr + ZG_KS * LD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // this warp's 32 (rows) x 16 (columns) tile; PIPE: the 12 tiles with 16 wn + 15 >= 32 wm
  const int wm = PIPE ? (warp < 6 ? 0 : (warp < 10 ? 1 : 2)) : warp / (2 * T);
  const int wn = PIPE ? (warp < 6 ? warp : (warp < 10 ? warp - 4 : warp - 6)) : warp % (2 * T);
  const int fr = lane >> 2, fk = lane & 3;
  // upper: the consumer reads C[i][j] for i <= j only (Hermitian result): warps whose tile lies
  // strictly below the diagonal only help staging
  const bool tile_live = PIPE || !upper || (wn * 16 + 15 >= wm * 32);
  const size_t cfg = blockIdx.x;
  const size_t dd = (size_t)d * d;
  const cplx *Ab = A + cfg * a_stride;
  const cplx *Bb = B + cfg * b_stride;
  double px = 0, py = 0, pz = 0;
  if (B_MUON) {
    px = 0.5 * pvec[cfg * 3];
    py = 0.5 * pvec[cfg * 3 + 1];
    pz = 0.5 * pvec[cfg * 3 + 2];
  }
  double cr[4][2][2], ci[4][2][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0;

  if (PIPE) {
    // element e of a slab: A element (k = e / DP, m = e % DP) [CONJ_A] and B element (k, n) likewise;
    // thread t owns e = t, t + NT, ... (4 of each)
    constexpr int EPT = (ZG_KS * DP + NT - 1) / NT;
    cplx ra[EPT], rb[EPT];
    // PAIRED (C = U^H (O U) with A and B the same matrix U): the K rows of a slab are taken as 8
    // rows with muon index 0 plus their 8 partner rows (muon index 1), so that both U rows a
    // B-operand element needs are A-operand elements of the SAME thread (elements q and q ^ 2):
    // one global load per element instead of three.
    const bool paired = B_MUON && CONJ_A && EPT == 4 && 2 * NT == 8 * DP && (const void *)Ab == (const void *)Bb &&
                        (mu.stride % 8) == 0 && (d % (2 * mu.stride)) == 0;
    auto fetch_paired = [&](int k0) {
      const int r0 = k0 >> 1;  // first of the 8 muon-index-0 rows of this slab, counted among those rows only
      const int blk = r0 / mu.stride, within = r0 - blk * mu.stride;
      const int row0 = blk * 2 * mu.stride + within;
#pragma unroll
      for (int q = 0; q < EPT; ++q) {
        const int e = tid + q * NT;
        const int k = e / DP, m = e - k * DP;
        const int row = row0 + (k & 7) + ((k >> 3) ? mu.stride : 0);
        ra[q] = (m < d) ? Ab[(size_t)row * d + m] : make_c(0.0, 0.0);
      }
#pragma unroll
      for (int q = 0; q < EPT; ++q) {
        const int mi = q >> 1;  // elements 0, 1: muon index 0; elements 2, 3: their partners
        const cplx u0 = ra[q], u1 = ra[q ^ 2];
        const double dg = mi ? -pz : pz;
        const double oy = mi ? py : -py;
        rb[q].x = dg * u0.x + px * u1.x - oy * u1.y;
        rb[q].y = dg * u0.y + px * u1.y + oy * u1.x;
      }
    };
    auto fetch = [&](int k0) {
      if (paired) {
        fetch_paired(k0);
        return;
      }
#pragma unroll
      for (int q = 0; q < EPT; ++q) {
        const int e = tid + q * NT;
        ra[q] = make_c(0.0, 0.0);
        rb[q] = make_c(0.0, 0.0);
        if (e < ZG_KS * DP) {
          if (CONJ_A) {
            const int k = e / DP, m = e - k * DP;
            if (k0 + k < d && m < d) ra[q] = Ab[(size_t)(k0 + k) * d + m];
          } else {
            const int m = e / ZG_KS, k = e - m * ZG_KS;
            if (k0 + k < d && m < d) ra[q] = Ab[(size_t)m * d + k0 + k];
          }
          const int k = e / DP, n = e - k * DP;
          const int kk = k0 + k;
          if (kk < d && n < d) {
            if (B_MUON) {
              const int m = (kk / mu.stride) & 1;
              const cplx u0 = Bb[(size_t)kk * d + n];
              const cplx u1 = Bb[(size_t)(m ? kk - mu.stride : kk + mu.stride) * d + n];
              const double dg = m ? -pz : pz;
              const double oy = m ? py : -py;
              rb[q].x = dg * u0.x + px * u1.x - oy * u1.y;
              rb[q].y = dg * u0.y + px * u1.y + oy * u1.x;
            } else {
              rb[q] = Bb[(size_t)kk * d + n];
            }
          }
        }
      }
    };
    auto deposit = [&]() {
#pragma unroll
      for (int q = 0; q < EPT; ++q) {
        const int e = tid + q * NT;
        if (e < ZG_KS * DP) {
          if (CONJ_A) {
            const int k = e / DP, m = e - k * DP;
            sAr[k * LD + m] = ra[q].x;
            sAi[k * LD + m] = ra[q].y;
          } else {
            const int m = e / ZG_KS, k = e - m * ZG_KS;
            sAr[m * LDK + k] = ra[q].x;
            sAi[m * LDK + k] = ra[q].y;
          }
          const int k = e / DP, n = e - k * DP;
          sBr[k * LD + n] = rb[q].x;
          sBi[k * LD + n] = rb[q].y;
        }
      }
    };
    fetch(0);
    deposit();
    __syncthreads();
    for (int k0 = 0; k0 < d; k0 += ZG_KS) {
      const bool more = k0 + ZG_KS < d;
      if (more) fetch(k0 + ZG_KS);  // global loads in flight during the DMMAs below
#pragma unroll
      for (int ks = 0; ks < ZG_KS; ks += 4) {
        double ar[4], ai[4], br[2], bi[2], nb[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int m = wm * 32 + i * 8 + fr;
          if (CONJ_A) {
            ar[i] = sAr[(ks + fk) * LD + m];
            ai[i] = -sAi[(ks + fk) * LD + m];  // conj
          } else {
            ar[i] = sAr[m * LDK + ks + fk];
            ai[i] = sAi[m * LDK + ks + fk];
          }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int n = wn * 16 + j * 8 + fr;
          br[j] = sBr[(ks + fk) * LD + n];
          bi[j] = sBi[(ks + fk) * LD + n];
          nb[j] = -bi[j];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            dmma884(cr[i][j][0], cr[i][j][1], ar[i], br[j]);
            dmma884(cr[i][j][0], cr[i][j][1], ai[i], nb[j]);
            dmma884(ci[i][j][0], ci[i][j][1], ar[i], bi[j]);
            dmma884(ci[i][j][0], ci[i][j][1], ai[i], br[j]);
          }
      }
      __syncthreads();
      if (more) {
        deposit();
        __syncthreads();
      }
    }
  } else {
  for (int k0 = 0; k0 < d; k0 += ZG_KS) {
    // ---- stage the K slab (planar re / im) ----
    if (CONJ_A) {
      for (int e = tid; e < ZG_KS * DP; e += NT) {
        const int k = e / DP, m = e - k * DP;
        cplx v = make_c(0.0, 0.0);
        if (k0 + k < d && m < d) v = Ab[(size_t)(k0 + k) * d + m];
        sAr[k * LD + m] = v.x;
        sAi[k * LD + m] = v.y;
      }
    } else {
      for (int e = tid; e < ZG_KS * DP; e += NT) {
        const int m = e / ZG_KS, k = e - m * ZG_KS;
        cplx v = make_c(0.0, 0.0);
        if (k0 + k < d && m < d) v = Ab[(size_t)m * d + k0 + k];
        sAr[m * LDK + k] = v.x;
        sAi[m * LDK + k] = v.y;
      }
    }
    for (int e = tid; e < ZG_KS * DP; e += NT) {
      const int k = e / DP, n = e - k * DP;
      cplx v = make_c(0.0, 0.0);
      const int kk = k0 + k;
      if (kk < d && n < d) {
        if (B_MUON) {
          // (O U)[kk][n] = o_diag U[kk][n] + o_off U[kk ^ muon][n]
          const int m = (kk / mu.stride) & 1;
          const cplx u0 = Bb[(size_t)kk * d + n];
          const cplx u1 = Bb[(size_t)(m ? kk - mu.stride : kk + mu.stride) * d + n];
          const double dg = m ? -pz : pz;
          const double oy = m ? py : -py;  // o_off = px -/+ i py
          v.x = dg * u0.x + px * u1.x - oy * u1.y;
          v.y = dg * u0.y + px * u1.y + oy * u1.x;
        } else {
          v = Bb[(size_t)kk * d + n];
        }
      }
      sBr[k * LD + n] = v.x;
      sBi[k * LD + n] = v.y;
    }
    __syncthreads();
    if (tile_live) {
#pragma unroll
    for (int ks = 0; ks < ZG_KS; ks += 4) {
      // T = 1 (d <= 32, e.g. the 24 x 24 systems of the ALC scans): skip the k steps and the 8-row / 8-column
      // tiles that lie entirely in the zero padding (uniform branches): 144 instead of 256 DMMAs per warp at d = 24
      if (T == 1 && k0 + ks >= d) continue;
      double ar[4], ai[4], br[2], bi[2], nb[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = wm * 32 + i * 8 + fr;
        if (CONJ_A) {
          ar[i] = sAr[(ks + fk) * LD + m];
          ai[i] = -sAi[(ks + fk) * LD + m];  // conj
        } else {
          ar[i] = sAr[m * LDK + ks + fk];
          ai[i] = sAi[m * LDK + ks + fk];
        }
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int n = wn * 16 + j * 8 + fr;
        br[j] = sBr[(ks + fk) * LD + n];
        bi[j] = sBi[(ks + fk) * LD + n];
        nb[j] = -bi[j];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (T == 1 && (wm * 32 + i * 8 >= d || wn * 16 + j * 8 >= d)) continue;
          dmma884(cr[i][j][0], cr[i][j][1], ar[i], br[j]);  // + Ar Br
          dmma884(cr[i][j][0], cr[i][j][1], ai[i], nb[j]);  // - Ai Bi
          dmma884(ci[i][j][0], ci[i][j][1], ar[i], bi[j]);  // + Ar Bi
          dmma884(ci[i][j][0], ci[i][j][1], ai[i], br[j]);  // + Ai Br
        }
    }
    }
    __syncthreads();
  }
  }
  if (EPI >= 4) {
    // ---- integral epilogue (all tiles are live: the callers pass upper = 0) ----
    const double twopi = 6.283185307179586476925286766559;
    const double *lc = ie.lam + cfg * d;
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int m = wm * 32 + i * 8 + fr;
        const int n = wn * 16 + j * 8 + fk * 2;
        if (m < d) {
          const double lm = lc[m];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (n + e < d) {
              cplx v = make_c(cr[i][j][e], ci[i][j][e]);
              if (EPI == 4)
                v = make_c(cnorm2(v) * scale, 0.0);
              else
                v = cmulc(v, D[cfg * dd + (size_t)m * d + n + e]);
              const double om = twopi * (lm - lc[n + e]);
              acc += (v.x * ie.it + v.y * om) / (ie.it * ie.it + om * om);
            }
          }
        }
      }
    acc = warp_sum(acc);
    double *red = sAr;  // the K loop ended with a barrier: the staging buffers are free
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int q = 0; q < NT / 32; ++q) t += red[q];
      atomicAdd(&ie.out[ie.slot[cfg]], ie.wgt[cfg] * t * ie.it);
    }
    return;
  }
  if (!tile_live) return;
  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int m = wm * 32 + i * 8 + fr;
      const int n = wn * 16 + j * 8 + fk * 2;
      if (m < d) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          if (n + e < d) {
            cplx v = make_c(cr[i][j][e], ci[i][j][e]);
            const size_t idx = cfg * dd + (size_t)m * d + n + e;
            if (EPI == 1) v = make_c(cnorm2(v) * scale, 0.0);
            if (EPI == 2) v = cadd(v, D[idx]);
            if (EPI == 3) v = cmulc(v, D[idx]);
            C[idx] = v;
          }
        }
      }
    }
}

template <int T, bool CONJ_A>
inline size_t zgemm_dmma_smem() {
  constexpr int DP = 32 * T, LD = DP + 4, LDK = ZG_KS + 4;
  return (2 * (CONJ_A ? ZG_KS * LD : DP * LDK) + 2 * ZG_KS * LD) * sizeof(double);
}

// host launcher; returns false if d is outside the tensor-pipe kernel's range (d <= 96)
template <bool CONJ_A, int EPI, bool B_MUON>
inline bool launch_zgemm_dmma(int d, int64_t n, const cplx *A, size_t as, const cplx *B, size_t bs, cplx *C,
                              double scale, const cplx *D, MuonObs mu, const double *pvec, cudaStream_t st,
                              bool upper = false, IntEpi ie = IntEpi(), bool pipe = true) {
  // pipe (option "zgemm_pipe"): 12-warp register-prefetch variant for upper-triangle outputs at d in (64, 96]
  if (d > 96) return false;
  if (d <= 32) {
    const size_t sm = zgemm_dmma_smem<1, CONJ_A>();
    zgemm_dmma_kernel<1, CONJ_A, EPI, B_MUON><<<(unsigned)n, 64, sm, st>>>(d, A, as, B, bs, C, scale, D, mu, pvec, upper ? 1 : 0, ie);
  } else if (d <= 64) {
    const size_t sm = zgemm_dmma_smem<2, CONJ_A>();
    cudaFuncSetAttribute(zgemm_dmma_kernel<2, CONJ_A, EPI, B_MUON>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    zgemm_dmma_kernel<2, CONJ_A, EPI, B_MUON><<<(unsigned)n, 256, sm, st>>>(d, A, as, B, bs, C, scale, D, mu, pvec, upper ? 1 : 0, ie);
  } else {
    const size_t sm = zgemm_dmma_smem<3, CONJ_A>();
    if (upper && pipe && EPI < 4) {
      cudaFuncSetAttribute(zgemm_dmma_kernel<3, CONJ_A, EPI, B_MUON, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      zgemm_dmma_kernel<3, CONJ_A, EPI, B_MUON, true><<<(unsigned)n, 384, sm, st>>>(d, A, as, B, bs, C, scale, D, mu, pvec, 1);
      return true;
    }
    cudaFuncSetAttribute(zgemm_dmma_kernel<3, CONJ_A, EPI, B_MUON>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    zgemm_dmma_kernel<3, CONJ_A, EPI, B_MUON><<<(unsigned)n, 576, sm, st>>>(d, A, as, B, bs, C, scale, D, mu, pvec, upper ? 1 : 0, ie);
  }
  return true;
}

}  // namespace musim
