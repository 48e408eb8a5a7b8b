// eigh_dispatch.cuh -- host-side launcher and workspace of the two batched eigensolvers.
#pragma once
#include "eigh_hql.cuh"
#include "eigh_jacobi.cuh"
#include "eigh_backwy.cuh"
#include "eigh_large.cuh"
#include "eigh_tdc.cuh"
#include "eigh_tridiag_rw.cuh"
#include "eigh_tridiag_hs.cuh"
#include "eigh_tridiag_warp.cuh"
#include "profiler.cuh"
#include "rotate.cuh"

namespace musim {

#define MUSIM_MAX_SMEM_OPTIN (227 * 1024)

enum { EIGH_AUTO = 0, EIGH_JACOBI = 1, EIGH_HQL = 2 };
// Kernel-selection options of the eigensolver (per handle; musim_set_option).  The defaults are the
// measured-fastest variants; the others stay as independent cross-checks covered by the parity tests.
struct EighOpts {
  bool reflect = true;         // "reflect": K4 applies the reflectors to Zt (d <= 96); 0: Q formed in K1 + GEMM
  bool tdc = true;             // "tdc": tridiagonal divide and conquer (32 < d <= 96) instead of QL + rotation replay
  bool back_wy_small = true;   // "back_wy_small": the compact-WY kernels also for 16 < d <= 32 (D = 24 / 32 instantiations, one CTA per matrix)
  bool back_wy = true;         // "back_wy": K4 in compact-WY blocks on the FP64 tensor pipe (32 < d <= 96); 0: level-2 reflector kernel
  bool tridiag_warp = true;    // "tridiag_warp": warp-per-matrix tridiagonalisation for d <= 32
  bool small24 = true;         // "small24": D = 16 / 24 instantiations of the replay / reflector kernels for d <= 16 / 24 (else D = 32)
  bool tridiag_fused = true;   // "tridiag_fused": warp kernel with the update of step k fused into the product of step k+1
  bool tridiag_phases = true;  // "tridiag_phases": K1 in up to three launches of decreasing size
  bool apply_warp = true;      // "apply_warp": rotation replay with one warp per CTA (d > 32)
  int tql_threads = 0;         // "tql_threads": matrices per block of the QL kernel (8, 16 or 32; 0 = auto: 32 for d <= 32, else 16)
  bool tridiag_rw = true;      // "tridiag_rw": rows-per-warp register tridiagonalisation (d <= 96); 0: shared-memory kernel
  bool tridiag_hsw = false;    // "tridiag_hsw": register-resident half-storage warp kernel for 8 < d <= 32 (and the last phase of larger d) instead of the shared-memory warp kernel: same speed at d = 32, 4 % slower at d = 24 (cross-check)
  bool tridiag_hs = true;      // "tridiag_hs": half-storage DMMA tridiagonalisation for the phases with a live block > 48 (48 < d <= 96)
  bool use_reflect(int d) const { return reflect && d >= 3 && d <= 96; }
};

inline int pick_eigh(long opt, int d) {
  if (opt == EIGH_JACOBI) return EIGH_JACOBI;
  if (opt == EIGH_HQL) return hql_supported(d) ? EIGH_HQL : EIGH_JACOBI;
  return hql_supported(d) ? EIGH_HQL : EIGH_JACOBI;
}

struct EighWs {
  int64_t cap = 0;
  int d = 0, method = 0;
  cplx *Vg = nullptr;
  double *dbuf[2] = {nullptr, nullptr}, *ebuf[2] = {nullptr, nullptr}, *Zt = nullptr;
  cplx *Awork = nullptr;  // column-major working matrices of the d > HQL_MAX_D path
  cplx *Q[2] = {nullptr, nullptr};  // [2]: stage A (tridiagonalisation) of the next launch group overlaps stage B
  bool dbl = false;
  cplx *Vp[2] = {nullptr, nullptr}, *tauv[2] = {nullptr, nullptr};  // packed reflectors + tau (d <= 96 path)
  // tridiagonal divide and conquer (eigh_tdc.cuh), 32 < d <= 96: the four leaves of every matrix as a batch
  // of 4 n independent 24 x 24 problems for the batched QL kernels
  double *l_hdr = nullptr, *l_d = nullptr, *l_e = nullptr, *l_lam = nullptr, *l_Z = nullptr;
  double2 *l_rot = nullptr;
  SweepIdx *l_swp = nullptr;
  int *l_nswp = nullptr;
  unsigned short *l_perm = nullptr;
  static constexpr size_t L_ROT_CAP = 2 * 24 * 24 + 64 + 14 * (6 * 24 + 16);
  static constexpr int L_SWP_CAP = 6 * 24 + 16;
  double *Timg = nullptr;  // compact-WY T factors in operand order (eigh_backwy.cuh), 32 < d <= 96
  size_t vcap = 0;
  double2 *rot = nullptr;
  SweepIdx *swp = nullptr;
  int *nswp = nullptr;
  unsigned short *perm = nullptr;
  size_t rot_cap = 0;
  int swp_cap = 0;

  // complex entries of Q per matrix: the eigenvector / phase buffer.  The trailing blocks handed between the K1 phases
  // need 72^2 + 48^2 entries for d > 48 and 64^2 + 32^2 for 32 < d <= 48; small systems only need d^2 (a fixed
  // floor would shrink the launch groups of the 10^7-configuration scans at d = 24 by 13 x)
  static size_t q_entries(int d) {
    const size_t dd = (size_t)d * d;
    return d > 48 ? std::max<size_t>(dd, 72 * 72 + 48 * 48) : (d > 32 ? std::max<size_t>(dd, 64 * 64 + 32 * 32) : dd);
  }

  static bool jacobi_vglobal(int d) { return eigh_jacobi_smem(d, false) > MUSIM_MAX_SMEM_OPTIN; }

  static size_t bytes_per_matrix(int method, int d) {
    const size_t dd = (size_t)d * d;
    if (method == EIGH_HQL)
      return (d > HQL_MAX_D ? (size_t)d * (d | 1) * sizeof(cplx) : 0) +
             2 * (2 * d * sizeof(double) + q_entries(d) * sizeof(cplx) + (dd / 2 + d) * sizeof(cplx)) + dd * sizeof(double) +
             (2 * dd + 64 + 14 * (6 * d + 16)) * sizeof(double2) + (6 * d + 16) * sizeof(SweepIdx) + sizeof(int) +
             d * sizeof(unsigned short);
    return jacobi_vglobal(d) ? (size_t)d * (d | 1) * sizeof(cplx) : 0;
  }

  void release() {
    dev_free(Vg);
    for (int i = 0; i < 2; ++i) {
      dev_free(dbuf[i]);
      dev_free(ebuf[i]);
      dev_free(Q[i]);
      dev_free(Vp[i]);
      dev_free(tauv[i]);
      dbuf[i] = ebuf[i] = nullptr;
      Q[i] = nullptr;
      Vp[i] = tauv[i] = nullptr;
    }
    dev_free(Zt);
    dev_free(Timg);
    Timg = nullptr;
    dev_free(l_hdr);
    dev_free(l_d);
    dev_free(l_e);
    dev_free(l_lam);
    dev_free(l_Z);
    dev_free(l_rot);
    dev_free(l_swp);
    dev_free(l_nswp);
    dev_free(l_perm);
    l_hdr = l_d = l_e = l_lam = l_Z = nullptr;
    l_rot = nullptr;
    l_swp = nullptr;
    l_nswp = nullptr;
    l_perm = nullptr;
    dev_free(Awork);
    Awork = nullptr;
    dev_free(rot);
    dev_free(swp);
    dev_free(nswp);
    dev_free(perm);
    Vg = nullptr;
    Zt = nullptr;
    rot = nullptr;
    swp = nullptr;
    nswp = nullptr;
    perm = nullptr;
    cap = 0;
  }

  cudaError_t ensure(int method_, int d_, int64_t n, bool dbl_ = false) {
    if (cap >= n && d == d_ && method == method_ && (dbl || !dbl_)) return cudaSuccess;
    release();
    d = d_;
    method = method_;
    dbl = dbl_;
    const size_t dd = (size_t)d * d;
    cudaError_t e = cudaSuccess;
#define EW_ALLOC(ptr, count)                                              \
  if (e == cudaSuccess) e = dev_malloc((void **)&ptr, (count) * sizeof(*ptr));
    if (method == EIGH_HQL) {
      rot_cap = 2 * dd + 64 + 14 * (size_t)(6 * d + 16);  // ~1.2 d^2 rotations observed + <= 14 padding entries per sweep; overflow is reported as ENOTCONV
      swp_cap = 6 * d + 16;
      vcap = (size_t)(d - 1) * (d - 2) / 2 + 8;
      for (int i = 0; i < (dbl ? 2 : 1); ++i) {
        EW_ALLOC(dbuf[i], (size_t)n * d);
        EW_ALLOC(ebuf[i], (size_t)n * d);
        EW_ALLOC(Q[i], (size_t)n * q_entries(d));
        EW_ALLOC(Vp[i], (size_t)n * vcap);
        EW_ALLOC(tauv[i], (size_t)n * d);
      }
      EW_ALLOC(Zt, (size_t)n * dd);
      if (d > 16 && d <= 96) EW_ALLOC(Timg, (size_t)n * (d <= 32 ? BackWyGeom<32>::TIMG : BackWyGeom<96>::TIMG));
      if (d > 32 && d <= 96) {
        EW_ALLOC(l_hdr, (size_t)n * 4);
        EW_ALLOC(l_d, (size_t)n * 4 * 24);
        EW_ALLOC(l_e, (size_t)n * 4 * 24);
        EW_ALLOC(l_lam, (size_t)n * 4 * 24);
        EW_ALLOC(l_Z, (size_t)n * 4 * 24 * 24);
        EW_ALLOC(l_rot, (size_t)n * 4 * L_ROT_CAP);
        EW_ALLOC(l_swp, (size_t)n * 4 * L_SWP_CAP);
        EW_ALLOC(l_nswp, (size_t)n * 4);
        EW_ALLOC(l_perm, (size_t)n * 4 * 24);
      }
      if (d > HQL_MAX_D) EW_ALLOC(Awork, (size_t)n * d * (d | 1));
      EW_ALLOC(rot, (size_t)n * rot_cap);
      EW_ALLOC(swp, (size_t)n * swp_cap);
      EW_ALLOC(nswp, (size_t)n);
      EW_ALLOC(perm, (size_t)n * d);
    } else if (jacobi_vglobal(d)) {
      EW_ALLOC(Vg, (size_t)n * d * (d | 1));
    }
#undef EW_ALLOC
    if (e == cudaSuccess) cap = n;
    return e;
  }
};

inline void launch_tridiag_hsw(int d, int64_t n, const cplx *H0, const cplx *Z, const double *B, const cplx *Ain, double *dd_,
                               double *ee_, cplx *vp_, size_t vcap, cplx *tt_, int dstride, int koff, cudaStream_t st) {
  const unsigned grid = (unsigned)((n + HSW_WARPS - 1) / HSW_WARPS);
  if (d <= 16)
    hql_tridiag_hsw_kernel<2><<<grid, 32 * HSW_WARPS, 0, st>>>(d, n, H0, Z, B, Ain, dd_, ee_, vp_, vcap, tt_, dstride, koff);
  else if (d <= 24)
    hql_tridiag_hsw_kernel<3><<<grid, 32 * HSW_WARPS, 0, st>>>(d, n, H0, Z, B, Ain, dd_, ee_, vp_, vcap, tt_, dstride, koff);
  else
    hql_tridiag_hsw_kernel<4><<<grid, 32 * HSW_WARPS, 0, st>>>(d, n, H0, Z, B, Ain, dd_, ee_, vp_, vcap, tt_, dstride, koff);
}

// Stage A of the Householder+QL solver: tridiagonalise + form Q into buffer set `buf`.
// (For Jacobi, stage A is the whole solver.)  Returns 0, a cudaError_t (> 0), or -5.
inline int launch_eigh_stageA(int method, int d, int64_t n, const cplx *H0, const cplx *Z, const double *B,
                              const cplx *Ain, double *lam, cplx *U, EighWs &ws, int buf, int *status,
                              cudaStream_t st, int64_t *launches, Profiler *prof, const EighOpts &o) {
  cudaError_t e;
  if (method == EIGH_HQL) {
    if (!hql_supported(d)) return -5;
    if (d > HQL_MAX_D) {  // working matrix in global memory (eigh_large.cuh)
      const HqlLargeGeom lg = hql_large_geom(d);
      const size_t sm = hql_tridiag_gmem_smem(d, lg);
      ProfScope ps(prof, st, PH_EIGH_TRIDIAG);
      if (Ain) {
        e = cudaFuncSetAttribute(hql_tridiag_gmem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return (int)e;
        hql_tridiag_gmem_kernel<false><<<(unsigned)n, lg.nth, sm, st>>>(d, lg.R, lg.G, H0, Z, B, Ain, ws.Awork, ws.dbuf[buf],
                                                                        ws.ebuf[buf], ws.Q[buf]);
      } else {
        e = cudaFuncSetAttribute(hql_tridiag_gmem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return (int)e;
        hql_tridiag_gmem_kernel<true><<<(unsigned)n, lg.nth, sm, st>>>(d, lg.R, lg.G, H0, Z, B, Ain, ws.Awork, ws.dbuf[buf],
                                                                       ws.ebuf[buf], ws.Q[buf]);
      }
      ++*launches;
      return (int)cudaGetLastError();
    }
    const HqlGeom g = hql_geom(d);
    const size_t smem = hql_tridiag_smem(d, g);
    if (smem > MUSIM_MAX_SMEM_OPTIN) return -5;
    ProfScope ps(prof, st, PH_EIGH_TRIDIAG);
    if (o.use_reflect(d) && o.tridiag_warp && o.tridiag_hsw && d > 8 && d <= 32) {
      launch_tridiag_hsw(d, n, H0, Z, B, Ain, ws.dbuf[buf], ws.ebuf[buf], ws.Vp[buf], ws.vcap, ws.tauv[buf], d, 0, st);
    } else if (o.use_reflect(d) && o.tridiag_warp && o.tridiag_fused && d <= 32) {
      const size_t sm = hql_tridiag_warpf_smem(d);
      cudaFuncSetAttribute(hql_tridiag_warpf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      hql_tridiag_warpf_kernel<<<(unsigned)((n + TRW_WARPS - 1) / TRW_WARPS), 32 * TRW_WARPS, sm, st>>>(
          d, n, H0, Z, B, Ain, ws.dbuf[buf], ws.ebuf[buf], ws.Vp[buf], ws.vcap, ws.tauv[buf], d, 0);
    } else if (o.use_reflect(d) && o.tridiag_warp && d <= 32) {
      const size_t sm = hql_tridiag_warp_smem(d);
      cudaFuncSetAttribute(hql_tridiag_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      hql_tridiag_warp_kernel<<<(unsigned)((n + TRW_WARPS - 1) / TRW_WARPS), 32 * TRW_WARPS, sm, st>>>(
          d, n, H0, Z, B, Ain, ws.dbuf[buf], ws.ebuf[buf], ws.Vp[buf], ws.vcap, ws.tauv[buf]);
    } else if (o.use_reflect(d) && o.tridiag_rw) {
      double *dd_ = ws.dbuf[buf], *ee_ = ws.ebuf[buf];
      cplx *vp_ = ws.Vp[buf], *tt_ = ws.tauv[buf];
      // phase buffers for the trailing blocks (Q is not used on the reflector path)
      cplx *A64 = ws.Q[buf], *A32 = ws.Q[buf] + (size_t)n * 64 * 64;
      {
      const unsigned g = (unsigned)n;
        if (d <= 32) {
          hql_tridiag_rw_kernel<32><<<g, 128, 0, st>>>(d, d, 0, d, H0, Z, B, Ain, dd_, ee_, vp_, ws.vcap, tt_, nullptr);
        } else if (!o.tridiag_phases) {
          if (d <= 64)
            hql_tridiag_rw_kernel<64><<<g, 256, 0, st>>>(d, d, 0, d, H0, Z, B, Ain, dd_, ee_, vp_, ws.vcap, tt_, nullptr);
          else
            hql_tridiag_rw_kernel<96><<<g, 384, 0, st>>>(d, d, 0, d, H0, Z, B, Ain, dd_, ee_, vp_, ws.vcap, tt_, nullptr);
        } else if (o.tridiag_hs && d > 48) {
          // live block d -> 72 -> 48 -> 32 (half-storage DMMA kernels, 2 / 4 / 5 matrices per SM) -> done (warp per matrix);
          // a 64 phase on 2 x 2-tile superblocks in between was slower (11.4 vs 11.0 ms at C5)
          cplx *bufs[2] = {ws.Q[buf], ws.Q[buf] + (size_t)n * 72 * 72};
          int cur = d, koff = 0, ib = 0, nl = 0;
          const cplx *in = Ain;
          bool first = true;
          for (;;) {
            const int nxt = cur > 72 ? 72 : (cur > 48 ? 48 : (cur > 32 ? 32 : 0));
            const int steps = nxt ? cur - nxt : cur;
            cplx *out = nxt ? bufs[ib] : nullptr;
            const cplx *h0 = first ? H0 : nullptr, *zz = first ? Z : nullptr;
            const double *bb = first ? B : nullptr;
            // (the next half-storage phase reads the lower triangle only -- if its 8 x 8 tiles line up with this one's)
            if (cur > 72)
              hql_tridiag_hs_kernel<12, 3><<<g, 32 * HsGeom<12, 3>::NW, 0, st>>>(cur, d, koff, steps, h0, zz, bb, in, dd_, ee_, vp_, ws.vcap, tt_, out, (steps & 7) ? 1 : 0);
            else if (cur > 48)
              hql_tridiag_hs_kernel<9, 3><<<g, 32 * HsGeom<9, 3>::NW, 0, st>>>(cur, d, koff, steps, h0, zz, bb, in, dd_, ee_, vp_, ws.vcap, tt_, out, (steps & 7) ? 1 : 0);
            else if (cur > 32)
              hql_tridiag_hs_kernel<6, 2><<<g, 32 * HsGeom<6, 2>::NW, 0, st>>>(cur, d, koff, steps, h0, zz, bb, in, dd_, ee_, vp_, ws.vcap, tt_, out, 1);
            else if (!first && o.tridiag_warp && o.tridiag_hsw && cur > 8) {  // last phase: warp per matrix, registers only
              launch_tridiag_hsw(cur, n, nullptr, nullptr, nullptr, in, dd_, ee_, vp_, ws.vcap, tt_, d, koff, st);
            } else if (!first && o.tridiag_warp && o.tridiag_fused) {  // last phase: warp per matrix, no block barriers
              const size_t sm = hql_tridiag_warpf_smem(cur);
              cudaFuncSetAttribute(hql_tridiag_warpf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
              hql_tridiag_warpf_kernel<<<(unsigned)((n + TRW_WARPS - 1) / TRW_WARPS), 32 * TRW_WARPS, sm, st>>>(
                  cur, n, nullptr, nullptr, nullptr, in, dd_, ee_, vp_, ws.vcap, tt_, d, koff);
            } else
              hql_tridiag_rw_kernel<32><<<g, 128, 0, st>>>(cur, d, koff, steps, h0, zz, bb, in, dd_, ee_, vp_, ws.vcap, tt_, out);
            ++nl;
            if (!nxt) break;
            koff += steps;
            cur = nxt;
            in = out;
            ib ^= 1;
            first = false;
          }
          *launches += nl - 1;
        } else if (d <= 64) {
          const int k1 = d - 32;
          hql_tridiag_rw_kernel<64><<<g, 256, 0, st>>>(d, d, 0, k1, H0, Z, B, Ain, dd_, ee_, vp_, ws.vcap, tt_, A32);
          hql_tridiag_rw_kernel<32><<<g, 128, 0, st>>>(32, d, k1, 32, nullptr, nullptr, nullptr, A32, dd_, ee_, vp_, ws.vcap, tt_, nullptr);
          ++*launches;
        } else {
          const int k1 = d - 64;
          hql_tridiag_rw_kernel<96><<<g, 384, 0, st>>>(d, d, 0, k1, H0, Z, B, Ain, dd_, ee_, vp_, ws.vcap, tt_, A64);
          hql_tridiag_rw_kernel<64><<<g, 256, 0, st>>>(64, d, k1, 32, nullptr, nullptr, nullptr, A64, dd_, ee_, vp_, ws.vcap, tt_, A32);
          hql_tridiag_rw_kernel<32><<<g, 128, 0, st>>>(32, d, k1 + 32, 32, nullptr, nullptr, nullptr, A32, dd_, ee_, vp_, ws.vcap, tt_, nullptr);
          *launches += 2;
        }
      }
    } else if (Ain) {
      e = cudaFuncSetAttribute(hql_tridiag_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      hql_tridiag_kernel<false><<<(unsigned)n, g.nth, smem, st>>>(d, g.R, g.G, H0, Z, B, Ain, ws.dbuf[buf], ws.ebuf[buf], ws.Q[buf],
                                                                  o.use_reflect(d) ? ws.Vp[buf] : nullptr, ws.vcap, ws.tauv[buf]);
    } else {
      e = cudaFuncSetAttribute(hql_tridiag_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      hql_tridiag_kernel<true><<<(unsigned)n, g.nth, smem, st>>>(d, g.R, g.G, H0, Z, B, Ain, ws.dbuf[buf], ws.ebuf[buf], ws.Q[buf],
                                                                 o.use_reflect(d) ? ws.Vp[buf] : nullptr, ws.vcap, ws.tauv[buf]);
    }
    ++*launches;
    return (int)cudaGetLastError();
  }
  // Jacobi
  const bool vglob = EighWs::jacobi_vglobal(d);
  const size_t smem = eigh_jacobi_smem(d, vglob);
  if (smem > MUSIM_MAX_SMEM_OPTIN) return -5;
  ProfScope ps(prof, st, PH_EIGH_JACOBI);
  int nth = 256;
  if (d * d <= 64)
    nth = 32;
  else if (d * d <= 256)
    nth = 64;
  else if (d * d <= 1024)
    nth = 128;
  if (Ain) {
    e = cudaFuncSetAttribute(eigh_jacobi_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    eigh_jacobi_kernel<false><<<(unsigned)n, nth, smem, st>>>(d, H0, Z, B, Ain, lam, U, status, 40, vglob ? ws.Vg : nullptr);
  } else {
    e = cudaFuncSetAttribute(eigh_jacobi_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    eigh_jacobi_kernel<true><<<(unsigned)n, nth, smem, st>>>(d, H0, Z, B, Ain, lam, U, status, 40, vglob ? ws.Vg : nullptr);
  }
  ++*launches;
  return (int)cudaGetLastError();
}

// Stage B (Householder+QL only): QL on (d, e), rotation replay, back-transformation U = Q Zt.
inline int launch_eigh_stageB(int method, int d, int64_t n, double *lam, cplx *U, EighWs &ws, int buf, int *status,
                              cudaStream_t st, int64_t *launches, Profiler *prof, bool sorted, const EighOpts &o) {
  if (method != EIGH_HQL) return 0;
  cudaError_t e;
  const bool use_tdc = o.tdc && o.use_reflect(d) && d > 32 && d <= 96;
  if (use_tdc) {
    ProfScope ps(prof, st, PH_EIGH_TDC);
    // leaves: prepare (scale, tear, pad to 24) -> batched QL (thread per leaf) -> rotation replay (thread per row)
    const int64_t nl = 4 * n;
    tdc_prep_kernel<<<(unsigned)((n + 3) / 4), 128, 0, st>>>(d, n, ws.dbuf[buf], ws.ebuf[buf], ws.l_hdr, ws.l_d, ws.l_e);
    {
      const size_t sm = hql_tql_smem(24, 32);
      cudaFuncSetAttribute(hql_tql_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      hql_tql_kernel<32><<<(unsigned)((nl + 31) / 32), 32, sm, st>>>(24, nl, ws.l_d, ws.l_e, ws.l_lam, ws.l_perm, ws.l_rot,
                                                                       EighWs::L_ROT_CAP, ws.l_swp, EighWs::L_SWP_CAP, ws.l_nswp,
                                                                       status, 0, 24);
      const size_t rsmem = hql_apply_reg_smem(24, 24, EighWs::L_SWP_CAP);
      cudaFuncSetAttribute(hql_apply_reg_kernel<24, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem);
      hql_apply_reg_kernel<24, 24><<<dim3((unsigned)nl, 1), 24, rsmem, st>>>(24, ws.l_rot, EighWs::L_ROT_CAP, ws.l_swp,
                                                                              EighWs::L_SWP_CAP, ws.l_nswp, ws.l_Z);
    }
    if (d <= 64) {
      const size_t sm = TdcGeom<64>::smem_bytes;
      cudaFuncSetAttribute(tdc_merge_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      tdc_merge_kernel<64><<<(unsigned)n, TdcGeom<64>::NT, sm, st>>>(d, ws.l_hdr, ws.l_lam, ws.l_Z, lam, ws.Zt);
    } else {
      const size_t sm = TdcGeom<96>::smem_bytes;
      cudaFuncSetAttribute(tdc_merge_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      tdc_merge_kernel<96><<<(unsigned)n, TdcGeom<96>::NT, sm, st>>>(d, ws.l_hdr, ws.l_lam, ws.l_Z, lam, ws.Zt);
    }
    *launches += 2;
    *launches += 2;
  }
  if (!use_tdc) {
    ProfScope ps(prof, st, PH_EIGH_TQL);
    int nt = o.tql_threads ? o.tql_threads : (d <= 32 ? 32 : 16);  // measured: C3 (d = 24) 12.0 -> 10.2 ms per 10^6 with 32
    while (nt > 8 && hql_tql_smem(d, nt) > 200 * 1024) nt >>= 1;  // (d, e) of nt matrices per CTA in shared memory
    const int dpad = (!sorted && d <= 96) ? (d <= 16 && o.small24 ? 16 : (d <= 24 && o.small24 ? 24 : (d <= 32 ? 32 : (d <= 64 ? 64 : 96)))) : 0;  // register replay kernel follows
    const unsigned tb = (unsigned)((n + nt - 1) / nt);
    const size_t sm = hql_tql_smem(d, nt);
#define TQL_LAUNCH(NT)                                                                                       \
  {                                                                                                          \
    e = cudaFuncSetAttribute(hql_tql_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);      \
    if (e != cudaSuccess) return (int)e;                                                                     \
    cudaFuncSetAttribute(hql_tql_kernel<NT>, cudaFuncAttributePreferredSharedMemoryCarveout,                 \
                         cudaSharedmemCarveoutMaxShared);                                                    \
    hql_tql_kernel<NT><<<tb, NT, sm, st>>>(d, n, ws.dbuf[buf], ws.ebuf[buf], lam, ws.perm, ws.rot, ws.rot_cap, \
                                           ws.swp, ws.swp_cap, ws.nswp, status, sorted ? 1 : 0, dpad);       \
  }
    if (nt == 8) TQL_LAUNCH(8) else if (nt == 16) TQL_LAUNCH(16) else TQL_LAUNCH(32)
#undef TQL_LAUNCH
  }
  if (!use_tdc) ++*launches;
  const size_t zsmem = hql_apply_smem(d, ws.swp_cap);
  if (!use_tdc && d <= HQL_MAX_D) {
    e = cudaFuncSetAttribute(hql_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zsmem);
    if (e != cudaSuccess) return (int)e;
  }
  const int ath = std::min(128, (d + 31) & ~31);
  if (use_tdc) {
    // eigenvectors of T already in ws.Zt
  } else if (!sorted && d <= 96) {
    // register-resident rows (static column indices): D = d rounded up to 32 / 64 / 96
    const int D = d <= 16 && o.small24 ? 16 : (d <= 24 && o.small24 ? 24 : (d <= 32 ? 32 : (d <= 64 ? 64 : 96)));
    const int nth = D < 32 ? D : (o.apply_warp ? 32 : D);  // option "apply_warp": one warp (32 rows of Z) per CTA
    const size_t rsmem = hql_apply_reg_smem(D, nth, ws.swp_cap);
    const dim3 grid((unsigned)n, D / nth);
    ProfScope ps(prof, st, PH_EIGH_APPLY);
#define APPLY_LAUNCH(DD, NTH)                                                                                   \
  {                                                                                                             \
    cudaFuncSetAttribute(hql_apply_reg_kernel<DD, NTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem); \
    hql_apply_reg_kernel<DD, NTH><<<grid, NTH, rsmem, st>>>(d, ws.rot, ws.rot_cap, ws.swp, ws.swp_cap, ws.nswp, ws.Zt); \
  }
    if (D == 16) APPLY_LAUNCH(16, 16)
    else if (D == 24) APPLY_LAUNCH(24, 24)
    else if (D == 32) APPLY_LAUNCH(32, 32)
    else if (D == 64 && nth == 32) APPLY_LAUNCH(64, 32)
    else if (D == 64) APPLY_LAUNCH(64, 64)
    else if (nth == 32) APPLY_LAUNCH(96, 32)
    else APPLY_LAUNCH(96, 96)
#undef APPLY_LAUNCH
  } else if (d > HQL_MAX_D) {
    ProfScope ps(prof, st, PH_EIGH_APPLY);
    const int rb = hql_apply_rows_rb(d);
    const size_t rsm = hql_apply_rows_smem(d, rb);
    e = cudaFuncSetAttribute(hql_apply_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsm);
    if (e != cudaSuccess) return (int)e;
    const dim3 grid((unsigned)n, (unsigned)((d + rb - 1) / rb));
    hql_apply_rows_kernel<<<grid, rb, rsm, st>>>(d, ws.rot, ws.rot_cap, ws.swp, ws.swp_cap, ws.nswp, ws.perm, ws.Zt);
  } else {
    ProfScope ps(prof, st, PH_EIGH_APPLY);
    hql_apply_kernel<<<(unsigned)n, ath, zsmem, st>>>(d, ws.rot, ws.rot_cap, ws.swp, ws.swp_cap, ws.nswp, ws.perm, ws.Zt);
  }
  if (!use_tdc) ++*launches;
  {
    ProfScope ps(prof, st, PH_EIGH_BACK);
    if (o.use_reflect(d) && o.back_wy && (d > 32 || (o.back_wy_small && d > 16))) {
      if (d <= 24) {
        const size_t sm = BackWyGeom<24>::smem_bytes;
        cudaFuncSetAttribute(hql_backwy_kernel<24, 1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        hql_tfactor_kernel<24><<<(unsigned)n, 96, BackWyGeom<24>::tf_smem_bytes, st>>>(d, ws.Vp[buf], ws.vcap, ws.tauv[buf], ws.Timg);
        hql_backwy_kernel<24, 1, 8><<<(unsigned)n, 96, sm, st>>>(d, ws.Zt, ws.Vp[buf], ws.vcap, ws.Timg, U);
      } else if (d <= 32) {
        const size_t sm = BackWyGeom<32>::smem_bytes;
        cudaFuncSetAttribute(hql_backwy_kernel<32, 1, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        hql_tfactor_kernel<32><<<(unsigned)n, 128, BackWyGeom<32>::tf_smem_bytes, st>>>(d, ws.Vp[buf], ws.vcap, ws.tauv[buf], ws.Timg);
        hql_backwy_kernel<32, 1, 6><<<(unsigned)n, 128, sm, st>>>(d, ws.Zt, ws.Vp[buf], ws.vcap, ws.Timg, U);
      } else if (d <= 64) {
        const size_t sm = BackWyGeom<64>::smem_bytes;
        cudaFuncSetAttribute(hql_backwy_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        cudaFuncSetAttribute(hql_tfactor_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BackWyGeom<64>::tf_smem_bytes);
        hql_tfactor_kernel<64><<<(unsigned)n, 256, BackWyGeom<64>::tf_smem_bytes, st>>>(d, ws.Vp[buf], ws.vcap, ws.tauv[buf], ws.Timg);
        hql_backwy_kernel<64, 2><<<(unsigned)(2 * n), 128, sm, st>>>(d, ws.Zt, ws.Vp[buf], ws.vcap, ws.Timg, U);
      } else {
        const size_t sm = BackWyGeom<96>::smem_bytes;
        cudaFuncSetAttribute(hql_backwy_kernel<96, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        cudaFuncSetAttribute(hql_tfactor_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BackWyGeom<96>::tf_smem_bytes);
        hql_tfactor_kernel<96><<<(unsigned)n, 384, BackWyGeom<96>::tf_smem_bytes, st>>>(d, ws.Vp[buf], ws.vcap, ws.tauv[buf], ws.Timg);
        hql_backwy_kernel<96, 2><<<(unsigned)(2 * n), 192, sm, st>>>(d, ws.Zt, ws.Vp[buf], ws.vcap, ws.Timg, U);
      }
      ++*launches;
    } else if (o.use_reflect(d)) {
      if (d <= 32) {
        const size_t sm = hql_reflect_smem(32);
        if (d <= 16 && o.small24) {
          const size_t sm16 = hql_reflect_smem(16);
          cudaFuncSetAttribute(hql_reflect_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm16);
          hql_reflect_kernel<16, 1><<<(unsigned)n, 16, sm16, st>>>(d, ws.Zt, ws.Vp[buf], ws.vcap, ws.tauv[buf], U);
        } else if (d <= 24 && o.small24) {  // D = 24 instantiation: 24 rows per thread instead of 32 (fewer registers, more warps)
          const size_t sm24 = hql_reflect_smem(24);
          cudaFuncSetAttribute(hql_reflect_kernel<24, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm24);
          hql_reflect_kernel<24, 1><<<(unsigned)n, 24, sm24, st>>>(d, ws.Zt, ws.Vp[buf], ws.vcap, ws.tauv[buf], U);
        } else if (d <= 24) {  // one thread per column (warp per matrix): the dot products are thread-local
          cudaFuncSetAttribute(hql_reflect_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
          hql_reflect_kernel<32, 1><<<(unsigned)n, 32, sm, st>>>(d, ws.Zt, ws.Vp[buf], ws.vcap, ws.tauv[buf], U);
        } else {
          cudaFuncSetAttribute(hql_reflect_kernel<32, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
          hql_reflect_kernel<32, 4><<<(unsigned)n, 128, sm, st>>>(d, ws.Zt, ws.Vp[buf], ws.vcap, ws.tauv[buf], U);
        }
      } else if (d <= 64) {
        const size_t sm = hql_reflect_smem(64);
        cudaFuncSetAttribute(hql_reflect_kernel<64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        hql_reflect_kernel<64, 4><<<(unsigned)n, 256, sm, st>>>(d, ws.Zt, ws.Vp[buf], ws.vcap, ws.tauv[buf], U);
      } else {
        const size_t sm = hql_reflect_smem(96);
        cudaFuncSetAttribute(hql_reflect_kernel<96, 8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        hql_reflect_kernel<96, 8, 2><<<(unsigned)n, 384, sm, st>>>(d, ws.Zt, ws.Vp[buf], ws.vcap, ws.tauv[buf], U);
      }
    } else {
      dim3 grid((d + 31) / 32, (d + 31) / 32, (unsigned)n);
      cgemm_realB_kernel<<<grid, 256, 0, st>>>(d, ws.Q[buf], ws.Zt, U);
    }
  }
  ++*launches;
  return (int)cudaGetLastError();
}

// Eigen-decompose n matrices: either H0 + B.Z (Ain == nullptr) or Ain[n].  lam ascending when
// `sorted`, U row-major with eigenvectors in columns.  Returns 0, a cudaError_t (> 0), or -5.
inline int launch_eigh(int method, int d, int64_t n, const cplx *H0, const cplx *Z, const double *B,
                       const cplx *Ain, double *lam, cplx *U, EighWs &ws, int *status, cudaStream_t st,
                       int64_t *launches, Profiler *prof, bool sorted = true, const EighOpts &o = EighOpts()) {
  int64_t dummy = 0;
  if (!launches) launches = &dummy;
  cudaError_t e = ws.ensure(method, d, n);
  if (e != cudaSuccess) return (int)e;
  int rc = launch_eigh_stageA(method, d, n, H0, Z, B, Ain, lam, U, ws, 0, status, st, launches, prof, o);
  if (rc) return rc;
  return launch_eigh_stageB(method, d, n, lam, U, ws, 0, status, st, launches, prof, sorted, o);
}

}  // namespace musim
