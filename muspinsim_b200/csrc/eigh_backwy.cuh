// eigh_backwy.cuh -- K4 on the FP64 tensor pipe: back-transformation U = Q Zt of the Hermitian
// eigensolver (second half of np.linalg.eigh, /root/reference/muspinsim/spinop.py:69; LAPACK
// zunmtr) with the Householder reflectors applied in COMPACT-WY BLOCKS of 8,
//
//     Q = Q_0 Q_1 ... Q_{NB-1},   Q_b = H_{8b} ... H_{8b+7} = I - V_b T_b V_b^H      (zlarft, forward / columnwise)
//     C <- C - V_b (T_b (V_b^H C))                                                   (zlarfb), b = NB-1 .. 0, C = Zt at the start
//
// so that every O(d^3) flop is a DMMA (mma.sync m8n8k4 f64).  The level-2 kernel it replaces
// (hql_reflect_kernel: one reflector at a time, dot product + axpy per column) ran one dependency
// chain per SM at 45 % of the FP64 pipe (profiles/r1_ncu_full_summary.md): 12.6 ms at C5.
//
// Mapping.  A warp owns columns 8w .. 8w+7 of C for ALL rows and keeps them in registers as the
// accumulator fragments of C^T (tile t = rows 8t .. 8t+7):
//     accumulator (m = lane/4, n = 2 (lane%4) + e)   <->   C[row 8t + n][column 8w + m].
// With C held transposed, all three products of a block are warp-local AND shuffle-free, because an
// accumulator fragment read slot by slot IS an A-operand fragment whose reduction index is
// permuted (slot e covers n = e, 2 + e, 4 + e, 6 + e):
//   1. W^T  = C^T conj(V_b)      M = column, N = reflector, K = row        A = C^T accumulators
//   2. W2^T = W^T T_b^T          M = column, N = reflector, K = reflector  A = W^T accumulators
//   3. C^T -= W2^T V_b^T         M = column, N = row,       K = reflector  A = W2^T accumulators
// Only the B operands come from shared memory.  V is staged once per CTA as planar re / im arrays,
// block b as Vs_b[reflector 0..7][row' 0 .. D - 8b) with a leading dimension = 4 (mod 16) and the
// rows of every group of 8 stored in the order pos = (0, 6, 1, 7, 2, 4, 3, 5): with that permutation
// the fragment loads of step 1 (rows 2j + e for j = lane%4) and of step 3 (rows lane/4, reflectors
// 4e + j) are both bank-conflict free.  The output columns of step 2 are assigned to reflectors
// 4e + j for the same reason (T is stored in the matching order).
//
// Two CTAs per matrix (HALVES = 2: each takes half of the column tiles), two CTAs per SM: the first
// version ran one 12-warp CTA per SM and ncu showed 17 % of the stall samples in the staging of V and
// 12 % in the T factors with the DMMA pipe idle; with two independent CTAs on an SM one stages while
// the other computes.  The T factors come from a separate, high-occupancy kernel:
//
// hql_tfactor_kernel: warp b forms the Gram matrix G = V_b^H V_b with DMMAs (A and B fragment are
// the same loaded value), then lane l < 8 runs row l of the zlarft recurrence
//     T[l][l] = tau_l,   T[l][i] = -tau_i sum_{q=l}^{i-1} T[l][q] G[q][i]    (rows are independent)
// and writes T in the operand order of step 2.
#pragma once
#include "common.cuh"
#include "eigh_hql.cuh"  // cp_async16
#include "polar.cuh"    // dmma884

namespace musim {

template <int D>
struct BackWyGeom {
  static constexpr int NB = D / 8;  // reflector blocks = row tiles = column tiles
  static constexpr int TLD = 12;    // leading dimension of the 8 x 8 T operand tiles
  static constexpr int TIMG = 2 * NB * 8 * TLD;  // doubles of the T image of one matrix (re plane, im plane)
  __host__ __device__ static constexpr int ld(int b) { return ((D - 8 * b + 15) & ~15) + 4; }  // = 4 (mod 16)
  __host__ __device__ static constexpr int voff(int b) {  // offset of block b in a plane (doubles)
    int o = 0;
    for (int q = 0; q < b; ++q) o += 8 * ld(q);
    return o;
  }
  static constexpr int VPLANE = voff(NB);
  static constexpr size_t smem_bytes = (size_t)(2 * VPLANE + TIMG) * sizeof(double);
  static constexpr size_t tf_smem_bytes = (size_t)((D - 1) * (D - 2) / 2 + 8) * sizeof(cplx);  // >= NB * 64 Gram entries
};

// ---------------------------------------------------------------------------------------
// T factors of all blocks of one matrix: one CTA (NB warps) per matrix.  The packed reflectors
// (71 KB at d = 96, contiguous) are staged into shared memory with 16-byte cp.async first: the
// version that gathered them straight from global memory (8 x 64-byte segments per load
// instruction, each warp waiting on its own dependent chain) ran 1.13 ms at C5, long_scoreboard
// 15 warps per issue (profiles/r2_ncu_k4.md).
// ---------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(4 * D, (D <= 32 ? 5 : 2))
hql_tfactor_kernel(int d, const cplx *__restrict__ Vp, size_t vcap, const cplx *__restrict__ tau,
                   double *__restrict__ Timg) {
  using G = BackWyGeom<D>;
  constexpr int NB = G::NB, TLD = G::TLD, NT = 4 * D;
  extern __shared__ __align__(16) unsigned char tf_smem[];
  cplx *sV = reinterpret_cast<cplx *>(tf_smem);  // packed reflectors; the Gram matrices re-use the front afterwards
  const int tid = threadIdx.x, lane = tid & 31, b = tid >> 5;
  const int fm = lane >> 2, fj = lane & 3;
  const size_t mat = blockIdx.x;
  {
    const int total = (d - 1) * (d - 2) / 2;
    const cplx *src = Vp + mat * vcap;
    for (int e = tid; e < total; e += NT) cp_async16(&sV[e], &src[e]);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
  }
  const int i = 8 * b + fm;  // this lane's reflector
  const bool valid = i < d - 1;
  const int mk = d - i - 2;
  // v_i[r] for r > i + 1 sits at base[r]
  const cplx *base = sV + (valid ? mk * (mk - 1) / 2 : 0) - (i + 2);
  double gr[2][2][2] = {}, gi[2][2][2] = {}, hr[2][2][2] = {}, hi[2][2][2] = {};  // [t parity][e][slot]: short dependent chains
#define TF_TILE(T_, TP_)                                                     \
  _Pragma("unroll") for (int e = 0; e < 2; ++e) {                            \
    const int r = 8 * (T_) + 4 * e + fj;                                     \
    cplx v = make_c(0.0, 0.0);                                               \
    if (valid && r < d) {                                                    \
      if (r > i + 1)                                                         \
        v = base[r];                                                         \
      else if (r == i + 1)                                                   \
        v = make_c(1.0, 0.0);                                                \
    }                                                                        \
    dmma884(gr[TP_][e][0], gr[TP_][e][1], v.x, v.x);                         \
    dmma884(hr[TP_][e][0], hr[TP_][e][1], v.y, v.y);                         \
    dmma884(gi[TP_][e][0], gi[TP_][e][1], v.x, v.y);                         \
    dmma884(hi[TP_][e][0], hi[TP_][e][1], -v.y, v.x);                        \
  }
  for (int t = b; t < NB; t += 2) {
    TF_TILE(t, 0)
    if (t + 1 < NB) {
      TF_TILE(t + 1, 1)
    }
  }
#undef TF_TILE
  __syncthreads();  // every warp is done with the staged reflectors
  cplx *gs = sV + b * 64;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const double re = (gr[0][0][s] + gr[0][1][s]) + (gr[1][0][s] + gr[1][1][s]) + (hr[0][0][s] + hr[0][1][s]) + (hr[1][0][s] + hr[1][1][s]);
    const double im = (gi[0][0][s] + gi[0][1][s]) + (gi[1][0][s] + gi[1][1][s]) + (hi[0][0][s] + hi[0][1][s]) + (hi[1][0][s] + hi[1][1][s]);
    gs[fm * 8 + 2 * fj + s] = make_c(re, im);
  }
  __syncwarp();
  double *Tre = Timg + mat * G::TIMG, *Tim = Tre + NB * 8 * TLD;
  if (lane < 8) {
    const int l = lane;
    cplx T[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const cplx tc = (8 * b + c < d - 1) ? tau[mat * d + 8 * b + c] : make_c(0.0, 0.0);
      cplx acc = make_c(0.0, 0.0);
#pragma unroll
      for (int q = 0; q < c; ++q)
        if (q >= l) cfma(acc, T[q], gs[q * 8 + c]);
      const cplx off = cmul(make_c(-tc.x, -tc.y), acc);
      T[c] = (c == l) ? tc : ((c > l) ? off : make_c(0.0, 0.0));
    }
    // operand order of step 2: B[k = (j, e) <-> reflector 2j + e][n = u' <-> reflector 4 (u' % 2) + u' / 2] = T[refl(u')][2j + e]
    const int up = 2 * (l & 3) + (l >> 2);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      Tre[(b * 8 + 4 * (c & 1) + (c >> 1)) * TLD + up] = T[c].x;
      Tim[(b * 8 + 4 * (c & 1) + (c >> 1)) * TLD + up] = T[c].y;
    }
  } else if (lane < 12) {  // the padding columns of the image (never read by a fragment load, copied by the consumer)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      Tre[(b * 8 + c) * TLD + lane] = 0.0;
      Tim[(b * 8 + c) * TLD + lane] = 0.0;
    }
  }
}

// ---------------------------------------------------------------------------------------
// The back-transformation proper.  HALVES CTAs per matrix, NB / HALVES warps each.
// ---------------------------------------------------------------------------------------
template <int D, int HALVES, int MINB = HALVES>
__global__ void __launch_bounds__(4 * D / HALVES, MINB)
hql_backwy_kernel(int d, const double *__restrict__ Zt, const cplx *__restrict__ Vp, size_t vcap,
                  const double *__restrict__ Timg, cplx *__restrict__ U) {
  using G = BackWyGeom<D>;
  constexpr int NB = G::NB, TLD = G::TLD, NT = 4 * D / HALVES, NW = NB / HALVES;
  static_assert(D % 8 == 0 && D <= 96 && NB % HALVES == 0, "D = 24, 32, 64 or 96");
  extern __shared__ __align__(16) unsigned char bw_smem[];
  double *Vre = reinterpret_cast<double *>(bw_smem);  // blocks b = 0 .. NB-1, [8][ld(b)] each
  double *Vim = Vre + G::VPLANE;
  double *Tre = Vim + G::VPLANE;                       // [NB][8][TLD]
  double *Tim = Tre + NB * 8 * TLD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fm = lane >> 2, fj = lane & 3;
  const size_t mat = blockIdx.x / HALVES;
  const int ctile = (blockIdx.x % HALVES) * NW + warp;  // this warp's column tile
  const size_t dd = (size_t)d * d;
  const cplx *myv = Vp + mat * vcap;

  // ---- C^T tiles of this warp's 8 columns (loads in flight while V is staged) ----
  double cr[NB][2], ci[NB][2];
  {
    const int n = 8 * ctile + fm;
#pragma unroll
    for (int t = 0; t < NB; ++t)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int r = 8 * t + 2 * fj + e;
        cr[t][e] = (r < d && n < d) ? Zt[mat * dd + (size_t)r * d + n] : 0.0;
        ci[t][e] = 0.0;
      }
  }
  // ---- stage V (unit diagonal and zeros made explicit) and the T image ----
  {
    const double *tg = Timg + mat * G::TIMG;
    for (int q = tid; q < G::TIMG / 2; q += NT) reinterpret_cast<double2 *>(Tre)[q] = reinterpret_cast<const double2 *>(tg)[q];
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    constexpr int dummy = 0;
    (void)dummy;
    const int wb = D - 8 * b;          // rows 8b .. D-1
    const int cnt = 8 * wb;            // elements of the block
    const int ldb = G::ld(b), ob = G::voff(b);
    constexpr int MAXB = (8 * D + NT - 1) / NT;
    cplx v[MAXB];
    int dst[MAXB];
#pragma unroll
    for (int u = 0; u < MAXB; ++u) {
      const int idx = tid + u * NT;
      dst[u] = -1;
      v[u] = make_c(0.0, 0.0);
      if (idx < cnt) {
        const int il = idx / wb, c = idx - il * wb;  // reflector 8b + il, row 8b + c
        const int i = 8 * b + il, r = 8 * b + c;
        if (i < d - 1 && r < d) {
          if (r > i + 1) {
            const int mk = d - i - 2;
            v[u] = myv[(size_t)mk * (mk - 1) / 2 + (r - i - 2)];
          } else if (r == i + 1) {
            v[u] = make_c(1.0, 0.0);
          }
        }
        dst[u] = ob + il * ldb + (c & ~7) + (int)((0x53427160u >> (4 * (c & 7))) & 7);
      }
    }
#pragma unroll
    for (int u = 0; u < MAXB; ++u)
      if (dst[u] >= 0) {
        Vre[dst[u]] = v[u].x;
        Vim[dst[u]] = v[u].y;
      }
  }
  __syncthreads();

  const int p3 = (0x53427160u >> (4 * fm)) & 7;  // row position of step 3's N index
#pragma unroll 1
  for (int b = NB - 1; b >= 0; --b) {
    const int ldb = G::ld(b);
    const double *vr0 = Vre + G::voff(b), *vi0 = Vim + G::voff(b);
    // 1. W^T = C^T conj(V_b): two accumulator sets (the two slots) halve the dependent DMMA chain
    double wr[2] = {0.0, 0.0}, wi[2] = {0.0, 0.0}, xr[2] = {0.0, 0.0}, xi[2] = {0.0, 0.0};
    {
      const double *vrb = vr0 + fm * ldb - 8 * b, *vib = vi0 + fm * ldb - 8 * b;
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
          const double vr = vr0[(4 * e + fj) * ldb + 8 * (t - b) + p3], vi = vi0[(4 * e + fj) * ldb + 8 * (t - b) + p3];
          dmma884(cr[t][0], cr[t][1], nyr[e], vr);
          dmma884(ci[t][0], ci[t][1], nyr[e], vi);
          dmma884(cr[t][0], cr[t][1], yi[e], vi);
          dmma884(ci[t][0], ci[t][1], nyi[e], vr);
        }
      }
    }
  }

  {
    const int n = 8 * ctile + fm;
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
