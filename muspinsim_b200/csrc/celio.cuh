// celio.cuh -- Celio's method (Trotter-split state-vector evolution) on the GPU, batched over the
// random initial states.  Replaces the reference's C++ extension for this path:
//   celio::evolve              /root/reference/muspinsim/cpp/celio.cpp:23-69     (time-step loop)
//   parallel::fast_evolve_ptr  /root/reference/muspinsim/cpp/parallel.cpp:236-266 (V <- (M (x) 1) V through an index map)
//   parallel::fast_measure_h_ptr                          parallel.cpp:130-168   (V^H (sigma (x) 1) V)
// called from CelioHamiltonian._fast_evolve_cpp (celio.py:433-476), which loops over `averages`
// random states one at a time; here every state is a CTA (or a slice of the grid) of ONE launch.
//
// A gate is a small dense matrix M (md x md: the exponential of one Hamiltonian contribution on the
// spins it couples) acting on groups of md amplitudes picked by an index map (the reference's swap
// trick: indices = transpose(arange(dim).reshape(dims), spin_order), celio.py:172-187):
//     for g < od = dim / md:   V[idx[i od + g]] <- sum_j M[i][j] V[idx[j od + g]]
// Two execution paths:
//   resident  dim <= 12 288 amplitudes (192 KB): one CTA per state keeps the state vector in SHARED
//             MEMORY for the whole run -- all nt x k x n_gates gate applications and the nt
//             measurements are one kernel, the state never touches HBM (examples/celio: mu + 4 x 51V,
//             dim = 8 192);
//   streamed  larger systems: the states live in global memory, one launch per gate over all states.
#pragma once
#include "common.cuh"

namespace musim {

#define CELIO_MAX_MD 64        // largest gate (two spins 7/2)
#define CELIO_SMEM_DIM 12288   // amplitudes of the shared-memory resident path

struct CelioGate {
  const cplx *M;        // md x md, row-major (device)
  const int *idx;       // [dim] (device)
  int md;
  long long od;
};

// one group g of one gate on the vector V (shared or global); MD > 0: gate size known at compile
// time (amplitudes in registers), MD == 0: any size up to CELIO_MAX_MD (local array)
template <int MD>
__device__ void celio_apply_group(cplx *V, const cplx *__restrict__ M, const int *__restrict__ idx,
                                                  int md, long long od, long long g) {
  if (MD > 0) {
    constexpr int N = MD > 0 ? MD : 1;
    cplx v[N];
    int at[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      at[j] = idx[(long long)j * od + g];
      v[j] = V[at[j]];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      cplx acc = make_c(0.0, 0.0);
#pragma unroll
      for (int j = 0; j < N; ++j) cfma(acc, M[i * N + j], v[j]);
      V[at[i]] = acc;  // every group touches its own md amplitudes: no hazard between groups
    }
  } else {
    cplx v[CELIO_MAX_MD];
    int at[CELIO_MAX_MD];
    for (int j = 0; j < md; ++j) {
      at[j] = idx[(long long)j * od + g];
      v[j] = V[at[j]];
    }
    for (int i = 0; i < md; ++i) {
      cplx acc = make_c(0.0, 0.0);
      for (int j = 0; j < md; ++j) cfma(acc, M[i * md + j], v[j]);
      V[at[i]] = acc;
    }
  }
}

// the large gates are separate functions (__noinline__): inlined into one kernel body their register
// arrays spilled 7 KB per thread
template <int MD>
__device__ __noinline__ void celio_apply_group_call(cplx *V, const cplx *__restrict__ M, const int *__restrict__ idx,
                                                    int md, long long od, long long g) {
  celio_apply_group<MD>(V, M, idx, md, od, g);
}

__device__ __forceinline__ void celio_apply_dispatch(cplx *V, const CelioGate &gt, long long g) {
  switch (gt.md) {
    case 2: celio_apply_group<2>(V, gt.M, gt.idx, 2, gt.od, g); break;
    case 3: celio_apply_group<3>(V, gt.M, gt.idx, 3, gt.od, g); break;
    case 4: celio_apply_group<4>(V, gt.M, gt.idx, 4, gt.od, g); break;
    case 6: celio_apply_group<6>(V, gt.M, gt.idx, 6, gt.od, g); break;
    case 8: celio_apply_group_call<8>(V, gt.M, gt.idx, 8, gt.od, g); break;
    case 16: celio_apply_group_call<16>(V, gt.M, gt.idx, 16, gt.od, g); break;
    default: celio_apply_group_call<0>(V, gt.M, gt.idx, gt.md, gt.od, g); break;
  }
}

// V^H (sigma (x) 1_half) V for a Hermitian 2 x 2 sigma: partial sum of this thread's groups
__device__ __forceinline__ double celio_measure_term(const cplx v0, const cplx v1, const cplx s00, const cplx s01,
                                                     const cplx s11) {
  // s00 |v0|^2 + s11 |v1|^2 + 2 Re(conj(v0) s01 v1)      (parallel.cpp:151-163: upper triangle doubled)
  const cplx t = cmul(s01, v1);
  return s00.x * cnorm2(v0) + s11.x * cnorm2(v1) + 2.0 * (v0.x * t.x + v0.y * t.y);
}

// ---- resident path: one CTA per state ----
__global__ void __launch_bounds__(512)
celio_resident_kernel(long long dim, long long half, const cplx *__restrict__ psi, cplx s00, cplx s01, cplx s11,
                      int k, int n_gates, const CelioGate *__restrict__ gates, int nt, double *__restrict__ results) {
  extern __shared__ __align__(16) unsigned char celio_smem[];
  cplx *V = reinterpret_cast<cplx *>(celio_smem);
  __shared__ double red[32];
  const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, wid = tid >> 5;
  const cplx *src = psi + (long long)blockIdx.x * dim;
  for (long long i = tid; i < dim; i += nth) V[i] = src[i];
  __syncthreads();
  for (int t = 0; t < nt; ++t) {
    double acc = 0.0;
    for (long long g = tid; g < half; g += nth) acc += celio_measure_term(V[g], V[half + g], s00, s01, s11);
    acc = warp_sum(acc);
    if (lane == 0) red[wid] = acc;
    __syncthreads();
    if (wid == 0) {
      double v = lane < (nth >> 5) ? red[lane] : 0.0;
      v = warp_sum(v);
      if (lane == 0) atomicAdd(&results[t], v);
    }
    __syncthreads();
    if (t == nt - 1) break;  // the reference evolves once more after the last measurement; the result is unused
    for (int rep = 0; rep < k; ++rep)
      for (int c = 0; c < n_gates; ++c) {
        const CelioGate gt = gates[c];
        for (long long g = tid; g < gt.od; g += nth) celio_apply_dispatch(V, gt, g);
        __syncthreads();
      }
  }
}

// ---- streamed path: states in global memory ----
__global__ void __launch_bounds__(256)
celio_gate_kernel(long long dim, int n_states, cplx *__restrict__ psi, CelioGate gt) {
  const long long tot = gt.od * (long long)n_states;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < tot; w += (long long)gridDim.x * blockDim.x) {
    const long long s = w / gt.od, g = w - s * gt.od;
    celio_apply_dispatch(psi + s * dim, gt, g);
  }
}

__global__ void __launch_bounds__(256)
celio_measure_kernel(long long dim, long long half, int n_states, const cplx *__restrict__ psi, cplx s00, cplx s01,
                     cplx s11, double *__restrict__ out, const int *__restrict__ step) {
  __shared__ double red[8];
  if (step) out += *step;  // replayed CUDA graph: the time index lives on the device (celio_tick_kernel)
  const long long tot = half * (long long)n_states;
  double acc = 0.0;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < tot; w += (long long)gridDim.x * blockDim.x) {
    const long long s = w / half, g = w - s * half;
    const cplx *V = psi + s * dim;
    acc += celio_measure_term(V[g], V[half + g], s00, s01, s11);
  }
  acc = warp_sum(acc);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) red[wid] = acc;
  __syncthreads();
  if (wid == 0) {
    double v = lane < 8 ? red[lane] : 0.0;
    v = warp_sum(v);
    if (lane == 0) atomicAdd(out, v);
  }
}

__global__ void celio_tick_kernel(int *step) { ++*step; }

// Host driver.  All pointers are HOST pointers; results[nt] is accumulated into (+=), summed over
// the states, exactly like repeated calls of the reference's celio_evolve on one results array.
// Returns 0, -1 (invalid), -2 (CUDA error), -5 (gate too large).
inline int celio_evolve_host(int device, long long dim, int n_states, const double *psi, const double *sigma,
                             long long half, int k, int n_gates, const int *mat_dim, const long long *other_dim,
                             const double *matrices, const long long *indices, int nt, double *results,
                             int64_t *launches, int force_streamed) {
  if (dim < 2 || n_states < 1 || !psi || !sigma || half * 2 != dim || k < 1 || n_gates < 0 || nt < 1 || !results ||
      (n_gates > 0 && (!mat_dim || !other_dim || !matrices || !indices)))
    return -1;
  for (int c = 0; c < n_gates; ++c) {
    if (mat_dim[c] < 1 || (long long)mat_dim[c] * other_dim[c] != dim) return -1;
    if (mat_dim[c] > CELIO_MAX_MD) return -5;
  }
  if (dim > 2147483647LL) return -5;
  int prev = -1;
  cudaGetDevice(&prev);
  if (cudaSetDevice(device) != cudaSuccess) return -2;
  cudaError_t e = cudaSuccess;
  cplx *dpsi = nullptr, *dM = nullptr;
  int *didx = nullptr;
  CelioGate *dg = nullptr;
  double *dres = nullptr;
  size_t msum = 0;
  for (int c = 0; c < n_gates; ++c) msum += (size_t)mat_dim[c] * mat_dim[c];
  std::vector<int> idx32((size_t)n_gates * dim);
  for (size_t i = 0; i < idx32.size(); ++i) {
    if (indices[i] < 0 || indices[i] >= dim) return -1;
    idx32[i] = (int)indices[i];
  }
#define CE(call) \
  if (e == cudaSuccess) e = (call)
  CE(dev_malloc((void **)&dpsi, (size_t)n_states * dim * sizeof(cplx)));
  CE(dev_malloc((void **)&dM, std::max<size_t>(1, msum) * sizeof(cplx)));
  CE(dev_malloc((void **)&didx, std::max<size_t>(1, idx32.size()) * sizeof(int)));
  CE(dev_malloc((void **)&dg, std::max(1, n_gates) * sizeof(CelioGate)));
  CE(dev_malloc((void **)&dres, (size_t)nt * sizeof(double)));
  CE(cudaMemcpy(dpsi, psi, (size_t)n_states * dim * sizeof(cplx), cudaMemcpyHostToDevice));
  if (msum) CE(cudaMemcpy(dM, matrices, msum * sizeof(cplx), cudaMemcpyHostToDevice));
  if (!idx32.empty()) CE(cudaMemcpy(didx, idx32.data(), idx32.size() * sizeof(int), cudaMemcpyHostToDevice));
  CE(cudaMemset(dres, 0, (size_t)nt * sizeof(double)));
  std::vector<CelioGate> hg(std::max(1, n_gates));
  {
    size_t mo = 0;
    for (int c = 0; c < n_gates; ++c) {
      hg[c].M = dM + mo;
      hg[c].idx = didx + (size_t)c * dim;
      hg[c].md = mat_dim[c];
      hg[c].od = other_dim[c];
      mo += (size_t)mat_dim[c] * mat_dim[c];
    }
  }
  CE(cudaMemcpy(dg, hg.data(), hg.size() * sizeof(CelioGate), cudaMemcpyHostToDevice));
  const cplx *sg = reinterpret_cast<const cplx *>(sigma);  // row-major 2 x 2
  const cplx s00 = sg[0], s01 = sg[1], s11 = sg[3];
  if (e == cudaSuccess) {
    if (dim <= CELIO_SMEM_DIM && !force_streamed) {
      const size_t sm = (size_t)dim * sizeof(cplx);
      e = cudaFuncSetAttribute(celio_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      if (e == cudaSuccess) {
        const int nth = dim >= 2048 ? 512 : (dim >= 512 ? 256 : 128);
        celio_resident_kernel<<<n_states, nth, sm>>>(dim, half, dpsi, s00, s01, s11, k, n_gates, dg, nt, dres);
        if (launches) ++*launches;
        e = cudaGetLastError();
      }
    } else {
      // Streamed path: one launch per gate over all states, one per measurement.  With small states the time
      // loop is launch bound (nt x (k n_gates + 1) launches of a few microseconds), so ONE time step is captured
      // as a CUDA graph -- measure (time index read from the device), tick, k x n_gates gate kernels -- and
      // replayed nt - 1 times; the last measurement is a plain launch.
      const long long work = (long long)n_states * dim;
      const int blocks = (int)std::min<long long>((work / 2 + 255) / 256, 148LL * 16);
      cudaStream_t cs = nullptr;
      cudaGraph_t graph = nullptr;
      cudaGraphExec_t gexec = nullptr;
      int *dstep = nullptr;
      e = cudaDeviceSynchronize();  // the uploads / memset above ran on the legacy stream; `cs` does not wait for it
      if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
      if (e == cudaSuccess) e = dev_malloc((void **)&dstep, sizeof(int));
      if (e == cudaSuccess) e = cudaMemsetAsync(dstep, 0, sizeof(int), cs);
      if (e == cudaSuccess && nt > 1) {
        e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
          celio_measure_kernel<<<blocks, 256, 0, cs>>>(dim, half, n_states, dpsi, s00, s01, s11, dres, dstep);
          celio_tick_kernel<<<1, 1, 0, cs>>>(dstep);
          for (int rep = 0; rep < k; ++rep)
            for (int c = 0; c < n_gates; ++c) {
              const long long tot = hg[c].od * (long long)n_states;
              const int gb = (int)std::min<long long>((tot + 255) / 256, 148LL * 16);
              celio_gate_kernel<<<gb, 256, 0, cs>>>(dim, n_states, dpsi, hg[c]);
            }
          e = cudaStreamEndCapture(cs, &graph);
        }
        if (e == cudaSuccess) e = cudaGraphInstantiate(&gexec, graph, 0);
        for (int t = 0; t < nt - 1 && e == cudaSuccess; ++t) e = cudaGraphLaunch(gexec, cs);
        if (launches) *launches += (int64_t)(nt - 1) * (2 + (int64_t)k * n_gates);
      }
      if (e == cudaSuccess) {
        celio_measure_kernel<<<blocks, 256, 0, cs>>>(dim, half, n_states, dpsi, s00, s01, s11, dres, dstep);
        if (launches) ++*launches;
        e = cudaGetLastError();
      }
      if (e == cudaSuccess) e = cudaStreamSynchronize(cs);
      if (gexec) cudaGraphExecDestroy(gexec);
      if (graph) cudaGraphDestroy(graph);
      dev_free(dstep);
      if (cs) cudaStreamDestroy(cs);
    }
  }
  std::vector<double> hres(nt);
  CE(cudaMemcpy(hres.data(), dres, (size_t)nt * sizeof(double), cudaMemcpyDeviceToHost));
#undef CE
  dev_free(dpsi);
  dev_free(dM);
  dev_free(didx);
  dev_free(dg);
  dev_free(dres);
  if (prev >= 0 && prev != device) cudaSetDevice(prev);
  if (e != cudaSuccess) return -2;
  for (int t = 0; t < nt; ++t) results[t] += hres[t];
  return 0;
}

}  // namespace musim
