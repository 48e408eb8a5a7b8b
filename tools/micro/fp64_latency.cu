// FP64 latency / issue probe on sm_100a: dependent chains of DFMA with ILP = 1, 2, 4, 8 in one warp,
// 1..4 warps per SM sub-partition.  usage: ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double *out, long long *cyc, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void chain_rsqrt(double *out, long long *cyc, int iters, double a) {
  double x = threadIdx.x + 2.0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) x = rsqrt(x) + a;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
void run(int threads, double *out, long long *cyc) {
  const int iters = 2000;
  chain<ILP><<<1, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  printf("threads %4d ILP %d: %.2f cycles per DFMA-step (per chain), %.2f cycles per warp-instruction per SMSP\n", threads, ILP,
         (double)h / (iters * 16), (double)h / (iters * 16 * ILP * ((threads + 127) / 128)));
}

int main() {
  double *out;
  long long *cyc;
  cudaMalloc(&out, 1 << 20);
  cudaMalloc(&cyc, 8);
  for (int threads : {32, 128, 256, 512}) {
    run<1>(threads, out, cyc);
    run<2>(threads, out, cyc);
    run<4>(threads, out, cyc);
    run<8>(threads, out, cyc);
  }
  chain_rsqrt<<<1, 32>>>(out, cyc, 2000, 1.5);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  printf("rsqrt+add dependent: %.1f cycles\n", (double)h / 8000);
  return 0;
}
