// common.cuh -- shared helpers for the musim sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace musim {

typedef double2 cplx;  // (x = re, y = im), 16-byte aligned -> LDG.128 / LDS.128

__host__ __device__ __forceinline__ cplx make_c(double r, double i) { return make_double2(r, i); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_c(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_c(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  return make_c(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
// a * conj(b)
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) {
  return make_c(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -a.x * b.y));
}
// conj(a) * b
__device__ __forceinline__ cplx ccmul(cplx a, cplx b) {
  return make_c(fma(a.x, b.x, a.y * b.y), fma(a.x, b.y, -a.y * b.x));
}
__device__ __forceinline__ cplx cscale(double s, cplx a) { return make_c(s * a.x, s * a.y); }
__device__ __forceinline__ cplx cconj(cplx a) { return make_c(a.x, -a.y); }
// acc += a*b
__device__ __forceinline__ void cfma(cplx &acc, cplx a, cplx b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}
// acc += conj(a)*b
__device__ __forceinline__ void ccfma(cplx &acc, cplx a, cplx b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(-a.y, b.x, acc.y);
}
__device__ __forceinline__ double cnorm2(cplx a) { return fma(a.x, a.x, a.y * a.y); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; `red` is >= 33 doubles of shared memory.  Result broadcast to all threads.
__device__ __forceinline__ double block_sum(double v, double *red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    double t = (lane < nw) ? red[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// exp(-2*pi*i*x) for x in "cycles", with exact range reduction (|x| may be 1e6).
__device__ __forceinline__ cplx cis_m2pi(double x) {
  double r = x - rint(x);  // exact for |x| < 2^52
  double s, c;
  sincospi(2.0 * r, &s, &c);
  return make_c(c, -s);
}

}  // namespace musim
