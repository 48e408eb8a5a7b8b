// common.cuh -- shared helpers for the musim sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <mutex>
#include <unordered_map>

#define MUSIM_MAX_DEVICES 64

namespace musim {

// Device memory of the library goes through a small per-device cache of freed blocks: a workspace that a
// destroyed handle returns is kept (by exact size) and handed to the next handle of the process -- a fresh
// ExperimentRunner, every run of a scan script -- instead of a cudaFree / cudaMalloc of tens of GB (measured
// at C5: 26-40 ms of a 77 ms ExperimentRunner(spec).run() and 60-90 ms per close()).  Same blocking semantics
// as cudaMalloc / cudaFree (a free waits for the device first).  musim_trim_pool() frees the cache; it is
// also emptied when an allocation fails or when it would exceed half of the device memory.  (The driver's
// stream-ordered pool was tried first: same warm behaviour, but 1.9 s for the first 26 GB.)
struct DevBlockCache {
  std::mutex mu;
  std::unordered_map<void *, size_t> live[MUSIM_MAX_DEVICES];          // blocks handed out
  std::unordered_multimap<size_t, void *> idle[MUSIM_MAX_DEVICES];     // freed blocks by size
  size_t idle_bytes[MUSIM_MAX_DEVICES] = {}, total_bytes[MUSIM_MAX_DEVICES] = {};
  static DevBlockCache &get() {
    static DevBlockCache c;
    return c;
  }
  void trim(int dev) {  // caller holds mu
    for (auto &kv : idle[dev]) cudaFree(kv.second);
    idle[dev].clear();
    idle_bytes[dev] = 0;
  }
};
inline cudaError_t dev_malloc(void **p, size_t bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (bytes == 0) bytes = 1;
  DevBlockCache &c = DevBlockCache::get();
  if (dev < 0 || dev >= MUSIM_MAX_DEVICES) return cudaMalloc(p, bytes);
  std::lock_guard<std::mutex> lk(c.mu);
  auto it = c.idle[dev].find(bytes);
  if (it != c.idle[dev].end()) {
    *p = it->second;
    c.idle[dev].erase(it);
    c.idle_bytes[dev] -= bytes;
  } else {
    e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) {
      (void)cudaGetLastError();
      c.trim(dev);
      e = cudaMalloc(p, bytes);
    }
    if (e != cudaSuccess) return e;
  }
  c.live[dev][*p] = bytes;
  return cudaSuccess;
}
inline cudaError_t dev_free(void *p) {
  if (!p) return cudaSuccess;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  DevBlockCache &c = DevBlockCache::get();
  if (dev < 0 || dev >= MUSIM_MAX_DEVICES) return cudaFree(p);
  std::lock_guard<std::mutex> lk(c.mu);
  auto it = c.live[dev].find(p);
  if (it == c.live[dev].end()) return cudaFree(p);  // not ours (or allocated under another current device)
  const size_t bytes = it->second;
  c.live[dev].erase(it);
  e = cudaDeviceSynchronize();  // as cudaFree: nothing on the device uses the block any more
  if (e != cudaSuccess) return e;
  if (c.total_bytes[dev] == 0) {  // (cudaMemGetInfo takes ~1 ms: once per device)
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) c.total_bytes[dev] = total_b;
  }
  if (c.idle_bytes[dev] + bytes > c.total_bytes[dev] / 2) {
    c.trim(dev);
    return cudaFree(p);
  }
  c.idle[dev].emplace(bytes, p);
  c.idle_bytes[dev] += bytes;
  return cudaSuccess;
}
inline cudaError_t dev_trim(int dev) {
  if (dev < 0 || dev >= MUSIM_MAX_DEVICES) return cudaErrorInvalidDevice;
  DevBlockCache &c = DevBlockCache::get();
  std::lock_guard<std::mutex> lk(c.mu);
  cudaError_t e = cudaDeviceSynchronize();
  c.trim(dev);
  return e;
}

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
