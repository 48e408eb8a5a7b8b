// profiler.cuh -- non-blocking per-kernel device timers (CUDA events on the launching stream).
// Events are recorded around each launch without synchronising; elapsed times are resolved
// lazily when queried, so the timed region of bench.py is not perturbed.
#pragma once
#include <cuda_runtime.h>
#include <string.h>

#include <vector>

namespace musim {

enum Phase {
  PH_EIGH_TRIDIAG = 0,
  PH_EIGH_TQL,
  PH_EIGH_APPLY,
  PH_EIGH_BACK,
  PH_EIGH_JACOBI,
  PH_ROTATE,
  PH_RHO0,
  PH_POLAR,
  PH_INTEGRAL,
  PH_LINDBLAD,
  PH_EIGH_TDC,
  PH_COUNT
};
static const char *const kPhaseNames[PH_COUNT] = {"eigh_tridiag", "eigh_tql", "eigh_apply", "eigh_back",
                                                  "eigh_jacobi",  "rotate",   "rho0",       "polar",
                                                  "integral",     "lindblad", "eigh_tdc"};

struct Profiler {
  bool on = false;
  struct Rec {
    int ph;
    cudaEvent_t a, b;
  };
  std::vector<Rec> pool;
  size_t used = 0;
  double ms[PH_COUNT] = {0};
  long count[PH_COUNT] = {0};

  void reset() {
    resolve();
    for (int i = 0; i < PH_COUNT; ++i) {
      ms[i] = 0.0;
      count[i] = 0;
    }
  }
  int begin(int ph, cudaStream_t st) {
    if (!on) return -1;
    if (used == pool.size()) {
      Rec r;
      r.ph = ph;
      cudaEventCreate(&r.a);
      cudaEventCreate(&r.b);
      pool.push_back(r);
    }
    pool[used].ph = ph;
    cudaEventRecord(pool[used].a, st);
    return (int)used++;
  }
  void end(int id, cudaStream_t st) {
    if (id >= 0) cudaEventRecord(pool[id].b, st);
  }
  void resolve() {
    for (size_t i = 0; i < used; ++i) {
      float t = 0.f;
      if (cudaEventSynchronize(pool[i].b) == cudaSuccess &&
          cudaEventElapsedTime(&t, pool[i].a, pool[i].b) == cudaSuccess) {
        ms[pool[i].ph] += t;
        ++count[pool[i].ph];
      }
    }
    used = 0;
  }
  void destroy() {
    for (auto &r : pool) {
      cudaEventDestroy(r.a);
      cudaEventDestroy(r.b);
    }
    pool.clear();
    used = 0;
  }
};

struct ProfScope {
  Profiler *p;
  cudaStream_t st;
  int id;
  ProfScope(Profiler *p_, cudaStream_t st_, int ph) : p(p_), st(st_), id(p_ ? p_->begin(ph, st_) : -1) {}
  ~ProfScope() {
    if (p) p->end(id, st);
  }
};

}  // namespace musim
