// lindblad.cuh -- placeholder until the Lindbladian path lands.
#pragma once
#include <string>
#include "common.cuh"
#include "rotate.cuh"
namespace musim {
inline int lindblad_run(int, const SpinTable &, int, const int *, const double *, const cplx *, const cplx *,
                        const cplx *, const cplx *, bool, int64_t, const double *, const double *,
                        const double *, const double *, const int32_t *, int, const double *, bool, double,
                        double, double, double *, void **, size_t *, long, cudaStream_t, int64_t *,
                        std::string &err) {
  err = "Lindbladian path not built yet";
  return -5;
}
}  // namespace musim
