// eigh_hql.cuh -- placeholder until the Householder + implicit-QL solver lands.
#pragma once
#include "common.cuh"
namespace musim {
inline bool hql_supported(int) { return false; }
inline int launch_eigh_hql(int, int64_t, const cplx *, const cplx *, const double *, const cplx *, double *,
                           cplx *, int *, cudaStream_t, int64_t *) {
  return -5;
}
}  // namespace musim
