/* C restatement of the reference's Cython kernel `parallel_fast_time_evolve`
 * (/root/reference/muspinsim/cython/parallel.pyx:16-68) -- TEST INFRASTRUCTURE (oracle).
 *
 *   res[i] = (1/d_o) * ( sum_j 0.5*A[j,j] + sum_{k<j} A[j,k]*cos(W[j,k]*t_i) )
 *
 * Same loop order as the reference (j outer, k inner, accumulation into res[i]) so the
 * floating-point summation order is identical.  The reference's `prange` over time points
 * becomes an OpenMP loop only when built with -fopenmp (the reference's default build has
 * OpenMP off, setup.py:25-27).
 */
#include <math.h>

void oracle_fast_time_evolve(const double *times, long nt, int other_dimension,
                             const double *A, const double *W, long d, double *res)
{
    const double one_over_d = 1.0 / (double)other_dimension;
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (long i = 0; i < nt; ++i) {
        double r = res[i];
        for (long j = 0; j < d; ++j) {
            r += 0.5 * A[j * d + j];
            for (long k = 0; k < j; ++k)
                r += A[j * d + k] * cos(W[j * d + k] * times[i]);
        }
        res[i] = r * one_over_d;
    }
}
