/* musim.h -- C ABI of the B200-native muspinsim hot path (libmusim.so).
 *
 * The reference (muspinsim v2.3.1) has no FFI for this path: the hot loop is the Python
 * `for cfg in self._config[rank::size]` of ExperimentRunner.run (muspinsim/experiment.py:358-382)
 * calling run_single (experiment.py:434-498) once per configuration snapshot, which in turn
 * calls Hermitian.diag (spinop.py:51-82), Operator.basis_change (spinop.py:330-355),
 * Hamiltonian.evolve / fast_evolve / integrate_decaying (hamiltonian.py:40-217), the Cython
 * kernel parallel_fast_time_evolve (cython/parallel.pyx:16-68) and Lindbladian.evolve /
 * integrate_decaying (lindbladian.py:43-173).  The entry points below replace that whole loop
 * with ONE batched call per group of configurations (a per-configuration FFI would only
 * re-create the reference's overhead).
 *
 * Conventions
 *   - plain C, no torch types; returns 0 on success or a negative MUSIM_E* code, never throws;
 *   - complex matrices are row-major, interleaved (re, im) doubles;
 *   - units as in the reference: H in MHz (frequency), B in T, gamma in MHz/T, times in us;
 *   - `musim_run` takes DEVICE pointers owned by the caller (e.g. torch tensors) and is
 *     asynchronous on `stream`; `musim_run_host` takes HOST pointers and includes the
 *     host<->device copies and a final synchronisation;
 *   - one handle per device; a handle is not thread-safe.
 */
#ifndef MUSIM_H
#define MUSIM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct musim_handle musim_handle;

/* error codes */
#define MUSIM_OK 0
#define MUSIM_EINVAL -1   /* bad argument (maps to ValueError on the Python side) */
#define MUSIM_ECUDA -2    /* CUDA runtime error; see musim_last_error */
#define MUSIM_ENOMEM -3
#define MUSIM_ENOTCONV -4 /* eigensolver did not converge */
#define MUSIM_EUNSUP -5   /* unsupported size / mode */

/* evaluation modes (which reference function each configuration would have gone through) */
#define MUSIM_MODE_EVOLVE 0       /* Hamiltonian.evolve, thermal rho0     hamiltonian.py:40-117 */
#define MUSIM_MODE_FAST 1         /* Hamiltonian.fast_evolve (T=inf|B=0)  hamiltonian.py:166-217 */
#define MUSIM_MODE_INTEGRAL 2     /* Hamiltonian.integrate_decaying / tau hamiltonian.py:119-164 */
#define MUSIM_MODE_LINDBLAD 3     /* Lindbladian.evolve                   lindbladian.py:43-111 */
#define MUSIM_MODE_LINDBLAD_INT 4 /* Lindbladian.integrate_decaying / tau lindbladian.py:113-173 */
#define MUSIM_MODE_INTEGRAL_FAST 5 /* integrate_decaying when the other spins are maximally mixed
                                      (T=inf|B=0): same value, weights |O'|^2/d_other */

/* Create a handle for one spin system on CUDA device `device`.
 *   d            Hilbert-space dimension (= prod dims)
 *   dims[n]      2I+1 of each spin, in Kronecker order (last spin fastest, spinsys.py:581-592)
 *   gammas[n]    gyromagnetic ratios, MHz/T (constants.py:12-53)
 *   muon_index   index of the muon in the spin list (spinsys.py:651)
 *   H0           d*d complex: field-independent Hamiltonian = MuonSpinSystem.hamiltonian
 *                (spinsys.py:613-626)
 *   Z            3*d*d complex: Z_a = sum_i gamma_i S_i^a so that Hz = sum_a B_a Z_a
 *                (ExperimentRunner.Hz, experiment.py:238-250)
 *   M            3*d*d complex: M_a = S_mu^a (x) 1 so that the observable is sum_a p_a M_a
 *                (MuonSpinSystem.muon_operator, spinsys.py:707-732)
 *   n_diss, diss_spin[], diss_rate[]   dissipation terms (simconfig.py:287-290;
 *                ExperimentRunner.dissipation_operators, experiment.py:271-325); 0 if none.
 * All pointers are HOST pointers and are copied. */
int musim_create(musim_handle **h, int device, int d, int n_spins, const int *dims,
                 const double *gammas, int muon_index, const double *H0, const double *Z,
                 const double *M, int n_diss, const int *diss_spin, const double *diss_rate);

/* Replace H0 / Z (resident-system fitting: FittingRunner re-creates the system per function
 * evaluation, fitting.py:126-135; here only the coupling-dependent matrices are re-uploaded). */
int musim_update_system(musim_handle *h, const double *H0, const double *Z);

/* Replace the three observable components M[3][d][d] (HOST): one device handle then serves every
 * operator of a per-call Hamiltonian.evolve / integrate_decaying(rho0, ..., operators)
 * (hamiltonian.py:40-164), which the reference evaluates operator by operator. */
int musim_update_observables(musim_handle *h, const double *M);

/* Use an explicit initial density matrix (d*d complex, HOST) for every configuration instead
 * of the thermal product state of ExperimentRunner.rho0 (experiment.py:170-236).  This is the
 * per-call boundary Hamiltonian.evolve(rho0, times, operators) (hamiltonian.py:40).  NULL
 * restores the thermal construction. */
int musim_set_rho0(musim_handle *h, const double *rho0);

/* Explicit, configuration-independent dissipators for the Lindbladian modes: n operators
 * A[n,d,d] (complex, HOST) with rates gamma[n], added to the field/temperature dependent ones
 * built from `diss_spin`.  This is the per-call boundary Lindbladian.from_hamiltonian(H,
 * dissipators=[(A, gamma), ...]) (lindbladian.py:18-41).  n = 0 clears them. */
int musim_set_dissipators(musim_handle *h, int n, const double *A, const double *gamma);

/* Tunables: "eigh" (0 auto, 1 Jacobi, 2 Householder+QL), "polar" (0 auto, 1 direct sincos,
 * 2 time-factorised), "chunk" (configurations per launch group, 0 auto). */
int musim_set_option(musim_handle *h, const char *key, long value);

/* Evaluate n_cfg configurations and ACCUMULATE  w[c] * signal_c  into out[slot[c], :].
 *   B[n,3], p[n,3], T[n]  field (T), muon polarisation and temperature (K, may be inf) of
 *                 each configuration AFTER the crystallite rotation of
 *                 ExperimentRunner.load_config (experiment.py:384-432)
 *   w[n]          orientation weight / avg_N (experiment.py:498, simconfig.py:368)
 *   slot[n]       row of `out` this configuration is accumulated into
 *                 (MuSpinConfig.store_time_slice, simconfig.py:347-368)
 *   times[nt]     HOST pointer; ignored (nt = 1) for the integral modes
 *   tau           decay time for the integral modes (MU_TAU, constants.py:14)
 *   out[n_slots, nt]   DEVICE, float64, += semantics
 * B, p, T, w, slot, out are DEVICE pointers. */
int musim_run(musim_handle *h, int mode, int64_t n_cfg, const double *B, const double *p,
              const double *T, const double *w, const int32_t *slot, int nt, const double *times,
              double tau, int n_slots, double *out, void *cuda_stream);

/* Same with HOST pointers for everything; `out` is read, accumulated into and written back. */
int musim_run_host(musim_handle *h, int mode, int64_t n_cfg, const double *B, const double *p,
                   const double *T, const double *w, const int32_t *slot, int nt,
                   const double *times, double tau, int n_slots, double *out);

/* Same as musim_run_host, but the configuration table is EXPANDED ON THE DEVICE from the axis
 * tables instead of being passed as n_cfg-long arrays (the reference builds one namedtuple per
 * configuration, MuSpinConfig.__getitem__ simconfig.py:497-519, and rotates B and p one at a time,
 * ExperimentRunner.load_config experiment.py:384-432; at 10^7 configurations that host step costs
 * more than the GPU evaluation).  Configuration i = 0 .. n_cfg-1 of this call is the global
 * configuration c = first + i * step (step = number of ranks, experiment.py:369) and
 *   idx_a = (c / div[a]) % len[a]      a = 0 polarisation, 1 field, 2 intrinsic field,
 *                                          3 orientation, 4 temperature
 *   B = R(q) B_lab + B_int,  p = R(q) p_lab,  T = Tv[idx_4],  w = ow[idx_3],
 *   slot = sum_a idx_a * slot_mult[a]
 * pol[len0,3] unit vectors, Blab[len1,3], Bint[len2,3], quat[len3,4] conjugate orientation
 * quaternions (w,x,y,z) (simconfig.py:596-613), ow[len3] orientation weights already divided by
 * avg_N, Tv[len4].  All pointers are HOST pointers. */
int musim_run_axes_host(musim_handle *h, int mode, int64_t n_cfg, int64_t first, int64_t step,
                        const int64_t *len, const int64_t *div, const int64_t *slot_mult,
                        const double *pol, const double *Blab, const double *Bint, const double *quat,
                        const double *ow, const double *Tv, int nt, const double *times, double tau,
                        int n_slots, double *out);

/* Batched complex-Hermitian eigensolver on its own (replaces np.linalg.eigh in
 * Hermitian.diag, spinop.py:69).  A[batch,d,d] complex row-major (only read), evals[batch,d]
 * ascending, evecs[batch,d,d] complex row-major with eigenvectors in COLUMNS (numpy
 * convention).  DEVICE pointers.  method: 0 auto, 1 Jacobi, 2 Householder+QL. */
int musim_eigh(int device, int d, int64_t batch, const double *A, double *evals, double *evecs,
               int method, void *cuda_stream);

/* Density matrices rho(t_k) = U [R0 .* exp(-2 pi i (l_i - l_j) t_k)] U^H, R0 = U^H rho0 U: the result of
 * Hamiltonian.evolve(rho0, times, operators=None) (hamiltonian.py:86-115, the branch without
 * operators).  evals[d], evecs[d,d] as musim_eigh returns them, rho0[d,d] complex, rho_t[nt,d,d] complex:
 * DEVICE pointers; times[nt] (microseconds) is a HOST pointer. */
int musim_evolve_rho(int device, int d, const double *evals, const double *evecs, const double *rho0,
                     int nt, const double *times, double *rho_t, void *cuda_stream);

/* Host-side tables of the type-1 NUFFT polarisation kernel (polar_nufft.cuh), computed without a
 * GPU: fine-grid size *M for nt time points, spreading width *w, polynomial degree *deg,
 * coef[w*(deg+1)] (per-tap monomial coefficients in y = 2x of the "exponential of semicircle"
 * kernel) and deconv[nt] (1 / kernel Fourier transform at mode k - nt/2).  Any pointer may be NULL. */
int musim_nufft_tables(int nt, int *M, int *w, int *deg, double *coef, double *deconv);

/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
int64_t musim_launch_count(musim_handle *h);

/* Milliseconds of device time spent in the named phase ("eigh", "rotate", "polar", ...)
 * during the last musim_run_host / musim_run with profiling enabled (option "profile" = 1).
 * The pseudo-phase "axes_resident_hits" returns a counter instead: how many musim_run_axes_host
 * calls found their configuration table already expanded on the device (resident-system
 * fitting, fitting.py:126-151); "lind_gemm_cfgs" the number of (d^2 x d^2 complex GEMM,
 * configuration) pairs the Lindbladian path has executed so far (matrix exponential by scaling
 * and squaring: the count depends on the norm of the super-operator). */
double musim_phase_ms(musim_handle *h, const char *phase);

/* FP64 peak micro-benchmarks used as roofline denominators (MEASURED_PEAKS.json has no FP64
 * entry).  kind 0: DFMA (vector pipe), 1: DMMA m8n8k4 (mma.sync f64).  Returns TFLOP/s. */
int musim_fp64_peak(int device, int kind, double *tflops);

const char *musim_last_error(musim_handle *h);
int musim_destroy(musim_handle *h);
/* Workspaces that a destroyed handle returns are kept in a per-device cache of the library (by exact
 * size), so that the next handle of the process (the reference builds a new ExperimentRunner per
 * input file / fit evaluation, muspinsim/__main__.py, fitting.py:126-151) does not pay cudaMalloc /
 * cudaFree of tens of GB again.  This call frees the cached blocks (the cache also empties itself when an
 * allocation fails or when it would hold more than half of the device memory). */
int musim_trim_pool(int device);
/* Celio's method (Phys. Rev. Lett. 56, 2720): Trotter-split evolution of `n_states` state vectors and
 * the muon polarisation <psi| sigma_mu (x) 1 |psi> at every time step, summed over the states into
 * results[num_times] (+=).  Replaces the reference's C++ extension call
 *     muspinsim.cpp.celio_evolve(num_times, psi, sigma_mu, half_dim, k, evol_contribs, results)
 * (cpp/celio.cpp:23-69; gate application parallel.cpp:236-266, measurement parallel.cpp:130-168),
 * which CelioHamiltonian._fast_evolve_cpp (celio.py:433-476) calls once per random initial state;
 * here all states go in one call.  All pointers are HOST pointers.
 *   psi        [n_states][dim] complex128, row-major
 *   sigma_mu   [2][2] complex128 (Hermitian)
 *   contribution c: matrix [mat_dim[c]][mat_dim[c]] complex128 (concatenated in `matrices`),
 *                   other_dim[c] = dim / mat_dim[c], indices[c][dim] (Celio_EvolveContrib.indices)
 *   flags      bit 0: force the streamed (global-memory) path also for small systems (cross-check)
 * Returns 0, MUSIM_EINVAL, MUSIM_ECUDA or MUSIM_EUNSUP (a gate larger than 64 x 64). */
int musim_celio_evolve(int device, int64_t dim, int n_states, const double *psi, const double *sigma_mu,
                       int64_t half_dim, int k, int n_contrib, const int32_t *mat_dim, const int64_t *other_dim,
                       const double *matrices, const int64_t *indices, int num_times, double *results, int flags);

/* Kernel launches issued by musim_celio_evolve so far (process-wide). */
int64_t musim_celio_launch_count(void);

int musim_version(void);

/* Number of CUDA devices visible to the process (0 if there is none): lets the reference-side
 * adapter map an MPI rank of `mpirun -n N muspinsim.mpi` (mpi.py:18-27) to a device without torch. */
int musim_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MUSIM_H */
