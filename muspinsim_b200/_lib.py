"""ctypes binding of libmusim.so (include/musim.h).

There is deliberately NO fallback: if the CUDA library is missing or a call fails, an
exception is raised.  PyTorch is used only by callers for device buffers / streams /
torch.distributed; nothing here depends on it.
"""

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmusim.so")

MODE_EVOLVE, MODE_FAST, MODE_INTEGRAL, MODE_LINDBLAD, MODE_LINDBLAD_INT, MODE_INTEGRAL_FAST = range(6)

_ERRORS = {-1: "EINVAL", -2: "ECUDA", -3: "ENOMEM", -4: "ENOTCONV", -5: "EUNSUP"}

# every symbol include/musim.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "musim_create",
    "musim_update_system",
    "musim_update_observables",
    "musim_set_rho0",
    "musim_set_dissipators",
    "musim_set_option",
    "musim_run",
    "musim_run_host",
    "musim_run_axes_host",
    "musim_nufft_tables",
    "musim_eigh",
    "musim_evolve_rho",
    "musim_launch_count",
    "musim_phase_ms",
    "musim_fp64_peak",
    "musim_trim_pool",
    "musim_celio_evolve",
    "musim_celio_launch_count",
    "musim_device_count",
    "musim_last_error",
    "musim_destroy",
    "musim_version",
]


class MusimError(RuntimeError):
    pass


_lib = None


def load():
    """Load libmusim.so; raise if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MusimError(
            "CUDA library %s is missing: build it with `make -C muspinsim_b200/csrc` "
            "(there is no CPU fallback)" % LIB_PATH
        )
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double
    lib.musim_create.argtypes = [ctypes.POINTER(vp), i32, i32, i32, vp, vp, i32, vp, vp, vp, i32, vp, vp]
    lib.musim_create.restype = i32
    lib.musim_update_system.argtypes = [vp, vp, vp]
    lib.musim_update_system.restype = i32
    lib.musim_update_observables.argtypes = [vp, vp]
    lib.musim_update_observables.restype = i32
    lib.musim_set_rho0.argtypes = [vp, vp]
    lib.musim_set_rho0.restype = i32
    lib.musim_set_dissipators.argtypes = [vp, i32, vp, vp]
    lib.musim_set_dissipators.restype = i32
    lib.musim_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_long]
    lib.musim_set_option.restype = i32
    lib.musim_run.argtypes = [vp, i32, i64, vp, vp, vp, vp, vp, i32, vp, dbl, i32, vp, vp]
    lib.musim_run.restype = i32
    lib.musim_run_host.argtypes = [vp, i32, i64, vp, vp, vp, vp, vp, i32, vp, dbl, i32, vp]
    lib.musim_run_host.restype = i32
    lib.musim_run_axes_host.argtypes = [vp, i32, i64, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, dbl, i32, vp]
    lib.musim_run_axes_host.restype = i32
    lib.musim_nufft_tables.argtypes = [i32, vp, vp, vp, vp, vp]
    lib.musim_nufft_tables.restype = i32
    lib.musim_eigh.argtypes = [i32, i32, i64, vp, vp, vp, i32, vp]
    lib.musim_eigh.restype = i32
    lib.musim_evolve_rho.argtypes = [i32, i32, vp, vp, vp, i32, vp, vp, vp]
    lib.musim_evolve_rho.restype = i32
    lib.musim_launch_count.argtypes = [vp]
    lib.musim_launch_count.restype = i64
    lib.musim_phase_ms.argtypes = [vp, ctypes.c_char_p]
    lib.musim_phase_ms.restype = dbl
    lib.musim_fp64_peak.argtypes = [i32, i32, ctypes.POINTER(dbl)]
    lib.musim_fp64_peak.restype = i32
    lib.musim_trim_pool.argtypes = [i32]
    lib.musim_trim_pool.restype = i32
    lib.musim_celio_evolve.argtypes = [i32, i64, i32, vp, vp, i64, i32, i32, vp, vp, vp, vp, i32, vp, i32]
    lib.musim_celio_evolve.restype = i32
    lib.musim_celio_launch_count.argtypes = []
    lib.musim_celio_launch_count.restype = i64
    lib.musim_device_count.argtypes = []
    lib.musim_device_count.restype = i32
    lib.musim_last_error.argtypes = [vp]
    lib.musim_last_error.restype = ctypes.c_char_p
    lib.musim_destroy.argtypes = [vp]
    lib.musim_destroy.restype = i32
    lib.musim_version.argtypes = []
    lib.musim_version.restype = i32
    _lib = lib
    return lib


def _c128(a):
    return np.ascontiguousarray(a, dtype=np.complex128)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return a.ctypes.data if a is not None else None


class Handle:
    """RAII wrapper around musim_handle for one spin system on one device."""

    def __init__(self, device, dims, gammas, muon_index, H0, Z, M, diss_spin=(), diss_rate=()):
        self._lib = load()
        self._h = ctypes.c_void_p()
        dims = np.ascontiguousarray(dims, dtype=np.int32)
        gammas = _f64(gammas)
        d = int(np.prod(dims))
        H0, Z, M = _c128(H0), _c128(Z), _c128(M)
        if H0.shape != (d, d) or Z.shape != (3, d, d) or M.shape != (3, d, d):
            raise ValueError("H0 must be (d,d); Z and M must be (3,d,d)")
        ds = np.ascontiguousarray(diss_spin, dtype=np.int32)
        dr = _f64(diss_rate)
        rc = self._lib.musim_create(
            ctypes.byref(self._h), int(device), d, len(dims), _ptr(dims), _ptr(gammas), int(muon_index),
            _ptr(H0), _ptr(Z), _ptr(M), len(ds), _ptr(ds) if len(ds) else None, _ptr(dr) if len(ds) else None,
        )
        self.d = d
        self.device = int(device)
        self._check(rc)

    def _check(self, rc):
        if rc == 0:
            return
        msg = ""
        if self._h:
            msg = (self._lib.musim_last_error(self._h) or b"").decode()
        text = "libmusim error %s (%d): %s" % (_ERRORS.get(rc, "?"), rc, msg)
        if rc == -1:
            raise ValueError(text)
        raise MusimError(text)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.musim_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key, value):
        self._check(self._lib.musim_set_option(self._h, key.encode(), int(value)))

    def set_rho0(self, rho0):
        if rho0 is None:
            self._check(self._lib.musim_set_rho0(self._h, None))
        else:
            r = _c128(rho0)
            if r.shape != (self.d, self.d):
                raise ValueError("rho0 must be (d,d)")
            self._check(self._lib.musim_set_rho0(self._h, _ptr(r)))

    def set_dissipators(self, ops, rates):
        """Explicit jump operators [(d,d) complex] with rates (Lindbladian.from_hamiltonian)."""
        if len(ops) == 0:
            self._check(self._lib.musim_set_dissipators(self._h, 0, None, None))
            return
        A = _c128(np.array([np.asarray(o) for o in ops]))
        g = _f64(rates)
        if A.shape != (len(g), self.d, self.d):
            raise ValueError("Invalid dissipation operator for this Lindbladian")
        self._check(self._lib.musim_set_dissipators(self._h, len(g), _ptr(A), _ptr(g)))

    def update_system(self, H0=None, Z=None):
        H0 = _c128(H0) if H0 is not None else None
        Z = _c128(Z) if Z is not None else None
        self._check(self._lib.musim_update_system(self._h, _ptr(H0), _ptr(Z)))

    def update_observables(self, M):
        M = _c128(M)
        if M.shape != (3, self.d, self.d):
            raise ValueError("M must be (3,d,d)")
        self._check(self._lib.musim_update_observables(self._h, _ptr(M)))

    def run_host(self, mode, B, p, T, w, slot, times, tau, out):
        """All numpy (host) arrays; `out` [n_slots, nt] float64 is accumulated into in place."""
        B, p, w = _f64(B), _f64(p), _f64(w)
        n = B.shape[0]
        T = _f64(T) if T is not None else None
        slot = np.ascontiguousarray(slot, dtype=np.int32)
        times = _f64(times) if times is not None else None
        if B.shape != (n, 3) or p.shape != (n, 3) or w.shape != (n,) or slot.shape != (n,) or (T is not None and T.shape != (n,)):
            raise ValueError("B and p must be [n,3]; T, w and slot must be [n]")
        if not (out.flags.c_contiguous and out.dtype == np.float64 and out.ndim == 2):
            raise ValueError("out must be a C-contiguous float64 [n_slots, nt] array")
        if n and (slot.min() < 0 or slot.max() >= out.shape[0]):
            raise ValueError("slot indices must lie in [0, n_slots)")
        nt = len(times) if times is not None else 1
        rc = self._lib.musim_run_host(
            self._h, int(mode), n, _ptr(B), _ptr(p), _ptr(T), _ptr(w), _ptr(slot), nt, _ptr(times),
            float(tau), out.shape[0], _ptr(out),
        )
        self._check(rc)
        return out

    def run_axes_host(self, mode, n_cfg, first, step, axes, times, tau, out):
        """Configurations first, first + step, ... (n_cfg of them) of the table described by
        `axes` (ConfigTable.axes_descriptor()), expanded on the device; `out` as in run_host."""
        times = _f64(times) if times is not None else None
        if not (out.flags.c_contiguous and out.dtype == np.float64 and out.ndim == 2):
            raise ValueError("out must be a C-contiguous float64 [n_slots, nt] array")
        nt = len(times) if times is not None else 1
        keep = [np.ascontiguousarray(axes[k], dtype=np.int64) for k in ("len", "div", "slot_mult")]
        tabs = [_f64(axes[k]) for k in ("pol", "Blab", "Bint", "quat", "ow", "Tv")]
        rc = self._lib.musim_run_axes_host(
            self._h, int(mode), int(n_cfg), int(first), int(step), _ptr(keep[0]), _ptr(keep[1]), _ptr(keep[2]),
            _ptr(tabs[0]), _ptr(tabs[1]), _ptr(tabs[2]), _ptr(tabs[3]), _ptr(tabs[4]), _ptr(tabs[5]),
            nt, _ptr(times), float(tau), out.shape[0], _ptr(out),
        )
        self._check(rc)
        return out

    def run_device(self, mode, n_cfg, B_ptr, p_ptr, T_ptr, w_ptr, slot_ptr, times, tau, n_slots, out_ptr, stream=0):
        """Device pointers (ints, e.g. torch.Tensor.data_ptr()); asynchronous on `stream`."""
        times = _f64(times) if times is not None else None
        nt = len(times) if times is not None else 1
        rc = self._lib.musim_run(
            self._h, int(mode), int(n_cfg), B_ptr, p_ptr, T_ptr, w_ptr, slot_ptr, nt, _ptr(times),
            float(tau), int(n_slots), out_ptr, stream,
        )
        self._check(rc)

    @property
    def launches(self):
        return int(self._lib.musim_launch_count(self._h))

    def phase_ms(self, name):
        return float(self._lib.musim_phase_ms(self._h, name.encode()))


def nufft_tables(nt):
    """(M, w, deg, coef[w, deg+1], deconv[nt]) of the NUFFT polarisation kernel (host only)."""
    lib = load()
    M, w, deg = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = lib.musim_nufft_tables(int(nt), ctypes.byref(M), ctypes.byref(w), ctypes.byref(deg), None, None)
    if rc:
        raise ValueError("musim_nufft_tables failed (%d)" % rc)
    coef = np.zeros((w.value, deg.value + 1))
    dec = np.zeros(int(nt))
    lib.musim_nufft_tables(int(nt), None, None, None, coef.ctypes.data, dec.ctypes.data)
    return M.value, w.value, deg.value, coef, dec


def celio_evolve(device, psi, sigma_mu, k, contribs, num_times, results, streamed=False):
    """Celio's method on the GPU for a batch of state vectors (include/musim.h: musim_celio_evolve).
    psi [n_states, dim] complex; contribs = [(matrix [md, md] complex, other_dim, indices [dim])];
    results [num_times] float64 is accumulated into (summed over the states)."""
    lib = load()
    psi = _c128(np.atleast_2d(psi))
    n_states, dim = psi.shape
    sig = _c128(sigma_mu)
    if sig.shape != (2, 2) or dim % 2:
        raise ValueError("sigma_mu must be 2x2 and the state dimension even")
    md = np.ascontiguousarray([np.asarray(m).shape[0] for m, _, _ in contribs], dtype=np.int32)
    od = np.ascontiguousarray([int(o) for _, o, _ in contribs], dtype=np.int64)
    mats = _c128(np.concatenate([np.asarray(m, dtype=complex).reshape(-1) for m, _, _ in contribs])) if contribs else None
    idx = np.ascontiguousarray(np.concatenate([np.asarray(i, dtype=np.int64).reshape(-1) for _, _, i in contribs])) if contribs else None
    if contribs and idx.size != len(contribs) * dim:
        raise ValueError("every contribution needs `dim` indices")
    if not (results.flags.c_contiguous and results.dtype == np.float64 and results.shape == (num_times,)):
        raise ValueError("results must be a C-contiguous float64 [num_times] array")
    rc = lib.musim_celio_evolve(int(device), dim, n_states, _ptr(psi), _ptr(sig), dim // 2, int(k), len(contribs),
                                _ptr(md) if len(contribs) else None, _ptr(od) if len(contribs) else None,
                                _ptr(mats), _ptr(idx), int(num_times), _ptr(results), 1 if streamed else 0)
    if rc == -1:
        raise ValueError("musim_celio_evolve: invalid arguments")
    if rc:
        raise MusimError("musim_celio_evolve failed: %s (%d)" % (_ERRORS.get(rc, "?"), rc))
    return results


def device_count():
    """Number of CUDA devices visible to the process (0 without a driver / GPU)."""
    return int(load().musim_device_count())


def trim_pool(device=0):
    """Return the library's cached device memory (workspaces of destroyed handles) to the driver."""
    rc = load().musim_trim_pool(int(device))
    if rc != 0:
        raise MusimError("musim_trim_pool failed: %s (%d)" % (_ERRORS.get(rc, "?"), rc))


def fp64_peak(device=0, kind=0):
    lib = load()
    v = ctypes.c_double()
    rc = lib.musim_fp64_peak(int(device), int(kind), ctypes.byref(v))
    if rc:
        raise MusimError("musim_fp64_peak failed (%d)" % rc)
    return v.value


def evolve_rho_device(device, d, evals_ptr, evecs_ptr, rho0_ptr, times, rho_t_ptr, stream=0):
    """rho(t) for every time (musim_evolve_rho): device pointers, `times` a host float64 array."""
    lib = load()
    t = np.ascontiguousarray(times, dtype=np.float64)
    rc = lib.musim_evolve_rho(int(device), int(d), evals_ptr, evecs_ptr, rho0_ptr, int(t.size), t.ctypes.data,
                              rho_t_ptr, stream)
    if rc:
        raise MusimError("musim_evolve_rho failed: %s (%d)" % (_ERRORS.get(rc, "?"), rc))


def eigh_device(device, d, batch, A_ptr, evals_ptr, evecs_ptr, method=0, stream=0):
    lib = load()
    rc = lib.musim_eigh(int(device), int(d), int(batch), A_ptr, evals_ptr, evecs_ptr, int(method), stream)
    if rc:
        raise MusimError("musim_eigh failed: %s (%d)" % (_ERRORS.get(rc, "?"), rc))
