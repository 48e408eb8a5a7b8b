"""Per-call boundary: `Hamiltonian` with the reference's method signatures, computing on the GPU.

Mirrors /root/reference/muspinsim/hamiltonian.py:19-217 (and Hermitian.diag, spinop.py:51-82):
same names, argument meaning, return shapes and exception types (validation.py:8-101).  Each
call is a batch of one through the same C ABI as the batched runner -- useful for tests and
for library users; the fast way to run many configurations is `ExperimentRunner`.
"""

from numbers import Number

import numpy as np

from . import _lib


def _dense(x):
    m = getattr(x, "matrix", x)
    if hasattr(m, "toarray"):
        m = m.toarray()
    return np.asarray(m, dtype=complex)


def validate_times(times):
    if not isinstance(times, np.ndarray):
        raise TypeError("times must be an array of values in microseconds")
    if len(times.shape) != 1:
        raise ValueError("times must be an array of values in microseconds")


class Hamiltonian:
    herm_tol = 1e-6  # spinop.py:86

    def __init__(self, matrix, dim=None, device=0):
        M = _dense(matrix)
        if M.ndim != 2 or M.shape[0] != M.shape[1]:
            raise ValueError("Matrix passed to Operator must be a square 2D array.")
        if not np.all(np.isclose(M, M.conj().T, atol=self.herm_tol)):
            raise ValueError("Operator must be hermitian")
        self._matrix = M
        self._dim = tuple(dim) if dim is not None else (M.shape[0],)
        if int(np.prod(self._dim)) != M.shape[0]:
            raise ValueError("Matrix size is not compatible with the dimension tuple")
        self._device = device
        self._eigh = None
        self._handles = {}

    @property
    def matrix(self):
        return self._matrix

    @property
    def dimension(self):
        return self._dim

    # ---- Hermitian.diag (spinop.py:51-82) -> musim_eigh ----
    def diag(self):
        if self._eigh is None:
            import torch

            d = self._matrix.shape[0]
            A = torch.from_numpy(np.ascontiguousarray(self._matrix)[None]).cuda(self._device)
            ev = torch.empty(1, d, dtype=torch.float64, device=A.device)
            U = torch.empty(1, d, d, dtype=torch.complex128, device=A.device)
            _lib.eigh_device(self._device, d, 1, A.data_ptr(), ev.data_ptr(), U.data_ptr(), 0,
                             torch.cuda.current_stream(A.device).cuda_stream)
            torch.cuda.synchronize(A.device)
            self._eigh = (ev[0].cpu().numpy(), U[0].cpu().numpy())
        return self._eigh

    def _get_handle(self, dims):
        """ONE device handle per Hamiltonian and dimension tuple: H0 is uploaded once, every call
        only replaces the observable (musim_update_observables) and rho0."""
        key = tuple(int(x) for x in dims)
        h = self._handles.get(key)
        if h is None:
            d = self._matrix.shape[0]
            zeros = np.zeros((3, d, d), dtype=complex)
            h = _lib.Handle(self._device, list(key), np.zeros(len(key)), 0, self._matrix, zeros, zeros)
            self._handles[key] = h
        return h

    def close(self):
        for h in self._handles.values():
            h.close()
        self._handles = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _expect(self, mode, rho0, times, tau, op):
        """<op>(t) for one operator.  The device path evaluates Re sum_ij rho'_ij O'_ji e^{..} from the
        i <= j triangle, which is the full (real) expectation value only for Hermitian O; a
        non-Hermitian operator (the reference's SpinOperator API allows e.g. S+) is split into its
        Hermitian and anti-Hermitian parts, O = Oh + i Oa, and evaluated as <Oh> + i <Oa>."""
        O = _dense(op)
        if O.shape != self._matrix.shape:
            raise ValueError("Incompatible operator dimension")
        Oh, Oa = 0.5 * (O + O.conj().T), -0.5j * (O - O.conj().T)
        parts = [Oh] if np.max(np.abs(Oa)) <= 1e-15 * max(1.0, np.max(np.abs(O))) else [Oh, Oa]
        h = self._get_handle([self._matrix.shape[0]])
        h.set_rho0(rho0)
        nt = len(times) if times is not None else 1
        res = np.zeros(nt, dtype=complex)
        one = np.array([[1.0, 0.0, 0.0]])
        for k, part in enumerate(parts):
            M = np.zeros((3,) + O.shape, dtype=complex)
            M[0] = part
            h.update_observables(M)
            out = np.zeros((1, nt))
            h.run_host(mode, np.zeros((1, 3)), one, np.array([np.inf]), np.array([1.0]), np.array([0]), times,
                       tau, out)
            res = res + (1j if k else 1.0) * out[0]
        return res

    def _check_rho0(self, rho0, name="rho0"):
        r = _dense(rho0)
        if r.shape != self._matrix.shape:
            raise ValueError("Incompatible rho0 dimension")
        # a DensityOperator is Hermitian by construction (spinop.py:431-453); the device path relies on it
        if not np.all(np.isclose(r, r.conj().T, atol=self.herm_tol)):
            raise ValueError("rho0 must be a Hermitian density matrix")
        return r

    def evolve(self, rho0, times, operators=None):
        """hamiltonian.py:40-117: expectation values [nt, n_ops] (complex)."""
        if operators is None:
            operators = []
        times = np.array(times)
        if not isinstance(operators, (list, tuple)):
            operators = [operators]
        validate_times(times)
        r = self._check_rho0(rho0)
        if len(operators) == 0:
            return self._evolve_rho(r, times.astype(float))
        cols = [self._expect(_lib.MODE_EVOLVE, r, times.astype(float), 1.0, o) for o in operators]
        return np.array(cols).T.astype(complex)

    def _evolve_rho(self, rho0, times):
        """hamiltonian.py:108-116: without operators the reference returns the density matrices themselves,
        rho(t) = V [rho0' .* exp(-2 pi i (l_i - l_j) t)] V^H; here as one complex array [nt, d, d]
        (musim_evolve_rho: two batched GEMMs over the time points on the device)."""
        import torch

        d = self._matrix.shape[0]
        dev = torch.device("cuda", self._device)
        A = torch.from_numpy(np.ascontiguousarray(self._matrix)[None]).to(dev)
        ev = torch.empty(1, d, dtype=torch.float64, device=dev)
        U = torch.empty(1, d, d, dtype=torch.complex128, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.eigh_device(self._device, d, 1, A.data_ptr(), ev.data_ptr(), U.data_ptr(), 0, stream)
        R = torch.from_numpy(np.ascontiguousarray(rho0, dtype=complex)).to(dev)
        out = torch.empty(len(times), d, d, dtype=torch.complex128, device=dev)
        _lib.evolve_rho_device(self._device, d, ev.data_ptr(), U.data_ptr(), R.data_ptr(), times, out.data_ptr(), stream)
        torch.cuda.synchronize(dev)
        return out.cpu().numpy()

    def integrate_decaying(self, rho0, tau, operators):
        """hamiltonian.py:119-164: sum_ab rho'_ab O'_ba / (1/tau + 2 pi i (l_a - l_b)) per operator."""
        if not isinstance(operators, (list, tuple)):
            operators = [operators]
        if not (isinstance(tau, Number) and np.isreal(tau) and tau > 0):
            raise ValueError("'tau' must be a real number > 0")
        if not operators:
            raise ValueError("At least one SpinOperator must be present in 'operators'")
        r = self._check_rho0(rho0)
        # the ABI returns the integral divided by tau (experiment.py:492-496)
        return np.array([self._expect(_lib.MODE_INTEGRAL, r, None, float(tau), o)[0] * tau for o in operators]).astype(complex)

    def fast_evolve(self, sigma_mu, times, other_dimension):
        """hamiltonian.py:166-217: muon first, other spins maximally mixed; result in [-0.5, 0.5]."""
        times = np.array(times)
        validate_times(times)
        sig = _dense(sigma_mu)
        d = self._matrix.shape[0]
        if sig.shape != (2, 2) or 2 * int(other_dimension) != d:
            raise ValueError("sigma_mu must be 2x2 and other_dimension half the total dimension")
        if not np.all(np.isclose(sig, sig.conj().T, atol=self.herm_tol)):
            raise ValueError("sigma_mu must be Hermitian")
        O = np.kron(0.5 * sig, np.eye(int(other_dimension)))
        h = self._get_handle([2, int(other_dimension)])
        M = np.zeros((3, d, d), dtype=complex)
        M[0] = O
        h.update_observables(M)
        out = np.zeros((1, len(times)))
        h.run_host(_lib.MODE_FAST, np.zeros((1, 3)), np.array([[1.0, 0, 0]]), None, np.array([1.0]),
                   np.array([0]), times.astype(float), 1.0, out)
        return out[0]
