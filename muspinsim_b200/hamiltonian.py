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

    def _handle(self, ops, muon_first_dims=None):
        d = self._matrix.shape[0]
        zeros = np.zeros((3, d, d), dtype=complex)
        M = zeros.copy()
        M[0] = ops
        dims = muon_first_dims or [d]
        return _lib.Handle(self._device, dims, np.zeros(len(dims)), 0, self._matrix, zeros, M)

    def _run(self, mode, rho0, times, tau, op):
        h = self._handle(_dense(op))
        try:
            if rho0 is not None:
                h.set_rho0(rho0)
            nt = len(times) if times is not None else 1
            out = np.zeros((1, nt))
            one = np.array([[1.0, 0.0, 0.0]])
            h.run_host(mode, np.zeros((1, 3)), one, np.array([np.inf]), np.array([1.0]), np.array([0]), times,
                       tau, out)
            return out[0]
        finally:
            h.close()

    def _check_rho0(self, rho0, name="rho0"):
        r = _dense(rho0)
        if r.shape != self._matrix.shape:
            raise ValueError("Incompatible rho0 dimension")
        return r

    def evolve(self, rho0, times, operators=None):
        """hamiltonian.py:40-117: expectation values [nt, n_ops] (complex; the imaginary part
        of a Hermitian observable's expectation is zero and is returned as such)."""
        if operators is None:
            operators = []
        times = np.array(times)
        if not isinstance(operators, (list, tuple)):
            operators = [operators]
        validate_times(times)
        if len(operators) == 0:
            raise NotImplementedError("density-matrix output (operators=None) is outside the hot path")
        r = self._check_rho0(rho0)
        cols = [self._run(_lib.MODE_EVOLVE, r, times.astype(float), 1.0, o) for o in operators]
        return np.array(cols).T.astype(complex)

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
        return np.array([self._run(_lib.MODE_INTEGRAL, r, None, float(tau), o)[0] * tau for o in operators]).astype(complex)

    def fast_evolve(self, sigma_mu, times, other_dimension):
        """hamiltonian.py:166-217: muon first, other spins maximally mixed; result in [-0.5, 0.5]."""
        times = np.array(times)
        validate_times(times)
        sig = _dense(sigma_mu)
        d = self._matrix.shape[0]
        if sig.shape != (2, 2) or 2 * int(other_dimension) != d:
            raise ValueError("sigma_mu must be 2x2 and other_dimension half the total dimension")
        O = np.kron(0.5 * sig, np.eye(int(other_dimension)))
        h = self._handle(O, muon_first_dims=[2, int(other_dimension)])
        try:
            out = np.zeros((1, len(times)))
            h.run_host(_lib.MODE_FAST, np.zeros((1, 3)), np.array([[1.0, 0, 0]]), None, np.array([1.0]),
                       np.array([0]), times.astype(float), 1.0, out)
            return out[0]
        finally:
            h.close()
