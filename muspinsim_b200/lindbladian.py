"""Per-call boundary for open-system dynamics: `Lindbladian` with the reference's signatures
(/root/reference/muspinsim/lindbladian.py:17-173), computed on the GPU without an
eigen-decomposition of the super-operator (see csrc/lindblad.cuh)."""

from numbers import Number

import numpy as np

from . import _lib
from .hamiltonian import Hamiltonian, _dense, validate_times


class Lindbladian:
    def __init__(self, H, dissipators=(), device=0):
        self._H = H
        self._dops = [(_dense(A), float(g)) for A, g in dissipators]
        self._device = device
        self._handle = None

    @classmethod
    def from_hamiltonian(cls, H, dissipators=()):
        """lindbladian.py:18-33."""
        if not isinstance(H, Hamiltonian):
            raise ValueError("Must use Hamiltonian to create Lindbladian")
        L = cls(H, [], H._device)
        for A, gamma in dissipators:
            L.add_dissipative_term(A, gamma)
        return L

    def add_dissipative_term(self, A, gamma=1.0):
        """lindbladian.py:35-41."""
        A = _dense(A)
        if A.shape != self._H.matrix.shape:
            raise ValueError("Invalid dissipation operator for this Lindbladian")
        self._dops.append((A, float(gamma)))

    @property
    def dimension(self):
        return self._H.dimension * 2

    def _get_handle(self):
        """ONE device handle per Lindbladian; calls only replace the observable, rho0 and jump operators."""
        if self._handle is None:
            d = self._H.matrix.shape[0]
            zeros = np.zeros((3, d, d), dtype=complex)
            self._handle = _lib.Handle(self._device, [d], [0.0], 0, self._H.matrix, zeros, zeros)
        return self._handle

    def close(self):
        if self._handle is not None:
            self._handle.close()
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _run(self, mode, rho0, times, tau, op):
        """<op> for one operator; non-Hermitian operators are evaluated as <Oh> + i <Oa> (the device
        path returns the real part of the trace, see Hamiltonian._expect)."""
        O = _dense(op)
        Oh, Oa = 0.5 * (O + O.conj().T), -0.5j * (O - O.conj().T)
        parts = [Oh] if np.max(np.abs(Oa)) <= 1e-15 * max(1.0, np.max(np.abs(O))) else [Oh, Oa]
        h = self._get_handle()
        h.set_rho0(rho0)
        h.set_dissipators([a for a, _ in self._dops], [g for _, g in self._dops])
        nt = len(times) if times is not None else 1
        res = np.zeros(nt, dtype=complex)
        for k, part in enumerate(parts):
            M = np.zeros((3,) + O.shape, dtype=complex)
            M[0] = part
            h.update_observables(M)
            out = np.zeros((1, nt))
            h.run_host(mode, np.zeros((1, 3)), np.array([[1.0, 0.0, 0.0]]), np.array([np.inf]), np.array([1.0]),
                       np.array([0]), times, tau, out)
            res = res + (1j if k else 1.0) * out[0]
        return res

    def evolve(self, rho0, times, operators=()):
        """lindbladian.py:43-111: expectation values [nt, n_ops]."""
        times = np.array(times)
        if not isinstance(operators, (list, tuple)):
            operators = [operators]
        validate_times(times)
        r = _dense(rho0)
        if r.shape != self._H.matrix.shape:
            raise ValueError("Incompatible rho0 dimension")
        if not np.all(np.isclose(r, r.conj().T, atol=Hamiltonian.herm_tol)):
            raise ValueError("rho0 must be a Hermitian density matrix")
        if any(_dense(o).shape != r.shape for o in operators):
            raise ValueError("Incompatible measure operator dimension")
        if len(operators) == 0:
            return self._evolve_rho(r, times.astype(float))
        cols = [self._run(_lib.MODE_LINDBLAD, r, times.astype(float), 1.0, o) for o in operators]
        return np.array(cols).T.astype(complex)

    def _evolve_rho(self, rho0, times):
        """lindbladian.py:103-109: without operators the reference returns the density matrices; here as
        [nt, d, d].  rho(t) stays Hermitian, so its d^2 real parameters are the expectation values of the
        Hermitian matrix units: rho_ii = <|i><i|>, Re rho_ij = <|i><j| + |j><i|> / 2,
        Im rho_ij = -<-i (|i><j| - |j><i|)> / 2 -- d^2 runs of the expectation-value path (a capability
        path for the small systems the Lindbladian is used with, not a tuned one)."""
        d = rho0.shape[0]
        out = np.zeros((len(times), d, d), dtype=complex)
        for i in range(d):
            E = np.zeros((d, d), dtype=complex)
            E[i, i] = 1.0
            out[:, i, i] = self._run(_lib.MODE_LINDBLAD, rho0, times, 1.0, E).real
            for j in range(i + 1, d):
                X = np.zeros((d, d), dtype=complex)
                X[i, j] = X[j, i] = 1.0
                Y = np.zeros((d, d), dtype=complex)
                Y[i, j], Y[j, i] = -1.0j, 1.0j
                re = 0.5 * self._run(_lib.MODE_LINDBLAD, rho0, times, 1.0, X).real
                im = -0.5 * self._run(_lib.MODE_LINDBLAD, rho0, times, 1.0, Y).real
                out[:, i, j] = re + 1j * im
                out[:, j, i] = re - 1j * im
        return out

    def integrate_decaying(self, rho0, tau, operators):
        """lindbladian.py:113-173."""
        if not isinstance(operators, (list, tuple)):
            operators = [operators]
        if not (isinstance(tau, Number) and np.isreal(tau) and tau > 0):
            raise ValueError("'tau' must be a real number > 0")
        if not operators:
            raise ValueError("At least one SpinOperator must be present in 'operators'")
        r = _dense(rho0)
        return np.array([self._run(_lib.MODE_LINDBLAD_INT, r, None, float(tau), o)[0] * tau for o in operators]).astype(complex)
