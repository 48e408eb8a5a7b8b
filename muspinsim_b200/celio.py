"""Celio's method (Phys. Rev. Lett. 56, 2720 (1986)) with the state-vector evolution on the GPU.

Mirrors /root/reference/muspinsim/celio.py: `CelioHamiltonian(terms, k, spinsys)` with
`_calc_H_contribs` (celio.py:73-144), `_calc_trotter_evol_op_contribs` (celio.py:146-205) and
`fast_evolve(sigma_mu, times, averages)` (celio.py:318-476) -- same names, argument meaning,
validation errors (validation.py:48-71) and random-state construction (`_compute_psi`,
celio.py:289-316: the reference draws `np.random.rand(half_dim)` per average from numpy's global
generator; so does this module, so that a seeded run reproduces the reference's numbers).

The gate matrices (exponentials of the small per-term Hamiltonians) are host set-up work; the hot
loop -- averages x times x k x gates applications of a gate to a 2^n-like state vector plus the
measurement -- is ONE call of `musim_celio_evolve` (csrc/celio.cuh), all random states at once.

A term is `(indices, matrix)`: `indices` the tuple of spin indices it couples ((i,), (i, j) or
(i, i) for a quadrupolar term) and `matrix` its operator on the product space of ONLY those spins,
Kronecker factors in ascending spin order (the reference builds it with
`spinsys.operator(..., include_only_given=True)`, spinsys.py:545-592).  As in the reference, the
index map of a contribution uses `spin_order = list(indices) + uninvolved` (celio.py:127-133), so
terms are expected with ascending indices -- what the reference's own term builders produce for a
muon-first system.
"""

import itertools

import numpy as np
import scipy.linalg as sla

from . import _lib


def validate_celio_params(terms, times):
    """validation.py:48-71."""
    if len(terms) == 0:
        raise ValueError("No interaction terms to evolve")
    if times[0] != 0:
        raise ValueError("Cannot use Celio's method with a non-zero start time")
    differences = np.diff(times)
    if not np.isclose(differences, differences[0]).all():
        raise ValueError("Cannot use Celio's method with uneven spacing between times")


def term_matrix(system, indices, tensor):
    """Operator of one interaction term on the spins it involves only (dense complex)."""
    tensor = np.asarray(tensor, dtype=float)
    idx = tuple(int(i) for i in indices)
    ops = system._local  # [(Sx, Sy, Sz)] per spin
    if len(idx) == 1:
        return sum(tensor[a] * ops[idx[0]][a] for a in range(3))
    i, j = idx
    if i == j:
        return sum(tensor[a, b] * (ops[i][a] @ ops[i][b]) for a in range(3) for b in range(3))
    out = 0
    for a in range(3):
        for b in range(3):
            # tensor[a, b] multiplies S_i^a S_j^b; the Kronecker product runs in ascending spin order
            fa, fb = (ops[i][a], ops[j][b]) if i < j else (ops[j][b], ops[i][a])
            out = out + tensor[a, b] * np.kron(fa, fb)
    return out


def terms_from_system(system):
    """[(indices, matrix)] of every interaction term of a muspinsim_b200 SpinSystem."""
    return [(tuple(ind), term_matrix(system, ind, ten)) for (_, ind, ten) in system._terms]


class CelioHContrib:
    """celio.py:28-52."""

    def __init__(self, matrix, other_dimension, spin_order, spin_dimensions):
        self.matrix = matrix
        self.other_dimension = int(other_dimension)
        self.spin_order = list(spin_order)
        self.spin_dimensions = list(spin_dimensions)


class CelioHamiltonian:
    def __init__(self, terms, k, spinsys, device=0):
        """terms: [(indices, matrix)]; k: Trotter factor; spinsys: object with `.dimension` (tuple of
        single-spin dimensions) and `.muon_index`."""
        self._terms = list(terms)
        self._k = int(k)
        self._spinsys = spinsys
        self._device = device

    def __add__(self, x):
        return CelioHamiltonian(self._terms + x._terms, self._k, self._spinsys, self._device)

    # ---- celio.py:73-144 ----
    def _calc_H_contribs(self):
        dims = tuple(int(n) for n in self._spinsys.dimension)
        n_spins = len(dims)
        contribs = []
        for i in range(n_spins):
            spin_ints = [t for t in self._terms if i == t[0][0]]
            other_spins = [s for s in range(n_spins) if s != i]
            if not spin_ints:
                continue
            for indices, group in itertools.groupby(spin_ints, lambda t: tuple(t[0])):
                group = list(group)
                H = sum(np.asarray(t[1], dtype=complex) for t in group)
                uninvolved = [s for s in other_spins if s not in indices]
                other_dimension = int(np.prod([dims[s] for s in uninvolved])) if uninvolved else 1
                ind = list(indices)
                if len(ind) == 2 and ind[0] == ind[1]:
                    ind.pop()  # quadrupolar term: the spin appears once in the ordering
                spin_order = ind + uninvolved
                contribs.append(CelioHContrib(H, other_dimension, spin_order, [dims[s] for s in spin_order]))
        return contribs

    # ---- celio.py:146-205 (the cpp=True branch: matrix, other dimension, index map) ----
    def _calc_trotter_evol_op_contribs(self, time_step):
        dims = tuple(int(n) for n in self._spinsys.dimension)
        total = int(np.prod(dims))
        out = []
        for c in self._calc_H_contribs():
            U = sla.expm(-2j * np.pi * np.asarray(c.matrix) * time_step / self._k)
            idx = np.transpose(np.arange(total, dtype=np.int64).reshape(dims), axes=c.spin_order).flatten()
            out.append((np.ascontiguousarray(U), c.other_dimension, idx))
        return out

    # ---- celio.py:289-316 ----
    @staticmethod
    def _compute_psi(mu_psi, half_dim):
        psi0 = np.exp(2j * np.pi * np.random.rand(half_dim))
        return np.kron(np.asarray(mu_psi), psi0) * (1.0 / np.sqrt(half_dim))

    def initial_states(self, sigma_mu, averages):
        """The `averages` random initial states the reference would draw, in order: [averages, dim]."""
        sig = _dense2(sigma_mu)
        evals, evecs = np.linalg.eig(sig + np.eye(2))  # celio.py:375-376
        mu_psi = evecs[:, 1] if evals[1] > 0.1 else evecs[:, 0]
        half_dim = int(np.prod(self._spinsys.dimension)) // 2
        return np.array([self._compute_psi(mu_psi, half_dim) for _ in range(averages)])

    # ---- celio.py:318-476 ----
    def fast_evolve(self, sigma_mu, times, averages, psi=None, streamed=False):
        """Muon polarisation (range -0.5 .. 0.5) at `times`, averaged over `averages` random initial
        states (or over the rows of `psi` if given: deterministic, for parity tests)."""
        times = np.array(times)
        if not isinstance(times, np.ndarray) or times.ndim != 1:
            raise ValueError("times must be an array of values in microseconds")
        validate_celio_params(self._terms, times)
        if psi is None and averages <= 0:
            raise ValueError("averages must be a positive integer")
        if self._spinsys.muon_index != 0:
            raise ValueError("Muon must be the first spin in the system in order to use the fast Celio method")
        time_step = times[1] - times[0]
        sig = _dense2(sigma_mu)
        if psi is None:
            psi = self.initial_states(sig, int(averages))
        psi = np.atleast_2d(np.asarray(psi, dtype=complex))
        contribs = self._calc_trotter_evol_op_contribs(time_step)
        results = np.zeros(times.shape[0], dtype=np.float64)
        _lib.celio_evolve(self._device, psi, sig, self._k, contribs, times.shape[0], results, streamed=streamed)
        return results * (1.0 / psi.shape[0]) * 0.5


def _dense2(sigma_mu):
    m = getattr(sigma_mu, "matrix", sigma_mu)
    if hasattr(m, "toarray"):
        m = m.toarray()
    m = np.asarray(m, dtype=complex)
    if m.shape != (2, 2):
        raise ValueError("sigma_mu must be a 2x2 matrix")
    return m
