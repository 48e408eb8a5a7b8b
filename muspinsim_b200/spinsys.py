"""Host-side spin system: the one-off setup that produces the constant dense operators the
GPU path consumes (H0, Z_a = sum_i gamma_i S_i^a, M_a = S_mu^a (x) 1).

Mirrors the public surface of the reference's `MuonSpinSystem`
(/root/reference/muspinsim/spinsys.py:159-760): same method names, argument meaning (0-based
spin indices, tensors in MHz, distances in Angstrom, fields in T) and error behaviour.  The
implementation is dense numpy (d <= a few hundred on this path); nothing here runs per
configuration.
"""

from numbers import Number

import numpy as np
import scipy.constants as cnst

from .constants import EFG_2_MHZ, gyromagnetic_ratio, parse_spin, quadrupole_moment, spin


def spin_operators(I):
    """(Sx, Sy, Sz) for spin I, basis ordered m = +I ... -I (spinop.py:14-41)."""
    n = int(round(2 * I + 1))
    m = I - np.arange(n)
    # <m+1|S+|m> = sqrt(I(I+1) - m(m+1)) on the first super-diagonal
    sp = np.zeros((n, n), dtype=complex)
    for a in range(n - 1):
        sp[a, a + 1] = np.sqrt(I * (I + 1) - m[a + 1] * (m[a + 1] + 1))
    sx = 0.5 * (sp + sp.T)
    sy = -0.5j * (sp - sp.T)
    sz = np.diag(m).astype(complex)
    return sx, sy, sz


class SpinSystem:
    def __init__(self, spins=None):
        spins = list(spins or [])
        self._spins = spins
        parsed = [parse_spin(s) for s in spins]
        self._gammas = np.array([gyromagnetic_ratio(e, i) for e, i in parsed], dtype=float)
        self._Qs = np.array([quadrupole_moment(e, i) for e, i in parsed], dtype=float)
        self._Is = np.array([spin(e, i) for e, i in parsed], dtype=float)
        self._dim = tuple(int(round(2 * I + 1)) for I in self._Is)
        self._local = [spin_operators(I) for I in self._Is]
        self._terms = []  # (label, indices, tensor)
        n = self.dim_total
        self._H = np.zeros((n, n), dtype=complex)

    # ---- properties (spinsys.py:215-247) ----
    @property
    def spins(self):
        return list(self._spins)

    @property
    def gammas(self):
        return self._gammas.copy()

    @property
    def Qs(self):
        return self._Qs.copy()

    @property
    def Is(self):
        return self._Is.copy()

    @property
    def dimension(self):
        return self._dim

    @property
    def dim_total(self):
        return int(np.prod(self._dim)) if self._dim else 1

    def gamma(self, i):
        return self._gammas[i]

    def Q(self, i):
        return self._Qs[i]

    def I(self, i):
        return self._Is[i]

    def __len__(self):
        return len(self._spins)

    # ---- operators ----
    def operator(self, terms=None):
        """Dense operator for {spin index: 'x'|'y'|'z'|'0' or a list of those (matrix product)};
        identity on every other spin; Kronecker order = spin order (spinsys.py:545-592)."""
        terms = terms or {}
        out = np.eye(1, dtype=complex)
        for i, n in enumerate(self._dim):
            sym = terms.get(i, "0")
            syms = sym if isinstance(sym, list) else [sym]
            loc = np.eye(n, dtype=complex)
            for s in syms:
                if s == "0":
                    continue
                loc = loc @ self._local[i]["xyz".index(s)]
            out = np.kron(out, loc)
        return out

    # ---- term builders (spinsys.py:249-442) ----
    def _check_index(self, i, name="i"):
        if i < 0 or i >= len(self._spins):
            raise ValueError(f"Invalid index {name}")

    def add_linear_term(self, i, vector, label="Single"):
        self._check_index(i)
        vector = np.array(vector, dtype=float)
        if vector.shape != (3,):
            raise ValueError("Tensor is not fully three-dimensional")
        for a in range(3):
            if vector[a] != 0.0:
                self._H += vector[a] * self.operator({i: "xyz"[a]})
        self._terms.append((label, (i,), vector))

    def add_bilinear_term(self, i, j, matrix, label="Double"):
        self._check_index(i)
        self._check_index(j, "j")
        matrix = np.array(matrix, dtype=float)
        if matrix.shape != (3, 3):
            raise ValueError("Tensor is not fully three-dimensional")
        if i == j:
            for a in range(3):
                for b in range(3):
                    if matrix[a, b] != 0.0:
                        self._H += matrix[a, b] * self.operator({i: ["xyz"[a], "xyz"[b]]})
        else:
            # sum_ab T_ab S_i^a S_j^b is built on the spins lo .. hi only and embedded once (two Kronecker
            # products with identities per TERM instead of one chain over all spins per tensor ELEMENT:
            # building the d = 96 benchmark system took 10 ms of host time per runner)
            lo, hi = (i, j) if i < j else (j, i)
            T = matrix if i < j else matrix.T  # T[a, b] multiplies S_lo^a S_hi^b
            mid = int(np.prod(self._dim[lo + 1:hi], dtype=np.int64)) if hi > lo + 1 else 1
            small = np.zeros((self._dim[lo] * mid * self._dim[hi],) * 2, dtype=complex)
            eye_mid = np.eye(mid, dtype=complex)
            for a in range(3):
                left = np.kron(self._local[lo][a], eye_mid)
                for b in range(3):
                    if T[a, b] != 0.0:
                        small += T[a, b] * np.kron(left, self._local[hi][b])
            nl = int(np.prod(self._dim[:lo], dtype=np.int64)) if lo > 0 else 1
            nr = int(np.prod(self._dim[hi + 1:], dtype=np.int64)) if hi + 1 < len(self._dim) else 1
            self._H += np.kron(np.kron(np.eye(nl, dtype=complex), small), np.eye(nr, dtype=complex))
        self._terms.append((label, (i, j), matrix))

    def add_zeeman_term(self, i, B):
        if isinstance(B, Number):
            B = [0, 0, B]
        return self.add_linear_term(i, np.array(B, dtype=float) * self.gamma(i), "Zeeman")

    def add_dipolar_term(self, i, j, r):
        if i == j:
            raise ValueError("Can not set up dipolar coupling with itself")
        r = np.array(r, dtype=float)
        rn = np.linalg.norm(r)
        D = -(np.eye(3) - 3.0 / rn**2.0 * np.outer(r, r))
        dij = -(cnst.mu_0 * cnst.hbar * (self.gamma(i) * self.gamma(j) * 1e6)) / (2 * (rn * 1e-10) ** 3)
        return self.add_bilinear_term(i, j, D * dij, "Dipolar")

    def add_quadrupolar_term(self, i, EFG):
        I = self.I(i)
        if I == 0.5:
            raise ValueError("Can not set up quadrupolar coupling for spin 1/2 particle")
        Qt = EFG_2_MHZ * self.Q(i) / (2 * I * (2 * I - 1)) * np.array(EFG, dtype=float)
        return self.add_bilinear_term(i, i, Qt, "Quadrupolar")

    @property
    def hamiltonian(self):
        """Field-independent Hamiltonian H0, dense complex (MHz).  spinsys.py:613-626."""
        return self._H.copy()


class MuonSpinSystem(SpinSystem):
    def __init__(self, spins=("mu", "e")):
        super().__init__(list(spins))
        if self._spins.count("mu") != 1:
            raise ValueError("Spins passed to MuonSpinSystem must contain exactly one muon")
        self._mu_i = self._spins.index("mu")
        self._e_i = {i for i, s in enumerate(self._spins) if s == "e"}

    @property
    def muon_index(self):
        return self._mu_i

    @property
    def elec_indices(self):
        return self._e_i

    def add_hyperfine_term(self, i, A, j=None):
        """spinsys.py:664-705."""
        if j is None:
            if len(self._e_i) > 1:
                raise ValueError("Must specify an electron index in system with multiple electrons")
            if len(self._e_i) == 0:
                raise ValueError("No electron in the system")
            j = next(iter(self._e_i))
        elif j not in self._e_i:
            raise ValueError("Second index in hyperfine coupling must refer to an electron")
        if i in self._e_i:
            raise ValueError("First index in hyperfine coupling must not refer to an electron")
        return self.add_bilinear_term(i, j, A, "Hyperfine")

    def muon_operator(self, v):
        """sum_a v_a S_mu^a (x) 1.  spinsys.py:707-732."""
        if len(v) != 3:
            raise ValueError("Vector passed to muon_operator must be three dimensional")
        M = self.muon_operators()
        return v[0] * M[0] + v[1] * M[1] + v[2] * M[2]

    def sigma_mu(self, v):
        """2x2 sum_a v_a sigma_a (no factor 1/2).  spinsys.py:734-760."""
        if len(v) != 3:
            raise ValueError("Vector passed to muon_operator must be three dimensional")
        sx, sy, sz = spin_operators(0.5)
        return 2.0 * (v[0] * sx + v[1] * sy + v[2] * sz)

    # ---- the constant operators of the GPU path ----
    def muon_operators(self):
        """M[3, d, d] = S_mu^a (x) 1."""
        return np.array([self.operator({self._mu_i: a}) for a in "xyz"])

    def zeeman_operators(self):
        """Z[3, d, d] = sum_i gamma_i S_i^a, so that Hz(B) = sum_a B_a Z_a (experiment.py:238-250)."""
        n = self.dim_total
        Z = np.zeros((3, n, n), dtype=complex)
        for i in range(len(self)):
            for a in range(3):
                Z[a] += self._gammas[i] * self.operator({i: "xyz"[a]})
        return Z


def system_from_spec(spec):
    """Build a MuonSpinSystem from a spec's `spins` / `couplings` (1-based indices, as in the
    .in file; simconfig.py:253-291).  Returns (system, {spin index: dissipation rate})."""
    sys_ = MuonSpinSystem(spec["spins"])
    dissip = {}
    for c in spec.get("couplings", []):
        i = c["i"] - 1
        j = c.get("j")
        j = j - 1 if j is not None else None
        if i < 0 or i >= len(sys_) or (j is not None and (j < 0 or j >= len(sys_))):
            raise ValueError("Out of range indices for coupling")
        t = c["type"]
        if t == "zeeman":
            sys_.add_zeeman_term(i, np.asarray(c["value"], dtype=float))
        elif t == "dipolar":
            sys_.add_dipolar_term(i, j, c["value"])
        elif t == "hyperfine":
            sys_.add_hyperfine_term(i, c["value"], j)
        elif t == "quadrupolar":
            sys_.add_quadrupolar_term(i, c["value"])
        elif t == "dissipation":
            dissip[i] = float(np.atleast_1d(c["value"])[0])
        else:
            raise ValueError(f"unknown {t} coupling")
    return sys_, dissip
