"""Celio's method, host side (no GPU): the contributions / gate matrices / index maps built by
muspinsim_b200.celio.CelioHamiltonian equal the reference's (celio.py:73-205), and the numpy oracle
of the C++ kernels (oracle.muspin_oracle.celio_evolve) is pinned against the reference's own compiled
extension (muspinsim.cpp.celio_evolve) on fixed initial states.  Reference-dependent parts are skipped
where oracle/_ref is absent."""
import numpy as np
import pytest

from muspinsim_b200 import celio
from muspinsim_b200.spinsys import MuonSpinSystem
from oracle import muspin_oracle as mo


def _systems(kind):
    """(ours, builder for the reference): mu + nuclei with dipolar, quadrupolar and Zeeman terms."""
    if kind == "mu_2V":
        spins = ["mu", "V", "V"]
    elif kind == "mu_e_H_N":
        spins = ["mu", "e", "H", ("N", 14)]
    else:
        spins = ["mu", "F", "F"]

    def build(sys_):
        if kind == "mu_2V":
            sys_.add_dipolar_term(0, 1, [0.0, 0.0, 1.6])
            sys_.add_dipolar_term(0, 2, [1.1, 0.3, -1.2])
            sys_.add_quadrupolar_term(1, [[0.2, 0.05, 0.0], [0.05, -0.5, 0.1], [0.0, 0.1, 0.3]])
            sys_.add_quadrupolar_term(2, [[-0.1, 0.0, 0.02], [0.0, 0.3, 0.0], [0.02, 0.0, -0.2]])
            sys_.add_zeeman_term(0, [0.0, 0.0, 0.01])
            sys_.add_zeeman_term(1, [0.0, 0.0, 0.01])
        elif kind == "mu_e_H_N":
            sys_.add_hyperfine_term(0, np.diag([10.0, 12.0, 15.0]) + 0.5)
            sys_.add_hyperfine_term(2, np.diag([3.0, 3.0, 4.0]))
            sys_.add_dipolar_term(0, 2, [1.5, 0.2, 0.4])
            sys_.add_dipolar_term(2, 3, [0.3, 1.1, 0.9])
            sys_.add_quadrupolar_term(3, [[0.3, 0.0, 0.1], [0.0, -0.6, 0.0], [0.1, 0.0, 0.3]])
            sys_.add_zeeman_term(1, [0.02, 0.0, 0.01])
        else:
            sys_.add_dipolar_term(0, 1, [0.0, 0.0, 1.17])
            sys_.add_dipolar_term(0, 2, [0.0, 0.0, -1.17])
            sys_.add_dipolar_term(1, 2, [0.0, 0.0, 2.34])
        return sys_

    return spins, build


def _reference():
    from oracle import ref_driver

    if not ref_driver.available():
        pytest.skip("oracle/_ref not present")
    ms = ref_driver._import()
    if not hasattr(np, "product"):  # the reference's celio.py still calls np.product (removed in numpy 2)
        np.product = np.prod
    return ms


@pytest.mark.parametrize("kind", ["mu_2V", "mu_e_H_N", "mu_F_F"])
def test_contributions_and_gates_match_the_reference(kind):
    ms = _reference()
    spins, build = _systems(kind)
    ours = build(MuonSpinSystem(spins))
    from muspinsim.spinsys import MuonSpinSystem as RefSystem

    ref_sys = build(RefSystem(spins, celio_k=4))
    Href = ref_sys.hamiltonian  # CelioHamiltonian (spinsys.py:613-626)
    H = celio.CelioHamiltonian(celio.terms_from_system(ours), 4, ours)
    a, b = H._calc_H_contribs(), Href._calc_H_contribs()
    assert len(a) == len(b) and len(a) > 0
    for ca, cb in zip(a, b):
        assert ca.other_dimension == cb.other_dimension
        assert ca.spin_order == list(cb.spin_order) and ca.spin_dimensions == list(cb.spin_dimensions)
        assert np.max(np.abs(np.asarray(ca.matrix) - cb.matrix.toarray())) < 1e-12
    dt = 0.05
    ga, gb = H._calc_trotter_evol_op_contribs(dt), Href._calc_trotter_evol_op_contribs(dt, True)
    for (U, od, idx), ref in zip(ga, gb):
        assert od == ref.other_dim and np.array_equal(idx, np.asarray(ref.indices))
        assert np.max(np.abs(U - np.asarray(ref.matrix))) < 1e-12


@pytest.mark.parametrize("kind", ["mu_2V", "mu_F_F"])
def test_numpy_oracle_of_the_cpp_kernels_is_pinned(kind):
    """oracle.celio_evolve vs muspinsim.cpp.celio_evolve on the same fixed state."""
    ms = _reference()
    from muspinsim.cpp import Celio_EvolveContrib, celio_evolve

    spins, build = _systems(kind)
    ours = build(MuonSpinSystem(spins))
    H = celio.CelioHamiltonian(celio.terms_from_system(ours), 3, ours)
    gates = H._calc_trotter_evol_op_contribs(0.04)
    dim = ours.dim_total
    rng = np.random.default_rng(3)
    psi = rng.normal(size=dim) + 1j * rng.normal(size=dim)
    psi /= np.linalg.norm(psi)
    sigma = ours.sigma_mu([0.3, -0.5, 0.8])
    nt = 25
    want = np.zeros(nt)
    celio_evolve(nt, psi.copy().reshape(-1, 1), np.ascontiguousarray(sigma), dim // 2, 3,
                 [Celio_EvolveContrib(U, od, idx.astype(np.uint64)) for (U, od, idx) in gates], want)
    got = np.zeros(nt)
    mo.celio_evolve(nt, psi.copy(), sigma, dim // 2, 3, gates, got)
    assert np.max(np.abs(got - want)) < 1e-12
    assert np.max(np.abs(want)) > 0.01


def test_validation_errors_mirror_the_reference():
    spins, build = _systems("mu_F_F")
    s = build(MuonSpinSystem(spins))
    H = celio.CelioHamiltonian(celio.terms_from_system(s), 2, s)
    sig = s.sigma_mu([1.0, 0.0, 0.0])
    with pytest.raises(ValueError):
        H.fast_evolve(sig, np.linspace(0.1, 1.0, 10), 4)  # non-zero start time
    with pytest.raises(ValueError):
        H.fast_evolve(sig, np.array([0.0, 0.1, 0.3]), 4)  # uneven spacing
    with pytest.raises(ValueError):
        H.fast_evolve(sig, np.linspace(0.0, 1.0, 10), 0)  # averages
    with pytest.raises(ValueError):
        celio.CelioHamiltonian([], 2, s).fast_evolve(sig, np.linspace(0.0, 1.0, 10), 4)  # no terms
    s2 = MuonSpinSystem(["F", "mu"])
    s2.add_dipolar_term(0, 1, [0.0, 0.0, 1.2])
    with pytest.raises(ValueError):
        celio.CelioHamiltonian(celio.terms_from_system(s2), 2, s2).fast_evolve(sig, np.linspace(0.0, 1.0, 10), 4)


def test_oracle_trotter_converges_to_the_exact_evolution():
    """Sanity of the whole construction without any reference: for growing k the Trotter evolution
    of a fixed state approaches exp(-2 pi i H t) applied to it."""
    spins, build = _systems("mu_F_F")
    s = build(MuonSpinSystem(spins))
    dim = s.dim_total
    rng = np.random.default_rng(0)
    psi = rng.normal(size=dim) + 1j * rng.normal(size=dim)
    psi /= np.linalg.norm(psi)
    sig = s.sigma_mu([1.0, 0.0, 0.0])
    O = np.kron(sig, np.eye(dim // 2))
    times = np.linspace(0.0, 2.0, 21)
    lam, U = np.linalg.eigh(s.hamiltonian)
    exact = []
    for t in times:
        v = U @ (np.exp(-2j * np.pi * lam * t) * (U.conj().T @ psi))
        exact.append(np.real(np.vdot(v, O @ v)))
    errs = []
    for k in (1, 4, 16):
        H = celio.CelioHamiltonian(celio.terms_from_system(s), k, s)
        got = np.zeros(len(times))
        mo.celio_evolve(len(times), psi.copy(), sig, dim // 2, k, H._calc_trotter_evol_op_contribs(times[1]), got)
        errs.append(np.max(np.abs(got - np.array(exact))))
    assert errs[2] < errs[1] < errs[0] and errs[2] < 5e-3
