"""Celio's method on the GPU (csrc/celio.cuh through musim_celio_evolve) against the reference's C++
extension (muspinsim.cpp.celio_evolve, when oracle/_ref travelled) and the numpy oracle of it, on
FIXED initial states (celio.py:309 randomises phases, so parity is defined per state), plus a seeded
run of the whole `fast_evolve` against the reference's.  Tolerance 1e-9 (FP64)."""
import numpy as np
import pytest

from test_celio_host import _systems

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _setup(kind, k, dt):
    from muspinsim_b200 import celio
    from muspinsim_b200.spinsys import MuonSpinSystem

    spins, build = _systems(kind)
    s = build(MuonSpinSystem(spins))
    H = celio.CelioHamiltonian(celio.terms_from_system(s), k, s)
    return s, H, H._calc_trotter_evol_op_contribs(dt)


@pytest.mark.parametrize("streamed", [False, True])
@pytest.mark.parametrize("kind,k", [("mu_F_F", 2), ("mu_2V", 3), ("mu_e_H_N", 4)])
def test_fixed_states_match_oracle_and_reference(kind, k, streamed):
    from muspinsim_b200 import _lib
    from oracle import muspin_oracle as mo
    from oracle import ref_driver

    s, H, gates = _setup(kind, k, 0.03)
    dim = s.dim_total
    rng = np.random.default_rng(11)
    n_states, nt = 5, 40
    psi = rng.normal(size=(n_states, dim)) + 1j * rng.normal(size=(n_states, dim))
    psi /= np.linalg.norm(psi, axis=1)[:, None]
    sigma = s.sigma_mu([0.3, -0.5, 0.8])
    want = np.zeros(nt)
    for st in range(n_states):
        mo.celio_evolve(nt, psi[st].copy(), sigma, dim // 2, k, gates, want)
    got = np.zeros(nt)
    l0 = _lib.load().musim_celio_launch_count()
    _lib.celio_evolve(0, psi, sigma, k, gates, nt, got, streamed=streamed)
    assert _lib.load().musim_celio_launch_count() > l0
    assert np.max(np.abs(got - want)) < TOL
    # accumulate semantics: a second call adds
    _lib.celio_evolve(0, psi, sigma, k, gates, nt, got, streamed=streamed)
    assert np.max(np.abs(got - 2 * want)) < TOL
    if ref_driver.available():
        ref_driver._import()
        if not hasattr(np, "product"):
            np.product = np.prod
        from muspinsim.cpp import Celio_EvolveContrib, celio_evolve

        ref = np.zeros(nt)
        contribs = [Celio_EvolveContrib(U, od, idx.astype(np.uint64)) for (U, od, idx) in gates]
        for st in range(n_states):
            celio_evolve(nt, psi[st].copy().reshape(-1, 1), np.ascontiguousarray(sigma), dim // 2, k, contribs, ref)
        assert np.max(np.abs(got - 2 * ref)) < TOL


def test_large_gate_and_resident_limit():
    """Two 51V (a 64 x 64 dipolar gate: the generic gate path) and the examples/celio-sized system
    mu + 4 x 51V (dim = 8 192: the shared-memory resident path at its design size)."""
    from muspinsim_b200 import _lib, celio
    from muspinsim_b200.spinsys import MuonSpinSystem
    from oracle import muspin_oracle as mo

    s = MuonSpinSystem(["mu", "V", "V"])
    s.add_dipolar_term(1, 2, [0.5, 0.4, 2.1])  # couples the two I = 7/2 nuclei: 64 x 64
    s.add_dipolar_term(0, 1, [0.0, 0.0, 1.5])
    H = celio.CelioHamiltonian(celio.terms_from_system(s), 2, s)
    gates = H._calc_trotter_evol_op_contribs(0.05)
    assert max(g[0].shape[0] for g in gates) == 64
    dim = s.dim_total
    rng = np.random.default_rng(2)
    psi = rng.normal(size=(2, dim)) + 1j * rng.normal(size=(2, dim))
    psi /= np.linalg.norm(psi, axis=1)[:, None]
    sigma = s.sigma_mu([1.0, 0.0, 0.0])
    want, got = np.zeros(12), np.zeros(12)
    for st in range(2):
        mo.celio_evolve(12, psi[st].copy(), sigma, dim // 2, 2, gates, want)
    _lib.celio_evolve(0, psi, sigma, 2, gates, 12, got)
    assert np.max(np.abs(got - want)) < TOL

    s = MuonSpinSystem(["mu", "V", "V", "V", "V"])
    pos = [[0.0, 0.0, 1.6], [1.6, 0.0, 0.0], [0.0, -1.6, 0.0], [-1.1, 1.1, 0.3]]
    for i, r in enumerate(pos):
        s.add_dipolar_term(0, i + 1, r)
        s.add_quadrupolar_term(i + 1, np.diag([0.2, 0.3, -0.5]))
    H = celio.CelioHamiltonian(celio.terms_from_system(s), 2, s)
    gates = H._calc_trotter_evol_op_contribs(0.1)
    dim = s.dim_total
    assert dim == 8192
    psi = H.initial_states(sigma, 3)
    want, got, got2 = np.zeros(10), np.zeros(10), np.zeros(10)
    for st in range(3):
        mo.celio_evolve(10, psi[st].copy(), sigma, dim // 2, 2, gates, want)
    _lib.celio_evolve(0, psi, sigma, 2, gates, 10, got)
    _lib.celio_evolve(0, psi, sigma, 2, gates, 10, got2, streamed=True)
    assert np.max(np.abs(got - want)) < TOL and np.max(np.abs(got2 - want)) < TOL


def test_seeded_fast_evolve_reproduces_the_reference():
    """CelioHamiltonian.fast_evolve with numpy's global generator seeded like the reference's run
    (celio.py:289-316 draws np.random.rand(half_dim) per average)."""
    from oracle import ref_driver

    if not ref_driver.available():
        pytest.skip("oracle/_ref not present")
    ref_driver._import()
    if not hasattr(np, "product"):
        np.product = np.prod
    from muspinsim.spinsys import MuonSpinSystem as RefSystem

    s, H, _ = _setup("mu_2V", 3, 0.05)
    spins, build = _systems("mu_2V")
    Href = build(RefSystem(spins, celio_k=3)).hamiltonian
    times = np.linspace(0.0, 1.0, 21)
    sig = s.sigma_mu([0.0, 0.6, 0.8])
    np.random.seed(1234)
    want = Href.fast_evolve(build(RefSystem(spins, celio_k=3)).sigma_mu([0.0, 0.6, 0.8]), times, 6, True)
    np.random.seed(1234)
    got = H.fast_evolve(sig, times, 6)
    assert got.shape == want.shape == (21,)
    assert np.max(np.abs(got - np.real(want))) < TOL


def test_experiment_runner_celio_keyword():
    """spec["celio"] = [k, averages]: ExperimentRunner evaluates every configuration with Celio's
    method on the GPU (experiment.py:454-470); with many averages and a fine Trotter step the powder
    signal approaches the exact (diagonalisation) one."""
    from muspinsim_b200 import ExperimentRunner

    spec = {"name": "celio_fmuf", "spins": ["mu", "F", "F"],
            "couplings": [{"type": "dipolar", "i": 1, "j": 2, "value": [0.0, 0.0, 1.17]},
                          {"type": "dipolar", "i": 1, "j": 3, "value": [0.0, 0.0, -1.17]},
                          {"type": "dipolar", "i": 2, "j": 3, "value": [0.0, 0.0, 2.34]}],
            "time": list(np.linspace(0.0, 4.0, 41)), "orientation": [[0.0, 0.0, 0.0, 1.0], [0.4, 0.9, 0.1, 2.0]],
            "field": [[0.0, 0.0, 0.002]]}
    exact = ExperimentRunner(spec, device=0).run()
    np.random.seed(7)
    approx = ExperimentRunner(dict(spec, celio=[8, 64]), device=0).run()
    assert approx.shape == exact.shape
    assert np.max(np.abs(approx - exact)) < 0.05  # statistical: 4-dimensional bath, 64 random states
    with pytest.raises(NotImplementedError):
        ExperimentRunner(dict(spec, celio=[8, 0]), device=0).run()  # density-matrix Celio stays with the reference
