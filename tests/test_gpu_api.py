"""The per-call boundary (muspinsim_b200.hamiltonian.Hamiltonian) with the reference's own
known-answer tests and error behaviour (tests/test_hamiltonian.py:10-128, validation.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sx_sz():
    from muspinsim_b200.spinsys import spin_operators

    sx, sy, sz = spin_operators(0.5)
    return sx, sz


def test_creation_errors():
    from muspinsim_b200.hamiltonian import Hamiltonian

    H = Hamiltonian(np.array([[1, 0], [0, -1]]))
    assert H.dimension == (2,)
    with pytest.raises(ValueError):
        Hamiltonian(np.array([[1, 1], [0, 1]]))


def test_diag():
    from muspinsim_b200.hamiltonian import Hamiltonian

    sx, sz = _sx_sz()
    evals, evecs = Hamiltonian(sx).diag()
    assert np.allclose(evals, [-0.5, 0.5], atol=1e-15)
    evecsT = np.array([[1.0, 1.0], [-1.0, 1.0]]) / 2**0.5
    assert np.allclose(abs(np.dot(evecs, evecsT)), np.eye(2))


def test_evolve_fast_evolve_integrate():
    from muspinsim_b200.hamiltonian import Hamiltonian

    sx, sz = _sx_sz()
    rho0 = 0.5 * np.eye(2) + sz
    t = np.linspace(0, 1, 100)
    H = Hamiltonian(sx)
    evol = H.evolve(rho0, t, [sz])
    assert evol.shape == (100, 1)
    assert np.allclose(evol[:, 0], 0.5 * np.cos(2 * np.pi * t), atol=1e-12)
    avg = H.integrate_decaying(rho0, 1.0, [sz])
    assert np.isclose(avg[0], 0.5 / (1.0 + 4 * np.pi**2), atol=1e-13)
    H2 = Hamiltonian(np.kron(sx, np.eye(2)))
    fe = H2.fast_evolve(2 * sz, t, 2)
    assert np.allclose(fe, 0.5 * np.cos(2 * np.pi * t), atol=1e-12)


def test_validation_errors():
    from muspinsim_b200.hamiltonian import Hamiltonian

    sx, sz = _sx_sz()
    H = Hamiltonian(sx)
    rho0 = 0.5 * np.eye(2) + sz
    with pytest.raises(ValueError):
        H.evolve(rho0, np.zeros((2, 2)), [sz])  # times not 1-D
    with pytest.raises(ValueError):
        H.integrate_decaying(rho0, -1.0, [sz])
    with pytest.raises(ValueError):
        H.integrate_decaying(rho0, 1.0, [])
    with pytest.raises(ValueError):
        H.evolve(np.eye(3), np.linspace(0, 1, 3), [sz])  # incompatible rho0


def test_lindbladian_closed_forms():
    """tests/test_lindbladian.py:40-80 and :130-159 of the reference, through the GPU path."""
    from muspinsim_b200.hamiltonian import Hamiltonian
    from muspinsim_b200.lindbladian import Lindbladian
    from muspinsim_b200.spinsys import spin_operators

    sx, sy, sz = spin_operators(0.5)
    sp, sm = sx + 1j * sy, sx - 1j * sy
    H = Hamiltonian(sz)
    rho0 = 0.5 * np.eye(2) + sx
    t = np.linspace(0, 1, 100)
    L = Lindbladian.from_hamiltonian(H)
    assert np.allclose(L.evolve(rho0, t, sx)[:, 0], 0.5 * np.cos(2 * np.pi * t), atol=1e-11)
    tau = 2.0
    assert abs(L.integrate_decaying(rho0, tau, sx)[0] - 0.5 * tau / (1 + 4 * np.pi**2 * tau**2)) < 1e-12
    for g in [1.0, 2.0, 5.0, 10.0]:
        L = Lindbladian.from_hamiltonian(H, [(sx, g)])
        ap = -0.5 * np.pi * g + ((0.5 * np.pi * g) ** 2 - 4 * np.pi**2 + 0j) ** 0.5
        am = -0.5 * np.pi * g - ((0.5 * np.pi * g) ** 2 - 4 * np.pi**2 + 0j) ** 0.5
        A = ap * am / (am - ap)
        solx = np.real(0.5 * A * (np.exp(ap * t) / ap - np.exp(am * t) / am))
        assert np.allclose(L.evolve(rho0, t, sx)[:, 0], solx, atol=1e-10)
        sol = np.real(0.5 * A * tau * (1 / ((1 - ap * tau) * ap) - 1 / ((1 - am * tau) * am)))
        assert abs(L.integrate_decaying(rho0, tau, sx)[0] - sol) < 1e-11
        L = Lindbladian.from_hamiltonian(H, [(sp, 1.5 * g), (sm, 0.5 * g)])
        ev = L.evolve(rho0, t, [sx, sz])
        assert np.allclose(ev[:, 0], 0.5 * np.cos(2 * np.pi * t) * np.exp(-2 * np.pi * g * t), atol=1e-10)
        assert np.allclose(ev[:, 1], 0.25 * (1 - np.exp(-4 * np.pi * g * t)), atol=1e-10)
    with pytest.raises(ValueError):
        Lindbladian.from_hamiltonian(sz)
    with pytest.raises(ValueError):
        L.evolve(np.eye(3), t, sx)


def test_resident_handle_across_runners_with_different_couplings():
    """Fitting loop (fitting.py:126-135 builds a new runner per function evaluation): runners
    sharing a handle cache reuse ONE device handle, re-uploading only H0 / Z; results equal
    those of fresh handles, also when the runners are used alternately."""
    import numpy as np

    from muspinsim_b200 import ExperimentRunner, workloads

    cache = {}
    specs = []
    for scale in (1.0, 1.3, 0.6):
        spec = workloads.c2_hfine_powder(n_orient=9, nt=100, n_h=1)
        for c in spec["couplings"]:
            c["value"] = np.asarray(c["value"]) * scale
        specs.append(spec)
    runners = [ExperimentRunner(s, device=0, handle_cache=cache) for s in specs]
    got = [r.run() for r in runners]
    assert len(cache) == 1 and all(r.handle is runners[0].handle for r in runners)
    for s, g in zip(specs, got):
        assert np.max(np.abs(ExperimentRunner(s, device=0).run() - g)) < 1e-13
    assert np.max(np.abs(runners[0].run() - got[0])) < 1e-13  # first runner again after the others
    assert np.max(np.abs(got[0] - got[1])) > 1e-3
