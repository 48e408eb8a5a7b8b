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
