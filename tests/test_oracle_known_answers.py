"""The reference's own analytic known-answer tests, restated against the oracle
(/root/reference/muspinsim/tests/...; file:line cited per test)."""
import numpy as np
import scipy.constants as cnst

from oracle import muspin_oracle as mo

SX, SY, SZ = mo.spin_matrices(0.5)


def test_spin_matrices():  # tests/test_spinop.py:41-62
    assert np.allclose(SX, [[0, 0.5], [0.5, 0]])
    assert np.allclose(SY, [[0, -0.5j], [0.5j, 0]])
    assert np.allclose(SZ, [[0.5, 0], [0, -0.5]])
    sx1, sy1, sz1 = mo.spin_matrices(1.0)
    assert np.allclose(sx1, np.array([[0, 1, 0], [1, 0, 1], [0, 1, 0]]) / 2**0.5)
    assert np.allclose(sz1, np.diag([1, 0, -1]))
    assert np.allclose(sx1 @ sy1 - sy1 @ sx1, 1j * sz1)


def test_evolve_precession():  # tests/test_hamiltonian.py:34-45
    t = np.linspace(0, 1, 100)
    rho0 = 0.5 * np.eye(2) + SZ
    evol = mo.evolve(SX.copy(), rho0, t, SZ)
    assert np.allclose(evol, 0.5 * np.cos(2 * np.pi * t))


def test_fast_evolve_precession():  # tests/test_hamiltonian.py:105-116
    t = np.linspace(0, 1, 100)
    H = np.kron(SX, np.eye(2))
    sig = 2 * SZ
    assert np.allclose(mo.fast_evolve(H, sig, t, 2), 0.5 * np.cos(2 * np.pi * t))


def test_integrate_decaying():  # tests/test_hamiltonian.py:118-128
    rho0 = 0.5 * np.eye(2) + SZ
    avg = mo.integrate_decaying(SX.copy(), rho0, 1.0, SZ)
    assert np.isclose(avg, 0.5 / (1.0 + 4 * np.pi**2))


def test_lindbladian_matrix():  # tests/test_lindbladian.py:10-38
    L = mo.superop_lindbladian(SZ.copy(), [])
    assert np.allclose(L, np.diag([0, -1j, 1j, 0]))


def test_lindblad_evolve_closed_forms():  # tests/test_lindbladian.py:40-80
    rho0 = 0.5 * np.eye(2) + SX
    t = np.linspace(0, 1, 100)
    L = mo.superop_lindbladian(SZ.copy(), [])
    assert np.allclose(mo.lindblad_evolve(L, rho0, t, SX), 0.5 * np.cos(2 * np.pi * t))
    sp, sm = SX + 1j * SY, SX - 1j * SY
    for g in [1.0, 2.0, 5.0, 10.0]:
        L = mo.superop_lindbladian(SZ.copy(), [(SX, g)])
        ap = -0.5 * np.pi * g + ((0.5 * np.pi * g) ** 2 - 4 * np.pi**2 + 0j) ** 0.5
        am = -0.5 * np.pi * g - ((0.5 * np.pi * g) ** 2 - 4 * np.pi**2 + 0j) ** 0.5
        A = ap * am / (am - ap)
        solx = np.real(0.5 * A * (np.exp(ap * t) / ap - np.exp(am * t) / am))
        assert np.allclose(mo.lindblad_evolve(L, rho0, t, SX), solx)
        L = mo.superop_lindbladian(SZ.copy(), [(sp, 1.5 * g), (sm, 0.5 * g)])
        assert np.allclose(mo.lindblad_evolve(L, rho0, t, SX), 0.5 * np.cos(2 * np.pi * t) * np.exp(-2 * np.pi * g * t))
        assert np.allclose(mo.lindblad_evolve(L, rho0, t, SZ), 0.25 * (1 - np.exp(-4 * np.pi * g * t)))


def test_lindblad_integrate_closed_forms():  # tests/test_lindbladian.py:130-159
    rho0 = 0.5 * np.eye(2) + SX
    tau = 2.0
    L = mo.superop_lindbladian(SZ.copy(), [])
    assert np.isclose(mo.lindblad_integrate(L, rho0, tau, SX), 0.5 * tau / (1 + 4 * np.pi**2 * tau**2))
    for g in [1.0, 2.0, 5.0, 10.0]:
        L = mo.superop_lindbladian(SZ.copy(), [(SX, g)])
        ap = -0.5 * np.pi * g + ((0.5 * np.pi * g) ** 2 - 4 * np.pi**2 + 0j) ** 0.5
        am = -0.5 * np.pi * g - ((0.5 * np.pi * g) ** 2 - 4 * np.pi**2 + 0j) ** 0.5
        A = ap * am / (am - ap)
        sol = np.real(0.5 * A * tau * (1 / ((1 - ap * tau) * ap) - 1 / ((1 - am * tau) * am)))
        assert np.isclose(mo.lindblad_integrate(L, rho0, tau, SX), sol)


def test_rho0_literals():  # tests/test_experiment.py:74-115
    s = mo.OracleSystem(["e", "mu"])
    rho0 = mo.rho0_matrix(s, np.zeros(3), np.array([1.0, 0, 0]), np.inf)
    assert np.allclose(rho0, [[0.25, 0.25, 0, 0], [0.25, 0.25, 0, 0], [0, 0, 0.25, 0.25], [0, 0, 0.25, 0.25]])
    T = 100
    B = np.array([0, 0, 2.0e-6 * cnst.k * T / (mo.ELEC_GAMMA * cnst.h)])
    rho0 = mo.rho0_matrix(s, B, np.array([0, 0, 1.0]), T)
    Z = np.exp([-1, 1])
    Z /= Z.sum()
    assert np.allclose(np.diag(rho0), [Z[0], 0, Z[1], 0])


def test_run_known_answers():  # tests/test_experiment.py:117-202
    t = np.linspace(0, 10, 100)
    res = mo.run_spec({"spins": ["e", "mu"], "time": t})
    assert np.all(res == 0.5)
    zee = [{"type": "zeeman", "i": 2, "value": [0, 0, 1.0 / mo.MU_GAMMA]}]
    res = mo.run_spec({"spins": ["e", "mu"], "time": t, "couplings": zee})
    assert np.allclose(res, 0.5 * np.cos(2 * np.pi * t))
    tau = mo.MU_TAU
    for key in ("field", "intrinsic_field"):
        res = mo.run_spec({"spins": ["e", "mu"], "couplings": zee, "y_axis": "integral", "x_axis": key,
                           key: [[0.0], [1.0]]})
        assert np.isclose(res[0], 0.5 / (1.0 + 4 * np.pi**2 * tau**2))


def test_dissipation_known_answers():  # tests/test_experiment.py:581-640
    g = 1.0
    t = np.linspace(0, 10, 101)
    res = mo.run_spec({"spins": ["mu"], "couplings": [{"type": "dissipation", "i": 1, "value": g}], "time": t})
    assert np.allclose(res, 0.5 * np.exp(-g * t))
    T = 0.1
    s = mo.build_system({"spins": ["mu"], "couplings": [{"type": "dissipation", "i": 1, "value": g}]})
    B = np.array([0, 0, -1.0])
    H = s.H0 + mo.zeeman_matrix(s, B)
    L = mo.superop_lindbladian(H, mo.dissipation_operators(s, B, T))
    rho0 = mo.rho0_matrix(s, B, np.array([1.0, 0, 0]), T)
    out = mo.lindblad_evolve(L, rho0, np.array([0.0, 20.0]), s.S(0, 2))
    Z = np.exp(-cnst.h * mo.MU_GAMMA * 1e6 / (cnst.k * T))
    assert np.isclose(np.real(out[-1]), 0.5 * (1 - Z) / (1 + Z))


def test_orientation_quaternions():  # tests/test_config.py:214-286, tests/test_utils.py:43-65
    q, w = mo.orientation_row([0.5 * np.pi, 0, 0, 2.0])
    assert np.allclose(q, [2**-0.5, 0, 0, -(2**-0.5)]) and w == 2.0
    rng = np.linspace(0, np.pi, 4)
    for a in rng:
        for b in rng:
            for c in rng:
                for mode, ax in (("zyz", [0, 1, 0]), ("zxz", [1, 0, 0])):
                    q2 = mo.quat_mul(mo.quat_mul(mo.quat_axis_angle([0, 0, 1], c), mo.quat_axis_angle(ax, b)),
                                     mo.quat_axis_angle([0, 0, 1], a))
                    q1, _ = mo.orientation_row([a, b, c], mode)
                    assert np.allclose(q1, q2 * [1, -1, -1, -1])
    theta, phi = 0.6 * np.pi, 0.4 * np.pi
    qc, _ = mo.orientation_row([theta, phi])
    q = qc * [1, -1, -1, -1]
    st, ct, sp, cp = np.sin(theta), np.cos(theta), np.sin(phi), np.cos(phi)
    assert np.allclose(mo.quat_rotmat(q) @ [0, 0, 1], [st * cp, st * sp, ct])
    assert np.allclose(mo.quat_rotmat(qc) @ [0, 0, 1], [-st * cp, st * sp, ct])
