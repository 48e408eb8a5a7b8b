"""Size-independent properties of the path at (near) BASELINE sizes, where the oracle would
take too long: t = 0 value, linearity in the weights, permutation invariance, agreement of the
fast and general formulations, agreement of the two polarisation kernels and of the two
eigensolvers, integral = quadrature of the asymmetry."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _runner(spec, **opts):
    from muspinsim_b200 import ExperimentRunner

    r = ExperimentRunner(spec, device=0)
    for k, v in opts.items():
        r.set_option(k, v)
    return r


def test_c2_full_size_properties():
    from muspinsim_b200 import workloads

    spec = workloads.c2_hfine_powder(n_orient=20000, nt=1000)
    a = _runner(spec).run()
    assert abs(a[0] - 0.5) < 1e-12  # P(0) = tr(rho0 O) = 1/2 for every orientation
    assert np.all(np.abs(a) <= 0.5 + 1e-12)
    b = _runner(spec, polar=1, chunk=4096).run()  # direct sincos kernel, different chunking
    assert np.max(np.abs(a - b)) < 1e-10
    # permutation invariance of the powder sum
    perm = np.random.default_rng(0).permutation(20000)
    spec2 = dict(spec, orientation=np.asarray(spec["orientation"])[perm])
    c = _runner(spec2).run()
    assert np.max(np.abs(a - c)) < 1e-11


def test_c5_properties_and_solver_agreement():
    from muspinsim_b200 import workloads

    spec = workloads.c5_large(n_orient=600, nt=1000)
    a = _runner(spec).run()
    assert abs(a[0] - 0.5) < 1e-12
    b = _runner(spec, eigh=1).run()  # Jacobi vs Householder+QL
    assert np.max(np.abs(a - b)) < 1e-10
    # general formulation at a temperature so high that rho_other = 1/d to 1e-13 must agree
    hot = dict(spec, temperature=[1e9])
    g = _runner(hot).run()
    assert np.max(np.abs(a - g)) < 1e-9
    # linearity: weighted sum of two half-tables equals the whole
    o = np.asarray(spec["orientation"])
    h1 = _runner(dict(spec, orientation=o[:300])).run()
    h2 = _runner(dict(spec, orientation=o[300:])).run()
    assert np.max(np.abs(0.5 * (h1 + h2) - a)) < 1e-12


def test_integral_equals_quadrature_of_asymmetry():
    """(1/tau) int_0^inf P(t) exp(-t/tau) dt by Simpson's rule on a fine grid vs the analytic
    integral operator (hamiltonian.py:119-164)."""
    from muspinsim_b200 import workloads
    from muspinsim_b200.constants import MU_TAU

    base = workloads.c3_alc(n_orient=4, n_field=3, extra_h=False)
    base["field"] = [[0.0, 0.0, b] for b in (0.001, 0.002, 0.003)]  # slow dynamics: quadrature converges
    base["couplings"] = [{"type": "hyperfine", "i": 2, "value": np.diag([3.0, 3.0, 5.0])}]
    ana = _runner(base).run()
    n = 200001
    t = np.linspace(0.0, 40 * MU_TAU, n)
    wq = np.ones(n)
    wq[1:-1:2], wq[2:-1:2] = 4.0, 2.0
    wq *= (t[1] - t[0]) / 3.0
    for k in range(3):
        s = dict(base, field=[base["field"][k]], x_axis="time", y_axis="asymmetry", time=t)
        p = _runner(s).run()
        num = np.sum(wq * p * np.exp(-t / MU_TAU)) / MU_TAU
        assert abs(num - ana[k]) < 1e-7


def test_c3_alc_scan_sum_rule():
    from muspinsim_b200 import workloads

    spec = workloads.c3_alc(n_orient=64, n_field=256)
    a = _runner(spec).run()
    assert a.shape == (256,) and np.all(np.isfinite(a))
    assert np.all(a <= 0.5 + 1e-12) and np.all(a >= -1e-12)  # longitudinal integral in [0, 1/2]
    b = _runner(spec, eigh=1).run()
    assert np.max(np.abs(a - b)) < 1e-9
