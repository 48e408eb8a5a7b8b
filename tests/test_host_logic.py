"""Host-side logic without a GPU: spin system builder, vectorised configuration table,
mode selection and result layout of ExperimentRunner, checked against the oracle / golden
vectors with the CPU stand-in handle of tests/helpers.py."""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from helpers import OracleHandle
from muspinsim_b200 import ExperimentRunner, MuonSpinSystem, _lib, configs
from muspinsim_b200.spinsys import spin_operators, system_from_spec
from oracle import muspin_oracle as mo


@pytest.mark.parametrize("name", golden_names())
def test_runner_host_logic_matches_golden(name):
    spec, want = load_golden(name)
    r = ExperimentRunner(spec)
    r._handle = OracleHandle(spec)
    got = r.run()
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) < 1e-10


@pytest.mark.parametrize("name", ["c2_fast_d16", "c5_fast_d96", "c3_alc_d24", "ground_state_T0", "c4_dissip_tf"])
def test_system_operators_match_oracle(name):
    spec, _ = load_golden(name)
    a = mo.build_system(spec)
    b, dis = system_from_spec(spec)
    assert np.allclose(a.H0, b.hamiltonian, atol=1e-12, rtol=0)
    v = np.array([0.3, -0.2, 0.9])
    assert np.allclose(a.muon_operator(v), b.muon_operator(v), atol=1e-15)
    assert np.allclose(mo.zeeman_matrix(a, v), np.tensordot(v, b.zeeman_operators(), 1), atol=1e-9, rtol=1e-15)
    assert np.allclose(a.sigma_mu(v), b.sigma_mu(v))
    assert dict(a.dissipation) == dis


def test_spin_operator_algebra():
    for I in (0.5, 1.0, 1.5, 3.5):
        sx, sy, sz = spin_operators(I)
        assert np.allclose(sx @ sy - sy @ sx, 1j * sz)
        assert np.allclose(sx @ sx + sy @ sy + sz @ sz, I * (I + 1) * np.eye(int(2 * I + 1)))
        ox, oy, oz = mo.spin_matrices(I)
        assert np.allclose(sx, ox) and np.allclose(sy, oy) and np.allclose(sz, oz)


def test_spinsys_errors_mirror_reference():
    with pytest.raises(ValueError):
        MuonSpinSystem(["e", "e"])  # exactly one muon (spinsys.py:646-649)
    s = MuonSpinSystem(["mu", "e", "e"])
    with pytest.raises(ValueError):
        s.add_hyperfine_term(0, np.eye(3))  # must name the electron (spinsys.py:684-689)
    with pytest.raises(ValueError):
        s.add_hyperfine_term(1, np.eye(3), 2)  # first index must not be an electron
    with pytest.raises(ValueError):
        s.add_dipolar_term(0, 0, [0, 0, 1])
    with pytest.raises(ValueError):
        s.add_quadrupolar_term(0, np.eye(3))  # spin 1/2
    with pytest.raises(ValueError):
        s.muon_operator([1, 0])
    with pytest.raises(ValueError):
        s.add_linear_term(5, [0, 0, 1])


def test_quaternion_conventions():
    # tests/test_config.py:231-233 and tests/test_utils.py:43-65 of the reference
    q, w = configs.orientation_table([[0.0, 0, 0, 1.0], [0.5 * np.pi, 0, 0, 2.0]])
    assert np.allclose(q[1], [2**-0.5, 0, 0, -(2**-0.5)]) and np.allclose(w, [1, 2])
    theta, phi = 0.6 * np.pi, 0.4 * np.pi
    qc, _ = configs.orientation_table([[theta, phi]])
    st, ct, sp, cp = np.sin(theta), np.cos(theta), np.sin(phi), np.cos(phi)
    R = configs.quat_rotation_matrices(qc)[0]
    assert np.allclose(R @ [0, 0, 1], [-st * cp, st * sp, ct])
    rng = np.random.default_rng(0)
    rows = rng.uniform(0, np.pi, size=(20, 3))
    for mode in ("zyz", "zxz"):
        qq, _ = configs.orientation_table(rows, mode)
        for k in range(20):
            q1, _ = mo.orientation_row(rows[k], mode)
            assert np.allclose(qq[k], q1, atol=1e-15)
    assert np.allclose(configs.eulrange(4), mo.eulrange(4))


def test_config_table_matches_oracle_enumeration():
    for name in ("alc_T_filerange", "intrinsic_scan_zxz", "time_averaged_vs_field", "polarization_filerange"):
        spec, want = load_golden(name)
        tab = configs.ConfigTable(spec)
        oc = mo.OracleConfig(spec)
        assert tab.n_cfg == len(oc.configurations)
        assert tab.results_shape == oc.results.shape
        assert tab.avg_N == oc.avg_N
        for idx in range(tab.n_cfg):
            snap = oc.snapshot(idx)
            q, w = snap["orient"]
            R = mo.quat_rotmat(q)
            B = R @ np.asarray(snap["B"]) + np.asarray(snap["intrinsic_B"])
            assert np.allclose(tab.B[idx], B, atol=1e-15)
            assert np.allclose(tab.p[idx], R @ np.asarray(snap["mupol"]), atol=1e-15)
            assert tab.T[idx] == snap["T"]
            assert np.isclose(tab.w[idx], w / oc.avg_N)


def test_config_errors():
    with pytest.raises(ValueError):
        configs.ConfigTable({"y_axis": "integral"})  # time as x axis with integral (simconfig.py:173-178)
    with pytest.raises(ValueError):
        configs.ConfigTable({"x_axis": "field"})  # x axis is not a range (simconfig.py:236-238)
    with pytest.raises(ValueError):
        configs.ConfigTable({"orientation": [[1.0]]})
    with pytest.raises(ValueError):
        configs.ConfigTable({"field": [[1.0, 2.0]]})


def test_weights_normalised_only_when_averaged():
    spec = {"orientation": [[0, 0, 0, 1.0], [1, 1, 1, 3.0]]}
    t = configs.ConfigTable(spec)
    assert np.allclose(t.w * t.avg_N, [0.5, 1.5])  # sum to N (simconfig.py:152-156)
    t = configs.ConfigTable(dict(spec, average_axes=["none"]))
    assert np.allclose(t.w, 1.0) and t.results_shape == (2, 101)


def test_fast_predicate():
    t = configs.ConfigTable({"temperature": [np.inf, 1.0], "field": [[0, 0, 0.1]], "average_axes": ["none"]})
    assert list(t.fast) == [True, False]
    t = configs.ConfigTable({"temperature": [1.0], "field": [[0, 0, 0.0]]})
    assert list(t.fast) == [True]  # B = 0 (experiment.py:413-418)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.MusimError):
        _lib.load()


@pytest.mark.parametrize("name", ["c2_fast_d16", "alc_T_filerange", "time_averaged_vs_field", "c4_dissip_tf",
                                  "polarization_filerange", "hfine_powder_eulrange3"])
def test_adapter_from_reference_runner(name):
    """The drop-in adapter on the reference's own parsed objects (needs oracle/_ref: build
    container only; skipped where the reference install did not travel)."""
    from oracle import ref_driver

    if not ref_driver.available():
        pytest.skip("oracle/_ref not present")
    from muspinsim_b200 import adapter

    spec, want = load_golden(name)
    ref_runner = ref_driver.make_runner(spec)
    r = adapter.runner_from_reference(ref_runner)
    r._handle = OracleHandle(spec)
    got = r.run()
    assert got.shape == want.shape and np.max(np.abs(got - want)) < 1e-10
    ours = system_from_spec(spec)[0]
    assert np.allclose(r.system.hamiltonian, ours.hamiltonian, atol=1e-12)
    assert np.allclose(r.system.zeeman_operators(), ours.zeeman_operators(), atol=1e-9, rtol=1e-15)


def test_axes_descriptor_matches_materialised_table():
    """Device-side expansion (musim_run_axes_host): the axis descriptor enumerates exactly the
    configurations of the materialised table (as a multiset of (B, p, T, w, slot) rows), with
    non-decreasing output rows, and the strided shards of two ranks partition it."""
    from muspinsim_b200.configs import ConfigTable

    rng = np.random.default_rng(0)
    spec = {
        "spins": ["mu", "e"],
        "polarization": [[1, 0, 0], [0, 1, 1]],
        "field": [[0.0, 0.0, b] for b in (0.0, 0.1, 0.25)],
        "intrinsic_field": [[0.0, 0.0, 0.0], [0.01, 0.0, 0.0]],
        "orientation": np.column_stack([rng.uniform(0, 6, 5), rng.uniform(0, 3, 5), rng.uniform(0, 6, 5), rng.uniform(0.5, 2, 5)]),
        "temperature": [np.inf, 1.0],
        "time": np.linspace(0, 1, 7),
        "x_axis": "field",
        "average_axes": ["orientation", "intrinsic_field"],
    }
    tab = ConfigTable(spec)
    B, p, T, w, slot = tab.expand_axes()
    assert len(B) == tab.n_cfg and np.all(np.diff(slot) >= 0)
    key = lambda B, p, T, w, s: sorted(map(tuple, np.round(np.column_stack([B, p, np.where(np.isinf(T), -1, T), w, s]), 12)))
    assert key(B, p, T, w, slot) == key(tab.B, tab.p, tab.T, tab.w, tab.slot)
    parts = [tab.expand_axes(r, 2) for r in range(2)]
    assert sum(len(x[0]) for x in parts) == tab.n_cfg
    assert key(*[np.concatenate([a[i] for a in parts]) for i in range(5)]) == key(B, p, T, w, slot)
    assert tab.uniform_fast() is None  # mixed temperatures / fields
    assert ConfigTable(dict(spec, temperature=[np.inf])).uniform_fast() is True
    assert ConfigTable(dict(spec, temperature=[2.0], field=[[0, 0, 0.1], [0, 0, 0.2]],
                            intrinsic_field=[[0, 0, 0]], average_axes=["orientation"])).uniform_fast() is False
    assert ConfigTable(dict(spec, temperature=[2.0], field=[[0, 0, 0.0], [0, 0, 0.2]],
                            intrinsic_field=[[0, 0, 0]], average_axes=["orientation"])).uniform_fast() is None


def test_zcw_generator_properties():
    """The reference pins only the row count and the vanishing P2 average of zcw(N)
    (tests/test_input.py:271-273, tests/test_utils.py:67-73)."""
    from muspinsim_b200.configs import ConfigTable, zcw

    for n in (1, 20, 100, 1000):
        rows = zcw(n)
        assert rows.shape[1] == 2 and len(rows) >= n
        assert abs(np.mean(3 * np.cos(rows[:, 0]) ** 2 - 1)) < 1e-3 if n >= 100 else True
    tab = ConfigTable({"spins": ["mu", "e"], "orientation": zcw(50), "time": np.linspace(0, 1, 5)})
    assert tab.n_cfg == len(zcw(50)) and abs(tab.w.sum() - 1.0) < 1e-12


def test_nufft_tables_reproduce_the_transform_on_the_host():
    """The library's NUFFT tables (per-tap kernel polynomials, deconvolution factors), used in a
    numpy restatement of the spread + FFT steps of polar_nufft.cuh, reproduce the direct sum
    sum_p c_p exp(i k theta_p) to 1e-11 * sum|c| -- no GPU involved."""
    from muspinsim_b200 import _lib

    rng = np.random.default_rng(3)
    for nt in (100, 1000):
        M, w, deg, coef, dec = _lib.nufft_tables(nt)
        assert M >= 2 * nt and M & (M - 1) == 0 and w == 12
        npt = 3000
        f = rng.normal(0, 200.0, npt)
        f[:300] = np.round(f[:300])  # exact grid hits and clusters
        c = rng.normal(size=npt) + 1j * rng.normal(size=npt)
        c /= np.abs(c).sum()
        dt = 0.01
        u = np.rint(f * dt) - f * dt  # cycles per step in [-1/2, 1/2]
        k = np.arange(nt)
        want = (c[None, :] * np.exp(2j * np.pi * np.outer(k, u))).sum(1)
        # half grid (polar_nufft.cuh): points mirrored to u in [0, 1/2] with conjugate strength, private
        # cells -marg .. M/2 + w/2, margins folded (conjugated) onto the cells they mirror
        cp = c * np.exp(2j * np.pi * (nt // 2) * u)
        cp = np.where(u < 0, np.conj(cp), cp)
        pos = np.abs(u) * M
        fl = np.floor(pos)
        y = 2.0 * (pos - fl) - 1.0
        marg = w // 2 - 1
        base = fl.astype(np.int64)  # private cell of tap 0 (= cell fl - marg, shifted by marg)
        priv = np.zeros(M // 2 + marg + 16, complex)
        for l in range(w):
            phi = np.polynomial.polynomial.polyval(y, coef[l], tensor=False)
            np.add.at(priv, base + l, cp * phi)
        grid = np.zeros(M, complex)
        for q, v in enumerate(priv):
            m = q - marg
            if m < 0:
                grid[-m] += np.conj(v)
            elif m > M // 2:
                grid[M - m] += np.conj(v)
            else:
                grid[m] += v
        F = np.fft.ifft(grid) * M
        got = (F[(k - nt // 2) % M] * dec).real
        want = want.real
        assert np.max(np.abs(got - want)) < 1e-11


def test_time_as_x_axis_and_averaged_axis_matches_reference():
    """x_axis = time with time ALSO in average_axes: the reference keeps t as the x range, sets
    _time_isavg and stores the time average in every x column (simconfig.py:172-191, 364-365)."""
    from oracle import ref_driver

    spec = {"name": "t_avg_x", "spins": ["mu", "e"],
            "couplings": [{"type": "hyperfine", "i": 1, "j": None, "value": [[5.0, 0, 0], [0, 5.0, 0], [0, 0, 5.0]]}],
            "time": list(np.linspace(0.0, 1.0, 7)), "x_axis": "time", "average_axes": ["orientation", "time"],
            "orientation": [[0.0, 0.0, 0.0, 1.0], [0.3, 0.4, 0.5, 1.0]]}
    r = ExperimentRunner(spec)
    r._handle = OracleHandle(spec)
    got = r.run()
    assert got.shape == (7,) and np.allclose(got, got[0])
    if ref_driver.available():
        want = ref_driver.run_reference(spec)
        assert want.shape == got.shape and np.max(np.abs(got - want)) < 1e-12


def test_run_host_validates_array_lengths_and_slots():
    """_lib.Handle.run_host rejects ragged configuration arrays before the C call (no GPU needed:
    the check runs on a handle-less instance)."""
    h = _lib.Handle.__new__(_lib.Handle)
    out = np.zeros((2, 3))
    t = np.linspace(0, 1, 3)
    B = np.zeros((4, 3))
    with pytest.raises(ValueError):
        _lib.Handle.run_host(h, 1, B, B, np.zeros(4), np.ones(3), np.zeros(4, int), t, 1.0, out)  # len(w) != n
    with pytest.raises(ValueError):
        _lib.Handle.run_host(h, 1, B, B, np.zeros(3), np.ones(4), np.zeros(4, int), t, 1.0, out)  # len(T) != n
    with pytest.raises(ValueError):
        _lib.Handle.run_host(h, 1, B, B, np.zeros(4), np.ones(4), np.array([0, 1, 2, 0]), t, 1.0, out)  # slot >= n_slots


@pytest.mark.parametrize("name", ["alc_T_filerange", "polarization_filerange", "time_averaged_vs_field", "c2_fast_d16"])
def test_output_side_matches_the_reference_dat_files(name, tmp_path):
    """ExperimentRunner.save_output writes the reference's files -- same names, same
    two columns (|B| as x for field axes), same header apart from the date line (simconfig.py:370-429)."""
    import os

    from oracle import ref_driver

    spec, want = load_golden(name)
    r = ExperimentRunner(spec)
    r._handle = OracleHandle(spec)
    r.run()
    ours = tmp_path / "ours"
    ours.mkdir()
    files = r.save_output(name="out", path=str(ours))
    assert len(files) == int(np.prod(r.config.results_shape[:-1])) and all(os.path.exists(f) for f in files)
    for f in files:
        data = np.loadtxt(f)
        assert data.shape == (r.config.results_shape[-1], 2)
    if not ref_driver.available():
        return
    theirs = tmp_path / "theirs"
    theirs.mkdir()
    rr = ref_driver.make_runner(spec)
    rr.run()
    rr.config.save_output(name="out", path=str(theirs))
    assert sorted(os.listdir(ours)) == sorted(os.listdir(theirs))
    for f in sorted(os.listdir(ours)):
        a, b = np.loadtxt(ours / f), np.loadtxt(theirs / f)
        assert a.shape == b.shape and np.max(np.abs(a - b)) < 1e-10
        ha = [ln for ln in open(ours / f) if ln.startswith("#")]
        hb = [ln for ln in open(theirs / f) if ln.startswith("#")]
        assert len(ha) == len(hb)
        for la, lb in zip(ha, hb):
            if "written on" not in la:
                assert la == lb


def test_results_function_callable_and_expression():
    """experiment.py:347-356: the results function sees x (x-axis values) and y (results)."""
    spec, want = load_golden("c1_hfine")
    for f in (lambda x, y: 2.0 * y * np.exp(-0.1 * x), "2*y*exp(-0.1*x)"):
        r = ExperimentRunner(dict(spec, results_function=f))
        r._handle = OracleHandle(spec)
        got = r.run()
        t = np.asarray(r.config.x_axis_values)
        assert np.max(np.abs(got - 2.0 * want * np.exp(-0.1 * t))) < 1e-10
