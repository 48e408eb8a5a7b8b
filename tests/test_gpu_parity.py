"""Parity of the CUDA path (through the C ABI) with the CPU oracle and with the golden vectors
generated from the unmodified reference.  Tolerance: 1e-9 absolute on the polarisation
(BASELINE.json north_star), FP64."""
import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _run(spec, **opts):
    from muspinsim_b200 import ExperimentRunner

    r = ExperimentRunner(spec, device=0)
    for k, v in opts.items():
        r.set_option(k, v)
    return r.run(), r


@pytest.mark.parametrize("name", golden_names())
def test_golden(name):
    spec, want = load_golden(name)
    got, _ = _run(spec)
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) < TOL


@pytest.mark.parametrize("opts", [{"polar": 1}, {"polar": 3}, {"eigh": 1}, {"eigh": 2, "polar": 2}, {"chunk": 3},
                                  {"polar_mma": 0}, {"rho0_dense": 1}, {"int_fused": 0}, {"zgemm_pipe": 0}, {"sorted": 1}, {"back_wy": 0}, {"tdc": 0}, {"tdc": 0, "back_wy": 0}, {"reflect": 0}, {"tridiag_rw": 0}, {"gemm": 1}, {"tridiag_warp": 0}, {"small24": 0}, {"tridiag_fused": 0}, {"tridiag_phases": 0}, {"tridiag_hs": 0}, {"tridiag_hsw": 1}, {"back_wy_small": 0}, {"apply_warp": 0}, {"tql_threads": 32}, {"tql_threads": 16}, {"tql_threads": 8}])
@pytest.mark.parametrize("name", ["c2_fast_d16", "c2_general_d8_T0p3", "c3_alc_d12", "c5_fast_d96",
                                  "polarization_filerange", "ground_state_T0"])
def test_golden_all_kernel_variants(name, opts):
    spec, want = load_golden(name)
    got, _ = _run(spec, **opts)
    assert np.max(np.abs(got - want)) < TOL


def test_nonuniform_times_take_direct_path_and_match():
    spec, want = load_golden("nonuniform_times")
    got, _ = _run(spec)
    assert np.max(np.abs(got - want)) < TOL
    with pytest.raises(ValueError):
        _run(spec, polar=2)  # the factorised kernel refuses non-uniform grids
    with pytest.raises(ValueError):
        _run(spec, polar=3)  # and so does the NUFFT kernel


@pytest.mark.parametrize("nt,t0", [(2, 0.0), (50, 0.0), (100, 0.3), (1000, 0.0), (1024, 1.7), (1025, 0.0), (4096, 0.0),
                                   (5000, 0.0)])
@pytest.mark.parametrize("temperature", [np.inf, 0.4])
def test_nufft_polarisation_matches_direct_kernel(nt, t0, temperature):
    """Type-1 NUFFT polarisation (polar=3; the default for uniform grids of >= 96 points) against
    the direct sincos kernel: fast and general paths, t0 != 0, grid sizes 64 .. 8192 (nt = 5000
    falls back to the time-factorised kernel)."""
    from muspinsim_b200 import workloads

    spec = workloads.c2_hfine_powder(n_orient=37, nt=nt, n_h=2, temperature=temperature)
    spec["time"] = t0 + np.linspace(0.0, 10.0, nt)
    want, _ = _run(spec, polar=1)
    got, _ = _run(spec, polar=3)
    assert np.max(np.abs(got - want)) < 1e-11
    auto, _ = _run(spec)
    assert np.max(np.abs(auto - want)) < 1e-11


def test_nufft_many_slots_degenerate_spectrum_and_oracle():
    """Field scan with time as a file range (one fine grid per output row), zero-field row
    (degenerate eigenvalues: every pair of a multiplet lands on the same cells), few
    configurations (pair lists split over warps); checked against the CPU oracle."""
    from muspinsim_b200 import workloads
    from oracle import muspin_oracle as mo

    spec = workloads.c2_hfine_powder(n_orient=5, nt=128, n_h=1)
    spec["field"] = [[0.0], [0.005], [0.02]]
    spec["x_axis"] = "field"
    want = mo.run_spec(spec, evolve_fn=mo.evolve_vectorised)
    got, _ = _run(spec, polar=3)
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) < TOL


@pytest.mark.parametrize("seed", range(4))
def test_seeded_against_oracle(seed):
    """Fresh seeded inputs (not in the golden set) against the oracle."""
    from muspinsim_b200 import workloads
    from oracle import muspin_oracle as mo

    rng = np.random.default_rng(seed)
    nt = int(rng.integers(1, 300))
    spec = workloads.c2_hfine_powder(n_orient=int(rng.integers(1, 40)), nt=nt, n_h=int(rng.integers(0, 3)),
                                     temperature=[np.inf, 0.5, 20.0, 0.0][seed], seed=100 + seed)
    spec["field"] = [[0.001 * seed, 0.0, 0.02]]
    want = mo.run_spec(spec, evolve_fn=mo.evolve_vectorised)
    got, _ = _run(spec)
    assert np.max(np.abs(got - want)) < TOL


def test_edge_cases():
    from oracle import muspin_oracle as mo

    t = np.linspace(0, 10, 100)
    # empty system: constant 0.5 (tests/test_experiment.py:117-133)
    got, _ = _run({"spins": ["e", "mu"], "time": t}, polar=2)
    assert np.max(np.abs(got - 0.5)) < 1e-14
    got, _ = _run({"spins": ["e", "mu"], "time": t})  # NUFFT path: aliasing error ~1e-11 * sum|w_ij|
    assert np.max(np.abs(got - 0.5)) < 1e-10
    # single spin, single time point ... and 1025 time points (two a-blocks in the factorised kernel)
    zee = [{"type": "zeeman", "i": 2, "value": [0, 0, 1.0 / mo.MU_GAMMA]}]
    for n in (2, 33, 1025, 2500):
        tt = np.linspace(0, 10, n)
        got, _ = _run({"spins": ["e", "mu"], "time": tt, "couplings": zee})
        assert np.max(np.abs(got - 0.5 * np.cos(2 * np.pi * tt))) < TOL
    # integral known answer (tests/test_experiment.py:135-202)
    got, _ = _run({"spins": ["e", "mu"], "couplings": zee, "y_axis": "integral", "x_axis": "field",
                   "field": [[0.0], [1.0]]})
    assert abs(got[0] - 0.5 / (1.0 + 4 * np.pi**2 * mo.MU_TAU**2)) < 1e-12


def test_device_pointer_entry_matches_host_entry():
    import torch

    from muspinsim_b200 import ExperimentRunner, _lib
    from muspinsim_b200.constants import MU_TAU

    spec, want = load_golden("c2_fast_d16")
    r = ExperimentRunner(spec, device=0)
    tab = r.config
    dev = torch.device("cuda", 0)
    t = {k: torch.from_numpy(np.ascontiguousarray(getattr(tab, k))).to(dev) for k in ("B", "p", "T", "w")}
    slot = torch.from_numpy(tab.slot).to(dev)
    out = torch.zeros(tab.n_slots, len(tab.times), dtype=torch.float64, device=dev)
    r.handle.run_device(_lib.MODE_FAST, tab.n_cfg, t["B"].data_ptr(), t["p"].data_ptr(), t["T"].data_ptr(),
                        t["w"].data_ptr(), slot.data_ptr(), tab.times, MU_TAU, tab.n_slots, out.data_ptr(),
                        torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.max(np.abs(tab.finish(out.cpu().numpy()) - want)) < TOL
    assert r.handle.launches > 0


def test_launch_groups_match_single_group():
    """The configuration table split into many launch groups (option chunk) gives the same answer
    as one group, checked against the oracle (not against the CUDA path itself)."""
    from muspinsim_b200 import workloads
    from oracle import muspin_oracle

    spec = workloads.c2_hfine_powder(n_orient=300, nt=120, n_h=1)
    want = muspin_oracle.run_spec(spec)
    a, _ = _run(spec, chunk=37)
    assert np.max(np.abs(a - want)) < TOL
    c, _ = _run(dict(spec, temperature=[0.7]), chunk=64)
    assert np.max(np.abs(c - muspin_oracle.run_spec(dict(spec, temperature=[0.7])))) < TOL


def test_accumulate_semantics_and_rank_shards():
    """out is accumulated into (+=); the round-robin shards of two ranks sum to the total."""
    from muspinsim_b200 import ExperimentRunner

    spec, want = load_golden("c2_fast_d16")
    r = ExperimentRunner(spec, device=0)
    a = r.run_partial(0, 2)
    b = r.run_partial(1, 2)
    assert np.max(np.abs(r.config.finish(a + b) - want)) < TOL


@pytest.mark.parametrize("temperature", [np.inf, 0.5])
def test_d128_system_matches_oracle(temperature):
    """mu + e + 5 x 1H (d = 128): beyond the single-SM eigensolver kernels; whole path against
    the oracle (fast and general evolve, and the integral)."""
    from muspinsim_b200 import workloads
    from oracle import muspin_oracle as mo

    spec = workloads.c2_hfine_powder(n_orient=3, nt=120, n_h=5, temperature=temperature)
    want = mo.run_spec(spec, evolve_fn=mo.evolve_vectorised)
    got, _ = _run(spec)
    assert np.max(np.abs(got - want)) < TOL
    spec_i = dict(spec, y_axis="integral", x_axis="field", field=[[0.0, 0.0, 0.01], [0.0, 0.0, 0.3]])
    want = mo.run_spec(spec_i)
    got, _ = _run(spec_i)
    assert np.max(np.abs(got - want)) < TOL


@pytest.mark.parametrize("spins,temperature", [(["mu", "e", "H"], 1.0), (["mu", "e", "14N"], np.inf),
                                               (["mu", "e", "H", "H"], 0.5)])
def test_lindbladian_beyond_shared_memory_matches_oracle(spins, temperature):
    """d = 8 keeps the 64 x 64 super-operator in shared memory; d = 12 (n = 144) and d = 16
    (n = 256) read it in place from global memory.  Evolution and integral against the oracle
    (which diagonalises L like the reference, lindbladian.py:87-99)."""
    from muspinsim_b200 import workloads
    from oracle import muspin_oracle as mo

    rng = np.random.default_rng(len(spins))
    cpl = [{"type": "hyperfine", "i": 1, "value": np.array([[5.0, 2, 3], [2, 5, 2], [3, 2, 5]])}]
    for k in range(3, len(spins) + 1):
        cpl.append({"type": "hyperfine", "i": k, "j": 2, "value": workloads._sym(rng, 3.0)})
    cpl.append({"type": "dissipation", "i": 1, "value": 0.2})
    cpl.append({"type": "dissipation", "i": 3, "value": 0.05})
    spec = {"name": "lind", "spins": spins, "couplings": cpl, "field": [[0.0, 0.0, 0.02]],
            "temperature": [temperature], "time": np.linspace(0.0, 4.0, 130),
            "orientation": workloads._euler_rows(np.random.default_rng(3), 3)}
    want = mo.run_spec(spec)
    got, _ = _run(spec)
    assert np.max(np.abs(got - want)) < TOL
    spec_i = dict(spec, y_axis="integral", x_axis="field", field=[[0.0, 0.0, 0.01], [0.0, 0.0, 0.05]])
    want = mo.run_spec(spec_i)
    got, _ = _run(spec_i)
    assert np.max(np.abs(got - want)) < TOL


def test_lindbladian_nonuniform_times_match_oracle():
    """The reference evaluates the dissipative evolution at arbitrary time rows
    (lindbladian.py:103-108); the GPU path then forms one matrix exponential per time point."""
    from muspinsim_b200 import workloads
    from oracle import muspin_oracle as mo

    spec = workloads.c4_fmuf_dissipation(n_orient=4, nt=10)
    spec["time"] = np.array([0.0, 0.013, 0.2, 0.21, 0.9, 1.7, 3.0, 3.001, 6.5, 8.0])
    want = mo.run_spec(spec)
    got, _ = _run(spec)
    assert np.max(np.abs(got - want)) < TOL


def _per_configuration_cpu(spec):
    """[n_cfg, nt] signal of every configuration from the unmodified reference (oracle/_ref,
    run_single of experiment.py:434-498) when it travelled to this box, else from the oracle port."""
    from oracle import muspin_oracle as mo
    from oracle import ref_driver

    if ref_driver.available():
        runner = ref_driver.make_runner(spec)
        return "reference", np.array([np.atleast_1d(runner.run_single(snap)) for snap in runner.config[:]])
    sys_ = mo.build_system(spec)
    cfg = mo.OracleConfig(spec)
    return "port", np.array([np.atleast_1d(mo.run_single(sys_, cfg, cfg.snapshot(i))) for i in range(len(cfg.configurations))])


def _per_configuration_gpu(spec):
    from muspinsim_b200 import ExperimentRunner
    from muspinsim_b200.constants import MU_TAU

    r = ExperimentRunner(spec, device=0)
    tab = r.config
    nt = 1 if tab.y == "integral" else len(tab.times)
    out = np.zeros((tab.n_cfg, nt))
    for mode, idx in r._modes(np.arange(tab.n_cfg)):
        part = np.zeros((len(idx), nt))
        r.handle.run_host(mode, tab.B[idx], tab.p[idx], tab.T[idx], tab.w[idx] * tab.avg_N, np.arange(len(idx)),
                          None if tab.y == "integral" else tab.times, MU_TAU, part)
        out[idx] = part
    return out


@pytest.mark.parametrize("case", ["c5_fast", "c5_general", "c2_fast", "c2_general", "c3_d24", "c4_lindblad"])
def test_baseline_size_systems_per_configuration_against_the_reference(case):
    """parity at the BASELINE systems and time grids (d = 96 / 32 / 24, nt = 1000,
    Lindbladian d = 8), configuration by configuration -- not only the powder average -- on a
    sample the CPU finishes in seconds."""
    from muspinsim_b200 import workloads

    spec = {
        "c5_fast": lambda: workloads.c5_large(n_orient=48, nt=1000),
        "c5_general": lambda: workloads.c5_large(n_orient=12, nt=1000, temperature=1.0),
        "c2_fast": lambda: workloads.c2_hfine_powder(n_orient=64, nt=1000),
        "c2_general": lambda: workloads.c2_hfine_powder(n_orient=24, nt=1000, temperature=1.0),
        "c3_d24": lambda: workloads.c3_alc(n_orient=6, n_field=16),
        "c4_lindblad": lambda: workloads.c4_fmuf_dissipation(n_orient=24, nt=1000),
    }[case]()
    kind, want = _per_configuration_cpu(spec)
    got = _per_configuration_gpu(spec)
    assert got.shape == want.shape
    err = np.max(np.abs(got - want))
    assert err < TOL, (case, kind, err)
