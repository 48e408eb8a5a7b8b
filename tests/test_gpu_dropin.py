"""The drop-in on hardware: the UNMODIFIED reference (oracle/_ref, travels to the GPU box) is
driven through its own public API -- MuSpinInput -> ExperimentRunner.run() -> save_output(), and
FittingRunner.run() -- once on its CPU path and once with `adapter.patch_reference()` routing
`ExperimentRunner.run` (experiment.py:358-382) to the CUDA library.  The `.dat` files and the
fitted parameters must agree.  Skipped where oracle/_ref is absent."""
import io
import os

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _need_ref():
    from oracle import ref_driver

    if not ref_driver.available():
        pytest.skip("oracle/_ref not present (reference install did not travel)")
    ref_driver._import()  # puts oracle/_ref and the third-party shims on sys.path
    return ref_driver


def _dat_files(path):
    out = {}
    for f in sorted(os.listdir(path)):
        if f.endswith(".dat"):
            out[f] = np.loadtxt(os.path.join(path, f))
    return out


@pytest.mark.parametrize("name", ["hfine_powder_eulrange3", "alc_T_filerange", "c4_dissip_tf", "c2_fast_d16",
                                  "time_averaged_vs_field"])
def test_patched_reference_runner_writes_the_same_dat_files(name, tmp_path):
    ref_driver = _need_ref()
    from muspinsim_b200 import _lib, adapter

    spec, want = load_golden(name)
    cpu_dir, gpu_dir = tmp_path / "cpu", tmp_path / "gpu"
    cpu_dir.mkdir()
    gpu_dir.mkdir()
    # reference, CPU path
    r_cpu = ref_driver.make_runner(spec)
    res_cpu = r_cpu.run()
    r_cpu.config.save_output(name="out", path=str(cpu_dir))
    # reference objects, CUDA path
    created = []
    orig_init = _lib.Handle.__init__

    def counting_init(self, *a, **k):
        created.append(1)
        return orig_init(self, *a, **k)

    _lib.Handle.__init__ = counting_init
    adapter._HANDLES.clear()
    adapter.patch_reference()
    try:
        r_gpu = ref_driver.make_runner(spec)
        res_gpu = r_gpu.run()
        r_gpu.config.save_output(name="out", path=str(gpu_dir))
        h = next(iter(adapter._HANDLES.values()))
        assert h.launches > 0  # the CUDA library did the work
    finally:
        adapter.unpatch_reference()
        _lib.Handle.__init__ = orig_init
    assert len(created) == 1
    assert res_gpu.shape == res_cpu.shape == want.shape
    assert np.max(np.abs(res_gpu - res_cpu)) < TOL
    a, b = _dat_files(cpu_dir), _dat_files(gpu_dir)
    assert list(a) == list(b) and len(a) >= 1
    for f in a:
        assert a[f].shape == b[f].shape
        assert np.max(np.abs(a[f] - b[f])) < TOL, f


def _fit_input(a_true, start, method):
    """mu + e with an isotropic hyperfine coupling A (the variable), zero field; the 'experiment'
    is the analytic zero-field signal of that system, 1/2 [1/2 + 1/2 cos(2 pi A t)]... evaluated by
    the reference itself at A = a_true so that no formula of ours enters the target."""
    ref_driver = _need_ref()
    ms = ref_driver._import()
    t = np.linspace(0.0, 0.12, 40)  # about one period: a single minimum between the bounds

    def text(data_block, a_line, with_fit=True):
        s = "spins\n    mu e\nhyperfine 1\n    {0} 0 0\n    0 {0} 0\n    0 0 {0}\n".format(a_line)
        if with_fit:
            s += "fitting_variables\n    A {0} 5.0 15.0\nfitting_method\n    {1}\nfitting_tolerance\n    1e-10\n".format(start, method)
            s += "fitting_data\n" + data_block + "\n"
        else:
            s += "time\n" + "\n".join("    %.17g" % x for x in t) + "\n"
        return s

    target_runner = ms.ExperimentRunner(ms.MuSpinInput(io.StringIO(text("", "%.17g" % a_true, with_fit=False))), {})
    y = target_runner.run()
    block = "\n".join("    %.17g %.17g" % (a, b) for a, b in zip(t, y))
    return ms, text(block, "A")


@pytest.mark.parametrize("method", ["least-squares", "nelder-mead"])
def test_fitting_runner_through_the_patched_path(method):
    """fitting.py:67-151: FittingRunner builds a new ExperimentRunner per function evaluation; with
    the patch every one of them runs on the SAME device handle (only H0 / Z are re-uploaded), and
    the fit converges to the parameters of the reference's own CPU fit."""
    from muspinsim_b200 import _lib, adapter

    ms, text = _fit_input(10.0, 9.0, method)
    sol_cpu = ms.FittingRunner(ms.MuSpinInput(io.StringIO(text))).run()

    created, runs = [], []
    orig_init, orig_axes, orig_host = _lib.Handle.__init__, _lib.Handle.run_axes_host, _lib.Handle.run_host

    def counting_init(self, *a, **k):
        created.append(1)
        return orig_init(self, *a, **k)

    def counting_axes(self, *a, **k):
        runs.append(1)
        return orig_axes(self, *a, **k)

    def counting_host(self, *a, **k):
        runs.append(1)
        return orig_host(self, *a, **k)

    _lib.Handle.__init__, _lib.Handle.run_axes_host, _lib.Handle.run_host = counting_init, counting_axes, counting_host
    adapter._HANDLES.clear()
    adapter.patch_reference()
    try:
        fr = ms.FittingRunner(ms.MuSpinInput(io.StringIO(text)))
        sol_gpu = fr.run()
    finally:
        adapter.unpatch_reference()
        _lib.Handle.__init__, _lib.Handle.run_axes_host, _lib.Handle.run_host = orig_init, orig_axes, orig_host
    assert len(created) == 1, "one device handle for the whole fit"
    assert len(runs) >= 3, "every function evaluation went through the CUDA library"
    assert abs(sol_cpu.x[0] - 10.0) < 1e-4
    assert abs(sol_gpu.x[0] - sol_cpu.x[0]) < 1e-6
    assert np.max(np.abs(fr._runner.config.results - ms.FittingRunner(ms.MuSpinInput(io.StringIO(text)))._ytarg)) < 1e-6


def test_patched_reference_runner_with_celio_averages():
    """`celio k averages` through the reference's own objects: the patched ExperimentRunner.run evaluates the
    random initial states as one batch on the GPU (experiment.py:454-470 loops over them with the C++
    extension).  Both draw the states from numpy's global generator in the same order, so a seeded run must
    reproduce the reference's result; without `averages` the density-matrix variant stays on the reference path."""
    ref_driver = _need_ref()
    if not hasattr(np, "product"):
        np.product = np.prod  # the reference's celio.py predates numpy 2 (test infrastructure only)
    from muspinsim_b200 import _lib, adapter

    spec = {"name": "celio_dropin", "spins": ["mu", "F", "F"],
            "couplings": [{"type": "dipolar", "i": 1, "j": 2, "value": [0.0, 0.0, 1.17]},
                          {"type": "dipolar", "i": 1, "j": 3, "value": [0.0, 0.0, -1.17]}],
            "field": [[0.0, 0.0, 0.002]], "time": np.linspace(0, 2.0, 21), "celio": [3, 5],
            "orientation": [[0.0, 0.0, 0.0], [0.3, 0.7, 1.1]], "orientation_mode": "zyz"}
    np.random.seed(5)
    want = ref_driver.make_runner(spec).run()
    r_gpu = ref_driver.make_runner(spec)
    adapter.patch_reference()
    try:
        l0 = _lib.load().musim_celio_launch_count()
        np.random.seed(5)
        got = r_gpu.run()
        assert _lib.load().musim_celio_launch_count() > l0, "the Celio run did not reach the CUDA library"
        # density-matrix variant (no averages): untouched reference path, no GPU launches
        spec2 = dict(spec, celio=[3], time=np.linspace(0, 0.5, 4), orientation=[[0.0, 0.0, 0.0]])
        l1 = _lib.load().musim_celio_launch_count()
        r2 = ref_driver.make_runner(spec2)
        try:
            r2.run()
        except Exception:
            pass  # (the reference's own slow path needs qutip; only the routing matters here)
        assert _lib.load().musim_celio_launch_count() == l1
    finally:
        adapter.unpatch_reference()
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) < TOL


def test_command_line_through_the_gpu_path(tmp_path):
    """`python -m muspinsim_b200 input.in -o out` = the reference's own CLI (muspinsim/__main__.py) with the hot
    loop on the GPU: the `.dat` file it writes equals the one `python -m muspinsim` writes from the same input."""
    import subprocess
    import sys

    _need_ref()
    from oracle.muspin_oracle import spec_to_infile

    spec, _ = load_golden("hfine_powder_eulrange3")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(root, "oracle", "_ref"), os.path.join(root, "oracle", "shims"), root,
                                         env.get("PYTHONPATH", "")])
    outs = {}
    for mod in ("muspinsim", "muspinsim_b200"):
        wd = tmp_path / mod
        wd.mkdir()
        (wd / "input.in").write_text(spec_to_infile(spec))
        p = subprocess.run([sys.executable, "-m", mod, str(wd / "input.in"), "-o", str(wd / "out")], env=env, cwd=str(wd),
                           capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        outs[mod] = _dat_files(str(wd / "out"))
    assert outs["muspinsim"] and sorted(outs["muspinsim"]) == sorted(outs["muspinsim_b200"])
    for f, ref in outs["muspinsim"].items():
        assert np.max(np.abs(outs["muspinsim_b200"][f] - ref)) < TOL
