"""Drive the UNMODIFIED reference (installed in oracle/_ref by oracle/build_ref.py, with the
third-party stand-ins of oracle/shims on the path) through its public API:
`.in` text -> MuSpinInput -> ExperimentRunner.run().  TEST INFRASTRUCTURE ONLY.

Used here (build container) to pin the numpy oracle and to generate tests/golden/*.npz, and
by `bench.py --impl reference` when oracle/_ref travelled to the GPU box.
"""
import io
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def available():
    return os.path.isdir(os.path.join(_HERE, "_ref", "muspinsim"))


def _import():
    for p in (os.path.join(_HERE, "_ref"), os.path.join(_HERE, "shims")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import logging
    import warnings

    warnings.filterwarnings("ignore")
    logging.disable(logging.WARNING)
    import muspinsim  # noqa: F401

    return muspinsim


def make_runner(spec):
    ms = _import()
    from .muspin_oracle import spec_to_infile

    infile = ms.MuSpinInput(io.StringIO(spec_to_infile(spec)))
    return ms.ExperimentRunner(infile, {})


def run_reference(spec, rank=0, size=1):
    """Return the reference's results array for `spec`; with size > 1 only the slice
    [rank::size] of the configuration list is evaluated (experiment.py:369), unsummed."""
    runner = make_runner(spec)
    if size == 1:
        return runner.run()
    cfg = runner.config
    for snap in cfg[rank::size]:
        cfg.store_time_slice(snap.id, runner.run_single(snap))
    return cfg.results
