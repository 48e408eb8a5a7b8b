"""Drop-in adapter: run the reference's own `ExperimentRunner` objects on the GPU path.

`runner_from_reference(ref_runner)` reads the (already parsed and validated) system and
configuration ranges out of a reference `muspinsim.ExperimentRunner` and returns the
equivalent `muspinsim_b200.ExperimentRunner`; `patch_reference()` replaces
`muspinsim.ExperimentRunner.run` so that the CLI, `FittingRunner` and library users call the GPU
path unchanged.  Nothing here imports the reference: it only duck-types the
objects it is handed.
"""

import numpy as np

from .configs import ConfigTable
from .experiment import ExperimentRunner
from .spinsys import MuonSpinSystem


class _ReferenceSystemView(MuonSpinSystem):
    """MuonSpinSystem whose operators are taken from a reference MuonSpinSystem."""

    def __init__(self, ref_runner):
        sysr = ref_runner.system
        self._spins = list(sysr.spins)
        self._gammas = np.array(sysr.gammas, dtype=float)
        self._Qs = np.array(sysr.Qs, dtype=float)
        self._Is = np.array(sysr.Is, dtype=float)
        self._dim = tuple(int(x) for x in sysr.dimension)
        self._mu_i = int(sysr.muon_index)
        self._e_i = set(sysr.elec_indices)
        # interaction terms (label, indices, tensor): what Celio's method builds its gates from (celio.py:73-140)
        self._terms = [(t.label, tuple(int(i) for i in t.indices), np.array(t.tensor, dtype=float))
                       for t in getattr(sysr, "_terms", [])]
        d = int(np.prod(self._dim))
        if getattr(ref_runner.config, "celio_k", 0):
            self._H = np.zeros((d, d), dtype=complex)  # Hsys is a CelioHamiltonian (no dense matrix); not needed
        else:
            H = ref_runner.Hsys.matrix  # spinsys.py:613-626 via experiment.py:119
            self._H = np.asarray(H.toarray() if hasattr(H, "toarray") else H, dtype=complex)
        self._ref = sysr
        from .spinsys import spin_operators

        self._local = [spin_operators(I) for I in self._Is]


def system_from_reference(ref_runner):
    return _ReferenceSystemView(ref_runner)


def table_from_reference(cfg):
    """MuSpinConfig (simconfig.py:82-321) -> ConfigTable, reusing its classified ranges."""
    def get(name):
        for od in (cfg._file_ranges, cfg._avg_ranges, cfg._x_range):
            if name in od and od[name] is not None:
                return list(od[name])
        return [cfg._constants[name]]

    orient = get("orient")
    vals = {
        "mupol": np.array(get("mupol"), dtype=float).reshape(-1, 3),
        "B": np.array(get("B"), dtype=float).reshape(-1, 3),
        "intrinsic_B": np.array(get("intrinsic_B"), dtype=float).reshape(-1, 3),
        "t": np.array(get("t"), dtype=float).reshape(-1),
        "orient": np.array([np.asarray(q.q, dtype=float) for (q, w) in orient]),
        "T": np.array(get("T"), dtype=float).reshape(-1),
    }
    ow = np.array([w for (q, w) in orient], dtype=float)  # already normalised (simconfig.py:152-159)
    x_name = list(cfg._x_range.keys())[0]
    avg = list(cfg._avg_ranges.keys()) + (["t"] if cfg._time_isavg and "t" not in cfg._avg_ranges else [])
    return ConfigTable.from_values(vals, ow, x_name, avg, cfg._y_axis)


# device handles shared by all runners built from reference objects: FittingRunner creates a new
# reference ExperimentRunner per function evaluation (fitting.py:126-135); with the cache only
# H0 / Z are re-uploaded and the workspaces stay resident
_HANDLES = {}


def runner_from_reference(ref_runner, device=None, comm=None):
    celio_k = int(getattr(ref_runner.config, "celio_k", 0) or 0)
    celio_avg = int(getattr(ref_runner.config, "celio_averages", 0) or 0)
    if celio_k and not celio_avg:
        # the density-matrix Trotter variant (celio.py:207-287) stays on the reference path
        raise NotImplementedError("Celio's method without random initial states stays on the reference path")
    dissip = {int(i): float(a) for i, a in ref_runner.config.dissipation_terms.items()}
    r = ExperimentRunner(system=system_from_reference(ref_runner), table=table_from_reference(ref_runner.config),
                         dissipation=dissip, device=device, comm=comm, handle_cache=None if celio_k else _HANDLES)
    if celio_k:  # `celio k averages` (experiment.py:454-470): batched state-vector evolution on the GPU
        r._celio_k, r._celio_averages = celio_k, celio_avg
    return r


def local_device():
    """CUDA device of this process: its node-local rank under torchrun / Open MPI / MVAPICH / Slurm
    (the reference's own launcher is `mpirun -n N muspinsim.mpi`), modulo the device count."""
    import os

    from . import _lib

    n = max(1, _lib.device_count())
    for key in ("LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID"):
        if key in os.environ:
            return int(os.environ[key]) % n
    return 0


def run_reference_runner(ref_runner, device=None, comm=None, mpi=None):
    """What the patched `ExperimentRunner.run` does: GPU evaluation of this rank's share of the
    configurations, the reference's own reduction (`mpi.sum_data`, mpi.py:104-112) and its own
    post-processing (experiment.py:373-382).  `mpi` is the reference's MPIController (rank, size,
    sum_data); None = single process."""
    runner = runner_from_reference(ref_runner, device if device is not None else local_device(), comm)
    if mpi is not None and getattr(mpi, "size", 1) > 1 and comm is None:
        out = runner.run_partial(mpi.rank, mpi.size)  # cfg[rank::size], experiment.py:369
        out = mpi.sum_data(out)
        results = runner.config.finish(out)
    else:
        results = runner.run()
    if not ref_runner._variables:
        results = ref_runner.apply_results_function(results, {})
    ref_runner._config.results = results
    return results


def patch_reference():
    """Monkey-patch muspinsim.ExperimentRunner.run (experiment.py:358-382) with the GPU path."""
    import muspinsim.experiment as mexp

    original = mexp.ExperimentRunner.run

    def run(self):
        if getattr(self.config, "celio_k", 0) and not getattr(self.config, "celio_averages", 0):
            return original(self)  # density-matrix Trotter variant: reference path
        return run_reference_runner(self, mpi=getattr(mexp, "mpi", None))

    run._musim_original = original
    mexp.ExperimentRunner.run = run
    return original


def unpatch_reference():
    import muspinsim.experiment as mexp

    orig = getattr(mexp.ExperimentRunner.run, "_musim_original", None)
    if orig is not None:
        mexp.ExperimentRunner.run = orig
