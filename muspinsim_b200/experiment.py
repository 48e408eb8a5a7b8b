"""ExperimentRunner: the reference's driver class with the per-configuration Python loop
replaced by batched calls into the CUDA library.

Reference: /root/reference/muspinsim/experiment.py:22-498.  `run()` keeps the reference's
contract: it returns (and stores in `config.results`) a float64 array of shape
[len(file_range_1), ..., len(x_range)], already divided by avg_N, summed over ranks.

Construction takes a "spec" (dict mirroring the `.in` keywords, see configs.default_spec) or a
prebuilt (system, ConfigTable) pair; `adapter.runner_from_reference` builds one from the
reference's own ExperimentRunner when that package is importable.
"""

import numpy as np

from . import _lib
from .configs import ConfigTable
from .constants import MU_TAU
from .spinsys import MuonSpinSystem, system_from_spec


class ExperimentRunner:
    def __init__(self, spec=None, system=None, table=None, dissipation=None, device=None, comm=None,
                 handle_cache=None):
        """
        spec        dict of `.in` keywords (spins, couplings, field, ..., see configs.default_spec)
        system      alternatively a prebuilt MuonSpinSystem (+ `table`, `dissipation`)
        device      CUDA device index (default: LOCAL_RANK or 0)
        comm        a `dist.Communicator` (shards configurations over ranks, one reduce at the
                    end -- the role of mpi.py in the reference); None = single process
        handle_cache  dict shared between runners: a runner whose spin system has the same spins
                    (dimensions, gyromagnetic ratios, muon, dissipation) as an earlier one reuses
                    that runner's device handle -- workspaces stay allocated, only H0 / Z are
                    re-uploaded (musim_update_system).  This is the resident-system fitting loop:
                    FittingRunner builds a new ExperimentRunner per function evaluation
                    (fitting.py:126-135)
        """
        if spec is not None:
            system, dissipation = system_from_spec({**{"spins": ["mu", "e"], "couplings": []}, **spec})
            table = ConfigTable(spec)
        if not isinstance(system, MuonSpinSystem) or table is None:
            raise TypeError("ExperimentRunner needs a spec or a (MuonSpinSystem, ConfigTable) pair")
        self._system = system
        self._table = table
        self._dissip = dict(dissipation or {})
        self._comm = comm
        if device is None:
            import os

            device = int(os.environ.get("LOCAL_RANK", "0"))
        self._device = device
        self._handle = None
        self._handle_cache = handle_cache
        self.results = None
        self._results_function = (spec or {}).get("results_function") if spec is not None else None
        # `celio k [averages]` keyword (simconfig.py:537-565): k = 0 means Celio's method is not used
        cel = list(np.atleast_1d((spec or {}).get("celio", [0])))
        self._celio_k = int(cel[0])
        self._celio_averages = int(cel[1]) if len(cel) > 1 else 0
        if self._celio_k < 0 or self._celio_averages < 0:
            raise ValueError("Value of k / averages for Celio's method must a postive integer or 0")
        self.options = {}
        self.device_expand = True  # expand the configuration table on the device when all configurations share a mode

    @property
    def system(self):
        return self._system

    @property
    def config(self):
        return self._table

    @property
    def handle(self):
        if self._handle is None and self._handle_cache is not None:
            key = (self._device, tuple(int(x) for x in self._system.dimension),
                   tuple(float(g) for g in self._system.gammas), int(self._system.muon_index),
                   tuple(sorted((int(k), float(v)) for k, v in self._dissip.items())))
            h = self._handle_cache.get(key)
            if h is not None:
                self._handle = h
                h.set_option("defaults", 1)  # an earlier runner's kernel options do not leak into this one
                for k, v in self.options.items():
                    h.set_option(k, v)
                return h
        if self._handle is None:
            ds = list(self._dissip.keys())
            dr = [self._dissip[k] for k in ds]
            self._handle = _lib.Handle(
                self._device,
                self._system.dimension,
                self._system.gammas,
                self._system.muon_index,
                self._system.hamiltonian,
                self._system.zeeman_operators(),
                self._system.muon_operators(),
                ds,
                dr,
            )
            for k, v in self.options.items():
                self._handle.set_option(k, v)
            if self._handle_cache is not None:
                if len(self._handle_cache) >= 4:  # a fit varies couplings, not the spin list
                    self._handle_cache.pop(next(iter(self._handle_cache)))
                self._handle_cache[key] = self._handle
        return self._handle

    def set_option(self, key, value):
        self.options[key] = value
        if self._handle is not None:
            self._handle.set_option(key, value)

    # --------------------------------------------------------------------------------------
    def _modes(self, sel):
        """Split the selected configurations by the reference function they would take
        (experiment.py:449-496)."""
        tab = self._table
        lind = len(self._dissip) > 0
        if tab.y == "asymmetry":
            if lind:
                return [(_lib.MODE_LINDBLAD, sel)]
            fast = tab.fast[sel]
            out = []
            if fast.any():
                out.append((_lib.MODE_FAST, sel[fast]))
            if (~fast).any():
                out.append((_lib.MODE_EVOLVE, sel[~fast]))
            return out
        if lind:
            return [(_lib.MODE_LINDBLAD_INT, sel)]
        fast = tab.fast[sel]
        out = []
        if fast.any():
            out.append((_lib.MODE_INTEGRAL_FAST, sel[fast]))
        if (~fast).any():
            out.append((_lib.MODE_INTEGRAL, sel[~fast]))
        return out

    def _uniform_mode(self):
        """The one mode all configurations take, or None if they differ (experiment.py:449-496)."""
        tab = self._table
        if len(self._dissip) > 0:
            return _lib.MODE_LINDBLAD if tab.y == "asymmetry" else _lib.MODE_LINDBLAD_INT
        fast = tab.uniform_fast()
        if fast is None:
            return None
        if tab.y == "asymmetry":
            return _lib.MODE_FAST if fast else _lib.MODE_EVOLVE
        return _lib.MODE_INTEGRAL_FAST if fast else _lib.MODE_INTEGRAL

    def run_partial(self, rank=0, size=1):
        """Accumulate this rank's share of the configurations (the reference's
        `self._config[mpi.rank :: mpi.size]`, experiment.py:369) and return the local
        [n_slots, nt] buffer (host)."""
        tab = self._table
        if tab.y == "asymmetry" and tab.x_name != "t" and not tab.time_isavg and tab._t_pos is None:
            raise ValueError("times must be an array of values in microseconds")
        nt = 1 if tab.y == "integral" else len(tab.times)
        out = np.zeros((tab.n_slots, nt))
        if self._celio_k:
            return self._run_partial_celio(rank, size, out)
        if self._handle_cache is not None and hasattr(self.handle, "update_system"):
            # a shared (cached) handle may hold another runner's couplings: re-upload H0 / Z
            self.handle.update_system(self._system.hamiltonian, self._system.zeeman_operators())
        if self.device_expand and hasattr(self.handle, "run_axes_host"):
            # every configuration takes the same reference function: expand the configuration
            # table on the device from the axis tables (no n_cfg-long host arrays at all)
            mode = self._uniform_mode()
            if mode is not None:
                n_loc = len(range(rank, tab.n_cfg, size))
                if n_loc:
                    self.handle.run_axes_host(mode, n_loc, rank, size, tab.axes_descriptor(),
                                              None if tab.y == "integral" else tab.times, MU_TAU, out)
                return out
        sel = np.arange(tab.n_cfg)[rank::size]
        if len(sel) == 0:
            return out
        for mode, idx in self._modes(sel):
            order = idx[np.argsort(tab.slot[idx], kind="stable")]  # group slots -> fewer flushes
            self.handle.run_host(
                mode,
                tab.B[order],
                tab.p[order],
                tab.T[order],
                tab.w[order],
                tab.slot[order],
                None if tab.y == "integral" else tab.times,
                MU_TAU,
                out,
            )
        return out

    def _run_partial_celio(self, rank, size, out):
        """Celio's method (experiment.py:454-470): every configuration is one call of the batched
        GPU state-vector evolution, all random initial states at once.  Hsys + Hz as lists of terms
        (spinsys.py:613-626, experiment.py:251-267): the system's interaction terms, then one Zeeman
        term per spin when the field is not zero."""
        from .celio import CelioHamiltonian, term_matrix, terms_from_system

        tab = self._table
        if self._dissip:
            raise NotImplementedError("Dissipation is not supported when using Celio's method")
        if tab.y != "asymmetry":
            raise NotImplementedError("Celio's method evaluates the time-domain signal only")
        if self._celio_averages <= 0:
            raise NotImplementedError("Celio's method without random initial states (density-matrix Trotter "
                                      "evolution, celio.py:207-287) stays with the reference")
        base = terms_from_system(self._system)
        for c in range(rank, tab.n_cfg, size):
            if tab.T[c] != np.inf:
                raise ValueError("The fast version of Celio's method requires T -> inf. Either remove the number "
                                 "of averages or don't use celio.")
            terms = list(base)
            B = tab.B[c]
            if not np.array_equal(B, [0, 0, 0]):
                for i in range(len(self._system)):
                    terms.append(((i,), term_matrix(self._system, (i,), B * self._system.gammas[i])))
            H = CelioHamiltonian(terms, self._celio_k, self._system, device=self._device)
            data = H.fast_evolve(self._system.sigma_mu(tab.p[c]), tab.times, self._celio_averages)
            out[tab.slot[c], :] += tab.w[c] * data
        return out

    def run(self):
        """Run every configuration, reduce over ranks, return results in the reference layout."""
        comm = self._comm
        rank, size = (comm.rank, comm.size) if comm is not None else (0, 1)
        out = self.run_partial(rank, size)
        if comm is not None:
            out = comm.sum_data(out)  # mpi.py:104-112
        results = self._table.finish(out)
        if self._results_function is not None:
            results = self.apply_results_function(results)
        self.results = results
        return self.results

    # ---- output side (experiment.py:347-356, simconfig.py:370-429) -----------------------
    def apply_results_function(self, results, variables=None):
        """The reference evaluates the `results_function` expression of the input file with the
        variables x (x-axis values) and y (results); here it is a callable f(x, y, **variables) or an
        expression string evaluated with numpy's functions in scope."""
        f = self._results_function
        x = self._table.x_axis_values
        if callable(f):
            y = f(x, results, **(variables or {}))
        else:
            scope = {k: getattr(np, k) for k in ("exp", "sin", "cos", "tan", "sqrt", "log", "abs", "pi", "arcsin",
                                                 "arccos", "arctan", "sinh", "cosh", "tanh")}
            scope.update(variables or {})
            scope.update(x=x, y=results)
            y = eval(compile(str(f), "<results_function>", "eval"), {"__builtins__": {}}, scope)  # noqa: S307
        return np.array(y, dtype=float).reshape(results.shape)

    def save_output(self, name=None, path=".", extension=".dat"):
        """Write the `.dat` files of the last run (MuSpinConfig.save_output, simconfig.py:370-429)."""
        if self.results is None:
            raise RuntimeError("run() has not been called")
        return self._table.save_output(self.results, name=name, path=path, extension=extension)
