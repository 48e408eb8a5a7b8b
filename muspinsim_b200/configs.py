"""Configuration table: array-based restatement of the reference's configuration expansion.

The reference materialises one Python namedtuple per configuration
(`MuSpinConfig.__getitem__`, /root/reference/muspinsim/simconfig.py:497-519) and rotates field
and polarisation one at a time with an `ase` quaternion (`ExperimentRunner.load_config`,
experiment.py:384-432).  Here the whole table is built as structure-of-arrays
(B_rot[n,3], p_rot[n,3], T[n], w[n], slot[n]) with vectorised numpy, which is what the GPU
consumes.  Range classification (x axis / averaged / file ranges), ordering and normalisation
follow simconfig.py:134-171, 300-316 and `store_time_slice` (simconfig.py:347-368).
"""

from collections import OrderedDict

import numpy as np
import scipy.constants as cnst

_KEYS = OrderedDict(
    [
        ("polarization", "mupol"),
        ("field", "B"),
        ("intrinsic_field", "intrinsic_B"),
        ("time", "t"),
        ("orientation", "orient"),
        ("temperature", "T"),
    ]
)  # simconfig.py:23-30


def default_spec():
    """Keyword defaults of the .in format (input/keyword.py:357-484)."""
    return {
        "name": "muspinsim",
        "spins": ["mu", "e"],
        "couplings": [],
        "polarization": [[1.0, 0.0, 0.0]],
        "field": [[0.0, 0.0, 0.0]],
        "intrinsic_field": [[0.0, 0.0, 0.0]],
        "time": np.linspace(0.0, 10.0, 101),
        "orientation": [[0.0, 0.0, 0.0]],
        "orientation_mode": "zyz",
        "temperature": [np.inf],
        "x_axis": "time",
        "y_axis": "asymmetry",
        "average_axes": ["orientation"],
        "celio": [0],  # [k] or [k, averages] (keyword.py `celio`; 0 = do not use Celio's method)
        "results_function": None,
    }


# ------------------------------------------------------------------------------------------
# quaternions, vectorised (ase.quaternions semantics; simconfig.py:596-613, utils.py:53-68)
# ------------------------------------------------------------------------------------------
def quat_mul(a, b):
    """Hamilton product of [...,4] arrays, q = (w, x, y, z)."""
    aw, ax, ay, az = np.moveaxis(a, -1, 0)
    bw, bx, by, bz = np.moveaxis(b, -1, 0)
    return np.stack(
        [
            aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by + ay * bw + az * bx - ax * bz,
            aw * bz + az * bw + ax * by - ay * bx,
        ],
        axis=-1,
    )


def _axis_quat(axis, theta):
    theta = np.asarray(theta, dtype=float)
    q = np.zeros(theta.shape + (4,))
    q[..., 0] = np.cos(theta / 2.0)
    q[..., 1 + axis] = np.sin(theta / 2.0)
    return q


def quat_from_euler(a, b, c, mode="zyz"):
    """q = q_z(c) * q_{y|x}(b) * q_z(a)."""
    if mode not in ("zyz", "zxz"):
        raise ValueError("Invalid Euler angles mode {0}".format(mode))
    qb = _axis_quat(1 if mode == "zyz" else 0, b)
    return quat_mul(quat_mul(_axis_quat(2, c), qb), _axis_quat(2, a))


def quat_rotation_matrices(q):
    """[...,4] -> [...,3,3] so that R @ v == Quaternion(q).rotate(v)."""
    w, x, y, z = np.moveaxis(q, -1, 0)
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = w * w + x * x - y * y - z * z
    R[..., 0, 1] = 2 * (x * y - w * z)
    R[..., 0, 2] = 2 * (x * z + w * y)
    R[..., 1, 0] = 2 * (x * y + w * z)
    R[..., 1, 1] = w * w - x * x + y * y - z * z
    R[..., 1, 2] = 2 * (y * z - w * x)
    R[..., 2, 0] = 2 * (x * z - w * y)
    R[..., 2, 1] = 2 * (y * z + w * x)
    R[..., 2, 2] = w * w - x * x - y * y + z * z
    return R


def orientation_table(rows, mode="zyz"):
    """`orientation` rows (2: polar theta,phi; 3: Euler; 4: Euler + weight) ->
    (conjugate quaternions [n,4], weights [n]).  simconfig.py:596-613."""
    rows = np.atleast_2d(np.asarray(rows, dtype=float))
    n, k = rows.shape
    w = np.ones(n)
    if k == 2:
        q = quat_from_euler(rows[:, 1], rows[:, 0], rows[:, 1], "zyz")
    elif k == 3:
        q = quat_from_euler(rows[:, 0], rows[:, 1], rows[:, 2], mode)
    elif k == 4:
        q = quat_from_euler(rows[:, 0], rows[:, 1], rows[:, 2], mode)
        w = rows[:, 3].copy()
    else:
        raise ValueError("Invalid orientation row")
    return q * np.array([1.0, -1.0, -1.0, -1.0]), w


def eulrange(N):
    """N^3 Euler-angle rows with weights sin(b).  utils.py:40-50."""
    N = int(N)
    a = np.linspace(0, 2 * np.pi, N)
    b = np.linspace(0, np.pi, N + 2)[1:-1]
    c = np.linspace(0, 2 * np.pi, N)
    a, b, c = np.array(np.meshgrid(a, b, c)).reshape((3, -1))
    return np.array([a, b, c, np.sin(b)]).T


def zcw(N, mode="sphere"):
    """At least N polar-angle rows (theta, phi) of the Zaremba-Conroy-Wolfsberg set (utils.py:34-37
    calls soprano.calculate.powder.ZCW(mode).get_orient_angles(N)[0]).  Restated from the published
    algorithm (Eden & Levitt, J. Magn. Reson. 132, 220 (1998)): N_m = g(m+2) points with
    g = 8, 13, 21, ...;  phi_j = 2 pi / c3 * frac(j g(m) / N_m),  theta_j = acos(c1 (c2 frac(j / N_m) - 1)).
    PARITY UNPINNED: Soprano is not available here and the reference's tests only check the row
    count and <3 cos^2 theta - 1> ~ 0 (tests/test_input.py:271-273, test_utils.py:67-73), both of
    which this satisfies; parity and benchmark inputs use explicit orientation rows."""
    c = {"sphere": (1.0, 2.0, 1.0), "hemisphere": (-1.0, 1.0, 1.0), "octant": (2.0, 1.0, 8.0)}[mode]
    g = [8, 13]
    m = 0
    while True:
        while len(g) <= m + 2:
            g.append(g[-1] + g[-2])
        if g[m + 2] >= N:
            break
        m += 1
    Nz, gm = g[m + 2], g[m]
    j = np.arange(Nz, dtype=float)
    phi = 2 * np.pi / c[2] * np.mod(j * gm / Nz, 1.0)
    theta = np.arccos(c[0] * (c[1] * np.mod(j / Nz, 1.0) - 1.0))
    return np.array([theta, phi]).T


def _vec3_rows(rows):
    out = []
    for r in rows:
        r = np.atleast_1d(np.asarray(r, dtype=float))
        if len(r) == 1:
            r = np.array([0.0, 0.0, r[0]])  # scalar field is along z (simconfig.py:571-586)
        elif len(r) != 3:
            raise ValueError("Invalid magnetic field value")
        out.append(r)
    return np.array(out)


def expand_from_axes(ax, n_cfg, first=0, step=1):
    """What expand_configs_kernel (csrc/musim.cu) computes, in numpy: configurations
    c = first + i * step -> (B[n,3], p[n,3], T[n], w[n], slot[n])."""
    c = first + step * np.arange(n_cfg, dtype=np.int64)
    idx = [(c // ax["div"][a]) % ax["len"][a] for a in range(5)]
    R = quat_rotation_matrices(ax["quat"])[idx[3]]
    B = np.einsum("nij,nj->ni", R, ax["Blab"][idx[1]]) + ax["Bint"][idx[2]]
    p = np.einsum("nij,nj->ni", R, ax["pol"][idx[0]])
    slot = sum(idx[a] * ax["slot_mult"][a] for a in range(5))
    return B, p, ax["Tv"][idx[4]], ax["ow"][idx[3]], np.asarray(slot, dtype=np.int32)


class ConfigTable:
    """All configurations of one simulation as arrays.

    Attributes
      B, p        [n,3] field / polarisation in the crystallite frame (rotated; intrinsic
                  field added unrotated, experiment.py:404-411)
      T, w        [n]   temperature, orientation weight / avg_N
      slot        [n]   row of the [n_slots, nt] accumulation buffer
      fast        [n]   bool: the reference's T=inf / B=0 predicate (experiment.py:413-418)
      times       [nt]  time axis (length 1 and unused for y = integral)
      results_shape     shape of the reference's results array (simconfig.py:293-297)
    """

    def __init__(self, spec):
        s = default_spec()
        s.update(spec)
        self.spec = s
        self.y = s["y_axis"]
        if self.y not in ("asymmetry", "integral"):
            raise ValueError("Invalid value '%s', accepts ['asymmetry', 'integral']" % self.y)
        try:
            xname = _KEYS[s["x_axis"]]
            avg = [_KEYS[a] for a in s["average_axes"] if a.lower() != "none"]
        except KeyError as exc:
            raise ValueError("Invalid axis name") from exc

        vals = {}
        pol = np.atleast_2d(np.asarray(s["polarization"], dtype=float))
        if pol.shape[1] != 3:
            raise ValueError("Invalid muon polarization direction")
        vals["mupol"] = pol / np.linalg.norm(pol, axis=1)[:, None]  # simconfig.py:588-594
        vals["B"] = _vec3_rows(s["field"])
        vals["intrinsic_B"] = _vec3_rows(s["intrinsic_field"])
        t = np.asarray(s["time"], dtype=float).reshape(-1)
        if self.y == "integral":
            t = np.array([np.inf])  # simconfig.py:145-149
        vals["t"] = t
        q, ow = orientation_table(s["orientation"], s["orientation_mode"])
        if "orient" in avg:
            ow = ow * (len(ow) / np.sum(ow))  # weights normalised to sum to N (simconfig.py:152-156)
        else:
            ow = np.ones(len(ow))
        vals["orient"] = q
        vals["T"] = np.asarray(s["temperature"], dtype=float).reshape(-1)
        self._build(vals, ow, xname, avg)

    @classmethod
    def from_values(cls, vals, orient_weights, x_name, avg_names, y_axis):
        """Build from already validated values: vals = {"mupol": [n,3] unit vectors, "B": [n,3],
        "intrinsic_B": [n,3], "t": [nt], "orient": [n,4] conjugate quaternions, "T": [n]},
        weights already normalised (used by adapter.table_from_reference)."""
        self = cls.__new__(cls)
        self.spec = None
        self.y = y_axis
        self._build({k: np.asarray(v, dtype=float) for k, v in vals.items()}, np.asarray(orient_weights, float),
                    x_name, list(avg_names))
        return self

    def _build(self, vals, ow, xname, avg):
        t = vals["t"]
        # classify ranges in the reference's keyword order
        self.x_name = xname
        self.file_ranges, self.avg_ranges = OrderedDict(), OrderedDict()
        x_len = None
        for cname in _KEYS.values():
            n = len(vals[cname])
            if n > 1:
                if cname == xname:
                    x_len = n
                elif cname in avg:
                    self.avg_ranges[cname] = n
                else:
                    self.file_ranges[cname] = n
        if self.y == "integral" and xname == "t":
            raise ValueError("Can not use time as X axis when evaluating integral of signal")
        if x_len is None:
            raise ValueError("Specified x axis is not a range")
        self.x_len = x_len
        self.time_isavg = "t" in avg
        self.times = t
        self.x_axis_values = vals[xname]
        self.results_shape = tuple(self.file_ranges.values()) + (x_len,)

        # enumerate product(file, avg, x) without the time axis (time is the inner dimension
        # of every evaluation: make_configs uses slice(None) for it, simconfig.py:300-309)
        axes = []  # (name, length, kind)
        for k, n in self.file_ranges.items():
            axes.append((k, n, "f"))
        for k, n in self.avg_ranges.items():
            axes.append((k, n, "a"))
        axes.append((xname, x_len, "x"))
        self._loop_axes = [(k, n, kind) for (k, n, kind) in axes if k != "t"]
        shape = [n for (_, n, _) in self._loop_axes]
        self.n_cfg = int(np.prod(shape)) if shape else 1
        self.avg_N = int(np.prod([n for (k, n, kind) in axes if kind == "a" and k != "t"])) if axes else 1
        # (if time is an averaged axis the reference counts it as ONE average configuration,
        #  because make_configs gives [slice(None)] for it, and averages the slice instead)

        # slot: index into the non-time dimensions of `results`, in results order
        self._slot_dims = [(k, n) for (k, n, kind) in axes if kind in "fx" and k != "t"]
        self._slot_shape = tuple(n for (_, n) in self._slot_dims) or (1,)
        self.n_slots = int(np.prod(self._slot_shape))
        # where the time axis sits in `results`
        names_f = list(self.file_ranges.keys())
        self._t_pos = names_f.index("t") if "t" in names_f else None
        self._vals = vals
        self._ow = np.asarray(ow, dtype=float)
        self._arrays = None  # B, p, T, w, slot, fast: built on first use (see axes_descriptor)

    # ---- per-configuration arrays (host expansion; the GPU path can expand on the device) ----
    def _materialise(self):
        if self._arrays is not None:
            return self._arrays
        vals, ow = self._vals, self._ow
        shape = [n for (_, n, _) in self._loop_axes]
        n_cfg = self.n_cfg
        grids = np.unravel_index(np.arange(n_cfg), shape) if shape else ()
        idx = {k: np.zeros(n_cfg, dtype=np.int64) for k in _KEYS.values()}
        for (k, _, _), g in zip(self._loop_axes, grids):
            idx[k] = g
        slot = np.zeros(n_cfg, dtype=np.int64)
        for k, n in self._slot_dims:
            slot = slot * n + idx[k]
        # rotate lab-frame field and polarisation by the stored (conjugate) quaternion
        R = quat_rotation_matrices(vals["orient"])[idx["orient"]]
        B = np.einsum("nij,nj->ni", R, vals["B"][idx["B"]]) + vals["intrinsic_B"][idx["intrinsic_B"]]
        p = np.einsum("nij,nj->ni", R, vals["mupol"][idx["mupol"]])
        T = vals["T"][idx["T"]].astype(float)
        w = ow[idx["orient"]] / self.avg_N
        Bn = np.linalg.norm(B, axis=1)
        with np.errstate(divide="ignore", invalid="ignore"):
            check = (cnst.e * (cnst.hbar**2) * Bn) / (2 * cnst.m_p * cnst.k * T)
        self._arrays = {"B": B, "p": p, "T": T, "w": w, "slot": slot.astype(np.int32), "fast": check == 0}
        return self._arrays

    B = property(lambda self: self._materialise()["B"])
    p = property(lambda self: self._materialise()["p"])
    T = property(lambda self: self._materialise()["T"])
    w = property(lambda self: self._materialise()["w"])
    slot = property(lambda self: self._materialise()["slot"])
    fast = property(lambda self: self._materialise()["fast"])  # experiment.py:413-418

    def uniform_fast(self):
        """True / False if EVERY configuration takes / does not take the reference's T = inf | B = 0
        fast path (experiment.py:413-418), None if they differ or it cannot be told from the axis
        tables alone (then the per-configuration arrays are needed)."""
        T = self._vals["T"]
        if np.all(np.isinf(T) & (T > 0)):
            return True
        if np.any(self._vals["intrinsic_B"] != 0.0):
            return None  # |R B_lab + B_int| depends on the orientation
        Bn = np.linalg.norm(self._vals["B"], axis=1)  # |R B_lab| = |B_lab|
        with np.errstate(divide="ignore", invalid="ignore"):
            check = (cnst.e * (cnst.hbar**2) * Bn[:, None]) / (2 * cnst.m_p * cnst.k * T[None, :])
        f = check == 0
        # |R B| is |B| only up to rounding: a field whose rotated norm could round to exactly 0 is 0
        if f.all():
            return True
        if not f.any():
            return False
        return None

    def axes_descriptor(self):
        """Axis tables for the device-side expansion (musim_run_axes_host): global configuration
        c -> idx_a = (c // div[a]) % len[a] for a = (mupol, B, intrinsic_B, orient, T), enumerated
        with the file axes slowest, then the x axis, then the averaged axes (so that the output
        row index is non-decreasing in c), and slot = sum_a idx_a * slot_mult[a]."""
        names = ["mupol", "B", "intrinsic_B", "orient", "T"]
        order = [(k, n) for (k, n, kind) in self._loop_axes if kind == "f"]
        order += [(k, n) for (k, n, kind) in self._loop_axes if kind == "x"]
        order += [(k, n) for (k, n, kind) in self._loop_axes if kind == "a"]
        ln = {k: 1 for k in names}
        div = {k: 1 for k in names}
        acc = 1
        for k, n in reversed(order):
            ln[k] = n
            div[k] = acc
            acc *= n
        smul = {k: 0 for k in names}
        acc = 1
        for k, n in reversed(self._slot_dims):
            smul[k] = acc
            acc *= n
        v = self._vals
        return {
            "len": np.array([ln[k] for k in names], dtype=np.int64),
            "div": np.array([div[k] for k in names], dtype=np.int64),
            "slot_mult": np.array([smul[k] for k in names], dtype=np.int64),
            "pol": np.ascontiguousarray(v["mupol"], dtype=float),
            "Blab": np.ascontiguousarray(v["B"], dtype=float),
            "Bint": np.ascontiguousarray(v["intrinsic_B"], dtype=float),
            "quat": np.ascontiguousarray(v["orient"], dtype=float),
            "ow": np.ascontiguousarray(self._ow / self.avg_N, dtype=float),
            "Tv": np.ascontiguousarray(v["T"], dtype=float),
        }

    def expand_axes(self, first=0, step=1):
        """numpy restatement of the device expansion kernel (tests; same enumeration order)."""
        return expand_from_axes(self.axes_descriptor(), len(range(first, self.n_cfg, step)), first, step)

    def finish(self, out):
        """[n_slots, nt] accumulation buffer -> the reference's results array layout."""
        out = np.asarray(out)
        if self.y == "integral":
            return out.reshape(self.results_shape)
        if self.time_isavg:
            avg = out.mean(axis=1)  # simconfig.py:364-365
            if self.x_name == "t":
                # time is both the x axis and an averaged axis: the reference stores the time average
                # in every x column (store_time_slice broadcasts the scalar over the slice)
                return np.repeat(avg[:, None], self.x_len, axis=1).reshape(self.results_shape)
            return avg.reshape(self.results_shape)
        if self.x_name == "t":
            return out.reshape(self.results_shape)
        if self._t_pos is None:
            # a single time value reaches the solver as a 0-d array (validation.py:19-20)
            raise ValueError("times must be an array of values in microseconds")
        full = out.reshape(self._slot_shape + (out.shape[1],))
        return np.moveaxis(full, -1, self._t_pos).reshape(self.results_shape)

    # ---- output side (simconfig.py:370-429) ----------------------------------------------
    def _print_value(self, key, idx):
        """Header text of one file-range value (the reference's _print_B / _print_orient, simconfig.py:638-647)."""
        v = self._vals[key][idx]
        if key == "B":
            return "{0} T".format(np.asarray(v))
        if key == "orient":
            a, b, c = quat_to_zyz(v)
            return "[ZYZ] a = {0:.1f} deg, b = {1:.1f} deg, c = {2:.1f} deg, weight = {3}".format(
                np.degrees(a), np.degrees(b), np.degrees(c), self._ow[idx])
        return np.asarray(v) if np.ndim(v) else float(v)

    def save_output(self, results, name=None, path=".", extension=".dat"):
        """Write one two-column file per file-range index tuple, `<name>[_i[_j...]].dat`, with the
        reference's header; the x column is |B| when the x axis is a field.  Returns the file names."""
        import datetime
        import os
        from itertools import product

        from .constants import REFERENCE_VERSION

        results = np.asarray(results)
        if results.shape != self.results_shape:
            raise ValueError("results do not have the shape of this configuration")
        if name is None:
            name = (self.spec or {}).get("name", "muspinsim")
        header = "MUSPINSIM v.{0}\nOutput file written on {1}\nParameters used:\n".format(
            REFERENCE_VERSION, datetime.datetime.now().ctime())
        x = np.asarray(self.x_axis_values)
        if self.x_name in ("B", "intrinsic_B"):
            x = np.linalg.norm(x, axis=-1)
        written = []
        keys = list(self.file_ranges.keys())
        for inds in product(*[range(n) for n in self.file_ranges.values()]):
            fname = os.path.join(path, "{0}{1}{2}".format(name, "".join("_%d" % i for i in inds), extension))
            hdr = header + "".join("\t{0:<20} = {1}\n".format(k, self._print_value(k, i)) for k, i in zip(keys, inds))
            data = np.zeros((len(x), 2))
            data[:, 0] = x
            data[:, 1] = results[inds]
            np.savetxt(fname, data, header=hdr)
            written.append(fname)
        return written


def quat_to_zyz(q):
    """Euler angles (a, b, c) with q = q_z(c) q_y(b) q_z(a) (ase Quaternion.euler_angles('zyz'))."""
    R = quat_rotation_matrices(np.asarray(q, dtype=float)[None])[0]
    b = np.arccos(np.clip(R[2, 2], -1.0, 1.0))
    if abs(np.sin(b)) > 1e-12:
        return np.arctan2(R[2, 1], -R[2, 0]), b, np.arctan2(R[1, 2], R[0, 2])
    a = np.arctan2(R[1, 0], R[0, 0]) if R[2, 2] > 0 else np.arctan2(-R[1, 0], -R[0, 0])
    return a, b, 0.0
