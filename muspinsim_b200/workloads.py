"""Synthetic problem definitions for the BASELINE.json configurations.

Each function returns a "spec": a plain dict whose keys map 1:1 onto the reference's `.in`
keywords (spins, couplings with 1-based indices, field, polarization, orientation rows,
temperature, time, x_axis, y_axis, average_axes).  Everything is generated from fixed seeds
with numpy.random.default_rng so the CPU oracle, the reference and the GPU path see
identical bits.  Orientations are always explicit rows (never zcw(N), whose exact grid
lives in a third-party package).
"""

import numpy as np


def _sym(rng, scale):
    a = rng.normal(0.0, scale, size=(3, 3))
    return 0.5 * (a + a.T)


def _sym_traceless(rng, scale):
    a = _sym(rng, scale)
    return a - np.eye(3) * np.trace(a) / 3.0


def _euler_rows(rng, n):
    """zyz Euler rows, isotropic: alpha, gamma ~ U[0, 2pi), cos(beta) ~ U[-1, 1]."""
    a = rng.uniform(0.0, 2 * np.pi, n)
    b = np.arccos(rng.uniform(-1.0, 1.0, n))
    c = rng.uniform(0.0, 2 * np.pi, n)
    return np.stack([a, b, c], axis=1)


def _polar_rows(rng, n):
    th = np.arccos(rng.uniform(-1.0, 1.0, n))
    ph = rng.uniform(0.0, 2 * np.pi, n)
    return np.stack([th, ph], axis=1)


def _rand_dir(rng):
    v = rng.normal(size=3)
    return v / np.linalg.norm(v)


def c1_hfine():
    """examples/hfine/hfine.in: mu + e isotropic hyperfine, single crystal, 100 time points."""
    return {
        "name": "c1_hfine",
        "spins": ["mu", "e"],
        "couplings": [{"type": "hyperfine", "i": 1, "value": np.eye(3) * 10.0}],
        "time": np.linspace(0.0, 0.1, 100),  # `range(0, 0.1)` defaults to 100 points
    }


def c2_hfine_powder(n_orient=20000, nt=1000, n_h=3, temperature=np.inf, seed=1):
    """examples/hfine_powder scaled: mu + e + n_h 1H (d = 4 * 2**n_h), powder average."""
    rng = np.random.default_rng(seed)
    spins = ["mu", "e"] + ["H"] * n_h
    cpl = [{"type": "hyperfine", "i": 1, "value": np.array([[5.0, 2, 3], [2, 5, 2], [3, 2, 5]])}]
    for k in range(n_h):
        cpl.append({"type": "hyperfine", "i": 3 + k, "j": 2, "value": _sym(rng, 5.0)})
    for k in range(n_h):
        cpl.append({"type": "dipolar", "i": 1, "j": 3 + k, "value": _rand_dir(rng) * rng.uniform(1.5, 2.5)})
    return {
        "name": "c2_hfine_powder_d%d" % (4 * 2**n_h),
        "spins": spins,
        "couplings": cpl,
        "field": [[0.0, 0.0, 0.01]],
        "temperature": [temperature],
        "orientation": _euler_rows(np.random.default_rng(seed + 1), n_orient),
        "time": np.linspace(0.0, 10.0, nt),
    }


def c3_alc(n_orient=5000, n_field=2000, extra_h=True, seed=3):
    """examples/alc scaled: avoided-level-crossing scan, integral of the longitudinal
    polarisation vs field; e + mu + 14N (+ 1H): d = 12 (24)."""
    rng = np.random.default_rng(seed)
    spins = ["e", "mu", "14N"] + (["H"] if extra_h else [])
    cpl = [
        {"type": "hyperfine", "i": 2, "value": np.array([[580.0, 5, 10], [5, 580, 9], [10, 9, 580]])},
        {"type": "hyperfine", "i": 3, "value": np.array([[150.0, 3, 4], [3, 150, 5], [4, 5, 150]])},
        {"type": "quadrupolar", "i": 3, "value": _sym_traceless(rng, 0.5)},
    ]
    if extra_h:
        cpl.append({"type": "hyperfine", "i": 4, "value": _sym(rng, 20.0)})
    return {
        "name": "c3_alc_d%d" % (24 if extra_h else 12),
        "spins": spins,
        "couplings": cpl,
        "polarization": [[0.0, 0.0, 1.0]],  # longitudinal (input/input.py:29-37 'alc')
        "field": [[0.0, 0.0, b] for b in np.linspace(1.8, 2.6, n_field)],
        "orientation": _polar_rows(np.random.default_rng(seed + 1), n_orient),
        "x_axis": "field",
        "y_axis": "integral",
    }


def c4_fmuf_dissipation(n_orient=10000, nt=1000, zero_field=False, seed=5):
    """examples/fluorine_dissipation scaled: F-mu-F with Lindbladian dissipation (d = 8)."""
    r = 0.82731493
    return {
        "name": "c4_fmuf_dissip" + ("_zf" if zero_field else "_tf"),
        "spins": ["mu", "F", "F"],
        "couplings": [
            {"type": "dipolar", "i": 1, "j": 2, "value": np.array([r, r, 0.0])},
            {"type": "dipolar", "i": 1, "j": 3, "value": np.array([-r, -r, 0.0])},
            {"type": "dissipation", "i": 2, "value": 0.1},
            {"type": "dissipation", "i": 3, "value": 0.1},
        ],
        "field": [[0.0, 0.0, 0.0]] if zero_field else [[1.27e-2, 1.27e-2, 1.27e-2]],
        "polarization": [[1.0, 0.0, 0.0]],
        "orientation": _euler_rows(np.random.default_rng(seed), n_orient),
        "time": np.linspace(0.0, 8.0, nt),
    }


def c5_large(n_orient=20000, nt=1000, temperature=np.inf, seed=6):
    """Synthetic large system: mu + e + 3 x 1H + 14N (d = 96), powder average."""
    rng = np.random.default_rng(seed)
    spins = ["mu", "e", "H", "H", "H", "14N"]
    cpl = [{"type": "hyperfine", "i": 1, "value": _sym(rng, 50.0) + 100.0 * np.eye(3)}]
    for k in (3, 4, 5, 6):
        cpl.append({"type": "hyperfine", "i": k, "j": 2, "value": _sym(rng, 10.0)})
    cpl.append({"type": "quadrupolar", "i": 6, "value": _sym_traceless(rng, 0.5)})
    return {
        "name": "c5_large_d96",
        "spins": spins,
        "couplings": cpl,
        "field": [[0.0, 0.0, 0.01]],
        "temperature": [temperature],
        "orientation": _euler_rows(np.random.default_rng(seed + 1), n_orient),
        "time": np.linspace(0.0, 10.0, nt),
    }


WORKLOADS = {
    "c1": c1_hfine,
    "c2": c2_hfine_powder,
    "c3": c3_alc,
    "c4": c4_fmuf_dissipation,
    "c5": c5_large,
}
