"""Golden parity cases: small seeded specs covering every branch of the hot path and the edge
cases of the configuration expansion.  TEST INFRASTRUCTURE.  Used by oracle/make_golden.py
(which runs the UNMODIFIED reference on them) and by tests/."""
import numpy as np

from muspinsim_b200 import workloads as wl


def cases():
    c = {}
    c["c1_hfine"] = wl.c1_hfine()
    c["c2_fast_d16"] = wl.c2_hfine_powder(n_orient=12, nt=200, n_h=2)
    c["c2_general_d8_T0p3"] = wl.c2_hfine_powder(n_orient=8, nt=64, n_h=1, temperature=0.3)
    c["c2_fast_d32"] = wl.c2_hfine_powder(n_orient=4, nt=1000, n_h=3)
    c["c3_alc_d12"] = wl.c3_alc(n_orient=6, n_field=9, extra_h=False)
    c["c3_alc_d24"] = wl.c3_alc(n_orient=3, n_field=5, extra_h=True)
    c["c4_dissip_tf"] = wl.c4_fmuf_dissipation(n_orient=4, nt=50)
    c["c4_dissip_zf"] = wl.c4_fmuf_dissipation(n_orient=4, nt=50, zero_field=True)
    c["c5_fast_d96"] = wl.c5_large(n_orient=2, nt=100)
    c["c5_general_d96_T1"] = wl.c5_large(n_orient=2, nt=40, temperature=1.0)

    # temperature file range x field x-axis, integral; muon not first
    s = wl.c3_alc(n_orient=3, n_field=4, extra_h=False)
    s["name"] = "alc_T_filerange"
    s["temperature"] = [0.05, 2.0, np.inf]
    s["field"] = [[0.0, 0.0, b] for b in (0.0, 0.5, 1.9, 2.1)]
    c["alc_T_filerange"] = s

    # weighted orientations (eulrange rows carry sin(b) weights), hfine_powder example
    from muspinsim_b200.configs import eulrange

    c["hfine_powder_eulrange3"] = {
        "name": "hfine_powder_eulrange3",
        "spins": ["mu", "e"],
        "couplings": [{"type": "hyperfine", "i": 1, "value": np.array([[5.0, 2, 3], [2, 5, 2], [3, 2, 5]])}],
        "field": [[0.0, 0.0, 0.01]],
        "time": np.linspace(0.0, 1.0, 100),
        "orientation": eulrange(3),
    }

    # intrinsic field scan (x axis), starting away from zero (see SURVEY.md appendix C for the
    # reference's stale-cache quirk when it starts at zero), zxz orientations
    rng = np.random.default_rng(11)
    c["intrinsic_scan_zxz"] = {
        "name": "intrinsic_scan_zxz",
        "spins": ["mu", "e", "H"],
        "couplings": [
            {"type": "hyperfine", "i": 1, "value": np.diag([30.0, 30.0, 50.0])},
            {"type": "hyperfine", "i": 3, "j": 2, "value": wl._sym(rng, 4.0)},
        ],
        "field": [[0.0, 0.002, 0.001]],
        "intrinsic_field": [[0.0, 0.0, b] for b in (0.001, 0.002, 0.004)],
        "polarization": [[0.0, 0.0, 1.0]],
        "orientation": rng.uniform(0, np.pi, size=(5, 3)),
        "orientation_mode": "zxz",
        "x_axis": "intrinsic_field",
        "y_axis": "integral",
    }

    # non-uniform explicit times not starting at zero
    c["nonuniform_times"] = dict(
        wl.c2_hfine_powder(n_orient=5, nt=10, n_h=1),
        name="nonuniform_times",
        time=np.array([0.013, 0.02, 0.5, 0.51, 1.7, 2.0, 3.3, 3.30001, 7.0, 9.99]),
    )

    # time averaged, field on the x axis, finite temperature
    c["time_averaged_vs_field"] = dict(
        wl.c2_hfine_powder(n_orient=3, nt=16, n_h=1, temperature=5.0),
        name="time_averaged_vs_field",
        field=[[0.0, 0.0, b] for b in (0.0, 0.01, 0.3)],
        x_axis="field",
        average_axes=["orientation", "time"],
    )

    # polarisation as a file range, single crystal
    c["polarization_filerange"] = {
        "name": "polarization_filerange",
        "spins": ["mu", "e"],
        "couplings": [{"type": "hyperfine", "i": 1, "value": np.array([[5.0, 2, 3], [2, 5, 2], [3, 2, 5]])}],
        "field": [[0.001, 0.0, 0.01]],
        "polarization": [[1.0, 0.0, 0.0], [0.0, 0.0, 1.0], [1.0, 1.0, 1.0]],
        "time": np.linspace(0.0, 2.0, 40),
        "orientation": [[0.3, 0.8, 1.1]],
    }

    # T = 0 (ground state only) and a quadrupolar nucleus, muon last
    c["ground_state_T0"] = {
        "name": "ground_state_T0",
        "spins": ["e", "2H", "mu"],
        "couplings": [
            {"type": "hyperfine", "i": 3, "value": np.diag([100.0, 100.0, 120.0])},
            {"type": "hyperfine", "i": 2, "value": np.diag([10.0, 12.0, 14.0])},
            {"type": "quadrupolar", "i": 2, "value": wl._sym_traceless(rng, 0.3)},
        ],
        "field": [[0.0, 0.0, 0.05]],
        "temperature": [0.0],
        "time": np.linspace(0.0, 0.5, 30),
        "orientation": rng.uniform(0, np.pi, size=(3, 2)),
    }

    # dissipation at finite temperature, single spin (tests/test_experiment.py:581-640)
    c["dissip_thermal"] = {
        "name": "dissip_thermal",
        "spins": ["mu"],
        "couplings": [{"type": "dissipation", "i": 1, "value": 1.0}],
        "field": [[0.0, 0.0, -1.0]],
        "temperature": [0.1],
        "polarization": [[1.0, 0.0, 1.0]],
        "time": np.linspace(0.0, 3.0, 31),
    }
    # dissipation, integral mode, field scan
    c["dissip_integral"] = dict(
        wl.c4_fmuf_dissipation(n_orient=3, nt=5),
        name="dissip_integral",
        field=[[0.0, 0.0, b] for b in (0.0, 0.005, 0.02)],
        polarization=[[0.0, 0.0, 1.0]],
        x_axis="field",
        y_axis="integral",
    )
    return c
