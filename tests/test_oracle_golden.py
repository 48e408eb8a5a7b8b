"""The numpy oracle against the golden vectors generated from the UNMODIFIED reference
(oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle import muspin_oracle as mo

TOL = 1e-11  # oracle and reference share numpy/LAPACK; differences are summation order only


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_golden(name):
    spec, want = load_golden(name)
    got = mo.run_spec(spec)
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) < TOL


def test_oracle_rank_slices_sum_to_total():
    spec, want = load_golden("c2_fast_d16")
    parts = [mo.run_spec(spec, rank=r, size=3) for r in range(3)]
    assert np.max(np.abs(sum(parts) - want)) < TOL


def test_general_path_equals_fast_path():
    spec, want = load_golden("c2_fast_d16")
    got = mo.run_spec(spec, force_general=True, evolve_fn=mo.evolve_vectorised)
    assert np.max(np.abs(got - want)) < 1e-12


def test_vectorised_evolve_equals_loop():
    spec, want = load_golden("c2_general_d8_T0p3")
    got = mo.run_spec(spec, evolve_fn=mo.evolve_vectorised)
    assert np.max(np.abs(got - want)) < TOL
