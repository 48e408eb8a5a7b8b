"""The C-ABI library loads and exports every symbol include/musim.h declares (no compute)."""
import ctypes
import os
import re

from muspinsim_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "musim.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(musim_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared() == sorted(_lib.EXPORTS)


def test_library_exports_every_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build with __graft_entry__.build()"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name


def test_version_and_argument_validation_without_gpu():
    lib = _lib.load()
    assert lib.musim_version() >= 1
    h = ctypes.c_void_p()
    # invalid arguments are rejected before any CUDA call
    assert lib.musim_create(ctypes.byref(h), 0, 0, 0, None, None, 0, None, None, None, 0, None, None) == -1
    # dissipation table: more entries than spins, or a count without arrays
    dims = (ctypes.c_int * 1)(2)
    gam = (ctypes.c_double * 1)(0.0)
    mat = (ctypes.c_double * 8)()
    m3 = (ctypes.c_double * 24)()
    assert lib.musim_create(ctypes.byref(h), 0, 2, 1, dims, gam, 0, mat, m3, m3, 17, None, None) == -1
    assert lib.musim_create(ctypes.byref(h), 0, 2, 1, dims, gam, 0, mat, m3, m3, 1, None, None) == -1
    assert lib.musim_create(ctypes.byref(h), 0, 2, 1, dims, gam, 0, mat, m3, m3, -1, None, None) == -1
    assert lib.musim_update_observables(None, None) == -1
    assert lib.musim_device_count() >= 0
    assert lib.musim_run(None, 0, 0, None, None, None, None, None, 0, None, 1.0, 1, None, None) == -1
    assert lib.musim_eigh(0, 0, 0, None, None, None, 0, None) == -1
    assert lib.musim_evolve_rho(0, 0, None, None, None, 0, None, None, None) == -1
    assert lib.musim_evolve_rho(0, 4, None, None, None, 3, None, None, None) == -1
    assert lib.musim_trim_pool(-1) == -1
    assert lib.musim_destroy(None) == 0
    assert lib.musim_run_axes_host(None, 0, 0, 0, 1, None, None, None, None, None, None, None, None, None, 0, None,
                                   1.0, 1, None) == -1
    assert lib.musim_nufft_tables(0, None, None, None, None, None) == -1
    M = ctypes.c_int()
    assert lib.musim_nufft_tables(1000, ctypes.byref(M), None, None, None, None) == 0 and M.value == 2048
