"""Isotope table lookup restating soprano.nmr.utils._get_isotope_data for the handful of
isotopes the parity/benchmark inputs use.  gamma in rad/(s T), Q in millibarn, I in hbar.
Pinned by the reference only for H (tests/test_constants.py:13-43)."""
import numpy as np

# element -> {isotope: (gamma, Q, I)}; first key = most abundant isotope
_TABLE = {
    "H": {1: (267522128.0, 0.0, 0.5), 2: (41066279.1, 2.86, 1.0)},
    "C": {12: (0.0, 0.0, 0.0), 13: (67282840.0, 0.0, 0.5)},
    "N": {14: (19337792.0, 20.44, 1.0), 15: (-27126180.4, 0.0, 0.5)},
    "F": {19: (251814800.0, 0.0, 0.5)},
    "V": {51: (70455117.0, -52.0, 3.5)},
    "Cu": {63: (71117890.0, -220.0, 1.5), 65: (76043500.0, -204.0, 1.5)},
}
_KEYS = {"gamma": 0, "Q": 1, "I": 2}


def _get_isotope_data(elems, key, isotopes=None, isotope_list=None, use_q_isotopes=False):
    out = []
    for i, el in enumerate(elems):
        if el not in _TABLE:
            raise RuntimeError("No NMR data on element {0}".format(el))
        iso = None
        if isotope_list is not None and isotope_list[i] is not None:
            iso = isotope_list[i]
        elif isotopes is not None and el in isotopes:
            iso = isotopes[el]
        if iso is None:
            iso = next(iter(_TABLE[el]))
        if iso not in _TABLE[el]:
            raise RuntimeError("No NMR data on isotope {0}{1}".format(iso, el))
        out.append(_TABLE[el][iso][_KEYS[key]])
    return np.array(out, dtype=float)
