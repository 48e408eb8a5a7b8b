"""Physical constants of the path, mirroring /root/reference/muspinsim/constants.py:10-108.

The reference reads nuclear data through the third-party Soprano package; the handful of
isotopes used by the example inputs and the benchmark workloads are tabulated here
(gamma in rad/(s T), Q in millibarn, spin in hbar).  More can be registered at run time with
`register_isotope`.
"""

import re

import numpy as np
from scipy import constants as cnst

ELEC_GAMMA = -28024.9514242  # MHz/T  (constants.py:12)
MU_GAMMA = -(ELEC_GAMMA / 206.7669883)  # MHz/T  (constants.py:13)
MU_TAU = 2.19703  # us     (constants.py:14)
EFG_2_MHZ = (
    cnst.physical_constants["atomic unit of electric field gradient"][0] * cnst.e * 1e-37 / cnst.h
)  # constants.py:20-25

_ISOTOPES = {
    "H": {1: (267522128.0, 0.0, 0.5), 2: (41066279.1, 2.86, 1.0)},
    "C": {12: (0.0, 0.0, 0.0), 13: (67282840.0, 0.0, 0.5)},
    "N": {14: (19337792.0, 20.44, 1.0), 15: (-27126180.4, 0.0, 0.5)},
    "F": {19: (251814800.0, 0.0, 0.5)},
    "V": {51: (70455117.0, -52.0, 3.5)},
    "Cu": {63: (71117890.0, -220.0, 1.5), 65: (76043500.0, -204.0, 1.5)},
}


def register_isotope(element, mass_number, gamma_rad_s_T, Q_mb, spin):
    _ISOTOPES.setdefault(element, {})[int(mass_number)] = (float(gamma_rad_s_T), float(Q_mb), float(spin))


def parse_spin(label):
    """'mu' | 'e' | 'F' | '14N' | ('N', 14) -> (element, isotope|None).  simconfig.py:523-533."""
    if isinstance(label, tuple):
        return label[0], label[1]
    m = re.match(r"([0-9]+)([A-Z][a-z]*|e)$", label)
    if m:
        return m.group(2), int(m.group(1))
    return label, None


def _iso(elem, iso):
    try:
        tab = _ISOTOPES[elem]
        return tab[iso if iso is not None else next(iter(tab))]
    except KeyError as exc:
        raise ValueError(f"Invalid isotope {iso} for element {elem}") from exc


def gyromagnetic_ratio(elem="mu", iso=None):
    """MHz/T (a frequency, not a pulsation).  constants.py:28-53."""
    if elem == "e":
        return ELEC_GAMMA
    if elem == "mu":
        return MU_GAMMA
    return _iso(elem, iso)[0] / (2e6 * np.pi)


def quadrupole_moment(elem="mu", iso=None):
    """constants.py:56-77 (value as tabulated: millibarn)."""
    if elem in ("e", "mu"):
        return 0
    return _iso(elem, iso)[1]


def spin(elem="mu", iso=None):
    """constants.py:80-108."""
    if elem == "mu":
        return 0.5
    if elem == "e":
        iso = iso or 1
        if iso < 1 or int(iso) != iso:
            raise ValueError(f"Invalid multiplicity {iso} for electron")
        return 0.5 * int(iso)
    return _iso(elem, iso)[2]

# version string the reference writes into its .dat headers (muspinsim/version.py:5)
REFERENCE_VERSION = "2.3.1"
