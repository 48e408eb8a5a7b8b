"""muspinsim_b200 -- B200-native (sm_100a) implementation of MuSpinSim's data-parallel hot path.

Host side: Python mirror of the reference's interface for this path (ExperimentRunner,
MuonSpinSystem, Hamiltonian, Lindbladian); arithmetic: hand-written CUDA behind the C ABI of
include/musim.h (muspinsim_b200/csrc/libmusim.so).  There is no CPU fallback.
"""

from .configs import ConfigTable  # noqa: F401
from .experiment import ExperimentRunner  # noqa: F401
from .spinsys import MuonSpinSystem, SpinSystem  # noqa: F401

__version__ = "0.1.0"
