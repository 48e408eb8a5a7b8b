"""Minimal stand-in for qutip==5.1.0 (absent from this image).  TEST INFRASTRUCTURE ONLY.
The reference's hot path only needs the three Pauli matrices as CSR (spinsys.py:754-756);
`Qobj` is only touched by Celio's method, which is out of scope."""
from .core.operators import sigmax, sigmay, sigmaz  # noqa: F401


class Qobj:  # pragma: no cover
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("qutip.Qobj is not available in the oracle shim")
