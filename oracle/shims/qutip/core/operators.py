import numpy as np
from scipy import sparse


class _Op:
    def __init__(self, m):
        self._m = sparse.csr_matrix(np.array(m, dtype=complex))

    def data_as(self, fmt="csr_matrix"):
        return self._m.copy()

    def full(self):
        return self._m.toarray()


def sigmax():
    return _Op([[0, 1], [1, 0]])


def sigmay():
    return _Op([[0, -1j], [1j, 0]])


def sigmaz():
    return _Op([[1, 0], [0, -1]])
