"""ZCW powder-averaging angles (Zaremba-Conroy-Wolfsberg), restated from the published
algorithm (Eden & Levitt, J. Magn. Reson. 132, 220 (1998)).  Parity UNPINNED: parity and
benchmark inputs use explicit orientation rows, never zcw(N)."""
import numpy as np


class ZCW:
    _modes = {"sphere": (1.0, 2.0, 1.0), "hemisphere": (-1.0, 1.0, 1.0), "octant": (2.0, 1.0, 8.0)}

    def __init__(self, mode="sphere"):
        self._c = self._modes[mode]

    @staticmethod
    def _g(m):
        g = [8, 13]
        while len(g) <= m:
            g.append(g[-1] + g[-2])
        return g[m]

    def get_orient_angles(self, N):
        m = 0
        while self._g(m + 2) < N:
            m += 1
        Nz = self._g(m + 2)
        gm = self._g(m)
        j = np.arange(Nz, dtype=float)
        c = self._c
        phi = 2 * np.pi / c[2] * np.mod(j * gm / Nz, 1.0)
        theta = np.arccos(c[0] * (c[1] * np.mod(j / Nz, 1.0) - 1.0))
        return np.array([theta, phi]).T, np.ones(Nz) / Nz
