"""Wall time of the streamed Celio path (launch bound: nt x (k n_gates + 1) small kernels) on a mu + 2 x 51V system,
forced through `streamed=True`; prints ms per call and launches per call."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from test_gpu_celio import _setup  # noqa: E402

from muspinsim_b200 import _lib  # noqa: E402
import os  # noqa: E402

if os.environ.get("MUSIM_LIB"):
    _lib.LIB_PATH = os.environ["MUSIM_LIB"]

s, H, gates = _setup("mu_2V", 3, 0.03)
dim = s.dim_total
rng = np.random.default_rng(3)
n_states, nt = 8, 400
psi = rng.normal(size=(n_states, dim)) + 1j * rng.normal(size=(n_states, dim))
psi /= np.linalg.norm(psi, axis=1)[:, None]
sigma = s.sigma_mu([0.3, -0.5, 0.8])
for streamed in (True, False):
    out = np.zeros(nt)
    _lib.celio_evolve(0, psi, sigma, 3, gates, nt, out, streamed=streamed)
    l0 = _lib.load().musim_celio_launch_count()
    t0 = time.perf_counter()
    for _ in range(5):
        _lib.celio_evolve(0, psi, sigma, 3, gates, nt, out, streamed=streamed)
    dt = (time.perf_counter() - t0) / 5
    print("streamed" if streamed else "resident", "dim", dim, "ms/call %.2f" % (dt * 1e3), "launches/call", (_lib.load().musim_celio_launch_count() - l0) // 5)
