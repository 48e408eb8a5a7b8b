"""Worker for tests/test_dist_gloo.py: one rank of a world_size-N gloo job on CPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import load_golden  # noqa: E402
from helpers import OracleHandle  # noqa: E402
from muspinsim_b200 import ExperimentRunner  # noqa: E402
from muspinsim_b200.dist import Communicator  # noqa: E402

name, outdir = sys.argv[1], sys.argv[2]
comm = Communicator(backend="gloo")
spec, want = load_golden(name)
spec = comm.broadcast_object(spec if comm.is_root else None)
r = ExperimentRunner(spec, comm=comm)
r._handle = OracleHandle(spec)
got = r.run()
part = sum(n for _, n in r._handle.calls)
total = comm.sum_data(np.array([float(part)]))[0]
np.savez(os.path.join(outdir, "rank%d.npz" % comm.rank), got=got, want=want, part=part, total=total,
         mx=comm.max_float(float(comm.rank)))
comm.barrier()
