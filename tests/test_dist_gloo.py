"""The N > 1 path on CPU: world_size-2 (and 3) gloo jobs.  Each rank takes the round-robin
shard cfg[rank::size] (experiment.py:369), one all-reduce merges the results
(mpi.py:104-112).  Compute is the CPU stand-in handle; this tests sharding + the collective."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("name,world", [("c2_fast_d16", 2), ("alc_T_filerange", 2), ("c4_dissip_tf", 3)])
def test_sharded_run_matches_golden(tmp_path, name, world):
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_gloo_worker.py"), name,
                                       str(tmp_path)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=300)
        assert p.returncode == 0, out.decode()[-2000:]
    parts = []
    for rank in range(world):
        z = np.load(tmp_path / ("rank%d.npz" % rank))
        assert np.max(np.abs(z["got"] - z["want"])) < 1e-10  # every rank holds the full sum
        assert z["mx"] == world - 1
        parts.append(int(z["part"]))
        total = int(z["total"])
    assert sum(parts) == total and max(parts) - min(parts) <= 1  # balanced round-robin shards
