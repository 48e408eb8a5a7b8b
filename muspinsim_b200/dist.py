"""Process-group plumbing: the role of /root/reference/muspinsim/mpi.py (MPIController).

One process per GPU.  Configurations are sharded round-robin exactly like the reference
(`self._config[mpi.rank :: mpi.size]`, experiment.py:369); the only collective of the path is
the final sum of the [n_slots, n_x] float64 results (`MPIController.sum_data` =
`comm.Reduce(SUM)`, mpi.py:104-112), done here with one torch.distributed all-reduce (NCCL over
NVLink on GPUs, gloo in the CPU tests).  The message is 8-16 KB, i.e. pure latency.
"""

import os

import numpy as np


class Communicator:
    def __init__(self, backend=None, device=None):
        import torch
        import torch.distributed as dist

        self._torch, self._dist = torch, dist
        if not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29512")
            kw = {}
            if backend == "nccl":
                local = int(os.environ.get("LOCAL_RANK", "0")) if device is None else int(device)
                torch.cuda.set_device(local)
                kw["device_id"] = torch.device("cuda", local)
            dist.init_process_group(
                backend=backend,
                rank=int(os.environ.get("RANK", "0")),
                world_size=int(os.environ.get("WORLD_SIZE", "1")),
                **kw,
            )
            self._owns_group = True
        self.backend = dist.get_backend()
        self.rank = dist.get_rank()
        self.size = dist.get_world_size()
        if device is None and self.backend == "nccl":
            device = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(device)
        self.device = device

    def close(self):
        if getattr(self, "_owns_group", False) and self._dist.is_initialized():
            self._dist.destroy_process_group()
            self._owns_group = False

    @property
    def is_root(self):
        return self.rank == 0

    def sum_data(self, data):
        """All ranks get the sum (a superset of the reference's Reduce-to-root)."""
        torch = self._torch
        t = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float64))
        if self.backend == "nccl":
            t = t.cuda(self.device)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def sum_tensor_(self, t):
        """In-place all-reduce of a device tensor (no host round trip)."""
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM)
        return t

    def max_float(self, x):
        torch = self._torch
        t = torch.tensor([float(x)], dtype=torch.float64)
        if self.backend == "nccl":
            t = t.cuda(self.device)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX)
        return float(t.item())

    def barrier(self):
        self._dist.barrier()

    def broadcast_object(self, obj, root=0):
        """mpi.py:54-102 broadcast*: pickled object broadcast for set-up data."""
        lst = [obj]
        self._dist.broadcast_object_list(lst, src=root)
        return lst[0]
