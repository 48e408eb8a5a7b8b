import sys, numpy as np, torch
sys.path.insert(0, ".")
from muspinsim_b200 import _lib
import os; _lib.LIB_PATH = os.environ.get("HSLIB", "tools/scratch/libmusim_timing.so")
def run(d, b):
    rng = np.random.default_rng(d)
    A = rng.normal(size=(b, d, d)) + 1j * rng.normal(size=(b, d, d))
    A = np.ascontiguousarray(A + np.conj(np.transpose(A, (0, 2, 1))))
    At = torch.from_numpy(A).cuda()
    ev = torch.zeros(b, d, dtype=torch.float64, device="cuda")
    U = torch.zeros(b, d, d, dtype=torch.complex128, device="cuda")
    _lib.eigh_device(0, d, b, At.data_ptr(), ev.data_ptr(), U.data_ptr(), 2)
    torch.cuda.synchronize()
    print("done", d, b, flush=True)
run(96, int(sys.argv[1]))
