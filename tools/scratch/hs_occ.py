import sys, numpy as np, torch
sys.path.insert(0, ".")
from muspinsim_b200 import _lib
def run(d, b):
    rng = np.random.default_rng(d)
    A = rng.normal(size=(b, d, d)) + 1j * rng.normal(size=(b, d, d))
    A = np.ascontiguousarray(A + np.conj(np.transpose(A, (0, 2, 1))))
    At = torch.from_numpy(A).cuda()
    ev = torch.zeros(b, d, dtype=torch.float64, device="cuda")
    U = torch.zeros(b, d, d, dtype=torch.complex128, device="cuda")
    for _ in range(2):
        _lib.eigh_device(0, d, b, At.data_ptr(), ev.data_ptr(), U.data_ptr(), 2)
    torch.cuda.synchronize()
for b in [148, 296, 592, 1184]:
    run(96, b)
