"""Quick check of musim_eigh against numpy for a list of sizes: python tools/eigh_check.py 24 64 96 (eigenvalue error, residual)."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from muspinsim_b200 import _lib
def run(d, b=3):
    rng = np.random.default_rng(d)
    A = rng.normal(size=(b, d, d)) + 1j * rng.normal(size=(b, d, d))
    A = np.ascontiguousarray(A + np.conj(np.transpose(A, (0, 2, 1))))
    At = torch.from_numpy(A).cuda()
    ev = torch.zeros(b, d, dtype=torch.float64, device="cuda")
    U = torch.zeros(b, d, d, dtype=torch.complex128, device="cuda")
    _lib.eigh_device(0, d, b, At.data_ptr(), ev.data_ptr(), U.data_ptr(), 2)
    torch.cuda.synchronize()
    ref = np.linalg.eigvalsh(A)
    evn = ev.cpu().numpy(); Un = U.cpu().numpy()
    print(d, "ev err", np.abs(evn - ref).max(), "resid", np.abs(A @ Un - Un * evn[:, None, :]).max())
for d in [int(a) for a in sys.argv[1:]]:
    run(d)
