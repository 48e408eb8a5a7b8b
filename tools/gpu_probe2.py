"""Eigensolver diagnosis probe."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from muspinsim_b200 import _lib

def run(A, method=1):
    batch, d, _ = A.shape
    At = torch.from_numpy(np.ascontiguousarray(A)).cuda()
    ev = torch.zeros(batch, d, dtype=torch.float64, device="cuda")
    U = torch.zeros(batch, d, d, dtype=torch.complex128, device="cuda")
    torch.cuda.synchronize()
    t0 = time.time()
    _lib.eigh_device(0, d, batch, At.data_ptr(), ev.data_ptr(), U.data_ptr(), method)
    torch.cuda.synchronize()
    return ev.cpu().numpy(), U.cpu().numpy(), time.time() - t0

rng = np.random.default_rng(0)
for d, batch in ((12, 64), (16, 64), (20, 64), (24, 8), (24, 64), (32, 64)):
    A = rng.normal(size=(batch, d, d)) + 1j * rng.normal(size=(batch, d, d))
    A = A + np.conj(np.transpose(A, (0, 2, 1)))
    for rep in range(2):
        ev, U, dt = run(A)
    ref, Uref = np.linalg.eigh(A)
    R = A @ U - U * ev[:, None, :]
    per_mat = np.abs(R).max(axis=(1, 2))
    print("d=%d batch=%d time %.2f ms; per-matrix resid: max %.2e; n_bad %d; first bad %s" % (
        d, batch, dt * 1e3, per_mat.max(), (per_mat > 1e-8).sum(), np.nonzero(per_mat > 1e-8)[0][:10]))
    if (per_mat > 1e-8).any():
        b = int(np.nonzero(per_mat > 1e-8)[0][0])
        percol = np.abs(R[b]).max(axis=0)
        print("   matrix %d: bad columns %s" % (b, np.nonzero(percol > 1e-8)[0]))
        # overlap with reference eigenvectors: which ref vector does each column match
        ov = np.abs(Uref[b].conj().T @ U[b])
        print("   argmax overlap per column:", ov.argmax(axis=0))
        print("   max overlap per column:", np.round(ov.max(axis=0), 3))
        print("   eval err per matrix", np.abs(ev[b] - ref[b]).max())
