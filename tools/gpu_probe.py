"""Development probe run on the GPU box: FP64 peaks, eigensolver accuracy, path parity, phase times."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from muspinsim_b200 import ExperimentRunner, _lib, workloads  # noqa: E402
from oracle import muspin_oracle as mo  # noqa: E402

out = {}
print("device:", torch.cuda.get_device_name(0))
for kind, name in ((0, "dfma"), (1, "dmma")):
    v = _lib.fp64_peak(0, kind)
    out["peak_%s_tflops" % name] = v
    print("peak %s: %.2f TFLOP/s" % (name, v))


def eigh_check(d, batch, method, seed=0, degenerate=False):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(batch, d, d)) + 1j * rng.normal(size=(batch, d, d))
    A = A + np.conj(np.transpose(A, (0, 2, 1)))
    if degenerate:
        # exactly repeated eigenvalues: Q diag(k // 4) Q^H
        for b in range(batch):
            q, _ = np.linalg.qr(A[b])
            lam = (np.arange(d) // 4).astype(float)
            A[b] = (q * lam) @ q.conj().T
    A = np.ascontiguousarray(A)
    At = torch.from_numpy(A).cuda()
    ev = torch.empty(batch, d, dtype=torch.float64, device="cuda")
    U = torch.empty(batch, d, d, dtype=torch.complex128, device="cuda")
    torch.cuda.synchronize()
    t0 = time.time()
    _lib.eigh_device(0, d, batch, At.data_ptr(), ev.data_ptr(), U.data_ptr(), method)
    torch.cuda.synchronize()
    dt = time.time() - t0
    ev, U = ev.cpu().numpy(), U.cpu().numpy()
    ref = np.linalg.eigvalsh(A)
    e_err = np.abs(ev - ref).max() / np.abs(ref).max()
    res = np.abs(A @ U - U * ev[:, None, :]).max() / np.abs(ref).max()
    orth = np.abs(np.conj(np.transpose(U, (0, 2, 1))) @ U - np.eye(d)).max()
    print("eigh d=%3d batch=%5d method=%d deg=%d: eval relerr %.2e resid %.2e orth %.2e  %.1f ms" % (d, batch, method, degenerate, e_err, res, orth, dt * 1e3))
    return e_err, res, orth


for d in (1, 2, 3, 4, 8, 12, 17, 24, 32, 64, 96, 120):
    for method in (1, 2):
        try:
            eigh_check(d, 64, method)
            eigh_check(d, 8, method, degenerate=True)
        except Exception as exc:
            print("eigh d=%d method=%d failed: %s" % (d, method, exc))

specs = [
    workloads.c1_hfine(),
    workloads.c2_hfine_powder(n_orient=16, nt=100, n_h=2),
    workloads.c2_hfine_powder(n_orient=8, nt=64, n_h=1, temperature=0.3),
    workloads.c2_hfine_powder(n_orient=8, nt=1000, n_h=3),
    workloads.c3_alc(n_orient=4, n_field=8, extra_h=False),
    workloads.c3_alc(n_orient=4, n_field=8, extra_h=True),
    workloads.c5_large(n_orient=4, nt=1000),
    workloads.c5_large(n_orient=4, nt=200, temperature=1.0),
    workloads.c4_fmuf_dissipation(n_orient=4, nt=100),
]
for spec in specs:
    try:
        want = mo.run_spec(spec, evolve_fn=mo.evolve_vectorised)
        for opts in ({}, {"polar": 1}, {"eigh": 1}):
            r = ExperimentRunner(spec, device=0)
            for k, v in opts.items():
                r.set_option(k, v)
            got = r.run()
            print("parity %-24s %-12s max|gpu-oracle| = %.3e" % (spec["name"], opts, np.abs(got - want).max()))
    except Exception as exc:
        print("parity %-24s FAILED: %s" % (spec["name"], exc))

# phase timing on the bench-size workloads
for spec in (workloads.c2_hfine_powder(n_orient=20000), workloads.c5_large(n_orient=4000)):
    r = ExperimentRunner(spec, device=0)
    r.set_option("profile", 1)
    r.run()
    t0 = time.time()
    r.run()
    dt = time.time() - t0
    ph = {k: r.handle.phase_ms(k) for k in ("eigh", "rotate", "rho0", "polar", "integral")}
    print("phases %s n=%d: %.1f ms total, %s" % (spec["name"], r.config.n_cfg, dt * 1e3, ph))
    out["phases_" + spec["name"]] = ph
    r.set_option("profile", 0)
    torch.cuda.synchronize()
    t0 = time.time()
    r.run()
    dt = time.time() - t0
    print("   unprofiled: %.1f ms -> %.0f eval/s" % (dt * 1e3, r.config.n_cfg / dt))

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
