"""Per-phase cycle counters of the half-storage tridiagonalisation kernel ("Why K1 is still ...").
Build an instrumented library first (the counters are compiled out of the product):
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DHS_TIMING \
       -shared -o /tmp/libmusim_timing.so muspinsim_b200/csrc/musim.cu
  HSLIB=/tmp/libmusim_timing.so python tools/k1_cycle_breakdown.py 148     # 148 matrices = one CTA per SM
Block 0 prints, per warp, the average cycles between the barriers of a Householder step."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from muspinsim_b200 import _lib
import os; _lib.LIB_PATH = os.environ.get("HSLIB", "/tmp/libmusim_timing.so")
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
