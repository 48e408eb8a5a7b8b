"""Where the wall time of ExperimentRunner(spec).run() goes at C5: constructor, handle, first run (workspaces), second run, close."""
import sys, time, types, numpy as np, torch
sys.path.insert(0, ".")
from muspinsim_b200 import ExperimentRunner
import bench
args = types.SimpleNamespace(workload="c5", n_orient=0, nt=0, general=False, scaling="strong")
spec = bench.make_spec(args, 1)
def T(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(3):
    t0 = T(); r = ExperimentRunner(spec, device=0); t1 = T()
    h = r.handle; t2 = T()
    out = r.run(); t3 = T()
    out = r.run(); t4 = T()
    r.handle.close(); t5 = T()
    print("ctor %.1f ms  handle %.1f ms  run1 %.1f ms  run2 %.1f ms  close %.1f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, (t5-t4)*1e3))
