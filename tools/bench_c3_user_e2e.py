"""User-level end-to-end time of the ALC scan (C3): ExperimentRunner(spec).run() including the
configuration-table step, with the table expanded on the device vs on the host."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from muspinsim_b200 import ExperimentRunner, workloads

n_orient = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
spec = workloads.c3_alc(n_orient=n_orient, n_field=2000)
ref = None
for expand in (True, False, True):
    t0 = time.time()
    r = ExperimentRunner(spec, device=0)
    r.device_expand = expand
    out = r.run()
    dt = time.time() - t0
    print("device_expand=%s  n_cfg=%d  construct+run %.3f s" % (expand, r.config.n_cfg, dt))
    if ref is None:
        ref = out
    else:
        print("   max |diff| vs first run: %.2e" % np.max(np.abs(out - ref)))
