import sys, time, numpy as np
sys.path.insert(0, '.')
from muspinsim_b200 import ExperimentRunner, workloads
for extra in (False,):
    spec = workloads.c3_alc(n_orient=500, n_field=2000, extra_h=extra)
    for opt in (1, 0):
        r = ExperimentRunner(spec, device=0); r.set_option("small24", opt)
        r.run(); t=time.time(); out=r.run(); dt=time.time()-t
        print("d=%d small24=%d  %.1f ms per 1e6 (host-timed)" % (int(np.prod(r.system.dimension)), opt, dt*1e3), float(out.sum()))
