"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref) on the golden
cases through its public API (.in text -> MuSpinInput -> ExperimentRunner.run()).
Run in the build container only:  python -m oracle.make_golden
Each file stores the spec (JSON) and the reference's results array."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import build_ref, ref_driver  # noqa: E402
from oracle.golden_cases import cases  # noqa: E402


def spec_to_json(spec):
    def conv(o):
        if isinstance(o, np.ndarray):
            return o.tolist()
        if isinstance(o, (np.floating, np.integer)):
            return o.item()
        raise TypeError(type(o))

    return json.dumps(spec, default=conv)


def spec_from_json(text):
    return json.loads(text)


def main():
    build_ref.build()
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for name, spec in cases().items():
        res = ref_driver.run_reference(spec)
        np.savez_compressed(os.path.join(out, name + ".npz"), spec=spec_to_json(spec), results=np.asarray(res))
        print("%-28s shape %-12s |max| %.6f" % (name, np.shape(res), np.abs(res).max()))


if __name__ == "__main__":
    main()
