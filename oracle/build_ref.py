"""Install the UNMODIFIED reference (/root/reference, read-only) into oracle/_ref/ (git-ignored).

TEST INFRASTRUCTURE.  The reference is a Python package with one Cython and one pybind11
extension; pip builds both from a scratch copy under /tmp because the source tree is
read-only.  No reference sources are copied into the repository history.  Third-party
packages the reference imports but this image lacks (ase, soprano, qutip) are stood in for by
oracle/shims/ (our own restatements; see each file's header).

Outcome recorded in DESIGN.md: installs cleanly with `--no-deps`; with the shims 152 of the
reference's 158 runnable tests pass, the 6 failures are all Celio's method (np.product was
removed in numpy 2; qutip.Qobj is stubbed) which is out of scope.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")


def build(force=False):
    if os.path.isdir(os.path.join(DST, "muspinsim")) and not force:
        return DST
    if not os.path.isdir(SRC):
        raise RuntimeError("reference sources not present (expected only in the build container)")
    tmp = "/tmp/musim_ref_src"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(SRC, tmp)
    shutil.rmtree(DST, ignore_errors=True)
    env = dict(os.environ, CC="/usr/bin/gcc", CXX="/usr/bin/g++")
    subprocess.check_call(
        [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
         "--find-links", "/opt/wheelhouse", "--target", DST, tmp],
        env=env,
    )
    return DST


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
