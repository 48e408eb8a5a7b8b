"""Minimal stand-in for the third-party `ase` package (ase==3.22.1 is pinned by the
reference's requirements.txt but is not installed in this image and not vendored under
/root/reference).  TEST INFRASTRUCTURE ONLY: it exists so that the unmodified reference
can be imported from oracle/_ref to generate golden vectors and CPU baselines.  Only
`ase.quaternions.Quaternion` is restated (from ase's published algorithm); it is pinned
by the reference's own tests/test_config.py:214-286 and tests/test_utils.py:43-65."""
__version__ = "3.22.1-shim"
