"""Minimal stand-in for Soprano==0.8.13 (absent from this image).  TEST INFRASTRUCTURE ONLY.
Only the two entry points the reference calls are restated: ZCW orientations (parity
UNPINNED: the reference's tests only check len>=N and <3cos^2-1>~0) and the isotope table."""
