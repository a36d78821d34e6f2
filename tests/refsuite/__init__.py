"""Runs the reference's own test-suite against the drop-in package (CPU only; see fakedev.py)."""
