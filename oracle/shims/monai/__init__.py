"""Minimal stand-in for the MONAI symbols the reference imports (test infrastructure only).

The real MONAI is not installable in this image (no network).  These shims exist so the
UNMODIFIED reference under /root/reference can be imported in this container to generate the
golden fixtures in tests/golden/ (see oracle/make_golden.py).  They are never imported by the
product package `medfusion_b200`.
"""
