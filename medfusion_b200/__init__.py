"""medfusion_b200 — B200-native (sm_100a) implementation of Medfusion's DDPM sampling hot path.

Python is the host layer only: the reference-compatible module API lives in `medfusion_b200.models`
and marshals torch tensors to the C-ABI library `csrc/libmedfusion_b200.so` (see include/medfusion_b200.h).
"""
__version__ = "0.1.0"
