"""medfusion_b200 — B200-native (sm_100a) implementation of Medfusion's DDPM sampling hot path.

Python is the host layer only: the reference-compatible module API lives in `medfusion_b200.models`
and marshals torch tensors to the C-ABI library `csrc/libmedfusion_b200.so` (see include/medfusion_b200.h).
"""
__version__ = "0.1.0"


def saturation_count(reset: bool = False, device=None) -> int:
    """Number of values the kernels had to clamp to the fp16 range of the split planes (+-65504, or non-finite) on
    `device` (default: current) since the last reset — mf_saturation_count.  Synchronises the current stream.  The
    reference computes in fp32 range: a non-zero count means parity with it was lost (DiffusionPipeline.denoise raises)."""
    import ctypes

    import torch

    from . import _lib
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    n = ctypes.c_ulonglong()
    with torch.cuda.device(dev):
        _lib.check(_lib.load().mf_saturation_count(ctypes.byref(n), 1 if reset else 0,
                                                   torch.cuda.current_stream(dev).cuda_stream), "mf_saturation_count")
    return int(n.value)
