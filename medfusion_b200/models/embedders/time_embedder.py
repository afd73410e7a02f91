"""Time-embedding *specifications* (reference: medical_diffusion/models/embedders/time_embedder.py).

In the reference these are nn.Modules that compute; here they only describe the embedding (widths,
frequency table) — the arithmetic runs inside the UNet engine's embedding kernels
(csrc/mf_kernels.cu: linear_small).  Class names and constructor arguments match the reference so
`UNet(time_embedder=TimeEmbbeding, time_embedder_kwargs={'emb_dim': 1024})` reads the same.
"""
from __future__ import annotations

import math

import torch


class SinusoidalPosEmb:
    """time_embedder.py:7-28.  Only the default cat(sin, cos) ordering is supported."""

    def __init__(self, emb_dim=16, downscale_freq_shift=1, max_period=10000, flip_sin_to_cos=False):
        if flip_sin_to_cos:
            raise NotImplementedError("flip_sin_to_cos=True is not supported by the B200 embedding kernel")
        if emb_dim % 2 != 0 or emb_dim < 4:
            raise ValueError("SinusoidalPosEmb emb_dim must be even (cat(sin, cos) halves)")
        self.emb_dim = emb_dim
        self.downscale_freq_shift = downscale_freq_shift
        self.max_period = max_period
        self.flip_sin_to_cos = flip_sin_to_cos

    def frequencies(self) -> torch.Tensor:
        """exp(-ln(max_period)/(half-shift) * k), evaluated with torch CPU fp32 ops exactly like
        time_embedder.py:17-18, so the device sin/cos arguments agree with the reference."""
        half = self.emb_dim // 2
        step = math.log(self.max_period) / (half - self.downscale_freq_shift)
        return torch.exp(-step * torch.arange(half))


class TimeEmbbeding:
    """time_embedder.py:52-75: sinusoidal -> Linear -> Swish -> Linear."""

    def __init__(self, emb_dim=64, pos_embedder=SinusoidalPosEmb, pos_embedder_kwargs=None, act_name=("SWISH", {})):
        if str(act_name[0] if isinstance(act_name, (tuple, list)) else act_name).lower() != "swish":
            raise NotImplementedError("only the Swish activation is implemented")
        kwargs = dict(pos_embedder_kwargs or {})
        self.emb_dim = emb_dim
        self.pos_emb_dim = kwargs.get("emb_dim", emb_dim // 4)
        kwargs["emb_dim"] = self.pos_emb_dim
        self.pos_embedder = pos_embedder(**kwargs)
