"""Deterministic synthetic parameters (there are no checkpoints offline).

`fill_(module_or_state_dict)` overwrites every tensor from a per-key seeded CPU generator, so the
reference modules (golden generation), the CPU oracle and the B200 engine can be given bit-identical
weights from nothing but the state_dict keys and shapes.  It also removes the reference's zero
initialisations (conv_blocks.py:336, unet2.py:213, latent_embedders.py:743) that would otherwise make
random-init outputs identically zero (SURVEY.md finding 4).
"""
from __future__ import annotations

import math
import zlib

import torch


def synth_tensor(key: str, shape, seed: int = 0) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed((zlib.crc32(key.encode()) + 7919 * seed) % (2 ** 31))
    shape = tuple(int(s) for s in shape)
    leaf = key.rsplit(".", 1)[-1]
    parent = key.rsplit(".", 2)[-2] if key.count(".") >= 1 else ""
    r = torch.randn(shape, generator=g, dtype=torch.float32)
    if parent in ("norm", "norm_x") and len(shape) == 1:
        return 1.0 + 0.1 * r if leaf == "weight" else 0.1 * r
    if parent == "embedding":
        return 0.5 * r
    if len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return r * (1.0 / math.sqrt(fan_in))
    return 0.05 * r


@torch.no_grad()
def fill_(target, seed: int = 0, skip_prefixes=()):
    """target: nn.Module (parameters are overwritten in place) or a dict name -> tensor."""
    items = target.state_dict().items() if hasattr(target, "state_dict") else target.items()
    for key, val in items:
        if not torch.is_floating_point(val) or key.startswith(tuple(skip_prefixes)):
            continue
        val.copy_(synth_tensor(key, val.shape, seed).to(val.device))
    return target
