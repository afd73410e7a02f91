"""Parameter containers that reproduce the reference's state_dict key hierarchy.

The C engine is the single source of truth for which parameters exist (names + shapes, enumerated
through mf_*_param_name/shape).  `build_param_tree` turns that flat list into nested nn.Module
containers so `state_dict()` / `load_state_dict()` use exactly the reference's dotted keys
(e.g. `in_blocks.3.0.block_seq.0.conv_res.weight`), without re-implementing the reference's
module classes.
"""
from __future__ import annotations

import math
from typing import Iterable, Tuple

import torch
import torch.nn as nn


class ParamGroup(nn.Module):
    """A bare container; children are further ParamGroups, leaves are nn.Parameters."""

    def extra_repr(self) -> str:
        return ", ".join(f"{k}{tuple(v.shape)}" for k, v in self._parameters.items())


def _init_leaf(name: str, shape: Tuple[int, ...], zero: bool) -> torch.Tensor:
    t = torch.empty(shape, dtype=torch.float32)
    leaf = name.rsplit(".", 1)[-1]
    parent = name.rsplit(".", 2)[-2] if name.count(".") >= 1 else ""
    if zero:
        return t.zero_()
    if parent == "norm":
        return t.fill_(1.0) if leaf == "weight" else t.zero_()
    if parent == "embedding":
        return t.normal_(0.0, 1.0)
    if len(shape) >= 2:  # conv / linear weight: torch default (kaiming_uniform, a=sqrt(5))
        fan_in = int(torch.tensor(shape[1:]).prod())
        bound = 1.0 / math.sqrt(fan_in)
        return t.uniform_(-bound, bound)
    return t.uniform_(-0.05, 0.05)  # biases (fan-in dependent in torch; magnitude is irrelevant here)


def build_param_tree(root: nn.Module, entries: Iterable[Tuple[str, Tuple[int, ...]]], zero_init=()):
    """Attach nested ParamGroups/Parameters named by `entries` to `root`."""
    for name, shape in entries:
        parts = name.split(".")
        mod = root
        for part in parts[:-1]:
            child = mod._modules.get(part)
            if child is None:
                child = ParamGroup()
                mod.add_module(part, child)
            mod = child
        zero = any(name.startswith(z) or (z.startswith("*") and z[1:] in name) for z in zero_init)
        mod.register_parameter(parts[-1], nn.Parameter(_init_leaf(name, tuple(shape), zero), requires_grad=False))


def param_signature(module: nn.Module):
    """Cheap change detector: (data_ptr, in-place version) of every parameter."""
    return tuple((p.data_ptr(), p._version) for p in module.parameters())
