"""Batch-sharded sampling across the GPUs of one node (not present in the reference).

One process per GPU (torchrun), full weight replica per rank, rank r owns samples [r*B/G, (r+1)*B/G).
There is NO per-step communication: samples are independent (GroupNorm is per-sample).  To stay
seed-compatible with the single-GPU / reference trajectory, every rank draws the FULL-batch noise
stream in the reference's order and keeps only its slice.  One all-gather of the decoded images ends
the call (NCCL over NVLink; gloo in the CPU unit tests of the slicing logic).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(batch: int, world: int, rank: int):
    if batch % world != 0:
        raise ValueError(f"batch {batch} is not divisible by world size {world}")
    per = batch // world
    return rank * per, (rank + 1) * per


def make_sharded_noise_fn(template: torch.Tensor, lo: int, hi: int):
    """Draw like `torch.randn_like(full_batch)` (same generator consumption as the reference), keep [lo:hi)."""

    def noise_fn(_x_local):
        return torch.randn_like(template)[lo:hi].contiguous()

    noise_fn.graph_safe = True   # pure device-side torch ops: may be captured into the per-timestep CUDA graph
    # stable identity for the CUDA-graph cache of denoise(): a fresh closure per call must not force a re-capture
    noise_fn.cache_key = ("sharded", lo, hi, tuple(template.shape), str(template.device))
    return noise_fn


def gather_batch(x_local: torch.Tensor, world: int) -> torch.Tensor:
    out = torch.empty((x_local.shape[0] * world, *x_local.shape[1:]), device=x_local.device, dtype=x_local.dtype)
    if dist.get_backend() == "gloo":  # CPU tests
        parts = [torch.empty_like(x_local) for _ in range(world)]
        dist.all_gather(parts, x_local.contiguous())
        return torch.cat(parts, dim=0)
    dist.all_gather_into_tensor(out, x_local.contiguous())
    return out


def sharded_sample(pipe, template: torch.Tensor, condition=None, **kwargs):
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("sample(shard=True) needs an initialised torch.distributed process group")
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_bounds(template.shape[0], world, rank)
    noise_fn = make_sharded_noise_fn(template, lo, hi)
    x_T = noise_fn(None)
    cond = None if condition is None else condition[lo:hi].contiguous()
    if kwargs.get("un_cond", None) is not None:
        kwargs = dict(kwargs, un_cond=kwargs["un_cond"][lo:hi].contiguous())
    x_local = pipe.denoise(x_T, condition=cond, _noise_fn=noise_fn, **kwargs)
    return gather_batch(x_local, world)
