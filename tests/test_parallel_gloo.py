"""CPU, world_size 2, gloo: the batch-shard logic (full-batch noise stream + slice, final all-gather)
reproduces the single-process result with the same seed."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from medfusion_b200.parallel import gather_batch, make_sharded_noise_fn, shard_bounds, sharded_sample


class _StubPipe:
    """Stands in for DiffusionPipeline.denoise: a deterministic function of x_T, the noise stream and cond."""

    def denoise(self, x_t, steps=3, condition=None, _noise_fn=None, **kw):
        for i in range(steps):
            x_t = 0.5 * x_t + (i + 1) * _noise_fn(x_t)
            if condition is not None:
                x_t = x_t + condition.view(-1, 1, 1, 1).float()
        return x_t * 2.0


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(1234)
    template = torch.zeros(6, 2, 4, 4)
    cond = torch.arange(6) % 2
    out = sharded_sample(_StubPipe(), template, cond, steps=3)
    lo, hi = shard_bounds(6, world, rank)
    fn = make_sharded_noise_fn(template, lo, hi)
    torch.manual_seed(99)
    g = gather_batch(fn(None), world)
    if rank == 0:
        q.put((out, g))
    dist.destroy_process_group()


def test_sharded_sampling_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out, g = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single-process reference with the same seeds
    torch.manual_seed(1234)
    template = torch.zeros(6, 2, 4, 4)
    cond = torch.arange(6) % 2
    x = torch.randn_like(template)
    ref = _StubPipe().denoise(x, steps=3, condition=cond, _noise_fn=lambda _x: torch.randn_like(template))
    assert torch.equal(out, ref)
    torch.manual_seed(99)
    assert torch.equal(g, torch.randn_like(template))


def test_shard_bounds():
    assert shard_bounds(512, 8, 3) == (192, 256)
    import pytest
    with pytest.raises(ValueError):
        shard_bounds(10, 4, 0)
