"""Bulk image generator — the production caller of the sampling hot path.

Restates /root/reference/scripts/helpers/sample_dataset.py:24-52: for a label, `n_samples` images are drawn in chunks
of `sample_batch` (200 there) with `pipeline.sample(len(chunk), img_size, guidance_scale=cfg, condition=…, un_cond=…,
steps=…)` under one `torch.manual_seed(0)`, converted `clip(-1,1) -> (x+1)/2*255 -> HWC -> uint8` (:47-50) and written
as `fake_{counter}.png` (:52).

What differs from the reference script is only where the time goes:
  * the uint8/HWC conversion is the epilogue of the VAE output head (`VAE.decode_uint8`, `mf_vae_decode_u8`), so a chunk
    leaves the GPU as B·H·W·C bytes instead of 4·B·C·H·W;
  * the device→host copy lands in one of two pinned buffers on a side stream and PNG encoding runs on a thread pool,
    both overlapping the sampling of the next chunk (the reference encodes serially between `sample()` calls).
File names, chunking, seeding order and pixel values are the reference's.
"""
from __future__ import annotations

import threading
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import torch


def chunks(lst, n):
    """Successive n-sized chunks of lst (sample_dataset.py:10-13)."""
    for i in range(0, len(lst), n):
        yield lst[i:i + n]


def _save_png(arr, path):
    from PIL import Image  # sample_dataset.py:7,51-52
    if arr.shape[-1] == 1:
        arr = arr[..., 0]
    Image.fromarray(arr).convert("RGB").save(path)


class _HostRing:
    """Two pinned staging buffers; a buffer is reused only after its PNG jobs have finished."""

    def __init__(self, shape):
        self.bufs = [torch.empty(shape, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.events = [torch.cuda.Event() for _ in range(2)]
        self.pending = [[], []]
        self.i = 0

    def acquire(self):
        k = self.i
        self.i ^= 1
        for f in self.pending[k]:
            f.result()
        self.pending[k] = []
        return k


def generate_dataset(pipeline, n_samples, path_out=None, label=None, un_cond_label="flip", steps=150, guidance_scale=1,
                     sample_batch=200, img_size=(8, 32, 32), seed=0, workers=8, sink=None, **sample_kwargs):
    """Generate `n_samples` images for one label; returns the number written.

    label=None -> unconditional (condition=None, un_cond=None).  un_cond_label: "flip" = `1-label` (sample_dataset.py:40),
    None, or an int.  `sink(counter, uint8_hwc_numpy)` replaces the PNG writer when given (tests, in-memory consumers);
    otherwise files go to `path_out/fake_{counter}.png`.
    """
    dev = pipeline.device
    if dev.type != "cuda":
        raise RuntimeError("medfusion_b200: generate_dataset needs the pipeline on a CUDA device (no CPU fallback)")
    if sink is None:
        if path_out is None:
            raise ValueError("path_out or sink required")
        path_out = Path(path_out)
        path_out.mkdir(parents=True, exist_ok=True)

        def sink(counter, arr):  # noqa: F811
            _save_png(arr, path_out / f"fake_{counter}.png")

    torch.manual_seed(seed)                                             # sample_dataset.py:36
    copy_stream = torch.cuda.Stream(device=dev)
    ring = None
    counter = 0
    lock = threading.Lock()
    errors = []

    def job(c, arr):
        try:
            sink(c, arr)
        except Exception as exc:  # surfaced after the loop; a failed write must not be silent
            with lock:
                errors.append(exc)

    # `flusher` waits for a chunk's copy event and fans its images out to the PNG `pool`; the sampling loop itself
    # only blocks when both staging buffers are still being encoded.
    with ThreadPoolExecutor(max_workers=max(1, workers)) as pool, ThreadPoolExecutor(max_workers=1) as flusher:
        for chunk in chunks(list(range(n_samples)), sample_batch):
            n = len(chunk)
            condition = torch.full((n,), label, device=dev, dtype=torch.long) if label is not None else None
            if label is None or un_cond_label is None:
                un_cond = None
            else:
                u = (1 - label) if un_cond_label == "flip" else int(un_cond_label)
                un_cond = torch.full((n,), u, device=dev, dtype=torch.long)
            img = pipeline.sample_uint8(n, img_size, guidance_scale=guidance_scale, condition=condition,
                                        un_cond=un_cond, steps=steps, **sample_kwargs)      # [n, H, W, C] uint8, device
            if ring is None:
                ring = _HostRing((sample_batch, *img.shape[1:]))
            k = ring.acquire()
            copy_stream.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(copy_stream):
                ring.bufs[k][:n].copy_(img, non_blocking=True)
                ring.events[k].record(copy_stream)
            img.record_stream(copy_stream)

            def flush(host=ring.bufs[k], ev=ring.events[k], n=n, base=counter):
                ev.synchronize()
                arr = host[:n].numpy()
                for f in [pool.submit(job, base + j, arr[j]) for j in range(n)]:
                    f.result()

            ring.pending[k] = [flusher.submit(flush)]
            counter += n
        if ring is not None:
            for k in (0, 1):
                for f in ring.pending[k]:
                    f.result()
    if errors:
        raise errors[0]
    return counter
