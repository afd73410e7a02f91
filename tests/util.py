"""Shared helpers for the parity tests (tolerance of the north star: rtol=1e-3, atol=1e-5, fp32)."""
import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL, ATOL = 1e-3, 1e-5


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def violations(got, ref, rtol=RTOL, atol=ATOL):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    d = (got - ref).abs()
    return int((d > atol + rtol * ref.abs()).sum()), float(d.max()), float(ref.abs().max())


def assert_close(got, ref, rtol=RTOL, atol=ATOL, what=""):
    assert got.shape == ref.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    assert torch.isfinite(got).all(), f"{what}: non-finite values"
    n, mx, rmax = violations(got, ref, rtol, atol)
    assert n == 0, f"{what}: {n}/{ref.numel()} elements outside rtol={rtol}, atol={atol} (max abs err {mx:.3e}, |ref|max {rmax:.3e})"


def unet_oracle_cfg(cfg):
    groups = dict(cfg.get("norm_name", ("GROUP", {"num_groups": 32}))[1]).get("num_groups", 32)
    return dict(hid_chs=list(cfg["hid_chs"]), strides=list(cfg["strides"]), groups=groups,
                num_res_blocks=cfg.get("num_res_blocks", 2), pos_emb_dim=cfg["time_embedder_kwargs"]["emb_dim"] // 4)


def vae_oracle_cfg(cfg):
    return dict(hid_chs=list(cfg["hid_chs"]), strides=list(cfg["strides"]), groups=8)


def synth_state_dict(keys, seed=0):
    from medfusion_b200.synthetic import synth_tensor
    return {k: synth_tensor(k, shape, seed) for k, shape in keys}


def make_unet(cfg, device=None):
    from medfusion_b200.models import UNet, TimeEmbbeding, LabelEmbedder
    from medfusion_b200.synthetic import fill_
    kw = {k: (dict(v) if isinstance(v, dict) else v) for k, v in cfg.items()}
    m = fill_(UNet(time_embedder=TimeEmbbeding, cond_embedder=LabelEmbedder, **kw))
    return m.to(device) if device is not None else m


def make_vae(cfg, device=None):
    from medfusion_b200.models import VAE
    from medfusion_b200.synthetic import fill_
    m = fill_(VAE(**cfg))
    return m.to(device) if device is not None else m


def to_uint8_hwc(images):
    """Host restatement of scripts/helpers/sample_dataset.py:47-50 for fp32 NCHW images (checker for the fused uint8 head)."""
    x = images.detach().float().clamp(-1, 1)
    x = (x + 1) / 2 * 255
    return x.permute(0, 2, 3, 1).to(torch.uint8)
