"""state_dict key lists of the reference modules for a config, reconstructed WITHOUT the reference: the
canonical fixtures store their own `keys`; for other configs we enumerate through the C engine registry
(which tests/test_abi.py checks against the stored reference key lists)."""
from util import make_unet, make_vae


def unet_keys(cfg):
    return [(k, tuple(v.shape)) for k, v in make_unet(cfg).state_dict().items()]


def vae_keys(cfg):
    keep = ("in_channels", "out_channels", "emb_channels", "spatial_dims", "hid_chs", "kernel_sizes", "strides",
            "deep_supervision", "use_attention")
    return [(k, tuple(v.shape)) for k, v in make_vae({k: v for k, v in cfg.items() if k in keep}).state_dict().items()]
