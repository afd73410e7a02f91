"""CPU: the C-ABI library loads, exports every symbol include/medfusion_b200.h declares, and the engine's
parameter registry is the reference's state_dict (names + shapes) — no compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

from util import load_golden, make_unet, make_vae

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from medfusion_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "medfusion_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(mf_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.mf_abi_version() == 3
    assert set(_lib.SIGNATURES) == set(declared)


def test_missing_library_fails_loudly(monkeypatch):
    from medfusion_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libmedfusion_b200.so")
    with pytest.raises(_lib.MedfusionLibError):
        _lib.load()


@pytest.mark.parametrize("fixture", ["unet_small.pt", "unet_canonical.pt", "unet_attn_small.pt"])
def test_unet_state_dict_matches_reference(fixture):
    g = load_golden(fixture)
    m = make_unet(g["cfg"])
    got = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    assert sorted(got) == sorted((k, tuple(s)) for k, s in g["keys"])
    # loading a reference-format state_dict works and is strict
    sd = {k: torch.zeros(s) for k, s in g["keys"]}
    m.load_state_dict(sd, strict=True)


def test_vae_accepts_full_reference_state_dict():
    g = load_golden("vae_canonical.pt")
    keep = ("in_channels", "out_channels", "emb_channels", "spatial_dims", "hid_chs", "kernel_sizes", "strides",
            "deep_supervision", "use_attention")
    m = make_vae({k: v for k, v in g["cfg"].items() if k in keep})
    # encoder + decoder: exactly the reference's parameter list, same order, same shapes
    # (the deep-supervision heads `outc_ver.*` only feed training losses and are not built)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == \
        [(k, tuple(s)) for k, s in g["keys"] if not k.startswith("outc_ver.")]
    ref = {k: tuple(s) for k, s in g["keys"]}
    m.load_state_dict({k: torch.zeros(s) for k, s in ref.items()}, strict=True)
    # training-side entries of a real checkpoint (LPIPS, loss modules) are ignored
    m.load_state_dict(dict({k: torch.zeros(s) for k, s in ref.items()},
                           **{"perceiver.net.lin0.model.1.weight": torch.zeros(1, 64, 1, 1)}), strict=True)


def test_scheduler_buffers_match_reference():
    from medfusion_b200.models import GaussianNoiseScheduler
    g = load_golden("sched.pt")
    s = GaussianNoiseScheduler(**g["sched"])
    sd = s.state_dict()
    assert list(sd.keys()) == list(g["buffers"].keys())
    for k, v in g["buffers"].items():
        assert torch.equal(sd[k], v), k


def test_plan_census_and_workspace():
    from medfusion_b200 import _lib
    g = load_golden("unet_canonical.pt")
    m = make_unet(g["cfg"])
    lib = _lib.load()
    nbytes = lib.mf_unet_workspace_bytes(m._h, 64, 32, 32)
    assert 100e6 < nbytes < 8e9
    info = m.plan_info()
    # canonical UNet: 51 convs; only the Cout=8 head is not on the tensor-core path
    assert info["tc_convs"] == 50 and info["simt_convs"] == 1
    assert lib.mf_op_conv_tc_supported(64, 32, 32, 256, 0, 256, 3, 1) == 1
    assert lib.mf_op_conv_tc_supported(64, 8, 8, 1024, 1024, 1024, 3, 1) == 1
    assert lib.mf_op_conv_tc_supported(64, 32, 32, 8, 0, 256, 3, 1) == 0      # Cin = 8 stem
    assert lib.mf_op_conv_tc_supported(64, 32, 32, 96, 0, 256, 3, 1) == 0     # channels must be multiples of 64
    assert lib.mf_op_conv_tc_supported(64, 16, 16, 256, 0, 256, 3, 2) == 1     # stride 2 (output 16x16)
    assert lib.mf_op_conv_tc_supported(64, 16, 16, 256, 256, 256, 3, 2) == 0   # stride 2 has no concat source
    assert lib.mf_op_conv_tc_supported(1, 4, 4, 256, 0, 256, 3, 1) == 0        # fewer than 32 pixels


def test_unsupported_options_raise():
    from medfusion_b200.models import UNet, VAE, DiffusionPipeline, GaussianNoiseScheduler
    with pytest.raises(NotImplementedError):
        UNet(in_ch=8, out_ch=8, spatial_dims=3)
    with pytest.raises(ValueError):
        UNet(in_ch=8, out_ch=8, spatial_dims=2, deep_supervision=False, use_attention="flash")
    with pytest.raises(NotImplementedError):
        VAE(spatial_dims=3)
    with pytest.raises(ValueError):
        DiffusionPipeline(noise_scheduler=GaussianNoiseScheduler, noise_estimator=UNet, estimator_objective="eps",
                          noise_estimator_kwargs=dict(in_ch=8, out_ch=8, spatial_dims=2, deep_supervision=False))


def test_cpu_tensors_are_rejected_not_emulated():
    g = load_golden("unet_small.pt")
    m = make_unet(g["cfg"])
    with pytest.raises(RuntimeError, match="CUDA"):
        m(g["x"], g["t"], g["cond"])


def test_compat_alias_resolves_reference_import_paths():
    import sys
    import medfusion_b200.compat as compat
    saved = {k: v for k, v in sys.modules.items() if k == "medical_diffusion" or k.startswith("medical_diffusion.")}
    for k in saved:
        del sys.modules[k]
    try:
        compat.install()
        from medical_diffusion.models.pipelines import DiffusionPipeline
        from medical_diffusion.models.estimators import UNet
        from medical_diffusion.models.embedders.latent_embedders import VAE
        import medfusion_b200.models as M
        assert DiffusionPipeline is M.DiffusionPipeline and UNet is M.UNet and VAE is M.VAE
    finally:
        for k in [k for k in sys.modules if k == "medical_diffusion" or k.startswith("medical_diffusion.")]:
            del sys.modules[k]
        sys.modules.update(saved)
