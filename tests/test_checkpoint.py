"""Checkpoint ingestion: Lightning-format files of the reference -> this runtime, with neither
Lightning nor the reference package importable.  The pickled `hyper_parameters` in tests/golden/ckpt_skeleton.pt were
written by the unmodified reference classes (oracle/make_golden.py:ckpt_fixture)."""
import sys

import pytest
import torch

from util import load_golden  # noqa: F401  (path setup)
from medfusion_b200.checkpoint import is_placeholder, load_checkpoint
from medfusion_b200.synthetic import synth_tensor
import os

SKELETON = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ckpt_skeleton.pt")


def _state_dict(keys, sched_buffers):
    sd = {}
    for k, shape, dtype in keys:
        if k.startswith("noise_scheduler."):
            sd[k] = sched_buffers[k[len("noise_scheduler."):]].clone()
        else:
            rel = k.split(".", 1)[1] if k.startswith(("noise_estimator.", "latent_embedder.")) else k
            sd[k] = synth_tensor(rel, shape)
    return sd


def write_checkpoints(tmp_path):
    """Materialise the two .ckpt files of a training run from the skeleton + synthetic per-key weights."""
    sk = load_checkpoint(SKELETON)
    vae_ckpt = {k: v for k, v in sk["vae"].items() if k != "keys"}
    vae_ckpt["state_dict"] = _state_dict(sk["vae"]["keys"], sk["sched_buffers"])
    pipe_ckpt = {k: v for k, v in sk["pipeline"].items() if k != "keys"}
    pipe_ckpt["state_dict"] = _state_dict(sk["pipeline"]["keys"], sk["sched_buffers"])
    # training-side entries a real run also carries
    pipe_ckpt["state_dict"]["latent_embedder.perceiver.net.lin0.model.1.weight"] = torch.zeros(1, 64, 1, 1)
    sk["pipeline"]["keys"].append(("latent_embedder.perceiver.net.lin0.model.1.weight", (1, 64, 1, 1), "torch.float32"))
    vae_ckpt["hyper_parameters"] = {k: v for k, v in vae_ckpt["hyper_parameters"].items() if not is_placeholder(v)}
    torch.save(vae_ckpt, tmp_path / "last_vae.ckpt")
    torch.save(pipe_ckpt, tmp_path / "last.ckpt")
    return tmp_path / "last.ckpt", tmp_path / "last_vae.ckpt", sk


def test_reference_pickles_resolve_without_the_reference_installed():
    assert "medical_diffusion" not in sys.modules and "pytorch_lightning" not in sys.modules
    sk = load_checkpoint(SKELETON)
    from medfusion_b200.models import (DiffusionPipeline, GaussianNoiseScheduler, LabelEmbedder, TimeEmbbeding, UNet,  # noqa
                                       VAE)
    hp = sk["pipeline"]["hyper_parameters"]
    assert hp["noise_estimator"] is UNet and hp["noise_scheduler"] is GaussianNoiseScheduler and hp["latent_embedder"] is VAE
    assert hp["noise_estimator_kwargs"]["time_embedder"] is TimeEmbbeding
    assert hp["noise_estimator_kwargs"]["cond_embedder"] is LabelEmbedder
    assert hp["optimizer"] is torch.optim.AdamW and hp["loss"] is torch.nn.L1Loss          # importable -> real classes
    assert is_placeholder(sk["vae"]["hyper_parameters"]["perceiver"])                       # lpips.LPIPS: inert stand-in
    assert "medical_diffusion" not in sys.modules and "lpips" not in sys.modules


def test_pipeline_load_from_checkpoint_restores_every_hot_path_tensor(tmp_path):
    from medfusion_b200.models import DiffusionPipeline, VAE
    pipe_path, vae_path, sk = write_checkpoints(tmp_path)
    # the path stored in the checkpoint points into the training machine's run directory
    with pytest.raises(FileNotFoundError):
        DiffusionPipeline.load_from_checkpoint(pipe_path)
    pipe = DiffusionPipeline.load_from_checkpoint(pipe_path, latent_embedder_checkpoint=str(vae_path))
    assert isinstance(pipe.latent_embedder, VAE) and pipe.estimator_objective == "x_T" and pipe.clip_x0 is False
    want = _state_dict(sk["pipeline"]["keys"], sk["sched_buffers"])
    got = pipe.state_dict()
    skipped = [k for k in want if k not in got]
    assert skipped == ["latent_embedder.perceiver.net.lin0.model.1.weight"]     # training-side only
    for k, v in got.items():
        assert torch.equal(v.cpu(), want[k]), k
    assert all(not p.requires_grad for p in pipe.latent_embedder.parameters())             # diffusion_pipeline.py:58-59
    # strict loading still rejects a damaged hot-path entry
    bad = dict(want)
    bad.pop("noise_estimator.outc.conv.conv.weight")
    with pytest.raises(RuntimeError):
        pipe.load_state_dict(bad)


def test_best_checkpoint_and_load_weights_helpers(tmp_path):
    from medfusion_b200.models import VAE
    _, vae_path, sk = write_checkpoints(tmp_path)
    run = tmp_path / "run"
    (run / "lightning_logs" / "version_0").mkdir(parents=True)
    (run / "epoch=3.ckpt").write_bytes(vae_path.read_bytes())
    VAE.save_best_checkpoint(run / "lightning_logs" / "version_0", run / "epoch=3.ckpt")   # model_base.py:49-52
    vae = VAE.load_best_checkpoint(run)
    ref = VAE.load_from_checkpoint(vae_path)
    for (k, a), (_, b) in zip(vae.state_dict().items(), ref.state_dict().items()):
        assert torch.equal(a, b), k
    # load_weights with a filter (model_base.py:77-83): only the selected keys change
    fresh = VAE(**{k: v for k, v in sk["vae"]["hyper_parameters"].items() if not is_placeholder(v)})
    before = {k: v.clone() for k, v in fresh.state_dict().items()}
    fresh.load_weights(ref.state_dict(), filter=lambda key: key.startswith("outc."))
    for k, v in fresh.state_dict().items():
        assert torch.equal(v, ref.state_dict()[k] if k.startswith("outc.") else before[k]), k
    fresh.load_pretrained(run)
    assert torch.equal(fresh.state_dict()["inc_dec.block_seq.0.basic_block.conv.weight"],
                       ref.state_dict()["inc_dec.block_seq.0.basic_block.conv.weight"])


def test_legacy_serialization_and_unknown_classes_in_containers(tmp_path):
    """Old (non-zip) torch.save files go through pickle_module.load; classes that cannot be imported — wherever they sit
    in the hyper-parameter tree — become inert placeholders instead of import errors."""
    import pickle
    import types

    mod = types.ModuleType("pytorch_lightning_fake_callbacks")

    class ModelCheckpoint:                      # stands in for a training-side class absent at sampling time
        def __init__(self, monitor):
            self.monitor = monitor
    ModelCheckpoint.__module__ = mod.__name__
    ModelCheckpoint.__qualname__ = "ModelCheckpoint"
    mod.ModelCheckpoint = ModelCheckpoint
    sys.modules[mod.__name__] = mod
    try:
        obj = {"state_dict": {"w": torch.arange(4.0)},
               "hyper_parameters": {"callbacks": [ModelCheckpoint("val/loss")], "loss": torch.nn.L1Loss, "lr": 1e-4}}
        torch.save(obj, tmp_path / "legacy.ckpt", _use_new_zipfile_serialization=False)
        torch.save(obj, tmp_path / "zip.ckpt")
    finally:
        del sys.modules[mod.__name__]
    for name in ("legacy.ckpt", "zip.ckpt"):
        with pytest.raises((ModuleNotFoundError, AttributeError, pickle.UnpicklingError)):
            torch.load(tmp_path / name, weights_only=False)                       # the stock loader cannot resolve it
        ck = load_checkpoint(tmp_path / name)
        assert torch.equal(ck["state_dict"]["w"], torch.arange(4.0))
        hp = ck["hyper_parameters"]
        assert is_placeholder(hp["callbacks"][0]) and hp["loss"] is torch.nn.L1Loss and hp["lr"] == 1e-4


def test_unpickler_does_not_resolve_arbitrary_importable_globals(tmp_path):
    """only an allow-list (torch, numpy, collections, plain containers, this package) is resolved; any other
    importable global inside a checkpoint becomes an inert placeholder instead of being imported and called."""
    import pickle

    class Evil:
        def __reduce__(self):
            import os
            return (os.system, ("echo pwned > %s" % (tmp_path / "pwned"),))

    p = tmp_path / "evil.ckpt"
    with open(p, "wb") as fh:
        pickle.dump({"state_dict": {}, "hyper_parameters": {"x": Evil()}}, fh)
    from medfusion_b200.checkpoint import _PickleModule
    with open(p, "rb") as fh:
        obj = _PickleModule.load(fh)
    assert not (tmp_path / "pwned").exists()
    assert is_placeholder(obj["hyper_parameters"]["x"])
