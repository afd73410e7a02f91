"""Checkpoint ingestion: Lightning `.ckpt` files written by the reference's training scripts -> this runtime
(reference: medical_diffusion/models/model_base.py:49-85, pipelines/diffusion_pipeline.py:56-60).

A reference checkpoint is a `torch.save`d dict {'state_dict', 'hyper_parameters', 'callbacks', 'optimizer_states', …}
whose `hyper_parameters` pickle *class objects* by qualified name (`medical_diffusion.models.estimators.unet2.UNet`,
`torch.optim.AdamW`, `torch.nn.L1Loss`, sometimes `lpips.LPIPS` / `pytorch_lightning.*`).  None of Lightning, MONAI or
the reference package is needed here: `load_checkpoint` unpickles with a resolver that
  * maps `medical_diffusion.*` classes to this package's class of the same name,
  * resolves an allow-list normally (torch.*, collections, numpy, pathlib and plain builtin containers),
  * and replaces everything else (training-side losses, Lightning callbacks, and any other importable global such as
    os.system or builtins.eval) by an inert placeholder class,
so the file loads without importing training-side code and without resolving arbitrary callables.  This narrows, but does
not remove, the usual pickle caveat: load checkpoints you trust.  `CheckpointMixin` gives the modules the
reference's loader methods with the reference's names and argument meaning.
"""
from __future__ import annotations

import importlib
import json
import pickle
from pathlib import Path

import torch

_PLACEHOLDERS: dict = {}


def _placeholder(module, name):
    key = (module, name)
    if key not in _PLACEHOLDERS:
        def _init(self, *a, **k):
            self.args, self.kwargs = a, k

        def _setstate(self, state):
            self.__dict__["state"] = state

        _PLACEHOLDERS[key] = type(name.rsplit(".", 1)[-1], (), {
            "__module__": module, "__init__": _init, "__setstate__": _setstate, "_medfusion_b200_placeholder": True,
            "__reduce_ex__": lambda self, p: (object.__new__, (type(self),)),
        })
    return _PLACEHOLDERS[key]


def is_placeholder(obj) -> bool:
    return bool(getattr(obj, "_medfusion_b200_placeholder", False))


def _resolve_reference_class(module, name):
    """`medical_diffusion.<path>.<Name>` -> medfusion_b200's class `<Name>` (same public name), else a placeholder."""
    from . import models
    leaf = name.rsplit(".", 1)[-1]
    target = "medfusion_b200" + module[len("medical_diffusion"):]
    for cand in (target, target.rsplit(".", 1)[0], "medfusion_b200.models"):
        try:
            mod = importlib.import_module(cand)
        except ImportError:
            continue
        if hasattr(mod, leaf):
            return getattr(mod, leaf)
    if hasattr(models, leaf):
        return getattr(models, leaf)
    return _placeholder(module, name)


# Globals a Lightning checkpoint of the reference legitimately contains: tensor rebuild helpers, containers, optimizer /
# loss / module CLASSES stored in `hyper_parameters` (only constructed if this package's constructors ask for them).
# Anything else (os.system, builtins.eval, subprocess.Popen, ...) resolves to an inert placeholder instead of being
# imported, so unpickling cannot call into arbitrary code.
_ALLOWED_MODULE_PREFIXES = ("torch", "collections", "numpy", "pathlib", "argparse", "datetime", "functools",
                            "medfusion_b200")   # this package's own classes (checkpoints re-saved through it)
_ALLOWED_BUILTINS = {"set", "frozenset", "slice", "tuple", "list", "dict", "int", "float", "bool", "str", "bytes",
                     "bytearray", "complex", "range", "object", "type"}
_FORBIDDEN = {("torch", "load"), ("torch", "save"), ("torch.serialization", "load"), ("functools", "partial"),
              ("torch.utils.cpp_extension", "load"), ("torch.hub", "load")}


def _allowed(module, name):
    if (module, name) in _FORBIDDEN:
        return False
    if module in ("builtins", "__builtin__"):
        return name in _ALLOWED_BUILTINS
    root = module.split(".", 1)[0]
    return root in _ALLOWED_MODULE_PREFIXES


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "medical_diffusion" or module.startswith("medical_diffusion."):
            return _resolve_reference_class(module, name)
        if not _allowed(module, name):
            return _placeholder(module, name)
        try:
            return super().find_class(module, name)
        except (ImportError, AttributeError):
            return _placeholder(module, name)


class _PickleModule:
    """The `pickle_module` protocol torch.load expects (Unpickler + load + the pickle constants it touches)."""
    __name__ = "medfusion_b200.checkpoint._PickleModule"
    Unpickler = _Unpickler
    Pickler = pickle.Pickler
    HIGHEST_PROTOCOL = pickle.HIGHEST_PROTOCOL
    dump = staticmethod(pickle.dump)
    dumps = staticmethod(pickle.dumps)

    @staticmethod
    def load(f, **kw):
        return _Unpickler(f, **kw).load()

    @staticmethod
    def loads(b, **kw):
        import io
        return _Unpickler(io.BytesIO(b), **kw).load()


def load_checkpoint(path, map_location=None):
    """torch.load of a reference checkpoint without Lightning / MONAI / the reference package installed."""
    return torch.load(str(path), map_location=map_location or "cpu", weights_only=False, pickle_module=_PickleModule)


def _ctor_kwargs(cls, hp):
    code = cls.__init__.__code__
    accepted = code.co_varnames[1:code.co_argcount + code.co_kwonlyargcount]
    kw = {}
    for k, v in hp.items():
        if k not in accepted:
            continue
        if is_placeholder(v) or (isinstance(v, type) and is_placeholder(v)):
            continue                      # training-side object (loss, perceiver): constructor default (None) is used
        kw[k] = v
    return kw


class CheckpointMixin:
    """Loader methods of the reference's `VeryBasicModel` (model_base.py:49-85) for nn.Modules of this package."""

    @classmethod
    def load_from_checkpoint(cls, path, map_location=None, strict=True, **overrides):
        """Lightning's classmethod: ctor(**hyper_parameters, **overrides) then load_state_dict(state_dict)."""
        ckpt = load_checkpoint(path, map_location)
        if not isinstance(ckpt, dict):
            raise ValueError(f"{path}: not a checkpoint dict")
        hp = dict(ckpt.get("hyper_parameters", {}) or {})
        hp.update(overrides)
        model = cls(**_ctor_kwargs(cls, hp))
        sd = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
        model.load_state_dict(sd, strict=strict)
        return model

    @classmethod
    def save_best_checkpoint(cls, path_checkpoint_dir, best_model_path):
        with open(Path(path_checkpoint_dir) / "best_checkpoint.json", "w") as f:        # model_base.py:49-52
            json.dump({"best_model_epoch": Path(best_model_path).name}, f)

    @classmethod
    def _get_best_checkpoint_path(cls, path_checkpoint_dir, version=0, **kwargs):
        path_version = "lightning_logs/version_" + str(version)                            # model_base.py:55-59
        with open(Path(path_checkpoint_dir) / path_version / "best_checkpoint.json", "r") as f:
            rel = Path(json.load(f)["best_model_epoch"])
        return Path(path_checkpoint_dir) / rel

    @classmethod
    def load_best_checkpoint(cls, path_checkpoint_dir, version=0, **kwargs):
        return cls.load_from_checkpoint(cls._get_best_checkpoint_path(path_checkpoint_dir, version), **kwargs)

    def load_pretrained(self, checkpoint_path, map_location=None, **kwargs):
        checkpoint_path = Path(checkpoint_path)                                            # model_base.py:66-75
        if checkpoint_path.is_dir():
            checkpoint_path = self._get_best_checkpoint_path(checkpoint_path, **kwargs)
        ckpt = load_checkpoint(checkpoint_path, map_location)
        return self.load_weights(ckpt["state_dict"], **kwargs)

    def load_weights(self, pretrained_weights, strict=True, **kwargs):
        flt = kwargs.get("filter", lambda key: key in pretrained_weights)                  # model_base.py:77-83
        init_weights = self.state_dict()
        init_weights.update({k: v for k, v in pretrained_weights.items() if flt(k)})
        self.load_state_dict(init_weights, strict=strict)
        return self
