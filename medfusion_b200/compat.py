"""Make `medical_diffusion.*` imports resolve to this package, so the reference's own scripts
(scripts/sample.py, scripts/helpers/sample_dataset.py, streamlit pages) run unmodified:

    import medfusion_b200.compat; medfusion_b200.compat.install()
    from medical_diffusion.models.pipelines import DiffusionPipeline   # -> medfusion_b200's
"""
from __future__ import annotations

import importlib
import sys
import types

_ALIASES = {
    "medical_diffusion.models": "medfusion_b200.models",
    "medical_diffusion.models.pipelines": "medfusion_b200.models.pipelines",
    "medical_diffusion.models.pipelines.diffusion_pipeline": "medfusion_b200.models.pipelines.diffusion_pipeline",
    "medical_diffusion.models.estimators": "medfusion_b200.models.estimators",
    "medical_diffusion.models.estimators.unet2": "medfusion_b200.models.estimators.unet",
    "medical_diffusion.models.embedders": "medfusion_b200.models.embedders",
    "medical_diffusion.models.embedders.time_embedder": "medfusion_b200.models.embedders.time_embedder",
    "medical_diffusion.models.embedders.cond_embedders": "medfusion_b200.models.embedders.cond_embedders",
    "medical_diffusion.models.embedders.latent_embedders": "medfusion_b200.models.embedders.latent_embedders",
    "medical_diffusion.models.noise_schedulers": "medfusion_b200.models.noise_schedulers",
    "medical_diffusion.models.noise_schedulers.gaussian_scheduler":
        "medfusion_b200.models.noise_schedulers.gaussian_scheduler",
}


def install(force: bool = False):
    if "medical_diffusion" in sys.modules and not force and not getattr(sys.modules["medical_diffusion"],
                                                                      "_medfusion_b200_alias", False):
        raise RuntimeError("a real `medical_diffusion` package is already imported; pass force=True to shadow it")
    root = types.ModuleType("medical_diffusion")
    root.__path__ = []  # mark as package
    root._medfusion_b200_alias = True
    sys.modules["medical_diffusion"] = root
    for alias, target in _ALIASES.items():
        mod = importlib.import_module(target)
        sys.modules[alias] = mod
        parent, _, leaf = alias.rpartition(".")
        setattr(sys.modules[parent], leaf, mod)
    return root
