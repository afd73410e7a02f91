from .diffusion_pipeline import DiffusionPipeline  # noqa: F401
