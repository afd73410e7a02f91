from .estimators import UNet  # noqa: F401
from .embedders import TimeEmbbeding, LabelEmbedder, SinusoidalPosEmb  # noqa: F401
from .embedders.latent_embedders import VAE, VQVAE  # noqa: F401
from .noise_schedulers import GaussianNoiseScheduler, BasicNoiseScheduler  # noqa: F401
from .pipelines import DiffusionPipeline  # noqa: F401
