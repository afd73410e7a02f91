from .gaussian_scheduler import BasicNoiseScheduler, GaussianNoiseScheduler  # noqa: F401
