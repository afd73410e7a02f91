from .time_embedder import TimeEmbbeding, SinusoidalPosEmb  # noqa: F401
from .cond_embedders import LabelEmbedder  # noqa: F401
