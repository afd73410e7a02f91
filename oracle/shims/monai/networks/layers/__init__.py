from .factories import Conv, Pool  # noqa: F401
