from .medianThreshold import medianThreshold  # noqa: F401
