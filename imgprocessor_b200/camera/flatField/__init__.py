from .flatFieldFromCloseDistance import flatFieldFromCloseDistance  # noqa: F401
