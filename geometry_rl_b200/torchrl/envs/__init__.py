from .transforms import NDVecNorm, ReshapeTransform  # noqa: F401
