"""TEST-ONLY stand-in for torch-geometric 2.5.2 (see ../README.md).  Not product code."""
from . import typing  # noqa: F401
from . import data  # noqa: F401
from . import nn  # noqa: F401
from . import transforms  # noqa: F401

__version__ = "2.5.2-shim"
