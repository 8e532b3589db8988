from . import nn, o3, util  # noqa: F401
__version__ = "oracle-shim"
