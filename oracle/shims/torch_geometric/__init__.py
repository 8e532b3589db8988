from . import data, loader, nn, utils  # noqa: F401
__version__ = "oracle-shim"
