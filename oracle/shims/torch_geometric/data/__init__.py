from oracle.thirdparty.pyg import Batch, Data  # noqa: F401
