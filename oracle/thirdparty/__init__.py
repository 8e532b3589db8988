"""Pure-PyTorch / numpy restatements of the third-party wheels the reference
imports (torch_scatter, torch_geometric, torch_cluster, e3nn, opt_einsum).

ORACLE / TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED against the real wheels
(none is installable in this image; see oracle/__init__.py)."""
from . import cluster, e3nn_nn, o3, pyg, scatter  # noqa: F401
from .einsum import contract  # noqa: F401
