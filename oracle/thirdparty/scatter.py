"""Restatement of ``torch_scatter.scatter`` / ``scatter_sum``.

ORACLE / TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED against the wheel.
Reference call sites: models/layers/egnn_layer.py:77,79,147;
models/layers/tfn_layer.py:87; models/mace_modules/blocks.py:261 (dead code).
"""
from __future__ import annotations

from typing import Optional

import torch


def _broadcast(index: torch.Tensor, src: torch.Tensor, dim: int) -> torch.Tensor:
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    for _ in range(index.dim(), src.dim()):
        index = index.unsqueeze(-1)
    return index.expand(src.size())


def scatter_sum(src, index, dim: int = -1, out: Optional[torch.Tensor] = None, dim_size: Optional[int] = None):
    index = _broadcast(index, src, dim)
    if out is None:
        size = list(src.size())
        if dim_size is not None:
            size[dim] = dim_size
        elif index.numel() == 0:
            size[dim] = 0
        else:
            size[dim] = int(index.max()) + 1
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


def scatter_mean(src, index, dim: int = -1, out=None, dim_size=None):
    out = scatter_sum(src, index, dim, out, dim_size)
    dim_size = out.size(dim)
    index_dim = dim
    if index_dim < 0:
        index_dim = index_dim + src.dim()
    if index.dim() <= index_dim:
        index_dim = index.dim() - 1
    ones = torch.ones(index.size(), dtype=src.dtype, device=src.device)
    count = scatter_sum(ones, index, index_dim, None, dim_size)
    count[count < 1] = 1
    count = _broadcast(count, out, dim)
    if out.is_floating_point():
        out.true_divide_(count)
    else:
        out.div_(count, rounding_mode="floor")
    return out


def scatter(src, index, dim: int = -1, out=None, dim_size=None, reduce: str = "sum"):
    if reduce in ("sum", "add"):
        return scatter_sum(src, index, dim, out, dim_size)
    if reduce == "mean":
        return scatter_mean(src, index, dim, out, dim_size)
    if reduce == "max":
        # torch_scatter.scatter_max()[0] [upstream-recalled]: rows that receive nothing are 0 (the kernel fills them after the
        # reduction: `out.masked_fill_(arg_out == src.size(dim), 0)`); the gradient goes to the arg-max entry
        assert out is None
        index = _broadcast(index, src, dim)
        size = list(src.size())
        size[dim] = dim_size if dim_size is not None else (int(index.max()) + 1 if index.numel() else 0)
        return torch.zeros(size, dtype=src.dtype, device=src.device).scatter_reduce(dim, index, src, "amax", include_self=False)
    raise ValueError(f"oracle scatter: reduce={reduce!r} is outside the hot path")
