"""``torch_scatter.scatter`` / ``scatter_sum`` / ``scatter_mean`` on the CSR kernels.

Same signature and semantics as the wheel for the cases the reference uses (2-D ``src``, reduction
along the node dimension; call sites models/layers/egnn_layer.py:77,79,147, models/layers/tfn_layer.py:87),
but deterministic: the index is counting-sorted once and reduced segment by segment, no atomics.
"""
from __future__ import annotations

from typing import Optional

import torch

from ._lib import GmpError, call, ptr
from .graph import CSR, build_csr


class _SegmentReduce(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, csr: CSR, mean: bool):
        src = src.contiguous()
        out = torch.empty(csr.n, src.shape[1], dtype=src.dtype, device=src.device)
        call("gmp_segment_reduce_f32", ptr(csr.rowptr), csr.perm_ptr, ptr(src), ptr(out), csr.n, src.shape[1], int(mean))
        ctx.csr, ctx.mean, ctx.E = csr, mean, src.shape[0]
        return out

    @staticmethod
    def backward(ctx, g):
        csr: CSR = ctx.csr
        g = g.contiguous()
        if ctx.mean:
            deg = (csr.rowptr[1:] - csr.rowptr[:-1]).clamp(min=1).to(g.dtype).unsqueeze(1)
            g = g / deg
        rows = csr.row_ids()
        gs = torch.empty(ctx.E, g.shape[1], dtype=g.dtype, device=g.device)
        call("gmp_gather_rows_f32", ptr(rows), ptr(g), ptr(gs), ctx.E, g.shape[1])
        if csr.perm is not None:  # sorted order -> caller's edge order
            out = torch.empty_like(gs)
            out[csr.perm.long()] = gs
            gs = out
        return gs, None, None


def segment_reduce(src: torch.Tensor, csr: CSR, reduce: str = "sum") -> torch.Tensor:
    assert reduce in ("sum", "add", "mean")
    return _SegmentReduce.apply(src, csr, reduce == "mean")


def scatter(src: torch.Tensor, index: torch.Tensor, dim: int = -1, out: Optional[torch.Tensor] = None,
            dim_size: Optional[int] = None, reduce: str = "sum") -> torch.Tensor:
    if out is not None:
        raise GmpError("gmp_b200.scatter: `out=` is not supported")
    if reduce not in ("sum", "add", "mean"):
        raise ValueError(f"gmp_b200.scatter: reduce={reduce!r} is outside the hot path")
    if src.dim() == 1:
        return scatter(src.unsqueeze(1), index, 0, None, dim_size, reduce).squeeze(1)
    if src.dim() != 2 or dim not in (0, -2) or index.dim() != 1 or src.dtype != torch.float32:
        raise GmpError("gmp_b200.scatter handles float32 [E,F] reduced along dim 0 with a 1-D index")
    n = dim_size if dim_size is not None else (int(index.max().item()) + 1 if index.numel() else 0)
    if n == 0:
        return src.new_zeros(0, src.shape[1])
    csr = build_csr(index, index, n)
    return segment_reduce(src, csr, reduce)


def scatter_sum(src, index, dim: int = -1, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, "sum")


def scatter_mean(src, index, dim: int = -1, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, "mean")
