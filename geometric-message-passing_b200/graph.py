"""Graph construction and CSR views on the GPU (host side of csrc/graph.cu).

``radius_graph`` mirrors ``torch_cluster.radius_graph`` (argument names, canonical CUDA edge order,
``max_num_neighbors`` truncation); the reference never calls it (models/schnet.py:66-72 bypasses
PyG's interaction graph) but BASELINE.json's north star requires it bit-exact.

``CSR`` is the destination-sorted view every fused layer aggregates over: edges stably sorted by
their aggregation index, so the segmented reduction is deterministic and needs no atomics
(replaces torch_scatter's atomicAdd at models/layers/egnn_layer.py:77,79 and tfn_layer.py:87).
"""
from __future__ import annotations

import weakref
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import call, ptr


def _i32(n, device):
    return torch.empty(n, dtype=torch.int32, device=device)


def exclusive_scan(counts: torch.Tensor) -> torch.Tensor:
    """int32[n] -> int64[n+1] exclusive prefix sum on the GPU."""
    n = counts.numel()
    out = torch.empty(n + 1, dtype=torch.int64, device=counts.device)
    ws = torch.empty((n + 1023) // 1024 + 1, dtype=torch.int64, device=counts.device)
    call("gmp_exclusive_scan_i32", ptr(counts), n, ptr(out), ptr(ws))
    return out


def radius_graph(x: torch.Tensor, r: float, batch: Optional[torch.Tensor] = None, loop: bool = False,
                 max_num_neighbors: int = 32, flow: str = "source_to_target",
                 method: str = "auto") -> torch.Tensor:
    """edge_index [2, E] int64 = [src; dst], dst-major, src ascending (torch_cluster CUDA order).

    method: "brute" scans each example in index order (torch_cluster's own algorithm);
            "cells" uses a uniform cell list (single example only; same output bit for bit);
            "auto" picks cells for one example with more than 4096 nodes."""
    assert flow == "source_to_target"
    assert x.dim() == 2 and x.shape[1] == 3 and x.dtype == torch.float32, "pos must be float32 [N,3]"
    x = x.contiguous()
    n, dev = x.shape[0], x.device
    if batch is None:
        gptr = torch.tensor([0, n], dtype=torch.int64, device=dev)
        ngraphs = 1
    else:
        assert batch.numel() == n
        ngraphs = int(batch.max().item()) + 1 if n else 1
        counts = torch.bincount(batch, minlength=ngraphs)
        gptr = torch.zeros(ngraphs + 1, dtype=torch.int64, device=dev)
        gptr[1:] = torch.cumsum(counts, 0)
    if method == "auto":
        method = "cells" if (ngraphs == 1 and n > 4096) else "brute"
    deg = _i32(n, dev)
    if method == "brute":
        call("gmp_radius_graph_count", ptr(x), ptr(gptr), ngraphs, n, float(r), max_num_neighbors, int(loop), ptr(deg))
        rowptr = exclusive_scan(deg)
        E = int(rowptr[-1].item())
        ei = torch.empty(2, E, dtype=torch.int64, device=dev)
        if E:
            call("gmp_radius_graph_fill", ptr(x), ptr(gptr), ngraphs, n, float(r), max_num_neighbors, int(loop),
                 ptr(rowptr), ptr(ei[0]), ptr(ei[1]))
        _tag_dst_sorted(ei)   # dst-major, ascending: Graph skips the permutation of the by_dst view
        return ei
    assert ngraphs == 1, "the cell-list path handles a single example"
    import ctypes as C
    lo = x.min(dim=0).values.cpu()
    hi = x.max(dim=0).values.cpu()
    cell = float(r) * 1.0001  # > r so that fp rounding of the cell coordinate can never hide a neighbour
    dims = [max(1, int((float(hi[k]) - float(lo[k])) / cell) + 1) for k in range(3)]
    origin = (C.c_float * 3)(*[float(v) for v in lo])
    cdims = (C.c_int32 * 3)(*dims)
    ncells = dims[0] * dims[1] * dims[2]
    cell_of, cell_nodes = _i32(n, dev), _i32(n, dev)
    cell_start, cursor = _i32(ncells + 1, dev), _i32(ncells, dev)
    call("gmp_cells_build", ptr(x), n, cell, origin, cdims, ptr(cell_of), ptr(cell_start), ptr(cell_nodes), ptr(cursor))
    call("gmp_radius_cells_count", ptr(x), n, float(r), cell, origin, cdims, ptr(cell_start), ptr(cell_nodes),
         max_num_neighbors, int(loop), ptr(deg))
    rowptr = exclusive_scan(deg)
    E = int(rowptr[-1].item())
    ei = torch.empty(2, E, dtype=torch.int64, device=dev)
    if E:
        call("gmp_radius_cells_fill", ptr(x), n, float(r), cell, origin, cdims, ptr(cell_start), ptr(cell_nodes),
             max_num_neighbors, int(loop), ptr(rowptr), ptr(ei[0]), ptr(ei[1]))
    _tag_dst_sorted(ei)
    return ei


def _tag_dst_sorted(ei: torch.Tensor) -> None:
    """Mark a radius_graph output as grouped by destination in ascending order.  The tag is bound to the tensor's
    version counter: an in-place edit of the edge list invalidates it."""
    ei._gmp_dst_sorted = True
    ei._gmp_dst_sorted_version = ei._version


def _is_tagged_dst_sorted(ei: torch.Tensor) -> bool:
    return bool(getattr(ei, "_gmp_dst_sorted", False)) and getattr(ei, "_gmp_dst_sorted_version", None) == ei._version


@dataclass
class CSR:
    """Edges stably sorted by `index` (the aggregation side); `col` is the other endpoint."""
    rowptr: torch.Tensor          # int32 [n+1]
    col: torch.Tensor             # int32 [E]   gather-side node of sorted edge k
    perm: Optional[torch.Tensor]  # int32 [E]   position of sorted edge k in the caller's edge order (None = identity)
    n: int
    E: int

    @property
    def perm_ptr(self):
        return ptr(self.perm)

    def inv_perm(self) -> Optional[torch.Tensor]:
        """int32[E]: sorted position of the edge with caller's id i (None when the CSR order is the caller's order); cached."""
        if self.perm is None:
            return None
        if getattr(self, "_inv_perm", None) is None:
            inv = torch.empty_like(self.perm)
            inv[self.perm.long()] = torch.arange(self.E, device=self.perm.device, dtype=torch.int32)
            self._inv_perm = inv
        return self._inv_perm

    def row_ids(self) -> torch.Tensor:
        """int32[E]: aggregation row of each sorted edge (cached)."""
        if getattr(self, "_row_ids", None) is None:
            deg = (self.rowptr[1:] - self.rowptr[:-1]).long()
            self._row_ids = torch.repeat_interleave(   # output_size: no read-back of deg.sum(), capturable in a CUDA graph
                torch.arange(self.n, device=self.rowptr.device, dtype=torch.int32), deg, output_size=self.E).contiguous()
        return self._row_ids


def build_csr(index: torch.Tensor, other: torch.Tensor, n: int, assume_sorted: Optional[bool] = None) -> CSR:
    """Stable counting sort of the edges by `index` (int64[E]); `other` is the opposite endpoint.

    assume_sorted=True: the caller guarantees `index` is non-decreasing (our own radius_graph output is dst-major):
    no permutation is stored.  None: no guarantee and no questions asked of the device -- the permutation is always built
    (the identity when the input happens to be sorted), because reading a sortedness flag back would stall the host once
    per graph and step.  False: check on the device and read the flag back (the exact perm-is-None-iff-sorted contract)."""
    assert index.dtype == torch.int64 and other.dtype == torch.int64
    index, other = index.contiguous(), other.contiguous()
    E, dev = index.numel(), index.device
    counts = _i32(max(n, 1), dev)
    call("gmp_csr_count", ptr(index), E, n, ptr(counts))
    rowptr = exclusive_scan(counts[:n]).to(torch.int32)
    col = _i32(E, dev)
    if assume_sorted is False:
        flag = _i32(1, dev)
        call("gmp_index_is_sorted", ptr(index), E, ptr(flag))
        assume_sorted = int(flag.item()) == 1
    if assume_sorted:
        perm = None
    else:
        perm, tmp, cursor = _i32(E, dev), _i32(E, dev), _i32(max(n, 1), dev)
        call("gmp_csr_fill", ptr(index), E, n, ptr(rowptr), ptr(cursor), ptr(tmp), ptr(perm))
    call("gmp_gather_i64_to_i32", ptr(other), ptr(perm), E, ptr(col))
    return CSR(rowptr, col, perm, n, E)


class Graph:
    """Both sorted views of one edge_index, built lazily and cached per edge_index tensor.

    by_dst: rows = edge_index[1] (EGNN / SchNet aggregate here, PyG flow source_to_target)
    by_src: rows = edge_index[0] (TFN / MACE aggregate here, models/layers/tfn_layer.py:83-87;
            also the transposed pass of the EGNN / SchNet backward)."""

    def __init__(self, edge_index: torch.Tensor, n: int):
        assert edge_index.dim() == 2 and edge_index.shape[0] == 2 and edge_index.dtype == torch.int64
        self._orig = edge_index                    # keeps the caller's storage alive while this Graph is cached (see get_graph)
        self.edge_index = edge_index.contiguous()
        self.n = n
        self.E = edge_index.shape[1]
        # radius_graph (ours, like torch_cluster's) emits edges grouped by destination in ascending order and says so
        self._dst_sorted = _is_tagged_dst_sorted(edge_index)
        self._by_dst: Optional[CSR] = None
        self._by_src: Optional[CSR] = None

    @property
    def by_dst(self) -> CSR:
        if self._by_dst is None:
            self._by_dst = build_csr(self.edge_index[1], self.edge_index[0], self.n, True if self._dst_sorted else None)
        return self._by_dst

    @property
    def by_src(self) -> CSR:
        if self._by_src is None:
            self._by_src = build_csr(self.edge_index[0], self.edge_index[1], self.n)
        return self._by_src


_GRAPH_CACHE: Dict[Tuple, Graph] = {}
_GRAPH_CACHE_SIZE = 4  # a training step touches one or two graphs; a small cache lets the allocator recycle their buffers


def get_graph(edge_index: torch.Tensor, n: int) -> Graph:
    """Graph for this edge_index, cached on (storage pointer, shape, strides, version, n) so that the layers of a
    model, which all receive the same tensor, sort it once.  The cached Graph holds a reference to the caller's tensor
    itself (not only to a contiguous copy of it), so its storage cannot be freed and recycled for a different edge list
    while the entry exists."""
    key = (edge_index.data_ptr(), tuple(edge_index.shape), tuple(edge_index.stride()), edge_index._version, n, str(edge_index.device))
    g = _GRAPH_CACHE.get(key)
    if g is None:
        while len(_GRAPH_CACHE) >= _GRAPH_CACHE_SIZE:
            _GRAPH_CACHE.pop(next(iter(_GRAPH_CACHE)))  # oldest entry
        g = Graph(edge_index, n)
        _GRAPH_CACHE[key] = g
    return g
