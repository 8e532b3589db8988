"""Restatement of ``torch_cluster.radius_graph`` in its CUDA canonical order.

ORACLE / TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: the reference never calls
``radius_graph`` (SchNetModel.forward bypasses PyG's interaction graph,
models/schnet.py:66-72); BASELINE.json's north star asks for torch_cluster
semantics, restated here from the published ``radius_cuda.cu`` algorithm:

* one query per node, candidates scanned in ascending index order inside the
  query's own example (``batch``);
* squared distance accumulated in fp32 over d = 0,1,2 with the products
  contracted into FMAs (``fmaf``), strict ``dist < r*r``;
* at most ``max_num_neighbors`` (+1 when ``loop=False``, for the self-match)
  hits are kept, lowest index first; self-loops are then dropped;
* result is dst-major, src ascending: ``edge_index = [src; dst]``.
"""
from __future__ import annotations

import numpy as np


def _fma32(a: np.ndarray, b: np.ndarray, c: np.ndarray) -> np.ndarray:
    """fp32 fused multiply-add, emulated exactly in fp64 (a*b is exact in fp64;
    the single fp64 rounding before the fp32 rounding can double-round only on
    ties that 24x24-bit products plus a 24-bit addend cannot produce when the
    exponents are within 29 bits, which the tests' coordinate ranges guarantee)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def sqdist_f32(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """dist = fma(dz,dz, fma(dy,dy, dx*dx)) in fp32, the order nvcc emits for
    ``for d: dist += (x[d]-y[d])*(x[d]-y[d])`` starting from 0."""
    d0 = (x[..., 0] - y[..., 0]).astype(np.float32)
    d1 = (x[..., 1] - y[..., 1]).astype(np.float32)
    d2 = (x[..., 2] - y[..., 2]).astype(np.float32)
    acc = (d0 * d0).astype(np.float32)
    acc = _fma32(d1, d1, acc)
    acc = _fma32(d2, d2, acc)
    return acc


def radius_graph(pos: np.ndarray, r: float, batch: np.ndarray | None = None, loop: bool = False,
                 max_num_neighbors: int = 32) -> np.ndarray:
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    n = pos.shape[0]
    if batch is None:
        batch = np.zeros(n, dtype=np.int64)
    batch = np.asarray(batch, dtype=np.int64)
    nb = int(batch.max()) + 1 if n else 0
    ptr = np.zeros(nb + 1, dtype=np.int64)
    np.add.at(ptr, batch + 1, 1)
    ptr = np.cumsum(ptr)
    r2 = np.float32(np.float32(r) * np.float32(r))
    cap = max_num_neighbors + (0 if loop else 1)
    src_all, dst_all = [], []
    for b in range(nb):
        lo, hi = int(ptr[b]), int(ptr[b + 1])
        p = pos[lo:hi]
        # rows = queries (dst), cols = candidates (src)
        d = sqdist_f32(p[None, :, :], p[:, None, :])
        hit = d < r2
        rank = np.cumsum(hit, axis=1)
        hit &= rank <= cap
        if not loop:
            np.fill_diagonal(hit, False)
        q, c = np.nonzero(hit)  # row-major: dst ascending, src ascending
        src_all.append(c + lo)
        dst_all.append(q + lo)
    if not src_all:
        return np.zeros((2, 0), dtype=np.int64)
    return np.stack([np.concatenate(src_all), np.concatenate(dst_all)]).astype(np.int64)


def csr_from_coo(index: np.ndarray, num_rows: int):
    """Stable counting sort of edges by ``index``: (rowptr[int32 n+1], perm[int32 E])."""
    index = np.asarray(index, dtype=np.int64)
    perm = np.argsort(index, kind="stable").astype(np.int32)
    counts = np.bincount(index, minlength=num_rows)
    rowptr = np.zeros(num_rows + 1, dtype=np.int32)
    rowptr[1:] = np.cumsum(counts)
    return rowptr, perm
