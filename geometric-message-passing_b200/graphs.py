"""Whole-step CUDA graphs ("use CUDA graphs to kill launch latency").

One SchNet training step at BASELINE config 2 is ~270 kernel launches (81 of them ours) for ~11 ms of device work, and the
Python + ATen + ctypes launch path needs about as long to issue them: the step is host-bound.  `GraphedStep` captures
`out = model(batch); loss(out).backward()` once into a `torch.cuda.CUDAGraph` and replays it, so a step costs one launch
on the host.  Everything on the path is capture-safe: the C-ABI entry points only enqueue kernels / memsets on the current
stream, the CSR build asks the device nothing (graph.build_csr), pooling takes `batch.num_graphs`.

Two modes:
* resident batch: the CSR views are built (and cached) before capture; a replay recomputes forward + backward.
* `rebuild_graph=True`: the batch tensors are static device buffers that `load(host_batch)` overwrites; the capture then
  contains the CSR sort as well, so a replay does everything a fresh batch of the same shapes needs.

Outputs of earlier eager steps must not be alive when a GraphedStep is built (their autograd graph pins the parameters'
gradient accumulators to the default stream, which invalidates the capture)."""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import _lib
from .data import Batch


class GraphedStep:
    def __init__(self, model: torch.nn.Module, batch, loss: Optional[Callable] = None, warmup: int = 3,
                 rebuild_graph: bool = False):
        assert all(t.is_cuda for t in _tensors(batch).values()), "GraphedStep needs a device-resident (static) batch"
        self.model, self.batch = model, batch
        self.rebuild_graph = rebuild_graph
        self.loss = loss or (lambda out: out.sum())
        self.params = [p for p in model.parameters() if p.requires_grad]
        # (the parameters' AccumulateGrad nodes may have been created by earlier eager steps on the default stream; autograd
        # warns when the side / capture stream meets them.  The accumulation itself is captured -- tests/test_gpu_schnet.py
        # checks the replayed gradients against eager ones for changing inputs.)
        warn = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if warn is not None:
            warn(False)
        try:
            self._capture(model, batch, warmup, rebuild_graph)
        finally:
            if warn is not None:
                warn(True)
        self.grads = [p.grad for p in self.params]

    def _capture(self, model, batch, warmup, rebuild_graph):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # warm-up off the capture stream: lazy inits, cuBLAS workspaces, autotuning
            for _ in range(max(1, warmup)):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for p in self.params:
            p.grad = None                      # the captured backward then allocates the gradients in the graph's pool
        if rebuild_graph:                      # new version counter -> the CSR cache misses -> the sort is captured too
            batch.edge_index.add_(0)
            # load() may later put ANY edge list into this buffer: the captured CSR build must not rely on the capture-time
            # tensor having been dst-sorted (a radius_graph output says so through this attribute), so the permutation of the
            # by_dst view is always built inside the graph
            batch.edge_index._gmp_dst_sorted = False
        k0, c0 = _lib.kernel_launches(), _lib.launches
        try:
            self.graph = torch.cuda.CUDAGraph(keep_graph=True)   # keeps the cudaGraph_t: kernel_nodes() counts its nodes
        except TypeError:
            self.graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(self.graph):
                self.out = model(batch)
                self.loss(self.out).backward()
        except RuntimeError as ex:
            raise RuntimeError(
                "GraphedStep: the capture was invalidated.  The usual cause is a live autograd graph from an earlier eager step "
                "(e.g. its output tensor is still referenced): it pins the parameters' AccumulateGrad nodes to the default stream, "
                "which cannot take part in a stream capture.  Drop those references (del out) before building a GraphedStep.") from ex
        self.kernels_per_replay = _lib.kernel_launches() - k0   # our kernels inside one replay (counted at capture)
        self.calls_per_replay = _lib.launches - c0

    def _eager(self):
        for p in self.params:
            p.grad = None
        out = self.model(self.batch)
        self.loss(out).backward()
        return out

    def load(self, host_batch) -> None:
        """Overwrite the static batch with a host batch of the same shapes (pinned memory makes the copies asynchronous).

        rebuild_graph=True: every field is replaced, edge_index included -- the replay re-sorts it.
        rebuild_graph=False: the replay runs against the CSR views built before capture, so the edge list is part of the
        captured state: node fields (atoms, pos, batch, ...) are replaced, `edge_index` is left alone, and a host batch
        that carries a different edge list raises (the comparison reads one flag back; pass a batch without `edge_index`
        to skip it)."""
        for k, dst in _tensors(self.batch).items():
            src = getattr(host_batch, k, None)
            if k == "edge_index" and not self.rebuild_graph:
                if src is not None and src is not dst:
                    same = src.shape == dst.shape and bool(torch.equal(src.to(dst.device), dst))
                    if not same:
                        raise ValueError("GraphedStep.load: this step was captured with rebuild_graph=False, i.e. for one fixed edge "
                                         "list; build it with rebuild_graph=True to load batches with other edges")
                continue
            assert src is not None, f"GraphedStep.load: the host batch has no field {k!r}"
            assert src.shape == dst.shape and src.dtype == dst.dtype, f"GraphedStep.load: {k} changed shape or dtype"
            dst.copy_(src, non_blocking=True)

    def kernel_nodes(self):
        """Number of kernel nodes of the captured graph (ours + ATen + cuBLAS), counted from the graph itself with
        cudaGraphGetNodes / cudaGraphNodeGetType; None when the handle or the CUDA bindings are unavailable."""
        try:
            from cuda.bindings import runtime as rt
            raw = self.graph.raw_cuda_graph()
            err, _, num = rt.cudaGraphGetNodes(raw, 0)
            if int(err) != 0:
                return None
            err, nodes, num = rt.cudaGraphGetNodes(raw, num)
            kernels = 0
            for nd in nodes[:num]:
                err, ty = rt.cudaGraphNodeGetType(nd)
                if int(err) == 0 and int(ty) == int(rt.cudaGraphNodeType.cudaGraphNodeTypeKernel):
                    kernels += 1
            return kernels
        except Exception:  # noqa: BLE001
            return None

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.out


def _tensors(batch):
    return batch.tensors() if isinstance(batch, Batch) else {k: v for k, v in vars(batch).items() if torch.is_tensor(v)}
