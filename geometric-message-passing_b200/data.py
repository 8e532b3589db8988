"""The data path into the layers: host batches -> device, double-buffered.

The reference feeds its models PyG `Batch` objects that a `DataLoader` builds on the host (`experiments/utils/train_utils.py`
`run_experiment`: `for batch in loader: batch = batch.to(device)`), i.e. one synchronous host->device copy per step in front
of the forward pass.  `DevicePrefetcher` keeps that calling convention (an object with `.atoms / .pos / .edge_index /
.batch`) and moves the copy of batch i+1 to a side stream so that it runs under the compute of batch i:
a B200 step of BASELINE config 2 is ~11.6 ms and its 32.7 MB of int64 indices and positions take ~1 ms over PCIe."""
from __future__ import annotations

from typing import Iterable, Iterator

import torch


class Batch:
    """Attribute bag with the reference's field names (`atoms`, `pos`, `edge_index`, `batch`)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def tensors(self):
        return {k: v for k, v in self.__dict__.items() if torch.is_tensor(v)}

    def pin_memory(self) -> "Batch":
        return Batch(**{k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in self.__dict__.items()})

    def to(self, device, non_blocking: bool = False) -> "Batch":
        return Batch(**{k: (v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v) for k, v in self.__dict__.items()})


class DevicePrefetcher:
    """Iterate over host batches (pinned memory for truly asynchronous copies), yielding device batches.  The upload of the
    next batch is issued on a side stream before the current one is handed out, so it overlaps whatever the caller
    launches for the current batch; the consumer stream waits on the upload's event.

    static=False: every batch gets fresh device tensors (registered with the consumer stream so that the caching
    allocator does not recycle them early); shapes may differ from batch to batch.
    static=True: two persistent sets of staging buffers, allocated for the shapes of the first batch and reused in turn (no
    allocator traffic at all; a batch of another shape raises).  A yielded batch stays valid until the next-but-one is
    requested -- which is what a consumer that copies it into its own static inputs (GraphedStep.load) needs."""

    def __init__(self, batches: Iterable, device: torch.device, static: bool = False):
        self.batches = batches
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.static = static
        self._stage = [None, None]
        self._free = [None, None]     # recorded on the consumer stream when it is done with that staging set

    def _upload(self, host, slot):
        if host is None:
            return None
        with torch.cuda.stream(self.stream):
            if self.static:
                if self._stage[slot] is None:
                    self._stage[slot] = host.to(self.device, non_blocking=True)
                else:
                    if self._free[slot] is not None:
                        self.stream.wait_event(self._free[slot])
                    for k, dst in self._stage[slot].tensors().items():
                        src = getattr(host, k)
                        if src.shape != dst.shape or src.dtype != dst.dtype:
                            raise ValueError(f"DevicePrefetcher(static=True): {k} changed shape or dtype")
                        dst.copy_(src, non_blocking=True)
                dev = self._stage[slot]
            else:
                dev = host.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return dev, ev

    def __iter__(self) -> Iterator:
        it = iter(self.batches)
        slot = 0
        nxt = self._upload(next(it, None), slot)
        while nxt is not None:
            dev, ev = nxt
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            if not self.static:
                for t in dev.tensors().values():
                    t.record_stream(cur)
            nxt = self._upload(next(it, None), slot ^ 1)   # in flight while the caller works on `dev`
            yield dev
            if self.static:   # whatever the consumer enqueued for `dev` precedes this event
                self._free[slot] = torch.cuda.Event()
                self._free[slot].record(torch.cuda.current_stream(self.device))
            slot ^= 1
