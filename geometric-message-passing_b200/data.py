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
    launches for the current batch; the consumer stream waits on the upload's event, and the tensors are registered with
    it so the caching allocator does not recycle them early."""

    def __init__(self, batches: Iterable, device: torch.device):
        self.batches = batches
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)

    def _upload(self, host):
        if host is None:
            return None
        with torch.cuda.stream(self.stream):
            dev = host.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return dev, ev

    def __iter__(self) -> Iterator:
        it = iter(self.batches)
        nxt = self._upload(next(it, None))
        while nxt is not None:
            dev, ev = nxt
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for t in dev.tensors().values():
                t.record_stream(cur)
            nxt = self._upload(next(it, None))   # in flight while the caller works on `dev`
            yield dev
