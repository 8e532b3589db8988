"""The data path into the layers: `Data` / `Batch.from_data_list` / `DataLoader` with the PyG field names
the reference's datasets and training loops use (experiments/utils/create_graphs.py:78-79, train_utils.py:24-35),
`to_undirected` / `coalesce` (bit-exact with torch_geometric.utils, on the GPU when the edge list lives there), and the
double-buffered host -> device upload.

The reference feeds its models PyG `Batch` objects that a `DataLoader` builds on the host (`experiments/utils/train_utils.py`
`run_experiment`: `for batch in loader: batch = batch.to(device)`), i.e. one synchronous host->device copy per step in front
of the forward pass.  `DevicePrefetcher` keeps that calling convention (an object with `.atoms / .pos / .edge_index /
.batch`) and moves the copy of batch i+1 to a side stream so that it runs under the compute of batch i:
a B200 step of BASELINE config 2 is ~11.6 ms and its 32.7 MB of int64 indices and positions take ~1 ms over PCIe."""
from __future__ import annotations

from typing import Iterable, Iterator, Optional

import torch


class Data:
    """One graph: an attribute bag with the reference's field names (`atoms`, `pos`, `edge_index`, `y`, ...), as
    `torch_geometric.data.Data(atoms=..., edge_index=..., pos=..., y=...)` at experiments/utils/create_graphs.py:78."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    @property
    def num_nodes(self) -> int:
        for k in ("atoms", "pos", "x"):
            v = self.__dict__.get(k)
            if torch.is_tensor(v):
                return v.shape[0]
        raise AttributeError("num_nodes: the graph has none of atoms / pos / x")

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    def tensors(self):
        return {k: v for k, v in self.__dict__.items() if torch.is_tensor(v)}

    def to(self, device, non_blocking: bool = False):
        return type(self)(**{k: (v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v) for k, v in self.__dict__.items()})

    def pin_memory(self):
        return type(self)(**{k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in self.__dict__.items()})


class Batch(Data):
    """Several graphs as one disconnected graph, with the reference's field names (`atoms`, `pos`, `edge_index`, `batch`,
    `num_graphs`, `y`): what `torch_geometric.loader.DataLoader` hands to `model(batch)` (train_utils.py:24-35)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    @staticmethod
    def from_data_list(data_list) -> "Batch":
        """PyG collate: node-level tensors are concatenated along dim 0, `edge_index` along dim 1 with the
        cumulative node count added, 0-d tensors become one entry each, `batch[i]` = graph of node i, `num_graphs`.
        Works on whatever device the graphs live on (device-resident datasets are batched without touching the host:
        the node offsets are Python ints known from the shapes)."""
        assert len(data_list) > 0
        keys = data_list[0].keys()
        counts = [d.num_nodes for d in data_list]
        offs = [0]
        for c in counts[:-1]:
            offs.append(offs[-1] + c)
        out = {}
        for k in keys:
            vals = [getattr(d, k) for d in data_list]
            if k == "edge_index":
                widths = [v.shape[1] for v in vals]
                ei = torch.cat(vals, dim=1)
                shift = torch.repeat_interleave(torch.tensor(offs, dtype=ei.dtype, device=ei.device),
                                                torch.tensor(widths, device=ei.device), output_size=sum(widths))
                out[k] = ei + shift
            elif torch.is_tensor(vals[0]):
                out[k] = torch.cat([v if v.dim() > 0 else v.reshape(1) for v in vals], dim=0)
            else:
                out[k] = vals
        ref = next(v for v in data_list[0].tensors().values())
        out["batch"] = torch.repeat_interleave(torch.arange(len(data_list), device=ref.device),
                                               torch.tensor(counts, device=ref.device), output_size=sum(counts))
        out["num_graphs"] = len(data_list)
        return Batch(**out)



class DataLoader:
    """`torch_geometric.loader.DataLoader(dataset, batch_size, shuffle)` for a list of `Data`: yields `Batch` objects.
    shuffle draws a fresh permutation per epoch from torch's global CPU generator, like torch.utils.data.RandomSampler."""

    def __init__(self, dataset, batch_size: int = 1, shuffle: bool = False, drop_last: bool = False):
        self.dataset, self.batch_size, self.shuffle, self.drop_last = list(dataset), batch_size, shuffle, drop_last

    def __len__(self) -> int:
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = len(self.dataset)
        order = torch.randperm(n).tolist() if self.shuffle else list(range(n))
        for i in range(0, n, self.batch_size):
            idx = order[i:i + self.batch_size]
            if self.drop_last and len(idx) < self.batch_size:
                return
            yield Batch.from_data_list([self.dataset[j] for j in idx])


def coalesce(edge_index: torch.Tensor, num_nodes: Optional[int] = None) -> torch.Tensor:
    """Sort the edges lexicographically by (row, col) and drop duplicates (torch_geometric.utils.coalesce without edge
    attributes).  CUDA edge lists: two stable counting-sort passes (by col, then by row: an LSD radix sort over the two
    node ids, csrc/graph.cu) + mark / scan / compact -- integer work only, bit-exact with the host result.
    Host edge lists (the reference builds its datasets on the CPU): the same ordering through torch.unique."""
    assert edge_index.dim() == 2 and edge_index.shape[0] == 2 and edge_index.dtype == torch.int64
    E = edge_index.shape[1]
    if E == 0:
        return edge_index.clone()
    if num_nodes is None:
        num_nodes = int(edge_index.max().item()) + 1
    if not edge_index.is_cuda:
        key = torch.unique(edge_index[0] * num_nodes + edge_index[1], sorted=True)
        return torch.stack([key // num_nodes, key % num_nodes], dim=0)
    from ._lib import call, ptr
    from .graph import build_csr, exclusive_scan
    row, col = edge_index[0].contiguous(), edge_index[1].contiguous()
    by_col = build_csr(col, row, num_nodes)                         # stable by col; .col = row in that order
    rows_a = by_col.col.long()
    by_row = build_csr(rows_a, rows_a, num_nodes)                   # stable by row on top: (row, col) lexicographic
    perm = by_col.perm[by_row.perm.long()].contiguous()             # int32: sorted position -> caller's edge id
    keep = torch.empty(E, dtype=torch.int32, device=edge_index.device)
    call("gmp_mark_unique_pairs", ptr(row), ptr(col), ptr(perm), E, ptr(keep))
    pos = exclusive_scan(keep)                                      # int64[E+1]
    total = int(pos[-1].item())                                     # the output shape is data dependent: one read-back
    out = torch.empty(2, total, dtype=torch.int64, device=edge_index.device)
    call("gmp_compact_pairs", ptr(row), ptr(col), ptr(perm), ptr(keep), ptr(pos), E, ptr(out[0]), ptr(out[1]))
    return out


def to_undirected(edge_index: torch.Tensor, num_nodes: Optional[int] = None) -> torch.Tensor:
    """torch_geometric.utils.to_undirected (experiments/utils/create_graphs.py:79): every edge and its
    reverse, coalesced -- so `edge_index[0]` comes out ascending, which is the order TFN / MACE aggregate in."""
    return coalesce(torch.cat([edge_index, edge_index.flip(0)], dim=1), num_nodes)


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
