"""Restatement of the PyG 2.3.1 pieces the reference calls.

ORACLE / TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED against the wheel.
Reference call sites: models/layers/egnn_layer.py:7,59 (MessagePassing.propagate);
models/schnet.py:5,9,41-54,72 (SchNet, InteractionBlock, CFConv);
models/egnn.py:54,83 / models/schnet.py:57,74 (global pooling);
experiments/kchains.ipynb:71-107 (Data, Batch, to_undirected).
"""
from __future__ import annotations

import inspect
import math
from typing import List, Optional

import torch
from torch.nn import Embedding, Linear, ModuleList, Sequential

from .scatter import scatter


# --------------------------------------------------------------------------- #
# MessagePassing.propagate (flow = source_to_target, node_dim = -2)
# --------------------------------------------------------------------------- #
class MessagePassing(torch.nn.Module):
    def __init__(self, aggr: Optional[str] = "add", flow: str = "source_to_target", node_dim: int = -2):
        super().__init__()
        assert flow == "source_to_target"
        self.aggr = aggr
        self.flow = flow
        self.node_dim = node_dim

    def propagate(self, edge_index: torch.Tensor, size=None, **kwargs):
        j, i = edge_index[0], edge_index[1]
        num_nodes = None
        for v in kwargs.values():
            if isinstance(v, torch.Tensor) and v.dim() >= 2:
                num_nodes = v.size(self.node_dim)
                break
        coll = {"index": i, "edge_index": edge_index, "dim_size": num_nodes, "ptr": None}
        coll.update(kwargs)
        msg_params = list(inspect.signature(self.message).parameters)
        msg_args = {}
        for name in msg_params:
            if name.endswith("_i") or name.endswith("_j"):
                base = kwargs[name[:-2]]
                msg_args[name] = base.index_select(self.node_dim, i if name.endswith("_i") else j)
            else:
                msg_args[name] = coll[name]
        out = self.message(**msg_args)
        aggr_params = list(inspect.signature(self.aggregate).parameters)[1:]
        out = self.aggregate(out, **{k: coll[k] for k in aggr_params if k in coll})
        upd_params = list(inspect.signature(self.update).parameters)[1:]
        return self.update(out, **{k: coll[k] for k in upd_params if k in coll})

    def message(self, x_j):
        return x_j

    def aggregate(self, inputs, index, ptr=None, dim_size=None):
        return scatter(inputs, index, dim=self.node_dim, dim_size=dim_size, reduce=self.aggr)

    def update(self, inputs):
        return inputs


# --------------------------------------------------------------------------- #
# Pooling, Data / Batch, to_undirected
# --------------------------------------------------------------------------- #
def global_add_pool(x, batch, size: Optional[int] = None):
    size = int(batch.max()) + 1 if size is None else size
    return scatter(x, batch, dim=-2, dim_size=size, reduce="sum")


def global_mean_pool(x, batch, size: Optional[int] = None):
    size = int(batch.max()) + 1 if size is None else size
    return scatter(x, batch, dim=-2, dim_size=size, reduce="mean")


def aggr_resolver(name: str):
    return {"add": global_add_pool, "sum": global_add_pool, "mean": global_mean_pool}[name]


def to_undirected(edge_index: torch.Tensor) -> torch.Tensor:
    ei = torch.cat([edge_index, edge_index.flip(0)], dim=1)
    n = int(ei.max()) + 1 if ei.numel() else 0
    key = torch.unique(ei[0] * n + ei[1], sorted=True)
    return torch.stack([key // n, key % n], dim=0)


class Data:
    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        for k in ("atoms", "pos", "x"):
            v = getattr(self, k, None)
            if v is not None:
                return v.shape[0]
        raise AttributeError("num_nodes")

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    def to(self, device):
        for k in self.keys():
            v = getattr(self, k)
            if isinstance(v, torch.Tensor):
                setattr(self, k, v.to(device))
        return self


class Batch(Data):
    @staticmethod
    def from_data_list(data_list: List[Data]) -> "Batch":
        out = Batch()
        keys = data_list[0].keys()
        offs, batch_vec, off = [], [], 0
        for g, d in enumerate(data_list):
            offs.append(off)
            batch_vec.append(torch.full((d.num_nodes,), g, dtype=torch.long))
            off += d.num_nodes
        for k in keys:
            vals = [getattr(d, k) for d in data_list]
            if k == "edge_index":
                setattr(out, k, torch.cat([v + o for v, o in zip(vals, offs)], dim=1))
            elif isinstance(vals[0], torch.Tensor):
                vals = [v if v.dim() > 0 else v.reshape(1) for v in vals]
                setattr(out, k, torch.cat(vals, dim=0))
            else:
                setattr(out, k, vals)
        out.batch = torch.cat(batch_vec)
        out.num_graphs = len(data_list)
        return out


# --------------------------------------------------------------------------- #
# SchNet (PyG 2.3.1 `torch_geometric.nn.models.schnet`)
# --------------------------------------------------------------------------- #
class ShiftedSoftplus(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.shift = math.log(2.0)

    def forward(self, x):
        return torch.nn.functional.softplus(x) - self.shift


class GaussianSmearing(torch.nn.Module):
    def __init__(self, start: float = 0.0, stop: float = 5.0, num_gaussians: int = 50):
        super().__init__()
        offset = torch.linspace(start, stop, num_gaussians)
        self.coeff = -0.5 / (offset[1] - offset[0]).item() ** 2
        self.register_buffer("offset", offset)

    def forward(self, dist):
        dist = dist.view(-1, 1) - self.offset.view(1, -1)
        return torch.exp(self.coeff * torch.pow(dist, 2))


class CFConv(MessagePassing):
    def __init__(self, in_channels, out_channels, num_filters, nn, cutoff):
        super().__init__(aggr="add")
        self.lin1 = Linear(in_channels, num_filters, bias=False)
        self.lin2 = Linear(num_filters, out_channels)
        self.nn = nn
        self.cutoff = cutoff
        self.reset_parameters()

    def reset_parameters(self):
        torch.nn.init.xavier_uniform_(self.lin1.weight)
        torch.nn.init.xavier_uniform_(self.lin2.weight)
        self.lin2.bias.data.fill_(0)

    def forward(self, x, edge_index, edge_weight, edge_attr):
        C = 0.5 * (torch.cos(edge_weight * math.pi / self.cutoff) + 1.0)
        W = self.nn(edge_attr) * C.view(-1, 1)
        x = self.lin1(x)
        x = self.propagate(edge_index, x=x, W=W)
        return self.lin2(x)

    def message(self, x_j, W):
        return x_j * W


class InteractionBlock(torch.nn.Module):
    def __init__(self, hidden_channels, num_gaussians, num_filters, cutoff):
        super().__init__()
        self.mlp = Sequential(Linear(num_gaussians, num_filters), ShiftedSoftplus(), Linear(num_filters, num_filters))
        self.conv = CFConv(hidden_channels, hidden_channels, num_filters, self.mlp, cutoff)
        self.act = ShiftedSoftplus()
        self.lin = Linear(hidden_channels, hidden_channels)
        self.reset_parameters()

    def reset_parameters(self):
        torch.nn.init.xavier_uniform_(self.mlp[0].weight)
        self.mlp[0].bias.data.fill_(0)
        torch.nn.init.xavier_uniform_(self.mlp[2].weight)
        self.mlp[2].bias.data.fill_(0)
        self.conv.reset_parameters()
        torch.nn.init.xavier_uniform_(self.lin.weight)
        self.lin.bias.data.fill_(0)

    def forward(self, x, edge_index, edge_weight, edge_attr):
        x = self.conv(x, edge_index, edge_weight, edge_attr)
        x = self.act(x)
        return self.lin(x)


class SchNet(torch.nn.Module):
    def __init__(self, hidden_channels=128, num_filters=128, num_interactions=6, num_gaussians=50, cutoff=10.0,
                 interaction_graph=None, max_num_neighbors=32, readout="add", dipole=False, mean=None, std=None,
                 atomref=None):
        super().__init__()
        assert not dipole and atomref is None
        self.hidden_channels, self.num_filters = hidden_channels, num_filters
        self.num_interactions, self.num_gaussians, self.cutoff = num_interactions, num_gaussians, cutoff
        self.readout = aggr_resolver(readout)
        self.mean, self.std, self.scale = mean, std, None
        self.embedding = Embedding(100, hidden_channels, padding_idx=0)
        self.distance_expansion = GaussianSmearing(0.0, cutoff, num_gaussians)
        self.interactions = ModuleList(
            [InteractionBlock(hidden_channels, num_gaussians, num_filters, cutoff) for _ in range(num_interactions)]
        )
        self.lin1 = Linear(hidden_channels, hidden_channels // 2)
        self.act = ShiftedSoftplus()
        self.lin2 = Linear(hidden_channels // 2, 1)
        self.reset_parameters()

    def reset_parameters(self):
        self.embedding.reset_parameters()
        for it in self.interactions:
            it.reset_parameters()
        torch.nn.init.xavier_uniform_(self.lin1.weight)
        self.lin1.bias.data.fill_(0)
        torch.nn.init.xavier_uniform_(self.lin2.weight)
        self.lin2.bias.data.fill_(0)
