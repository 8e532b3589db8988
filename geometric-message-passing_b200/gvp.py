"""GVP-GNN: drop-ins for ``models/layers/gvp_layer.py`` (GVP, LayerNorm, Dropout, GVPConv, GVPConvLayer) and
``models/gvpgnn.py`` (GVPGNNModel).  Features are pairs ``(s [n, ns], V [n, nv, 3])``.

Same constructor arguments, attribute names and ``state_dict`` keys as the reference (including its empty ``dummy_param``
entries).  What runs where: the per-edge geometric-vector-perceptron stack is dense row-wise work (gathered rows, library
GEMMs over the channel axes, elementwise gates) and stays in PyTorch; the message aggregation -- ``MessagePassing(aggr="mean")``
in the reference, i.e. a torch_scatter mean over the destination rows with atomics -- is the deterministic, atomics-free
segmented reduction over the destination-sorted CSR (csrc/segment.cu), like every other scatter in this package.  Not fused
further: the layer is a by-product of the scatter path (SURVEY.md 8f ranks it after the four model families of 8a).
"""
from __future__ import annotations

import functools
from typing import Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from .graph import get_graph
from .scatter import segment_reduce
from .schnet import global_add_pool, global_mean_pool
from .tfn import RadialEmbeddingBlock

SV = Tuple[torch.Tensor, torch.Tensor]


def _vnorm(x: torch.Tensor, dim: int = -1, keepdim: bool = False, eps: float = 1e-8, sqrt: bool = True) -> torch.Tensor:
    """gvp_layer.py:66-73: L2 norm (or its square) clamped from below by eps."""
    sq = x.square().sum(dim, keepdim).clamp(min=eps)
    return sq.sqrt() if sqrt else sq


def merge(s: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """gvp_layer.py:90-98: [n, ns] and [n, nv, 3] -> [n, ns + 3 nv]."""
    return torch.cat([s, v.reshape(v.shape[0], -1)], dim=-1)


def split(x: torch.Tensor, nv: int) -> SV:
    """gvp_layer.py:76-87: inverse of merge."""
    cut = x.shape[-1] - 3 * nv
    return x[..., :cut], x[..., cut:].reshape(x.shape[0], nv, 3)


def tuple_sum(*xs: SV) -> SV:
    return tuple(sum(parts) for parts in zip(*xs))


def tuple_cat(*xs: SV, dim: int = -1) -> SV:
    """gvp_layer.py:28-39: `dim` counts on the scalar tensors (so -1 means -2 for the vector tensors)."""
    dim %= xs[0][0].dim()
    return torch.cat([x[0] for x in xs], dim=dim), torch.cat([x[1] for x in xs], dim=dim)


def tuple_index(x: SV, idx) -> SV:
    return x[0][idx], x[1][idx]


class GVP(nn.Module):
    """Geometric vector perceptron (gvp_layer.py:101-170).  Scalars see the norms of `h_dim` mixed vector channels; vectors
    are mixed linearly (no bias: equivariance) and gated by a sigmoid of the scalar output (vector_gate) or of their norm."""

    def __init__(self, in_dims, out_dims, h_dim=None, activations=(F.relu, torch.sigmoid), vector_gate=True):
        super().__init__()
        self.si, self.vi = in_dims
        self.so, self.vo = out_dims
        self.vector_gate = vector_gate
        if self.vi:
            self.h_dim = h_dim or max(self.vi, self.vo)
            self.wh = nn.Linear(self.vi, self.h_dim, bias=False)
            self.ws = nn.Linear(self.h_dim + self.si, self.so)
            if self.vo:
                self.wv = nn.Linear(self.h_dim, self.vo, bias=False)
                if vector_gate:
                    self.wsv = nn.Linear(self.so, self.vo)
        else:
            self.ws = nn.Linear(self.si, self.so)
        self.scalar_act, self.vector_act = activations
        self.dummy_param = nn.Parameter(torch.empty(0))   # (reference state_dict key)

    def forward(self, x):
        v = None
        if self.vi:
            s, v_in = x
            vh = self.wh(v_in.transpose(-1, -2))                      # [n, 3, h]
            s = self.ws(torch.cat([s, _vnorm(vh, dim=-2)], dim=-1))
            if self.vo:
                v = self.wv(vh).transpose(-1, -2)                     # [n, vo, 3]
                if self.vector_gate:
                    gate = self.wsv(self.vector_act(s) if self.vector_act else s)
                    v = v * torch.sigmoid(gate).unsqueeze(-1)
                elif self.vector_act:
                    v = v * self.vector_act(_vnorm(v, dim=-1, keepdim=True))
        else:
            s = self.ws(x)
            if self.vo:
                v = torch.zeros(s.shape[0], self.vo, 3, device=self.dummy_param.device)
        if self.scalar_act:
            s = self.scalar_act(s)
        return (s, v) if self.vo else s


class _VDropout(nn.Module):
    """gvp_layer.py:173-195: whole vector channels are dropped together."""

    def __init__(self, drop_rate):
        super().__init__()
        self.drop_rate = drop_rate
        self.dummy_param = nn.Parameter(torch.empty(0))

    def forward(self, x):
        if not self.training:
            return x
        keep = 1 - self.drop_rate
        mask = torch.bernoulli(keep * torch.ones(x.shape[:-1], device=self.dummy_param.device)).unsqueeze(-1)
        return mask * x / keep


class Dropout(nn.Module):
    """gvp_layer.py:198-218."""

    def __init__(self, drop_rate):
        super().__init__()
        self.sdropout = nn.Dropout(drop_rate)
        self.vdropout = _VDropout(drop_rate)

    def forward(self, x):
        if isinstance(x, torch.Tensor):
            return self.sdropout(x)
        return self.sdropout(x[0]), self.vdropout(x[1])


class LayerNorm(nn.Module):
    """gvp_layer.py:221-243: nn.LayerNorm on the scalars, the vectors divided by the RMS of their (clamped) norms."""

    def __init__(self, dims):
        super().__init__()
        self.s, self.v = dims
        self.scalar_norm = nn.LayerNorm(self.s)

    def forward(self, x):
        if not self.v:
            return self.scalar_norm(x)
        s, v = x
        rms = _vnorm(v, dim=-1, keepdim=True, sqrt=False).mean(dim=-2, keepdim=True).sqrt()
        return self.scalar_norm(s), v / rms


class GVPConv(nn.Module):
    """gvp_layer.py:246-324: message = GVP stack over cat[(s_j, V_j), edge_attr, (s_i, V_i)], aggregated over the destination
    rows (edge_index[1]; PyG flow source_to_target) by mean (or add).  Does not do the residual / feed-forward: GVPConvLayer."""

    def __init__(self, in_dims, out_dims, edge_dims, n_layers=3, module_list=None, aggr="mean",
                 activations=(F.relu, torch.sigmoid), vector_gate=True):
        super().__init__()
        if aggr not in ("mean", "add", "sum"):
            raise NotImplementedError(f"aggr={aggr!r}: the segmented reduction implements mean / add")
        self.aggr = aggr
        self.si, self.vi = in_dims
        self.so, self.vo = out_dims
        self.se, self.ve = edge_dims
        mk = functools.partial(GVP, activations=activations, vector_gate=vector_gate)
        cat_dims = (2 * self.si + self.se, 2 * self.vi + self.ve)
        stack = list(module_list or [])
        if not stack:
            if n_layers == 1:
                stack = [mk(cat_dims, (self.so, self.vo), activations=(None, None))]
            else:
                stack = [mk(cat_dims, out_dims)] + [mk(out_dims, out_dims) for _ in range(n_layers - 2)]
                stack.append(mk(out_dims, out_dims, activations=(None, None)))
        self.message_func = nn.Sequential(*stack)

    def forward(self, x: SV, edge_index: torch.Tensor, edge_attr: SV) -> SV:
        s, v = x
        j, i = edge_index[0], edge_index[1]
        msg = self.message_func(tuple_cat((s[j], v[j]), edge_attr, (s[i], v[i])))
        graph = get_graph(edge_index, s.shape[0])
        out = segment_reduce(merge(*msg), graph.by_dst, "mean" if self.aggr == "mean" else "sum")
        return split(out, self.vo)


class GVPConvLayer(nn.Module):
    """gvp_layer.py:327-438: x <- LN(x + drop(conv(x))), x <- LN(x + drop(ff(x)))."""

    def __init__(self, node_dims, edge_dims, n_message=3, n_feedforward=2, drop_rate=0.1, autoregressive=False,
                 activations=(F.relu, torch.sigmoid), vector_gate=True, residual=True):
        super().__init__()
        self.conv = GVPConv(node_dims, node_dims, edge_dims, n_message, aggr="add" if autoregressive else "mean",
                            activations=activations, vector_gate=vector_gate)
        mk = functools.partial(GVP, activations=activations, vector_gate=vector_gate)
        self.norm = nn.ModuleList([LayerNorm(node_dims) for _ in range(2)])
        self.dropout = nn.ModuleList([Dropout(drop_rate) for _ in range(2)])
        if n_feedforward == 1:
            ff = [mk(node_dims, node_dims, activations=(None, None))]
        else:
            hid = (4 * node_dims[0], 2 * node_dims[1])
            ff = [mk(node_dims, hid)] + [mk(hid, hid) for _ in range(n_feedforward - 2)] + [mk(hid, node_dims, activations=(None, None))]
        self.ff_func = nn.Sequential(*ff)
        self.residual = residual

    def forward(self, x: SV, edge_index, edge_attr: SV, autoregressive_x: Optional[SV] = None, node_mask=None) -> SV:
        if autoregressive_x is not None:
            # messages along src < dst use the current embeddings, the others `autoregressive_x`; sums divided by the in-degree
            src, dst = edge_index
            fwd = src < dst
            dh = tuple_sum(self.conv(x, edge_index[:, fwd], tuple_index(edge_attr, fwd)),
                           self.conv(autoregressive_x, edge_index[:, ~fwd], tuple_index(edge_attr, ~fwd)))
            deg = torch.bincount(dst, minlength=dh[0].shape[0]).clamp(min=1).to(dh[0].dtype).unsqueeze(-1)
            dh = (dh[0] / deg, dh[1] / deg.unsqueeze(-1))
        else:
            dh = self.conv(x, edge_index, edge_attr)
        full = None
        if node_mask is not None:
            full = x
            x, dh = tuple_index(x, node_mask), tuple_index(dh, node_mask)
        x = self.norm[0](tuple_sum(x, self.dropout[0](dh))) if self.residual else dh
        dh = self.ff_func(x)
        x = self.norm[1](tuple_sum(x, self.dropout[1](dh))) if self.residual else dh
        if full is not None:
            full[0][node_mask], full[1][node_mask] = x[0], x[1]
            x = full
        return x


class GVPGNNModel(nn.Module):
    """models/gvpgnn.py:9-126."""

    def __init__(self, r_max: float = 10.0, num_bessel: int = 8, num_polynomial_cutoff: int = 5, num_layers: int = 5, in_dim=1,
                 out_dim=1, s_dim: int = 128, v_dim: int = 16, s_dim_edge: int = 32, v_dim_edge: int = 1, pool: str = "sum",
                 residual: bool = True, equivariant_pred: bool = False):
        super().__init__()
        self.r_max, self.num_layers, self.equivariant_pred = r_max, num_layers, equivariant_pred
        self.s_dim, self.v_dim = s_dim, v_dim
        acts = (F.relu, None)
        node_dims, edge_dims = (s_dim, v_dim), (s_dim_edge, v_dim_edge)
        self.emb_in = nn.Embedding(in_dim, s_dim)
        self.W_v = nn.Sequential(LayerNorm((s_dim, 0)), GVP((s_dim, 0), node_dims, activations=(None, None), vector_gate=True))
        self.radial_embedding = RadialEmbeddingBlock(r_max=r_max, num_bessel=num_bessel, num_polynomial_cutoff=num_polynomial_cutoff)
        self.W_e = nn.Sequential(LayerNorm((self.radial_embedding.out_dim, 1)),
                                 GVP((self.radial_embedding.out_dim, 1), edge_dims, activations=(None, None), vector_gate=True))
        self.layers = nn.ModuleList(GVPConvLayer(node_dims, edge_dims, activations=acts, vector_gate=True, residual=residual)
                                    for _ in range(num_layers))
        self.pool = {"mean": global_mean_pool, "sum": global_add_pool}[pool]
        if equivariant_pred:
            self.pred = nn.Linear(s_dim + v_dim * 3, out_dim)
        else:
            self.pred = nn.Sequential(nn.Linear(s_dim, s_dim), nn.ReLU(), nn.Linear(s_dim, out_dim))

    def forward(self, batch):
        ei = batch.edge_index
        vec = batch.pos[ei[0]] - batch.pos[ei[1]]
        length = torch.linalg.norm(vec, dim=-1, keepdim=True)
        h_v = self.W_v(self.emb_in(batch.atoms))
        h_e = self.W_e((self.radial_embedding(length), torch.nan_to_num(vec / length).unsqueeze(-2)))
        for layer in self.layers:
            h_v = layer(h_v, ei, h_e)
        out = self.pool(merge(*h_v), batch.batch, getattr(batch, "num_graphs", None))
        if not self.equivariant_pred:
            out = out[:, :self.s_dim]
        return self.pred(out)
