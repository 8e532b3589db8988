"""The ACEsuit-style MACE interaction blocks: drop-ins for ``models/mace_modules/blocks.py:136-530``.

Every block is: ``linear_up`` (o3.Linear) -> gather at the sender -> ``conv_tp`` (a 'uvu' tensor product with the edge
attributes, one weight per (edge, path, channel) from a radial MLP) -> ``scatter_sum`` at the receiver -> ``linear`` ->
division by ``avg_num_neighbors`` -> the ``skip_tp`` selector product with the one-hot node attributes.  Constructor
arguments, attribute names, ``forward(node_attrs, node_feats, edge_attrs, edge_feats, edge_index)`` and the
``state_dict`` keys (``linear_up.weight``, ``conv_tp_weights.layer{0..3}.weight`` / ``conv_tp_weights.weights``,
``linear.weight``, ``skip_tp.weight``) are the reference's.

The per-edge part -- gather, 'uvu' product, segmented sum over the receiver rows -- is one kernel per direction
(csrc/uvu.cu) when node features, edge attributes and target are ``C x (0e+1o+2e)`` / ``0e+1o+2e`` (the reference
models' l_max = 2 shape: 11 paths); the message ``mji [E, 35 C]`` of the reference is never stored.  Other irreps take
an unfused path (torch einsum per path + the deterministic segmented reduction).  The per-edge radial weights
``[E, 11 C]`` are what the reference's radial MLP produces and stay a tensor (they are the MLP's output, 44 B per
channel and edge); gradients w.r.t. the edge attributes (spherical harmonics) are not produced (as in
TensorProductConvLayer).
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch
from torch import nn

from . import _lib
from ._lib import call, ptr
from .graph import get_graph
from .irreps import Irrep, Irreps, wigner_3j
from .mace import EquivariantLinear, reshape_irreps
from .scatter import segment_reduce

_SILU_2MOM = 1.6791767923989418   # e3nn normalize2mom(silu): 1/sqrt(E_{z~N(0,1)} silu(z)^2) on its seeded 1e6 samples


def _sorted_irreps(irreps: Irreps):
    """e3nn ``Irreps.sort()``: (sorted irreps, p, inv) with p[old index] = new index; stable for equal irreps."""
    order = sorted(range(len(irreps)), key=lambda i: ((irreps[i][1].l, -irreps[i][1].p), i))   # 0e < 0o < 1e ... as e3nn: (l, -p)?
    # e3nn orders irreps by (l, p) with p = -(-1)^l first?  Its Irrep.__lt__ is: (l, -p*(-1)^l) -- 0e, 0o, 1o, 1e, 2e, 2o, ...
    order = sorted(range(len(irreps)), key=lambda i: ((irreps[i][1].l, -irreps[i][1].p * (-1) ** irreps[i][1].l), i))
    p = [0] * len(order)
    for pos, i in enumerate(order):
        p[i] = pos
    out = Irreps("")
    out.items = [irreps[i] for i in order]
    return out, tuple(p), tuple(order)


def _product_irreps(a: Irrep, b: Irrep) -> List[Irrep]:
    return [Irrep(l, a.p * b.p) for l in range(abs(a.l - b.l), a.l + b.l + 1)]


def tp_out_irreps_with_instructions(irreps1, irreps2, target_irreps) -> Tuple[Irreps, List[tuple]]:
    """models/mace_modules/irreps_tools.py:14-44: one 'uvu' instruction per (in1 block, in2 block, l_out in target), the
    output blocks sorted by irrep so that the following o3.Linear sees them grouped."""
    irreps1, irreps2, target = Irreps(str(irreps1)), Irreps(str(irreps2)), Irreps(str(target_irreps))
    tset = {ir for _, ir in target}
    out_list, instructions = [], []
    for i, (mul, ir_in) in enumerate(irreps1):
        for j, (_, ir_edge) in enumerate(irreps2):
            for ir_out in _product_irreps(ir_in, ir_edge):
                if ir_out in tset:
                    k = len(out_list)
                    out_list.append((mul, ir_out))
                    instructions.append((i, j, k, "uvu", True))
    mid = Irreps("")
    mid.items = out_list
    mid, permut, _ = _sorted_irreps(mid)
    return mid, [(i1, i2, permut[io], mode, train) for i1, i2, io, mode, train in instructions]


def linear_out_irreps(irreps, target_irreps) -> Irreps:
    """models/mace_modules/irreps_tools.py:47-62."""
    irreps, target = Irreps(str(irreps)), Irreps(str(target_irreps))
    out = []
    for _, ir_in in irreps:
        hit = [(mul, ir_out) for mul, ir_out in target if ir_in == ir_out]
        if not hit:
            raise RuntimeError(f"{ir_in} not in {target}")
        out.append(hit[0])
    r = Irreps("")
    r.items = out
    return r


class FullyConnectedNet(nn.Sequential):
    """e3nn ``nn.FullyConnectedNet(hs, act)``: bias-free layers ``act(x @ W / sqrt(h_in))`` with the second-moment
    normalised activation, last layer linear (blocks.py:243-246); keys ``layer{i}.weight`` ([h_in, h_out], randn)."""

    class _Layer(nn.Module):
        def __init__(self, h_in, h_out, act_cst):
            super().__init__()
            self.weight = nn.Parameter(torch.randn(h_in, h_out))
            self.h_in, self.act_cst = h_in, act_cst

        def forward(self, x):
            x = x @ (self.weight / self.h_in ** 0.5)
            return torch.nn.functional.silu(x) * self.act_cst if self.act_cst is not None else x

    def __init__(self, hs, act=None):
        super().__init__()
        if act is not None and not isinstance(act, nn.SiLU) and act is not torch.nn.functional.silu:
            raise NotImplementedError("gmp_b200 FullyConnectedNet: the reference's blocks use SiLU")
        self.hs = list(hs)
        for i, (h1, h2) in enumerate(zip(self.hs, self.hs[1:])):
            last = i == len(self.hs) - 2
            setattr(self, f"layer{i}", FullyConnectedNet._Layer(h1, h2, None if (last or act is None) else _SILU_2MOM))


class TensorProductWeightsBlock(nn.Module):
    """models/mace_modules/blocks.py:177-203 (element-dependent radial weights)."""

    def __init__(self, num_elements: int, num_edge_feats: int, num_feats_out: int):
        super().__init__()
        w = torch.empty(num_elements, num_edge_feats, num_feats_out)
        nn.init.xavier_uniform_(w)
        self.weights = nn.Parameter(w)

    def forward(self, sender_or_receiver_node_attrs, edge_feats):
        return torch.einsum("be,ba,aek->bk", edge_feats, sender_or_receiver_node_attrs, self.weights)


class FullyConnectedTensorProduct(nn.Module):
    """e3nn ``o3.FullyConnectedTensorProduct(in1, in2, out)`` with internal shared weights (the ``skip_tp`` selector of
    the blocks; in2 = the one-hot node attributes, ``num_elements x 0e``): 'uvw' paths, e3nn normalisation.
    Node-level (N rows), plain einsums."""

    def __init__(self, irreps_in1, irreps_in2, irreps_out):
        super().__init__()
        self.irreps_in1, self.irreps_in2, self.irreps_out = Irreps(str(irreps_in1)), Irreps(str(irreps_in2)), Irreps(str(irreps_out))
        self.paths = []
        for i1, (m1, ir1) in enumerate(self.irreps_in1):
            for i2, (m2, ir2) in enumerate(self.irreps_in2):
                for io, (mo, iro) in enumerate(self.irreps_out):
                    if iro in _product_irreps(ir1, ir2):
                        self.paths.append((i1, i2, io))
        fan = {}
        for i1, i2, io in self.paths:
            fan[io] = fan.get(io, 0) + self.irreps_in1[i1][0] * self.irreps_in2[i2][0]
        self.path_weight = [math.sqrt(self.irreps_out[io][1].dim / fan[io]) for _, _, io in self.paths]
        self.weight_numel = sum(self.irreps_in1[i1][0] * self.irreps_in2[i2][0] * self.irreps_out[io][0] for i1, i2, io in self.paths)
        self.weight = nn.Parameter(torch.randn(self.weight_numel))
        for n, (i1, i2, io) in enumerate(self.paths):
            cg = wigner_3j(self.irreps_in1[i1][1].l, self.irreps_in2[i2][1].l, self.irreps_out[io][1].l)
            self.register_buffer(f"_cg{n}", torch.from_numpy(cg).float(), persistent=False)

    def forward(self, x1, x2):
        N = x1.shape[0]
        o1, o2 = self.irreps_in1.offsets(), self.irreps_in2.offsets()
        outs = [None] * len(self.irreps_out)
        off = 0
        for n, (i1, i2, io) in enumerate(self.paths):
            (m1, ir1), (m2, ir2), (mo, iro) = self.irreps_in1[i1], self.irreps_in2[i2], self.irreps_out[io]
            W = self.weight[off:off + m1 * m2 * mo].view(m1, m2, mo)
            off += m1 * m2 * mo
            a = x1[:, o1[i1]:o1[i1] + m1 * ir1.dim].reshape(N, m1, ir1.dim)
            b = x2[:, o2[i2]:o2[i2] + m2 * ir2.dim].reshape(N, m2, ir2.dim)
            t = torch.einsum("ijk,nui,nvj->nuvk", getattr(self, f"_cg{n}"), a, b)
            r = torch.einsum("uvw,nuvk->nwk", W, t).reshape(N, mo * iro.dim) * self.path_weight[n]
            outs[io] = r if outs[io] is None else outs[io] + r
        for k, (mo, iro) in enumerate(self.irreps_out):
            if outs[k] is None:
                outs[k] = x1.new_zeros(N, mo * iro.dim)
        return torch.cat(outs, dim=-1)


def _is_l2_shape(feats: Irreps, sh: Irreps, target: Irreps) -> int:
    """C when (feats, sh, target) = (C x (0e+1o+2e), 0e+1o+2e, any multiplicities of 0e+1o+2e), else 0."""
    want = [(0, 1), (1, -1), (2, 1)]
    ok = (len(feats) == 3 and len(sh) == 3 and [(ir.l, ir.p) for _, ir in feats] == want and [(ir.l, ir.p) for _, ir in sh] == want
          and all(m == 1 for m, _ in sh) and len({m for m, _ in feats}) == 1 and {(ir.l, ir.p) for _, ir in target} == set(want))
    return feats[0][0] if ok and feats[0][0] <= 128 else 0


class _UVUConvFn(torch.autograd.Function):
    """message[i] = sum_{e: receiver_e = i} uvu(node_feats[sender_e], edge_attrs_e, w_e)   (blocks.py:446-452)."""

    @staticmethod
    def forward(ctx, x, edge_attrs, w, graph, C):
        x, edge_attrs, w = x.contiguous(), edge_attrs.contiguous(), w.contiguous()
        csr = graph.by_dst                       # rows = receivers (edge_index[1]), col = senders
        out = torch.empty(graph.n, 35 * C, dtype=x.dtype, device=x.device)
        call("gmp_uvu_conv_fwd", ptr(csr.rowptr), ptr(csr.col), csr.perm_ptr, graph.n, graph.E, ptr(x), ptr(edge_attrs), ptr(w), C, ptr(out))
        ctx.save_for_backward(x, edge_attrs, w)
        ctx.graph, ctx.C = graph, C
        return out

    @staticmethod
    def backward(ctx, g):
        x, edge_attrs, w = ctx.saved_tensors
        graph, C = ctx.graph, ctx.C
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("gmp_b200 uvu convolution: gradients w.r.t. the edge attributes (spherical harmonics) are not built")
        g = g.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            t = graph.by_src                     # rows = senders, col = receivers
            dx = torch.empty_like(x)
            call("gmp_uvu_conv_dx", ptr(t.rowptr), ptr(t.col), t.perm_ptr, graph.n, graph.E, ptr(g), ptr(edge_attrs), ptr(w), C, ptr(dx))
        if ctx.needs_input_grad[2]:
            dw = torch.empty_like(w)
            ei = graph.edge_index
            call("gmp_uvu_conv_dw", ptr(ei[0]), ptr(ei[1]), graph.E, ptr(x), ptr(g), ptr(edge_attrs), C, ptr(dw))
        return dx, None, dw, None, None


class UVUTensorProduct(nn.Module):
    """``o3.TensorProduct(node_feats, edge_attrs, irreps_mid, instructions='uvu', shared_weights=False)`` fused with the
    sender gather and the receiver scatter_sum."""

    def __init__(self, irreps_in1, irreps_in2, target_irreps):
        super().__init__()
        self.irreps_in1, self.irreps_in2 = Irreps(str(irreps_in1)), Irreps(str(irreps_in2))
        self.irreps_out, self.instructions = tp_out_irreps_with_instructions(self.irreps_in1, self.irreps_in2, target_irreps)
        self.weight_numel = sum(self.irreps_in1[i1][0] * self.irreps_in2[i2][0] for i1, i2, _, _, _ in self.instructions)
        self.fused_C = _is_l2_shape(self.irreps_in1, self.irreps_in2, Irreps(str(target_irreps)))
        for n, (i1, i2, io, _, _) in enumerate(self.instructions):
            cg = wigner_3j(self.irreps_in1[i1][1].l, self.irreps_in2[i2][1].l, self.irreps_out[io][1].l)
            # every mid block is fed by exactly one path: e3nn's 'element' normalisation is sqrt(2 l_out + 1)
            self.register_buffer(f"_cg{n}", torch.from_numpy(cg * math.sqrt(self.irreps_out[io][1].dim)).float(), persistent=False)

    def forward(self, node_feats, edge_index, edge_attrs, tp_weights):
        graph = get_graph(edge_index, node_feats.shape[0])
        if self.fused_C and node_feats.is_cuda:
            return _UVUConvFn.apply(node_feats, edge_attrs, tp_weights, graph, self.fused_C)
        # unfused: per-path einsum on the gathered rows, then the deterministic segmented sum at the receivers
        sender = edge_index[0]
        xs = node_feats[sender]
        E = xs.shape[0]
        o1, o2 = self.irreps_in1.offsets(), self.irreps_in2.offsets()
        outs = [None] * len(self.irreps_out)
        off = 0
        for n, (i1, i2, io, _, _) in enumerate(self.instructions):
            (m1, ir1), (m2, ir2) = self.irreps_in1[i1], self.irreps_in2[i2]
            a = xs[:, o1[i1]:o1[i1] + m1 * ir1.dim].reshape(E, m1, ir1.dim)
            b = edge_attrs[:, o2[i2]:o2[i2] + m2 * ir2.dim].reshape(E, m2, ir2.dim)
            wgt = tp_weights[:, off:off + m1 * m2].reshape(E, m1, m2)
            off += m1 * m2
            t = torch.einsum("ijk,eui,evj->euvk", getattr(self, f"_cg{n}"), a, b)
            outs[io] = torch.einsum("euv,euvk->euk", wgt, t).reshape(E, -1)
        mji = torch.cat(outs, dim=-1)
        return segment_reduce(mji, graph.by_dst, "sum")


class InteractionBlock(nn.Module):
    """models/mace_modules/blocks.py:136-171 (constructor contract of every block)."""

    def __init__(self, node_attrs_irreps, node_feats_irreps, edge_attrs_irreps, edge_feats_irreps, target_irreps, hidden_irreps,
                 avg_num_neighbors: float) -> None:
        super().__init__()
        self.node_attrs_irreps, self.node_feats_irreps = Irreps(str(node_attrs_irreps)), Irreps(str(node_feats_irreps))
        self.edge_attrs_irreps, self.edge_feats_irreps = Irreps(str(edge_attrs_irreps)), Irreps(str(edge_feats_irreps))
        self.target_irreps, self.hidden_irreps = Irreps(str(target_irreps)), Irreps(str(hidden_irreps))
        self.avg_num_neighbors = avg_num_neighbors
        self._setup()

    # shared by every variant: linear_up, conv_tp, radial weights, linear
    def _common(self, element_dependent: bool, out_from_target: bool) -> None:
        self.linear_up = EquivariantLinear(self.node_feats_irreps, self.node_feats_irreps)
        self.conv_tp = UVUTensorProduct(self.node_feats_irreps, self.edge_attrs_irreps, self.target_irreps)
        if element_dependent:
            self.conv_tp_weights = TensorProductWeightsBlock(self.node_attrs_irreps.num_irreps, self.edge_feats_irreps.num_irreps,
                                                             self.conv_tp.weight_numel)
        else:
            self.conv_tp_weights = FullyConnectedNet([self.edge_feats_irreps.num_irreps] + 3 * [64] + [self.conv_tp.weight_numel], nn.SiLU())
        irreps_mid = self.conv_tp.irreps_out.simplify()
        self.irreps_out = self.target_irreps if out_from_target else linear_out_irreps(irreps_mid, self.target_irreps).simplify()
        self.linear = EquivariantLinear(irreps_mid, self.irreps_out)

    def _message(self, node_attrs, node_feats, edge_attrs, edge_feats, edge_index):
        node_feats = self.linear_up(node_feats)
        if isinstance(self.conv_tp_weights, TensorProductWeightsBlock):
            tp_weights = self.conv_tp_weights(node_attrs[edge_index[0]], edge_feats)
        else:
            tp_weights = self.conv_tp_weights(edge_feats)
        message = self.conv_tp(node_feats, edge_index, edge_attrs, tp_weights)       # gather + uvu + scatter_sum, fused
        return self.linear(message) / self.avg_num_neighbors


class ResidualElementDependentInteractionBlock(InteractionBlock):
    """blocks.py:206-273."""

    def _setup(self) -> None:
        self._common(element_dependent=True, out_from_target=False)
        self.skip_tp = FullyConnectedTensorProduct(self.node_feats_irreps, self.node_attrs_irreps, self.irreps_out)

    def forward(self, node_attrs, node_feats, edge_attrs, edge_feats, edge_index):
        sc = self.skip_tp(node_feats, node_attrs)
        return self._message(node_attrs, node_feats, edge_attrs, edge_feats, edge_index) + sc


class AgnosticNonlinearInteractionBlock(InteractionBlock):
    """blocks.py:276-327."""

    def _setup(self) -> None:
        self._common(element_dependent=False, out_from_target=False)
        self.skip_tp = FullyConnectedTensorProduct(self.irreps_out, self.node_attrs_irreps, self.irreps_out)

    def forward(self, node_attrs, node_feats, edge_attrs, edge_feats, edge_index):
        return self.skip_tp(self._message(node_attrs, node_feats, edge_attrs, edge_feats, edge_index), node_attrs)


class AgnosticResidualNonlinearInteractionBlock(InteractionBlock):
    """blocks.py:330-393."""

    def _setup(self) -> None:
        self._common(element_dependent=False, out_from_target=False)
        self.skip_tp = FullyConnectedTensorProduct(self.node_feats_irreps, self.node_attrs_irreps, self.irreps_out)

    def forward(self, node_attrs, node_feats, edge_attrs, edge_feats, edge_index):
        sc = self.skip_tp(node_feats, node_attrs)
        return self._message(node_attrs, node_feats, edge_attrs, edge_feats, edge_index) + sc


class RealAgnosticInteractionBlock(InteractionBlock):
    """blocks.py:396-459: returns (message reshaped to [N, channels, (l_max+1)^2], None)."""

    def _setup(self) -> None:
        self._common(element_dependent=False, out_from_target=True)
        self.skip_tp = FullyConnectedTensorProduct(self.irreps_out, self.node_attrs_irreps, self.irreps_out)
        self.reshape = reshape_irreps(self.irreps_out)

    def forward(self, node_attrs, node_feats, edge_attrs, edge_feats, edge_index) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        message = self.skip_tp(self._message(node_attrs, node_feats, edge_attrs, edge_feats, edge_index), node_attrs)
        return self.reshape(message), None


class RealAgnosticResidualInteractionBlock(InteractionBlock):
    """blocks.py:462-530: returns (message reshaped, sc)."""

    def _setup(self) -> None:
        self._common(element_dependent=False, out_from_target=True)
        self.skip_tp = FullyConnectedTensorProduct(self.node_feats_irreps, self.node_attrs_irreps, self.hidden_irreps)
        self.reshape = reshape_irreps(self.irreps_out)

    def forward(self, node_attrs, node_feats, edge_attrs, edge_feats, edge_index) -> Tuple[torch.Tensor, torch.Tensor]:
        sc = self.skip_tp(node_feats, node_attrs)
        return self.reshape(self._message(node_attrs, node_feats, edge_attrs, edge_feats, edge_index)), sc
