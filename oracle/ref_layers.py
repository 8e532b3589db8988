"""CPU restatement of the reference's own hot-path code (EGNN / SchNet / TFN / MACE).

ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function
cites the reference file:line it follows.  Parameter names and shapes equal the
reference modules', so a reference ``state_dict`` loads unchanged; the golden
vectors in tests/golden/ (made by running the unmodified reference under
oracle/shims) pin this file.

Written functionally on purpose: each layer is a pure function of
``(params, inputs)`` plus a thin ``nn.Module`` that owns the parameters.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F
from torch import nn

from .thirdparty import o3
from .thirdparty.e3nn_nn import BatchNorm as E3BatchNorm
from .thirdparty.e3nn_nn import Gate as E3Gate
from .thirdparty.e3nn_nn import Activation as E3Activation
from .thirdparty.einsum import contract
from .thirdparty.pyg import SchNet as _PygSchNet
from .thirdparty.pyg import global_add_pool, global_mean_pool
from .thirdparty.scatter import scatter

_ACT = {"relu": nn.ReLU, "swish": nn.SiLU}
_NORM = {"layer": nn.LayerNorm, "batch": nn.BatchNorm1d}
_POOL = {"sum": global_add_pool, "mean": global_mean_pool}


# =========================================================================== #
# EGNN   (models/layers/egnn_layer.py:7-89, models/egnn.py:8-87)
# =========================================================================== #
def _mlp(dims_in: int, d: int, act: str, norm: str, tail: Optional[int] = None) -> nn.Sequential:
    """Linear-norm-act-Linear(-norm-act | ->tail)  (egnn_layer.py:28-48)."""
    mods: List[nn.Module] = [nn.Linear(dims_in, d), _NORM[norm](d), _ACT[act]()]
    if tail is None:
        mods += [nn.Linear(d, d), _NORM[norm](d), _ACT[act]()]
    else:
        mods += [nn.Linear(d, tail)]
    return nn.Sequential(*mods)


def egnn_edge_message(layer: "EGNNLayer", h, pos, edge_index):
    """egnn_layer.py:62-72 with PyG's gather (A.2): j = edge_index[0], i = edge_index[1]."""
    j, i = edge_index[0], edge_index[1]
    delta = pos[i] - pos[j]
    dist = delta.norm(dim=-1, keepdim=True)
    m = layer.mlp_msg(torch.cat([h[i], h[j], dist], dim=-1))
    return m, delta * layer.mlp_pos(m)


class EGNNLayer(nn.Module):
    def __init__(self, emb_dim, activation="relu", norm="layer", aggr="add"):
        super().__init__()
        self.emb_dim, self.aggr = emb_dim, aggr
        self.mlp_msg = _mlp(2 * emb_dim + 1, emb_dim, activation, norm)
        self.mlp_pos = _mlp(emb_dim, emb_dim, activation, norm, tail=1)
        self.mlp_upd = _mlp(2 * emb_dim, emb_dim, activation, norm)

    def forward(self, h, pos, edge_index):
        m, shift = egnn_edge_message(self, h, pos, edge_index)
        idx = edge_index[1]
        # egnn_layer.py:77,79 (no dim_size: rows = idx.max()+1, SURVEY A.1 gotcha)
        m_aggr = scatter(m, idx, dim=-2, reduce=self.aggr)
        p_aggr = scatter(shift, idx, dim=-2, reduce="mean")
        # egnn_layer.py:84-85
        return self.mlp_upd(torch.cat([h, m_aggr], dim=-1)), pos + p_aggr


class MPNNLayer(nn.Module):
    """egnn_layer.py:92-155 (same scatter, no geometry)."""

    def __init__(self, emb_dim, activation="relu", norm="layer", aggr="add"):
        super().__init__()
        self.emb_dim, self.aggr = emb_dim, aggr
        self.mlp_msg = _mlp(2 * emb_dim, emb_dim, activation, norm)
        self.mlp_upd = _mlp(2 * emb_dim, emb_dim, activation, norm)

    def forward(self, h, edge_index):
        j, i = edge_index[0], edge_index[1]
        m = self.mlp_msg(torch.cat([h[i], h[j]], dim=-1))
        return self.mlp_upd(torch.cat([h, scatter(m, i, dim=-2, reduce=self.aggr)], dim=-1))


class EGNNModel(nn.Module):
    def __init__(self, num_layers=5, emb_dim=128, in_dim=1, out_dim=1, activation="relu", norm="layer",
                 aggr="sum", pool="sum", residual=True, equivariant_pred=False):
        super().__init__()
        self.equivariant_pred, self.residual = equivariant_pred, residual
        self.emb_in = nn.Embedding(in_dim, emb_dim)
        self.convs = nn.ModuleList([EGNNLayer(emb_dim, activation, norm, aggr) for _ in range(num_layers)])
        self.pool = _POOL[pool]
        if equivariant_pred:
            self.pred = nn.Linear(emb_dim + 3, out_dim)
        else:
            self.pred = nn.Sequential(nn.Linear(emb_dim, emb_dim), nn.ReLU(), nn.Linear(emb_dim, out_dim))

    def forward(self, batch):
        # egnn.py:66-87
        h, pos = self.emb_in(batch.atoms), batch.pos
        for conv in self.convs:
            dh, pos = conv(h, pos, batch.edge_index)
            h = h + dh if self.residual else dh
        feats = torch.cat([h, pos], dim=-1) if self.equivariant_pred else h
        return self.pred(self.pool(feats, batch.batch))


# =========================================================================== #
# SchNet   (models/schnet.py:9-80 over PyG 2.3.1 SchNet)
# =========================================================================== #
class SchNetModel(_PygSchNet):
    def __init__(self, hidden_channels=128, in_dim=1, out_dim=1, num_filters=128, num_layers=6,
                 num_gaussians=50, cutoff=10, max_num_neighbors=32, pool="sum"):
        super().__init__(hidden_channels, num_filters, num_layers, num_gaussians, cutoff,
                         interaction_graph=None, max_num_neighbors=max_num_neighbors, readout=pool)
        self.pool = _POOL[pool]
        self.lin2 = nn.Linear(hidden_channels // 2, out_dim)  # schnet.py:60 (default init)

    def forward(self, batch):
        # schnet.py:62-80
        h = self.embedding(batch.atoms)
        row, col = batch.edge_index
        d = (batch.pos[row] - batch.pos[col]).norm(dim=-1)
        rbf = self.distance_expansion(d)
        for block in self.interactions:
            h = h + block(h, batch.edge_index, d, rbf)
        out = self.pool(h, batch.batch)
        return self.lin2(self.act(self.lin1(out)))


# =========================================================================== #
# Radial basis   (models/mace_modules/radial.py:12-81, blocks.py:84-96)
# =========================================================================== #
def bessel_basis(x: torch.Tensor, r_max: float, num_basis: int) -> torch.Tensor:
    """radial.py:20-29,44-46: sqrt(2/r_max) * sin(n*pi*x/r_max) / x, n = 1..num_basis."""
    w = (math.pi / r_max) * torch.linspace(1.0, num_basis, num_basis, dtype=x.dtype, device=x.device)
    pref = torch.tensor(math.sqrt(2.0 / r_max), dtype=x.dtype, device=x.device)
    return pref * (torch.sin(w * x) / x)


def polynomial_cutoff(x: torch.Tensor, r_max: float, p: float) -> torch.Tensor:
    """radial.py:71-78."""
    p = torch.tensor(float(p), dtype=x.dtype, device=x.device)
    r = torch.tensor(float(r_max), dtype=x.dtype, device=x.device)
    u = x / r
    env = (1.0 - ((p + 1.0) * (p + 2.0) / 2.0) * torch.pow(u, p)
           + p * (p + 2.0) * torch.pow(u, p + 1) - (p * (p + 1.0) / 2) * torch.pow(u, p + 2))
    return env * (x < r)


class RadialEmbeddingBlock(nn.Module):
    def __init__(self, r_max: float, num_bessel: int, num_polynomial_cutoff: int):
        super().__init__()
        self.r_max, self.num_bessel, self.p = float(r_max), num_bessel, num_polynomial_cutoff
        self.out_dim = num_bessel

    def forward(self, edge_lengths):  # [E,1] -> [E,num_bessel]
        return bessel_basis(edge_lengths, self.r_max, self.num_bessel) * polynomial_cutoff(
            edge_lengths, self.r_max, self.p)


# =========================================================================== #
# Irreps helpers   (models/mace_modules/irreps_tools.py:64-97)
# =========================================================================== #
def irreps2gate(irreps):
    """irreps_tools.py:82-97: split into (scalars 0e, gates one-0e-per-gated-channel, gated l>0 or odd)."""
    irreps = o3.Irreps(irreps)
    scal = o3.Irreps([(m, ir) for m, ir in irreps if ir.l == 0 and ir.p == 1]).simplify()
    gated = o3.Irreps([(m, ir) for m, ir in irreps if not (ir.l == 0 and ir.p == 1)]).simplify()
    gates = o3.Irreps([(m, "0e") for m, _ in gated]).simplify() if gated.dim > 0 else o3.Irreps([])
    return scal, gates, gated


def reshape_irreps_fn(x: torch.Tensor, irreps) -> torch.Tensor:
    """irreps_tools.py:69-79: [N, sum mul*d] -> [N, mul, sum d]."""
    out, ix = [], 0
    for mul, ir in o3.Irreps(irreps):
        out.append(x[:, ix:ix + mul * ir.dim].reshape(x.shape[0], mul, ir.dim))
        ix += mul * ir.dim
    return torch.cat(out, dim=-1)


class reshape_irreps(nn.Module):
    def __init__(self, irreps):
        super().__init__()
        self.irreps = o3.Irreps(irreps)

    def forward(self, x):
        return reshape_irreps_fn(x, self.irreps)


# =========================================================================== #
# TFN conv layer   (models/layers/tfn_layer.py:8-93)
# =========================================================================== #
class TensorProductConvLayer(nn.Module):
    def __init__(self, in_irreps, out_irreps, sh_irreps, edge_feats_dim, mlp_dim, aggr="add",
                 batch_norm=False, gate=False):
        super().__init__()
        self.in_irreps, self.sh_irreps = o3.Irreps(str(in_irreps)), o3.Irreps(str(sh_irreps))
        out_irreps = o3.Irreps(str(out_irreps))
        self.aggr = aggr
        self.gate = None
        if gate:  # tfn_layer.py:45-63
            scal, gates, gated = irreps2gate(out_irreps)
            if gated.num_irreps == 0:
                self.gate = E3Activation(out_irreps, acts=[F.silu])
            else:
                self.gate = E3Gate(scal, [F.silu for _ in scal], gates, [torch.sigmoid for _ in gates], gated)
                out_irreps = self.gate.irreps_in
        self.out_irreps = out_irreps
        self.tp = o3.FullyConnectedTensorProduct(self.in_irreps, self.sh_irreps, out_irreps, shared_weights=False)
        self.fc = nn.Sequential(nn.Linear(edge_feats_dim, mlp_dim), nn.ReLU(), nn.Linear(mlp_dim, self.tp.weight_numel))
        self.batch_norm = E3BatchNorm(out_irreps) if batch_norm else None

    def forward(self, node_attr, edge_index, edge_sh, edge_feat):
        src, dst = edge_index  # tfn_layer.py:83: gather at ei[1], reduce at ei[0]
        msg = self.tp(node_attr[dst], edge_sh, self.fc(edge_feat))
        out = scatter(msg, src, dim=0, reduce=self.aggr)
        if self.gate is not None:
            out = self.gate(out)
        if self.batch_norm is not None:
            out = self.batch_norm(out)
        return out


def first_node_pooling(x, batch, size=None):
    """models/tfn.py:13-40: picks the first node of every graph except graph 0's
    (the comparison ``batch - shifted == 1`` is False at index 0, where shifted = -1
    ... it is True: 0 - (-1) == 1), i.e. the first node of every graph."""
    shifted = torch.cat([batch[-1:], batch[:-1]])
    shifted[0] = -1
    return x[(batch - shifted) == 1]


def edge_geometry(pos, edge_index, sh_mod, radial_mod):
    """models/tfn.py:171-175 == models/mace.py:170-174."""
    vec = pos[edge_index[0]] - pos[edge_index[1]]
    length = torch.linalg.norm(vec, dim=-1, keepdim=True)
    return sh_mod(vec), radial_mod(length)


class _EquivariantBase(nn.Module):
    def _common(self, r_max, num_bessel, num_polynomial_cutoff, max_ell, emb_dim, hidden_irreps, in_dim, out_dim,
                equivariant_pred):
        self.radial_embedding = RadialEmbeddingBlock(r_max, num_bessel, num_polynomial_cutoff)
        self.sh_irreps = o3.Irreps.spherical_harmonics(max_ell)
        self.spherical_harmonics = o3.SphericalHarmonics(self.sh_irreps, normalize=True, normalization="component")
        self.emb_in = nn.Embedding(in_dim, emb_dim)
        if hidden_irreps is None:
            hidden_irreps = (self.sh_irreps * emb_dim).sort()[0].simplify()
        self.hidden_irreps = o3.Irreps(str(hidden_irreps))
        if equivariant_pred:
            self.pred = nn.Linear(self.hidden_irreps.dim, out_dim)
        else:
            self.pred = nn.Sequential(nn.Linear(emb_dim, emb_dim), nn.ReLU(), nn.Linear(emb_dim, out_dim))


class TFNModel(_EquivariantBase):
    def __init__(self, r_max=10.0, num_bessel=8, num_polynomial_cutoff=5, max_ell=2, num_layers=5, emb_dim=64,
                 hidden_irreps=None, mlp_dim=256, in_dim=1, out_dim=1, aggr="sum", pool="first", gate=True,
                 batch_norm=False, residual=True, equivariant_pred=False):
        super().__init__()
        self.emb_dim, self.residual, self.equivariant_pred = emb_dim, residual, equivariant_pred
        self._common(r_max, num_bessel, num_polynomial_cutoff, max_ell, emb_dim, hidden_irreps, in_dim, out_dim,
                     equivariant_pred)
        ins = [o3.Irreps(f"{emb_dim}x0e")] + [self.hidden_irreps] * (num_layers - 1)
        self.convs = nn.ModuleList([
            TensorProductConvLayer(i, self.hidden_irreps, self.sh_irreps, num_bessel, mlp_dim, aggr, batch_norm, gate)
            for i in ins])
        self.pool = {**_POOL, "first": first_node_pooling}[pool]

    def forward(self, batch):
        # tfn.py:166-190
        h = self.emb_in(batch.atoms)
        sh, rbf = edge_geometry(batch.pos, batch.edge_index, self.spherical_harmonics, self.radial_embedding)
        for conv in self.convs:
            upd = conv(h, batch.edge_index, sh, rbf)
            h = upd + F.pad(h, (0, upd.shape[-1] - h.shape[-1])) if self.residual else upd
        out = self.pool(h, batch.batch)
        if not self.equivariant_pred:
            out = out[:, :self.emb_dim]
        return self.pred(out)


# =========================================================================== #
# MACE   (models/mace.py:16-190, blocks.py:99-135, symmetric_contraction.py, cg.py)
# =========================================================================== #
def _wigner_nj(irrepss, dtype):
    """cg.py:19-88 ('component' normalisation, no mid filter): iterated real-CG coupling.
    Returns a list of (ir_out, basis tensor [d_out, dim, ..., dim]) sorted by ir_out (stable)."""
    irrepss = [o3.Irreps(i) for i in irrepss]
    if len(irrepss) == 1:
        (irreps,) = irrepss
        eye, ret, i = torch.eye(irreps.dim, dtype=dtype), [], 0
        for mul, ir in irreps:
            for _ in range(mul):
                ret.append((ir, eye[i:i + ir.dim]))
                i += ir.dim
        return ret
    *left, right = irrepss
    ret = []
    for ir_left, C_left in _wigner_nj(left, dtype):
        i = 0
        for mul, ir in right:
            for ir_out in ir_left * ir:
                C = o3.wigner_3j(ir_out.l, ir_left.l, ir.l, dtype=dtype) * ir_out.dim ** 0.5
                C = torch.einsum("jk,ijl->ikl", C_left.flatten(1), C)
                C = C.reshape(ir_out.dim, *(x.dim for x in left), ir.dim)
                for u in range(mul):
                    E = torch.zeros(ir_out.dim, *(x.dim for x in left), right.dim, dtype=dtype)
                    E[..., i + u * ir.dim:i + (u + 1) * ir.dim] = C
                    ret.append((ir_out, E))
            i += mul * ir.dim
    return sorted(ret, key=lambda t: t[0])


def u_matrix_real(irreps_in, ir_out, correlation: int, dtype=None) -> torch.Tensor:
    """cg.py:91-133 for a single output irrep: stack of all coupling paths into
    ``ir_out``, shape [d_out, dim^nu, k] squeezed when d_out == 1."""
    dtype = dtype or torch.get_default_dtype()
    ir_out = o3.Irrep(ir_out)
    basis = [B for ir, B in _wigner_nj([o3.Irreps(irreps_in)] * correlation, dtype) if ir == ir_out]
    return torch.stack([b.squeeze() for b in basis], dim=-1)


class Contraction(nn.Module):
    """symmetric_contraction.py:88-188, element_dependent=False branch."""

    def __init__(self, irreps_in, irrep_out, correlation: int):
        super().__init__()
        irreps_in = o3.Irreps(irreps_in)
        self.num_features = irreps_in.count((0, 1))
        coupling = o3.Irreps([ir for _, ir in irreps_in])
        self.correlation = correlation
        self.weights = nn.ParameterDict()
        for nu in range(1, correlation + 1):
            U = u_matrix_real(coupling, irrep_out, nu)
            self.register_buffer(f"U_matrix_{nu}", U)
            k = U.shape[-1]
            self.weights[str(nu)] = nn.Parameter(torch.randn(k, self.num_features) / k)

    def forward(self, x, y=None):
        nu = self.correlation
        out = contract("...ik,kc,bci->bc...", getattr(self, f"U_matrix_{nu}"), self.weights[str(nu)], x)
        for nu in range(self.correlation - 1, 0, -1):
            c = contract("...k,kc->c...", getattr(self, f"U_matrix_{nu}"), self.weights[str(nu)]) + out
            out = contract("bc...i,bci->bc...", c, x)
        return out.reshape(out.shape[0], -1)


class SymmetricContraction(nn.Module):
    """symmetric_contraction.py:21-85."""

    def __init__(self, irreps_in, irreps_out, correlation, element_dependent=False, num_elements=None, **_):
        super().__init__()
        assert not element_dependent, "reference uses element_dependent=False (models/mace.py:119)"
        self.irreps_in, self.irreps_out = o3.Irreps(str(irreps_in)), o3.Irreps(str(irreps_out))
        self.contractions = nn.ModuleDict({
            str(mi): Contraction(self.irreps_in, mi.ir, correlation) for mi in self.irreps_out})

    def forward(self, x, y=None):
        return torch.cat([self.contractions[str(mi)](x, y) for mi in self.irreps_out], dim=-1)


class EquivariantProductBasisBlock(nn.Module):
    """blocks.py:99-135."""

    def __init__(self, node_feats_irreps, target_irreps, correlation, element_dependent=True, use_sc=True,
                 batch_norm=False, num_elements=None):
        super().__init__()
        self.use_sc = use_sc
        self.symmetric_contractions = SymmetricContraction(node_feats_irreps, target_irreps, correlation,
                                                           element_dependent=element_dependent,
                                                           num_elements=num_elements)
        self.linear = o3.Linear(target_irreps, target_irreps)
        self.batch_norm = E3BatchNorm(target_irreps) if batch_norm else None

    def forward(self, node_feats, sc, node_attrs=None):
        out = self.linear(self.symmetric_contractions(node_feats, node_attrs))
        if self.batch_norm is not None:
            out = self.batch_norm(out)
        return out + sc if self.use_sc else out


class MACEModel(_EquivariantBase):
    def __init__(self, r_max=10.0, num_bessel=8, num_polynomial_cutoff=5, max_ell=2, correlation=3, num_layers=5,
                 emb_dim=64, hidden_irreps=None, mlp_dim=256, in_dim=1, out_dim=1, aggr="sum", pool="sum",
                 batch_norm=True, residual=True, equivariant_pred=False):
        super().__init__()
        self.emb_dim, self.residual, self.equivariant_pred = emb_dim, residual, equivariant_pred
        self._common(r_max, num_bessel, num_polynomial_cutoff, max_ell, emb_dim, hidden_irreps, in_dim, out_dim,
                     equivariant_pred)
        ins = [o3.Irreps(f"{emb_dim}x0e")] + [self.hidden_irreps] * (num_layers - 1)
        self.convs = nn.ModuleList([
            TensorProductConvLayer(i, self.hidden_irreps, self.sh_irreps, num_bessel, mlp_dim, aggr, batch_norm, False)
            for i in ins])
        self.reshapes = nn.ModuleList([reshape_irreps(self.hidden_irreps) for _ in ins])
        self.prods = nn.ModuleList([
            EquivariantProductBasisBlock(self.hidden_irreps, self.hidden_irreps, correlation, element_dependent=False,
                                         num_elements=in_dim, use_sc=residual) for _ in ins])
        self.pool = _POOL[pool]

    def forward(self, batch):
        # mace.py:165-190
        h = self.emb_in(batch.atoms)
        sh, rbf = edge_geometry(batch.pos, batch.edge_index, self.spherical_harmonics, self.radial_embedding)
        for conv, reshape, prod in zip(self.convs, self.reshapes, self.prods):
            upd = conv(h, batch.edge_index, sh, rbf)
            sc = F.pad(h, (0, upd.shape[-1] - h.shape[-1]))
            h = prod(reshape(upd), sc, None)
        out = self.pool(h, batch.batch)
        if not self.equivariant_pred:
            out = out[:, :self.emb_dim]
        return self.pred(out)


# =========================================================================== #
# Fixtures from the reference experiments
# =========================================================================== #

# ----------------------------------------------------------------------------------------------------------------------
# ACEsuit-style interaction blocks -- models/mace_modules/blocks.py:136-530
# ----------------------------------------------------------------------------------------------------------------------
def tp_out_irreps_with_instructions(irreps1, irreps2, target_irreps):
    """irreps_tools.py:14-44: one 'uvu' instruction per (feature block, edge block, admissible l_out in the target); the
    output blocks are then sorted by irrep and the instruction's output index follows the permutation."""
    irreps1, irreps2, target_irreps = o3.Irreps(irreps1), o3.Irreps(irreps2), o3.Irreps(target_irreps)
    blocks, ins = [], []
    for i, (mul, ir_in) in enumerate(irreps1):
        for j, (_, ir_edge) in enumerate(irreps2):
            for ir_out in ir_in * ir_edge:
                if ir_out in target_irreps:
                    ins.append((i, j, len(blocks), "uvu", True))
                    blocks.append((mul, ir_out))
    mid, permut, _ = o3.Irreps(blocks).sort()
    return mid, [(a, b, permut[c], mode, train) for a, b, c, mode, train in ins]


def linear_out_irreps(irreps, target_irreps):
    """irreps_tools.py:47-62."""
    out = []
    for _, ir_in in o3.Irreps(irreps):
        hit = [(mul, ir_out) for mul, ir_out in o3.Irreps(target_irreps) if ir_in == ir_out]
        if not hit:
            raise RuntimeError(f"{ir_in} not in {target_irreps}")
        out.append(hit[0])
    return o3.Irreps(out)


class TensorProductWeightsBlock(nn.Module):
    """blocks.py:177-203."""

    def __init__(self, num_elements, num_edge_feats, num_feats_out):
        super().__init__()
        w = torch.empty(num_elements, num_edge_feats, num_feats_out)
        nn.init.xavier_uniform_(w)
        self.weights = nn.Parameter(w)

    def forward(self, sender_or_receiver_node_attrs, edge_feats):
        return torch.einsum("be,ba,aek->bk", edge_feats, sender_or_receiver_node_attrs, self.weights)


class InteractionBlock(nn.Module):
    """The five concrete blocks of blocks.py:206-530 differ in three choices, given here as ``kind``:
    ================================================  ==========  ==================  =========================================
    kind (reference class, blocks.py line)            radial net  irreps_out          skip_tp / return
    ================================================  ==========  ==================  =========================================
    residual_element (ResidualElementDependent, 206)  per element linear_out_irreps   msg + skip_tp(node_feats, attrs)
    agnostic_nonlinear (AgnosticNonlinear, 276)       MLP         linear_out_irreps   skip_tp(msg, attrs)
    agnostic_residual_nonlinear (..., 330)            MLP         linear_out_irreps   msg + skip_tp(node_feats, attrs)
    real_agnostic (RealAgnostic, 396)                 MLP         target_irreps       (reshape(skip_tp(msg, attrs)), None)
    real_agnostic_residual (..., 462)                 MLP         target_irreps       (reshape(msg), skip_tp(node_feats, attrs) -> hidden)
    ================================================  ==========  ==================  =========================================
    Common body (e.g. :440-455): linear_up -> conv_tp(node_feats[sender], edge_attrs, w(edge_feats)) -> scatter_sum over the
    receivers with dim_size = N -> linear -> / avg_num_neighbors."""

    def __init__(self, kind, node_attrs_irreps, node_feats_irreps, edge_attrs_irreps, edge_feats_irreps, target_irreps,
                 hidden_irreps, avg_num_neighbors):
        super().__init__()
        from .thirdparty.e3nn_nn import FullyConnectedNet
        I = o3.Irreps
        self.kind, self.avg_num_neighbors = kind, avg_num_neighbors
        node_attrs_irreps, node_feats_irreps, target_irreps = I(node_attrs_irreps), I(node_feats_irreps), I(target_irreps)
        self.linear_up = o3.Linear(node_feats_irreps, node_feats_irreps, internal_weights=True, shared_weights=True)
        irreps_mid, instructions = tp_out_irreps_with_instructions(node_feats_irreps, edge_attrs_irreps, target_irreps)
        self.conv_tp = o3.TensorProduct(node_feats_irreps, I(edge_attrs_irreps), irreps_mid, instructions=instructions,
                                        shared_weights=False, internal_weights=False)
        n_edge = I(edge_feats_irreps).num_irreps
        if kind == "residual_element":
            self.conv_tp_weights = TensorProductWeightsBlock(node_attrs_irreps.num_irreps, n_edge, self.conv_tp.weight_numel)
        else:
            self.conv_tp_weights = FullyConnectedNet([n_edge] + 3 * [64] + [self.conv_tp.weight_numel], F.silu)
        irreps_mid = irreps_mid.simplify()
        self.irreps_out = target_irreps if kind.startswith("real") else linear_out_irreps(irreps_mid, target_irreps).simplify()
        self.linear = o3.Linear(irreps_mid, self.irreps_out, internal_weights=True, shared_weights=True)
        skip_in = self.irreps_out if kind in ("agnostic_nonlinear", "real_agnostic") else node_feats_irreps
        skip_out = I(hidden_irreps) if kind == "real_agnostic_residual" else self.irreps_out
        self.skip_tp = o3.FullyConnectedTensorProduct(skip_in, node_attrs_irreps, skip_out)

    def forward(self, node_attrs, node_feats, edge_attrs, edge_feats, edge_index):
        sender, receiver = edge_index
        n = node_feats.shape[0]
        sc = None
        if self.kind in ("residual_element", "agnostic_residual_nonlinear", "real_agnostic_residual"):
            sc = self.skip_tp(node_feats, node_attrs)
        node_feats = self.linear_up(node_feats)
        if self.kind == "residual_element":
            w = self.conv_tp_weights(node_attrs[sender], edge_feats)
        else:
            w = self.conv_tp_weights(edge_feats)
        mji = self.conv_tp(node_feats[sender], edge_attrs, w)
        message = scatter(mji, receiver, dim=0, dim_size=n, reduce="sum")
        message = self.linear(message) / self.avg_num_neighbors
        if self.kind == "real_agnostic":
            return reshape_irreps_fn(self.skip_tp(message, node_attrs), self.irreps_out), None
        if self.kind == "real_agnostic_residual":
            return reshape_irreps_fn(message, self.irreps_out), sc
        if self.kind == "agnostic_nonlinear":
            return self.skip_tp(message, node_attrs)
        return message + sc



# ----------------------------------------------------------------------------------------------------------------------
# GVP-GNN -- models/layers/gvp_layer.py:101-438, models/gvpgnn.py:9-126
# ----------------------------------------------------------------------------------------------------------------------
def _clamped_norm(x, axis=-1, keepdims=False, eps=1e-8, sqrt=True):
    """gvp_layer.py:66-73."""
    out = torch.clamp(torch.sum(torch.square(x), axis, keepdims), min=eps)
    return torch.sqrt(out) if sqrt else out


class GVP(nn.Module):
    """gvp_layer.py:101-170.  Parameter names (wh, ws, wv, wsv, dummy_param) are the reference's."""

    def __init__(self, in_dims, out_dims, h_dim=None, activations=(F.relu, torch.sigmoid), vector_gate=True):
        super().__init__()
        (self.si, self.vi), (self.so, self.vo) = in_dims, out_dims
        self.vector_gate, (self.scalar_act, self.vector_act) = vector_gate, activations
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
        self.dummy_param = nn.Parameter(torch.empty(0))

    def forward(self, x):
        if not self.vi:                                            # :161-164
            s = self.ws(x)
            v = torch.zeros(s.shape[0], self.vo, 3) if self.vo else None
        else:                                                      # :146-160
            s, v = x
            vh = torch.einsum("hc,ncx->nxh", self.wh.weight, v)
            s = self.ws(torch.cat([s, _clamped_norm(vh, axis=-2)], -1))
            if self.vo:
                v = torch.einsum("oh,nxh->nox", self.wv.weight, vh)
                if self.vector_gate:
                    v = v * torch.sigmoid(self.wsv(self.vector_act(s) if self.vector_act else s)).unsqueeze(-1)
                elif self.vector_act:
                    v = v * self.vector_act(_clamped_norm(v, axis=-1, keepdims=True))
        if self.scalar_act:
            s = self.scalar_act(s)
        return (s, v) if self.vo else s


class GVPLayerNorm(nn.Module):
    """gvp_layer.py:221-243."""

    def __init__(self, dims):
        super().__init__()
        self.s, self.v = dims
        self.scalar_norm = nn.LayerNorm(self.s)

    def forward(self, x):
        if not self.v:
            return self.scalar_norm(x)
        s, v = x
        vn = torch.sqrt(torch.mean(_clamped_norm(v, axis=-1, keepdims=True, sqrt=False), dim=-2, keepdim=True))
        return self.scalar_norm(s), v / vn


class _GVPDropout(nn.Module):
    """gvp_layer.py:173-218 in eval mode (identity); keeps the reference's parameter key."""

    def __init__(self, drop_rate):
        super().__init__()
        self.sdropout = nn.Dropout(drop_rate)
        self.vdropout = nn.Module()
        self.vdropout.dummy_param = nn.Parameter(torch.empty(0))

    def forward(self, x):
        assert not self.training, "the oracle restates the deterministic (eval) path of the GVP dropout"
        return x


class GVPConvLayer(nn.Module):
    """gvp_layer.py:246-438 (GVPConv inlined as `conv`): messages over cat[(s_j, V_j), edge, (s_i, V_i)] (:319-324), scatter mean
    over edge_index[1] with dim_size = N (PyG propagate), then the two residual + LayerNorm steps (:432-435)."""

    def __init__(self, node_dims, edge_dims, n_message=3, n_feedforward=2, drop_rate=0.1, activations=(F.relu, torch.sigmoid),
                 vector_gate=True, residual=True):
        super().__init__()
        mk = lambda i, o, **kw: GVP(i, o, **{"activations": activations, "vector_gate": vector_gate, **kw})
        (si, vi), (se, ve) = node_dims, edge_dims
        cat_dims = (2 * si + se, 2 * vi + ve)
        self.conv = nn.Module()
        if n_message == 1:
            msg = [mk(cat_dims, node_dims, activations=(None, None))]
        else:
            msg = [mk(cat_dims, node_dims)] + [mk(node_dims, node_dims) for _ in range(n_message - 2)] + [mk(node_dims, node_dims, activations=(None, None))]
        self.conv.message_func = nn.Sequential(*msg)
        self.norm = nn.ModuleList([GVPLayerNorm(node_dims) for _ in range(2)])
        self.dropout = nn.ModuleList([_GVPDropout(drop_rate) for _ in range(2)])
        if n_feedforward == 1:
            ff = [mk(node_dims, node_dims, activations=(None, None))]
        else:
            hid = (4 * si, 2 * vi)
            ff = [mk(node_dims, hid)] + [mk(hid, hid) for _ in range(n_feedforward - 2)] + [mk(hid, node_dims, activations=(None, None))]
        self.ff_func = nn.Sequential(*ff)
        self.residual, self.vo = residual, vi

    def forward(self, x, edge_index, edge_attr):
        s, v = x
        j, i = edge_index[0], edge_index[1]
        ms, mv = self.conv.message_func((torch.cat([s[j], edge_attr[0], s[i]], -1), torch.cat([v[j], edge_attr[1], v[i]], -2)))
        agg = scatter(torch.cat([ms, mv.reshape(mv.shape[0], -1)], -1), i, dim=0, dim_size=s.shape[0], reduce="mean")
        dh = (agg[:, :-3 * self.vo], agg[:, -3 * self.vo:].reshape(-1, self.vo, 3))
        x = self.norm[0]((s + dh[0], v + dh[1])) if self.residual else dh
        dh = self.ff_func(x)
        return self.norm[1]((x[0] + dh[0], x[1] + dh[1])) if self.residual else dh


class GVPGNNModel(nn.Module):
    """models/gvpgnn.py:9-126."""

    def __init__(self, r_max=10.0, num_bessel=8, num_polynomial_cutoff=5, num_layers=5, in_dim=1, out_dim=1, s_dim=128, v_dim=16,
                 s_dim_edge=32, v_dim_edge=1, pool="sum", residual=True, equivariant_pred=False):
        super().__init__()
        self.s_dim, self.equivariant_pred = s_dim, equivariant_pred
        self.emb_in = nn.Embedding(in_dim, s_dim)
        self.W_v = nn.Sequential(GVPLayerNorm((s_dim, 0)), GVP((s_dim, 0), (s_dim, v_dim), activations=(None, None)))
        self.radial_embedding = RadialEmbeddingBlock(r_max, num_bessel, num_polynomial_cutoff)
        self.W_e = nn.Sequential(GVPLayerNorm((num_bessel, 1)), GVP((num_bessel, 1), (s_dim_edge, v_dim_edge), activations=(None, None)))
        self.layers = nn.ModuleList(GVPConvLayer((s_dim, v_dim), (s_dim_edge, v_dim_edge), activations=(F.relu, None), residual=residual)
                                    for _ in range(num_layers))
        self.pool = _POOL[pool]
        self.pred = (nn.Linear(s_dim + v_dim * 3, out_dim) if equivariant_pred else
                     nn.Sequential(nn.Linear(s_dim, s_dim), nn.ReLU(), nn.Linear(s_dim, out_dim)))

    def forward(self, batch):
        vec = batch.pos[batch.edge_index[0]] - batch.pos[batch.edge_index[1]]        # :102-103
        length = torch.linalg.norm(vec, dim=-1, keepdim=True)
        h_v = self.W_v(self.emb_in(batch.atoms))
        h_e = self.W_e((self.radial_embedding(length), torch.nan_to_num(vec / length).unsqueeze(-2)))
        for layer in self.layers:
            h_v = layer(h_v, batch.edge_index, h_e)
        out = self.pool(torch.cat([h_v[0], h_v[1].reshape(h_v[1].shape[0], -1)], -1), batch.batch)
        return self.pred(out if self.equivariant_pred else out[:, :self.s_dim])


def create_kchains(k: int):
    """experiments/kchains.ipynb:71-107: two (k+2)-node chains whose k centre nodes sit at
    (0, 5i, 0); the first end node is at (-4,-3,0) in graph 0 and (+4,-3,0) in graph 1;
    positions are centred on their mean; edges are the undirected chain."""
    from .thirdparty.pyg import Data, to_undirected
    assert k >= 2
    n = k + 2
    chain = to_undirected(torch.stack([torch.arange(0, n - 1), torch.arange(1, n)]).long())
    graphs = []
    for label, x0 in enumerate((-4.0, 4.0)):
        pos = torch.tensor([[x0, -3.0, 0.0]] + [[0.0, 5.0 * i, 0.0] for i in range(k)]
                           + [[4.0, 5.0 * (k - 1) + 3.0, 0.0]], dtype=torch.float32)
        pos = pos - pos.mean(dim=0)
        graphs.append(Data(atoms=torch.zeros(n, dtype=torch.long), edge_index=chain.clone(), pos=pos,
                           y=torch.tensor([label])))
    return graphs
