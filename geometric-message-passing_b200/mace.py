"""MACE on the fused kernels: drop-in for ``models/mace.py`` and the pieces of ``models/mace_modules`` it uses
(``reshape_irreps``, ``SymmetricContraction`` / ``Contraction``, ``EquivariantProductBasisBlock``, e3nn ``o3.Linear``).

``state_dict`` keys equal the reference's: ``symmetric_contractions.contractions.<mul>x<ir>.weights.{1,2,3}`` (+ the
``U_matrix_*`` buffers), ``linear.weight``.  The symmetric contraction runs as one kernel per direction
(csrc/symcontract.cu) over per-channel monomial coefficients that a tiny differentiable einsum builds from the
weights; the interaction is the same fused tensor-product convolution TFN uses (gate=False, e3nn BatchNorm).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Union

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from .embed import embedding_lookup
from . import _lib
from ._lib import call, ptr
from .irreps import Irrep, Irreps, monomials, symmetrise_u, u_matrix_real
from .irreps import hidden_irreps as default_hidden_irreps
from .schnet import global_add_pool, global_mean_pool
from .tfn import BatchNorm, RadialEmbeddingBlock, SphericalHarmonics, TensorProductConvLayer, edge_geometry


class reshape_irreps(nn.Module):
    """models/mace_modules/irreps_tools.py:64-79: [N, sum mul*d] -> [N, mul, sum d]."""

    def __init__(self, irreps):
        super().__init__()
        self.irreps = Irreps(str(irreps))

    def forward(self, tensor: torch.Tensor) -> torch.Tensor:
        ix, out = 0, []
        batch = tensor.shape[0]
        for mul, ir in self.irreps:
            out.append(tensor[:, ix:ix + mul * ir.dim].reshape(batch, mul, ir.dim))
            ix += mul * ir.dim
        return torch.cat(out, dim=-1)


class _SymContractFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, coef, mono, out_map, out_len):
        x, coef = x.contiguous(), coef.contiguous()
        N, C, D = x.shape
        K, M = coef.shape[1], coef.shape[2]
        out = torch.empty(N, out_len, dtype=x.dtype, device=x.device)
        call("gmp_symcontract_fwd", ptr(x), ptr(coef), ptr(mono), ptr(out_map), N, C, D, K, M, ptr(out), out_len)
        ctx.save_for_backward(x, coef, mono, out_map)
        ctx.out_len = out_len
        return out

    @staticmethod
    def backward(ctx, g):
        x, coef, mono, out_map = ctx.saved_tensors
        N, C, D = x.shape
        K, M = coef.shape[1], coef.shape[2]
        g = g.contiguous()
        nparts = _lib.lib().gmp_symcontract_bwd_num_parts(N)
        dx = torch.empty_like(x)
        parts = torch.empty(nparts, C * K * M, dtype=x.dtype, device=x.device)
        call("gmp_symcontract_bwd", ptr(x), ptr(coef), ptr(mono), ptr(out_map), N, C, D, K, M, ptr(g), ctx.out_len, ptr(dx),
             ptr(parts))
        dcoef = torch.empty(C * K * M, dtype=x.dtype, device=x.device)
        call("gmp_reduce_partials_f32", ptr(parts), nparts, C * K * M, ptr(dcoef))
        return dx, dcoef.view(C, K, M), None, None, None


class Contraction(nn.Module):
    """Parameter holder with the reference's layout (symmetric_contraction.py:88-148, element_dependent=False)."""

    def __init__(self, irreps_in: Irreps, irrep_out: Irrep, correlation: int, num_features: int):
        super().__init__()
        self.correlation, self.num_features = correlation, num_features
        coupling = Irreps("+".join(f"1x{ir}" for _, ir in irreps_in))
        self.weights = nn.ParameterDict()
        D = coupling.dim
        for nu in range(1, correlation + 1):
            U = u_matrix_real(coupling, irrep_out, nu)  # [d_out, D^nu, k]
            k = U.shape[-1]
            self.register_buffer(f"U_matrix_{nu}", torch.from_numpy(U.squeeze(0) if U.shape[0] == 1 else U).float())
            # symmetric-monomial form used by the kernel (constructor-time constant, not part of the state_dict)
            self.register_buffer(f"_Usym_{nu}", torch.from_numpy(symmetrise_u(U, nu, D)).float(), persistent=False)
            self.weights[str(nu)] = nn.Parameter(torch.randn(k, num_features) / k)

    def coefficients(self) -> torch.Tensor:
        """[C, d_out, M] over monomials ordered degree 1, 2, ..., correlation."""
        return torch.cat([torch.einsum("kme,ec->ckm", getattr(self, f"_Usym_{nu}"), self.weights[str(nu)])
                          for nu in range(1, self.correlation + 1)], dim=2)


class SymmetricContraction(nn.Module):
    """models/mace_modules/symmetric_contraction.py:21-85 (shared weights, not element dependent)."""

    def __init__(self, irreps_in, irreps_out, correlation: Union[int, Dict[str, int]], irrep_normalization="component",
                 path_normalization="element", internal_weights=None, shared_weights=None, element_dependent=None,
                 num_elements=None):
        super().__init__()
        if element_dependent:
            raise NotImplementedError("element_dependent=True is unused by the reference models (models/mace.py:119)")
        if not isinstance(correlation, int):
            raise NotImplementedError("per-irrep correlation dicts are not built")
        if not 1 <= correlation <= 3:
            raise NotImplementedError("correlation order 1..3 (monomial kernel covers degree <= 3)")
        self.irreps_in, self.irreps_out = Irreps(str(irreps_in)), Irreps(str(irreps_out))
        self.correlation = correlation
        C = sum(m for m, ir in self.irreps_in if ir.l == 0 and ir.p == 1)
        self.num_features = C
        self.contractions = nn.ModuleDict({f"{m}x{ir}": Contraction(self.irreps_in, ir, correlation, C) for m, ir in self.irreps_out})
        D = sum(ir.dim for _, ir in self.irreps_in)
        self.D = D
        omap, off = [], 0
        for m, ir in self.irreps_out:
            assert m == C, "symmetric contraction keeps the channel count"
            omap += [(off, ir.dim, k) for k in range(ir.dim)]
            off += ir.dim
        self.register_buffer("_mono", torch.from_numpy(monomials(D, correlation)), persistent=False)
        self.register_buffer("_out_map", torch.tensor(omap, dtype=torch.int32), persistent=False)

    def forward(self, x: torch.Tensor, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        coef = torch.cat([self.contractions[f"{m}x{ir}"].coefficients() for m, ir in self.irreps_out], dim=1)  # [C, K, M]
        return _SymContractFn.apply(x, coef, self._mono, self._out_map, self.irreps_out.dim)


class EquivariantLinear(nn.Module):
    """e3nn o3.Linear(irreps, irreps) with internal shared weights, no bias: per irrep block a
    [mul_in, mul_out] channel mix scaled by 1/sqrt(fan_in); weight is one flat vector in instruction order."""

    def __init__(self, irreps_in, irreps_out):
        super().__init__()
        self.irreps_in, self.irreps_out = Irreps(str(irreps_in)), Irreps(str(irreps_out))
        self.instructions = [(i, o) for i, (_, a) in enumerate(self.irreps_in) for o, (_, b) in enumerate(self.irreps_out) if a == b]
        self.fan_in = {}
        for i, o in self.instructions:
            self.fan_in[o] = self.fan_in.get(o, 0) + self.irreps_in[i][0]
        self.weight_numel = sum(self.irreps_in[i][0] * self.irreps_out[o][0] for i, o in self.instructions)
        self.weight = nn.Parameter(torch.randn(self.weight_numel))

    def forward(self, x):
        N = x.shape[0]
        offs_in = self.irreps_in.offsets()
        outs = [None] * len(self.irreps_out)
        off = 0
        for i, o in self.instructions:
            mi, ir = self.irreps_in[i]
            mo, _ = self.irreps_out[o]
            W = self.weight[off:off + mi * mo].view(mi, mo)
            off += mi * mo
            xi = x[:, offs_in[i]:offs_in[i] + mi * ir.dim].reshape(N, mi, ir.dim)
            r = torch.einsum("uw,nud->nwd", W, xi).reshape(N, mo * ir.dim) * (1.0 / math.sqrt(self.fan_in[o]))
            outs[o] = r if outs[o] is None else outs[o] + r
        for k, (mo, ir) in enumerate(self.irreps_out):
            if outs[k] is None:
                outs[k] = x.new_zeros(N, mo * ir.dim)
        return torch.cat(outs, dim=-1)


class EquivariantProductBasisBlock(nn.Module):
    """models/mace_modules/blocks.py:99-135."""

    def __init__(self, node_feats_irreps, target_irreps, correlation, element_dependent: bool = True, use_sc: bool = True,
                 batch_norm: bool = False, num_elements: Optional[int] = None):
        super().__init__()
        self.use_sc = use_sc
        self.symmetric_contractions = SymmetricContraction(node_feats_irreps, target_irreps, correlation,
                                                           element_dependent=element_dependent, num_elements=num_elements)
        self.linear = EquivariantLinear(target_irreps, target_irreps)
        self.batch_norm = BatchNorm(target_irreps) if batch_norm else None

    def forward(self, node_feats, sc, node_attrs=None):
        node_feats = self.symmetric_contractions(node_feats, node_attrs)
        out = self.linear(node_feats)
        if self.batch_norm is not None:
            out = self.batch_norm(out)
        if self.use_sc:
            out = out + sc
        return out


class MACEModel(nn.Module):
    """models/mace.py:16-190."""

    def __init__(self, r_max: float = 10.0, num_bessel: int = 8, num_polynomial_cutoff: int = 5, max_ell: int = 2,
                 correlation: int = 3, num_layers: int = 5, emb_dim: int = 64, hidden_irreps=None, mlp_dim: int = 256,
                 in_dim: int = 1, out_dim: int = 1, aggr: str = "sum", pool: str = "sum", batch_norm: bool = True,
                 residual: bool = True, equivariant_pred: bool = False, precision: str = "fp32"):
        super().__init__()
        self.r_max, self.max_ell, self.num_layers, self.emb_dim, self.mlp_dim = r_max, max_ell, num_layers, emb_dim, mlp_dim
        self.residual, self.batch_norm, self.equivariant_pred = residual, batch_norm, equivariant_pred
        self.radial_embedding = RadialEmbeddingBlock(r_max, num_bessel, num_polynomial_cutoff)
        sh_irreps = Irreps.spherical_harmonics(max_ell)
        self.spherical_harmonics = SphericalHarmonics(max_ell)
        self.emb_in = torch.nn.Embedding(in_dim, emb_dim)
        hidden = default_hidden_irreps(max_ell, emb_dim) if hidden_irreps is None else Irreps(str(hidden_irreps))
        self.hidden_irreps = hidden
        self.convs, self.prods, self.reshapes = torch.nn.ModuleList(), torch.nn.ModuleList(), torch.nn.ModuleList()
        for layer in range(num_layers):
            in_irreps = Irreps(f"{emb_dim}x0e") if layer == 0 else hidden
            self.convs.append(TensorProductConvLayer(in_irreps, hidden, sh_irreps, self.radial_embedding.out_dim, mlp_dim,
                                                     aggr, batch_norm, False, precision))
            self.reshapes.append(reshape_irreps(hidden))
            self.prods.append(EquivariantProductBasisBlock(hidden, hidden, correlation, element_dependent=False,
                                                           num_elements=in_dim, use_sc=residual))
        self.pool = {"mean": global_mean_pool, "sum": global_add_pool}[pool]
        if equivariant_pred:
            self.pred = torch.nn.Linear(hidden.dim, out_dim)
        else:
            self.pred = torch.nn.Sequential(torch.nn.Linear(emb_dim, emb_dim), torch.nn.ReLU(), torch.nn.Linear(emb_dim, out_dim))

    def forward(self, batch):
        h = embedding_lookup(self.emb_in, batch.atoms)
        edge_sh, edge_feats = edge_geometry(batch.pos, batch.edge_index, self.max_ell, self.radial_embedding)
        for conv, reshape, prod in zip(self.convs, self.reshapes, self.prods):
            h_update = conv(h, batch.edge_index, edge_sh, edge_feats)
            sc = F.pad(h, (0, h_update.shape[-1] - h.shape[-1]))
            h = prod(reshape(h_update), sc, None)
        out = self.pool(h, batch.batch, getattr(batch, "num_graphs", None))
        if not self.equivariant_pred:
            out = out[:, :self.emb_dim]
        return self.pred(out)
