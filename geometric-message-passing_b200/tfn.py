"""TFN / MACE tensor-product convolution on the fused kernels: drop-in for
``models/layers/tfn_layer.py`` (``TensorProductConvLayer``), ``models/tfn.py`` (``TFNModel``) and the e3nn
pieces they instantiate (``FullyConnectedTensorProduct`` with per-edge weights, ``Gate``, ``BatchNorm``,
``SphericalHarmonics``), plus ``RadialEmbeddingBlock`` (models/mace_modules/blocks.py:84-96).

``fc(edge_feat)`` -- the [E, weight_numel] tensor that dominates the reference layer -- is generated slice by
slice in shared memory and consumed in place (csrc/tpconv.cu).  Irreps arguments may be e3nn ``Irreps`` objects
or their strings.  ``state_dict`` keys: ``fc.0.*``, ``fc.2.*``, ``batch_norm.{weight,bias,running_mean,running_var}``.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from .embed import embedding_lookup
from . import _lib
from ._lib import GmpError, call, ptr
from .graph import Graph, get_graph
from .irreps import NORM2MOM, Irreps, TPPath, fctp_paths, gate_split, wigner_3j
from .irreps import hidden_irreps as default_hidden_irreps
from .schnet import global_add_pool, global_mean_pool

_PREC = {"fp32": _lib.FP32_STRICT, "bf16": _lib.BF16_TC}
_SLICE = 128


# ------------------------------------------------------------------------------------------------
# node-side equivariant non-linearities (Gate / Activation: fused elementwise kernels, csrc/gate.cu)
# ------------------------------------------------------------------------------------------------
class _GateFn(torch.autograd.Function):
    """[scalars | gates | gated] -> [silu(s) c | gated * sigmoid(gate) c]: one elementwise kernel each way (csrc/gate.cu)."""

    @staticmethod
    def forward(ctx, x, expand, gstart, gdim, ns: int, ng: int, nv: int):
        x = x.contiguous()
        out = torch.empty(x.shape[0], ns + nv, dtype=x.dtype, device=x.device)
        call("gmp_gate_fwd", ptr(x), ptr(expand), x.shape[0], ns, ng, nv, NORM2MOM["silu"], NORM2MOM["sigmoid"], ptr(out))
        ctx.save_for_backward(x, expand, gstart, gdim)
        ctx.dims = (ns, ng, nv)
        return out

    @staticmethod
    def backward(ctx, g):
        x, expand, gstart, gdim = ctx.saved_tensors
        ns, ng, nv = ctx.dims
        dx = torch.empty_like(x)
        call("gmp_gate_bwd", ptr(x), ptr(g.contiguous()), ptr(expand), ptr(gstart), ptr(gdim), x.shape[0], ns, ng, nv,
             NORM2MOM["silu"], NORM2MOM["sigmoid"], ptr(dx))
        return dx, None, None, None, None, None, None


class Gate(nn.Module):
    """e3nn.nn.Gate with silu scalars / sigmoid gates (models/layers/tfn_layer.py:45-63): input laid out as
    [scalars | gates | gated]; out = [silu(s) * c_silu | gated_u * sigmoid(gate_u) * c_sigmoid].  Fused: one kernel
    forward, one backward (fp32, CUDA, 2-D input); other inputs take the equivalent torch expression."""

    def __init__(self, irreps_scalars, irreps_gates, irreps_gated):
        super().__init__()
        self.irreps_scalars, self.irreps_gates, self.irreps_gated = Irreps(irreps_scalars), Irreps(irreps_gates), Irreps(irreps_gated)
        assert self.irreps_gates.num_irreps == self.irreps_gated.num_irreps
        self.irreps_in = (self.irreps_scalars + self.irreps_gates + self.irreps_gated).simplify()
        self.irreps_out = (self.irreps_scalars + self.irreps_gated)
        expand, gstart, gdim = [], [], []
        g = 0
        for m, ir in self.irreps_gated:
            for _ in range(m):
                gstart.append(len(expand))
                gdim.append(ir.dim)
                expand += [g] * ir.dim
                g += 1
        self.register_buffer("_expand", torch.tensor(expand, dtype=torch.long), persistent=False)
        self.register_buffer("_expand32", torch.tensor(expand, dtype=torch.int32), persistent=False)
        self.register_buffer("_gstart", torch.tensor(gstart, dtype=torch.int32), persistent=False)
        self.register_buffer("_gdim", torch.tensor(gdim, dtype=torch.int32), persistent=False)

    def forward(self, x):
        ns, ng = self.irreps_scalars.dim, self.irreps_gates.dim
        if x.is_cuda and x.dim() == 2 and x.dtype == torch.float32:
            return _GateFn.apply(x, self._expand32, self._gstart, self._gdim, ns, ng, self.irreps_gated.dim)
        s, g, v = x[..., :ns], x[..., ns:ns + ng], x[..., ns + ng:]
        s = F.silu(s) * NORM2MOM["silu"]
        if ng == 0:
            return s
        g = torch.sigmoid(g) * NORM2MOM["sigmoid"]
        return torch.cat([s, v * g.index_select(-1, self._expand)], dim=-1)


class ScalarActivation(nn.Module):
    """e3nn.nn.Activation(out_irreps, [silu]) for an all-scalar output (tfn_layer.py:52-53)."""

    def forward(self, x):
        if x.is_cuda and x.dim() == 2 and x.dtype == torch.float32:
            return _GateFn.apply(x, None, None, None, x.shape[1], 0, 0)
        return F.silu(x) * NORM2MOM["silu"]


class BatchNorm(nn.Module):
    """e3nn.nn.BatchNorm(irreps) (eps 1e-5, momentum 0.1, affine, reduce='mean', 'component').
    `process_group`: when set (graph-sharded data parallel), batch statistics are all-reduced so that the result
    equals the single-process reference."""

    def __init__(self, irreps, eps=1e-5, momentum=0.1, affine=True, process_group=None):
        super().__init__()
        self.irreps = Irreps(irreps)
        self.eps, self.momentum, self.affine = eps, momentum, affine
        self.process_group = process_group
        num_scalar = sum(m for m, ir in self.irreps if ir.l == 0 and ir.p == 1)
        num_features = self.irreps.num_irreps
        self.register_buffer("running_mean", torch.zeros(num_scalar))
        self.register_buffer("running_var", torch.ones(num_features))
        if affine:
            self.weight = nn.Parameter(torch.ones(num_features))
            self.bias = nn.Parameter(torch.zeros(num_scalar))

    def _mean0(self, t: torch.Tensor, count: int) -> torch.Tensor:
        """mean over dim 0, across ranks when a process group is attached."""
        if self.process_group is None:
            return t.mean(0)
        import torch.distributed.nn.functional as dist_fn  # autograd-aware collectives
        s = t.sum(0)
        n = torch.tensor([float(count)], device=t.device, dtype=t.dtype)
        s, n = dist_fn.all_reduce(s, group=self.process_group), dist_fn.all_reduce(n, group=self.process_group)
        return s / n

    def forward(self, x):
        B = x.shape[0]
        new_means, new_vars, fields = [], [], []
        ix = irm = irv = iw = ib = 0
        for mul, ir in self.irreps:
            d = ir.dim
            field = x[:, ix:ix + mul * d].reshape(B, mul, d)
            ix += mul * d
            scalar = ir.l == 0 and ir.p == 1
            if scalar:
                if self.training:
                    mean = self._mean0(field.reshape(B, mul), B)
                    new_means.append((1 - self.momentum) * self.running_mean[irm:irm + mul] + self.momentum * mean.detach())
                else:
                    mean = self.running_mean[irm:irm + mul]
                irm += mul
                field = field - mean.reshape(1, mul, 1)
            if self.training:
                norm = self._mean0(field.pow(2).mean(2), B)
                new_vars.append((1 - self.momentum) * self.running_var[irv:irv + mul] + self.momentum * norm.detach())
            else:
                norm = self.running_var[irv:irv + mul]
            irv += mul
            scale = (norm + self.eps).pow(-0.5)
            if self.affine:
                scale = scale * self.weight[iw:iw + mul]
                iw += mul
            field = field * scale.reshape(1, mul, 1)
            if self.affine and scalar:
                field = field + self.bias[ib:ib + mul].reshape(mul, 1)
                ib += mul
            fields.append(field.reshape(B, mul * d))
        if self.training:
            with torch.no_grad():
                if new_means:
                    self.running_mean.copy_(torch.cat(new_means))
                self.running_var.copy_(torch.cat(new_vars))
        return torch.cat(fields, dim=1)


# ------------------------------------------------------------------------------------------------
# edge prologue
# ------------------------------------------------------------------------------------------------
class RadialEmbeddingBlock(nn.Module):
    """models/mace_modules/blocks.py:84-96 (BesselBasis x PolynomialCutoff, radial.py)."""

    def __init__(self, r_max: float, num_bessel: int, num_polynomial_cutoff: int):
        super().__init__()
        self.r_max, self.num_bessel, self.p = float(r_max), int(num_bessel), float(num_polynomial_cutoff)
        self.out_dim = num_bessel

    def forward(self, edge_lengths: torch.Tensor) -> torch.Tensor:  # [E,1] -> [E,num_bessel]
        x = edge_lengths
        w = (math.pi / self.r_max) * torch.linspace(1.0, self.num_bessel, self.num_bessel, dtype=x.dtype, device=x.device)
        bessel = math.sqrt(2.0 / self.r_max) * (torch.sin(w * x) / x)
        u, p = x / self.r_max, self.p
        env = (1.0 - ((p + 1.0) * (p + 2.0) / 2.0) * torch.pow(u, p) + p * (p + 2.0) * torch.pow(u, p + 1)
               - (p * (p + 1.0) / 2) * torch.pow(u, p + 2))
        return bessel * (env * (x < self.r_max))


class SphericalHarmonics(nn.Module):
    """e3nn.o3.SphericalHarmonics(irreps.spherical_harmonics(l), normalize=True, normalization='component'), l <= 2."""

    def __init__(self, max_ell: int):
        super().__init__()
        if max_ell > 2:
            raise NotImplementedError("spherical harmonics are built for max_ell <= 2 (the BASELINE configs)")
        self.max_ell = max_ell
        self.irreps_out = Irreps.spherical_harmonics(max_ell)

    def forward(self, v: torch.Tensor) -> torch.Tensor:
        v = F.normalize(v, dim=-1)
        x, y, z = v[..., 0], v[..., 1], v[..., 2]
        out = [torch.ones_like(x)]
        if self.max_ell >= 1:
            s3 = math.sqrt(3.0)
            out += [s3 * x, s3 * y, s3 * z]
        if self.max_ell >= 2:
            s3, s5 = math.sqrt(3.0), math.sqrt(5.0)
            out += [s5 * s3 * x * z, s5 * s3 * x * y, s5 * (y * y - 0.5 * (x * x + z * z)), s5 * s3 * y * z,
                    s5 * (s3 / 2.0) * (z * z - x * x)]
        return torch.stack(out, dim=-1)


def edge_geometry(pos: torch.Tensor, edge_index: torch.Tensor, max_ell: int, radial: RadialEmbeddingBlock):
    """(edge_sh [E,(L+1)^2], edge_feats [E,num_bessel]) in one kernel (models/tfn.py:171-175)."""
    if pos.requires_grad:
        raise NotImplementedError("gradients w.r.t. positions through the TFN/MACE edge prologue are not built "
                                  "(the reference trains without them)")
    pos, ei = pos.contiguous(), edge_index.contiguous()
    E = ei.shape[1]
    sh = torch.empty(E, (max_ell + 1) ** 2, dtype=pos.dtype, device=pos.device)
    rbf = torch.empty(E, radial.num_bessel, dtype=pos.dtype, device=pos.device)
    call("gmp_edge_geometry_fwd", ptr(pos), ptr(ei[0]), ptr(ei[1]), E, max_ell, radial.r_max, radial.num_bessel, radial.p,
         ptr(sh), ptr(rbf))
    return sh, rbf


# ------------------------------------------------------------------------------------------------
# tensor-product tables
# ------------------------------------------------------------------------------------------------
def _pow2_chunk(m: int, cap: int = 64) -> int:
    c = 1
    while c * 2 <= cap and m % (c * 2) == 0:
        c *= 2
    return c


class TensorProductPlan:
    """Device tables for the contract / wgrad kernels of one FullyConnectedTensorProduct(in, sh, out)."""

    def __init__(self, irreps_in, irreps_sh, irreps_out):
        self.irreps_in, self.irreps_sh, self.irreps_out = Irreps(irreps_in), Irreps(irreps_sh), Irreps(irreps_out)
        self.paths, self.weight_numel = fctp_paths(self.irreps_in, self.irreps_sh, self.irreps_out)
        for p in self.paths:
            if max(p.l_in, p.l_out) > 2 or p.l_sh > 3:
                raise NotImplementedError("fused tensor product: irreps up to l = 2 (sh up to l = 3)")
            if p.mul_in > 128 and _pow2_chunk(p.mul_in) < 8:
                raise NotImplementedError("fused tensor product: multiplicities need a power-of-two factor >= 8 above 128")
        cg: List[np.ndarray] = []
        off = 0

        def add_cg(t: np.ndarray) -> int:
            nonlocal off
            cg.append(t.astype(np.float32).reshape(-1))
            o = off
            off += t.size
            return o

        fwd_cg = [add_cg(p.coeff * wigner_3j(p.l_in, p.l_sh, p.l_out)) for p in self.paths]                     # [i][j][k]
        bwd_cg = [add_cg(p.coeff * wigner_3j(p.l_in, p.l_sh, p.l_out).transpose(2, 1, 0)) for p in self.paths]  # [k][j][i]
        self.fwd = self._contract_tables(
            [(m, ir.dim, o) for (m, ir), o in zip(self.irreps_out, self.irreps_out.offsets())],
            lambda p: p.i_out,
            lambda p, k: dict(w_off=p.w_off, stride_a=p.mul_out, stride_b=1, MA=p.mul_in, v_off=p.in_off, DA=2 * p.l_in + 1,
                              sh_off=p.sh_off, DS=2 * p.l_sh + 1, cg_off=fwd_cg[k]))
        self.bwd = self._contract_tables(
            [(m, ir.dim, o) for (m, ir), o in zip(self.irreps_in, self.irreps_in.offsets())],
            lambda p: p.i_in,
            lambda p, k: dict(w_off=p.w_off, stride_a=1, stride_b=p.mul_out, MA=p.mul_out, v_off=p.out_off, DA=2 * p.l_out + 1,
                              sh_off=p.sh_off, DS=2 * p.l_sh + 1, cg_off=bwd_cg[k]))
        units = []
        for k, p in enumerate(self.paths):
            rows = p.mul_in * p.mul_out
            for r0 in range(0, rows, 64):
                units.append([p.w_off + r0, min(64, rows - r0), r0, p.mul_out, p.in_off, 2 * p.l_in + 1, p.out_off,
                              2 * p.l_out + 1, p.sh_off, 2 * p.l_sh + 1, fwd_cg[k], 0])
        self.wunits = np.asarray(units, dtype=np.int32).reshape(-1, 12)
        self.cg = np.concatenate(cg) if cg else np.zeros(1, np.float32)
        # tensor-core path (csrc/tpconv_tc.cu): a = summed multiplicity index, b = kept one
        self.tc_fwd = self._tc_tables([
            dict(w_off=p.w_off, stride_a=p.mul_out, stride_b=1, MA=p.mul_in, MB=p.mul_out, v_off=p.in_off, DA=2 * p.l_in + 1,
                 DB=2 * p.l_out + 1, sh_off=p.sh_off, DS=2 * p.l_sh + 1, cg_off=fwd_cg[k], r_off=p.out_off)
            for k, p in enumerate(self.paths)])
        self.tc_bwd = self._tc_tables([
            dict(w_off=p.w_off, stride_a=1, stride_b=p.mul_out, MA=p.mul_out, MB=p.mul_in, v_off=p.out_off, DA=2 * p.l_out + 1,
                 DB=2 * p.l_in + 1, sh_off=p.sh_off, DS=2 * p.l_sh + 1, cg_off=bwd_cg[k], r_off=p.in_off)
            for k, p in enumerate(self.paths)])
        self._dev = {}

    def _contract_tables(self, blocks, block_of, pass_of):
        passes, blks, unit = [], [], 0
        for bi, (MB, DB, r_off) in enumerate(blocks):
            mine = [(k, p) for k, p in enumerate(self.paths) if block_of(p) == bi]
            begin = len(passes)
            mcs = []
            for k, p in mine:
                d = pass_of(p, k)
                MC = _pow2_chunk(d["MA"])
                mcs.append(MC)
                for a0 in range(0, d["MA"], MC):
                    passes.append([d["w_off"], d["stride_a"], d["stride_b"], a0, MC, d["v_off"], d["DA"], d["sh_off"],
                                   d["DS"], d["cg_off"], 0, 0])
            WS = max(1, min(16, 80 // DB, _SLICE // max(mcs) if mcs else 16, MB))
            blks.append([r_off, MB, DB, WS, begin, len(passes), unit, 0])
            unit += (MB + WS - 1) // WS
        return dict(passes=np.asarray(passes, dtype=np.int32).reshape(-1, 12), blocks=np.asarray(blks, dtype=np.int32).reshape(-1, 8),
                    nblocks=len(blks), nunits=unit)

    @staticmethod
    def _tc_tables(roles):
        """y-groups (path x range of the summed index), N-tiles (256 generated weights each, in the order the
        kernel consumes them), and the per-path tables of the bias term (struct layouts: csrc/tpconv_tc.cu)."""
        ygroups, ntiles, wtiles, ypaths, zent, bias = [], [], [], [], [], []
        y_off = z_off = 0
        for r in roles:
            DA, DB, MA, MB = r["DA"], r["DB"], r["MA"], r["MB"]
            xm = DA < DB
            M = DA if xm else DB
            WS = 32 if (not xm and DB == 1) else 8
            MC = 256 // WS
            AR = 64 if M == 1 else 32
            for A0 in range(0, MA, AR):
                ARv = min(AR, MA - A0)
                nsub, nslices = -(-ARv // MC), -(-MB // WS)
                ygroups.append([r["v_off"], DA, DB, r["DS"], r["sh_off"], r["cg_off"], A0, ARv, r["r_off"], MB, nslices, nsub,
                                len(ntiles), 0, 0, 0])
                for sl in range(nslices):
                    for q in range(nsub):
                        ntiles.append([r["w_off"], r["stride_a"], r["stride_b"], A0 + q * MC, A0 + ARv, sl * WS, MB, WS])
                        wtiles.append(ntiles[-1] + [r["v_off"], DA, DB, r["DS"], r["sh_off"], r["cg_off"], r["r_off"], 0])
            ypaths.append([r["v_off"], DA, DB, MA, y_off, z_off, 0, 0])
            for i in range(DA):
                for k in range(DB):
                    zent.append([r["sh_off"], r["DS"], r["cg_off"] + i * r["DS"] * DB + k, DB])
            bias.append(dict(y_off=y_off, MA=MA, MB=MB, DB=DB, r_off=r["r_off"], w_off=r["w_off"], stride_a=r["stride_a"],
                             stride_b=r["stride_b"]))
            y_off += MA * DB
            z_off += DA * DB
        i32 = lambda a, w: np.asarray(a, dtype=np.int32).reshape(-1, w)
        return dict(ygroups=i32(ygroups, 16), ntiles=i32(ntiles, 8), wtiles=i32(wtiles, 16), ypaths=i32(ypaths, 8), zent=i32(zent, 4), bias=bias,
                    y_len=y_off, npairs=sum(r["MA"] for r in roles))

    def device(self, dev):
        key = str(dev)
        if key not in self._dev:
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            d = dict(cg=t(self.cg), wunits=t(self.wunits),
                     fwd_passes=t(self.fwd["passes"]), fwd_blocks=t(self.fwd["blocks"]),
                     bwd_passes=t(self.bwd["passes"]), bwd_blocks=t(self.bwd["blocks"]))
            for name, tab in (("tc_fwd", self.tc_fwd), ("tc_bwd", self.tc_bwd)):
                for k in ("ygroups", "ntiles", "wtiles", "ypaths", "zent"):
                    d[f"{name}_{k}"] = t(tab[k])
            self._dev[key] = d
        return self._dev[key]


def _tc_images(csr, E, edge_feat, w1, b1, w2, tab, which, d):
    """bf16 UMMA operand images: the hidden layer of fc per 128-edge tile of `csr`'s order, and w2 per N-tile."""
    L = _lib.lib()
    H, R = w1.shape[0], w1.shape[1]
    if H % 64 != 0 or H > 256:
        raise NotImplementedError("precision='bf16': mlp_dim must be 64, 128, 192 or 256")
    NT = tab["ntiles"].shape[0]
    hid_img = torch.empty(max(int(L.gmp_tp_tc_hid_bytes(E, H)), 16), dtype=torch.uint8, device=w1.device)
    w2_img = torch.empty(max(int(L.gmp_tp_tc_w2_bytes(NT, H)), 16), dtype=torch.uint8, device=w1.device)
    call("gmp_tp_tc_pack_hid", csr.perm_ptr, E, ptr(edge_feat), R, ptr(w1), ptr(b1), H, ptr(hid_img))
    call("gmp_tp_tc_pack_w2", ptr(w2), H, ptr(d[which + "_ntiles"]), NT, ptr(w2_img))
    return hid_img, w2_img


def _tc_contract(csr, n, E, V, r_len, edge_sh, edge_feat, w1, b1, w2, b2, plan: "TensorProductPlan", which: str, d, images=None):
    """res[n] = sum_{e in CSR row n} sum_a (T_e[a,b] + b2[a,b]) Y_e[a,k] on the tensor cores (bf16 operands, fp32
    accumulation): gmp_tp_tc_contract for the T part, gmp_tp_ysum + node-level GEMMs (cuBLAS) for the bias part."""
    tab = plan.tc_fwd if which == "tc_fwd" else plan.tc_bwd
    L = _lib.lib()
    H, S = w1.shape[0], edge_sh.shape[1]
    hid_img, w2_img = images if images is not None else _tc_images(csr, E, edge_feat, w1, b1, w2, tab, which, d)
    res = torch.empty(n, r_len, dtype=torch.float32, device=V.device)
    head = torch.empty(int(L.gmp_tp_tc_num_chunks(E)), r_len, dtype=torch.float32, device=V.device)
    call("gmp_tp_tc_contract", ptr(csr.rowptr), ptr(csr.col), csr.perm_ptr, n, E, ptr(V), V.shape[1], ptr(res), r_len, ptr(head),
         ptr(edge_sh), S, ptr(hid_img), ptr(w2_img), ptr(d[which + "_ygroups"]), tab["ygroups"].shape[0], tab["ntiles"].shape[0], H,
         ptr(d["cg"]))
    ys = _tc_ysum(csr, n, E, V, edge_sh, tab, which, d)
    for bs in tab["bias"]:
        Bm = b2.as_strided((bs["MA"], bs["MB"]), (bs["stride_a"], bs["stride_b"]), b2.storage_offset() + bs["w_off"])
        ysp = ys[:, bs["y_off"]:bs["y_off"] + bs["MA"] * bs["DB"]].view(n, bs["MA"], bs["DB"])
        blk = res[:, bs["r_off"]:bs["r_off"] + bs["MB"] * bs["DB"]].view(n, bs["MB"], bs["DB"])
        blk += torch.einsum("nak,ab->nbk", ysp, Bm)
    return res, hid_img, ys


def _tc_wgrad(graph, x, g, edge_sh, edge_feat, w1, b1, w2, b2, plan: "TensorProductPlan", d, hid_img, ys):
    """Parameter gradients of fc on the tensor cores: dW2 = dT^T hid (gmp_tp_tc_dw2), dhid = dT W2 (gmp_tp_tc_dhid,
    masked by the ReLU and written as dL/d(pre-activation) per edge), db2 through the node-level aggregate YS,
    dW1 / db1 as plain GEMM / column sum of the pre-activation gradient."""
    tab, csr = plan.tc_fwd, graph.by_src
    n, E, H, R, S = graph.n, graph.E, w1.shape[0], w1.shape[1], edge_sh.shape[1]
    NT = tab["ntiles"].shape[0]
    w2_img = torch.empty(max(int(_lib.lib().gmp_tp_tc_w2_bytes(NT, H)), 16), dtype=torch.uint8, device=x.device)
    call("gmp_tp_tc_pack_w2", ptr(w2), H, ptr(d["tc_fwd_ntiles"]), NT, ptr(w2_img))
    dpre = torch.zeros(E, H, dtype=torch.float32, device=x.device) if E == 0 else torch.empty(E, H, dtype=torch.float32, device=x.device)
    rowid = csr.row_ids()
    call("gmp_tp_tc_dhid", ptr(csr.rowptr), ptr(csr.col), csr.perm_ptr, ptr(rowid), n, E, ptr(x), x.shape[1], ptr(g), g.shape[1], ptr(edge_sh), S,
         ptr(edge_feat), R, ptr(w1), ptr(b1), ptr(w2_img), ptr(d["tc_fwd_ygroups"]), tab["ygroups"].shape[0], NT, H, ptr(d["cg"]),
         ptr(dpre))
    if E > 0:
        # edge groups: the grid NT * G should fill whole waves of SMs (each CTA needs a full SM)
        sms = torch.cuda.get_device_properties(x.device).multi_processor_count
        ntile_e = -(-E // 128)
        G = min(range(1, max(1, min(8, ntile_e // 32)) + 1), key=lambda k: (-(-NT * k // sms)) / k)
        parts = torch.empty(G, w2.shape[0], H, dtype=torch.float32, device=x.device)
        call("gmp_tp_tc_dw2", ptr(csr.rowptr), ptr(csr.col), csr.perm_ptr, ptr(rowid), n, E, ptr(x), x.shape[1], ptr(g), g.shape[1],
             ptr(edge_sh), S, ptr(hid_img), ptr(d["tc_fwd_wtiles"]), NT, H, ptr(d["cg"]), w2.shape[0], G, ptr(parts))
        dW2 = parts[0] if G == 1 else parts.sum(0)
    else:
        dW2 = torch.zeros_like(w2)
    db2 = torch.zeros_like(b2)
    for bs in tab["bias"]:
        ysp = ys[:, bs["y_off"]:bs["y_off"] + bs["MA"] * bs["DB"]].view(n, bs["MA"], bs["DB"])
        gb = g[:, bs["r_off"]:bs["r_off"] + bs["MB"] * bs["DB"]].view(n, bs["MB"], bs["DB"])
        db2.as_strided((bs["MA"], bs["MB"]), (bs["stride_a"], bs["stride_b"]), bs["w_off"]).copy_(torch.einsum("nak,nbk->ab", ysp, gb))
    dW1 = dpre.t() @ edge_feat
    db1 = dpre.sum(0)
    return dW1, db1, dW2, db2


def _tc_ysum(csr, n, E, V, edge_sh, tab, which, d):
    """YS[n][path, a, k] = sum_{e in row n} sum_i V[col_e][a, i] Z_e[i, k]  (fp32; the bias term and db2 are linear in it)."""
    ys = torch.empty(n, max(tab["y_len"], 1), dtype=torch.float32, device=V.device)
    call("gmp_tp_ysum", ptr(csr.rowptr), ptr(csr.col), csr.perm_ptr, n, E, ptr(V), V.shape[1], ptr(edge_sh), edge_sh.shape[1],
         ptr(d[which + "_ypaths"]), tab["ypaths"].shape[0], tab["npairs"], ptr(d[which + "_zent"]), tab["zent"].shape[0],
         ptr(d["cg"]), ptr(ys), ys.shape[1])
    return ys


class _TPConvFn(torch.autograd.Function):
    """(node_attr, edge_sh, edge_feat, fc weights) -> sum_{e: edge_index[0][e] = n} TP(node_attr[edge_index[1][e]], sh_e; fc(feat_e))."""

    @staticmethod
    def forward(ctx, x, edge_sh, edge_feat, w1, b1, w2, b2, graph: Graph, plan: TensorProductPlan, precision: int):
        x, edge_sh, edge_feat = x.contiguous(), edge_sh.contiguous(), edge_feat.contiguous()
        w1, b1, w2, b2 = (t.contiguous() for t in (w1, b1, w2, b2))
        d = plan.device(x.device)
        csr = graph.by_src  # rows = edge_index[0] (aggregation), col = edge_index[1] (gather)
        H, R, S = w1.shape[0], w1.shape[1], edge_sh.shape[1]
        ctx.save_for_backward(x, edge_sh, edge_feat, w1, b1, w2, b2)
        ctx.graph, ctx.plan, ctx.precision = graph, plan, precision
        if precision == _lib.BF16_TC:
            # the hidden-layer image (bf16, E x mlp_dim) and the node aggregate YS are kept for the weight-gradient kernels
            out, ctx.hid_img, ctx.ys = _tc_contract(csr, graph.n, graph.E, x, plan.irreps_out.dim, edge_sh, edge_feat, w1, b1, w2, b2,
                                                    plan, "tc_fwd", d)
            return out
        out = torch.empty(graph.n, plan.irreps_out.dim, dtype=x.dtype, device=x.device)
        call("gmp_tp_contract", ptr(csr.rowptr), ptr(csr.col), csr.perm_ptr, graph.n, graph.E, ptr(x), x.shape[1], ptr(out),
             out.shape[1], ptr(edge_sh), S, ptr(edge_feat), R, ptr(w1), ptr(b1), ptr(w2), ptr(b2), H, ptr(d["fwd_passes"]),
             ptr(d["fwd_blocks"]), plan.fwd["nblocks"], plan.fwd["nunits"], ptr(d["cg"]), precision)
        return out

    @staticmethod
    def backward(ctx, g):
        x, edge_sh, edge_feat, w1, b1, w2, b2 = ctx.saved_tensors
        graph, plan, precision = ctx.graph, ctx.plan, ctx.precision
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            raise NotImplementedError("gradients w.r.t. edge_sh / edge_feat are not built into the fused tensor-product "
                                      "convolution (the reference never needs them: positions carry no gradient)")
        g = g.contiguous()
        d = plan.device(x.device)
        H, R, S = w1.shape[0], w1.shape[1], edge_sh.shape[1]
        dx = None
        if ctx.needs_input_grad[0]:
            t = graph.by_dst  # rows = edge_index[1] (where node_attr was gathered), col = edge_index[0]
            if precision == _lib.BF16_TC:
                dx = _tc_contract(t, graph.n, graph.E, g, x.shape[1], edge_sh, edge_feat, w1, b1, w2, b2, plan, "tc_bwd", d)[0]
            else:
                dx = torch.empty_like(x)
                call("gmp_tp_contract", ptr(t.rowptr), ptr(t.col), t.perm_ptr, graph.n, graph.E, ptr(g), g.shape[1], ptr(dx),
                     dx.shape[1], ptr(edge_sh), S, ptr(edge_feat), R, ptr(w1), ptr(b1), ptr(w2), ptr(b2), H, ptr(d["bwd_passes"]),
                     ptr(d["bwd_blocks"]), plan.bwd["nblocks"], plan.bwd["nunits"], ptr(d["cg"]), precision)
        if precision == _lib.BF16_TC:
            dW1, db1, dW2, db2 = _tc_wgrad(graph, x, g, edge_sh, edge_feat, w1, b1, w2, b2, plan, d, ctx.hid_img, ctx.ys)
            ctx.hid_img = ctx.ys = None
            return dx, None, None, dW1, db1, dW2, db2, None, None, None
        csr = graph.by_src
        nunits = plan.wunits.shape[0]
        plen = _lib.lib().gmp_tp_wgrad_part_len(H)
        dW2, db2 = torch.empty_like(w2), torch.empty_like(b2)
        parts = torch.empty(max(nunits, 1), plen, dtype=x.dtype, device=x.device)
        call("gmp_tp_wgrad", ptr(csr.rowptr), ptr(csr.col), csr.perm_ptr, graph.n, graph.E, ptr(x), x.shape[1], ptr(g), g.shape[1],
             ptr(edge_sh), S, ptr(edge_feat), R, ptr(w1), ptr(b1), ptr(w2), H, ptr(d["wunits"]), nunits, ptr(d["cg"]), ptr(dW2),
             ptr(db2), ptr(parts), _lib.FP32_STRICT)
        red = torch.empty(plen, dtype=x.dtype, device=x.device)
        call("gmp_reduce_partials_f32", ptr(parts), nunits, plen, ptr(red))
        dW1 = red[:H * 16].view(H, 16)[:, :R].contiguous()
        db1 = red[H * 16:H * 16 + H]
        return dx, None, None, dW1, db1, dW2, db2, None, None, None


class _TPInfo(nn.Module):
    """Stands where e3nn's FullyConnectedTensorProduct module sits (``layer.tp``): no parameters
    (shared_weights=False), exposes weight_numel and the instruction list."""

    def __init__(self, plan: TensorProductPlan):
        super().__init__()
        self.plan = plan
        self.weight_numel = plan.weight_numel
        self.instructions = plan.paths

    def extra_repr(self):
        return f"{self.plan.irreps_in} x {self.plan.irreps_sh} -> {self.plan.irreps_out} | {len(self.plan.paths)} paths | {self.weight_numel} weights"


class TensorProductConvLayer(nn.Module):
    """models/layers/tfn_layer.py:8-93."""

    def __init__(self, in_irreps, out_irreps, sh_irreps, edge_feats_dim, mlp_dim, aggr="add", batch_norm=False, gate=False,
                 precision: str = "fp32"):
        super().__init__()
        self.in_irreps, self.sh_irreps = Irreps(str(in_irreps)), Irreps(str(sh_irreps))
        out_irreps = Irreps(str(out_irreps))
        self.edge_feats_dim, self.aggr, self.precision = edge_feats_dim, aggr, precision
        if aggr not in ("add", "sum", "mean"):
            raise NotImplementedError(f"aggr={aggr!r}: the fused reduction implements add/sum/mean")
        if gate:
            scal, gates, gated = gate_split(out_irreps)
            if gated.num_irreps == 0:
                self.gate = ScalarActivation()
            else:
                self.gate = Gate(scal, gates, gated)
                out_irreps = self.gate.irreps_in
        else:
            self.gate = None
        self.out_irreps = out_irreps
        self.tp = _TPInfo(TensorProductPlan(self.in_irreps, self.sh_irreps, out_irreps))
        self.fc = nn.Sequential(nn.Linear(edge_feats_dim, mlp_dim), nn.ReLU(), nn.Linear(mlp_dim, self.tp.weight_numel))
        self.batch_norm = BatchNorm(out_irreps) if batch_norm else None

    def forward(self, node_attr, edge_index, edge_sh, edge_feat):
        graph = get_graph(edge_index, node_attr.shape[0])
        out = _TPConvFn.apply(node_attr, edge_sh, edge_feat, self.fc[0].weight, self.fc[0].bias, self.fc[2].weight,
                              self.fc[2].bias, graph, self.tp.plan, _PREC[self.precision])
        if self.aggr == "mean":
            rp = graph.by_src.rowptr
            out = out / (rp[1:] - rp[:-1]).clamp(min=1).to(out.dtype).unsqueeze(1)
        if self.gate is not None:
            out = self.gate(out)
        if self.batch_norm is not None:
            out = self.batch_norm(out)
        return out


def first_node_pooling(x, batch, size=None):
    """models/tfn.py:13-40: the first node of every graph."""
    shifted = torch.cat([batch[-1:], batch[:-1]])
    shifted[0] = -1
    return x[(batch - shifted) == 1]


class TFNModel(nn.Module):
    """models/tfn.py:42-190."""

    def __init__(self, r_max: float = 10.0, num_bessel: int = 8, num_polynomial_cutoff: int = 5, max_ell: int = 2,
                 num_layers: int = 5, emb_dim: int = 64, hidden_irreps=None, mlp_dim: int = 256, in_dim: int = 1,
                 out_dim: int = 1, aggr: str = "sum", pool: str = "first", gate: bool = True, batch_norm: bool = False,
                 residual: bool = True, equivariant_pred: bool = False, precision: str = "fp32"):
        super().__init__()
        self.r_max, self.max_ell, self.num_layers, self.emb_dim, self.mlp_dim = r_max, max_ell, num_layers, emb_dim, mlp_dim
        self.residual, self.batch_norm, self.gate, self.equivariant_pred = residual, batch_norm, gate, equivariant_pred
        self.radial_embedding = RadialEmbeddingBlock(r_max, num_bessel, num_polynomial_cutoff)
        sh_irreps = Irreps.spherical_harmonics(max_ell)
        self.spherical_harmonics = SphericalHarmonics(max_ell)
        self.emb_in = torch.nn.Embedding(in_dim, emb_dim)
        hidden = default_hidden_irreps(max_ell, emb_dim) if hidden_irreps is None else Irreps(str(hidden_irreps))
        self.hidden_irreps = hidden
        ins = [Irreps(f"{emb_dim}x0e")] + [hidden] * (num_layers - 1)
        self.convs = torch.nn.ModuleList([
            TensorProductConvLayer(i, hidden, sh_irreps, self.radial_embedding.out_dim, mlp_dim, aggr, batch_norm, gate, precision)
            for i in ins])
        self.pool = {"mean": global_mean_pool, "sum": global_add_pool, "first": first_node_pooling}[pool]
        if equivariant_pred:
            self.pred = torch.nn.Linear(hidden.dim, out_dim)
        else:
            self.pred = torch.nn.Sequential(torch.nn.Linear(emb_dim, emb_dim), torch.nn.ReLU(), torch.nn.Linear(emb_dim, out_dim))

    def forward(self, batch):
        h = embedding_lookup(self.emb_in, batch.atoms)
        edge_sh, edge_feats = edge_geometry(batch.pos, batch.edge_index, self.max_ell, self.radial_embedding)
        for conv in self.convs:
            h_update = conv(h, batch.edge_index, edge_sh, edge_feats)
            h = h_update + F.pad(h, (0, h_update.shape[-1] - h.shape[-1])) if self.residual else h_update
        out = self.pool(h, batch.batch, getattr(batch, "num_graphs", None))
        if not self.equivariant_pred:
            out = out[:, :self.emb_dim]
        return self.pred(out)
