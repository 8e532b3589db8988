"""Restatement of ``e3nn.nn.{Activation, Gate, BatchNorm}``.

ORACLE / TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED against the wheel; pinned
by tests/test_oracle_thirdparty.py (normalize2mom constants, gate layout,
batch-norm statistics).  Reference call sites: models/layers/tfn_layer.py:45-63,80;
models/mace_modules/blocks.py:124.
"""
from __future__ import annotations

import torch

from .o3 import ElementwiseTensorProduct, Irreps


def _second_moment_constant(f) -> float:
    """e3nn ``normalize2mom``: 1/sqrt(E_{z~N(0,1)} f(z)^2) on 1e6 seeded fp64 samples."""
    gen = torch.Generator(device="cpu").manual_seed(0)
    z = torch.randn(1_000_000, generator=gen, dtype=torch.float64)
    with torch.no_grad():
        return f(z).pow(2).mean().pow(-0.5).item()


class normalize2mom:
    _cache = {}

    def __init__(self, f):
        key = getattr(f, "__name__", repr(f))
        if key not in normalize2mom._cache:
            normalize2mom._cache[key] = _second_moment_constant(f)
        self.cst = normalize2mom._cache[key]
        self.f = f
        self._is_id = abs(self.cst - 1) < 1e-4

    def __call__(self, x):
        return self.f(x) if self._is_id else self.f(x).mul(self.cst)


class Activation(torch.nn.Module):
    def __init__(self, irreps_in, acts):
        super().__init__()
        self.irreps_in = Irreps(irreps_in)
        acts = list(acts)
        if len(acts) != len(self.irreps_in):
            raise ValueError(f"Irreps in and number of activation functions does not match: {len(acts)}, "
                             f"({self.irreps_in}, {acts})")
        self.acts = [normalize2mom(a) if a is not None else None for a in acts]
        for (mul, ir), a in zip(self.irreps_in, self.acts):
            if a is not None and ir.l != 0:
                raise ValueError("Activation: cannot apply an activation function to a non-scalar input.")
        self.irreps_out = self.irreps_in

    def forward(self, x):
        outs, i = [], 0
        for (mul, ir), a in zip(self.irreps_in, self.acts):
            blk = x[..., i:i + mul * ir.dim]
            outs.append(a(blk) if a is not None else blk)
            i += mul * ir.dim
        return torch.cat(outs, dim=-1) if outs else x


class Gate(torch.nn.Module):
    def __init__(self, irreps_scalars, act_scalars, irreps_gates, act_gates, irreps_gated):
        super().__init__()
        self.irreps_scalars = Irreps(irreps_scalars)
        self.irreps_gates = Irreps(irreps_gates)
        self.irreps_gated = Irreps(irreps_gated)
        assert self.irreps_gates.num_irreps == self.irreps_gated.num_irreps
        self.irreps_in = (self.irreps_scalars + self.irreps_gates + self.irreps_gated).simplify()
        self.act_scalars = Activation(self.irreps_scalars, act_scalars)
        self.act_gates = Activation(self.irreps_gates, act_gates)
        self.mul = ElementwiseTensorProduct(self.irreps_gated, self.irreps_gates)
        self.irreps_out = self.irreps_scalars + self.mul.irreps_out

    def forward(self, x):
        ns, ng = self.irreps_scalars.dim, self.irreps_gates.dim
        scalars, gates, gated = x[..., :ns], x[..., ns:ns + ng], x[..., ns + ng:]
        scalars = self.act_scalars(scalars)
        if gates.shape[-1]:
            gates = self.act_gates(gates)
            gated = self.mul(gated, gates)
            return torch.cat([scalars, gated], dim=-1)
        return scalars


class BatchNorm(torch.nn.Module):
    def __init__(self, irreps, eps=1e-5, momentum=0.1, affine=True, reduce="mean", instance=False,
                 normalization="component"):
        super().__init__()
        self.irreps = Irreps(irreps)
        self.eps, self.momentum, self.affine = eps, momentum, affine
        assert reduce == "mean" and not instance and normalization == "component"
        num_scalar = sum(mul for mul, ir in self.irreps if ir.is_scalar())
        num_features = self.irreps.num_irreps
        self.register_buffer("running_mean", torch.zeros(num_scalar))
        self.register_buffer("running_var", torch.ones(num_features))
        if affine:
            self.weight = torch.nn.Parameter(torch.ones(num_features))
            self.bias = torch.nn.Parameter(torch.zeros(num_scalar))

    def forward(self, x):
        batch, dim = x.shape[0], x.shape[-1]
        x = x.reshape(batch, -1, dim)
        new_means, new_vars, fields = [], [], []
        ix = irm = irv = iw = ib = 0
        for mul, ir in self.irreps:
            d = ir.dim
            field = x[:, :, ix:ix + mul * d].reshape(batch, -1, mul, d)
            ix += mul * d
            if ir.is_scalar():
                if self.training:
                    mean = field.mean([0, 1]).reshape(mul)
                    new_means.append((1 - self.momentum) * self.running_mean[irm:irm + mul]
                                     + self.momentum * mean.detach())
                else:
                    mean = self.running_mean[irm:irm + mul]
                irm += mul
                field = field - mean.reshape(-1, 1, mul, 1)
            if self.training:
                norm = field.pow(2).mean(3).mean(1).mean(0)
                new_vars.append((1 - self.momentum) * self.running_var[irv:irv + mul]
                                + self.momentum * norm.detach())
            else:
                norm = self.running_var[irv:irv + mul]
            irv += mul
            scale = (norm + self.eps).pow(-0.5)
            if self.affine:
                scale = scale * self.weight[iw:iw + mul]
                iw += mul
            field = field * scale.reshape(-1, 1, mul, 1)
            if self.affine and ir.is_scalar():
                field = field + self.bias[ib:ib + mul].reshape(mul, 1)
                ib += mul
            fields.append(field.reshape(batch, -1, mul * d))
        if self.training:
            if new_means:
                self.running_mean.copy_(torch.cat(new_means))
            self.running_var.copy_(torch.cat(new_vars))
        return torch.cat(fields, dim=2).reshape(batch, dim) if x.shape[1] == 1 else torch.cat(fields, dim=2)


class FullyConnectedNet(torch.nn.Sequential):
    """e3nn 0.4.4 ``nn.FullyConnectedNet(hs, act=None, variance_in=1, variance_out=1, out_act=False)`` (SURVEY.md 8f.2:
    the radial MLP of the ACEsuit interaction blocks, models/mace_modules/blocks.py:243-246, 299-302, 428-431).
    [upstream-recalled] A stack of bias-free layers, weight ~ randn(h_in, h_out); a layer computes
    ``act(x @ (W / sqrt(h_in * var_in))) * sqrt(var_out)`` with ``act = normalize2mom(act)`` (second moment 1 under a
    standard normal input), the last one without activation unless ``out_act``: ``x @ (W / sqrt(h_in * var_in / var_out))``.
    State-dict keys ``layer{i}.weight``."""

    class _Layer(torch.nn.Module):
        def __init__(self, h_in, h_out, act, var_in, var_out):
            super().__init__()
            self.weight = torch.nn.Parameter(torch.randn(h_in, h_out))
            self.act, self.h_in, self.var_in, self.var_out = act, h_in, var_in, var_out

        def forward(self, x):
            if self.act is not None:
                x = x @ (self.weight / (self.h_in * self.var_in) ** 0.5)
                return self.act(x) * self.var_out ** 0.5
            return x @ (self.weight / (self.h_in * self.var_in / self.var_out) ** 0.5)

    def __init__(self, hs, act=None, variance_in=1, variance_out=1, out_act=False):
        super().__init__()
        self.hs = list(hs)
        act = normalize2mom(act) if act is not None else None
        var_in = variance_in
        for i, (h1, h2) in enumerate(zip(self.hs, self.hs[1:])):
            last = i == len(self.hs) - 2
            a = act if (not last or out_act) else None
            setattr(self, f"layer{i}", FullyConnectedNet._Layer(h1, h2, a, var_in, variance_out if last else 1))
            var_in = 1
