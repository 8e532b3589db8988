"""Restatement of the parts of ``e3nn.o3`` the reference calls.

ORACLE / TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED against the real wheel
(e3nn 0.4.4 pinned in the reference README.md:53, 0.5.1 printed in
experiments/rotsym.ipynb:50): restated from the published algorithm and
pinned by the closed-form tests in tests/test_oracle_thirdparty.py.

Reference call sites: models/tfn.py:110-113,128; models/mace.py:82-85,103;
models/layers/tfn_layer.py:47-80; models/mace_modules/blocks.py:121-123;
models/mace_modules/cg.py:52; models/mace_modules/irreps_tools.py.
"""
from __future__ import annotations

import collections
import math
from fractions import Fraction
from functools import lru_cache
from math import factorial
from typing import List, Tuple

import torch


# --------------------------------------------------------------------------- #
# Irrep / Irreps
# --------------------------------------------------------------------------- #
class Irrep(tuple):
    """(l, p) with p = +1 (even) / -1 (odd).  Ordering is plain tuple order."""

    def __new__(cls, l, p=None):
        if p is None:
            if isinstance(l, Irrep):
                return l
            if isinstance(l, _MulIr):
                return l.ir
            if isinstance(l, str):
                s = l.strip()
                p = {"e": 1, "o": -1, "y": None}[s[-1]]
                l = int(s[:-1])
                if p is None:
                    p = (-1) ** l
            elif isinstance(l, tuple):
                l, p = l
        assert isinstance(l, int) and l >= 0, l
        assert p in (-1, 1), p
        return super().__new__(cls, (l, p))

    @property
    def l(self) -> int:  # noqa: E743
        return self[0]

    @property
    def p(self) -> int:
        return self[1]

    @property
    def dim(self) -> int:
        return 2 * self.l + 1

    def is_scalar(self) -> bool:
        return self.l == 0 and self.p == 1

    def __repr__(self):
        return f"{self.l}{'e' if self.p == 1 else 'o'}"

    def __mul__(self, other):
        other = Irrep(other)
        p = self.p * other.p
        for l in range(abs(self.l - other.l), self.l + other.l + 1):
            yield Irrep(l, p)

    def __rmul__(self, mul):
        assert isinstance(mul, int)
        return Irreps([(mul, self)])

    def __add__(self, other):
        return Irreps(self) + Irreps(other)

    def __contains__(self, _):
        raise NotImplementedError

    def __len__(self):
        raise NotImplementedError


class _MulIr(tuple):
    def __new__(cls, mul, ir=None):
        if ir is None:
            mul, ir = mul
        assert isinstance(mul, int)
        return super().__new__(cls, (mul, Irrep(ir)))

    @property
    def mul(self) -> int:
        return self[0]

    @property
    def ir(self) -> Irrep:
        return self[1]

    @property
    def dim(self) -> int:
        return self.mul * self.ir.dim

    def __repr__(self):
        return f"{self.mul}x{self.ir}"


class Irreps(tuple):
    def __new__(cls, irreps=None):
        if isinstance(irreps, Irreps):
            return super().__new__(cls, irreps)
        out = []
        if isinstance(irreps, Irrep):
            out.append(_MulIr(1, irreps))
        elif isinstance(irreps, _MulIr):
            out.append(irreps)
        elif isinstance(irreps, str):
            if irreps.strip() != "":
                for piece in irreps.split("+"):
                    piece = piece.strip()
                    if "x" in piece:
                        mul, ir = piece.split("x")
                        out.append(_MulIr(int(mul), Irrep(ir)))
                    else:
                        out.append(_MulIr(1, Irrep(piece)))
        elif irreps is None:
            pass
        else:
            for item in irreps:
                if isinstance(item, _MulIr):
                    out.append(item)
                elif isinstance(item, Irrep):
                    out.append(_MulIr(1, item))
                elif isinstance(item, str):
                    out.append(_MulIr(1, Irrep(item)))
                elif len(item) == 2 and isinstance(item[0], int) and not isinstance(item[1], int):
                    out.append(_MulIr(item[0], Irrep(item[1])))
                elif len(item) == 2:
                    # ambiguous (int, int): e3nn reads a bare pair as an Irrep (l, p)
                    out.append(_MulIr(1, Irrep(item)))
                else:
                    raise ValueError(f"cannot parse {item!r}")
        return super().__new__(cls, out)

    @staticmethod
    def spherical_harmonics(lmax: int, p: int = -1) -> "Irreps":
        return Irreps([(1, (l, p ** l)) for l in range(lmax + 1)])

    def slices(self):
        s, i = [], 0
        for mul_ir in self:
            s.append(slice(i, i + mul_ir.dim))
            i += mul_ir.dim
        return s

    @property
    def dim(self) -> int:
        return sum(mul_ir.dim for mul_ir in self)

    @property
    def num_irreps(self) -> int:
        return sum(mul for mul, _ in self)

    @property
    def ls(self) -> List[int]:
        return [ir.l for mul, ir in self for _ in range(mul)]

    @property
    def lmax(self) -> int:
        return max(self.ls)

    def count(self, ir) -> int:
        ir = Irrep(ir)
        return sum(mul for mul, ir_this in self if ir == ir_this)

    def __contains__(self, ir) -> bool:
        ir = Irrep(ir)
        return any(ir == ir_this for _, ir_this in self)

    def __getitem__(self, i):
        x = super().__getitem__(i)
        if isinstance(i, slice):
            return Irreps(x)
        return x

    def __add__(self, other):
        return Irreps(tuple.__add__(self, Irreps(other)))

    def __mul__(self, n: int):
        assert isinstance(n, int)
        return Irreps(tuple.__mul__(self, n))

    def __rmul__(self, n: int):
        return self.__mul__(n)

    def simplify(self) -> "Irreps":
        out: List[Tuple[int, Irrep]] = []
        for mul, ir in self:
            if out and out[-1][1] == ir:
                out[-1] = (out[-1][0] + mul, ir)
            elif mul > 0:
                out.append((mul, ir))
        return Irreps(out)

    def remove_zero_multiplicities(self) -> "Irreps":
        return Irreps([(mul, ir) for mul, ir in self if mul > 0])

    def sort(self):
        Ret = collections.namedtuple("sort", ["irreps", "p", "inv"])
        out = sorted((ir, i, mul) for i, (mul, ir) in enumerate(self))
        inv = tuple(i for _, i, _ in out)
        p = [0] * len(inv)
        for pos, i in enumerate(inv):
            p[i] = pos
        irreps = Irreps([(mul, ir) for ir, _, mul in out])
        return Ret(irreps, tuple(p), inv)

    def __repr__(self):
        return "+".join(f"{mul_ir}" for mul_ir in self)


# --------------------------------------------------------------------------- #
# Real Wigner 3j (e3nn 0.5.x `_wigner.py` algorithm)
# --------------------------------------------------------------------------- #
def _su2_cg_coeff(j1, m1, j2, m2, j3, m3) -> float:
    if m3 != m1 + m2:
        return 0.0
    vmin = int(max(-j1 + j2 + m3, -j1 + m1, 0))
    vmax = int(min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3))

    def f(n):
        return factorial(round(n))

    C = (
        (2.0 * j3 + 1.0)
        * Fraction(
            f(j3 + j1 - j2) * f(j3 - j1 + j2) * f(j1 + j2 - j3) * f(j3 + m3) * f(j3 - m3),
            f(j1 + j2 + j3 + 1) * f(j1 - m1) * f(j1 + m1) * f(j2 - m2) * f(j2 + m2),
        )
    ) ** 0.5
    S = 0
    for v in range(vmin, vmax + 1):
        S += (-1) ** int(v + j2 + m2) * Fraction(
            f(j2 + j3 + m1 - v) * f(j1 - m1 + v),
            f(v) * f(j3 - j1 + j2 - v) * f(j3 + m3 - v) * f(v + j1 - j2 - m3),
        )
    return float(C * S)


def _su2_cg(j1: int, j2: int, j3: int) -> torch.Tensor:
    mat = torch.zeros(2 * j1 + 1, 2 * j2 + 1, 2 * j3 + 1, dtype=torch.float64)
    if abs(j1 - j2) <= j3 <= j1 + j2:
        for m1 in range(-j1, j1 + 1):
            for m2 in range(-j2, j2 + 1):
                if abs(m1 + m2) <= j3:
                    mat[j1 + m1, j2 + m2, j3 + m1 + m2] = _su2_cg_coeff(j1, m1, j2, m2, j3, m1 + m2)
    return mat


def _real_to_complex(l: int) -> torch.Tensor:
    q = torch.zeros(2 * l + 1, 2 * l + 1, dtype=torch.complex128)
    s2 = 1 / math.sqrt(2)
    for m in range(-l, 0):
        q[l + m, l + abs(m)] = s2
        q[l + m, l - abs(m)] = -1j * s2
    q[l, l] = 1
    for m in range(1, l + 1):
        q[l + m, l + abs(m)] = (-1) ** m * s2
        q[l + m, l - abs(m)] = 1j * (-1) ** m * s2
    return (-1j) ** l * q


@lru_cache(maxsize=None)
def _wigner_3j_f64(l1: int, l2: int, l3: int) -> torch.Tensor:
    Q1, Q2, Q3 = _real_to_complex(l1), _real_to_complex(l2), _real_to_complex(l3)
    C = _su2_cg(l1, l2, l3).to(torch.complex128)
    C = torch.einsum("ij,kl,mn,ikn->jlm", Q1, Q2, torch.conj(Q3.T), C)
    assert torch.all(torch.abs(C.imag) < 1e-5)
    C = C.real
    return C / C.norm()


def wigner_3j(l1: int, l2: int, l3: int, dtype=None, device=None) -> torch.Tensor:
    assert abs(l2 - l3) <= l1 <= l2 + l3
    if dtype is None:
        dtype = torch.get_default_dtype()
    return _wigner_3j_f64(l1, l2, l3).to(dtype=dtype, device=device).clone()


# --------------------------------------------------------------------------- #
# Real spherical harmonics, l <= 3 (e3nn `_spherical_harmonics`)
# --------------------------------------------------------------------------- #
def _raw_sh(lmax: int, x, y, z) -> torch.Tensor:
    """Unit-norm-on-the-sphere real SH polynomial blocks, e3nn component order."""
    out = [torch.ones_like(x)]
    if lmax >= 1:
        out += [x, y, z]
    if lmax >= 2:
        s3 = math.sqrt(3.0)
        x2z2 = x * x + z * z
        out += [s3 * x * z, s3 * x * y, y * y - 0.5 * x2z2, s3 * y * z, (s3 / 2.0) * (z * z - x * x)]
    if lmax >= 3:
        # e3nn recursion for l = 3 (used only by oracle-side extras; hot path is l<=2)
        s3 = math.sqrt(3.0)
        sh_2_0, sh_2_1, sh_2_2, sh_2_3, sh_2_4 = out[4:9]
        x2z2 = x * x + z * z
        y2 = y * y
        c = math.sqrt(5.0 / 6.0)
        out += [
            c * (sh_2_0 * z + sh_2_4 * x),
            math.sqrt(5.0) * sh_2_0 * y,
            math.sqrt(3.0 / 8.0) * (4.0 * y2 - x2z2) * x,
            0.5 * y * (2.0 * y2 - 3.0 * x2z2),
            math.sqrt(3.0 / 8.0) * z * (4.0 * y2 - x2z2),
            math.sqrt(5.0) * sh_2_4 * y,
            c * (sh_2_4 * z - sh_2_0 * x),
        ]
    if lmax >= 4:
        raise NotImplementedError("oracle spherical harmonics are restated for l <= 3 only")
    return torch.stack(out, dim=-1)


class SphericalHarmonics(torch.nn.Module):
    def __init__(self, irreps_out, normalize: bool, normalization: str = "integral", irreps_in=None):
        super().__init__()
        if isinstance(irreps_out, int):
            irreps_out = Irreps([(1, (irreps_out, (-1) ** irreps_out))])
        self.irreps_out = Irreps(irreps_out)
        self._ls = [ir.l for mul, ir in self.irreps_out for _ in range(mul)]
        self._lmax = max(self._ls)
        self.normalize = normalize
        self.normalization = normalization
        assert normalization in ("integral", "component", "norm")

    def forward(self, v: torch.Tensor) -> torch.Tensor:
        if self.normalize:
            v = torch.nn.functional.normalize(v, dim=-1)
        sh = _raw_sh(self._lmax, v[..., 0], v[..., 1], v[..., 2])
        sh = torch.cat([sh[..., l * l:(l + 1) * (l + 1)] for l in self._ls], dim=-1)
        if self.normalization == "integral":
            f = torch.cat([torch.full((2 * l + 1,), math.sqrt(2 * l + 1) / math.sqrt(4 * math.pi), dtype=sh.dtype,
                                      device=sh.device) for l in self._ls])
            sh = sh * f
        elif self.normalization == "component":
            f = torch.cat([torch.full((2 * l + 1,), math.sqrt(2 * l + 1), dtype=sh.dtype, device=sh.device)
                           for l in self._ls])
            sh = sh * f
        return sh


def spherical_harmonics(l, x, normalize, normalization="integral"):
    return SphericalHarmonics(l, normalize, normalization)(x)


# --------------------------------------------------------------------------- #
# Tensor products
# --------------------------------------------------------------------------- #
Instruction = collections.namedtuple("Instruction", "i_in1 i_in2 i_out mode has_weight path_weight path_shape")


class TensorProduct(torch.nn.Module):
    """'uvw' / 'uvu' / 'uuu' paths, e3nn default normalisations
    (irrep_normalization='component', path_normalization='element')."""

    def __init__(self, irreps_in1, irreps_in2, irreps_out, instructions, shared_weights=None,
                 internal_weights=None, irrep_normalization="component", path_normalization="element"):
        super().__init__()
        self.irreps_in1 = Irreps(irreps_in1)
        self.irreps_in2 = Irreps(irreps_in2)
        self.irreps_out = Irreps(irreps_out)
        ins = []
        for i in instructions:
            i1, i2, io, mode, has_w = i[:5]
            pw = i[5] if len(i) > 5 else 1.0
            m1, m2, mo = self.irreps_in1[i1].mul, self.irreps_in2[i2].mul, self.irreps_out[io].mul
            shape = {"uvw": (m1, m2, mo), "uvu": (m1, m2), "uvv": (m1, m2), "uuw": (m1, mo),
                     "uuu": (m1,), "uvuv": (m1, m2)}[mode]
            ins.append(Instruction(i1, i2, io, mode, has_w, pw, shape))

        def num_elements(i):
            m1, m2 = self.irreps_in1[i.i_in1].mul, self.irreps_in2[i.i_in2].mul
            return {"uvw": m1 * m2, "uvu": m2, "uvv": m1, "uuw": m1, "uuu": 1, "uvuv": 1}[i.mode]

        normed = []
        for i in ins:
            ir_out = self.irreps_out[i.i_out].ir
            ir1, ir2 = self.irreps_in1[i.i_in1].ir, self.irreps_in2[i.i_in2].ir
            assert ir_out in list(ir1 * ir2)
            alpha = {"component": ir_out.dim, "norm": ir1.dim * ir2.dim, "none": 1}[irrep_normalization]
            if path_normalization == "element":
                x = sum(num_elements(j) for j in ins if j.i_out == i.i_out)
            elif path_normalization == "path":
                x = num_elements(i) * len([j for j in ins if j.i_out == i.i_out])
            else:
                x = 1
            alpha = alpha / x if x > 0 else alpha
            alpha *= i.path_weight
            normed.append(i._replace(path_weight=math.sqrt(alpha)))
        self.instructions = normed
        self.weight_numel = sum(math.prod(i.path_shape) for i in self.instructions if i.has_weight)
        if shared_weights is False and internal_weights is None:
            internal_weights = False
        if shared_weights is None:
            shared_weights = True
        if internal_weights is None:
            internal_weights = shared_weights and self.weight_numel > 0
        self.shared_weights = shared_weights
        self.internal_weights = internal_weights
        if internal_weights and self.weight_numel > 0:
            self.weight = torch.nn.Parameter(torch.randn(self.weight_numel))

    def forward(self, x1: torch.Tensor, x2: torch.Tensor, weight: torch.Tensor = None) -> torch.Tensor:
        if weight is None and self.weight_numel > 0:
            weight = self.weight
        lead = x1.shape[:-1]
        x1 = x1.reshape(-1, x1.shape[-1])
        x2 = x2.reshape(-1, x2.shape[-1])
        Z = x1.shape[0]
        if weight is not None and not self.shared_weights:
            weight = weight.reshape(-1, self.weight_numel)
        s1, s2 = self.irreps_in1.slices(), self.irreps_in2.slices()
        outs = [None] * len(self.irreps_out)
        off = 0
        for ins in self.instructions:
            m1, ir1 = self.irreps_in1[ins.i_in1]
            m2, ir2 = self.irreps_in2[ins.i_in2]
            mo, iro = self.irreps_out[ins.i_out]
            a = x1[:, s1[ins.i_in1]].reshape(Z, m1, ir1.dim)
            b = x2[:, s2[ins.i_in2]].reshape(Z, m2, ir2.dim)
            w3j = wigner_3j(ir1.l, ir2.l, iro.l, dtype=x1.dtype, device=x1.device)
            w = None
            if ins.has_weight:
                n = math.prod(ins.path_shape)
                if self.shared_weights:
                    w = weight[off:off + n].reshape(ins.path_shape)
                else:
                    w = weight[:, off:off + n].reshape((-1,) + tuple(ins.path_shape))
                off += n
            z = "" if (w is None or self.shared_weights) else "z"
            # contraction order: inputs x CG first ([Z,u,v,k], small), weights last -- contracting the per-edge weights
            # with the CG tensor first (einsum's left-to-right default) materialises [Z,u,v,w,i,j,k], ~2 GB per path
            # and 1000 edges at C = 64.  Same sum, reassociated (the golden fixtures pin it at 2e-6).
            if ins.mode == "uvw":
                xx = torch.einsum("zui,zvj->zuvij", a, b)
                t = torch.einsum("ijk,zuvij->zuvk", w3j, xx)
                r = torch.einsum(f"{z}uvw,zuvk->zwk", w, t)
            elif ins.mode == "uvu":
                xx = torch.einsum("zui,zvj->zuvij", a, b)
                t = torch.einsum("ijk,zuvij->zuvk", w3j, xx)
                r = (torch.einsum(f"{z}uv,zuvk->zuk", w, t) if w is not None else t.sum(dim=2))
            elif ins.mode == "uuu":
                r = (torch.einsum(f"{z}u,ijk,zui,zuj->zuk", w, w3j, a, b) if w is not None
                     else torch.einsum("ijk,zui,zuj->zuk", w3j, a, b))
            else:
                raise NotImplementedError(ins.mode)
            r = ins.path_weight * r.reshape(Z, mo * iro.dim)
            outs[ins.i_out] = r if outs[ins.i_out] is None else outs[ins.i_out] + r
        for k, (mo, iro) in enumerate(self.irreps_out):
            if outs[k] is None:
                outs[k] = x1.new_zeros(Z, mo * iro.dim)
        return torch.cat(outs, dim=-1).reshape(lead + (self.irreps_out.dim,))


class FullyConnectedTensorProduct(TensorProduct):
    def __init__(self, irreps_in1, irreps_in2, irreps_out, irrep_normalization=None, path_normalization=None, **kw):
        i1, i2, io = Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out)
        instr = [
            (a, b, c, "uvw", True, 1.0)
            for a, (_, ir_1) in enumerate(i1)
            for b, (_, ir_2) in enumerate(i2)
            for c, (_, ir_out) in enumerate(io)
            if ir_out in list(ir_1 * ir_2)
        ]
        super().__init__(i1, i2, io, instr,
                         irrep_normalization=irrep_normalization or "component",
                         path_normalization=path_normalization or "element", **kw)


class ElementwiseTensorProduct(TensorProduct):
    def __init__(self, irreps_in1, irreps_in2, filter_ir_out=None, **kw):
        i1, i2 = Irreps(irreps_in1).simplify(), Irreps(irreps_in2).simplify()
        assert i1.num_irreps == i2.num_irreps
        i1, i2 = list(i1), list(i2)
        k = 0
        while k < len(i1):
            (m1, ir1), (m2, ir2) = i1[k], i2[k]
            if m1 < m2:
                i2[k] = (m1, ir2)
                i2.insert(k + 1, (m2 - m1, ir2))
            if m2 < m1:
                i1[k] = (m2, ir1)
                i1.insert(k + 1, (m1 - m2, ir1))
            k += 1
        out, instr = [], []
        for k, ((mul, ir1), (mul2, ir2)) in enumerate(zip(i1, i2)):
            assert mul == mul2
            for ir in Irrep(ir1) * Irrep(ir2):
                if filter_ir_out is not None and ir not in filter_ir_out:
                    continue
                instr.append((k, k, len(out), "uuu", False))
                out.append((mul, ir))
        super().__init__(Irreps(i1), Irreps(i2), Irreps(out), instr, **kw)


class Linear(torch.nn.Module):
    """e3nn o3.Linear, internal shared weights, no bias."""

    def __init__(self, irreps_in, irreps_out, internal_weights=True, shared_weights=True, **_):
        super().__init__()
        self.irreps_in = Irreps(irreps_in)
        self.irreps_out = Irreps(irreps_out)
        self.instructions = [
            (i, o)
            for i, (_, ir_in) in enumerate(self.irreps_in)
            for o, (_, ir_out) in enumerate(self.irreps_out)
            if ir_in == ir_out
        ]
        self.fan_in = {}
        for i, o in self.instructions:
            self.fan_in[o] = self.fan_in.get(o, 0) + self.irreps_in[i].mul
        self.weight_numel = sum(self.irreps_in[i].mul * self.irreps_out[o].mul for i, o in self.instructions)
        self.weight = torch.nn.Parameter(torch.randn(self.weight_numel))

    def forward(self, x: torch.Tensor, weight: torch.Tensor = None) -> torch.Tensor:
        w = self.weight if weight is None else weight
        Z = x.shape[0]
        s_in = self.irreps_in.slices()
        outs = [None] * len(self.irreps_out)
        off = 0
        for i, o in self.instructions:
            mi, ir = self.irreps_in[i]
            mo, _ = self.irreps_out[o]
            blk = w[off:off + mi * mo].reshape(mi, mo)
            off += mi * mo
            xi = x[:, s_in[i]].reshape(Z, mi, ir.dim)
            r = torch.einsum("uw,zui->zwi", blk, xi) * (1.0 / math.sqrt(self.fan_in[o]))
            r = r.reshape(Z, mo * ir.dim)
            outs[o] = r if outs[o] is None else outs[o] + r
        for k, (mo, iro) in enumerate(self.irreps_out):
            if outs[k] is None:
                outs[k] = x.new_zeros(Z, mo * iro.dim)
        return torch.cat(outs, dim=-1)


# --------------------------------------------------------------------------- #
# Rotations (test helpers: e3nn `rand_matrix`, Wigner D from the SH basis)
# --------------------------------------------------------------------------- #
def rand_matrix(*shape, generator=None, dtype=torch.float64) -> torch.Tensor:
    """Random proper rotation(s) via QR (test helper; e3nn samples Euler angles)."""
    a = torch.randn(*shape, 3, 3, generator=generator, dtype=dtype)
    q, r = torch.linalg.qr(a)
    q = q * torch.sign(torch.diagonal(r, dim1=-2, dim2=-1)).unsqueeze(-2)
    det = torch.linalg.det(q)
    q[..., :, 0] = q[..., :, 0] * det.unsqueeze(-1)
    return q


def wigner_D_from_R(l: int, R: torch.Tensor, parity: int = 1) -> torch.Tensor:
    """D^l(R) in the real SH basis of `_raw_sh`, by least squares on random points.

    For an improper R (det = -1) the irrep's parity p enters as p**1 on top of the
    polynomial action: D = (p * (-1)^l if det<0) * D_poly ... handled by the caller
    through `irrep_D`."""
    g = torch.Generator().manual_seed(1234 + l)
    pts = torch.nn.functional.normalize(torch.randn(64, 3, generator=g, dtype=torch.float64), dim=-1)
    Y = _raw_sh(max(l, 0), pts[:, 0], pts[:, 1], pts[:, 2])[:, l * l:(l + 1) * (l + 1)]
    pr = pts @ R.to(torch.float64).T
    Yr = _raw_sh(max(l, 0), pr[:, 0], pr[:, 1], pr[:, 2])[:, l * l:(l + 1) * (l + 1)]
    # Y(R x) = Y(x) D^T  ->  D^T = lstsq(Y, Yr)
    Dt = torch.linalg.lstsq(Y, Yr).solution
    return Dt.T


def irrep_D(ir, R: torch.Tensor) -> torch.Tensor:
    """Representation matrix of an O(3) element on irrep (l, p)."""
    ir = Irrep(ir)
    R = R.to(torch.float64)
    det = torch.linalg.det(R)
    if det < 0:
        return wigner_D_from_R(ir.l, -R) * ir.p
    return wigner_D_from_R(ir.l, R)


def irreps_D(irreps, R: torch.Tensor) -> torch.Tensor:
    irreps = Irreps(irreps)
    blocks = []
    for mul, ir in irreps:
        D = irrep_D(ir, R)
        blocks += [D] * mul
    return torch.block_diag(*blocks)
