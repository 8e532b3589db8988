"""Product-side O(3) bookkeeping: irreps strings, real Clebsch-Gordan tensors, tensor-product path tables.

The reference passes ``e3nn.o3.Irreps`` objects around (models/tfn.py:110-128, models/mace.py:82-149);
the fused layers accept those or their ``str()`` ("64x0e+64x1o+64x2e").  Everything here is
constructor-time host code (numpy, float64); the kernels consume the flat tables it produces.
Conventions are e3nn's: irrep blocks concatenated, each [mul, 2l+1] row-major; real
basis ordered m = -l..l with the l=1 block equal to (x, y, z); ``wigner_3j`` normalised to unit Frobenius
norm; FullyConnectedTensorProduct paths 'uvw' with 'component' / 'element' normalisation.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from fractions import Fraction
from functools import lru_cache
from typing import List, Tuple

import numpy as np


@dataclass(frozen=True)
class Irrep:
    l: int
    p: int  # +1 even, -1 odd

    @property
    def dim(self) -> int:
        return 2 * self.l + 1

    def __str__(self):
        return f"{self.l}{'e' if self.p == 1 else 'o'}"


class Irreps:
    """Ordered list of (mul, Irrep)."""

    def __init__(self, spec):
        self.items: List[Tuple[int, Irrep]] = []
        if isinstance(spec, Irreps):
            self.items = list(spec.items)
            return
        text = str(spec).strip()
        if text:
            for piece in text.split("+"):
                piece = piece.strip()
                mul, ir = piece.split("x") if "x" in piece else ("1", piece)
                self.items.append((int(mul), Irrep(int(ir[:-1]), 1 if ir[-1] == "e" else -1)))

    def __iter__(self):
        return iter(self.items)

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]

    @property
    def dim(self) -> int:
        return sum(m * ir.dim for m, ir in self.items)

    @property
    def num_irreps(self) -> int:
        return sum(m for m, _ in self.items)

    def offsets(self) -> List[int]:
        out, o = [], 0
        for m, ir in self.items:
            out.append(o)
            o += m * ir.dim
        return out

    def simplify(self) -> "Irreps":
        out: List[Tuple[int, Irrep]] = []
        for m, ir in self.items:
            if out and out[-1][1] == ir:
                out[-1] = (out[-1][0] + m, ir)
            elif m > 0:
                out.append((m, ir))
        r = Irreps("")
        r.items = out
        return r

    def __add__(self, other):
        r = Irreps("")
        r.items = self.items + Irreps(other).items
        return r

    def __str__(self):
        return "+".join(f"{m}x{ir}" for m, ir in self.items)

    __repr__ = __str__

    @staticmethod
    def spherical_harmonics(lmax: int) -> "Irreps":
        return Irreps("+".join(f"1x{l}{'e' if l % 2 == 0 else 'o'}" for l in range(lmax + 1)))


def hidden_irreps(max_ell: int, emb_dim: int) -> Irreps:
    """(sh_irreps * emb_dim).sort().simplify() of models/tfn.py:119-122."""
    return Irreps("+".join(f"{emb_dim}x{l}{'e' if l % 2 == 0 else 'o'}" for l in range(max_ell + 1)))


# ------------------------------------------------------------------------------------------------
# real Wigner 3j: SU(2) Clebsch-Gordan (Racah's formula) rotated into the real basis
# ------------------------------------------------------------------------------------------------
def _cg_su2(j1, m1, j2, m2, j3, m3) -> float:
    if m3 != m1 + m2:
        return 0.0
    f = math.factorial
    pref = Fraction((2 * j3 + 1) * f(j3 + j1 - j2) * f(j3 - j1 + j2) * f(j1 + j2 - j3) * f(j3 + m3) * f(j3 - m3),
                    f(j1 + j2 + j3 + 1) * f(j1 - m1) * f(j1 + m1) * f(j2 - m2) * f(j2 + m2))
    lo = max(-j1 + j2 + m3, -j1 + m1, 0)
    hi = min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3)
    total = Fraction(0)
    for v in range(lo, hi + 1):
        total += (-1) ** (v + j2 + m2) * Fraction(f(j2 + j3 + m1 - v) * f(j1 - m1 + v),
                                                  f(v) * f(j3 - j1 + j2 - v) * f(j3 + m3 - v) * f(v + j1 - j2 - m3))
    return float(pref) ** 0.5 * float(total)


def _real_basis(l: int) -> np.ndarray:
    q = np.zeros((2 * l + 1, 2 * l + 1), dtype=np.complex128)
    s = 1 / math.sqrt(2)
    for m in range(-l, 0):
        q[l + m, l + abs(m)] = s
        q[l + m, l - abs(m)] = -1j * s
    q[l, l] = 1
    for m in range(1, l + 1):
        q[l + m, l + abs(m)] = (-1) ** m * s
        q[l + m, l - abs(m)] = 1j * (-1) ** m * s
    return (-1j) ** l * q


@lru_cache(maxsize=None)
def wigner_3j(l1: int, l2: int, l3: int) -> np.ndarray:
    """Real, unit-Frobenius-norm coupling tensor [2l1+1, 2l2+1, 2l3+1] (e3nn's o3.wigner_3j)."""
    assert abs(l2 - l3) <= l1 <= l2 + l3
    c = np.zeros((2 * l1 + 1, 2 * l2 + 1, 2 * l3 + 1))
    for m1 in range(-l1, l1 + 1):
        for m2 in range(-l2, l2 + 1):
            if abs(m1 + m2) <= l3:
                c[l1 + m1, l2 + m2, l3 + m1 + m2] = _cg_su2(l1, m1, l2, m2, l3, m1 + m2)
    q1, q2, q3 = _real_basis(l1), _real_basis(l2), _real_basis(l3)
    r = np.einsum("ij,kl,mn,ikn->jlm", q1, q2, np.conj(q3.T), c.astype(np.complex128))
    assert np.abs(r.imag).max() < 1e-9
    r = r.real
    return r / np.linalg.norm(r)


# ------------------------------------------------------------------------------------------------
# FullyConnectedTensorProduct path table
# ------------------------------------------------------------------------------------------------
@dataclass
class TPPath:
    i_in: int
    i_sh: int
    i_out: int
    mul_in: int
    mul_out: int
    l_in: int
    l_sh: int
    l_out: int
    in_off: int      # float offset of the input block in a node row
    sh_off: int      # float offset of the sh block in an edge_sh row
    out_off: int     # float offset of the output block
    w_off: int       # offset of this path's [mul_in, 1, mul_out] weight block in the per-edge weight vector
    coeff: float     # sqrt((2 l_out + 1) / sum_{paths into the same output} mul_in * mul_sh)


def fctp_paths(irreps_in, irreps_sh, irreps_out) -> Tuple[List[TPPath], int]:
    """Instructions of e3nn.o3.FullyConnectedTensorProduct(in, sh, out, shared_weights=False), in e3nn's
    order (for in, for sh, for out), with their normalisation; returns (paths, weight_numel)."""
    a, b, c = Irreps(irreps_in), Irreps(irreps_sh), Irreps(irreps_out)
    ao, bo, co = a.offsets(), b.offsets(), c.offsets()
    raw = []
    for i, (m1, ir1) in enumerate(a):
        for j, (m2, ir2) in enumerate(b):
            for k, (mo, iro) in enumerate(c):
                if iro.p == ir1.p * ir2.p and abs(ir1.l - ir2.l) <= iro.l <= ir1.l + ir2.l:
                    raw.append((i, j, k))
    fan = {}
    for i, j, k in raw:
        fan[k] = fan.get(k, 0) + a[i][0] * b[j][0]
    paths, w = [], 0
    for i, j, k in raw:
        (m1, ir1), (m2, ir2), (mo, iro) = a[i], b[j], c[k]
        if m2 != 1:
            raise NotImplementedError("the fused tensor product handles sh irreps of multiplicity 1 (spherical harmonics)")
        paths.append(TPPath(i, j, k, m1, mo, ir1.l, ir2.l, iro.l, ao[i], bo[j], co[k], w, math.sqrt(iro.dim / fan[k])))
        w += m1 * m2 * mo
    return paths, w


def gate_split(irreps) -> Tuple[Irreps, Irreps, Irreps]:
    """models/mace_modules/irreps_tools.py:82-97."""
    irreps = Irreps(irreps)
    scal, gated = Irreps(""), Irreps("")
    scal.items = [(m, ir) for m, ir in irreps if ir.l == 0 and ir.p == 1]
    gated.items = [(m, ir) for m, ir in irreps if not (ir.l == 0 and ir.p == 1)]
    scal, gated = scal.simplify(), gated.simplify()
    gates = Irreps("")
    gates.items = [(m, Irrep(0, 1)) for m, _ in gated]
    return scal, gates.simplify(), gated


# e3nn normalize2mom constants (1 / sqrt(E_{z~N(0,1)} f(z)^2), e3nn's seeded 1e6-sample estimate)
NORM2MOM = {"silu": 1.6791767923989418, "sigmoid": 1.8467055342154763}


# ------------------------------------------------------------------------------------------------
# MACE generalised Clebsch-Gordan ("U matrices", models/mace_modules/cg.py:19-133) and their symmetric
# monomial form
# ------------------------------------------------------------------------------------------------
def _coupled_bases(irreps_list: List[Irreps]):
    """Iterated coupling of the components of irreps_list[0] x ... x irreps_list[-1] ('component' normalisation).
    Returns [(Irrep, basis [d_out, dim_1, ..., dim_n])], stably sorted by (l, p) at every level, as the reference does."""
    if len(irreps_list) == 1:
        (irreps,) = irreps_list
        eye, out, i = np.eye(irreps.dim), [], 0
        for mul, ir in irreps:
            for _ in range(mul):
                out.append((ir, eye[i:i + ir.dim]))
                i += ir.dim
        return out
    *left, right = irreps_list
    out = []
    for ir_left, C_left in _coupled_bases(left):
        i = 0
        for mul, ir in right:
            for l in range(abs(ir_left.l - ir.l), ir_left.l + ir.l + 1):
                ir_out = Irrep(l, ir_left.p * ir.p)
                C = wigner_3j(ir_out.l, ir_left.l, ir.l) * math.sqrt(ir_out.dim)
                C = np.einsum("jk,ijl->ikl", C_left.reshape(C_left.shape[0], -1), C)
                C = C.reshape((ir_out.dim,) + tuple(x.dim for x in left) + (ir.dim,))
                for u in range(mul):
                    E = np.zeros((ir_out.dim,) + tuple(x.dim for x in left) + (right.dim,))
                    E[..., i + u * ir.dim:i + (u + 1) * ir.dim] = C
                    out.append((ir_out, E))
            i += mul * ir.dim
    return sorted(out, key=lambda t: (t[0].l, t[0].p))  # Python's sort is stable, like the reference's


def u_matrix_real(irreps_in, ir_out: Irrep, correlation: int) -> np.ndarray:
    """U_nu for one output irrep: [d_out, dim, ..., dim (nu times), k] (not squeezed), float64."""
    irreps_in = Irreps(irreps_in)
    bases = [B for ir, B in _coupled_bases([irreps_in] * correlation) if ir == ir_out]
    return np.stack(bases, axis=-1)


def monomials(D: int, nu: int) -> np.ndarray:
    """Symmetric monomials of degree 1..nu in D variables as index triples padded with D (the constant 1)."""
    out = []
    for deg in range(1, nu + 1):
        if deg == 1:
            out += [(i, D, D) for i in range(D)]
        elif deg == 2:
            out += [(i, j, D) for i in range(D) for j in range(i, D)]
        else:
            out += [(i, j, k) for i in range(D) for j in range(i, D) for k in range(j, D)]
    return np.asarray(out, dtype=np.int32)


def symmetrise_u(U: np.ndarray, nu: int, D: int) -> np.ndarray:
    """Sum a U_nu [d_out, D^nu, k] over index permutations onto the degree-nu monomials: [d_out, n_mono(nu), k]."""
    mons = [m for m in monomials(D, nu) if (m != D).sum() == nu]
    index = {tuple(m[:nu]): n for n, m in enumerate(mons)}
    out = np.zeros((U.shape[0], len(mons), U.shape[-1]))
    for idx in np.ndindex(*([D] * nu)):
        out[:, index[tuple(sorted(idx))], :] += U[(slice(None),) + idx + (slice(None),)]
    return out
