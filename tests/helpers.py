"""Shared test helpers: golden-fixture loading and comparison, synthetic graphs."""
from __future__ import annotations

import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name: str) -> dict:
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """Normwise relative error max|a-b| / max(|b|, tiny): the north star's '1e-5 relative'."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.numel() == 0:
        return 0.0
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def check_against_digest(t: torch.Tensor, ref, tol: float, what: str = ""):
    """`ref` is a tensor or a digest dict written by tests/golden/make_golden.py."""
    if isinstance(ref, dict) and ref.get("digest"):
        assert tuple(t.shape) == tuple(ref["shape"]), what
        flat = t.detach().double().cpu().reshape(-1)
        sign = torch.randint(0, 2, (flat.numel(),), generator=_gen(flat.numel())).double() * 2 - 1
        scale = max(ref["norm"], 1e-30)
        assert abs(flat.norm().item() - ref["norm"]) <= tol * scale, (what, flat.norm().item(), ref["norm"])
        # sums of n terms: allow sqrt(n) growth of the per-element error
        slack = tol * scale * max(1.0, flat.numel() ** 0.5)
        assert abs(flat.sum().item() - ref["sum"]) <= slack, (what, "sum")
        assert abs((flat * sign).sum().item() - ref["proj"]) <= slack, (what, "proj")
        head = ref["head"].double()
        assert (flat[:64] - head).abs().max().item() <= tol * max(head.abs().max().item(), scale / flat.numel() ** 0.5), \
            (what, "head")
    else:
        e = rel_err(t, ref)
        assert e <= tol, (what, e)


def load_params(module: torch.nn.Module, state: dict):
    """Load a reference state_dict; every parameter of `module` must be present in it."""
    own = dict(module.named_parameters())
    missing = [k for k in own if k not in state]
    assert not missing, f"parameters absent from the reference state_dict: {missing}"
    module.load_state_dict(state, strict=False)
    return module


def densify(sp: dict) -> torch.Tensor:
    t = torch.zeros(sp["shape"], dtype=sp["val"].dtype)
    t[tuple(sp["idx"].long().T)] = sp["val"]
    return t


def random_clouds(num_graphs, nodes, box, r, seed, max_nb=64, atoms_hi=1):
    """Same generator as tests/golden/make_golden.py."""
    from oracle.thirdparty import cluster
    g = _gen(seed)
    pos = torch.rand(num_graphs * nodes, 3, generator=g) * box
    batch = torch.arange(num_graphs).repeat_interleave(nodes)
    ei = torch.from_numpy(cluster.radius_graph(pos.numpy(), r, batch.numpy(), False, max_nb))
    atoms = (torch.randint(1, atoms_hi, (num_graphs * nodes,), generator=g) if atoms_hi > 1
             else torch.zeros(num_graphs * nodes, dtype=torch.long))
    return dict(atoms=atoms, pos=pos, edge_index=ei, batch=batch)


class Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)
