"""Closed-form / property tests pinning the third-party semantics the oracle restates
(SURVEY.md §4 item 2, Appendix A).  CPU only, fp64 where it matters."""
import math

import pytest
import torch

from oracle import ref_layers as R
from oracle.thirdparty import cluster, e3nn_nn, o3, pyg, scatter
from tests.helpers import Bag, random_clouds


def _unit(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn(n, 3, generator=g, dtype=torch.float64), dim=-1)


def test_wigner_3j_known_values():
    eps = torch.zeros(3, 3, 3, dtype=torch.float64)
    for i, j, k, s in [(0, 1, 2, 1), (1, 2, 0, 1), (2, 0, 1, 1), (0, 2, 1, -1), (2, 1, 0, -1), (1, 0, 2, -1)]:
        eps[i, j, k] = s
    assert (o3.wigner_3j(1, 1, 1, dtype=torch.float64) - eps / math.sqrt(6)).abs().max() < 1e-14
    assert (o3.wigner_3j(1, 1, 0, dtype=torch.float64)[..., 0] - torch.eye(3, dtype=torch.float64) / math.sqrt(3)).abs().max() < 1e-14
    for l in (0, 1, 2):
        w = o3.wigner_3j(l, 0, l, dtype=torch.float64)[:, 0, :]
        assert (w - torch.eye(2 * l + 1, dtype=torch.float64) / math.sqrt(2 * l + 1)).abs().max() < 1e-14


@pytest.mark.parametrize("ls", [(1, 1, 0), (1, 1, 1), (1, 1, 2), (2, 2, 2), (1, 2, 1), (2, 2, 0), (2, 1, 2), (2, 2, 1)])
def test_wigner_3j_invariance(ls):
    R_ = o3.rand_matrix(generator=torch.Generator().manual_seed(5))
    Da, Db, Dc = [o3.wigner_D_from_R(l, R_) for l in ls]
    C = o3.wigner_3j(*ls, dtype=torch.float64)
    assert (torch.einsum("ia,jb,kc,abc->ijk", Da, Db, Dc, C) - C).abs().max() < 1e-12
    assert abs(C.norm().item() - 1) < 1e-12


def test_spherical_harmonics_norm_and_cg_consistency():
    v = _unit(500)
    Y = o3._raw_sh(2, v[:, 0], v[:, 1], v[:, 2])
    for l in range(3):
        assert (Y[:, l * l:(l + 1) ** 2].pow(2).sum(-1) - 1).abs().max() < 1e-12
    y2 = torch.einsum("ijk,zi,zj->zk", o3.wigner_3j(1, 1, 2, dtype=torch.float64), v, v)
    assert (y2 - math.sqrt(2 / 15) * Y[:, 4:9]).abs().max() < 1e-12
    sh = o3.SphericalHarmonics("1x0e+1x1o+1x2e", True, "component")(3.7 * v)
    assert (sh[:, 1:4] - math.sqrt(3) * v).abs().max() < 1e-12
    assert (sh[:, 4:9].pow(2).sum(-1) - 5).abs().max() < 1e-12
    assert sh[:, 0].eq(1).all()
    assert o3.SphericalHarmonics("1x0e+1x1o+1x2e", True, "component")(torch.zeros(1, 3))[0, 1:].abs().max() == 0


def test_normalize2mom_constants():
    assert abs(e3nn_nn.normalize2mom(torch.nn.functional.silu).cst - 1.6791767923989418) < 1e-12
    assert abs(e3nn_nn.normalize2mom(torch.sigmoid).cst - 1.8467055342154763) < 1e-12


def test_irreps_algebra():
    sh = o3.Irreps.spherical_harmonics(2)
    assert str(sh) == "1x0e+1x1o+1x2e"
    hid = (sh * 64).sort()[0].simplify()
    assert str(hid) == "64x0e+64x1o+64x2e" and hid.dim == 576
    assert str(o3.Irreps("0e")) == "1x0e"
    s, g, gd = R.irreps2gate(hid)
    assert (str(s), str(g), str(gd)) == ("64x0e", "128x0e", "64x1o+64x2e")
    gate = e3nn_nn.Gate(s, [torch.nn.functional.silu], g, [torch.sigmoid], gd)
    assert str(gate.irreps_in) == "192x0e+64x1o+64x2e" and str(gate.irreps_out) == "64x0e+64x1o+64x2e"


def test_weight_numel_at_baseline_configs():
    sh = o3.Irreps.spherical_harmonics(2)
    tfn = R.TensorProductConvLayer("64x0e+64x1o+64x2e", "64x0e+64x1o+64x2e", sh, 8, 256, gate=True)
    assert tfn.tp.weight_numel == 69632 and len(tfn.tp.instructions) == 11
    tfn0 = R.TensorProductConvLayer("64x0e", "64x0e+64x1o+64x2e", sh, 8, 256, gate=True)
    assert tfn0.tp.weight_numel == 20480
    c = {i.i_out: i.path_weight for i in tfn.tp.instructions}
    assert abs(c[0] - math.sqrt(1 / (3 * 64))) < 1e-12 and abs(c[1] - math.sqrt(3 / (4 * 64))) < 1e-12
    assert abs(c[2] - math.sqrt(5 / (4 * 64))) < 1e-12


def test_gate_is_plain_product():
    gate = e3nn_nn.Gate("2x0e", [torch.nn.functional.silu], "2x0e", [torch.sigmoid], "1x1o+1x2e")
    x = torch.randn(5, 2 + 2 + 3 + 5, dtype=torch.float64)
    y = gate(x)
    cs, cg = 1.6791767923989418, 1.8467055342154763
    assert (y[:, :2] - cs * torch.nn.functional.silu(x[:, :2])).abs().max() < 1e-12
    assert (y[:, 2:5] - x[:, 4:7] * (cg * torch.sigmoid(x[:, 2:3]))).abs().max() < 1e-12
    assert (y[:, 5:10] - x[:, 7:12] * (cg * torch.sigmoid(x[:, 3:4]))).abs().max() < 1e-12


def test_scatter_semantics():
    src = torch.arange(12.0).reshape(6, 2)
    idx = torch.tensor([0, 0, 3, 3, 3, 1])
    out = scatter.scatter(src, idx, dim=0, reduce="sum")
    assert out.shape == (4, 2) and out[2].abs().sum() == 0  # no dim_size -> max+1 rows, empty row = 0
    mean = scatter.scatter(src, idx, dim=-2, reduce="mean", dim_size=6)
    assert mean.shape == (6, 2) and torch.equal(mean[3], src[2:5].mean(0)) and mean[5].abs().sum() == 0


def test_radius_graph_canonical_order_and_truncation():
    g = torch.Generator().manual_seed(0)
    pos = torch.rand(40, 3, generator=g).numpy() * 2
    batch = torch.arange(2).repeat_interleave(20).numpy()
    ei = cluster.radius_graph(pos, 0.9, batch, False, 64)
    key = ei[1] * 40 + ei[0]
    assert (key[1:] > key[:-1]).all()  # dst-major, src ascending
    assert (batch[ei[0]] == batch[ei[1]]).all() and (ei[0] != ei[1]).all()
    s = set(map(tuple, ei.T.tolist()))
    assert all((b, a) in s for a, b in s)  # symmetric when nothing is truncated
    ei3 = cluster.radius_graph(pos, 5.0, batch, False, 3)  # everything within r: first hits in index order
    for q in range(40):
        srcs = ei3[0][ei3[1] == q].tolist()
        lo = 20 * (q // 20)
        first4 = list(range(lo, lo + 4))
        assert srcs == [c for c in first4 if c != q][: (3 if q in first4 else 4)]


def _rot_case(seed):
    g = torch.Generator().manual_seed(seed)
    Rm = o3.rand_matrix(generator=g)
    if seed % 2:
        Rm = -Rm  # improper
    return Rm, torch.randn(3, generator=g, dtype=torch.float64)


@pytest.mark.parametrize("seed", [0, 1])
def test_oracle_equivariance_fp64(seed):
    """EGNN (h invariant, pos equivariant), SchNet (invariant), TFN/MACE conv (D^l-equivariant)."""
    torch.manual_seed(0)
    d = random_clouds(3, 10, 3.0, 2.0, 31)
    pos, ei = d["pos"].double(), d["edge_index"]
    Rm, t = _rot_case(seed)
    pos2 = pos @ Rm.T + t
    n = pos.shape[0]
    egnn = R.EGNNLayer(16).double()
    h = torch.randn(n, 16, dtype=torch.float64)
    h1, p1 = egnn(h, pos, ei)
    h2, p2 = egnn(h, pos2, ei)
    assert (h1 - h2).abs().max() < 1e-10 and (p1 @ Rm.T + t - p2).abs().max() < 1e-10

    sch = R.SchNetModel(hidden_channels=16, num_filters=16, num_layers=2, cutoff=5.0).double()
    b1 = Bag(atoms=torch.ones(n, dtype=torch.long), pos=pos, edge_index=ei, batch=d["batch"])
    b2 = Bag(atoms=b1.atoms, pos=pos2, edge_index=ei, batch=d["batch"])
    assert (sch(b1) - sch(b2)).abs().max() < 1e-10

    hid = o3.Irreps("4x0e+4x1o+4x2e")
    shm = o3.SphericalHarmonics(o3.Irreps.spherical_harmonics(2), True, "component")
    rad = R.RadialEmbeddingBlock(2.0, 8, 5)
    for kw in (dict(gate=True), dict(gate=False, batch_norm=True)):
        conv = R.TensorProductConvLayer(hid, hid, shm.irreps_out, 8, 16, **kw).double()
        x = torch.randn(n, hid.dim, dtype=torch.float64)
        D_in = o3.irreps_D(hid, Rm)
        D_out = o3.irreps_D(conv.gate.irreps_out if conv.gate is not None else conv.out_irreps, Rm)
        s1, r1 = R.edge_geometry(pos, ei, shm, rad)
        s2, r2 = R.edge_geometry(pos2, ei, shm, rad)
        y1 = conv(x, ei, s1, r1)
        y2 = conv(x @ D_in.T, ei, s2, r2)
        assert (y1 @ D_out.T - y2).abs().max() < 1e-9
    prod = R.EquivariantProductBasisBlock(hid, hid, 3, element_dependent=False, use_sc=False).double()
    for c in prod.symmetric_contractions.contractions.values():
        for nu in (1, 2, 3):
            setattr(c, f"U_matrix_{nu}", R.u_matrix_real("1x0e+1x1o+1x2e", c.U_matrix_1.shape[0] == 9 and "0e" or
                                                         {3: "1o", 5: "2e"}[c.U_matrix_1.shape[0]], nu, torch.float64))
    x = torch.randn(n, hid.dim, dtype=torch.float64)
    D = o3.irreps_D(hid, Rm)
    y1 = prod(R.reshape_irreps_fn(x, hid), None, None)
    y2 = prod(R.reshape_irreps_fn(x @ D.T, hid), None, None)
    assert (y1 @ D.T - y2).abs().max() < 1e-9


def test_permutation_equivariance_egnn():
    torch.manual_seed(1)
    d = random_clouds(2, 8, 3.0, 2.0, 33)
    n = d["pos"].shape[0]
    perm = torch.randperm(n)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(n)
    layer = R.EGNNLayer(8).double()
    h, pos = torch.randn(n, 8, dtype=torch.float64), d["pos"].double()
    h1, p1 = layer(h, pos, d["edge_index"])
    h2, p2 = layer(h[perm], pos[perm], inv[d["edge_index"]])
    assert (h1[perm] - h2).abs().max() < 1e-12 and (p1[perm] - p2).abs().max() < 1e-12
