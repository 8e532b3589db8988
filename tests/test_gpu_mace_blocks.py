"""the ACEsuit-style MACE interaction blocks (models/mace_modules/blocks.py:206-530) on the fused 'uvu'
kernels (csrc/uvu.cu), through the C ABI: golden fixtures made by the unmodified reference, and the oracle at the model
width on shuffled / ragged graphs."""
import pytest
import torch

from oracle import ref_layers as R
from tests.helpers import check_against_digest, load_golden, load_params, random_clouds, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5   # fp32 vs the reference (north star: 1e-5 relative)

KINDS = {"ResidualElementDependentInteractionBlock": "residual_element", "AgnosticNonlinearInteractionBlock": "agnostic_nonlinear",
         "AgnosticResidualNonlinearInteractionBlock": "agnostic_residual_nonlinear", "RealAgnosticInteractionBlock": "real_agnostic",
         "RealAgnosticResidualInteractionBlock": "real_agnostic_residual"}


def _outs(out):
    return [o for o in out if o is not None] if isinstance(out, tuple) else [out]


@pytest.mark.parametrize("cls", sorted(KINDS))
def test_interaction_block_golden(cls):
    import gmp_b200
    fx = load_golden("mace_interaction_" + cls)
    m = load_params(getattr(gmp_b200, cls)(**fx["ctor"]), fx["state"]).cuda()
    assert m.conv_tp.fused_C == 8 and str(m.conv_tp.irreps_out) == fx["extra"]["irreps_mid"]
    assert [tuple(t[:3]) for t in m.conv_tp.instructions] == [tuple(t) for t in fx["extra"]["instructions"]]
    i = {k: v.cuda() for k, v in fx["inputs"].items()}
    x, ef = i["node_feats"].requires_grad_(True), i["edge_feats"].requires_grad_(True)
    outs = _outs(m(i["node_attrs"], x, i["edge_attrs"], ef, i["edge_index"]))
    assert len(outs) == len(fx["outputs"])
    for o, ref in zip(outs, fx["outputs"]):
        assert rel_err(o, ref) <= TOL
    params = dict(m.named_parameters())
    loss = sum((o * c.cuda()).sum() for o, c in zip(outs, fx["cotangent"]))
    gs = torch.autograd.grad(loss, [x, ef] + list(params.values()))
    got = dict(zip(["input.node_feats", "input.edge_feats"] + [f"param.{k}" for k in params], gs))
    for k, ref in fx["grads"].items():
        check_against_digest(got[k].cpu(), ref, 10 * TOL, k)


@pytest.mark.parametrize("C,graphs,nodes,shuffle", [(128, 3, 40, True), (64, 1, 1, False), (32, 4, 33, True), (40, 2, 20, True)])
def test_real_agnostic_block_vs_oracle(C, graphs, nodes, shuffle):
    """RealAgnosticInteractionBlock (blocks.py:396-459) at the MACE widths; edge order shuffled (perm path of the CSR),
    a single isolated node (E = 0), nodes without in-edges."""
    import gmp_b200
    ir = f"{C}x0e+{C}x1o+{C}x2e"
    ctor = dict(node_attrs_irreps="4x0e", node_feats_irreps=ir, edge_attrs_irreps="1x0e+1x1o+1x2e", edge_feats_irreps="8x0e",
                target_irreps=ir, hidden_irreps=ir, avg_num_neighbors=9.0)
    torch.manual_seed(C)
    ref = R.InteractionBlock("real_agnostic", **ctor)
    mine = gmp_b200.RealAgnosticInteractionBlock(**ctor)
    load_params(mine, ref.state_dict())
    mine = mine.cuda()
    d = random_clouds(graphs, nodes, 3.0, 1.6, 7)
    ei = d["edge_index"]
    if shuffle:
        ei = ei[:, torch.randperm(ei.shape[1], generator=torch.Generator().manual_seed(3))]
    N, E = d["pos"].shape[0], ei.shape[1]
    g = torch.Generator().manual_seed(11)
    attrs = torch.eye(4)[torch.randint(0, 4, (N,), generator=g)]
    x = torch.randn(N, 9 * C, generator=g)
    vec = d["pos"][ei[0]] - d["pos"][ei[1]]
    esh = R.o3.SphericalHarmonics("1x0e+1x1o+1x2e", normalize=True, normalization="component")(vec) if E else torch.zeros(0, 9)
    eft = R.RadialEmbeddingBlock(2.0, 8, 5)(vec.norm(dim=-1, keepdim=True)) if E else torch.zeros(0, 8)
    cot = torch.randn(N, C, 9, generator=g)
    xr, er = x.clone().requires_grad_(True), eft.clone().requires_grad_(True)
    out_r, none_r = ref(attrs, xr, esh, er, ei)
    pr = dict(ref.named_parameters())
    gr = torch.autograd.grad((out_r * cot).sum(), [xr, er] + list(pr.values()), allow_unused=True)
    xm, em = x.cuda().requires_grad_(True), eft.cuda().requires_grad_(True)
    out_m, none_m = mine(attrs.cuda(), xm, esh.cuda(), em, ei.cuda())
    assert none_r is None and none_m is None
    assert rel_err(out_m, out_r) <= TOL
    pm = dict(mine.named_parameters())
    gm = torch.autograd.grad((out_m * cot.cuda()).sum(), [xm, em] + [pm[k] for k in pr], allow_unused=True)
    for name, a, b in zip(["node_feats", "edge_feats"] + list(pr), gm, gr):
        if b is None or E == 0 and a is None:
            continue
        assert rel_err(a, b) <= 10 * TOL, name


def test_uvu_conv_deterministic():
    import gmp_b200
    C = 128
    d = random_clouds(8, 48, 3.0, 1.5, 5)
    ei = d["edge_index"].cuda()
    N, E = d["pos"].shape[0], ei.shape[1]
    tp = gmp_b200.UVUTensorProduct(f"{C}x0e+{C}x1o+{C}x2e", "1x0e+1x1o+1x2e", f"{C}x0e+{C}x1o+{C}x2e").cuda()
    g = torch.Generator().manual_seed(1)
    x, sh, w = torch.randn(N, 9 * C, generator=g).cuda(), torch.randn(E, 9, generator=g).cuda(), torch.randn(E, 11 * C, generator=g).cuda()
    a, b = tp(x, ei, sh, w), tp(x, ei, sh, w)
    assert a.shape == (N, 35 * C) and torch.equal(a, b)


def test_interaction_block_equivariance():
    """RealAgnosticInteractionBlock on the fused uvu kernels: the output transforms with D^l under a random O(3) element
    (proper rotations and reflections), error not above 2x the oracle's (SURVEY.md 8c equivariance bar)."""
    import gmp_b200
    from oracle.thirdparty import o3
    C = 16
    ir = f"{C}x0e+{C}x1o+{C}x2e"
    ctor = dict(node_attrs_irreps="2x0e", node_feats_irreps=ir, edge_attrs_irreps="1x0e+1x1o+1x2e", edge_feats_irreps="8x0e",
                target_irreps=ir, hidden_irreps=ir, avg_num_neighbors=6.0)
    torch.manual_seed(4)
    ref = R.InteractionBlock("real_agnostic", **ctor)
    mine = load_params(gmp_b200.RealAgnosticInteractionBlock(**ctor), ref.state_dict()).cuda()
    d = random_clouds(3, 12, 3.0, 1.8, 9)
    pos, ei = d["pos"], d["edge_index"]
    n = pos.shape[0]
    g = torch.Generator().manual_seed(2)
    attrs = torch.eye(2)[torch.randint(0, 2, (n,), generator=g)]
    x = torch.randn(n, 9 * C, generator=g)
    shm, rad = o3.SphericalHarmonics(o3.Irreps("1x0e+1x1o+1x2e"), True, "component"), R.RadialEmbeddingBlock(2.0, 8, 5)
    rs = R.reshape_irreps(ir)
    inv_rs = lambda t: torch.cat([t[:, :, 0:1].reshape(n, -1), t[:, :, 1:4].reshape(n, -1), t[:, :, 4:9].reshape(n, -1)], dim=1)
    worst_r = worst_m = 0.0
    for seed in range(4):
        Rm = o3.rand_matrix(generator=torch.Generator().manual_seed(seed)).float() * (-1 if seed % 2 else 1)
        D = o3.irreps_D(ir, Rm.double()).float()

        def run(block, p, xx, dev):
            vec = p[ei[0]] - p[ei[1]]
            with torch.no_grad():
                out, _ = block(attrs.to(dev), xx.to(dev), shm(vec).to(dev), rad(vec.norm(dim=-1, keepdim=True)).to(dev), ei.to(dev))
            return inv_rs(out.cpu())     # [n, C, 9] -> e3nn layout, on which D acts

        worst_r = max(worst_r, rel_err(run(ref, pos @ Rm.T, x @ D.T, "cpu"), run(ref, pos, x, "cpu") @ D.T))
        worst_m = max(worst_m, rel_err(run(mine, pos @ Rm.T, x @ D.T, "cuda"), run(mine, pos, x, "cuda") @ D.T))
    assert worst_m <= max(2 * worst_r, 5e-6), (worst_m, worst_r)
