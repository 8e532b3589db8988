"""GPU parity: MACE symmetric contraction / product basis block / model (through the C ABI) vs the golden
vectors of the unmodified reference and the CPU oracle.  fp32; 1e-5 normwise relative."""
import pytest
import torch

from oracle import ref_layers as R
from oracle.thirdparty import o3
from tests.helpers import Bag, check_against_digest, load_golden, load_params, random_clouds, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def test_product_block_golden():
    import gmp_b200
    fx = load_golden("mace_product_block")
    m = load_params(gmp_b200.EquivariantProductBasisBlock(**fx["ctor"]), fx["state"]).cuda()
    i = fx["inputs"]
    x, sc = i["node_feats"].cuda().requires_grad_(True), i["sc"].cuda().requires_grad_(True)
    out = m(x, sc, None)
    assert rel_err(out, fx["outputs"][0]) <= TOL
    params = dict(m.named_parameters())
    gs = torch.autograd.grad((out * fx["cotangent"][0].cuda()).sum(), [x, sc] + list(params.values()))
    got = dict(zip(["input.node_feats", "input.sc"] + [f"param.{k}" for k in params], gs))
    for k, v in got.items():
        check_against_digest(v.cpu(), fx["grads"][k], 2 * TOL, k)


@pytest.mark.parametrize("C,N,corr", [(128, 200, 3), (16, 1000, 2), (8, 5, 1), (32, 777, 3), (3, 1, 3)])
def test_symmetric_contraction_vs_oracle(C, N, corr):
    """models/mace_modules/symmetric_contraction.py:81-85, 169-185 against the oracle.  correlation 3 on l <= 2 features is
    the model shape and takes the unrolled kernels (ragged node counts: 777 = 3 tiles + 9, a single node); the other
    cases take the generic ones."""
    import gmp_b200
    assert gmp_b200._lib.lib().gmp_symcontract_fast_path(C, 9, 9, {1: 9, 2: 54, 3: 219}[corr], 9 * C) == int(corr == 3)
    ir = f"{C}x0e+{C}x1o+{C}x2e"
    torch.manual_seed(C)
    ref = R.SymmetricContraction(ir, ir, corr)
    mine = gmp_b200.SymmetricContraction(ir, ir, corr)
    mine.load_state_dict(ref.state_dict())
    mine = mine.cuda()
    for (k, a), (_, b) in zip(sorted(ref.named_buffers()), sorted((k, v) for k, v in mine.named_buffers() if "U_matrix" in k)):
        assert rel_err(b, a) <= 1e-6, k
    x = torch.randn(N, C, 9, generator=torch.Generator().manual_seed(1))
    cot = torch.randn(N, 9 * C, generator=torch.Generator().manual_seed(2))
    xr = x.clone().requires_grad_(True)
    out_r = ref(xr, None)
    gr = torch.autograd.grad((out_r * cot).sum(), [xr] + list(ref.parameters()))
    xc = x.cuda().requires_grad_(True)
    out = mine(xc, None)
    gm = torch.autograd.grad((out * cot.cuda()).sum(), [xc] + list(mine.parameters()))
    assert rel_err(out, out_r) <= TOL
    for a, b, name in zip(gm, gr, ["x"] + [k for k, _ in ref.named_parameters()]):
        assert rel_err(a, b) <= 2 * TOL, name
    out2 = mine(xc, None)
    gm2 = torch.autograd.grad((out2 * cot.cuda()).sum(), [xc] + list(mine.parameters()))
    assert torch.equal(out, out2) and all(torch.equal(a, b) for a, b in zip(gm, gm2))


def test_mace_model_golden():
    import gmp_b200
    fx = load_golden("mace_model")
    m = load_params(gmp_b200.MACEModel(**fx["ctor"]), fx["state"]).cuda()
    m.train()
    i = fx["inputs"]
    b = Bag(atoms=i["atoms"].cuda(), pos=i["pos"].cuda(), edge_index=i["edge_index"].cuda(), batch=i["batch"].cuda())
    out = m(b)
    assert rel_err(out, fx["outputs"][0]) <= 5 * TOL
    params = dict(m.named_parameters())
    gs = torch.autograd.grad((out * fx["cotangent"][0].cuda()).sum(), list(params.values()), allow_unused=True)
    for (k, _), g in zip(params.items(), gs):
        ref = fx["grads"][f"param.{k}"]
        if ref is None:
            continue
        check_against_digest(g.cpu(), ref, 1e-4, k)


def test_mace_conv_equivariance():
    """Fused TP conv + product block transform with D^l under a random O(3) element, error <= 2x the oracle's."""
    import gmp_b200
    d = random_clouds(3, 10, 3.0, 2.0, 31)
    pos, ei = d["pos"], d["edge_index"]
    n = pos.shape[0]
    hid = "8x0e+8x1o+8x2e"
    sh_ir = "1x0e+1x1o+1x2e"
    torch.manual_seed(0)
    refc = R.TensorProductConvLayer(hid, hid, sh_ir, 8, 64)
    refp = R.EquivariantProductBasisBlock(hid, hid, 3, element_dependent=False, use_sc=False)
    minec = gmp_b200.TensorProductConvLayer(hid, hid, sh_ir, 8, 64)
    minec.load_state_dict(refc.state_dict(), strict=False)
    minep = gmp_b200.EquivariantProductBasisBlock(hid, hid, 3, element_dependent=False, use_sc=False)
    minep.load_state_dict(refp.state_dict(), strict=False)
    minec, minep = minec.cuda(), minep.cuda()
    rad_r, shm = R.RadialEmbeddingBlock(2.0, 8, 5), o3.SphericalHarmonics(o3.Irreps(sh_ir), True, "component")
    rad = gmp_b200.RadialEmbeddingBlock(2.0, 8, 5)
    rs_r, rs_m = R.reshape_irreps(hid), gmp_b200.reshape_irreps(hid)
    x = torch.randn(n, 72)
    worst_r = worst_m = 0.0
    for seed in range(4):
        gen = torch.Generator().manual_seed(seed)
        Rm = o3.rand_matrix(generator=gen).float() * (-1 if seed % 2 else 1)
        D = o3.irreps_D(hid, Rm.double()).float()
        pos2 = pos @ Rm.T
        with torch.no_grad():
            def run_ref(p, xx):
                s, f = R.edge_geometry(p, ei, shm, rad_r)
                return refp(rs_r(refc(xx, ei, s, f)), None, None)

            def run_mine(p, xx):
                s, f = gmp_b200.edge_geometry(p.cuda(), ei.cuda(), 2, rad)
                return minep(rs_m(minec(xx.cuda(), ei.cuda(), s, f)), None, None).cpu()
            worst_r = max(worst_r, rel_err(run_ref(pos2, x @ D.T), run_ref(pos, x) @ D.T))
            worst_m = max(worst_m, rel_err(run_mine(pos2, x @ D.T), run_mine(pos, x) @ D.T))
    assert worst_m <= max(2 * worst_r, 5e-6), (worst_m, worst_r)
