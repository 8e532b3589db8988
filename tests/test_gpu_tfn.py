"""GPU parity: fused TFN / MACE tensor-product convolution (through the C ABI) vs the golden vectors of the
unmodified reference and the CPU oracle.  fp32 strict; tolerance 1e-5 normwise relative (2e-5 for parameter
gradients, which are sums over all edges of products of O(1) terms)."""
import pytest
import torch

from oracle import ref_layers as R
from oracle.thirdparty import o3
from tests.helpers import Bag, check_against_digest, load_golden, load_params, random_clouds, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize("name", ["tfn_conv_first", "tfn_conv_hidden", "mace_conv_hidden", "tfn_conv_mean_nogate"])
def test_tp_conv_layer_golden(name):
    import gmp_b200
    fx = load_golden(name)
    m = load_params(gmp_b200.TensorProductConvLayer(**fx["ctor"]), fx["state"]).cuda()
    m.train()
    assert str(m.out_irreps) == fx["extra"]["tp_out_irreps"] and m.tp.weight_numel == fx["extra"]["weight_numel"]
    i = fx["inputs"]
    x = i["node_attr"].cuda().requires_grad_(True)
    out = m(x, i["edge_index"].cuda(), i["edge_sh"].cuda(), i["edge_feat"].cuda())
    assert rel_err(out, fx["outputs"][0]) <= TOL
    params = dict(m.named_parameters())
    gs = torch.autograd.grad((out * fx["cotangent"][0].cuda()).sum(), [x] + list(params.values()))
    got = dict(zip(["input.node_attr"] + [f"param.{k}" for k in params], gs))
    for k, v in got.items():
        check_against_digest(v.cpu(), fx["grads"][k], 2 * TOL, k)
    if fx["extra"]["bn_running_var"] is not None:
        assert rel_err(m.batch_norm.running_var, fx["extra"]["bn_running_var"]) <= TOL


def test_edge_geometry_kernel_golden():
    import gmp_b200
    fx = load_golden("edge_geometry")
    vec = fx["inputs"]["vec"]
    n = vec.shape[0]
    pos = torch.cat([vec, torch.zeros(1, 3)]).cuda()  # edge e: src = e, dst = n (origin) -> pos[src]-pos[dst] = vec[e]
    ei = torch.stack([torch.arange(n), torch.full((n,), n)]).cuda()
    sh, rbf = gmp_b200.edge_geometry(pos, ei, 2, gmp_b200.RadialEmbeddingBlock(2.0, 8, 5))
    assert rel_err(sh, fx["outputs"]["sh"]) <= 2e-6 and rel_err(rbf, fx["outputs"]["rbf"]) <= 5e-6
    assert float(rbf[2].abs().max()) == 0.0


@pytest.mark.parametrize("C,gate,bn,shuffle", [(64, True, False, True), (16, False, True, False)])
def test_tp_conv_vs_oracle_random(C, gate, bn, shuffle):
    """TFN config width (C = 64: weight_numel 69 632) on a small cloud; shuffled edge_index exercises perm."""
    import gmp_b200
    d = random_clouds(3, 12, 3.0, 1.9, 300 + C)
    ei, pos = d["edge_index"], d["pos"]
    if shuffle:
        ei = ei[:, torch.randperm(ei.shape[1], generator=torch.Generator().manual_seed(1))]
    n = pos.shape[0]
    assert int(ei[0].max()) == n - 1
    hid = f"{C}x0e+{C}x1o+{C}x2e"
    sh_ir = "1x0e+1x1o+1x2e"
    torch.manual_seed(C)
    ref = R.TensorProductConvLayer(hid, hid, sh_ir, 8, 256, gate=gate, batch_norm=bn)
    shm = o3.SphericalHarmonics(o3.Irreps(sh_ir), True, "component")
    esh, eft = R.edge_geometry(pos, ei, shm, R.RadialEmbeddingBlock(2.0, 8, 5))
    x = torch.randn(n, 9 * C, generator=torch.Generator().manual_seed(2))
    xr = x.clone().requires_grad_(True)
    out_r = ref(xr, ei, esh, eft)
    cot = torch.randn(out_r.shape, generator=torch.Generator().manual_seed(3))
    gr = torch.autograd.grad((out_r * cot).sum(), [xr] + list(ref.parameters()))

    mine = gmp_b200.TensorProductConvLayer(hid, hid, sh_ir, 8, 256, gate=gate, batch_norm=bn)
    mine.load_state_dict(ref.state_dict(), strict=False)
    mine = mine.cuda()
    esh_c, eft_c = gmp_b200.edge_geometry(pos.cuda(), ei.cuda(), 2, gmp_b200.RadialEmbeddingBlock(2.0, 8, 5))
    assert rel_err(esh_c, esh) <= 2e-6 and rel_err(eft_c, eft) <= 5e-6
    xc = x.cuda().requires_grad_(True)
    out = mine(xc, ei.cuda(), esh_c, eft_c)
    gm = torch.autograd.grad((out * cot.cuda()).sum(), [xc] + list(mine.parameters()))
    assert rel_err(out, out_r) <= TOL
    for a, b_, name in zip(gm, gr, ["node_attr"] + [k for k, _ in ref.named_parameters()]):
        assert rel_err(a, b_) <= 2 * TOL, name
    out2 = mine(xc, ei.cuda(), esh_c, eft_c)
    assert torch.equal(out, out2)


def test_tfn_model_golden():
    import gmp_b200
    fx = load_golden("tfn_model")
    m = load_params(gmp_b200.TFNModel(**fx["ctor"]), fx["state"]).cuda()
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


@pytest.mark.parametrize("n,C", [(1, 8), (333, 64), (5000, 16)])
def test_fused_gate_matches_oracle_gate(n, C):
    """e3nn Gate as built by models/layers/tfn_layer.py:45-63 (silu scalars, sigmoid gates): the fused kernels (csrc/gate.cu)
    against the oracle's Gate, forward and input gradient; and the all-scalar Activation (:52-53)."""
    import gmp_b200
    from oracle.thirdparty import e3nn_nn, o3
    from oracle import ref_layers as R
    scal, gates, gated = R.irreps2gate(o3.Irreps(f"{C}x0e+{C}x1o+{C}x2e"))
    ref = e3nn_nn.Gate(scal, [torch.nn.functional.silu], gates, [torch.sigmoid], gated)
    mine = gmp_b200.Gate(str(scal), str(gates), str(gated)).cuda()
    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, ref.irreps_in.dim, generator=g) * 2
    cot = torch.randn(n, ref.irreps_out.dim, generator=g)
    xr = x.clone().requires_grad_(True)
    out_r = ref(xr)
    (gr,) = torch.autograd.grad((out_r * cot).sum(), xr)
    xm = x.cuda().requires_grad_(True)
    out_m = mine(xm)
    (gm,) = torch.autograd.grad((out_m * cot.cuda()).sum(), xm)
    assert rel_err(out_m, out_r) <= 1e-5 and rel_err(gm, gr) <= 1e-5
    act = gmp_b200.tfn.ScalarActivation()
    xs = x[:, :C].contiguous()
    ref_a = e3nn_nn.Activation(o3.Irreps(f"{C}x0e"), [torch.nn.functional.silu])
    xr2, xm2 = xs.clone().requires_grad_(True), xs.cuda().requires_grad_(True)
    o_r, o_m = ref_a(xr2), act(xm2)
    (g2r,) = torch.autograd.grad(o_r.sum(), xr2)
    (g2m,) = torch.autograd.grad(o_m.sum(), xm2)
    assert rel_err(o_m, o_r) <= 1e-5 and rel_err(g2m, g2r) <= 1e-5
