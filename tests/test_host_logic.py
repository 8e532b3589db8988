"""CPU: product-side host logic (irreps parsing, CG tables, tensor-product plans, Gate / BatchNorm /
radial / SH modules, which are plain torch ops) against the oracle.  No CUDA kernels are launched."""
import numpy as np
import pytest
import torch

from oracle import ref_layers as R
from oracle.thirdparty import e3nn_nn, o3
from tests.helpers import load_golden


def test_wigner_3j_matches_oracle():
    from gmp_b200.irreps import wigner_3j
    for l1 in range(3):
        for l2 in range(4):
            for l3 in range(abs(l1 - l2), min(l1 + l2, 2) + 1):
                ref = o3.wigner_3j(l1, l2, l3, dtype=torch.float64).numpy()
                assert np.abs(wigner_3j(l1, l2, l3) - ref).max() < 1e-12


@pytest.mark.parametrize("cin,cout,gate", [("64x0e+64x1o+64x2e", "64x0e+64x1o+64x2e", True), ("64x0e", "64x0e+64x1o+64x2e", True),
                                           ("128x0e+128x1o+128x2e", "128x0e+128x1o+128x2e", False), ("8x0e+8x1o+8x2e", "8x0e+8x1o+8x2e", True)])
def test_tensor_product_plan_matches_e3nn_instructions(cin, cout, gate):
    import gmp_b200
    sh = "1x0e+1x1o+1x2e"
    mine = gmp_b200.TensorProductConvLayer(cin, cout, sh, 8, 64, gate=gate)
    ref = R.TensorProductConvLayer(cin, cout, sh, 8, 64, gate=gate)
    assert str(mine.out_irreps) == str(ref.out_irreps) and mine.tp.weight_numel == ref.tp.weight_numel
    assert set(mine.state_dict()) == set(ref.state_dict())
    off = 0
    for p, ins in zip(mine.tp.plan.paths, ref.tp.instructions):
        assert (p.i_in, p.i_sh, p.i_out) == (ins.i_in1, ins.i_in2, ins.i_out) and abs(p.coeff - ins.path_weight) < 1e-12
        assert p.w_off == off
        off += p.mul_in * p.mul_out
    plan = mine.tp.plan
    for tab in (plan.fwd, plan.bwd):
        for r_off, MB, DB, WS, pb, pe, ub, _ in tab["blocks"]:
            assert WS * DB <= 80 and WS >= 1
            for ps in tab["passes"][pb:pe]:
                assert ps[4] * WS <= 128  # MC * WS columns fit the 128-wide slice
    # every W2 row belongs to exactly one wgrad unit
    cover = np.zeros(plan.weight_numel, dtype=np.int32)
    for u in plan.wunits:
        cover[u[0]:u[0] + u[1]] += 1
    assert (cover == 1).all()


def test_config_sizes():
    import gmp_b200
    sh = "1x0e+1x1o+1x2e"
    assert gmp_b200.TensorProductPlan("64x0e+64x1o+64x2e", sh, "192x0e+64x1o+64x2e").weight_numel == 69632
    assert gmp_b200.TensorProductPlan("128x0e+128x1o+128x2e", sh, "128x0e+128x1o+128x2e").weight_numel == 180224


def test_gate_batchnorm_radial_sh_modules_match_oracle():
    import gmp_b200
    from gmp_b200.irreps import NORM2MOM, gate_split
    assert abs(NORM2MOM["silu"] - e3nn_nn.normalize2mom(torch.nn.functional.silu).cst) < 1e-12
    assert abs(NORM2MOM["sigmoid"] - e3nn_nn.normalize2mom(torch.sigmoid).cst) < 1e-12
    s, g, gd = gate_split("8x0e+8x1o+8x2e")
    assert (str(s), str(g), str(gd)) == ("8x0e", "16x0e", "8x1o+8x2e")
    ref_gate = e3nn_nn.Gate(s.__str__(), [torch.nn.functional.silu], str(g), [torch.sigmoid], str(gd))
    mine_gate = gmp_b200.Gate(s, g, gd)
    x = torch.randn(7, ref_gate.irreps_in.dim, dtype=torch.float64)
    assert (mine_gate(x) - ref_gate(x)).abs().max() < 1e-12
    ir = "8x0e+8x1o+8x2e"
    rb, mb = e3nn_nn.BatchNorm(ir).double(), gmp_b200.BatchNorm(ir).double()
    with torch.no_grad():
        w, b = torch.randn(24, dtype=torch.float64), torch.randn(8, dtype=torch.float64)
        rb.weight.copy_(w), mb.weight.copy_(w), rb.bias.copy_(b), mb.bias.copy_(b)
    x = torch.randn(33, 72, dtype=torch.float64)
    for mode in ("train", "eval"):
        getattr(rb, mode)(), getattr(mb, mode)()
        assert (rb(x) - mb(x)).abs().max() < 1e-12
        assert (rb.running_var - mb.running_var).abs().max() < 1e-12 and (rb.running_mean - mb.running_mean).abs().max() < 1e-12
    fx = load_golden("edge_geometry")
    vec = fx["inputs"]["vec"]
    rad = gmp_b200.RadialEmbeddingBlock(2.0, 8, 5)
    assert (rad(vec.norm(dim=-1, keepdim=True)) - fx["outputs"]["rbf"]).abs().max() < 2e-6
    assert (gmp_b200.SphericalHarmonics(2)(vec) - fx["outputs"]["sh"]).abs().max() < 2e-6
