"""Node-side dense chains on tcgen05 (csrc/node_chain.cu, through the C ABI) against plain PyTorch fp32 of the same ops.
Tolerance: the bf16-operand mode's 1e-2 normwise (BASELINE.json north_star: "1e-2 where bf16 MLP inputs are used")."""
import math

import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-2


def _ssp(x):
    return torch.nn.functional.softplus(x) - math.log(2.0)


@pytest.mark.parametrize("n", [1, 300, 128 * 5 + 17, 40000])
@pytest.mark.parametrize("act", [None, "relu", "silu", "ssp"])
def test_single_stage(n, act):
    from gmp_b200 import nodechain as nc
    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, 128, generator=g).cuda()
    w = (torch.randn(128, 128, generator=g) / 128 ** 0.5).cuda()
    b = torch.randn(128, generator=g).cuda()
    res = torch.randn(n, 128, generator=g).cuda()
    out, o16 = nc.linear(x, w, b, act=act, res=res, want_bf16=True)
    ref = torch.nn.functional.linear(x, w, b)
    ref = {None: lambda v: v, "relu": torch.relu, "silu": torch.nn.functional.silu, "ssp": _ssp}[act](ref) + res
    assert rel_err(out, ref) <= TOL
    assert rel_err(o16.float(), ref) <= 2 * TOL
    out2, _ = nc.linear(x, w, b, act=act, res=res)
    assert torch.equal(out, out2)   # run-to-run bitwise


@pytest.mark.parametrize("n", [7, 1000])
@pytest.mark.parametrize("act", ["relu", "silu"])
def test_egnn_update_chain(n, act):
    """mlp_upd of models/layers/egnn_layer.py:41-48 on cat[h, agg] (two K = 128 sources), residual of models/egnn.py:76."""
    from gmp_b200 import nodechain as nc
    g = torch.Generator().manual_seed(3)
    h, agg = torch.randn(n, 128, generator=g).cuda(), torch.randn(n, 128, generator=g).cuda() * 3
    w0 = (torch.randn(128, 256, generator=g) / 16).cuda()
    w1 = (torch.randn(128, 128, generator=g) / 11).cuda()
    b0, b1, g0, g1, be0, be1 = (torch.randn(128, generator=g).cuda() for _ in range(6))
    pre0, pre1 = torch.empty(n, 128, device="cuda"), torch.empty(n, 128, device="cuda")
    out = torch.empty(n, 128, device="cuda")
    nc.run(h, [nc.stage(nc.pack_w(w0), b0, ln=(g0, be0, 1e-5), act=act, out_pre=pre0),
               nc.stage(nc.pack_w(w1), b1, ln=(g1, be1, 1e-5), act=act, add_res=h, out_f32=out, out_pre=pre1)], a1=agg)
    f = torch.relu if act == "relu" else torch.nn.functional.silu
    r0 = torch.nn.functional.linear(torch.cat([h, agg], 1), w0, b0)
    a0 = f(torch.nn.functional.layer_norm(r0, (128,), g0, be0, 1e-5))
    r1 = torch.nn.functional.linear(a0, w1, b1)
    ref = f(torch.nn.functional.layer_norm(r1, (128,), g1, be1, 1e-5)) + h
    assert rel_err(pre0, r0) <= TOL
    assert rel_err(pre1, r1) <= 2 * TOL
    assert rel_err(out, ref) <= 3 * TOL    # two LayerNorms amplify the bf16 operand rounding of the first stage


@pytest.mark.parametrize("n", [129, 5000])
def test_schnet_chains(n):
    """The forward chain agg -> lin2 -> ssp -> lin -> + h -> next lin1 and the transposed chain of its backward pass."""
    from gmp_b200 import nodechain as nc
    g = torch.Generator().manual_seed(5)
    agg, h = torch.randn(n, 128, generator=g).cuda(), torch.randn(n, 128, generator=g).cuda()
    w2, wl, w1n = ((torch.randn(128, 128, generator=g) / 11).cuda() for _ in range(3))
    b2, bl = torch.randn(128, generator=g).cuda(), torch.randn(128, generator=g).cuda()
    y, hn = torch.empty(n, 128, device="cuda"), torch.empty(n, 128, device="cuda")
    x1n = torch.empty(n, 128, device="cuda", dtype=torch.bfloat16)
    nc.run(agg, [nc.stage(nc.pack_w(w2), b2, act="ssp", out_f32=y), nc.stage(nc.pack_w(wl), bl, add_res=h, out_f32=hn),
                 nc.stage(nc.pack_w(w1n), out_bf16=x1n)])
    y_r = _ssp(torch.nn.functional.linear(agg, w2, b2))
    hn_r = torch.nn.functional.linear(y_r, wl, bl) + h
    x1n_r = torch.nn.functional.linear(hn_r, w1n)
    assert rel_err(y, y_r) <= TOL and rel_err(hn, hn_r) <= TOL and rel_err(x1n.float(), x1n_r) <= 2 * TOL
    # backward: dx1' -> lin1^T -> + G -> [G'] -> lin^T -> * ssp'(Y) -> [dT] -> lin2^T -> [dAgg]
    dx1n, G = torch.randn(n, 128, generator=g).cuda(), torch.randn(n, 128, generator=g).cuda()
    gt, dT, dagg = (torch.empty(n, 128, device="cuda") for _ in range(3))
    dagg16 = torch.empty(n, 128, device="cuda", dtype=torch.bfloat16)
    nc.run(dx1n, [nc.stage(nc.pack_w(w1n, True), add_res=G, out_f32=gt),
                  nc.stage(nc.pack_w(wl, True), mul_aux=y_r, mul_mode=nc.MUL_DSSP, out_f32=dT),
                  nc.stage(nc.pack_w(w2, True), out_f32=dagg, out_bf16=dagg16)])
    gt_r = dx1n @ w1n + G
    dT_r = (gt_r @ wl) * torch.sigmoid(torch.nn.functional.linear(agg, w2, b2))
    dagg_r = dT_r @ w2
    assert rel_err(gt, gt_r) <= TOL and rel_err(dT, dT_r) <= TOL and rel_err(dagg, dagg_r) <= 2 * TOL
    assert rel_err(dagg16.float(), dagg_r) <= 2 * TOL


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 4097, 131072])
def test_wgrad_bulk_pipeline(n):
    """dW = g^T x, db = column sums of g (nn.Linear parameter gradients, 128 x 128) on the bulk-copy pipelined reduction kernel:
    ragged row counts around the 32-row tile, fewer tiles than CTAs, the config-2 node count."""
    from gmp_b200 import nodechain as nc
    gen = torch.Generator().manual_seed(n)
    g, x = torch.randn(n, 128, generator=gen).cuda(), torch.randn(n, 128, generator=gen).cuda()
    dw, db = nc.wgrad(g, x)
    ref_w, ref_b = g.double().t() @ x.double(), g.double().sum(0)
    assert rel_err(dw, ref_w) <= TOL and rel_err(db, ref_b) <= 1e-5
    dw2, db2 = nc.wgrad(g, x)
    assert torch.equal(dw, dw2) and torch.equal(db, db2)
