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


def _emulate_tc_contract(tab, cg, rowidx, colidx, V, sh, T, b2, n, r_len):
    """numpy restatement of what csrc/tpconv_tc.cu computes from the host tables: N-tile columns -> W2 rows,
    factor per y-group, per-edge results summed into rows; plus the bias term through YS (gmp_tp_ysum + GEMM)."""
    res = np.zeros((n, r_len))
    E = len(rowidx)
    nt_iter = iter(tab["ntiles"])
    for (v_off, DA, DB, DS, sh_off, cg_off, A0, AR, r_off, MB, nslices, nsub, nt_begin, _, _, _) in tab["ygroups"]:
        Zc = cg[cg_off:cg_off + DA * DS * DB].reshape(DA, DS, DB)
        Z = np.einsum("ej,ijk->eik", sh[:, sh_off:sh_off + DS], Zc)
        WS = 32 if (DA >= DB and DB == 1) else 8
        for sl in range(nslices):
            for q in range(nsub):
                w_off, sa, sb, a_begin, a_end, b0, b_end, ws = next(nt_iter)
                assert ws == WS and b0 == sl * WS and a_begin == A0 + q * (256 // WS) and a_end == A0 + AR and b_end == MB
                for c in range(256):
                    a, b = a_begin + c // WS, b0 + c % WS
                    if a >= a_end or b >= b_end:
                        continue
                    y = np.einsum("ei,eik->ek", V[colidx, v_off + a * DA:v_off + (a + 1) * DA], Z)
                    np.add.at(res, (rowidx[:, None], (r_off + b * DB + np.arange(DB))[None, :]), T[:, w_off + a * sa + b * sb, None] * y)
    assert next(nt_iter, None) is None
    # bias: YS[n][path][a][k], then sum_a b2[a,b] YS[n,a,k]
    for (v_off, DA, DB, MA, y_off, z_off, _, _), bs in zip(tab["ypaths"], tab["bias"]):
        zs = tab["zent"][z_off:z_off + DA * DB]
        Z = np.stack([sum(sh[:, so + j] * cg[base + j * st] for j in range(ds)) for so, ds, base, st in zs], 1).reshape(E, DA, DB)
        Ye = np.einsum("eai,eik->eak", V[colidx, v_off:v_off + MA * DA].reshape(E, MA, DA), Z)
        YS = np.zeros((n, MA, DB))
        np.add.at(YS, rowidx, Ye)
        Bm = np.array([[b2[bs["w_off"] + a * bs["stride_a"] + b * bs["stride_b"]] for b in range(bs["MB"])] for a in range(MA)])
        res[:, bs["r_off"]:bs["r_off"] + bs["MB"] * DB] += np.einsum("nak,ab->nbk", YS, Bm).reshape(n, -1)
    return res


@pytest.mark.parametrize("cin,cout", [("8x0e+8x1o+8x2e", "24x0e+8x1o+8x2e"), ("40x0e+12x1o", "12x0e+40x1o+4x2e")])
def test_tc_tables_reproduce_the_tensor_product(cin, cout):
    """The tensor-core decomposition (y-groups, N-tiles, YS bias term) is exact algebra: forward and feature gradient."""
    import gmp_b200
    sh_ir = "1x0e+1x1o+1x2e"
    plan = gmp_b200.TensorProductPlan(cin, sh_ir, cout)
    tp = o3.FullyConnectedTensorProduct(cin, sh_ir, cout, shared_weights=False)
    g = torch.Generator().manual_seed(5)
    n, E = 7, 23
    src, dst = torch.randint(0, n, (E,), generator=g), torch.randint(0, n, (E,), generator=g)
    x = torch.randn(n, plan.irreps_in.dim, dtype=torch.float64, generator=g).requires_grad_(True)
    sh = torch.randn(E, 9, dtype=torch.float64, generator=g)
    T = torch.randn(E, plan.weight_numel, dtype=torch.float64, generator=g)
    b2 = torch.randn(plan.weight_numel, dtype=torch.float64, generator=g)
    tp = tp.double()
    out = torch.zeros(n, plan.irreps_out.dim, dtype=torch.float64).index_add_(0, src, tp(x[dst], sh, T + b2))
    cot = torch.randn(out.shape, dtype=torch.float64, generator=g)
    (dx,) = torch.autograd.grad((out * cot).sum(), [x])
    cg = plan.cg.astype(np.float64)
    mine = _emulate_tc_contract(plan.tc_fwd, cg, src.numpy(), dst.numpy(), x.detach().numpy(), sh.numpy(), T.numpy(), b2.numpy(),
                                n, plan.irreps_out.dim)
    assert np.abs(mine - out.detach().numpy()).max() <= 1e-5 * np.abs(out.detach().numpy()).max()
    mine_dx = _emulate_tc_contract(plan.tc_bwd, cg, dst.numpy(), src.numpy(), cot.numpy(), sh.numpy(), T.numpy(), b2.numpy(),
                                   n, plan.irreps_in.dim)
    assert np.abs(mine_dx - dx.numpy()).max() <= 1e-5 * np.abs(dx.numpy()).max()


def test_batch_bag_keeps_the_reference_field_names():
    """gmp_b200.Batch: the attribute bag the models read (`atoms`, `pos`, `edge_index`, `batch`)."""
    import gmp_b200
    b = gmp_b200.Batch(atoms=torch.zeros(4, dtype=torch.long), pos=torch.randn(4, 3),
                       edge_index=torch.tensor([[0, 1], [1, 0]]), batch=torch.zeros(4, dtype=torch.long), name="x")
    c = b.to("cpu")
    assert set(c.tensors()) == {"atoms", "pos", "edge_index", "batch"} and c.name == "x"
    assert torch.equal(c.pos, b.pos) and c.edge_index.dtype == torch.long


def test_data_batch_dataloader_match_the_pyg_restatement():
    """Data / Batch.from_data_list / DataLoader / to_undirected on host tensors (where the reference builds
    its datasets, experiments/utils/create_graphs.py:78-79) against oracle/thirdparty/pyg.py."""
    import gmp_b200
    from oracle.thirdparty import pyg
    g = torch.Generator().manual_seed(0)
    for n, E in ((50, 300), (7, 0), (3, 40)):
        ei = torch.randint(0, n, (2, E), generator=g)
        got = gmp_b200.to_undirected(ei)
        assert torch.equal(got, pyg.to_undirected(ei))
        if E:
            assert bool((got[0][1:] >= got[0][:-1]).all())           # row-ascending, duplicates dropped
            assert torch.equal(gmp_b200.coalesce(got), got)          # idempotent
    ds = [gmp_b200.Data(atoms=torch.randint(0, 5, (n,), generator=g), pos=torch.randn(n, 3, generator=g),
                        edge_index=torch.randint(0, n, (2, 2 * n), generator=g), y=torch.tensor(float(n)))
          for n in (5, 7, 3, 9, 4)]
    b = gmp_b200.Batch.from_data_list(ds)
    ob = pyg.Batch.from_data_list([pyg.Data(**d.__dict__) for d in ds])
    for k in ("atoms", "pos", "edge_index", "y", "batch"):
        assert torch.equal(getattr(b, k), getattr(ob, k)), k
    assert b.num_graphs == ob.num_graphs == 5 and b.num_nodes == 28
    loader = gmp_b200.DataLoader(ds, batch_size=2)
    assert len(loader) == 3 and [x.num_graphs for x in loader] == [2, 2, 1]
    assert sum(x.num_nodes for x in gmp_b200.DataLoader(ds, batch_size=2, shuffle=True)) == 28
    assert len(gmp_b200.DataLoader(ds, batch_size=2, drop_last=True)) == 2


def test_uvu_tables_match():
    """csrc/uvu_cg.cuh is generated (scripts/gen_uvu_cg.py) from the same wigner_3j the oracle pins; the committed header
    must be the current render, and the instruction list / mid irreps must equal the reference run's (golden `extra`)."""
    import importlib.util
    import os
    from tests.helpers import load_golden
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_uvu_cg", os.path.join(root, "scripts", "gen_uvu_cg.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert open(os.path.join(root, "geometric-message-passing_b200", "csrc", "uvu_cg.cuh")).read() == mod.render()
    from gmp_b200.mace_blocks import linear_out_irreps, tp_out_irreps_with_instructions
    fx = load_golden("mace_interaction_RealAgnosticInteractionBlock")
    mid, ins = tp_out_irreps_with_instructions(fx["ctor"]["node_feats_irreps"], fx["ctor"]["edge_attrs_irreps"], fx["ctor"]["target_irreps"])
    assert str(mid) == fx["extra"]["irreps_mid"]
    assert [tuple(t[:3]) for t in ins] == [tuple(t) for t in fx["extra"]["instructions"]]
    assert str(linear_out_irreps(mid.simplify(), fx["ctor"]["target_irreps"]).simplify()) == fx["extra"]["irreps_out"]


def test_interaction_block_state_dict_keys_match_the_reference():
    import gmp_b200
    from tests.helpers import load_golden
    for cls in ("ResidualElementDependentInteractionBlock", "AgnosticNonlinearInteractionBlock", "AgnosticResidualNonlinearInteractionBlock",
                "RealAgnosticInteractionBlock", "RealAgnosticResidualInteractionBlock"):
        fx = load_golden("mace_interaction_" + cls)
        m = getattr(gmp_b200, cls)(**fx["ctor"])
        ref = {k: tuple(v.shape) for k, v in fx["state"].items() if "output_mask" not in k and not k.startswith("conv_tp.")}
        own = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert own == ref, (cls, set(own) ^ set(ref))


@pytest.mark.parametrize("world,n,r", [(2, 1000, 1.0), (3, 500, 2.5), (5, 300, 6.0), (4, 4, 1.0)])
def test_slab_partition_peer_tables_are_mutually_consistent(world, n, r):
    """The peer-memory halo exchange reads, on rank p, rows out of rank q's buffers at offsets p computes by itself
    (gmp_b200.distributed.slab_partition: peer_send_start, peer_halo_offset).  They must equal what q does with its own
    tables: the start of q's send range towards p, and where p's rows sit inside q's [left halo | right halo] rows --
    also when slabs are thinner than the radius (halo rows from several ranks, world = 5 here)."""
    import gmp_b200
    g = torch.Generator().manual_seed(world * 1000 + n)
    xs = torch.sort(torch.rand(n, generator=g) * 20.0).values
    parts = [gmp_b200.slab_partition(xs, r, p, world) for p in range(world)]
    assert len({pt.n_own_max for pt in parts}) == 1 and len({pt.halo_max for pt in parts}) == 1
    assert parts[0].n_own_max == max(pt.n_own for pt in parts)
    assert parts[0].halo_max == max(max(pt.n_left + pt.n_right for pt in parts), 1)
    for p in range(world):
        for q in range(world):
            if p == q:
                continue
            a, b = parts[q].send_ranges[p]
            assert parts[p].recv_counts[q] == b - a
            if b > a:
                assert parts[p].peer_send_start[q] == a
            # rows owned by p inside q's halo rows: q's halo is [from rank 0 | from rank 1 | ...] without q itself
            off = sum(parts[q].recv_counts[p2] for p2 in range(p) if p2 != q)
            assert parts[p].peer_halo_offset[q] == off


def test_gvp_state_dict_keys_match_the_reference():
    import gmp_b200
    from tests.helpers import load_golden
    fx = load_golden("gvp_model")
    own = {k: tuple(v.shape) for k, v in gmp_b200.GVPGNNModel(**fx["ctor"]).state_dict().items()}
    ref = {k: tuple(v.shape) for k, v in fx["state"].items()}
    assert own == ref, set(own) ^ set(ref)
