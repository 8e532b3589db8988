"""Oracle (oracle/ref_layers.py) vs the golden vectors produced by running the
UNMODIFIED reference (tests/golden/make_golden.py).  CPU only.

Tolerance: both sides run the same ATen fp32 ops in (almost) the same order, so
1e-6 normwise relative is expected; 2e-6 is asserted.
"""
import pytest
import torch

from oracle import ref_layers as R
from oracle.thirdparty import o3
from tests.helpers import Bag, check_against_digest, densify, load_golden, load_params, rel_err

TOL = 2e-6


def _run(module, args, wrt, cots):
    for t in wrt.values():
        t.requires_grad_(True)
    out = module(*args)
    outs = list(out) if isinstance(out, (tuple, list)) else [out]
    loss = sum((o * c).sum() for o, c in zip(outs, cots))
    params = dict(module.named_parameters())
    grads = torch.autograd.grad(loss, list(wrt.values()) + list(params.values()), allow_unused=True)
    names = [f"input.{k}" for k in wrt] + [f"param.{k}" for k in params]
    return outs, dict(zip(names, grads))


def _check(fx, outs, grads, tol=TOL):
    for o, ref in zip(outs, fx["outputs"]):
        assert rel_err(o, ref) <= tol
    for name, ref in fx["grads"].items():
        if ref is None:
            assert grads.get(name) is None or float(grads[name].abs().max()) == 0.0, name
            continue
        check_against_digest(grads[name], ref, 10 * tol, name)


@pytest.mark.parametrize("name", ["egnn_layer_relu_add", "egnn_layer_swish_mean"])
def test_egnn_layer(name):
    fx = load_golden(name)
    m = load_params(R.EGNNLayer(**fx["ctor"]), fx["state"])
    i = fx["inputs"]
    h, pos = i["h"].clone(), i["pos"].clone()
    outs, grads = _run(m, (h, pos, i["edge_index"]), {"h": h, "pos": pos}, fx["cotangent"])
    _check(fx, outs, grads)


def test_mpnn_layer():
    fx = load_golden("mpnn_layer")
    m = load_params(R.MPNNLayer(**fx["ctor"]), fx["state"])
    h = fx["inputs"]["h"].clone()
    outs, grads = _run(m, (h, fx["inputs"]["edge_index"]), {"h": h}, fx["cotangent"])
    _check(fx, outs, grads)


def test_egnn_model_config1_kchains():
    fx = load_golden("egnn_model_kchains")
    m = load_params(R.EGNNModel(**fx["ctor"]), fx["state"])
    i = fx["inputs"]
    # the fixture itself is the restated k-chains generator, replicated 32x (BASELINE config 1)
    pos = i["pos"].clone()
    b = Bag(atoms=i["atoms"], pos=pos, edge_index=i["edge_index"], batch=i["batch"])
    assert pos.shape[0] == 384 and i["edge_index"].shape[1] == 640
    outs, grads = _run(m, (b,), {"pos": pos}, fx["cotangent"])
    _check(fx, outs, grads, tol=1e-5)


def test_schnet_model():
    fx = load_golden("schnet_model")
    m = load_params(R.SchNetModel(**fx["ctor"]), fx["state"])
    i = fx["inputs"]
    pos = i["pos"].clone()
    b = Bag(atoms=i["atoms"], pos=pos, edge_index=i["edge_index"], batch=i["batch"])
    outs, grads = _run(m, (b,), {"pos": pos}, fx["cotangent"])
    _check(fx, outs, grads)


def test_schnet_interaction():
    from oracle.thirdparty.pyg import InteractionBlock
    fx = load_golden("schnet_interaction")
    m = load_params(InteractionBlock(**fx["ctor"]), fx["state"])
    i = fx["inputs"]
    x, ew, ea = i["x"].clone(), i["edge_weight"].clone(), i["edge_attr"].clone()
    outs, grads = _run(m, (x, i["edge_index"], ew, ea), {"x": x, "edge_weight": ew, "edge_attr": ea}, fx["cotangent"])
    _check(fx, outs, grads)


def test_edge_geometry():
    fx = load_golden("edge_geometry")
    c = fx["ctor"]
    vec = fx["inputs"]["vec"]
    rad = R.RadialEmbeddingBlock(c["r_max"], c["num_bessel"], c["num_polynomial_cutoff"])
    sh = o3.SphericalHarmonics(o3.Irreps.spherical_harmonics(c["max_ell"]), True, "component")
    assert rel_err(rad(vec.norm(dim=-1, keepdim=True)), fx["outputs"]["rbf"]) <= TOL
    assert rel_err(sh(vec), fx["outputs"]["sh"]) <= TOL
    assert float(fx["outputs"]["rbf"][2].abs().max()) == 0.0  # beyond r_max


@pytest.mark.parametrize("name", ["tfn_conv_first", "tfn_conv_hidden", "mace_conv_hidden", "tfn_conv_mean_nogate"])
def test_tp_conv_layer(name):
    fx = load_golden(name)
    m = load_params(R.TensorProductConvLayer(**fx["ctor"]), fx["state"])
    m.train()
    assert str(m.out_irreps) == fx["extra"]["tp_out_irreps"]
    assert m.tp.weight_numel == fx["extra"]["weight_numel"]
    i = fx["inputs"]
    x, sh, ft = i["node_attr"].clone(), i["edge_sh"].clone(), i["edge_feat"].clone()
    outs, grads = _run(m, (x, i["edge_index"], sh, ft), {"node_attr": x, "edge_sh": sh, "edge_feat": ft},
                       fx["cotangent"])
    _check(fx, outs, grads)
    if fx["extra"]["bn_running_var"] is not None:
        assert rel_err(m.batch_norm.running_var, fx["extra"]["bn_running_var"]) <= TOL
        assert rel_err(m.batch_norm.running_mean, fx["extra"]["bn_running_mean"]) <= 1e-5


def test_mace_product_block():
    fx = load_golden("mace_product_block")
    m = load_params(R.EquivariantProductBasisBlock(**fx["ctor"]), fx["state"])
    i = fx["inputs"]
    x, sc = i["node_feats"].clone(), i["sc"].clone()
    outs, grads = _run(m, (x, sc, None), {"node_feats": x, "sc": sc}, fx["cotangent"])
    _check(fx, outs, grads)


def test_u_matrices():
    fx = load_golden("u_matrices")["outputs"]
    ks = {"0e": (1, 3, 11), "1o": (1, 4, 21), "2e": (1, 4, 23)}
    for ir in ("0e", "1o", "2e"):
        for nu in (1, 2, 3):
            ref = densify(fx[f"{ir}.{nu}"])
            mine = R.u_matrix_real("1x0e+1x1o+1x2e", ir, nu, dtype=torch.float32)
            assert mine.shape == ref.shape and mine.shape[-1] == ks[ir][nu - 1]
            assert (mine - ref).abs().max().item() <= 1e-6


def test_irreps_tools():
    fx = load_golden("irreps_tools")
    t = fx["inputs"]["t"]
    assert torch.equal(R.reshape_irreps("8x0e+8x1o+8x2e")(t), fx["outputs"]["reshaped"])
    assert [str(s) for s in R.irreps2gate("8x0e+8x1o+8x2e")] == fx["outputs"]["gate_split"]
    b = fx["inputs"]["batch"]
    assert torch.equal(R.first_node_pooling(torch.arange(9.0).unsqueeze(1), b), fx["outputs"]["first_pool"])


@pytest.mark.parametrize("name,cls", [("tfn_model", "TFNModel"), ("mace_model", "MACEModel")])
def test_equivariant_models(name, cls):
    fx = load_golden(name)
    m = load_params(getattr(R, cls)(**fx["ctor"]), fx["state"])
    m.train()
    i = fx["inputs"]
    b = Bag(atoms=i["atoms"], pos=i["pos"].clone(), edge_index=i["edge_index"], batch=i["batch"])
    outs, grads = _run(m, (b,), {}, fx["cotangent"])
    _check(fx, outs, grads, tol=1e-5)


_KINDS = {"ResidualElementDependentInteractionBlock": "residual_element", "AgnosticNonlinearInteractionBlock": "agnostic_nonlinear",
          "AgnosticResidualNonlinearInteractionBlock": "agnostic_residual_nonlinear", "RealAgnosticInteractionBlock": "real_agnostic",
          "RealAgnosticResidualInteractionBlock": "real_agnostic_residual"}


@pytest.mark.parametrize("cls", sorted(_KINDS))
def test_mace_interaction_block(cls):
    """the oracle's restatement of models/mace_modules/blocks.py:206-530 against the reference run."""
    fx = load_golden("mace_interaction_" + cls)
    m = load_params(R.InteractionBlock(_KINDS[cls], **fx["ctor"]), fx["state"])
    assert str(m.conv_tp.irreps_out) == fx["extra"]["irreps_mid"] and m.conv_tp.weight_numel == fx["extra"]["weight_numel"]
    assert [(i.i_in1, i.i_in2, i.i_out) for i in m.conv_tp.instructions] == [tuple(t) for t in fx["extra"]["instructions"]]
    i = fx["inputs"]
    x, ef = i["node_feats"].clone(), i["edge_feats"].clone()

    def fwd(*a):
        out = m(*a)
        return tuple(o for o in out if o is not None) if isinstance(out, tuple) else out

    class W(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.m = m

        def forward(self, *a):
            return fwd(*a)

    outs, grads = _run(W(), (i["node_attrs"], x, i["edge_attrs"], ef, i["edge_index"]), {"node_feats": x, "edge_feats": ef}, fx["cotangent"])
    _check(fx, outs, {k.replace("param.m.", "param."): v for k, v in grads.items()})


def test_gvp_conv_layer():
    """oracle GVPConvLayer against the reference run (models/layers/gvp_layer.py:327-438, eval mode)."""
    import torch.nn.functional as F
    fx = load_golden("gvp_conv_layer")
    c = fx["ctor"]
    m = R.GVPConvLayer(c["node_dims"], c["edge_dims"], drop_rate=c["drop_rate"], activations=(F.relu, None), vector_gate=c["vector_gate"],
                       residual=c["residual"])
    m = load_params(m, fx["state"]).eval()
    i = fx["inputs"]
    s, v, es, ev = (i[k].clone() for k in ("s", "v", "edge_s", "edge_v"))

    class W(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.m = m

        def forward(self, s_, v_, es_, ev_, ei):
            return self.m((s_, v_), ei, (es_, ev_))

    outs, grads = _run(W(), (s, v, es, ev, i["edge_index"]), {"s": s, "v": v, "edge_s": es, "edge_v": ev}, fx["cotangent"])
    _check(fx, outs, {k.replace("param.m.", "param."): g for k, g in grads.items()})


def test_gvp_model():
    fx = load_golden("gvp_model")
    m = load_params(R.GVPGNNModel(**fx["ctor"]), fx["state"]).eval()
    i = fx["inputs"]
    b = Bag(atoms=i["atoms"], pos=i["pos"].clone(), edge_index=i["edge_index"], batch=i["batch"])
    outs, grads = _run(m, (b,), {}, fx["cotangent"])
    _check(fx, outs, grads)
