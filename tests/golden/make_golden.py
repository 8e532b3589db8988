"""Generate the golden vectors in tests/golden/*.pt by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

How: ``oracle/shims`` provides stand-ins for the wheels the reference imports
but this image cannot install (e3nn, torch_scatter, torch_geometric,
opt_einsum); they re-export the pure-torch restatements in
``oracle/thirdparty``.  With the shims and ``/root/reference`` on ``sys.path``
the reference's own modules (models/layers/*.py, models/*.py,
models/mace_modules/*.py) import and run unchanged.  What these vectors pin is
therefore the reference-owned code; the third-party semantics stay
"upstream-recalled" (oracle/__init__.py).

Each fixture is a dict: ``ctor`` kwargs, ``state`` (state_dict), ``inputs``,
``outputs``, ``cotangent`` and ``grads`` (d<out,cot>/d inputs and parameters).
"""
from __future__ import annotations

import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
sys.path.insert(0, REF)

# `models/__init__.py` imports DimeNet/SphereNet/GVP (out of scope, need sympy/torch_sparse):
# register `models` as a bare namespace so that only the hot-path submodules load.
for name, sub in (("models", "models"), ("models.layers", "models/layers")):
    pkg = types.ModuleType(name)
    pkg.__path__ = [os.path.join(REF, sub)]
    sys.modules[name] = pkg

from models.egnn import EGNNModel  # noqa: E402
from models.layers.egnn_layer import EGNNLayer, MPNNLayer  # noqa: E402
from models.layers.tfn_layer import TensorProductConvLayer  # noqa: E402
from models.mace import MACEModel  # noqa: E402
from models.mace_modules import blocks as ref_blocks  # noqa: E402
from models.mace_modules.blocks import EquivariantProductBasisBlock, RadialEmbeddingBlock  # noqa: E402
from models.mace_modules.cg import U_matrix_real  # noqa: E402
from models.mace_modules.irreps_tools import irreps2gate, reshape_irreps  # noqa: E402
from models.mace_modules.symmetric_contraction import SymmetricContraction  # noqa: E402
from models.schnet import SchNetModel  # noqa: E402
from models.tfn import TFNModel, first_node_pooling  # noqa: E402

import e3nn  # noqa: E402  (the shim)
from oracle.ref_layers import create_kchains  # noqa: E402  (fixture restated from experiments/kchains.ipynb:71-107)
from oracle.thirdparty import cluster  # noqa: E402
from oracle.thirdparty.pyg import Batch, Data  # noqa: E402


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def random_clouds(num_graphs, nodes, box, r, seed, max_nb=64, atoms_hi=1):
    g = _gen(seed)
    pos = torch.rand(num_graphs * nodes, 3, generator=g) * box
    batch = torch.arange(num_graphs).repeat_interleave(nodes)
    ei = torch.from_numpy(cluster.radius_graph(pos.numpy(), r, batch.numpy(), False, max_nb))
    atoms = torch.randint(1, atoms_hi, (num_graphs * nodes,), generator=g) if atoms_hi > 1 else torch.zeros(
        num_graphs * nodes, dtype=torch.long)
    return Data(atoms=atoms, pos=pos, edge_index=ei, batch=batch)


def run(module, args, wrt, seed, out_index=None):
    """Forward + backward of <out, cotangent>; returns outputs, cotangents, grads."""
    # snapshot parameters/buffers BEFORE the forward (batch norm updates its running stats in it)
    module._state0 = {k: v.detach().clone() for k, v in module.state_dict().items()}
    for t in wrt.values():
        t.requires_grad_(True)
    out = module(*args)
    outs = list(out) if isinstance(out, (tuple, list)) else [out]
    g = _gen(seed)
    cots = [torch.randn(o.shape, generator=g) for o in outs]
    loss = sum((o * c).sum() for o, c in zip(outs, cots))
    params = dict(module.named_parameters())
    grads = torch.autograd.grad(loss, list(wrt.values()) + list(params.values()), allow_unused=True)
    names = [f"input.{k}" for k in wrt] + [f"param.{k}" for k in params]
    return ([o.detach().clone() for o in outs], cots,
            {n: (None if gr is None else digest(gr.detach())) for n, gr in zip(names, grads)})


def digest(t: torch.Tensor, keep: int = 4096):
    """Small tensors are stored whole; big parameter gradients as a digest
    (sum, L2 norm, a seeded +-1 projection and the first 64 entries) so fixtures stay small."""
    if t.numel() <= keep:
        return t.clone()
    flat = t.reshape(-1).double()
    sign = torch.randint(0, 2, (flat.numel(),), generator=_gen(flat.numel())).double() * 2 - 1
    return {"digest": True, "shape": tuple(t.shape), "sum": flat.sum().item(), "norm": flat.norm().item(),
            "proj": (flat * sign).sum().item(), "head": t.reshape(-1)[:64].clone()}


def strip_buffers(state):
    """Drop deterministic constructor-time buffers (U matrices, Bessel constants)."""
    return {k: v for k, v in state.items()
            if "U_matrix" not in k and "bessel" not in k and "cutoff_fn" not in k and not k.endswith("r_max")}


def sparse(t: torch.Tensor):
    nz = t.nonzero()
    return {"shape": tuple(t.shape), "idx": nz.to(torch.int16), "val": t[tuple(nz.T)].clone()}


def save(name, **payload):
    path = os.path.join(HERE, name + ".pt")
    torch.save(payload, path)
    print(f"{name:28s} {os.path.getsize(path) / 1024:8.1f} KiB")


def main():
    torch.manual_seed(0)

    # ---- config 1: EGNN 5x128 on k-chains (k=4), batch 64 graphs ---------------------------------
    graphs = create_kchains(4) * 32
    b = Batch.from_data_list(graphs)
    torch.manual_seed(0)
    m = EGNNModel(num_layers=5, emb_dim=128, in_dim=1, out_dim=2)
    pos = b.pos.clone()
    b.pos = pos
    outs, cots, grads = run(m, (b,), {"pos": pos}, 1)
    save("egnn_model_kchains", ctor=dict(num_layers=5, emb_dim=128, in_dim=1, out_dim=2), state=m._state0,
         inputs=dict(atoms=b.atoms, pos=pos.detach(), edge_index=b.edge_index, batch=b.batch),
         outputs=outs, cotangent=cots, grads=grads)

    # ---- EGNN layer on random clouds (asymmetric degrees, isolated tail nodes kept out) ----------
    d = random_clouds(6, 10, 3.0, 1.6, 2)
    for act, norm, aggr in (("relu", "layer", "add"), ("swish", "layer", "mean")):
        torch.manual_seed(3)
        layer = EGNNLayer(64, activation=act, norm=norm, aggr=aggr)
        h = torch.randn(d.pos.shape[0], 64, generator=_gen(4))
        pos = d.pos.clone()
        # make sure the highest-numbered node receives an edge (reference omits dim_size)
        assert int(d.edge_index[1].max()) == d.pos.shape[0] - 1
        outs, cots, grads = run(layer, (h, pos, d.edge_index), {"h": h, "pos": pos}, 5)
        save(f"egnn_layer_{act}_{aggr}", ctor=dict(emb_dim=64, activation=act, norm=norm, aggr=aggr),
             state=layer._state0, inputs=dict(h=h.detach(), pos=pos.detach(), edge_index=d.edge_index),
             outputs=outs, cotangent=cots, grads=grads)

    torch.manual_seed(6)
    layer = MPNNLayer(32)
    h = torch.randn(d.pos.shape[0], 32, generator=_gen(7))
    outs, cots, grads = run(layer, (h, d.edge_index), {"h": h}, 8)
    save("mpnn_layer", ctor=dict(emb_dim=32), state=layer._state0,
         inputs=dict(h=h.detach(), edge_index=d.edge_index), outputs=outs, cotangent=cots, grads=grads)

    # ---- SchNet ----------------------------------------------------------------------------------
    d = random_clouds(5, 12, 6.0, 5.0, 9, max_nb=32, atoms_hi=10)
    torch.manual_seed(10)
    m = SchNetModel(hidden_channels=64, num_filters=64, num_layers=3, num_gaussians=50, cutoff=5.0, out_dim=2)
    pos = d.pos.clone()
    d.pos = pos
    outs, cots, grads = run(m, (d,), {"pos": pos}, 11)
    save("schnet_model", ctor=dict(hidden_channels=64, num_filters=64, num_layers=3, num_gaussians=50, cutoff=5.0,
                                   out_dim=2), state=m._state0,
         inputs=dict(atoms=d.atoms, pos=pos.detach(), edge_index=d.edge_index, batch=d.batch),
         outputs=outs, cotangent=cots, grads=grads)

    torch.manual_seed(12)
    m = SchNetModel(hidden_channels=128, num_filters=128, num_layers=1, num_gaussians=50, cutoff=5.0)
    blk = m.interactions[0]
    x = torch.randn(d.pos.shape[0], 128, generator=_gen(13))
    row, col = d.edge_index
    ew = (d.pos.detach()[row] - d.pos.detach()[col]).norm(dim=-1)
    ea = m.distance_expansion(ew).detach()
    ew = ew.clone()
    outs, cots, grads = run(blk, (x, d.edge_index, ew, ea), {"x": x, "edge_weight": ew, "edge_attr": ea}, 14)
    save("schnet_interaction", ctor=dict(hidden_channels=128, num_gaussians=50, num_filters=128, cutoff=5.0),
         state=blk._state0, inputs=dict(x=x.detach(), edge_index=d.edge_index, edge_weight=ew.detach(),
                                             edge_attr=ea.detach()),
         outputs=outs, cotangent=cots, grads=grads)

    # ---- edge geometry: radial embedding + spherical harmonics -------------------------------------
    vec = torch.randn(64, 3, generator=_gen(15)) * 0.8
    vec[0] = torch.tensor([0.0, 0.0, 1.0])
    vec[1] = torch.tensor([1.0, 0.0, 0.0])
    vec[2] = torch.tensor([0.0, -2.5, 0.0])  # beyond r_max = 2: cutoff -> 0
    length = vec.norm(dim=-1, keepdim=True)
    rad = RadialEmbeddingBlock(r_max=2.0, num_bessel=8, num_polynomial_cutoff=5)
    sh = e3nn.o3.SphericalHarmonics(e3nn.o3.Irreps.spherical_harmonics(2), normalize=True, normalization="component")
    save("edge_geometry", inputs=dict(vec=vec), outputs=dict(rbf=rad(length), sh=sh(vec)),
         ctor=dict(r_max=2.0, num_bessel=8, num_polynomial_cutoff=5, max_ell=2))

    # ---- TFN conv layer (gate) and MACE conv layer (batch norm), C = 8 ---------------------------
    d = random_clouds(4, 12, 3.0, 2.0, 16)
    assert int(d.edge_index[0].max()) == d.pos.shape[0] - 1
    sh_ir = e3nn.o3.Irreps.spherical_harmonics(2)
    hid = e3nn.o3.Irreps("8x0e+8x1o+8x2e")
    vec = d.pos[d.edge_index[0]] - d.pos[d.edge_index[1]]
    esh = sh(vec)
    rbf = rad(vec.norm(dim=-1, keepdim=True))
    for tag, in_ir, kw in (("tfn_conv_first", e3nn.o3.Irreps("8x0e"), dict(gate=True)),
                           ("tfn_conv_hidden", hid, dict(gate=True)),
                           ("mace_conv_hidden", hid, dict(gate=False, batch_norm=True)),
                           ("tfn_conv_mean_nogate", hid, dict(gate=False, aggr="mean"))):
        torch.manual_seed(17)
        layer = TensorProductConvLayer(in_irreps=in_ir, out_irreps=hid, sh_irreps=sh_ir, edge_feats_dim=8,
                                       mlp_dim=64, aggr=kw.get("aggr", "add"), batch_norm=kw.get("batch_norm", False),
                                       gate=kw.get("gate", False))
        x = torch.randn(d.pos.shape[0], in_ir.dim, generator=_gen(18))
        e_sh, e_ft = esh.clone(), rbf.clone()
        outs, cots, grads = run(layer, (x, d.edge_index, e_sh, e_ft),
                                {"node_attr": x, "edge_sh": e_sh, "edge_feat": e_ft}, 19)
        save(tag, ctor=dict(in_irreps=str(in_ir), out_irreps=str(hid), sh_irreps=str(sh_ir), edge_feats_dim=8,
                            mlp_dim=64, **kw),
             state=layer._state0, inputs=dict(node_attr=x.detach(), edge_index=d.edge_index,
                                                   edge_sh=e_sh.detach(), edge_feat=e_ft.detach()),
             outputs=outs, cotangent=cots, grads=grads,
             extra=dict(tp_out_irreps=str(layer.out_irreps), weight_numel=layer.tp.weight_numel,
                        bn_running_mean=None if layer.batch_norm is None else layer.batch_norm.running_mean.clone(),
                        bn_running_var=None if layer.batch_norm is None else layer.batch_norm.running_var.clone()))

    # ---- MACE product basis block, symmetric contraction, U matrices ------------------------------
    torch.manual_seed(20)
    blk = EquivariantProductBasisBlock(node_feats_irreps=hid, target_irreps=hid, correlation=3,
                                       element_dependent=False, num_elements=1, use_sc=True)
    xf = torch.randn(24, 8, 9, generator=_gen(21))
    sc = torch.randn(24, hid.dim, generator=_gen(22))
    outs, cots, grads = run(blk, (xf, sc, None), {"node_feats": xf, "sc": sc}, 23)
    save("mace_product_block", ctor=dict(node_feats_irreps=str(hid), target_irreps=str(hid), correlation=3,
                                         element_dependent=False, num_elements=1, use_sc=True),
         state=strip_buffers(blk._state0), inputs=dict(node_feats=xf.detach(), sc=sc.detach()),
         outputs=outs, cotangent=cots, grads=grads)

    u = {}
    for ir in ("0e", "1o", "2e"):
        for nu in (1, 2, 3):
            u[f"{ir}.{nu}"] = sparse(U_matrix_real(irreps_in=e3nn.o3.Irreps("1x0e+1x1o+1x2e"),
                                                   irreps_out=e3nn.o3.Irreps(ir), correlation=nu,
                                                   dtype=torch.float32)[-1])
    save("u_matrices", outputs=u)

    rs = reshape_irreps(hid)
    t = torch.randn(5, hid.dim, generator=_gen(24))
    sg = irreps2gate(e3nn.o3.Irreps("8x0e+8x1o+8x2e"))
    bvec = torch.tensor([0, 0, 0, 1, 1, 2, 2, 2, 2])
    save("irreps_tools", inputs=dict(t=t, batch=bvec), outputs=dict(
        reshaped=rs(t), gate_split=[str(s) for s in sg],
        first_pool=first_node_pooling(torch.arange(9.0).unsqueeze(1), bvec)))

    # ---- whole models, small --------------------------------------------------------------------
    d = random_clouds(4, 10, 3.0, 2.0, 25)
    for tag, cls, ctor in (("tfn_model", TFNModel, dict(r_max=2.0, max_ell=2, num_layers=2, emb_dim=4, mlp_dim=64, out_dim=2)),
                           ("mace_model", MACEModel, dict(r_max=2.0, max_ell=2, correlation=3, num_layers=2,
                                                          emb_dim=4, mlp_dim=64, out_dim=2))):
        torch.manual_seed(26)
        m = cls(**ctor)
        m.train()
        pos = d.pos.clone()
        d.pos = pos
        outs, cots, grads = run(m, (d,), {}, 27)
        save(tag, ctor=ctor, state=strip_buffers(m._state0),
             inputs=dict(atoms=d.atoms, pos=pos.detach(), edge_index=d.edge_index, batch=d.batch),
             outputs=outs, cotangent=cots, grads=grads)

    interaction_blocks()
    gvp_fixtures()


def interaction_blocks():
    """the five ACEsuit-style interaction blocks (models/mace_modules/blocks.py:206-530), C = 8 channels,
    3 elements, on a 2-cloud radius graph."""
    I = e3nn.o3.Irreps
    C = 8
    feats, sh_ir = I(f"{C}x0e+{C}x1o+{C}x2e"), I("1x0e+1x1o+1x2e")
    d = random_clouds(2, 10, 3.0, 2.0, 31)
    N, E = d.pos.shape[0], d.edge_index.shape[1]
    vec = d.pos[d.edge_index[0]] - d.pos[d.edge_index[1]]
    sh = e3nn.o3.SphericalHarmonics(sh_ir, normalize=True, normalization="component")
    rad = RadialEmbeddingBlock(r_max=2.0, num_bessel=8, num_polynomial_cutoff=5)
    for cls in ("ResidualElementDependentInteractionBlock", "AgnosticNonlinearInteractionBlock", "AgnosticResidualNonlinearInteractionBlock",
                "RealAgnosticInteractionBlock", "RealAgnosticResidualInteractionBlock"):
        torch.manual_seed(32)
        ctor = dict(node_attrs_irreps="3x0e", node_feats_irreps=str(feats), edge_attrs_irreps=str(sh_ir), edge_feats_irreps="8x0e",
                    target_irreps=str(feats), hidden_irreps=str(feats), avg_num_neighbors=5.5)
        blk = getattr(ref_blocks, cls)(**{k: (I(v) if k.endswith("irreps") else v) for k, v in ctor.items()})
        g = _gen(33)
        attrs = torch.eye(3)[torch.randint(0, 3, (N,), generator=g)]
        x = torch.randn(N, feats.dim, generator=g)
        e_sh, e_ft = sh(vec).clone(), rad(vec.norm(dim=-1, keepdim=True)).clone()
        outs, cots, grads = run(_NoNone(blk), (attrs, x, e_sh, e_ft, d.edge_index), {"node_feats": x, "edge_feats": e_ft}, 34)
        grads = {k.replace("param.m.", "param."): v for k, v in grads.items()}
        save("mace_interaction_" + cls, cls=cls, ctor=ctor, state={k[2:]: v for k, v in blk._wrap_state0.items()},
             inputs=dict(node_attrs=attrs, node_feats=x.detach(), edge_attrs=e_sh.detach(), edge_feats=e_ft.detach(), edge_index=d.edge_index),
             outputs=outs, cotangent=cots, grads=grads,
             extra=dict(irreps_mid=str(blk.conv_tp.irreps_out), weight_numel=blk.conv_tp.weight_numel, irreps_out=str(blk.irreps_out),
                        instructions=[(i.i_in1, i.i_in2, i.i_out) for i in blk.conv_tp.instructions]))


def gvp_fixtures():
    """GVPConvLayer (models/layers/gvp_layer.py:327-438) and GVPGNNModel (models/gvpgnn.py), eval mode (dropout off)."""
    import models.layers.gvp_layer as ref_gvp
    from models.gvpgnn import GVPGNNModel
    d = random_clouds(3, 12, 3.0, 2.0, 41)
    N, E = d.pos.shape[0], d.edge_index.shape[1]
    g = _gen(42)
    torch.manual_seed(43)
    ctor = dict(node_dims=(16, 4), edge_dims=(8, 1), drop_rate=0.1, vector_gate=True, residual=True)
    layer = ref_gvp.GVPConvLayer(activations=(torch.nn.functional.relu, None), **ctor)
    layer.eval()
    s, v = torch.randn(N, 16, generator=g), torch.randn(N, 4, 3, generator=g)
    es, ev = torch.randn(E, 8, generator=g), torch.randn(E, 1, 3, generator=g)

    class Wrap(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, s_, v_, es_, ev_, ei):
            return self.m((s_, v_), ei, (es_, ev_))

    w = Wrap(layer)
    w.eval()
    outs, cots, grads = run(w, (s, v, es, ev, d.edge_index), {"s": s, "v": v, "edge_s": es, "edge_v": ev}, 44)
    save("gvp_conv_layer", ctor=ctor, state={k[2:]: t for k, t in w._state0.items()},
         inputs=dict(s=s.detach(), v=v.detach(), edge_s=es.detach(), edge_v=ev.detach(), edge_index=d.edge_index),
         outputs=outs, cotangent=cots, grads={k.replace("param.m.", "param."): t for k, t in grads.items()})

    torch.manual_seed(45)
    mctor = dict(r_max=2.0, num_layers=2, s_dim=16, v_dim=4, s_dim_edge=8, v_dim_edge=1, out_dim=2)
    m = GVPGNNModel(**mctor)
    m.eval()
    pos = d.pos.clone()
    d.pos = pos
    outs, cots, grads = run(m, (d,), {}, 46)
    save("gvp_model", ctor=mctor, state=strip_buffers(m._state0),
         inputs=dict(atoms=d.atoms, pos=pos.detach(), edge_index=d.edge_index, batch=d.batch), outputs=outs, cotangent=cots, grads=grads)


class _NoNone(torch.nn.Module):
    """Drops the ``None`` some blocks return as their second output (run() takes tensors only)."""

    def __init__(self, m):
        super().__init__()
        self.m = m

    def forward(self, *a):
        out = self.m(*a)
        self.m._wrap_state0 = self._state0
        return tuple(o for o in out if o is not None) if isinstance(out, tuple) else out


if __name__ == "__main__":
    main()
