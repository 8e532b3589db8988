"""GPU parity: fused SchNet CFConv / InteractionBlock / SchNetModel (through the C ABI) against the
golden vectors of the unmodified reference and against the CPU oracle on seeded random inputs.

Tolerance (fp32 strict mode): 1e-5 normwise relative, the north star's bound."""
import pytest
import torch

from tests.helpers import Bag, check_against_digest, load_golden, load_params, random_clouds, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _grads(module, outs, cots, wrt):
    loss = sum((o * c).sum() for o, c in zip(outs, cots))
    params = dict(module.named_parameters())
    gs = torch.autograd.grad(loss, list(wrt.values()) + list(params.values()), allow_unused=True)
    names = [f"input.{k}" for k in wrt] + [f"param.{k}" for k in params]
    return dict(zip(names, gs))


def _check_fixture(fx, outs, grads, tol=TOL):
    for o, ref in zip(outs, fx["outputs"]):
        assert rel_err(o, ref) <= tol
    for name, ref in fx["grads"].items():
        if ref is None:
            continue
        got = grads[name]
        assert got is not None, name
        check_against_digest(got.cpu(), ref, 10 * tol, name)


def test_interaction_block_golden_materialised_edge_attr():
    import gmp_b200
    fx = load_golden("schnet_interaction")
    m = load_params(gmp_b200.InteractionBlock(**fx["ctor"]), fx["state"]).cuda()
    i = fx["inputs"]
    x, ew, ea = (i[k].cuda().requires_grad_(True) for k in ("x", "edge_weight", "edge_attr"))
    out = m(x, i["edge_index"].cuda(), ew, ea)
    grads = _grads(m, [out], [c.cuda() for c in fx["cotangent"]], {"x": x, "edge_weight": ew, "edge_attr": ea})
    _check_fixture(fx, [out], grads)


def test_schnet_model_golden_fused_distance_expansion():
    import gmp_b200
    fx = load_golden("schnet_model")
    m = load_params(gmp_b200.SchNetModel(**fx["ctor"]), fx["state"]).cuda()
    i = fx["inputs"]
    pos = i["pos"].cuda().requires_grad_(True)
    b = Bag(atoms=i["atoms"].cuda(), pos=pos, edge_index=i["edge_index"].cuda(), batch=i["batch"].cuda())
    out = m(b)
    grads = _grads(m, [out], [c.cuda() for c in fx["cotangent"]], {"pos": pos})
    _check_fixture(fx, [out], grads)


@pytest.mark.parametrize("F,graphs,nodes,shuffle", [(128, 24, 32, False), (64, 7, 19, True), (128, 3, 40, True)])
def test_interaction_block_vs_oracle_random(F, graphs, nodes, shuffle):
    """Seeded random molecules; optionally a shuffled (unsorted) edge_index and isolated atoms."""
    import gmp_b200
    from oracle.thirdparty.pyg import GaussianSmearing, InteractionBlock
    d = random_clouds(graphs, nodes, 8.0, 5.0, 40 + F, max_nb=32)
    ei, pos = d["edge_index"], d["pos"]
    if shuffle:
        g = torch.Generator().manual_seed(1)
        ei = ei[:, torch.randperm(ei.shape[1], generator=g)]
        keep = (ei[0] >= 3) & (ei[1] >= 3)  # atoms 0..2 become isolated: empty CSR rows
        ei = ei[:, keep]
    torch.manual_seed(F)
    ref = InteractionBlock(F, 50, F, 5.0)
    with torch.no_grad():
        for p in ref.parameters():
            if p.dim() == 1:
                p.normal_(0, 0.3)
    sm = GaussianSmearing(0.0, 5.0, 50)
    x = torch.randn(pos.shape[0], F, generator=torch.Generator().manual_seed(2))
    ew = (pos[ei[0]] - pos[ei[1]]).norm(dim=-1)
    cot = torch.randn(pos.shape[0], F, generator=torch.Generator().manual_seed(3))

    xr, ewr = x.clone().requires_grad_(True), ew.clone().requires_grad_(True)
    out_r = ref(xr, ei, ewr, sm(ewr))
    gr = torch.autograd.grad((out_r * cot).sum(), [xr, ewr] + list(ref.parameters()))

    mine = gmp_b200.InteractionBlock(F, 50, F, 5.0)
    mine.load_state_dict(ref.state_dict())
    mine = mine.cuda()
    smc = gmp_b200.GaussianSmearing(0.0, 5.0, 50).cuda()
    xc, ewc = x.cuda().requires_grad_(True), ew.cuda().requires_grad_(True)
    out = mine(xc, ei.cuda(), ewc, smc.lazy())
    gm = torch.autograd.grad((out * cot.cuda()).sum(), [xc, ewc] + list(mine.parameters()))
    assert rel_err(out, out_r) <= TOL
    for a, b_, name in zip(gm, gr, ["x", "edge_weight"] + [n for n, _ in ref.named_parameters()]):
        assert rel_err(a, b_) <= 5 * TOL, name
    # bitwise run-to-run determinism (atomics-free reduction)
    out2 = mine(xc, ei.cuda(), ewc, smc.lazy())
    gm2 = torch.autograd.grad((out2 * cot.cuda()).sum(), [xc, ewc] + list(mine.parameters()))
    assert torch.equal(out, out2) and all(torch.equal(a, b_) for a, b_ in zip(gm, gm2))


def test_cfconv_no_edges():
    import gmp_b200
    m = gmp_b200.InteractionBlock(64, 50, 64, 5.0).cuda()
    x = torch.randn(10, 64, device="cuda", requires_grad=True)
    ei = torch.zeros(2, 0, dtype=torch.long, device="cuda")
    out = m(x, ei, torch.zeros(0, device="cuda"), gmp_b200.GaussianSmearing(0.0, 5.0, 50).cuda().lazy())
    # agg = 0 -> lin(ssp(lin2(0)))
    ref = m.lin(m.act(m.conv.lin2(torch.zeros(10, 64, device="cuda"))))
    assert torch.allclose(out, ref)
    out.sum().backward()
    assert x.grad is not None and float(x.grad.abs().max()) == 0.0


def test_k0_gather_mul_segsum():
    import gmp_b200
    from gmp_b200._lib import call, ptr
    d = random_clouds(5, 30, 6.0, 3.0, 77)
    ei = d["edge_index"].cuda()
    n, E, F = 150, ei.shape[1], 128
    g = gmp_b200.get_graph(ei, n).by_dst
    x = torch.randn(n, F, device="cuda")
    w = torch.randn(E, F, device="cuda")
    out = torch.empty(n, F, device="cuda")
    call("gmp_gather_mul_segsum_f32", ptr(g.rowptr), ptr(g.col), g.perm_ptr, ptr(x), ptr(w), ptr(out), n, F)
    ref = torch.zeros(n, F, dtype=torch.float64).index_add_(0, ei[1].cpu(), (x[ei[0]] * w).double().cpu())
    assert rel_err(out, ref) <= 1e-6


@pytest.mark.gpu
def test_device_prefetcher_matches_direct_upload():
    """gmp_b200.DevicePrefetcher (next batch uploaded on a side stream under the current batch's compute): same
    batches, same order, same model outputs as a synchronous `.to(device)` per step."""
    import gmp_b200
    from tests.helpers import random_clouds
    torch.manual_seed(0)
    model = gmp_b200.SchNetModel(hidden_channels=128, num_filters=128, num_layers=2, num_gaussians=50, cutoff=5.0).cuda()
    hosts = []
    for seed in range(5):
        d = random_clouds(6, 20, 8.0, 5.0, 100 + seed, max_nb=32)
        atoms = torch.randint(1, 10, (d["pos"].shape[0],), generator=torch.Generator().manual_seed(seed))
        hosts.append(gmp_b200.Batch(atoms=atoms, pos=d["pos"], batch=d["batch"], edge_index=d["edge_index"]).pin_memory())
    with torch.no_grad():
        direct = [model(h.to("cuda")).cpu() for h in hosts]
        fetched = [model(b).cpu() for b in gmp_b200.DevicePrefetcher(hosts, "cuda")]
    assert len(fetched) == len(direct)
    for a, b in zip(fetched, direct):
        assert torch.equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graphed_step_matches_eager(precision):
    """gmp_b200.GraphedStep: forward + backward captured into one CUDA graph.  A replay reproduces the eager outputs and
    gradients bit for bit (same kernels, same order); in rebuild mode a host batch of the same shapes but different
    contents (atoms, positions, shuffled edge order) is loaded into the static buffers and the replay -- which then
    contains the CSR sort -- equals an eager step on that batch."""
    import gmp_b200
    torch.manual_seed(0)
    model = gmp_b200.SchNetModel(hidden_channels=128, num_filters=128, num_layers=2, num_gaussians=50, cutoff=5.0,
                                 precision=precision).cuda()
    params = [p for p in model.parameters()]
    d = random_clouds(6, 24, 8.0, 5.0, 11, max_nb=32)
    n = d["pos"].shape[0]
    gen = torch.Generator().manual_seed(3)

    def eager(b):
        for p in params:
            p.grad = None
        out = model(b)
        out.sum().backward()
        return out.detach().clone(), [p.grad.detach().clone() for p in params]

    b1 = gmp_b200.Batch(atoms=torch.randint(1, 10, (n,), generator=gen), pos=d["pos"], batch=d["batch"],
                        edge_index=d["edge_index"], num_graphs=6)
    perm = torch.randperm(d["edge_index"].shape[1], generator=gen)
    b2 = gmp_b200.Batch(atoms=torch.randint(1, 10, (n,), generator=gen), pos=d["pos"] + 0.05 * torch.randn(n, 3, generator=gen),
                        batch=d["batch"], edge_index=d["edge_index"][:, perm].contiguous(), num_graphs=6)
    out1, g1 = eager(b1.to("cuda"))
    out2, g2 = eager(b2.to("cuda"))

    gs = gmp_b200.GraphedStep(model, b1.to("cuda"), warmup=2)                       # resident batch
    assert gs.kernels_per_replay > 0
    kn = gs.kernel_nodes()     # counted from the captured cudaGraph_t itself: ours + ATen + cuBLAS
    assert kn is None or kn >= gs.kernels_per_replay
    for _ in range(2):
        assert torch.equal(gs.replay(), out1)
        assert all(torch.equal(a, b) for a, b in zip(gs.grads, g1))

    gs = gmp_b200.GraphedStep(model, b1.to("cuda"), warmup=2, rebuild_graph=True)    # static buffers, CSR sort captured
    for host, out_ref, g_ref in ((b2, out2, g2), (b1, out1, g1), (b2, out2, g2)):
        gs.load(host.pin_memory())
        out = gs.replay()
        assert torch.equal(out, out_ref)
        assert all(torch.equal(a, b) for a, b in zip(gs.grads, g_ref))
    # the same through the double-buffered uploader (two persistent staging sets, upload of batch i+1 under step i)
    seq = [(b1, out1, g1), (b2, out2, g2), (b2, out2, g2), (b1, out1, g1), (b2, out2, g2)]
    hosts = [h.pin_memory() for h, _, _ in seq]
    for staged, (_, out_ref, g_ref) in zip(gmp_b200.DevicePrefetcher(hosts, "cuda", static=True), seq):
        gs.load(staged)
        out = gs.replay()
        assert torch.equal(out, out_ref)
        assert all(torch.equal(a, b) for a, b in zip(gs.grads, g_ref))


def test_graphed_step_load_guards_the_edge_list():
    """a step captured for one fixed edge list (rebuild_graph=False) must not silently accept another one; with
    rebuild_graph=True a non-dst-sorted edge list loaded into a buffer that came from radius_graph must give the same
    result as the eager model (the captured CSR build may not trust the capture-time sortedness tag)."""
    import gmp_b200
    d = random_clouds(6, 20, 8.0, 5.0, 321, max_nb=32, atoms_hi=10)
    n = d["pos"].shape[0]
    torch.manual_seed(0)
    model = gmp_b200.SchNetModel(hidden_channels=128, num_filters=128, num_layers=2, num_gaussians=50, cutoff=5.0).cuda()
    ei = gmp_b200.radius_graph(d["pos"].cuda(), 5.0, d["batch"].cuda(), max_num_neighbors=32)
    perm = torch.randperm(ei.shape[1], generator=torch.Generator().manual_seed(1)).cuda()
    shuffled = ei[:, perm].contiguous()
    mk = lambda e: gmp_b200.Batch(atoms=d["atoms"].cuda(), pos=d["pos"].cuda(), batch=d["batch"].cuda(), edge_index=e, num_graphs=6)
    fixed = gmp_b200.GraphedStep(model, mk(ei.clone()), warmup=1)
    with pytest.raises(ValueError):
        fixed.load(mk(shuffled))
    fixed.load(mk(ei.clone()))                       # same edges: accepted, node fields refreshed
    out_fixed = fixed.replay().clone()
    del fixed
    static = mk(ei)                                  # the radius_graph output itself (tagged dst-sorted) becomes the static buffer
    gs = gmp_b200.GraphedStep(model, static, warmup=1, rebuild_graph=True)
    gs.load(mk(shuffled))
    out = gs.replay().clone()
    grads = [g.clone() for g in gs.grads]
    del gs
    for p in model.parameters():
        p.grad = None
    ref = model(mk(shuffled))
    ref.sum().backward()
    assert rel_err(out, ref) <= 1e-5 and rel_err(out_fixed, ref) <= 1e-5
    for g, p in zip(grads, model.parameters()):
        assert rel_err(g, p.grad) <= 5e-5
