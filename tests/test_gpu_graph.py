"""GPU: graph construction and CSR sort, bit-exact against the numpy oracle (integer path)."""
import numpy as np
import pytest
import torch

from oracle.thirdparty import cluster

pytestmark = pytest.mark.gpu


def _cloud(n, box, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 3, generator=g) * box


def _margin_ok(pos, batch, r, ulps=8):
    """No pair within a few ulp of r^2: then the edge set cannot depend on rounding conventions."""
    p = pos.numpy()
    ok = True
    for b in np.unique(batch):
        q = p[batch == b]
        d = cluster.sqdist_f32(q[None], q[:, None])
        ok &= not np.any(np.abs(d - np.float32(r * r)) <= ulps * np.spacing(np.float32(r * r)))
    return ok


@pytest.mark.parametrize("graphs,nodes,box,r,maxnb", [(8, 32, 8.0, 5.0, 32), (4, 64, 4.0, 2.0, 64), (1, 1, 1.0, 1.0, 4),
                                                      (3, 50, 2.0, 1.9, 8), (6, 20, 3.0, 1.0, 32)])
def test_radius_graph_bruteforce_bit_exact(graphs, nodes, box, r, maxnb):
    import gmp_b200
    pos = _cloud(graphs * nodes, box, 7)
    batch = torch.arange(graphs).repeat_interleave(nodes)
    assert _margin_ok(pos, batch.numpy(), r)
    ref = cluster.radius_graph(pos.numpy(), r, batch.numpy(), False, maxnb)
    got = gmp_b200.radius_graph(pos.cuda(), r, batch.cuda(), max_num_neighbors=maxnb, method="brute")
    assert got.dtype == torch.int64 and np.array_equal(got.cpu().numpy(), ref)
    ref_l = cluster.radius_graph(pos.numpy(), r, batch.numpy(), True, maxnb)
    got_l = gmp_b200.radius_graph(pos.cuda(), r, batch.cuda(), loop=True, max_num_neighbors=maxnb, method="brute")
    assert np.array_equal(got_l.cpu().numpy(), ref_l)


@pytest.mark.parametrize("n,box,r,maxnb", [(3000, 10.0, 1.0, 64), (2500, 6.0, 1.0, 8), (500, 1.5, 1.0, 128)])
def test_radius_graph_cells_equals_bruteforce(n, box, r, maxnb):
    import gmp_b200
    pos = _cloud(n, box, 11)
    ref = cluster.radius_graph(pos.numpy(), r, None, False, maxnb)
    got = gmp_b200.radius_graph(pos.cuda(), r, None, max_num_neighbors=maxnb, method="cells")
    brute = gmp_b200.radius_graph(pos.cuda(), r, None, max_num_neighbors=maxnb, method="brute")
    assert torch.equal(got, brute)
    assert np.array_equal(got.cpu().numpy(), ref)


def test_radius_graph_empty_and_isolated():
    import gmp_b200
    pos = torch.tensor([[0.0, 0, 0], [10.0, 0, 0], [20.0, 0, 0]]).cuda()
    ei = gmp_b200.radius_graph(pos, 1.0, None)
    assert ei.shape == (2, 0)


@pytest.mark.parametrize("E,n,seed", [(0, 5, 0), (1, 1, 1), (1000, 37, 2), (5000, 4000, 3), (4096, 3, 4)])
def test_csr_stable_sort_bit_exact(E, n, seed):
    import gmp_b200
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, n, (E,), generator=g)
    other = torch.randint(0, n, (E,), generator=g)
    rowptr, perm = cluster.csr_from_coo(idx.numpy(), n)
    csr = gmp_b200.build_csr(idx.cuda(), other.cuda(), n)
    assert np.array_equal(csr.rowptr.cpu().numpy(), rowptr)
    got_perm = np.arange(E, dtype=np.int32) if csr.perm is None else csr.perm.cpu().numpy()
    assert np.array_equal(got_perm, perm)
    assert np.array_equal(csr.col.cpu().numpy(), other.numpy()[perm].astype(np.int32))
    # size-independent property: sortedness + permutation
    assert np.all(np.diff(idx.numpy()[got_perm]) >= 0) and np.array_equal(np.sort(got_perm), np.arange(E))


@pytest.mark.parametrize("E,n,hot", [(1_000_000, 1, 1.0), (300_000, 7, 0.0), (1_200_000, 5000, 0.5), (70_000, 3, 0.9)])
def test_csr_stable_sort_long_rows(E, n, hot):
    """Rows far longer than a warp's rank sort can handle (pooling one large graph puts every node into one
    row): the block-wide radix sort of csrc/graph.cu.  `hot` = fraction of the edges that land in row 0; the rest is
    spread uniformly, so short and long rows mix.  Bit-exact against the numpy stable sort; must finish in seconds."""
    import time
    import gmp_b200
    g = torch.Generator().manual_seed(E % 977)
    idx = torch.randint(0, n, (E,), generator=g)
    idx[torch.rand(E, generator=g) < hot] = 0
    other = torch.randint(0, 1000, (E,), generator=g)
    rowptr, perm = cluster.csr_from_coo(idx.numpy(), n)
    ic, oc = idx.cuda(), other.cuda()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    csr = gmp_b200.build_csr(ic, oc, n)
    torch.cuda.synchronize()
    assert time.perf_counter() - t0 < 5.0
    assert np.array_equal(csr.rowptr.cpu().numpy(), rowptr)
    assert np.array_equal(csr.perm.cpu().numpy(), perm)
    assert np.array_equal(csr.col.cpu().numpy(), other.numpy()[perm].astype(np.int32))


def test_global_pool_single_large_graph():
    """global_add_pool / global_mean_pool over one graph of 2^20 nodes (config-5 shape through EGNNModel's read-out)."""
    import gmp_b200
    n = 1 << 20
    x = torch.randn(n, 16, generator=torch.Generator().manual_seed(3)).cuda()
    batch = torch.zeros(n, dtype=torch.long, device="cuda")
    s = gmp_b200.global_add_pool(x, batch, 1)
    m = gmp_b200.global_mean_pool(x, batch, 1)
    ref = x.double().sum(0, keepdim=True)
    assert (s.double() - ref).abs().max() <= 1e-5 * ref.abs().max().clamp_min(1.0) * 40
    assert (m.double() - ref / n).abs().max() <= 1e-6


def test_csr_sorted_input_is_identity():
    import gmp_b200
    pos = _cloud(256, 4.0, 5).cuda()
    ei = gmp_b200.radius_graph(pos, 1.2, None, max_num_neighbors=128)  # no truncation -> symmetric
    g = gmp_b200.get_graph(ei, 256)
    assert g.by_dst.perm is None  # radius_graph output is already dst-major
    s = g.by_src
    # symmetric graph: the src-sorted view has the same row structure
    assert torch.equal(s.rowptr, g.by_dst.rowptr) and torch.equal(s.col, g.by_dst.col)


@pytest.mark.parametrize("F,reduce", [(128, "sum"), (64, "mean"), (3, "mean"), (1, "sum"), (576, "sum")])
def test_scatter_matches_oracle(F, reduce):
    import gmp_b200
    from oracle.thirdparty.scatter import scatter as oscatter
    g = torch.Generator().manual_seed(F)
    E, n = 3000, 257
    idx = torch.randint(0, n - 5, (E,), generator=g)  # last rows empty -> zeros with dim_size
    src = torch.randn(E, F, generator=g)
    ref = oscatter(src.double(), idx, dim=0, dim_size=n, reduce=reduce)
    s = src.cuda().requires_grad_(True)
    out = gmp_b200.scatter(s, idx.cuda(), dim=0, dim_size=n, reduce=reduce)
    assert out.shape == (n, F)
    assert (out.detach().cpu().double() - ref).abs().max() <= 1e-5 * max(1.0, ref.abs().max().item())
    out2 = gmp_b200.scatter(s, idx.cuda(), dim=0, dim_size=n, reduce=reduce)
    assert torch.equal(out, out2)  # deterministic, bit for bit
    cot = torch.randn(n, F, generator=g)
    (gs,) = torch.autograd.grad((out * cot.cuda()).sum(), s)
    sr = src.double().requires_grad_(True)
    (gr,) = torch.autograd.grad((oscatter(sr, idx, dim=0, dim_size=n, reduce=reduce) * cot.double()).sum(), sr)
    assert (gs.cpu().double() - gr).abs().max() <= 1e-6 * max(1.0, gr.abs().max().item())
    # no dim_size: rows = index.max()+1 (torch_scatter semantics)
    assert gmp_b200.scatter(s, idx.cuda(), dim=0, reduce=reduce).shape[0] == int(idx.max()) + 1


@pytest.mark.parametrize("n,E,dup", [(50, 300, False), (2000, 100000, True), (5, 0, False), (1, 64, True), (300000, 1200000, False)])
def test_to_undirected_coalesce_on_gpu_bit_exact(n, E, dup):
    """to_undirected / coalesce of a CUDA edge list (two stable counting-sort passes + mark / scan / compact,
    csrc/graph.cu) against the oracle's torch_geometric restatement: same edges, same (row, col) order; Batch.from_data_list
    on device-resident graphs equals the host collate."""
    import gmp_b200
    from oracle.thirdparty import pyg
    g = torch.Generator().manual_seed(n + E)
    ei = torch.randint(0, n, (2, E), generator=g)
    if dup and E:
        ei = torch.cat([ei, ei[:, : E // 3], ei[:, : E // 5].flip(0)], dim=1)     # duplicates and pre-existing reverses
    want = pyg.to_undirected(ei)
    got = gmp_b200.to_undirected(ei.cuda(), n)
    assert got.is_cuda and torch.equal(got.cpu(), want)
    assert torch.equal(gmp_b200.coalesce(got, n), got)                              # idempotent
    # size-independent properties: strictly increasing keys, symmetric edge set
    if got.shape[1]:
        key = got[0] * n + got[1]
        assert bool((key[1:] > key[:-1]).all())
        assert torch.equal(gmp_b200.coalesce(got.flip(0), n), got)


def test_collate_on_device_matches_host():
    import gmp_b200
    g = torch.Generator().manual_seed(4)
    ds = [gmp_b200.Data(atoms=torch.randint(0, 5, (n,), generator=g), pos=torch.randn(n, 3, generator=g),
                        edge_index=gmp_b200.to_undirected(torch.randint(0, n, (2, 3 * n), generator=g)), y=torch.tensor(float(n)))
          for n in (6, 8, 5, 12)]
    host = gmp_b200.Batch.from_data_list(ds)
    dev = gmp_b200.Batch.from_data_list([d.to("cuda") for d in ds])
    for k in ("atoms", "pos", "edge_index", "y", "batch"):
        assert getattr(dev, k).is_cuda and torch.equal(getattr(dev, k).cpu(), getattr(host, k)), k
    assert dev.num_graphs == 4
    # the collated batch drives a model end to end
    torch.manual_seed(0)
    model = gmp_b200.EGNNModel(num_layers=2, emb_dim=64, in_dim=5, out_dim=2).cuda()
    assert model(dev).shape == (4, 2)
