"""CPU, world_size 2 over gloo: the host-side logic of the graph-sharded multi-GPU path.
* e3nn BatchNorm with a process group reproduces the single-process statistics and gradients;
* the flat gradient all-reduce equals the single-process gradient of the concatenated batch."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gmp_b200
    torch.manual_seed(0)
    ir = "4x0e+4x1o+4x2e"
    x_full = torch.randn(24, 36, dtype=torch.float64)
    cot = torch.randn(24, 36, dtype=torch.float64)
    w, b = torch.randn(12, dtype=torch.float64), torch.randn(4, dtype=torch.float64)
    sl = slice(rank * 12, (rank + 1) * 12)
    bn = gmp_b200.BatchNorm(ir, process_group=dist.group.WORLD).double()
    with torch.no_grad():
        bn.weight.copy_(w), bn.bias.copy_(b)
    x = x_full[sl].clone().requires_grad_(True)
    y = bn(x)
    (y * cot[sl]).sum().backward()
    flat = torch.cat([p.grad.reshape(-1) for p in bn.parameters()])
    dist.all_reduce(flat)  # what bench.py does with the model gradients after backward
    q.put((rank,) + tuple(v.detach().numpy().copy() for v in (y, x.grad, flat, bn.running_var, bn.running_mean)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_batchnorm_and_gradient_allreduce_match_single_process():
    import gmp_b200
    port = 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=100) for _ in range(2)], key=lambda t: t[0])
    res = [(r[0],) + tuple(torch.from_numpy(v) for v in r[1:]) for r in res]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    torch.manual_seed(0)
    ir = "4x0e+4x1o+4x2e"
    x_full = torch.randn(24, 36, dtype=torch.float64)
    cot = torch.randn(24, 36, dtype=torch.float64)
    w, b = torch.randn(12, dtype=torch.float64), torch.randn(4, dtype=torch.float64)
    bn = gmp_b200.BatchNorm(ir).double()
    with torch.no_grad():
        bn.weight.copy_(w), bn.bias.copy_(b)
    x = x_full.clone().requires_grad_(True)
    y = bn(x)
    (y * cot).sum().backward()
    flat = torch.cat([p.grad.reshape(-1) for p in bn.parameters()])
    y2 = torch.cat([res[0][1], res[1][1]])
    gx2 = torch.cat([res[0][2], res[1][2]])
    assert (y2 - y.detach()).abs().max() < 1e-12
    assert (gx2 - x.grad).abs().max() < 1e-12
    for r in res:
        assert (r[3] - flat).abs().max() < 1e-12
        assert (r[4] - bn.running_var).abs().max() < 1e-12 and (r[5] - bn.running_mean).abs().max() < 1e-12


# ------------------------------------------------------------------------------------------------
# destination-partitioned single graph: slab partition + halo exchange (gloo send/recv), CPU
# ------------------------------------------------------------------------------------------------
def _toy_layer(x_loc, pos_loc, ei, W, n_rows):
    """messages W x_j * |pos_i - pos_j| summed at i (plain torch; stands in for a fused layer on the CPU)."""
    d = (pos_loc[ei[1]] - pos_loc[ei[0]]).norm(dim=-1, keepdim=True)
    m = (x_loc[ei[0]] @ W) * d
    return torch.zeros(n_rows, W.shape[1], dtype=x_loc.dtype).index_add_(0, ei[1], m)


def _halo_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np
    import gmp_b200
    from oracle.thirdparty import cluster
    g = torch.Generator().manual_seed(0)
    pos = torch.rand(300, 3, generator=g, dtype=torch.float64) * torch.tensor([6.0, 2.0, 2.0], dtype=torch.float64)
    pos = pos[torch.argsort(pos[:, 0])]
    x = torch.randn(300, 5, generator=g, dtype=torch.float64)
    W = torch.randn(5, 4, generator=g, dtype=torch.float64).requires_grad_(True)
    r = 0.9
    part = gmp_b200.slab_partition(pos[:, 0], r, rank, world)
    pos_loc = pos[part.local_global]
    ei = torch.from_numpy(cluster.radius_graph(pos_loc.float().numpy(), r, None, False, 128))
    ei = ei[:, (ei[1] >= part.n_left) & (ei[1] < part.n_left + part.n_own)]
    x_own = x[part.own_lo:part.own_hi].clone().requires_grad_(True)
    x_loc = gmp_b200.halo_exchange(x_own, part)
    assert torch.equal(x_loc.detach(), x[part.local_global])          # halo rows are the owners' rows
    out = _toy_layer(x_loc, pos_loc, ei, W, part.n_local)[part.own_slice]
    cot = torch.randn(300, 4, generator=g, dtype=torch.float64)[part.own_lo:part.own_hi]
    (out * cot).sum().backward()
    gmp_b200.allreduce_gradients([W])
    q.put((rank, out.detach().numpy().copy(), x_own.grad.numpy().copy(), W.grad.numpy().copy(), ei.shape[1]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("world", [2, 3])
def test_slab_partition_halo_exchange_matches_single_process(world):
    from oracle.thirdparty import cluster
    port = 31500 + (os.getpid() + world) % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_halo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=100) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    pos = torch.rand(300, 3, generator=g, dtype=torch.float64) * torch.tensor([6.0, 2.0, 2.0], dtype=torch.float64)
    pos = pos[torch.argsort(pos[:, 0])]
    x = torch.randn(300, 5, generator=g, dtype=torch.float64).requires_grad_(True)
    W = torch.randn(5, 4, generator=g, dtype=torch.float64).requires_grad_(True)
    ei = torch.from_numpy(cluster.radius_graph(pos.float().numpy(), 0.9, None, False, 128))
    out = _toy_layer(x, pos, ei, W, 300)
    cot = torch.randn(300, 4, generator=g, dtype=torch.float64)
    (out * cot).sum().backward()
    assert sum(r[4] for r in res) == ei.shape[1]                        # every edge owned by exactly one rank
    assert (torch.cat([torch.from_numpy(r[1]) for r in res]) - out.detach()).abs().max() < 1e-12
    assert (torch.cat([torch.from_numpy(r[2]) for r in res]) - x.grad).abs().max() < 1e-12
    for r in res:
        assert (torch.from_numpy(r[3]) - W.grad).abs().max() < 1e-11
