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
