"""GPU, >= 2 devices (run with `gpurun --gpus 2`): the destination-partitioned EGNN with NCCL halo exchange equals
the single-GPU layer stack on the same graph (forward and gradients), SURVEY.md §8e row 2."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def _data():
    g = torch.Generator().manual_seed(0)
    pos = torch.rand(6000, 3, generator=g) * torch.tensor([24.0, 5.0, 5.0])
    pos = pos[torch.argsort(pos[:, 0])].contiguous()
    h = torch.randn(6000, 128, generator=g)
    cot_h, cot_p = torch.randn(6000, 128, generator=g), torch.randn(6000, 3, generator=g)
    return pos, h, cot_h, cot_p


def _worker(rank, world, port, q, halo="nccl", precision="fp32", act="relu"):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import gmp_b200
    dev = torch.device("cuda", rank)
    pos, h, cot_h, cot_p = (t.to(dev) for t in _data())
    part = gmp_b200.slab_partition(pos[:, 0], 1.0, rank, world)
    ei = gmp_b200.distributed.local_radius_graph(pos[part.local_global], 1.0, part)
    torch.manual_seed(1)
    gmp_b200.set_fast_matmul(precision == "bf16")
    model = gmp_b200.PartitionedEGNN(num_layers=2, emb_dim=128, halo=halo, precision=precision, activation=act).to(dev)
    h_own = h[part.own_lo:part.own_hi].clone().requires_grad_(True)
    p_own = pos[part.own_lo:part.own_hi].clone().requires_grad_(True)
    ho, po = model(h_own, p_own, ei, part)
    ((ho * cot_h[part.own_lo:part.own_hi]).sum() + (po * cot_p[part.own_lo:part.own_hi]).sum()).backward()
    params = list(model.parameters())
    gmp_b200.allreduce_gradients(params)
    torch.cuda.synchronize()
    q.put((rank, ho.detach().cpu().numpy(), po.detach().cpu().numpy(), h_own.grad.cpu().numpy(), p_own.grad.cpu().numpy(),
           [p.grad.cpu().numpy() for p in params], int(ei.shape[1])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("halo,precision", [("nccl", "fp32"), ("peer", "fp32"), ("peer", "bf16"), ("fused", "bf16")])
def test_partitioned_egnn_matches_single_gpu(halo, precision):
    """halo = "peer": the halo rows are pulled out of the owners' symmetric-memory buffers by csrc/halo.cu (P2P loads over
    NVLink) instead of ncclSend / ncclRecv; same numbers either way (the local edge order equals the global one).
    halo = "fused" (bf16 mode): the feature rows are not exchanged at all -- the edge kernels' gather reads the halo sources'
    projected rows out of the neighbouring rank's memory; compared with the same bf16 model on one GPU (SiLU, so that no ReLU
    kink amplifies the one reassociation: halo gradients are summed before instead of after the W0b^T product)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import gmp_b200
    from tests.helpers import rel_err
    world = 2
    port = 33500 + os.getpid() % 2000 + {"nccl": 0, "peer": 7, "fused": 14}[halo] + (3 if precision == "bf16" else 0)
    act = "swish" if precision == "bf16" else "relu"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, halo, precision, act)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=250) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    dev = torch.device("cuda", 0)
    pos, h, cot_h, cot_p = (t.to(dev) for t in _data())
    ei = gmp_b200.radius_graph(pos, 1.0, None, max_num_neighbors=128)
    assert sum(r[6] for r in res) == ei.shape[1]
    torch.manual_seed(1)
    gmp_b200.set_fast_matmul(precision == "bf16")
    model = gmp_b200.PartitionedEGNN(num_layers=2, emb_dim=128, precision=precision, activation=act).to(dev)
    part1 = gmp_b200.slab_partition(pos[:, 0], 1.0, 0, 1)
    hh, pp = h.clone().requires_grad_(True), pos.clone().requires_grad_(True)
    ho, po = model(hh, pp, ei, part1)
    ((ho * cot_h).sum() + (po * cot_p).sum()).backward()
    cat = lambda k: torch.cat([torch.from_numpy(r[k]) for r in res])
    # bf16: the same kernels on the same rows, but the tiles are cut elsewhere and the tensor core's accumulation inside a
    # K block (the one-hot segment sums) is not a sequential fp32 sum: partitioned and whole-graph runs agree to ~3e-4
    t_out, t_grad = (1e-5, 5e-5) if precision == "fp32" else (2e-3, 1e-2)
    assert rel_err(cat(1), ho) <= t_out and rel_err(cat(2) - pos.cpu(), po - pos) <= t_out
    assert rel_err(cat(3), hh.grad) <= t_grad and rel_err(cat(4), pp.grad) <= t_grad
    for gp, p in zip(res[0][5], model.parameters()):
        assert rel_err(torch.from_numpy(gp), p.grad) <= t_grad
    gmp_b200.set_fast_matmul(False)
