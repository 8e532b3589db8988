"""BASELINE configs 3 and 4 at full size, whole model forward + backward:
  tfn  : TFNModel(max_ell=2, emb_dim=64, num_layers=4, r_max=2.0), 2048 clouds x 64 points        (config 3)
  mace : MACEModel(max_ell=2, correlation=3, emb_dim=128, num_layers=2, r_max=2.0), 1024 clouds x 64 (config 4)
Graph-sharded over the ranks (strong scaling: the cloud count is fixed, ranks split it), gradients all-reduced over
NCCL, e3nn BatchNorm statistics all-reduced in forward (MACE default) so the sharded result equals the single-process one.
Launch: python -m torch.distributed.run --nproc-per-node N scripts/bench_config34.py tfn|mace [fp32|bf16] [steps]
Prints one JSON line (rank 0): edges/s per layer forward+backward (whole job), max over ranks."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import gmp_b200
from tests.helpers import Bag

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
which = sys.argv[1] if len(sys.argv) > 1 else "mace"
precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"
K = int(sys.argv[3]) if len(sys.argv) > 3 else 3
clouds = {"tfn": 2048, "mace": 1024}[which]
mine = clouds // world
g = torch.Generator().manual_seed(0)
pos_all = torch.rand(clouds * 64, 3, generator=g) * 4.0
pos = pos_all[rank * mine * 64:(rank + 1) * mine * 64].contiguous().to(dev)
batch = torch.arange(mine).repeat_interleave(64).to(dev)
ei = gmp_b200.radius_graph(pos, 2.0, batch, max_num_neighbors=64)
atoms = torch.zeros(pos.shape[0], dtype=torch.long, device=dev)
gmp_b200.set_fast_matmul(precision == "bf16")   # bf16 mode: node-side library GEMMs on TF32 tensor cores
torch.manual_seed(0)
if which == "tfn":
    model = gmp_b200.TFNModel(r_max=2.0, max_ell=2, emb_dim=64, num_layers=4, precision=precision).to(dev)
    layers = 4
else:
    model = gmp_b200.MACEModel(r_max=2.0, max_ell=2, correlation=3, emb_dim=128, num_layers=2, precision=precision).to(dev)
    layers = 2
if world > 1:
    for m in model.modules():
        if isinstance(m, gmp_b200.tfn.BatchNorm):
            m.process_group = dist.group.WORLD
params = [p for p in model.parameters()]
b = Bag(atoms=atoms, pos=pos, edge_index=ei, batch=batch)


def step():
    for p in params:
        p.grad = None
    model(b).sum().backward()
    if world > 1:
        gmp_b200.allreduce_gradients(params)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(2):
    step()
barrier()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(K):
    step()
e.record()
barrier()
t = torch.tensor([s.elapsed_time(e) / K], device=dev, dtype=torch.float64)
E_all = torch.tensor([float(ei.shape[1])], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(E_all)
if rank == 0:
    ms = t.item()
    print(json.dumps({"workload": f"{which.upper()} model, BASELINE config {3 if which == 'tfn' else 4}: {clouds} clouds x 64 points, "
                                  f"{layers} layers, graph-sharded x{world}", "n_gpus": world,
                      "precision": "fp32-strict" if precision == "fp32" else "bf16 tcgen05 (1e-2)", "nodes": clouds * 64,
                      "edges": int(E_all.item()), "ms_per_step": ms, "edges_per_s_per_layer": E_all.item() * layers / (ms * 1e-3),
                      "scaling": "strong"}))
if world > 1:
    dist.destroy_process_group()
