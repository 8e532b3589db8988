"""BASELINE config 5 (reduced N): EGNN layer stack on ONE random 3D radius graph (density 8 per unit volume, r = 1,
~31 neighbours), destination-partitioned over the ranks with NCCL halo exchange.  Strong scaling: the graph is fixed,
ranks split it.  Launch: python -m torch.distributed.run --nproc-per-node N scripts/bench_config5.py [log2_nodes] [layers] [fp32|bf16]
Prints one JSON line (rank 0): edges/s per layer forward+backward, max over ranks."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import gmp_b200

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 4
precision = sys.argv[3] if len(sys.argv) > 3 else "fp32"   # "fp32" (strict, 1e-5) | "bf16" (tcgen05 edge GEMMs, 1e-2)
n = 2 ** log2n
side = (n / 8.0) ** (1.0 / 3.0)
g = torch.Generator().manual_seed(0)
pos = torch.rand(n, 3, generator=g) * side
pos = pos[torch.argsort(pos[:, 0])].contiguous().to(dev)   # spatial sort: slabs along x
part = gmp_b200.slab_partition(pos[:, 0], 1.0, rank, world)
ei = gmp_b200.distributed.local_radius_graph(pos[part.local_global].contiguous(), 1.0, part)
E_loc = torch.tensor([float(ei.shape[1])], device=dev, dtype=torch.float64)
gmp_b200.set_fast_matmul(precision == "bf16")   # bf16 mode: node-side library GEMMs on TF32 tensor cores
torch.manual_seed(0)
model = gmp_b200.PartitionedEGNN(num_layers=layers, emb_dim=128, precision=precision).to(dev)
params = list(model.parameters())
h_own = torch.randn(part.n_own, 128, device=dev)
p_own = pos[part.own_lo:part.own_hi].clone()


def step():
    for p in params:
        p.grad = None
    ho, po = model(h_own, p_own, ei, part)
    (ho.sum() + po.sum()).backward()
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
K = 3
s.record()
for _ in range(K):
    step()
e.record()
barrier()
t = torch.tensor([s.elapsed_time(e) / K], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(E_loc)
if rank == 0:
    ms = t.item()
    print(json.dumps({"workload": f"EGNN {layers} layers d=128 on one radius graph N=2^{log2n}, r=1, density 8 (config 5 geometry), "
                                  "destination-partitioned slabs + NCCL halo exchange", "n_gpus": world, "precision": "fp32-strict" if precision == "fp32" else "bf16 tcgen05 (1e-2)",
                      "nodes": n, "edges": int(E_loc.item()), "halo_nodes_rank0": part.n_left + part.n_right,
                      "ms_per_step": ms, "edges_per_s_per_layer": E_loc.item() * layers / (ms * 1e-3), "scaling": "strong"}))
if world > 1:
    dist.destroy_process_group()
