"""e2e step variants for BASELINE config 2: synchronous upload vs DevicePrefetcher (timing experiment)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import gmp_b200
dev = torch.device("cuda")
CFG = bench.CFG
gmp_b200.set_fast_matmul(True)
torch.manual_seed(0)
model = gmp_b200.SchNetModel(hidden_channels=CFG["hidden"], num_filters=CFG["filters"], num_layers=CFG["layers"],
                             num_gaussians=CFG["gaussians"], cutoff=CFG["cutoff"], precision="bf16").to(dev)
atoms, pos, batch = bench.synth(CFG["molecules"], seed=0)
ei = gmp_b200.radius_graph(pos.to(dev), CFG["cutoff"], batch.to(dev), max_num_neighbors=CFG["max_num_neighbors"]).cpu()
host = gmp_b200.Batch(atoms=atoms, pos=pos, batch=batch, edge_index=ei, num_graphs=CFG["molecules"]).pin_memory()
out_host = torch.empty(CFG["molecules"], 1).pin_memory()


def step(b):
    for p in model.parameters():
        p.grad = None
    out = model(b)
    out.sum().backward()
    return out


def sync_steps(n):
    for _ in range(n):
        out = step(host.to(dev, non_blocking=True))
        out_host.copy_(out.detach(), non_blocking=True)


def pre_steps(n):
    for bb in gmp_b200.DevicePrefetcher((host for _ in range(n)), dev):
        out = step(bb)
        out_host.copy_(out.detach(), non_blocking=True)


def resident_steps(n):
    b = host.to(dev)
    for _ in range(n):
        step(b)


for name, fn in (("resident", resident_steps), ("sync upload", sync_steps), ("prefetcher", pre_steps), ("sync upload", sync_steps),
                 ("prefetcher", pre_steps)):
    fn(5)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s.record()
    fn(20)
    e.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"{name:12s} {s.elapsed_time(e) / 20:.3f} ms/step (host enqueue {(t1 - t0) * 50:.3f} ms/step)", flush=True)
