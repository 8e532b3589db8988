"""Where the end-to-end step's extra millisecond goes: replay of the resident graph vs the rebuild graph vs upload + load."""
import os, sys
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
ei = gmp_b200.radius_graph(pos.to(dev), CFG["cutoff"], batch.to(dev), max_num_neighbors=CFG["max_num_neighbors"])
res = gmp_b200.Batch(atoms=atoms.to(dev), pos=pos.to(dev), batch=batch.to(dev), edge_index=ei, num_graphs=CFG["molecules"])
host = gmp_b200.Batch(atoms=atoms, pos=pos, batch=batch, edge_index=ei.cpu(), num_graphs=CFG["molecules"]).pin_memory()
gs = gmp_b200.GraphedStep(model, res)
gs2 = gmp_b200.GraphedStep(model, host.to(dev), warmup=2, rebuild_graph=True)
out_host = torch.empty(CFG["molecules"], 1).pin_memory()


def T(name, fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    print(f"{name:42s} {s.elapsed_time(e) / n:.3f} ms", flush=True)


T("replay, resident graph", gs.replay)
T("replay, rebuild graph (CSR sort inside)", gs2.replay)
T("upload (H2D, main stream) + load", lambda: gs2.load(host))
T("upload + load + replay + read-back", lambda: (gs2.load(host), out_host.copy_(gs2.replay().detach(), non_blocking=True)))


def pre(n=20):
    for staged in gmp_b200.DevicePrefetcher((host for _ in range(n)), dev, static=True):
        gs2.load(staged)
        out_host.copy_(gs2.replay().detach(), non_blocking=True)


for _ in range(2):
    pre(5)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    pre(20)
    e.record()
    torch.cuda.synchronize()
    print(f"{'prefetcher(static) + load + replay':42s} {s.elapsed_time(e) / 20:.3f} ms", flush=True)
