"""Three SchNet steps (BASELINE config 2, bf16 mode) and nothing else: for an ncu launch list of the step."""
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
atoms, pos, batch = (t.to(dev) for t in bench.synth(CFG["molecules"], seed=0))
ei = gmp_b200.radius_graph(pos, CFG["cutoff"], batch, max_num_neighbors=CFG["max_num_neighbors"])
b = bench.Bag(atoms=atoms, pos=pos, edge_index=ei, batch=batch, num_graphs=CFG["molecules"])
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    for p in model.parameters():
        p.grad = None
    model(b).sum().backward()
torch.cuda.synchronize()
