"""BASELINE config 1 (EGNN, 5 layers, hidden 128, k-chains k = 4, 64 graphs: N = 384, E = 640) on one GPU: the step is pure
launch latency, so this is the case for replaying it as one CUDA graph.  Inputs: the committed golden
fixture (tests/golden/egnn_model_kchains.pt).  usage: python scripts/bench_config1.py [fp32|bf16]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gmp_b200
from tests.helpers import load_golden, load_params

precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
fx = load_golden("egnn_model_kchains")
ctor = dict(fx["ctor"])
model = load_params(gmp_b200.EGNNModel(**ctor), fx["state"]).cuda()
if precision == "bf16":
    for layer in model.convs:
        layer.precision = "bf16"
i = fx["inputs"]
b = gmp_b200.Batch(atoms=i["atoms"].cuda(), pos=i["pos"].cuda(), edge_index=i["edge_index"].cuda(), batch=i["batch"].cuda(),
                   num_graphs=int(i["batch"].max()) + 1)
params = [p for p in model.parameters()]


def eager():
    for p in params:
        p.grad = None
    out = model(b)
    out.sum().backward()
    return out


def timeit(fn, n):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


out_e = eager().detach().clone()
g_e = [p.grad.detach().clone() for p in params]
ms_eager = timeit(eager, 200)
gs = gmp_b200.GraphedStep(model, b)
out_g = gs.replay()
same = torch.equal(out_g, out_e) and all(torch.equal(a, c) for a, c in zip(gs.grads, g_e))
ms_graph = timeit(gs.replay, 2000)
E, L = b.edge_index.shape[1], ctor.get("num_layers", 5)
print(json.dumps({"workload": "EGNN 5 layers d=128, k-chains k=4 x 64 graphs (BASELINE config 1)", "precision": precision,
                  "nodes": int(b.pos.shape[0]), "edges": E, "ms_per_step_eager": ms_eager, "ms_per_step_cuda_graph": ms_graph,
                  "speedup": ms_eager / ms_graph, "edges_per_s_per_layer_eager": E * L / (ms_eager * 1e-3),
                  "edges_per_s_per_layer_cuda_graph": E * L / (ms_graph * 1e-3), "kernels_per_replay": gs.kernels_per_replay,
                  "replay_bit_identical_to_eager": bool(same)}))
