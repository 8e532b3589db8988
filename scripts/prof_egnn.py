"""EGNN layer (bf16 tcgen05 edge kernels) at the config-5 geometry, 2^log2n nodes: forward and forward+backward times,
and a target for `ncu -k regex:egnn_`.   usage: python scripts/prof_egnn.py [log2n=18] [relu|swish] [reps=5]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import gmp_b200

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 18
act = sys.argv[2] if len(sys.argv) > 2 else "relu"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda")
gmp_b200.set_fast_matmul(True)
pos = bench.synth_cube(log2n).to(dev)
ei = gmp_b200.radius_graph(pos, 1.0, None, max_num_neighbors=128)
N, E = pos.shape[0], ei.shape[1]
torch.manual_seed(0)
layer = gmp_b200.EGNNLayer(128, activation=act, precision="bf16").to(dev)
h = torch.randn(N, 128, device=dev, requires_grad=True)
p = pos.clone().requires_grad_(True)


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def fwd():
    with torch.no_grad():
        layer(h, p, ei)


def step():
    layer.zero_grad(set_to_none=True)
    h.grad = p.grad = None
    out, pp = layer(h, p, ei)
    (out.sum() + pp.sum()).backward()


ms_f, ms_s = timed(fwd), timed(step)
print(f"EGNNLayer(128, {act}) bf16 N={N} E={E}: fwd {ms_f:.3f} ms, fwd+bwd {ms_s:.3f} ms, {E / (ms_s * 1e-3):.3e} edges/s")
