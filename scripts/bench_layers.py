"""Per-layer forward+backward timings of the fused layers at the BASELINE config shapes: fp32-strict kernels (reduced
graph counts where the full size would take minutes) and the tcgen05 path (precision="bf16", full config sizes).
Prints one JSON line per layer.
usage: python scripts/bench_layers.py [egnn] [tfn] [mace] [tfn_tc] [mace_tc] [symc]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gmp_b200

dev = torch.device("cuda")
which = sys.argv[1:] or ["egnn", "tfn", "mace", "symc"]


def timeit(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def clouds(graphs, nodes, box, seed=0):
    g = torch.Generator().manual_seed(seed)
    pos = (torch.rand(graphs * nodes, 3, generator=g) * box).to(dev)
    batch = torch.arange(graphs).repeat_interleave(nodes).to(dev)
    return pos, batch


if "egnn" in which:
    # config 5 geometry: density 8 / unit volume, r = 1 (~31 neighbours), one connected cube; N = 2^18 here
    n_side = 32.0
    g = torch.Generator().manual_seed(0)
    pos = (torch.rand(2 ** 18, 3, generator=g) * n_side).to(dev)
    ei = gmp_b200.radius_graph(pos, 1.0, None, max_num_neighbors=128)
    N, E = pos.shape[0], ei.shape[1]
    layer = gmp_b200.EGNNLayer(128).to(dev)
    h = torch.randn(N, 128, device=dev, requires_grad=True)
    p = pos.clone().requires_grad_(True)

    def step():
        layer.zero_grad(set_to_none=True)
        out, pp = layer(h, p, ei)
        (out.sum() + pp.sum()).backward()
    ms = timeit(step)
    flops = 3 * (E * (2 * 2 * 128 * 128 + 4 * 128) + N * 6 * 128 * 128)  # fused form: two 128x128 edge GEMMs
    print(json.dumps({"layer": "EGNNLayer(128) fp32-strict", "workload": f"one radius graph N={N} E={E} (config-5 geometry)",
                      "ms_fwd_bwd": ms, "edges_per_s": E / (ms * 1e-3), "tflops_fp32": flops / (ms * 1e-3) / 1e12}))

for name, C, graphs in (("tfn", 64, 64), ("mace", 128, 32)):
    if name not in which:
        continue
    pos, batch = clouds(graphs, 64, 4.0)
    ei = gmp_b200.radius_graph(pos, 2.0, batch, max_num_neighbors=64)
    N, E = pos.shape[0], ei.shape[1]
    hid = f"{C}x0e+{C}x1o+{C}x2e"
    conv = gmp_b200.TensorProductConvLayer(hid, hid, "1x0e+1x1o+1x2e", 8, 256, gate=(name == "tfn"), batch_norm=(name == "mace")).to(dev)
    sh, ft = gmp_b200.edge_geometry(pos, ei, 2, gmp_b200.RadialEmbeddingBlock(2.0, 8, 5))
    x = torch.randn(N, 9 * C, device=dev, requires_grad=True)

    def step():
        conv.zero_grad(set_to_none=True)
        conv(x, ei, sh, ft).sum().backward()
    ms = timeit(step, warm=1, reps=3)
    wn = conv.tp.weight_numel
    print(json.dumps({"layer": f"TensorProductConvLayer C={C} ({name} config, layer>=1) fp32-strict",
                      "workload": f"{graphs} clouds x 64 nodes, N={N} E={E}, weight_numel={wn}", "ms_fwd_bwd": ms,
                      "edges_per_s": E / (ms * 1e-3), "tflops_fp32": 4 * 2 * 256 * wn * E / (ms * 1e-3) / 1e12}))

PEAK_TF = 1366.5  # MEASURED_PEAKS.json bf16_tflops_sustained (kernels timed inside a long step)
for name, C, graphs in (("tfn_tc", 64, 2048), ("mace_tc", 128, 1024)):
    if name not in which:
        continue
    pos, batch = clouds(graphs, 64, 4.0)
    ei = gmp_b200.radius_graph(pos, 2.0, batch, max_num_neighbors=64)
    N, E = pos.shape[0], ei.shape[1]
    hid = f"{C}x0e+{C}x1o+{C}x2e"
    conv = gmp_b200.TensorProductConvLayer(hid, hid, "1x0e+1x1o+1x2e", 8, 256, gate=(name == "tfn_tc"), batch_norm=(name == "mace_tc"),
                                           precision="bf16").to(dev)
    sh, ft = gmp_b200.edge_geometry(pos, ei, 2, gmp_b200.RadialEmbeddingBlock(2.0, 8, 5))
    x = torch.randn(N, 9 * C, device=dev, requires_grad=True)

    def fwd():
        return conv(x, ei, sh, ft)

    def step():
        conv.zero_grad(set_to_none=True)
        fwd().sum().backward()
    with torch.no_grad():
        ms_f = timeit(fwd, warm=2, reps=5)
    ms = timeit(step, warm=2, reps=5)
    wn = conv.tp.weight_numel
    gemm = 2 * 256 * wn * E  # one pass of fc's second Linear
    print(json.dumps({"layer": f"TensorProductConvLayer C={C} ({name[:-3]} config, layer>=1) bf16 tcgen05 (1e-2)",
                      "workload": f"{graphs} clouds x 64 nodes (BASELINE config size), N={N} E={E}, weight_numel={wn}",
                      "ms_fwd": ms_f, "ms_fwd_bwd": ms, "edges_per_s": E / (ms * 1e-3), "edges_per_s_fwd": E / (ms_f * 1e-3),
                      "tflops_fwd": gemm / (ms_f * 1e-3) / 1e12,
                      "tflops_fwd_bwd_algorithmic_3x": 3 * gemm / (ms * 1e-3) / 1e12,
                      "tflops_fwd_bwd_executed_4x": 4 * gemm / (ms * 1e-3) / 1e12,
                      "roofline": {"bound": "tensor", "peak": PEAK_TF, "unit": "TFLOP/s", "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained",
                                   "frac_fwd": gemm / (ms_f * 1e-3) / 1e12 / PEAK_TF,
                                   "frac_fwd_bwd_algorithmic": 3 * gemm / (ms * 1e-3) / 1e12 / PEAK_TF,
                                   "frac_fwd_bwd_executed": 4 * gemm / (ms * 1e-3) / 1e12 / PEAK_TF}}))

if "symc" in which:
    N, C = 65536, 128
    ir = f"{C}x0e+{C}x1o+{C}x2e"
    blk = gmp_b200.EquivariantProductBasisBlock(ir, ir, 3, element_dependent=False, use_sc=True).to(dev)
    x = torch.randn(N, C, 9, device=dev, requires_grad=True)
    sc = torch.randn(N, 9 * C, device=dev)

    def step():
        blk.zero_grad(set_to_none=True)
        blk(x, sc, None).sum().backward()
    ms = timeit(step)
    print(json.dumps({"layer": "EquivariantProductBasisBlock C=128 corr=3 (config 4 node side)", "workload": f"N={N}",
                      "ms_fwd_bwd": ms, "nodes_per_s": N / (ms * 1e-3)}))

if "uvu" in which:
    # the ACEsuit-style interaction on the config-4 graph (1024 clouds x 64 points, C = 128): the fused uvu
    # kernels alone (HBM-bound: per edge 11 C fp32 radial weights streamed once, a 9 C row gathered; per node 35 C written)
    # and the whole RealAgnosticInteractionBlock
    graphs, C = 1024, 128
    pos, batch = clouds(graphs, 64, 4.0)
    ei = gmp_b200.radius_graph(pos, 2.0, batch, max_num_neighbors=64)
    N, E = pos.shape[0], ei.shape[1]
    ir = f"{C}x0e+{C}x1o+{C}x2e"
    tp = gmp_b200.UVUTensorProduct(ir, "1x0e+1x1o+1x2e", ir).to(dev)
    x = torch.randn(N, 9 * C, device=dev, requires_grad=True)
    sh, _ = gmp_b200.edge_geometry(pos, ei, 2, gmp_b200.RadialEmbeddingBlock(2.0, 8, 5))
    w = torch.randn(E, 11 * C, device=dev, requires_grad=True)
    HBM = 6545.0

    def fwd():
        return tp(x, ei, sh, w)

    def step():
        x.grad = w.grad = None
        fwd().sum().backward()
    with torch.no_grad():
        ms_f = timeit(fwd)
    ms = timeit(step)
    b_f = E * (11 * C * 4 + 36 + 8) + N * (35 * C * 4 + 9 * C * 4)
    b_all = b_f + E * (2 * 11 * C * 4 + 2 * 36) + N * (2 * 35 * C * 4 + 9 * C * 4)   # dx pass re-reads w, dw pass writes [E, 11 C]
    print(json.dumps({"layer": "uvu tensor product + receiver sum (csrc/uvu.cu), C = 128", "workload": f"N={N} E={E}",
                      "ms_fwd": ms_f, "ms_fwd_bwd": ms, "edges_per_s": E / (ms * 1e-3),
                      "roofline": {"bound": "hbm", "peak": HBM, "unit": "GB/s", "frac_fwd": b_f / (ms_f * 1e-3) / 1e9 / HBM,
                                   "frac_fwd_bwd": b_all / (ms * 1e-3) / 1e9 / HBM, "algorithmic_bytes_fwd": b_f}}))
    blk = gmp_b200.RealAgnosticInteractionBlock(node_attrs_irreps="4x0e", node_feats_irreps=ir, edge_attrs_irreps="1x0e+1x1o+1x2e",
                                                edge_feats_irreps="8x0e", target_irreps=ir, hidden_irreps=ir, avg_num_neighbors=17.0).to(dev)
    attrs = torch.eye(4, device=dev)[torch.randint(0, 4, (N,), device=dev)]
    feats = gmp_b200.RadialEmbeddingBlock(2.0, 8, 5)((pos[ei[0]] - pos[ei[1]]).norm(dim=-1, keepdim=True))

    def bstep():
        blk.zero_grad(set_to_none=True)
        x.grad = None
        blk(attrs, x, sh, feats, ei)[0].sum().backward()
    msb = timeit(bstep)
    print(json.dumps({"layer": "RealAgnosticInteractionBlock C = 128 (blocks.py:396-459)", "workload": f"N={N} E={E}", "ms_fwd_bwd": msb,
                      "edges_per_s": E / (msb * 1e-3)}))

if "gvp" in which:
    # GVPConvLayer at the reference defaults (s 128, v 16, edge 32 / 1) on the config-2 graph
    pos, batch = clouds(4096, 32, 8.0)
    ei = gmp_b200.radius_graph(pos, 5.0, batch, max_num_neighbors=32)
    N, E = pos.shape[0], ei.shape[1]
    import torch.nn.functional as F_
    layer = gmp_b200.GVPConvLayer((128, 16), (32, 1), activations=(F_.relu, None)).to(dev).eval()
    s = torch.randn(N, 128, device=dev, requires_grad=True)
    v = torch.randn(N, 16, 3, device=dev, requires_grad=True)
    es, ev = torch.randn(E, 32, device=dev), torch.randn(E, 1, 3, device=dev)

    def gstep():
        layer.zero_grad(set_to_none=True)
        s.grad = v.grad = None
        os_, ov = layer((s, v), ei, (es, ev))
        (os_.sum() + ov.sum()).backward()
    msg = timeit(gstep)
    print(json.dumps({"layer": "GVPConvLayer (128, 16) / (32, 1), unfused perceptron stack + deterministic mean aggregation",
                      "workload": f"N={N} E={E}", "ms_fwd_bwd": msg, "edges_per_s": E / (msg * 1e-3)}))
