"""Run one fused kernel standalone at the BASELINE config-2 size (for ncu captures and quick timing).
usage: python scripts/prof_kernel.py schnet_fwd|schnet_fwd2|schnet_fwd2k|schnet_bwd|schnet_bwd2 [fp32|bf16] [reps]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import gmp_b200
from gmp_b200._lib import SchnetFilter, call, ptr

which = sys.argv[1] if len(sys.argv) > 1 else "schnet_fwd"
prec = {"fp32": 0, "bf16": 1}[sys.argv[2] if len(sys.argv) > 2 else "bf16"]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda")
atoms, pos, batch = bench.synth(4096, 0)
pos, batch = pos.to(dev), batch.to(dev)
ei = gmp_b200.radius_graph(pos, 5.0, batch, max_num_neighbors=32)
N, E, F = pos.shape[0], ei.shape[1], 128
g = gmp_b200.get_graph(ei, N)
csr = g.by_dst
torch.manual_seed(0)
blk = gmp_b200.InteractionBlock(128, 50, 128, 5.0).to(dev)
sm = gmp_b200.GaussianSmearing(0.0, 5.0, 50).to(dev)
w1, b1, w2, b2 = blk.mlp[0].weight, blk.mlp[0].bias, blk.mlp[2].weight, blk.mlp[2].bias
filt = SchnetFilter(ptr(w1), ptr(b1), ptr(w2), ptr(b2), 50, F, 5.0, ptr(sm.offset), sm.coeff)
ew = gmp_b200.edge_length(pos, g)
x1 = torch.randn(N, F, device=dev)
agg = torch.empty(N, F, device=dev)
gout = torch.randn(N, F, device=dev)
lib = gmp_b200._lib.lib()
parts = torch.empty(lib.gmp_schnet_bwd_num_parts(E), lib.gmp_schnet_bwd_part_len(50, F), device=dev)
x1b = x1.to(torch.bfloat16)
head = torch.empty(lib.gmp_schnet_tc2_num_chunks(E), F, device=dev)
rowid = csr.row_ids()
keep = torch.empty(E, F, dtype=torch.bfloat16, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
times = []
for it in range(reps + 2):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    if which == "schnet_fwd2k":   # the training variant: also stores the per-edge filter values
        call("gmp_schnet_cfconv_fwd_tc2_keep", ptr(csr.rowptr), ptr(csr.col), csr.perm_ptr, ptr(rowid), N, E, ptr(ew), ptr(x1b),
             C.byref(filt), ptr(agg), ptr(head), ptr(keep), ptr(g.by_src.inv_perm()))
    elif which == "schnet_fwd2":
        call("gmp_schnet_cfconv_fwd_tc2", ptr(csr.rowptr), ptr(csr.col), csr.perm_ptr, ptr(rowid), N, E, ptr(ew), ptr(x1b),
             C.byref(filt), ptr(agg), ptr(head))
    elif which == "schnet_bwd2":
        call("gmp_schnet_cfconv_bwd_tc2", ptr(csr.rowptr), ptr(csr.col), csr.perm_ptr, ptr(rowid), N, E, ptr(ew), ptr(x1b),
             C.byref(filt), ptr(gout), ptr(parts), parts.shape[0])
    elif which == "schnet_fwd":
        call("gmp_schnet_cfconv_fwd", ptr(csr.rowptr), ptr(csr.col), csr.perm_ptr, N, E, ptr(ew), None, ptr(x1), C.byref(filt),
             ptr(agg), prec)
    else:
        call("gmp_schnet_cfconv_bwd", ptr(csr.rowptr), ptr(csr.col), csr.perm_ptr, N, E, ptr(ew), None, ptr(x1), C.byref(filt),
             ptr(gout), ptr(parts), None, None, prec)
    e.record()
    torch.cuda.synchronize()
    if it >= 2:
        times.append(s.elapsed_time(e))
print(f"{which} prec={prec} N={N} E={E}: {sum(times) / len(times):.4f} ms/launch (min {min(times):.4f})")
