"""Where the host-buffer (e2e) step spends its time: H2D copies, CSR builds, model step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, gmp_b200
dev = torch.device("cuda")
atoms, pos, batch = bench.synth(4096, 0)
ei = gmp_b200.radius_graph(pos.to(dev), 5.0, batch.to(dev), max_num_neighbors=32)
ei_p, pos_p, atoms_p, batch_p = ei.cpu().pin_memory(), pos.pin_memory(), atoms.pin_memory(), batch.pin_memory()
def T(fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("h2d edge_index 29MB: %.2f ms" % T(lambda: ei_p.to(dev, non_blocking=True)))
print("h2d rest: %.2f ms" % T(lambda: (pos_p.to(dev, non_blocking=True), atoms_p.to(dev, non_blocking=True), batch_p.to(dev, non_blocking=True))))
N = pos.shape[0]
print("build_csr by_dst (sorted): %.2f ms" % T(lambda: gmp_b200.build_csr(ei[1], ei[0], N)))
print("build_csr by_src (unsorted): %.2f ms" % T(lambda: gmp_b200.build_csr(ei[0], ei[1], N)))
b = batch.to(dev)
print("build_csr batch: %.2f ms" % T(lambda: gmp_b200.build_csr(b, b, 4096)))
from gmp_b200._lib import call, ptr
idx = ei[0].contiguous()
counts = torch.empty(N, dtype=torch.int32, device=dev)
print(" csr_count: %.2f ms" % T(lambda: call("gmp_csr_count", ptr(idx), idx.numel(), N, ptr(counts))))
rowptr = gmp_b200.graph.exclusive_scan(counts).to(torch.int32)
print(" scan: %.2f ms" % T(lambda: gmp_b200.graph.exclusive_scan(counts)))
perm, tmp, cur = (torch.empty(idx.numel(), dtype=torch.int32, device=dev) for _ in range(3))
cursor = torch.empty(N, dtype=torch.int32, device=dev)
print(" csr_fill: %.2f ms" % T(lambda: call("gmp_csr_fill", ptr(idx), idx.numel(), N, ptr(rowptr), ptr(cursor), ptr(tmp), ptr(perm))))
