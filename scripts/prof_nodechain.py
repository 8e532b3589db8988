"""Time the node-side chain kernel (csrc/node_chain.cu) at the config-2 node count.  usage: python scripts/prof_nodechain.py [n=131072] [reps=20]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gmp_b200 import nodechain as nc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
A, H = torch.randn(n, 128, generator=g).to(dev), torch.randn(n, 128, generator=g).to(dev)
W = [(torch.randn(128, 128, generator=g) / 11).to(dev) for _ in range(3)]
b = torch.randn(128, generator=g).to(dev)
o = [torch.empty(n, 128, device=dev) for _ in range(3)]
o16 = torch.empty(n, 128, device=dev, dtype=torch.bfloat16)
imgs = [nc.pack_w(w) for w in W]
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)   # 256 MB > L2


def timed(fn, name, bytes_):
    """`reps` back-to-back launches between one event pair (the working set of a launch exceeds the L2, and a launch is
    longer than the host-side cost of issuing it, so neither the cache nor the CPU flatters / limits the figure)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.zero_()
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    print(f"{name}: {ms * 1e3:.1f} us  ({bytes_ / ms / 1e6:.0f} GB/s algorithmic)")


row = n * 512
S1 = [nc.stage(imgs[0], b, act="ssp", out_f32=o[0])]
S2 = [nc.stage(imgs[0], out_bf16=o16)]
S3 = [nc.stage(imgs[0], b, act="ssp", out_f32=o[0]), nc.stage(imgs[1], b, add_res=H, out_f32=o[1]), nc.stage(imgs[2], out_bf16=o16)]
S4 = [nc.stage(imgs[0], add_res=H, out_f32=o[0]), nc.stage(imgs[1], mul_aux=o[2], mul_mode=nc.MUL_DSSP, out_f32=o[1]),
      nc.stage(imgs[2], out_f32=o[2], out_bf16=o16)]
timed(lambda: nc.run(A, S1), "1 stage  (read 1, write 1 fp32, ssp)", 2 * row)
timed(lambda: nc.run(A, S2), "1 stage  (read 1, write bf16)", 1.5 * row)
timed(lambda: nc.run(A, S3), "SchNet forward chain (read 2, write 2 fp32 + 1 bf16)", 4.5 * row)
timed(lambda: nc.run(A, S4), "SchNet backward chain (read 3, write 3 fp32 + 1 bf16)", 6.5 * row)
torch.backends.cuda.matmul.allow_tf32 = True
timed(lambda: torch.nn.functional.linear(A, W[0], b), "cuBLAS TF32 linear alone", 2 * row)
gg, xx = torch.randn(n, 128, device=dev), torch.randn(n, 128, device=dev)
timed(lambda: nc.wgrad(gg, xx), "wgrad + reduce (read 2)", 2 * row)
