"""Debug driver for the tcgen05 EGNN edge kernels: small case vs the fp32 kernels, then timing at size."""
import sys, time
import torch
sys.path.insert(0, ".")
import gmp_b200
from tests.helpers import random_clouds, rel_err


def make(n_side, n, seed=0):
    g = torch.Generator().manual_seed(seed)
    pos = (torch.rand(n, 3, generator=g) * n_side).cuda()
    ei = gmp_b200.radius_graph(pos, 1.0, None, max_num_neighbors=128)
    return pos, ei


def check(act="relu", aggr="add", n=3000, side=7.0, bwd=True):
    pos, ei = make(side, n)
    torch.manual_seed(1)
    m32 = gmp_b200.EGNNLayer(128, activation=act, aggr=aggr).cuda()
    with torch.no_grad():
        for p in m32.parameters():
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.2)
    m16 = gmp_b200.EGNNLayer(128, activation=act, aggr=aggr, precision="bf16").cuda()
    m16.load_state_dict(m32.state_dict())
    h = torch.randn(n, 128, device="cuda")
    h32, h16 = h.clone().requires_grad_(True), h.clone().requires_grad_(True)
    p32, p16 = pos.clone().requires_grad_(True), pos.clone().requires_grad_(True)
    o32, q32 = m32(h32, p32, ei)
    o16, q16 = m16(h16, p16, ei)
    torch.cuda.synchronize()
    print(f"{act}/{aggr} N={n} E={ei.shape[1]}  fwd rel err h {rel_err(o16, o32):.2e}  pos {rel_err(q16 - pos, q32 - pos):.2e}", flush=True)
    if bwd:
        c1, c2 = torch.randn_like(o32), torch.randn_like(q32)
        g32 = torch.autograd.grad((o32 * c1).sum() + (q32 * c2).sum(), [h32, p32] + list(m32.parameters()))
        g16 = torch.autograd.grad((o16 * c1).sum() + (q16 * c2).sum(), [h16, p16] + list(m16.parameters()))
        torch.cuda.synchronize()
        names = ["h", "pos"] + [k for k, _ in m32.named_parameters()]
        for a, b, k in zip(g16, g32, names):
            print(f"   grad {k:22s} {rel_err(a, b):.2e}", flush=True)


def timing(n=2 ** 18, side=32.0, bwd=True):
    gmp_b200.set_fast_matmul(True)
    pos, ei = make(side, n)
    E = ei.shape[1]
    for prec in ("bf16",):
        m = gmp_b200.EGNNLayer(128, precision=prec).cuda()
        h = torch.randn(n, 128, device="cuda", requires_grad=True)
        p = pos.clone().requires_grad_(True)
        for what in (("fwd", "fwd+bwd") if bwd else ("fwd",)):
            ts = []
            for _ in range(4):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                if what == "fwd":
                    with torch.no_grad():
                        m(h, p, ei)
                else:
                    o, q = m(h, p, ei)
                    (o.sum() + q.sum()).backward()
                torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
            ms = min(ts) * 1e3
            print(f"{prec} N={n} E={E} {what}: {ms:.2f} ms  {E / ms * 1e3:.3e} edges/s", flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "check"
    if mode == "check":
        check("relu", "add", bwd=len(sys.argv) > 2)
        check("swish", "mean", n=1000, side=5.0, bwd=len(sys.argv) > 2)
    else:
        timing(bwd=len(sys.argv) > 2)
    if mode == "prof":
        pos, ei = make(32.0, 2 ** 18)
        m = gmp_b200.EGNNLayer(128, precision="bf16").cuda()
        h = torch.randn(pos.shape[0], 128, device="cuda", requires_grad=True)
        p = pos.clone().requires_grad_(True)
        for _ in range(2):
            o, q = m(h, p, ei)
            (o.sum() + q.sum()).backward()
        torch.cuda.synchronize()
