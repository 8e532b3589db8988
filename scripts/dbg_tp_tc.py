"""Debug driver for the tcgen05 tensor-product convolution: small case vs the fp32 kernels, then timing at size."""
import sys, time
import torch
sys.path.insert(0, ".")
import gmp_b200
from tests.helpers import random_clouds, rel_err


def run(C, graphs, nodes, mlp=256, gate=True, time_it=False, bwd=True):
    d = random_clouds(graphs, nodes, 3.0 if not time_it else 4.0, 1.9 if not time_it else 2.0, 700 + C)
    ei, pos = d["edge_index"].cuda(), d["pos"].cuda()
    hid, sh_ir = f"{C}x0e+{C}x1o+{C}x2e", "1x0e+1x1o+1x2e"
    torch.manual_seed(C)
    m32 = gmp_b200.TensorProductConvLayer(hid, hid, sh_ir, 8, mlp, gate=gate).cuda()
    with torch.no_grad():
        m32.fc[2].bias.normal_(0, 0.05)
    m16 = gmp_b200.TensorProductConvLayer(hid, hid, sh_ir, 8, mlp, gate=gate, precision="bf16").cuda()
    m16.load_state_dict(m32.state_dict())
    esh, eft = gmp_b200.edge_geometry(pos, ei, 2, gmp_b200.RadialEmbeddingBlock(2.0, 8, 5))
    x = torch.randn(pos.shape[0], 9 * C, device="cuda")
    x16 = x.clone().requires_grad_(True)
    E = ei.shape[1]
    print(f"C={C} N={pos.shape[0]} E={E} numel={m16.tp.weight_numel}", flush=True)
    o16 = m16(x16, ei, esh, eft)
    torch.cuda.synchronize()
    print("  tc forward done", flush=True)
    if not time_it:
        x32 = x.clone().requires_grad_(True)
        o32 = m32(x32, ei, esh, eft)
        print("  fwd rel err", rel_err(o16, o32), flush=True)
        if bwd:
            cot = torch.randn_like(o32)
            g32 = torch.autograd.grad((o32 * cot).sum(), [x32] + list(m32.parameters()))
            g16 = torch.autograd.grad((o16 * cot).sum(), [x16] + list(m16.parameters()))
            torch.cuda.synchronize()
            for a, b, name in zip(g16, g32, ["node_attr"] + [k for k, _ in m32.named_parameters()]):
                print("  grad", name, rel_err(a, b), flush=True)
    else:
        for what in ("fwd", "fwd+bwd") if bwd else ("fwd",):
            ts = []
            for it in range(4):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                o = m16(x16, ei, esh, eft)
                if what != "fwd":
                    o.square().sum().backward()
                torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
            ms = min(ts) * 1e3
            fl = 2.0 * mlp * m16.tp.weight_numel * E * (1 if what == "fwd" else 4)
            print(f"  {what}: {ms:.2f} ms  {E / ms * 1e3:.3e} edges/s  fc.2-equivalent {fl / ms * 1e-9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "check"
    if mode == "prof":
        C = int(sys.argv[2]) if len(sys.argv) > 2 else 64
        d = random_clouds(256 if C == 64 else 128, 64, 4.0, 2.0, 700 + C)
        ei, pos = d["edge_index"].cuda(), d["pos"].cuda()
        hid = f"{C}x0e+{C}x1o+{C}x2e"
        m = gmp_b200.TensorProductConvLayer(hid, hid, "1x0e+1x1o+1x2e", 8, 256, gate=C == 64, precision="bf16").cuda()
        esh, eft = gmp_b200.edge_geometry(pos, ei, 2, gmp_b200.RadialEmbeddingBlock(2.0, 8, 5))
        x = torch.randn(pos.shape[0], 9 * C, device="cuda", requires_grad=True)
        for _ in range(2):
            m(x, ei, esh, eft).square().sum().backward()
        torch.cuda.synchronize()
    elif mode == "check":
        run(16, 3, 12, mlp=64, gate=False)
        run(64, 6, 24)
        run(128, 40, 16, gate=False)
    else:
        run(64, 256, 64, time_it=True, bwd=len(sys.argv) > 2)
        run(128, 128, 64, gate=False, time_it=True, bwd=len(sys.argv) > 2)
