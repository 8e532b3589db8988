"""GPU: the tcgen05 / TMEM path.  (1) hardware self test of the hand-built UMMA descriptors and 128-byte swizzle;
(2) the bf16 tensor-core CFConv forward against the fp32-strict kernel: 1e-2 normwise relative, the tolerance
BASELINE.json states for bf16 MLP inputs."""
import ctypes as C

import pytest
import torch

from tests.helpers import random_clouds, rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]


@pytest.mark.parametrize("K", [64, 128])
def test_umma_selftest(K):
    from gmp_b200._lib import call, ptr
    g = torch.Generator().manual_seed(K)
    A = torch.randn(128, K, generator=g).cuda()
    B = torch.randn(128, K, generator=g).cuda()
    out = torch.zeros(128, 128, device="cuda")
    call("gmp_umma_selftest", ptr(A), ptr(B), ptr(out), K)
    torch.cuda.synchronize()
    ref = A.bfloat16().double() @ B.bfloat16().double().T
    assert rel_err(out, ref) <= 1e-5


@pytest.mark.parametrize("N", [64, 128])
def test_umma_selftest_mn_major(N):
    from gmp_b200._lib import call, ptr
    g = torch.Generator().manual_seed(N)
    A = torch.randn(128, 128, generator=g).cuda()   # [k, m]
    B = torch.randn(128, N, generator=g).cuda()     # [k, n]
    out = torch.zeros(128, N, device="cuda")
    call("gmp_umma_selftest_mn", ptr(A), ptr(B), ptr(out), N)
    torch.cuda.synchronize()
    ref = A.bfloat16().double().T @ B.bfloat16().double()
    assert rel_err(out, ref) <= 1e-5


@pytest.mark.parametrize("graphs,nodes,lazy", [(64, 32, True), (9, 21, False)])
def test_cfconv_bf16_tc_vs_fp32(graphs, nodes, lazy):
    import gmp_b200
    d = random_clouds(graphs, nodes, 8.0, 5.0, 500 + graphs, max_nb=32)
    ei, pos = d["edge_index"].cuda(), d["pos"].cuda()
    torch.manual_seed(0)
    m32 = gmp_b200.InteractionBlock(128, 50, 128, 5.0).cuda()
    with torch.no_grad():
        for p in m32.parameters():
            if p.dim() == 1:
                p.normal_(0, 0.3)
    m16 = gmp_b200.InteractionBlock(128, 50, 128, 5.0, precision="bf16").cuda()
    m16.load_state_dict(m32.state_dict())
    sm = gmp_b200.GaussianSmearing(0.0, 5.0, 50).cuda()
    x = torch.randn(pos.shape[0], 128, device="cuda")
    ew = (pos[ei[0]] - pos[ei[1]]).norm(dim=-1)
    attr = sm.lazy() if lazy else sm(ew)
    x32, x16 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    o32, o16 = m32(x32, ei, ew, attr), m16(x16, ei, ew, attr)
    assert rel_err(o16, o32) <= 1e-2
    cot = torch.randn_like(o32)
    g32 = torch.autograd.grad((o32 * cot).sum(), [x32] + list(m32.parameters()))
    g16 = torch.autograd.grad((o16 * cot).sum(), [x16] + list(m16.parameters()))
    for a, b in zip(g16, g32):
        assert rel_err(a, b) <= 1e-2
    # deterministic
    assert torch.equal(o16, m16(x16, ei, ew, attr))


@pytest.mark.parametrize("C,gate,graphs,nodes,shuffle,mlp", [(64, True, 6, 24, True, 256), (16, False, 3, 12, False, 64),
                                                            (128, False, 40, 16, False, 256)])
def test_tp_conv_bf16_tc_vs_fp32(C, gate, graphs, nodes, shuffle, mlp):
    """TFN / MACE tensor-product convolution: tcgen05 path (bf16 operands of fc's second Linear, bf16 factor) against
    the fp32-strict kernels, forward, d/d node_attr and parameter gradients: 1e-2 normwise relative."""
    import gmp_b200
    d = random_clouds(graphs, nodes, 3.0, 1.9, 700 + C)
    ei, pos = d["edge_index"], d["pos"]
    if shuffle:
        ei = ei[:, torch.randperm(ei.shape[1], generator=torch.Generator().manual_seed(1))]
    ei, pos = ei.cuda(), pos.cuda()
    hid, sh_ir = f"{C}x0e+{C}x1o+{C}x2e", "1x0e+1x1o+1x2e"
    torch.manual_seed(C)
    m32 = gmp_b200.TensorProductConvLayer(hid, hid, sh_ir, 8, mlp, gate=gate).cuda()
    with torch.no_grad():
        m32.fc[2].bias.normal_(0, 0.05)
    m16 = gmp_b200.TensorProductConvLayer(hid, hid, sh_ir, 8, mlp, gate=gate, precision="bf16").cuda()
    m16.load_state_dict(m32.state_dict())
    esh, eft = gmp_b200.edge_geometry(pos, ei, 2, gmp_b200.RadialEmbeddingBlock(2.0, 8, 5))
    x = torch.randn(pos.shape[0], 9 * C, device="cuda")
    x32, x16 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    o32, o16 = m32(x32, ei, esh, eft), m16(x16, ei, esh, eft)
    torch.cuda.synchronize()
    assert rel_err(o16, o32) <= 1e-2
    cot = torch.randn_like(o32)
    g32 = torch.autograd.grad((o32 * cot).sum(), [x32] + list(m32.parameters()))
    g16 = torch.autograd.grad((o16 * cot).sum(), [x16] + list(m16.parameters()))
    for a, b, name in zip(g16, g32, ["node_attr"] + [k for k, _ in m32.named_parameters()]):
        assert rel_err(a, b) <= 1e-2, name
    assert torch.equal(o16, m16(x16, ei, esh, eft))  # deterministic


def _l2_rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("act,aggr,n,side,fused", [("swish", "mean", 1000, 5.0, True), ("swish", "add", 3000, 7.0, True),
                                                   ("swish", "add", 3000, 7.0, False), ("relu", "add", 3000, 7.0, True)])
def test_egnn_bf16_tc_vs_fp32(act, aggr, n, side, fused, monkeypatch):
    """EGNN layer, tcgen05 edge kernels (forward; single-pass backward with per-edge scratch, or the two recompute
    passes) against the fp32-strict kernels.
    Smooth activation: every output and gradient within 1e-2 (normwise, max).  ReLU: the forward holds 1e-2; its
    gradients are compared in the L2 norm with a 0.1 bound, because a bf16-level perturbation of a pre-activation that
    sits at the kink flips that unit's derivative -- the fp32 reference shows the same sensitivity (mlp_upd, a pure
    fp32 torch path, moves by 2-3e-2 when its input moves by 2e-3)."""
    import gmp_b200
    if not fused:   # force the two-pass (recompute) backward that runs when the per-edge scratch would not fit
        monkeypatch.setattr(gmp_b200.egnn, "_FUSED_BWD_SCRATCH_BYTES", 0)
    g = torch.Generator().manual_seed(n)
    pos = (torch.rand(n, 3, generator=g) * side).cuda()
    ei = gmp_b200.radius_graph(pos, 1.0, None, max_num_neighbors=128)
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
    assert rel_err(o16, o32) <= 1e-2 and rel_err(q16 - pos, q32 - pos) <= 1e-2
    c1, c2 = torch.randn_like(o32), torch.randn_like(q32)
    g32 = torch.autograd.grad((o32 * c1).sum() + (q32 * c2).sum(), [h32, p32] + list(m32.parameters()))
    g16 = torch.autograd.grad((o16 * c1).sum() + (q16 * c2).sum(), [h16, p16] + list(m16.parameters()))
    names = ["h", "pos"] + [k for k, _ in m32.named_parameters()]
    for a, b, k in zip(g16, g32, names):
        if act == "swish":
            assert rel_err(a, b) <= 1e-2, k
        else:
            assert _l2_rel(a, b) <= 0.1, k
    o16b, q16b = m16(h16, p16, ei)
    assert torch.equal(o16, o16b) and torch.equal(q16, q16b)  # deterministic


@pytest.mark.parametrize("n,deg,shuffle", [(5000, 2, True), (300, 40, False), (70000, 9, True), (64, 0, False)])
def test_cfconv_pipelined_forward_edge_cases(n, deg, shuffle):
    """The pipelined three-MMA CFConv forward (csrc/schnet_tc2.cu) against the fp32 kernel on graphs that exercise its
    corners: low degree (more than 32 destination rows per 128 edges: tiles are cut), isolated nodes (empty rows),
    rows straddling tile and CTA boundaries (head buffer + fix-up; 70 000 x 9 edges = enough tiles for every SM),
    shuffled edge order (perm), high degree (one row spanning tiles), and the empty graph."""
    import gmp_b200
    g = torch.Generator().manual_seed(n + deg)
    E = n * deg
    dst = torch.randint(0, max(n - n // 7, 1), (E,), generator=g)      # the last n/7 nodes never receive an edge
    src = torch.randint(0, n, (E,), generator=g)
    if not shuffle:
        dst = dst.sort().values
    ei = torch.stack([src, dst]).cuda()
    torch.manual_seed(0)
    m32 = gmp_b200.InteractionBlock(128, 50, 128, 5.0).cuda()
    with torch.no_grad():
        for p in m32.parameters():
            if p.dim() == 1:
                p.normal_(0, 0.3)
    m16 = gmp_b200.InteractionBlock(128, 50, 128, 5.0, precision="bf16").cuda()
    m16.load_state_dict(m32.state_dict())
    sm = gmp_b200.GaussianSmearing(0.0, 5.0, 50).cuda()
    x = torch.randn(n, 128, device="cuda")
    ew = torch.rand(E, generator=g).cuda() * 5.0
    x32, x16 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    o32, o16 = m32(x32, ei, ew, sm.lazy()), m16(x16, ei, ew, sm.lazy())
    assert rel_err(o16, o32) <= 1e-2
    cot = torch.randn_like(o32)
    (g32,) = torch.autograd.grad((o32 * cot).sum(), [x32])
    (g16,) = torch.autograd.grad((o16 * cot).sum(), [x16])   # d/dx1 runs the same kernel over the transposed CSR
    assert rel_err(g16, g32) <= 1e-2
    assert torch.equal(o16, m16(x16, ei, ew, sm.lazy()))


@pytest.mark.timeout(180)
def test_cfconv_pipelined_one_tile_chunks_stress():
    """Regression for a rare hang of the pipelined CFConv forward on chunks that hold a single 128-edge tile: the
    end-of-stream arrival of the softplus warps could complete a second phase of the barrier the G2 warp was still
    waiting on (a waiter two phases behind never wakes).  12 000 edges = 94 one-tile chunks; both CSR directions,
    launched back to back many times through the C ABI."""
    import ctypes as C
    import gmp_b200
    from gmp_b200._lib import SchnetFilter, call, ptr
    n, deg = 300, 40
    g = torch.Generator().manual_seed(n + deg)
    E = n * deg
    dst = torch.randint(0, n - n // 7, (E,), generator=g).sort().values
    src = torch.randint(0, n, (E,), generator=g)
    gr = gmp_b200.get_graph(torch.stack([src, dst]).cuda(), n)
    torch.manual_seed(0)
    blk = gmp_b200.InteractionBlock(128, 50, 128, 5.0).cuda()
    sm = gmp_b200.GaussianSmearing(0.0, 5.0, 50).cuda()
    w1, b1, w2, b2 = blk.mlp[0].weight, blk.mlp[0].bias, blk.mlp[2].weight, blk.mlp[2].bias
    filt = SchnetFilter(ptr(w1), ptr(b1), ptr(w2), ptr(b2), 50, 128, 5.0, ptr(sm.offset), sm.coeff)
    ew = torch.rand(E, generator=g).cuda() * 5.0
    x1b = torch.randn(n, 128, device="cuda").to(torch.bfloat16)
    head = torch.empty(gmp_b200._lib.lib().gmp_schnet_tc2_num_chunks(E), 128, device="cuda")
    first = {}
    for it in range(1500):
        for name, csr in (("dst", gr.by_dst), ("src", gr.by_src)):
            agg = torch.empty(n, 128, device="cuda")
            call("gmp_schnet_cfconv_fwd_tc2", ptr(csr.rowptr), ptr(csr.col), csr.perm_ptr, ptr(csr.row_ids()), n, E, ptr(ew),
                 ptr(x1b), C.byref(filt), ptr(agg), ptr(head))
            if it % 250 == 0:
                torch.cuda.synchronize()
                assert torch.equal(agg, first.setdefault(name, agg))   # and bit-identical every time
    torch.cuda.synchronize()


@pytest.mark.parametrize("case", ["empty", "isolated", "shuffled_small_mul"])
def test_tp_conv_bf16_tc_corner_cases(case):
    """tcgen05 tensor-product convolution on degenerate inputs: no edges at all, nodes without edges (rows that must
    come out as the bias-free zero), multiplicities that do not fill an N-tile (12x0e+12x1o -> zero-padded columns)
    with a shuffled edge list."""
    import gmp_b200
    g = torch.Generator().manual_seed(7)
    n = 40
    if case == "empty":
        ei = torch.zeros(2, 0, dtype=torch.long)
        irr_in = irr_out = "16x0e+16x1o+16x2e"
    elif case == "isolated":
        src = torch.randint(0, 20, (150,), generator=g)
        dst = torch.randint(0, 20, (150,), generator=g)
        ei = torch.stack([src, dst])            # nodes 20..39 never appear
        irr_in = irr_out = "16x0e+16x1o+16x2e"
    else:
        src = torch.randint(0, n, (333,), generator=g)
        dst = torch.randint(0, n, (333,), generator=g)
        ei = torch.stack([src, dst])
        irr_in, irr_out = "12x0e+12x1o", "12x0e+12x1o+4x2e"
    ei = ei.cuda()
    E = ei.shape[1]
    sh_ir = "1x0e+1x1o+1x2e"
    torch.manual_seed(3)
    m32 = gmp_b200.TensorProductConvLayer(irr_in, irr_out, sh_ir, 8, 64).cuda()
    with torch.no_grad():
        m32.fc[2].bias.normal_(0, 0.1)
    m16 = gmp_b200.TensorProductConvLayer(irr_in, irr_out, sh_ir, 8, 64, precision="bf16").cuda()
    m16.load_state_dict(m32.state_dict())
    vec = torch.randn(E, 3, generator=g)
    esh = gmp_b200.SphericalHarmonics(2)(vec).cuda()
    eft = torch.rand(E, 8, generator=g).cuda()
    x = torch.randn(n, m32.in_irreps.dim, generator=g).cuda()
    x32, x16 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    o32, o16 = m32(x32, ei, esh, eft), m16(x16, ei, esh, eft)
    assert o16.shape == o32.shape
    if E == 0:
        assert float(o16.detach().abs().max()) == 0.0
    else:
        assert rel_err(o16, o32) <= 1e-2
    cot = torch.randn_like(o32)
    g32 = torch.autograd.grad((o32 * cot).sum(), [x32] + list(m32.parameters()), allow_unused=True)
    g16 = torch.autograd.grad((o16 * cot).sum(), [x16] + list(m16.parameters()), allow_unused=True)
    for a, b in zip(g16, g32):
        if b is None or float(b.abs().max()) == 0.0:
            assert a is None or float(a.abs().max()) == 0.0
        else:
            assert rel_err(a, b) <= 1e-2


def test_egnn_bf16_tc_empty_and_isolated():
    """tcgen05 EGNN kernels with no edges at all, and with nodes that receive no edge (rows must come out as zeros)."""
    import gmp_b200
    torch.manual_seed(0)
    m32 = gmp_b200.EGNNLayer(128, activation="swish").cuda()   # smooth activation: gradients comparable at 1e-2
    m16 = gmp_b200.EGNNLayer(128, activation="swish", precision="bf16").cuda()
    m16.load_state_dict(m32.state_dict())
    n = 50
    h, pos = torch.randn(n, 128, device="cuda"), torch.randn(n, 3, device="cuda")
    for ei in (torch.zeros(2, 0, dtype=torch.long, device="cuda"),
               torch.stack([torch.arange(0, 20), torch.arange(1, 21)]).cuda()):   # a chain on nodes 0..20, nodes 21..49 isolated
        h32, h16 = h.clone().requires_grad_(True), h.clone().requires_grad_(True)
        p32, p16 = pos.clone().requires_grad_(True), pos.clone().requires_grad_(True)
        o32, q32 = m32(h32, p32, ei)
        o16, q16 = m16(h16, p16, ei)
        assert rel_err(o16, o32) <= 1e-2 and rel_err(q16, q32) <= 1e-2
        g32 = torch.autograd.grad(o32.sum() + q32.sum(), [h32, p32])
        g16 = torch.autograd.grad(o16.sum() + q16.sum(), [h16, p16])
        for a, b in zip(g16, g32):
            assert rel_err(a, b) <= 1e-2


@pytest.mark.parametrize("n,deg,xbf16", [(700, 9, True), (300, 40, False), (50, 0, True)])
def test_gather_mul_segsum_wbf16(n, deg, xbf16):
    """K0 with a bf16 per-edge factor (dL/dx1 of the CFConv from kept filter values): against the same sum in torch on the
    bf16-rounded inputs -- the kernel only differs in the order of the fp32 additions."""
    import gmp_b200
    from gmp_b200._lib import call, ptr
    g = torch.Generator().manual_seed(n + deg)
    E = n * deg
    src, dst = torch.randint(0, n, (E,), generator=g), torch.randint(0, max(n - n // 5, 1), (E,), generator=g)
    gr = gmp_b200.get_graph(torch.stack([src, dst]).cuda(), n)
    csr = gr.by_src
    x = torch.randn(n, 128, generator=g).cuda()
    w = torch.randn(E, 128, generator=g).cuda().to(torch.bfloat16)
    xin = x.to(torch.bfloat16) if xbf16 else x
    out = torch.full((n, 128), float("nan"), device="cuda")
    call("gmp_gather_mul_segsum_wbf16", ptr(csr.rowptr), ptr(csr.col), csr.perm_ptr, ptr(xin), int(xbf16), ptr(w), ptr(out), n, 128)
    ref = torch.zeros(n, 128, device="cuda", dtype=torch.float64)
    ref.index_add_(0, src.cuda(), (xin.double()[dst.cuda()] * w.double()))
    assert torch.isfinite(out).all()
    assert (out.double() - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item())


def test_cfconv_kept_filter_values():
    """The training variant of the pipelined forward stores W(e) C(e) per edge (caller's edge order): against the filter MLP in
    torch (fp32), 1e-2; and the aggregate is bit-identical to the plain variant's."""
    import ctypes as C
    import gmp_b200
    from gmp_b200._lib import SchnetFilter, call, ptr
    n, E = 900, 11000
    g = torch.Generator().manual_seed(5)
    ei = torch.stack([torch.randint(0, n, (E,), generator=g), torch.randint(0, n, (E,), generator=g)]).cuda()
    gr = gmp_b200.get_graph(ei, n)
    csr = gr.by_dst
    torch.manual_seed(0)
    blk = gmp_b200.InteractionBlock(128, 50, 128, 5.0).cuda()
    sm = gmp_b200.GaussianSmearing(0.0, 5.0, 50).cuda()
    w1, b1, w2, b2 = blk.mlp[0].weight, blk.mlp[0].bias, blk.mlp[2].weight, blk.mlp[2].bias
    filt = SchnetFilter(ptr(w1), ptr(b1), ptr(w2), ptr(b2), 50, 128, 5.0, ptr(sm.offset), sm.coeff)
    ew = torch.rand(E, generator=g).cuda() * 5.5      # some edges beyond the cutoff of the cosine window? no: C(d) is smooth, > 5 wraps
    ew = ew.clamp(max=5.0)
    x1b = torch.randn(n, 128, device="cuda").to(torch.bfloat16)
    head = torch.empty(gmp_b200._lib.lib().gmp_schnet_tc2_num_chunks(E), 128, device="cuda")
    agg0, agg1 = torch.empty(n, 128, device="cuda"), torch.empty(n, 128, device="cuda")
    keep = torch.full((E, 128), float("nan"), device="cuda", dtype=torch.bfloat16)
    call("gmp_schnet_cfconv_fwd_tc2", ptr(csr.rowptr), ptr(csr.col), csr.perm_ptr, ptr(csr.row_ids()), n, E, ptr(ew), ptr(x1b),
         C.byref(filt), ptr(agg0), ptr(head))
    call("gmp_schnet_cfconv_fwd_tc2_keep", ptr(csr.rowptr), ptr(csr.col), csr.perm_ptr, ptr(csr.row_ids()), n, E, ptr(ew), ptr(x1b),
         C.byref(filt), ptr(agg1), ptr(head), ptr(keep), None)
    assert torch.equal(agg0, agg1)
    # the same values at caller-chosen rows (the model passes the position of each edge in the source-sorted CSR)
    inv = gr.by_src.inv_perm()
    keep2 = torch.full((E, 128), float("nan"), device="cuda", dtype=torch.bfloat16)
    call("gmp_schnet_cfconv_fwd_tc2_keep", ptr(csr.rowptr), ptr(csr.col), csr.perm_ptr, ptr(csr.row_ids()), n, E, ptr(ew), ptr(x1b),
         C.byref(filt), ptr(agg1), ptr(head), ptr(keep2), ptr(inv))
    assert torch.equal(agg0, agg1) and torch.equal(keep2[inv.long()], keep)
    with torch.no_grad():
        rbf = torch.exp(sm.coeff * (ew[:, None] - sm.offset[None, :]) ** 2)
        h = torch.nn.functional.softplus(rbf @ w1.t() + b1) - 0.6931471805599453
        ref = (h @ w2.t() + b2) * (0.5 * (torch.cos(ew * 3.141592653589793 / 5.0) + 1.0))[:, None]
    assert torch.isfinite(keep.float()).all()
    assert rel_err(keep.float(), ref) <= 1e-2
