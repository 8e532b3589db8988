"""CPU: how far the ORACLE's own gradients move when its inputs are rounded to bf16 (relative 2^-9).

This is the reference-side half of the bf16 tolerance statement ("EGNN ReLU"): with SiLU the EGNN layer
(models/layers/egnn_layer.py:62-86) is 1e-2-stable under such a change in every gradient; with ReLU (the reference
default) it is not -- a LayerNorm output within 2^-9 of zero flips that unit's derivative, and the fp32 reference's
gradients move by a few 1e-2 in the L2 norm.  tests/test_gpu_tc.py::test_egnn_bf16_tc_vs_oracle measures this same
quantity on its own inputs and bounds the tcgen05 kernels by a small multiple of it."""
import torch

from oracle import ref_layers as R
from oracle.thirdparty import cluster


def _l2(a, b):
    return ((a - b).double().norm() / b.double().norm().clamp_min(1e-30)).item()


def _grads(layer, h, pos, ei, c1, c2):
    hh, pp = h.clone().requires_grad_(True), pos.clone().requires_grad_(True)
    o, q = layer(hh, pp, ei)
    return o.detach(), torch.autograd.grad((o * c1).sum() + (q * c2).sum(), [hh, pp] + list(layer.parameters()))


def test_oracle_relu_gradients_are_not_bf16_stable_but_silu_are():
    n, side = 1200, 5.0
    g = torch.Generator().manual_seed(n)
    pos = torch.rand(n, 3, generator=g) * side
    ei = torch.from_numpy(cluster.radius_graph(pos.numpy(), 1.0, None, False, 128))
    assert int(ei[1].max()) == n - 1
    h = torch.randn(n, 128, generator=g)
    c1, c2 = torch.randn(n, 128, generator=g), torch.randn(n, 3, generator=g)
    moved = {}
    for act in ("relu", "swish"):
        torch.manual_seed(1)
        layer = R.EGNNLayer(128, act, "layer", "add")
        o0, g0 = _grads(layer, h, pos, ei, c1, c2)
        o1, g1 = _grads(layer, h.bfloat16().float(), pos, ei, c1, c2)
        assert _l2(o1, o0) <= 5e-3                       # the forward is stable for both activations
        moved[act] = max(_l2(a, b) for a, b in zip(g1[:1] + g1[2:6], g0[:1] + g0[2:6]))   # dL/dh and mlp_msg.0/1 grads
    assert moved["swish"] <= 5e-3, moved
    assert moved["relu"] >= 1e-2, moved                  # ReLU: the reference itself moves by more than the bf16 tolerance
    assert moved["relu"] >= 5 * moved["swish"], moved
