"""Node-side dense chains (csrc/node_chain.cu): nn.Linear layers of width 128 with their elementwise neighbours fused into
tcgen05 epilogues -- the node half of the SchNet interaction (PyG blocks built at models/schnet.py:41-54), EGNN's ``mlp_upd``
and P / Q projections (models/layers/egnn_layer.py:41-48, 62-72, 82-86).  bf16 operands, fp32 accumulation: the GMP_BF16_TC
precision mode (1e-2); the fp32-strict mode keeps ``torch.nn.functional.linear``."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import NodeStage, call, ptr

ACT = {None: 0, "none": 0, "ssp": 1, "relu": 2, "silu": 3, "swish": 3}
MUL_PLAIN, MUL_DSSP = 0, 1


def pack_w(weight: torch.Tensor, transpose: bool = False) -> torch.Tensor:
    """bf16 operand image(s) of an nn.Linear weight [out, in]; ``transpose`` gives the image of W^T (for dx = g W)."""
    out_dim, in_dim = weight.shape
    nbytes = _lib.lib().gmp_node_w_image_bytes(out_dim, in_dim, int(transpose))
    if nbytes <= 0:
        raise _lib.GmpError(f"node chain: weight {tuple(weight.shape)} (transpose={transpose}) is not a 128-row operand")
    img = torch.empty(nbytes, dtype=torch.uint8, device=weight.device)
    call("gmp_node_pack_w", ptr(weight.detach().contiguous()), out_dim, in_dim, int(transpose), ptr(img))
    return img


def pack_w_batch(weights: Sequence[torch.Tensor], transposes: Sequence[bool]):
    """Operand images of many [128, 128] weights in ONE launch; returns the images (views of one buffer) in order."""
    n = len(weights)
    assert n == len(transposes) and all(tuple(w.shape) == (128, 128) for w in weights)
    ws = [w.detach().contiguous() for w in weights]
    buf = torch.empty(n, 32768, dtype=torch.uint8, device=ws[0].device)
    call("gmp_node_pack_w_batch", (C.c_void_p * n)(*[w.data_ptr() for w in ws]), (C.c_int32 * n)(*[int(t) for t in transposes]),
         (C.c_void_p * n)(*[buf[i].data_ptr() for i in range(n)]), n)
    return [buf[i] for i in range(n)]


def stage(w_img: torch.Tensor, bias: Optional[torch.Tensor] = None, ln: Optional[Sequence] = None, act: Optional[str] = None,
          mul_aux: Optional[torch.Tensor] = None, mul_mode: int = MUL_PLAIN, add_res: Optional[torch.Tensor] = None,
          out_f32: Optional[torch.Tensor] = None, out_bf16: Optional[torch.Tensor] = None, out_pre: Optional[torch.Tensor] = None):
    """One stage description; ``ln`` = (weight, bias, eps).  The returned struct keeps its tensors alive (``_keep``)."""
    g, b, eps = (ln[0], ln[1], float(ln[2])) if ln is not None else (None, None, 0.0)
    for t_ in (out_f32, out_pre, mul_aux, add_res):
        assert t_ is None or (t_.dtype == torch.float32 and t_.shape[-1] == 128)
    assert out_bf16 is None or out_bf16.dtype == torch.bfloat16
    st = NodeStage(ptr(w_img), ptr(bias), ptr(g), ptr(b), eps, ACT[act], ptr(mul_aux), mul_mode, ptr(add_res), ptr(out_f32),
                   ptr(out_bf16), ptr(out_pre))
    st._keep = (w_img, bias, g, b, mul_aux, add_res, out_f32, out_bf16, out_pre)   # the struct holds raw pointers only
    return st


def run(a0: torch.Tensor, stages: Sequence[NodeStage], a1: Optional[torch.Tensor] = None) -> None:
    assert a0.dtype == torch.float32 and a0.dim() == 2 and a0.shape[1] == 128 and (a1 is None or a1.shape == a0.shape)
    arr = (NodeStage * len(stages))(*stages)
    call("gmp_node_chain_tc", ptr(a0), ptr(a1), a0.shape[0], len(stages), C.byref(arr))


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, act: Optional[str] = None,
           res: Optional[torch.Tensor] = None, want_bf16: bool = False, out16: Optional[torch.Tensor] = None):
    """act(x W^T + b) (+ res) as one launch; returns (fp32 result, bf16 copy or None).  No autograd.
    out16: caller-provided bf16 [n,128] destination of the copy (e.g. a peer-visible symmetric-memory buffer)."""
    out = torch.empty(x.shape[0], 128, dtype=torch.float32, device=x.device)
    o16 = out16 if out16 is not None else (torch.empty(x.shape[0], 128, dtype=torch.bfloat16, device=x.device) if want_bf16 else None)
    assert o16 is None or (o16.dtype == torch.bfloat16 and tuple(o16.shape) == (x.shape[0], 128) and o16.is_contiguous())
    img = pack_w(weight)
    x2 = None
    if weight.shape[1] == 256:
        x, x2 = x[:, :128].contiguous(), x[:, 128:].contiguous()
    run(x.contiguous(), [stage(img, bias, act=act, add_res=res, out_f32=out, out_bf16=o16)], x2)
    return out, o16


def wgrad(g: torch.Tensor, x: torch.Tensor):
    """(dW [128,128], db [128]) = (g^T x, column sums of g) over all rows: tensor-core reduction kernel + fixed-order sum."""
    n = g.shape[0]
    nparts, plen = _lib.lib().gmp_linear_wgrad_num_parts(n), 128 * 128 + 128
    parts = torch.empty(nparts, plen, dtype=torch.float32, device=g.device)
    call("gmp_linear_wgrad_tc", ptr(g), ptr(x), n, 128, 128, ptr(parts))
    red = torch.empty(plen, dtype=torch.float32, device=g.device)
    call("gmp_reduce_partials_f32", ptr(parts), nparts, plen, ptr(red))
    return red[:128 * 128].view(128, 128), red[128 * 128:]


class ChainLinearFn(torch.autograd.Function):
    """y = x W^T (+ b) for a 128 x 128 weight on the chain kernel, optionally with a bf16 copy of y as a second
    (non-differentiable) output; dx on the chain kernel (transposed image), dW / db on the reduction kernel."""

    @staticmethod
    def forward(ctx, x, w, b, want_bf16: bool, out16=None):
        x = x.contiguous()
        y, y16 = linear(x, w, b, want_bf16=want_bf16, out16=out16)
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        if y16 is None or out16 is not None:     # (a caller-provided buffer is not handed back: the caller holds it)
            y16 = x.new_empty(0, dtype=torch.bfloat16)
        ctx.mark_non_differentiable(y16)
        return y, y16

    @staticmethod
    def backward(ctx, g, _g16):
        x, w = ctx.saved_tensors
        g = g.contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            run(g, [stage(pack_w(w, True), out_f32=dx)])
        dw = db = None
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw, db = wgrad(g, x)
        return dx, dw, (db if ctx.has_bias else None), None, None


def ln_act_bwd(g, pre, gamma, beta, eps, act: str, want_act: bool):
    """(d_pre, recomputed activations or None, d gamma, d beta) of a = act(LayerNorm(pre) * gamma + beta)."""
    n = pre.shape[0]
    d_pre = torch.empty_like(pre)
    a = torch.empty_like(pre) if want_act else None
    nparts = _lib.lib().gmp_ln_act_bwd_num_parts(n)
    parts = torch.empty(nparts, 256, dtype=torch.float32, device=pre.device)
    call("gmp_ln_act_bwd", ptr(g), ptr(pre), ptr(gamma), ptr(beta), float(eps), ACT[act], n, ptr(d_pre), ptr(a), ptr(parts))
    red = torch.empty(256, dtype=torch.float32, device=pre.device)
    call("gmp_reduce_partials_f32", ptr(parts), nparts, 256, ptr(red))
    return d_pre, a, red[:128], red[128:]


class EGNNUpdateFn(torch.autograd.Function):
    """``mlp_upd(cat[h, msg_aggr])`` of models/layers/egnn_layer.py:41-48, 82-86 (Linear(2d, d), LayerNorm, act, Linear(d, d),
    LayerNorm, act; d = 128) as one chain launch; the backward pass is two LayerNorm/activation kernels, three chain
    launches for the data gradients and three weight-gradient reductions.  Kept for the backward: the two pre-LayerNorm
    tensors (the activations are recomputed from them)."""

    @staticmethod
    def forward(ctx, h, agg, w0, b0, g0, be0, w1, b1, g1, be1, act: str, eps: float):
        h, agg = h.contiguous(), agg.contiguous()
        n = h.shape[0]
        train = any(ctx.needs_input_grad)
        pre0 = torch.empty_like(h) if train else None
        pre1 = torch.empty_like(h) if train else None
        out = torch.empty_like(h)
        run(h, [stage(pack_w(w0), b0, ln=(g0, be0, eps), act=act, out_pre=pre0),
                stage(pack_w(w1), b1, ln=(g1, be1, eps), act=act, out_f32=out, out_pre=pre1)], a1=agg)
        ctx.save_for_backward(h, agg, w0, g0, be0, w1, g1, be1, pre0, pre1)
        ctx.act, ctx.eps = act, eps
        return out

    @staticmethod
    def backward(ctx, g):
        h, agg, w0, g0, be0, w1, g1, be1, pre0, pre1 = ctx.saved_tensors
        act, eps = ctx.act, ctx.eps
        g = g.contiguous()
        dpre1, _, dg1, dbe1 = ln_act_bwd(g, pre1, g1, be1, eps, act, False)
        da0 = torch.empty_like(h)
        run(dpre1, [stage(pack_w(w1, True), out_f32=da0)])
        dpre0, a0, dg0, dbe0 = ln_act_bwd(da0, pre0, g0, be0, eps, act, True)
        dw1, db1 = wgrad(dpre1, a0)
        dh, dagg = torch.empty_like(h), torch.empty_like(h)
        w0c = w0.detach()
        run(dpre0, [stage(pack_w(w0c[:, :128].contiguous(), True), out_f32=dh)])
        run(dpre0, [stage(pack_w(w0c[:, 128:].contiguous(), True), out_f32=dagg)])
        dwa, db0 = wgrad(dpre0, h)
        dwb, _ = wgrad(dpre0, agg)
        return dh, dagg, torch.cat([dwa, dwb], dim=1), db0, dg0, dbe0, dw1, db1, dg1, dbe1, None, None
