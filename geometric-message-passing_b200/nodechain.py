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
           res: Optional[torch.Tensor] = None, want_bf16: bool = False):
    """act(x W^T + b) (+ res) as one launch; returns (fp32 result, bf16 copy or None).  No autograd."""
    out = torch.empty(x.shape[0], 128, dtype=torch.float32, device=x.device)
    o16 = torch.empty(x.shape[0], 128, dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    img = pack_w(weight)
    x2 = None
    if weight.shape[1] == 256:
        x, x2 = x[:, :128].contiguous(), x[:, 128:].contiguous()
    run(x.contiguous(), [stage(img, bias, act=act, add_res=res, out_f32=out, out_bf16=o16)], x2)
    return out, o16
