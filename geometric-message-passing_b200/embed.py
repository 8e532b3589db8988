"""Atom-type embedding lookup whose backward asks the device nothing.

ATen's `embedding_dense_backward` sorts the indices and reads the number of distinct ones back to the host: one stream
synchronisation per training step (which also makes the step impossible to capture in a CUDA graph), and it accumulates
with atomics.  The vocabulary here is tiny (atom types: <= 100 rows), so the weight gradient is simply
one_hot(idx)^T @ grad -- one library GEMM, deterministic, no read-back.  Forward and values are those of
`torch.nn.Embedding` (including `padding_idx`: that row receives no gradient), which stays the parameter container so
that state_dict keys match the reference (`models/schnet.py` via PyG `SchNet.embedding`, `models/egnn.py:30` `emb_in`)."""
from __future__ import annotations

import torch


class _EmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, idx, padding_idx):
        ctx.save_for_backward(idx)
        ctx.rows, ctx.padding_idx = weight.shape[0], padding_idx
        return weight.index_select(0, idx)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        onehot = torch.zeros(idx.numel(), ctx.rows, dtype=g.dtype, device=g.device)
        onehot.scatter_(1, idx.view(-1, 1), 1.0)
        gw = onehot.t().mm(g.reshape(idx.numel(), -1))
        if ctx.padding_idx is not None:
            gw[ctx.padding_idx].zero_()
        return gw, None, None


def embedding_lookup(emb: torch.nn.Embedding, idx: torch.Tensor) -> torch.Tensor:
    if not idx.is_cuda or idx.dim() != 1 or emb.max_norm is not None or emb.sparse or emb.scale_grad_by_freq:
        return emb(idx)
    return _EmbedFn.apply(emb.weight, idx, emb.padding_idx)
