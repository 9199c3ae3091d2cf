"""``nn.Linear`` whose weight gradient is the slab-split product of ``ops.gram`` (same parameters, same state_dict keys).

Only the backward differs from ``torch.nn.Linear``: dW = dY^T X over rows = batch x time is computed as a batched
split-K product, which reads the two tall operands once at HBM speed instead of through one under-parallelised GEMM."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..ops.gram import colsum, gram


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return F.linear(x, w, b)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        g2 = gy.reshape(-1, gy.shape[-1])
        gx = (g2 @ w).view(x.shape) if ctx.needs_input_grad[0] else None
        gw = gram(g2.contiguous(), x.reshape(-1, x.shape[-1]).contiguous()) if ctx.needs_input_grad[1] else None
        gb = colsum(g2) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return gx, gw, gb


class Linear(nn.Linear):
    def forward(self, x):
        if x.is_cuda and x.dtype == torch.float32 and torch.is_grad_enabled() and x.numel() // x.shape[-1] >= 4096:
            return _LinearFn.apply(x, self.weight, self.bias)
        return super().forward(x)
