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
    """On the large-batch CUDA training path, feature counts that are not a multiple of 4 floats (47 / 94 inputs, the
    5-way head) are zero-padded to one: with 16-byte-aligned rows cuBLAS runs its sm_100 TMA kernels instead of the
    ``align1`` sm_80 ones for the forward, the input gradient and the weight gradient (2.0 -> ~1.0 ms per QMIX update).
    The padding goes through ``F.pad`` / slicing, so autograd hands back gradients of the parameters' own shapes.  An
    input that already carries the padding columns (``in_features`` rounded up to 4; they must be zero) is used as is."""

    def forward(self, x):
        if x.is_cuda and x.dtype == torch.float32 and torch.is_grad_enabled() and x.numel() // x.shape[-1] >= 4096:
            K, N = self.in_features, self.out_features
            pk, pn = (-K) % 4, (-N) % 4
            w, b = self.weight, self.bias
            if pk or pn:
                if x.shape[-1] == K:
                    x = F.pad(x, (0, pk)) if pk else x
                w = F.pad(w, (0, pk, 0, pn))
                b = F.pad(b, (0, pn)) if (b is not None and pn) else b
            y = _LinearFn.apply(x, w, b)
            return y[..., :N] if pn else y
        if x.shape[-1] != self.in_features:                      # pre-padded input on a path that does not pad
            x = x[..., :self.in_features]
        return super().forward(x)
