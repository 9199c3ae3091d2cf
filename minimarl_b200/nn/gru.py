"""Single-layer GRU whose sequential part runs in the tcgen05 kernels of csrc/gru_tc.cu, as an autograd op.

``fused_gru(x, h0, w_ih, w_hh, b_ih, b_hh)`` has nn.GRU's semantics (gate order r, z, n; ``x`` [T, R, I],
``h0`` [R, H]); the row-parallel GEMMs (input projection, dx, dW_ih, dW_hh) stay in cuBLAS, the recurrence and
its BPTT are one kernel launch each.  ``GRU`` is an nn.GRU subclass (same parameters / state_dict keys) that
uses it for CUDA float32 inputs with hidden size 32 or 64.
"""
import torch
import torch.nn as nn

from .. import _lib
from ..ops.gram import colsum, gram


class _FusedGRUFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, h0, w_ih, w_hh, b_ih, b_hh):
        T, R, I = x.shape
        H = w_hh.shape[1]
        x2 = x.reshape(T * R, I)
        gi = torch.addmm(b_ih, x2, w_ih.t())                                   # [T*R, 3H]  (cuBLAS)
        z_ = dict(dtype=torch.float32, device=x.device)
        h_seq = torch.empty((T, R, H), **z_)
        need_grad = any(ctx.needs_input_grad)
        saved = [torch.empty((T, R, H), **z_) for _ in range(4)] if need_grad else [None] * 4
        h0c = h0.contiguous() if h0 is not None else None
        _lib.check(_lib.lib().mm_gru_seq_fwd(_lib.ptr(gi), _lib.ptr(h0c), _lib.ptr(w_hh.contiguous()), _lib.ptr(b_hh.contiguous()),
                                             _lib.ptr(h_seq), *[_lib.ptr(s) for s in saved], T, R, H, _lib.stream()))
        if need_grad:
            ctx.save_for_backward(x2, h0c if h0c is not None else torch.empty(0, device=x.device), w_ih, w_hh, h_seq, *saved)
            ctx.dims = (T, R, I, H, h0 is not None)
        return h_seq

    @staticmethod
    def backward(ctx, dh_seq):
        x2, h0, w_ih, w_hh, h_seq, r, z, n, ghn = ctx.saved_tensors
        T, R, I, H, has_h0 = ctx.dims
        z_ = dict(dtype=torch.float32, device=x2.device)
        dgi = torch.empty((T, R, 3 * H), **z_)
        dghn = torch.empty((T, R, H), **z_)
        dh0 = torch.empty((R, H), **z_)
        _lib.check(_lib.lib().mm_gru_seq_bwd(_lib.ptr(dh_seq.contiguous()), _lib.ptr(r), _lib.ptr(z), _lib.ptr(n), _lib.ptr(ghn),
                                             _lib.ptr(h_seq), _lib.ptr(h0 if has_h0 else None), _lib.ptr(w_hh.contiguous()),
                                             _lib.ptr(dgi), _lib.ptr(dghn), _lib.ptr(dh0), T, R, H, _lib.stream()))
        dgi2 = dgi.view(T * R, 3 * H)
        dx = (dgi2 @ w_ih).view(T, R, I) if ctx.needs_input_grad[0] else None
        dw_ih = gram(dgi2, x2.reshape(T * R, I).contiguous())
        db_ih = colsum(dgi2)
        # h_{t-1} for every step: h0 (or zeros) followed by h_seq[:-1]
        h_prev = torch.empty((T, R, H), **z_)
        if has_h0:
            h_prev[0].copy_(h0)
        else:
            h_prev[0].zero_()
        h_prev[1:].copy_(h_seq[:-1])
        hp2 = h_prev.view(T * R, H)
        dghn2 = dghn.view(T * R, H)
        dw_hh_all = gram(dgi2, hp2)                      # the n rows of this product are discarded: one pass over hp2
        dw_hh = torch.cat((dw_hh_all[:2 * H], gram(dghn2, hp2)), dim=0)
        db_hh = torch.cat((db_ih[:2 * H], colsum(dghn2)))
        return dx, (dh0 if (has_h0 and ctx.needs_input_grad[1]) else None), dw_ih, dw_hh, db_ih, db_hh


def fused_gru(x, h0, w_ih, w_hh, b_ih, b_hh):
    """Returns (output [T, R, H], h_n [R, H])."""
    out = _FusedGRUFn.apply(x.contiguous(), h0, w_ih, w_hh, b_ih, b_hh)
    return out, out[-1]


class GRU(nn.GRU):
    """nn.GRU with the fused recurrence on CUDA (one layer, unidirectional, hidden 32 / 64, float32)."""

    def forward(self, x, hx=None):
        # TF32 operands, like cuDNN's own GRU when torch.backends.cudnn.allow_tf32 is on (the default); switching that
        # flag off selects PyTorch's fp32 path here as well
        if (x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and self.num_layers == 1 and not self.bidirectional
                and not self.batch_first and self.bias and self.hidden_size in (32, 64)
                and torch.backends.cudnn.allow_tf32):
            h0 = hx[0] if hx is not None else None
            out, hn = fused_gru(x, h0, self.weight_ih_l0, self.weight_hh_l0, self.bias_ih_l0, self.bias_hh_l0)
            return out, hn.unsqueeze(0)
        return super().forward(x, hx)
