"""Row-parallel half of the agent network's differentiated pass as ONE op (``mm_agent_mlp_fwd`` / ``_bwd``, tcgen05, fp16
operands with fp32 accumulation): ``LayerNorm -> [Linear ReLU LayerNorm] x2 -> GRU input projection`` for hidden size 64.

What ``MLPBase`` / ``MLPLayer`` (offpolicy/utils/algorithm_utils/mlp.py:7-89) and the input half of ``nn.GRU`` in ``RNNLayer``
(rnn.py:4-47) compute when ``QMix.train_policy_on_batch`` (offpolicy/algorithms/qmix/qmix.py:80-210) differentiates the live
network.  The op returns ``gi`` [rows, 192], which ``gru_from_gi`` feeds to the recurrence kernels of csrc/gru_tc.cu; the
backward takes ``dgi`` and contracts every weight gradient inside the kernel (no cuBLAS, no LayerNorm / ReLU / bias launches).
LayerNorm affines are folded into the Linear that follows for the kernel (W' = W diag(g), b' = b + W beta) and unfolded here.
"""
import ctypes as C

import torch

from .. import _lib
from ..ops.gram import colsum, gram
from ._direct import grad_targets

H, K1 = 64, 48
_DIMS = None
_SCALES = {}          # (data_ptr, numel) of a dgi tensor -> its power-of-two gradient scale, from the GRU backward to agent_mlp's


def dims():
    global _DIMS
    if _DIMS is None:
        d = (C.c_int32 * 8)()
        _lib.check(_lib.lib().mm_agent_mlp_dims(d))
        _DIMS = dict(zip("n_params saved_row_floats acc_cols tile_rows K1 H x0_halves_per_tile _".split(), list(d)))
    return _DIMS


def agent_mlp_fits(in_dim, hidden, layer_N=1):
    return hidden == H and 1 <= in_dim <= K1 and layer_N == 1


_GRAD_LAYOUT = {}


def _grad_layout(D):
    if D not in _GRAD_LAYOUT:
        o = (C.c_int32 * 13)()
        _lib.check(_lib.lib().mm_agent_mlp_grad_layout(D, o))
        _GRAD_LAYOUT[D] = list(o)
    return _GRAD_LAYOUT[D]


def _ptr_array(params):
    arr = (C.c_void_p * len(params))()
    for i, t in enumerate(params):
        assert t.is_contiguous() and t.dtype == torch.float32
        arr[i] = t.data_ptr()
    return arr


class _AgentMlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, obs, *params):                 # params: ln0_w, ln0_b, w1, b1, ln1_w, ln1_b, w2, b2, ln2_w, ln2_b, wih, bih
        rows, D = obs.shape
        d = dims()
        param_objs = list(params)                      # the Parameters themselves (their .grad may take the gradients directly)
        params = [t.detach().contiguous() for t in params]
        packed = torch.empty(d["n_params"], dtype=torch.float32, device=obs.device)
        _lib.check(_lib.lib().mm_agent_mlp_pack(_ptr_array(params), _lib.ptr(packed), D, _lib.stream()))
        gi = torch.empty((rows, 3 * H), dtype=torch.float32, device=obs.device)
        train = any(ctx.needs_input_grad[1:])
        saved_rows = saved_x0 = None
        if train:
            n_tiles = (rows + d["tile_rows"] - 1) // d["tile_rows"]
            saved_rows = torch.empty((rows, d["saved_row_floats"]), dtype=torch.float32, device=obs.device)
            saved_x0 = torch.empty((n_tiles, d["x0_halves_per_tile"]), dtype=torch.float16, device=obs.device)
        _lib.check(_lib.lib().mm_agent_mlp_fwd(_lib.ptr(packed), _lib.ptr(obs), _lib.ptr(gi), _lib.ptr(saved_rows),
                                               _lib.ptr(saved_x0), rows, D, _lib.stream()))
        if train:
            ctx.save_for_backward(packed, saved_rows, saved_x0, *params)
            ctx.dims = (rows, D)
            ctx.param_objs = param_objs
        return gi

    @staticmethod
    def backward(ctx, dgi):
        packed, saved_rows, saved_x0 = ctx.saved_tensors[:3]
        params = list(ctx.saved_tensors[3:])
        rows, D = ctx.dims
        d = dims()
        dgi = dgi.contiguous()
        lib = _lib.lib()
        scale = _SCALES.pop((dgi.data_ptr(), dgi.numel()), None)          # left by the GRU backward that produced this dgi
        if scale is None:
            scale = torch.empty(2, dtype=torch.float32, device=dgi.device)
            _lib.check(lib.mm_ppo_trunk2_scale(_lib.ptr(dgi), dgi.numel(), _lib.ptr(scale), _lib.stream()))
        n_ctas = lib.mm_agent_mlp_bwd_ctas(rows)
        acc = torch.empty((n_ctas, d["tile_rows"], d["acc_cols"]), dtype=torch.float32, device=dgi.device)
        _lib.check(lib.mm_agent_mlp_bwd(_lib.ptr(packed), _lib.ptr(saved_rows), _lib.ptr(saved_x0), _lib.ptr(dgi), _lib.ptr(scale),
                                        _lib.ptr(acc), n_ctas, rows, _lib.stream()))
        # one launch: sum over the CTAs, 1 / scale, LayerNorm affines unfolded into the gradients of the module's own tensors
        off = _grad_layout(D)
        red = torch.empty(d["tile_rows"] * d["acc_cols"], dtype=torch.float32, device=dgi.device)
        targets = grad_targets(ctx.param_objs, ctx.needs_input_grad[1:])
        if targets is not None:
            gptrs = (C.c_void_p * len(targets))(*[g.data_ptr() for g in targets])
            _lib.check(lib.mm_agent_mlp_unfold_into(_lib.ptr(acc), n_ctas, _lib.ptr(scale), _lib.ptr(red), _ptr_array(params), gptrs, D,
                                                    _lib.stream()))
            return (None,) * (1 + len(params))
        flat = torch.empty(off[-1], dtype=torch.float32, device=dgi.device)
        _lib.check(lib.mm_agent_mlp_unfold(_lib.ptr(acc), n_ctas, _lib.ptr(scale), _lib.ptr(red), _ptr_array(params), _lib.ptr(flat),
                                           D, _lib.stream()))
        return (None,) + tuple(flat[off[i]: off[i + 1]].view_as(params[i]) if ctx.needs_input_grad[1 + i] else None
                               for i in range(len(params)))


def agent_mlp_gi(obs2d, ln0, fc1, ln1, fc2, ln2, w_ih, b_ih):
    """obs2d [rows, D]; ln*: (weight, bias) of the LayerNorms, fc*: (weight, bias) of the Linears -> gi [rows, 192]."""
    return _AgentMlpFn.apply(obs2d.contiguous(), ln0[0], ln0[1], fc1[0], fc1[1], ln1[0], ln1[1], fc2[0], fc2[1], ln2[0], ln2[1],
                             w_ih, b_ih)


class _GruFromGiFn(torch.autograd.Function):
    """The recurrence of ``nn.GRU`` on precomputed input projections (``mm_gru_seq_fwd`` / ``_bwd``): the op of nn/gru.py without
    its row-parallel input half, which ``agent_mlp_gi`` has fused into the layers below."""

    @staticmethod
    def forward(ctx, gi, h0, w_hh, b_hh, T, R):
        Hh = w_hh.shape[1]
        z_ = dict(dtype=torch.float32, device=gi.device)
        h_seq = torch.empty((T, R, Hh), **z_)
        need = any(ctx.needs_input_grad)
        saved = [torch.empty((T, R, Hh), **z_) for _ in range(4)] if need else [None] * 4
        h0c = h0.contiguous() if h0 is not None else None
        _lib.check(_lib.lib().mm_gru_seq_fwd(_lib.ptr(gi), _lib.ptr(h0c), _lib.ptr(w_hh.contiguous()), _lib.ptr(b_hh.contiguous()),
                                             _lib.ptr(h_seq), *[_lib.ptr(s) for s in saved], T, R, Hh, _lib.stream()))
        if need:
            ctx.save_for_backward(h0c if h0c is not None else torch.empty(0, device=gi.device), w_hh, h_seq, *saved)
            ctx.dims = (T, R, Hh, h0 is not None)
            ctx.param_objs = [w_hh, b_hh]
        return h_seq

    @staticmethod
    def backward(ctx, dh_seq):
        h0, w_hh, h_seq, r, z, n, ghn = ctx.saved_tensors
        T, R, Hh, has_h0 = ctx.dims
        z_ = dict(dtype=torch.float32, device=h_seq.device)
        dgi = torch.empty((T * R, 3 * Hh), **z_)
        dghn = torch.empty((T, R, Hh), **z_)
        dh0 = torch.empty((R, Hh), **z_)
        scale = torch.empty(2, **z_)                     # [0]: power-of-two scale of dgi for the fp16 consumers, [1]: max |dgi|
        _lib.check(_lib.lib().mm_gru_seq_bwd_scaled(_lib.ptr(dh_seq.contiguous()), _lib.ptr(r), _lib.ptr(z), _lib.ptr(n), _lib.ptr(ghn),
                                                    _lib.ptr(h_seq), _lib.ptr(h0 if has_h0 else None), _lib.ptr(w_hh.contiguous()),
                                                    _lib.ptr(dgi), _lib.ptr(dghn), _lib.ptr(dh0), _lib.ptr(scale), T, R, Hh,
                                                    _lib.stream()))
        _SCALES[(dgi.data_ptr(), dgi.numel())] = scale  # agent_mlp's backward receives this dgi next
        lib = _lib.lib()
        if Hh == H:
            # dW_hh / db_hh on tcgen05 straight from dgi, dgh_n and the (shifted) h_seq: no h_prev copy, no cuBLAS, no column sums
            n_ctas = lib.mm_gru_wgrad_ctas(T * R)
            acc = torch.empty((n_ctas, 128, 3 * Hh), **z_)
            targets = grad_targets(ctx.param_objs, ctx.needs_input_grad[2:4])
            dw_hh, db_hh = targets if targets is not None else (torch.empty((3 * Hh, Hh), **z_), torch.empty(3 * Hh, **z_))
            _lib.check(lib.mm_gru_wgrad(_lib.ptr(dgi), _lib.ptr(dghn), _lib.ptr(h_seq), _lib.ptr(h0 if has_h0 else None), _lib.ptr(scale),
                                        _lib.ptr(acc), n_ctas, _lib.ptr(dw_hh), _lib.ptr(db_hh), int(targets is not None), T, R, Hh,
                                        _lib.stream()))
            if targets is not None:
                dw_hh = db_hh = None                     # already added to the parameters' .grad
            return dgi, (dh0 if (has_h0 and ctx.needs_input_grad[1]) else None), dw_hh, db_hh, None, None
        h_prev = torch.empty((T, R, Hh), **z_)
        if has_h0:
            h_prev[0].copy_(h0)
        else:
            h_prev[0].zero_()
        h_prev[1:].copy_(h_seq[:-1])
        hp2 = h_prev.view(T * R, Hh)
        dghn2 = dghn.view(T * R, Hh)
        dw_hh_all = gram(dgi, hp2)                       # the n rows of this product are discarded: one pass over hp2
        dw_hh = torch.cat((dw_hh_all[:2 * Hh], gram(dghn2, hp2)), dim=0)
        db_hh = torch.cat((colsum(dgi)[:2 * Hh], colsum(dghn2)))          # b_hh's r and z gradients are b_ih's
        return dgi, (dh0 if (has_h0 and ctx.needs_input_grad[1]) else None), dw_hh, db_hh, None, None


def gru_from_gi(gi, h0, w_hh, b_hh, T, R):
    """gi [T * R, 3H] (input projections incl. b_ih) -> h_seq [T, R, H]."""
    return _GruFromGiFn.apply(gi, h0, w_hh, b_hh, T, R)


class _LnHeadFn(torch.autograd.Function):
    """``Linear(H, NA)(LayerNorm(H)(h))`` on the GRU outputs (mm_ln_head_fwd / _bwd): the NA = 5 head is a bandwidth problem, one
    pass over the rows per direction instead of LayerNorm kernels + three padded cuBLAS GEMMs + a bias column sum."""

    @staticmethod
    def forward(ctx, h, ln_w, ln_b, wq, bq):
        M, Hh = h.shape
        NA = wq.shape[0]
        h = h.contiguous()
        ps = [t.detach().contiguous() for t in (ln_w, ln_b, wq, bq)]
        z = dict(dtype=torch.float32, device=h.device)
        q, mean, rstd = torch.empty((M, NA), **z), torch.empty(M, **z), torch.empty(M, **z)
        _lib.check(_lib.lib().mm_ln_head_fwd(_lib.ptr(h), *[_lib.ptr(t) for t in ps], _lib.ptr(q), _lib.ptr(mean), _lib.ptr(rstd),
                                             M, Hh, NA, _lib.stream()))
        ctx.save_for_backward(h, mean, rstd, *ps)
        ctx.param_objs = [ln_w, ln_b, wq, bq]
        return q

    @staticmethod
    def backward(ctx, dq):
        h, mean, rstd, ln_w, ln_b, wq, bq = ctx.saved_tensors
        M, Hh = h.shape
        NA = wq.shape[0]
        z = dict(dtype=torch.float32, device=h.device)
        dh = torch.empty_like(h)
        targets = grad_targets(ctx.param_objs, ctx.needs_input_grad[1:])
        if targets is not None:                          # the kernel accumulates: straight into the parameters' .grad
            g_lnw, g_lnb, g_wq, g_bq = targets
            _lib.check(_lib.lib().mm_ln_head_bwd(_lib.ptr(dq.contiguous()), _lib.ptr(h), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(ln_w),
                                                 _lib.ptr(ln_b), _lib.ptr(wq), _lib.ptr(dh), _lib.ptr(g_wq), _lib.ptr(g_bq),
                                                 _lib.ptr(g_lnw), _lib.ptr(g_lnb), M, Hh, NA, _lib.stream()))
            return dh, None, None, None, None
        flat = torch.zeros(NA * Hh + NA + 2 * Hh, **z)                    # dwq | dbq | dln_w | dln_b, accumulated by the kernel
        dwq, dbq, dg, db = flat[:NA * Hh].view(NA, Hh), flat[NA * Hh:NA * Hh + NA], flat[NA * Hh + NA:NA * Hh + NA + Hh], flat[NA * Hh + NA + Hh:]
        _lib.check(_lib.lib().mm_ln_head_bwd(_lib.ptr(dq.contiguous()), _lib.ptr(h), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(ln_w),
                                             _lib.ptr(ln_b), _lib.ptr(wq), _lib.ptr(dh), _lib.ptr(flat), _lib.ptr(flat[NA * Hh:]),
                                             _lib.ptr(flat[NA * Hh + NA:]), _lib.ptr(flat[NA * Hh + NA + Hh:]), M, Hh, NA, _lib.stream()))
        return dh, dg, db, dwq, dbq


def ln_head_fits(hidden, n_out):
    return hidden in (32, 64) and 1 <= n_out <= 8 and (hidden * n_out) % 4 == 0


def ln_head(h, ln_w, ln_b, wq, bq):
    """h [..., H] -> Linear(H, NA)(LayerNorm(H)(h)) [..., NA]."""
    out = _LnHeadFn.apply(h.reshape(-1, h.shape[-1]), ln_w, ln_b, wq, bq)
    return out.view(*h.shape[:-1], wq.shape[0])
