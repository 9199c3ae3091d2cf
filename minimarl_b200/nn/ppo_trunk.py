"""Fused recurrent trunk for the PPO update: ``LayerNorm -> [Linear ReLU LayerNorm] x2 -> GRU -> LayerNorm -> Linear``
over data chunks as ONE differentiable op (``mm_ppo_trunk_fwd`` / ``mm_ppo_trunk_bwd``, tcgen05).

It computes what ``R_Actor.evaluate_actions`` / ``R_Critic.forward`` compute in the chunked (recurrent-generator)
mode (mappo/algorithms/r_actor_critic.py:95-146,189-208; rnn.py:24-80): rows are data chunks, the carried state is
multiplied by ``masks[t]`` before every step.  The forward stores the operands of the weight gradients (normalised
activations, LayerNorm scales, ReLU masks), the backward kernel re-computes only the GRU cell per step, walks the chunk in
reverse and returns the pre-activation gradients G; all
weight and bias gradients are blocks of ONE large-K TF32 GEMM ``G^T X`` (X carries a ones column), the LayerNorm gain
and head gradients come from column sums reduced inside the kernel.  ``LayerNorm(D)``'s affine is folded into the first Linear for the kernel and unfolded here.
"""
import ctypes as C
import os

import torch

from .. import _lib
from ..ops.gram import gram
from ._direct import direct_param_grads, grad_targets  # noqa: F401  (re-exported)

_OFF = None
_NAMES = "w1 w2 wih whh small b1 ln1_w ln1_b b2 ln2_w ln2_b bih bhh ln3_w ln3_b wo bo total".split()
H, K1, NOMAX = 32, 48, 8


def offsets():
    global _OFF
    if _OFF is None:
        o = (C.c_int32 * 18)()
        _lib.check(_lib.lib().mm_ppo_trunk_param_offsets(o))
        _OFF = dict(zip(_NAMES, list(o)))
    return _OFF


def trunk_fits(in_dim, hidden, n_out):
    return hidden == H and 1 <= in_dim <= K1 and 1 <= n_out <= NOMAX


# order of the parameter tensors taken by ppo_trunk()
PARAM_ORDER = ("ln0_w", "ln0_b", "w1", "b1", "ln1_w", "ln1_b", "w2", "b2", "ln2_w", "ln2_b", "wih", "whh", "bih", "bhh",
               "ln3_w", "ln3_b", "wo", "bo")


def _pack(p, D, NO):
    o = offsets()
    dev = p["w1"].device
    buf = torch.zeros(o["total"], dtype=torch.float32, device=dev)
    buf[o["w1"]: o["w1"] + H * K1].view(H, K1)[:, :D] = p["w1"] * p["ln0_w"]
    buf[o["b1"]: o["b1"] + H] = p["b1"] + p["w1"] @ p["ln0_b"]
    for k, n in (("w2", H * H), ("wih", 3 * H * H), ("whh", 3 * H * H), ("ln1_w", H), ("ln1_b", H), ("b2", H), ("ln2_w", H),
                 ("ln2_b", H), ("bih", 3 * H), ("bhh", 3 * H), ("ln3_w", H), ("ln3_b", H)):
        buf[o[k]: o[k] + n] = p[k].reshape(-1)
    buf[o["wo"]: o["wo"] + NO * H] = p["wo"].reshape(-1)
    buf[o["bo"]: o["bo"] + NO] = p["bo"]
    return buf


_DIMS = None


def dims():
    global _DIMS
    if _DIMS is None:
        d = (C.c_int32 * 8)()
        _lib.check(_lib.lib().mm_ppo_trunk_dims(d))
        _DIMS = dict(zip("XS GS n_acc S dbo g2 g1 NOMAX".split(), list(d)))
    return _DIMS


class _PPOTrunkFnV1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, obs, h0, masks, rows, *params):
        p = dict(zip(PARAM_ORDER, params))
        if rows is None:
            L, R, D = obs.shape
            row_index, step_rows = None, 0
        else:                                   # obs / h0 / masks are row tables, rows = (first row of every chunk, rows per step, L)
            row_index, step_rows, L = rows
            R, D = row_index.numel(), obs.shape[-1]
        NO = p["wo"].shape[0]
        with torch.no_grad():
            packed = _pack(p, D, NO)
        z = dict(dtype=torch.float32, device=obs.device)
        out = torch.empty((L, R, NO), **z)
        train = any(ctx.needs_input_grad[4:])
        X = None
        if train:
            XS = dims()["XS"]
            # columns 144..151 are written by the kernel: the ones column (bias gradients come out of the weight-gradient
            # GEMM), 1 / sigma of the two LayerNorms and the ReLU masks for the backward kernel
            X = torch.empty((L * R, XS), **z)
        _lib.check(_lib.lib().mm_ppo_trunk_fwd(_lib.ptr(packed), _lib.ptr(obs), _lib.ptr(row_index), int(step_rows), _lib.ptr(h0),
                                               _lib.ptr(masks), _lib.ptr(out), None, _lib.ptr(X), L, R, D, NO, _lib.stream()))
        if train:
            ctx.save_for_backward(packed, masks, X, row_index if row_index is not None else packed.new_empty(0), *params)
            ctx.dims = (L, R, D, NO, int(step_rows), row_index is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        packed, masks, X, row_index = ctx.saved_tensors[:4]
        p = dict(zip(PARAM_ORDER, ctx.saved_tensors[4:]))
        L, R, D, NO, step_rows, lazy = ctx.dims
        d = dims()
        z = dict(dtype=torch.float32, device=X.device)
        G = torch.empty((L * R, d["GS"]), **z)
        n_part = _lib.lib().mm_ppo_trunk_max_ctas()
        partial = torch.zeros((n_part, d["n_acc"]), **z)
        _lib.check(_lib.lib().mm_ppo_trunk_bwd(_lib.ptr(packed), _lib.ptr(row_index if lazy else None), step_rows,
                                               _lib.ptr(masks), _lib.ptr(X),
                                               _lib.ptr(dout.contiguous()), _lib.ptr(G), _lib.ptr(partial), n_part, L, R, D, NO,
                                               _lib.stream()))
        acc = partial.sum(0)
        W = gram(G, X)                                           # [GS, XS]: every weight / bias gradient is a block of it
        S = acc[d["S"]: d["S"] + NOMAX * H].view(NOMAX, H)[:NO]   # sum_r dout[r, o] * xhat3[r, j]
        dbo = acc[d["dbo"]: d["dbo"] + NO]
        gw1f, gb1f = W[0:32, 0:D], W[0:32, 144]                   # folded first Linear
        # X holds the NORMALISED activations xh1 / xh2 (x = xh * gain + offset is what the layer consumed):
        #   dW = G^T x = (G^T xh) * gain + outer(db, offset)
        gb2, gbih = W[32:64, 144], W[64:160, 144]
        gw2 = W[32:64, 48:80] * p["ln1_w"] + torch.outer(gb2, p["ln1_b"])
        gwih = W[64:160, 80:112] * p["ln2_w"] + torch.outer(gbih, p["ln2_b"])
        gwhh = torch.cat([W[64:128, 112:144], W[160:192, 112:144]], 0)
        gbhh = torch.cat([gbih[:2 * H], W[160:192, 144]])
        w1, wo = p["w1"], p["wo"]
        grads = dict(
            # unfold LayerNorm(D): w1f = w1 * g0, b1f = b1 + w1 @ b0  =>  dw1 = dw1f * g0 + outer(db1f, b0)
            ln0_w=(gw1f * w1).sum(0), ln0_b=w1.t() @ gb1f, w1=gw1f * p["ln0_w"] + torch.outer(gb1f, p["ln0_b"]), b1=gb1f,
            ln1_w=acc[d["g1"]: d["g1"] + H], ln1_b=gb2 @ p["w2"], w2=gw2, b2=gb2,
            ln2_w=acc[d["g2"]: d["g2"] + H], ln2_b=gbih @ p["wih"], wih=gwih, whh=gwhh, bih=gbih, bhh=gbhh,
            ln3_w=(wo * S).sum(0), ln3_b=dbo @ wo, wo=S * p["ln3_w"] + torch.outer(dbo, p["ln3_b"]), bo=dbo)
        return (None, None, None, None) + tuple(grads[k] if ctx.needs_input_grad[4 + i] else None
                                          for i, k in enumerate(PARAM_ORDER))


_DIMS2 = None


def dims2():
    global _DIMS2
    if _DIMS2 is None:
        d = (C.c_int32 * 8)()
        _lib.check(_lib.lib().mm_ppo_trunk2_dims(d))
        _DIMS2 = dict(zip("n_params saved_bytes acc_cols acc_a_cols tile_rows K1 NOMAX H".split(), list(d)))
    return _DIMS2


def obs_table(obs2d, masks=None, out=None):
    """LayerNorm (no affine) of every observation row, fp16, zero padded to 48 columns: [rows, D] -> [rows, 48].  Computed once
    per update; both networks and every epoch gather their chunk rows from it (mm_ppo_obs_table).  With D < 48 the row's mask
    rides in column 47 (``masks`` [rows]): the kernels then read it from the sector their gather fetches anyway."""
    obs2d = obs2d.contiguous()
    rows, D = obs2d.shape
    if masks is not None and D >= K1:
        masks = None
    if out is None:
        out = torch.empty((rows, K1), dtype=torch.float16, device=obs2d.device)
    assert out.shape == (rows, K1) and out.dtype == torch.float16 and out.is_contiguous()
    _lib.check(_lib.lib().mm_ppo_obs_table(_lib.ptr(obs2d), _lib.ptr(masks), _lib.ptr(out), rows, D, _lib.stream()))
    return out


def _ptr_array(params):
    """HOST array of the 18 device pointers (PARAM_ORDER) the pack / unfold launches take."""
    arr = (C.c_void_p * len(PARAM_ORDER))()
    for i, t in enumerate(params):
        assert t.is_contiguous() and t.dtype == torch.float32
        arr[i] = t.data_ptr()
    return arr


def _pack2(params, D, NO):
    """Kernel parameters with every LayerNorm affine folded into the Linear that follows it (W' = W diag(g), b' = b + W beta):
    one launch (mm_ppo_trunk2_pack)."""
    out = torch.empty(dims2()["n_params"], dtype=torch.float32, device=params[0].device)
    _lib.check(_lib.lib().mm_ppo_trunk2_pack(_ptr_array(params), _lib.ptr(out), D, NO, _lib.stream()))
    return out


_GRAD_LAYOUT = {}


def _grad_layout(D, NO):
    if (D, NO) not in _GRAD_LAYOUT:
        o = (C.c_int32 * (len(PARAM_ORDER) + 1))()
        _lib.check(_lib.lib().mm_ppo_trunk2_grad_layout(D, NO, o))
        _GRAD_LAYOUT[(D, NO)] = list(o)
    return _GRAD_LAYOUT[(D, NO)]


class _PPOTrunkFn(torch.autograd.Function):
    """mm_ppo_trunk2_fwd / mm_ppo_trunk2_bwd: fp16 operands on tcgen05, thread = row, weight gradients inside the backward."""

    @staticmethod
    def forward(ctx, xhat, h0, masks, first, step_rows, L, D, *params):
        # masks None and D < 48: the masks ride in column 47 of xhat
        p = dict(zip(PARAM_ORDER, params))
        R, NO = first.numel(), p["wo"].shape[0]
        d = dims2()
        params = [t.detach().contiguous() for t in params]
        packed = _pack2(params, D, NO)
        out = torch.empty((L, R, NO), dtype=torch.float32, device=xhat.device)
        train = any(ctx.needs_input_grad[7:])
        saved = None
        if train:
            n_tiles = (R + d["tile_rows"] - 1) // d["tile_rows"]
            saved = torch.empty(n_tiles * L * d["saved_bytes"], dtype=torch.uint8, device=xhat.device)
        in_table = int(masks is None and D < K1)
        _lib.check(_lib.lib().mm_ppo_trunk2_fwd(_lib.ptr(packed), _lib.ptr(xhat), _lib.ptr(first), int(step_rows), _lib.ptr(h0),
                                                _lib.ptr(masks), in_table, _lib.ptr(out), None, _lib.ptr(saved), L, R, NO,
                                                _lib.stream()))
        if train:
            ctx.save_for_backward(packed, xhat, first, masks if masks is not None else packed.new_empty(0), saved, *params)
            ctx.dims = (L, R, D, NO, int(step_rows), masks is not None)
            ctx.param_objs = list(p.values())            # the Parameters themselves (their .grad may take the gradients directly)
        return out

    @staticmethod
    def backward(ctx, dout):
        packed, xhat, first, masks, saved = ctx.saved_tensors[:5]
        p = dict(zip(PARAM_ORDER, ctx.saved_tensors[5:]))
        L, R, D, NO, step_rows, has_masks = ctx.dims
        d = dims2()
        dout = dout.contiguous()
        # gradients travel as fp16 operands: a power-of-two scale puts max |dout| at 64 (fp16 overflows at 65 504, keeps 11 bits
        # down to 6e-5); the weight gradients come back multiplied by it
        scale = torch.empty(2, dtype=torch.float32, device=dout.device)
        _lib.check(_lib.lib().mm_ppo_trunk2_scale(_lib.ptr(dout), dout.numel(), _lib.ptr(scale), _lib.stream()))
        n_ctas = _lib.lib().mm_ppo_trunk2_bwd_ctas(R)
        acc = torch.empty((n_ctas, d["tile_rows"], d["acc_cols"]), dtype=torch.float32, device=dout.device)
        _lib.check(_lib.lib().mm_ppo_trunk2_bwd(_lib.ptr(packed), _lib.ptr(xhat), _lib.ptr(first), step_rows,
                                                _lib.ptr(masks if has_masks else None), int(not has_masks and D < K1),
                                                _lib.ptr(saved), _lib.ptr(dout),
                                                _lib.ptr(scale), _lib.ptr(acc), n_ctas, L, R, NO, _lib.stream()))
        # one launch: sum over the CTAs, 1 / scale, LayerNorm affines unfolded into the gradients of the module's own tensors
        off = _grad_layout(D, NO)
        red = torch.empty(d["tile_rows"] * d["acc_cols"], dtype=torch.float32, device=dout.device)
        params = [p[k].detach().contiguous() for k in PARAM_ORDER]
        targets = grad_targets(ctx.param_objs, ctx.needs_input_grad[7:])
        if targets is not None:
            gptrs = (C.c_void_p * len(PARAM_ORDER))(*[g.data_ptr() for g in targets])
            _lib.check(_lib.lib().mm_ppo_trunk2_unfold_into(_lib.ptr(acc), n_ctas, _lib.ptr(scale), _lib.ptr(red), _ptr_array(params),
                                                            gptrs, D, NO, _lib.stream()))
            return (None,) * (7 + len(PARAM_ORDER))
        flat = torch.empty(off[-1], dtype=torch.float32, device=dout.device)
        _lib.check(_lib.lib().mm_ppo_trunk2_unfold(_lib.ptr(acc), n_ctas, _lib.ptr(scale), _lib.ptr(red),
                                                   _ptr_array(params), _lib.ptr(flat), D, NO, _lib.stream()))
        grads = {k: flat[off[i]: off[i + 1]].view_as(p[k]) for i, k in enumerate(PARAM_ORDER)}
        return (None,) * 7 + tuple(grads[k] if ctx.needs_input_grad[7 + i] else None for i, k in enumerate(PARAM_ORDER))


class ChunkRows:
    """Data chunks addressed in place: ``table`` [rows, D] is the rollout buffer's storage seen as rows, step ``l`` of
    chunk ``c`` is row ``first[c] + l * step_rows``.  What recurrent_generator would otherwise gather into [L * Nc, D].
    ``cache`` (a dict shared by the minibatches of one update) keeps the normalised fp16 observation table of this storage."""

    def __init__(self, table, first, step_rows, L, cache=None, key=None):
        self.table, self.first, self.step_rows, self.L = table, first, int(step_rows), int(L)
        self.cache, self.key = cache, key

    def xhat(self, masks=None):
        """The normalised fp16 table of this storage, with ``masks`` [rows] embedded (kept in ``cache`` for the whole update)."""
        if self.cache is None:
            return obs_table(self.table, masks)
        t = self.cache.get(self.key)
        if t is None:
            t = self.cache[self.key] = obs_table(self.table, masks)
        return t

    def size(self, dim):
        return (self.L * self.first.numel(), self.table.shape[-1])[dim]

    def gather(self):
        idx = self.first.unsqueeze(0) + torch.arange(self.L, device=self.first.device).unsqueeze(1) * self.step_rows
        return self.table[idx.reshape(-1)]


def ppo_trunk_v1(obs, h0, masks, params):
    """The first-generation op (tf32 tiles, 8-column thread slices, G^T X on cuBLAS); kept for A/B measurements."""
    if isinstance(obs, ChunkRows):
        return _PPOTrunkFnV1.apply(obs.table, h0.table, masks.table.reshape(-1), (obs.first.contiguous(), obs.step_rows, obs.L),
                                   *[params[k] for k in PARAM_ORDER])
    h0 = h0.contiguous()
    masks = masks.contiguous() if masks is not None else None
    return _PPOTrunkFnV1.apply(obs.contiguous(), h0, masks, None, *[params[k] for k in PARAM_ORDER])


def ppo_trunk(obs, h0, masks, params):
    """obs [L, R, D], h0 [R, 32], masks [L, R] or None -- or all three as ChunkRows over the rollout buffer's storage (h0 with
    L = 1) -- params: dict keyed by PARAM_ORDER -> head output [L, R, NO]."""
    if os.environ.get("MM_PPO_TRUNK_V1"):
        return ppo_trunk_v1(obs, h0, masks, params)
    ps = [params[k] for k in PARAM_ORDER]
    if isinstance(obs, ChunkRows):
        if not isinstance(h0, ChunkRows):        # only the observations are addressed in place: gather would defeat it
            raise TypeError("with ChunkRows observations, h0 and masks must be ChunkRows of the same row map")
        assert h0.first is obs.first and masks.first is obs.first and masks.step_rows == obs.step_rows
        D = obs.table.shape[-1]
        mt = masks.table.reshape(-1)
        return _PPOTrunkFn.apply(obs.xhat(mt), h0.table.contiguous(), mt if D >= K1 else None, obs.first.contiguous(),
                                 obs.step_rows, obs.L, D, *ps)
    L, R, D = obs.shape
    first = torch.arange(R, device=obs.device, dtype=torch.int64)
    if masks is None:
        masks = torch.ones((L, R), dtype=torch.float32, device=obs.device)
    masks = masks.contiguous().reshape(-1)
    return _PPOTrunkFn.apply(obs_table(obs.reshape(L * R, D), masks), h0.contiguous(), masks if D >= K1 else None, first, R, L, D,
                             *ps)
