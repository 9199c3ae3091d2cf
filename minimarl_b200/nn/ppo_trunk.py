"""Fused recurrent trunk for the PPO update: ``LayerNorm -> [Linear ReLU LayerNorm] x2 -> GRU -> LayerNorm -> Linear``
over data chunks as ONE differentiable op (``mm_ppo_trunk_fwd`` / ``mm_ppo_trunk_bwd``, tcgen05).

It computes what ``R_Actor.evaluate_actions`` / ``R_Critic.forward`` compute in the chunked (recurrent-generator)
mode (mappo/algorithms/r_actor_critic.py:95-146,189-208; rnn.py:24-80): rows are data chunks, the carried state is
multiplied by ``masks[t]`` before every step.  The forward stores the operands of the weight gradients, the backward
kernel re-computes the forward per step, walks the chunk in reverse and returns the pre-activation gradients; the
weight gradients are four large-K TF32 GEMMs of those, every bias / LayerNorm / head-weight gradient is reduced
inside the kernel.  ``LayerNorm(D)``'s affine is folded into the first Linear for the kernel and unfolded here.
"""
import ctypes as C

import torch

from .. import _lib

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


class _PPOTrunkFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, obs, h0, masks, *params):
        p = dict(zip(PARAM_ORDER, params))
        L, R, D = obs.shape
        NO = p["wo"].shape[0]
        with torch.no_grad():
            packed = _pack(p, D, NO)
        z = dict(dtype=torch.float32, device=obs.device)
        out = torch.empty((L, R, NO), **z)
        train = any(ctx.needs_input_grad[3:])
        if train:
            x0, x1 = torch.empty((L, R, K1), **z), torch.empty((L, R, H), **z)
            x2, hp = torch.empty((L, R, H), **z), torch.empty((L, R, H), **z)
        else:
            x0 = x1 = x2 = hp = None
        _lib.check(_lib.lib().mm_ppo_trunk_fwd(_lib.ptr(packed), _lib.ptr(obs), _lib.ptr(h0), _lib.ptr(masks), _lib.ptr(out),
                                               None, _lib.ptr(x0), _lib.ptr(x1), _lib.ptr(x2), _lib.ptr(hp), L, R, D, NO,
                                               _lib.stream()))
        if train:
            ctx.save_for_backward(packed, obs, masks, x0, x1, x2, hp, p["w1"], p["ln0_w"], p["ln0_b"])
            ctx.dims = (L, R, D, NO)
        return out

    @staticmethod
    def backward(ctx, dout):
        packed, obs, masks, x0, x1, x2, hp, w1, ln0_w, ln0_b = ctx.saved_tensors
        L, R, D, NO = ctx.dims
        o = offsets()
        n_small = o["total"] - o["small"]
        z = dict(dtype=torch.float32, device=obs.device)
        gA1, gA2 = torch.empty((L * R, H), **z), torch.empty((L * R, H), **z)
        gGi, gGhn = torch.empty((L * R, 3 * H), **z), torch.empty((L * R, H), **z)
        n_part = _lib.lib().mm_ppo_trunk_max_ctas()
        partial = torch.zeros((n_part, n_small), **z)
        _lib.check(_lib.lib().mm_ppo_trunk_bwd(_lib.ptr(packed), _lib.ptr(obs), _lib.ptr(masks), _lib.ptr(hp),
                                               _lib.ptr(dout.contiguous()), _lib.ptr(gA1), _lib.ptr(gA2), _lib.ptr(gGi),
                                               _lib.ptr(gGhn), _lib.ptr(partial), n_part, L, R, D, NO, _lib.stream()))
        small = partial.sum(0)
        s = lambda k, n: small[o[k] - o["small"]: o[k] - o["small"] + n]
        gw1f = gA1.t() @ x0.view(L * R, K1)                       # gradient of the folded first Linear [H, K1]
        gb1f = s("b1", H)
        gw2 = gA2.t() @ x1.view(L * R, H)
        gwih = gGi.t() @ x2.view(L * R, H)
        hp2 = hp.view(L * R, H)
        gwhh = torch.cat([gGi[:, :2 * H].t() @ hp2, gGhn.t() @ hp2], 0)
        gbih = s("bih", 3 * H)
        gbhh = torch.cat([gbih[:2 * H], s("bhh", 3 * H)[2 * H:]])
        # unfold LayerNorm(D): w1f = w1 * g0, b1f = b1 + w1 @ b0  =>  dw1 = dw1f * g0 + outer(db1f, b0)
        gw1f = gw1f[:, :D]
        grads = dict(ln0_w=(gw1f * w1).sum(0), ln0_b=w1.t() @ gb1f, w1=gw1f * ln0_w + torch.outer(gb1f, ln0_b), b1=gb1f,
                     ln1_w=s("ln1_w", H), ln1_b=s("ln1_b", H), w2=gw2, b2=s("b2", H), ln2_w=s("ln2_w", H),
                     ln2_b=s("ln2_b", H), wih=gwih, whh=gwhh, bih=gbih, bhh=gbhh, ln3_w=s("ln3_w", H),
                     ln3_b=s("ln3_b", H), wo=s("wo", NOMAX * H).view(NOMAX, H)[:NO], bo=s("bo", NOMAX)[:NO])
        return (None, None, None) + tuple(grads[k] if ctx.needs_input_grad[3 + i] else None
                                          for i, k in enumerate(PARAM_ORDER))


def ppo_trunk(obs, h0, masks, params):
    """obs [L, R, D], h0 [R, 32], masks [L, R] or None, params: dict keyed by PARAM_ORDER -> head output [L, R, NO]."""
    obs, h0 = obs.contiguous(), h0.contiguous()
    masks = masks.contiguous() if masks is not None else None
    return _PPOTrunkFn.apply(obs, h0, masks, *[params[k] for k in PARAM_ORDER])
